#!/bin/bash
# Round-2 GPU-box session (run with gpurun --gpus 8): where is the end-to-end wall of the box?
# Topology, host memory bandwidth, aggregate H2D/D2H ceilings over 1..8 GPUs for several staging
# strategies, and the CLI over FMD_DEVICES=1,2,4,8 GPUs.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
{
  echo "== nproc $(nproc)"; lscpu | grep -E 'Model name|Socket|NUMA|Thread|Core|^CPU\(s\)|L3' ; free -g | head -2
  echo "== numa nodes"; ls /sys/devices/system/node/ 2>/dev/null | tr '\n' ' '; echo
  for d in /sys/bus/pci/devices/*; do
    if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-4)" = "0x03" ]; then
      echo "$(basename $d) numa=$(cat $d/numa_node) cpus=$(cat $d/local_cpulist) link=$(cat $d/current_link_speed 2>/dev/null) x$(cat $d/current_link_width 2>/dev/null)"; fi; done
  echo "== thp: $(cat /sys/kernel/mm/transparent_hugepage/enabled 2>/dev/null)  hugepages: $(grep -E 'HugePages_Total|Hugepagesize' /proc/meminfo | tr '\n' ' ')"
  echo "== /dev/shm"; df -h /dev/shm | tail -1
  nvidia-smi topo -m 2>&1 | head -40
  echo "== iommu groups: $(ls /sys/kernel/iommu_groups 2>/dev/null | wc -l)   hypervisor: $(grep -c hypervisor /proc/cpuinfo)"
} > $O/r02_topology.txt 2>&1
./tools/hostbw 1,2,4,8,16,32 > $O/r02_hostbw.txt 2>&1
{
  for w in threads procs; do ./tools/pciebench_multi --workers $w --alloc pinned; done
  ./tools/pciebench_multi --workers threads --alloc wc
  ./tools/pciebench_multi --workers threads --alloc thp
  ./tools/pciebench_multi --workers threads --alloc hugetlb --counts 8
  ./tools/pciebench_multi --workers threads --alloc pinned --slice 128 --counts 1,8
  ./tools/pciebench_multi --workers threads --alloc pinned --slice 4 --counts 1,8
} > $O/r02_pcie_multi.txt 2>&1
# ---- the CLI over 1, 2, 4, 8 GPUs: 8 GiB file in /dev/shm -> file in /dev/shm, order checked by cmp
python - <<'PY' > $O/r02_cli_multi.txt 2>&1
import os, subprocess, sys, time
sys.path.insert(0, ".")
import numpy as np
from demodulator_b200 import synth
base = synth.config_signal("cfg2", 64)            # 16 MiB of cfg2 signal
src, n_rep = "/dev/shm/fmd_in.u8", 512            # 8 GiB
with open(src, "wb") as f:
    for _ in range(n_rep):
        f.write(base.tobytes())
exe = "demodulator_b200/bin/demodulator"
ndev = int(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU "))
ref_out = None
for devs in [1, 2, 4, 8]:
    if devs > ndev: break
    for out in ("/dev/null", "/dev/shm/fmd_out_%d.f32" % devs):
        env = dict(os.environ, FMD_DEVICES=",".join(map(str, range(devs))), FMD_STATS="1")
        t0 = time.time()
        r = subprocess.run([exe, "-i", src, "-o", out, "-S96000", "-l6500", "-L12500"], env=env, capture_output=True, text=True)
        print(f"devices={devs} out={out} rc={r.returncode} wall={time.time()-t0:.2f}s :: {r.stderr.strip()}", flush=True)
    cur = "/dev/shm/fmd_out_%d.f32" % devs
    if ref_out is None: ref_out = cur
    else:
        same = subprocess.run(["cmp", ref_out, cur]).returncode == 0
        print(f"   output of {devs} GPUs identical to 1 GPU: {same}", flush=True)
        os.unlink(cur)
# the same through a stdin pipe (cfg5-style stream), 8 GPUs vs 1
for devs in [1, ndev]:
    env = dict(os.environ, FMD_DEVICES=",".join(map(str, range(devs))), FMD_STATS="1")
    t0 = time.time()
    r = subprocess.run(f"cat {src} | {exe} -i - -o /dev/null -S96000 -l6500 -L12500", shell=True, env=env, capture_output=True, text=True)
    print(f"stdin pipe devices={devs} rc={r.returncode} wall={time.time()-t0:.2f}s :: {r.stderr.strip()}", flush=True)
for p in (src, ref_out):
    if p and os.path.exists(p): os.unlink(p)
PY
echo done
