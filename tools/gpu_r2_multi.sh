#!/bin/bash
# Round-2 GPU-box session on N >= 2 GPUs: the multi-GPU tests (CLI over all GPUs, sharded channels,
# -Q segments), the bench line at N ranks, the CLI's throughput over 1..N GPUs.
cd "$(dirname "$0")/.." || exit 1
O=gpurun_out; mkdir -p $O
N=$(nvidia-smi -L | wc -l)
timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_quality_gpu.py -q -m gpu > $O/r02_pytest_multi_gpu.log 2>&1; echo "pytest rc=$? on $N GPUs" >> $O/r02_pytest_multi_gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > $O/r02_bench_n$N.json 2> $O/r02_bench_n$N.err; echo "bench rc=$?" >> $O/r02_bench_n$N.err
python - <<'PY' > $O/r02_cli_multi_n$N.txt 2>&1
import os, subprocess, sys, time
sys.path.insert(0, ".")
import numpy as np
from demodulator_b200 import synth
ndev = int(subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout.count("GPU "))
exe = "demodulator_b200/bin/demodulator"
src = "/dev/shm/fmdin.u8"
base = synth.config_signal("cfg3", 64)
with open(src, "wb") as f:
    for _ in range(512):          # 8 GiB
        f.write(base.tobytes())
def run(desc, cmd, env):
    t0 = time.time()
    r = subprocess.run(cmd, shell=True, env=dict(os.environ, FMD_STATS="1", **env), capture_output=True, text=True)
    print(f"{desc}: rc={r.returncode} wall={time.time()-t0:.2f}s :: {r.stderr.strip().splitlines()[-1] if r.stderr.strip() else ''}", flush=True)
counts = [c for c in (1, 2, 4, 8) if c <= ndev]
for c in counts:
    devs = ",".join(map(str, range(c)))
    ends = c in (counts[0], counts[-1])
    run(f"cfg3 file->/dev/null      GPUs={c}", f"{exe} -i {src} -o /dev/null -S48000 -l15000", dict(FMD_DEVICES=devs))
    run(f"-Q cfg3 file->/dev/null   GPUs={c}", f"{exe} -i {src} -o /dev/null -Q -r2400000 -S48000 -l15000", dict(FMD_DEVICES=devs))
    if not ends:
        continue
    run(f"cfg3 file->file (mmap)    GPUs={c}", f"{exe} -i {src} -o /dev/shm/fmdout.f32 -S48000 -l15000", dict(FMD_DEVICES=devs))
    run(f"cfg3 file->file (pwrite)  GPUs={c}", f"{exe} -i {src} -o /dev/shm/fmdout.f32 -S48000 -l15000", dict(FMD_DEVICES=devs, FMD_NO_MMAP="1"))
    run(f"cfg3 stdin pipe->/dev/null GPUs={c}", f"cat {src} | {exe} -i - -o /dev/null -S48000 -l15000", dict(FMD_DEVICES=devs))
    run(f"-Q cfg2 file->/dev/null   GPUs={c}", f"{exe} -i {src} -o /dev/null -Q -r192000 -S96000 -l6500 -L12500", dict(FMD_DEVICES=devs))
for th in (4, 16):
    run(f"cfg3 file->/dev/null GPUs={ndev} FMD_IO_THREADS={th}", f"{exe} -i {src} -o /dev/null -S48000 -l15000", dict(FMD_DEVICES="all", FMD_IO_THREADS=str(th)))
for p in (src, "/dev/shm/fmdout.f32"):
    if os.path.exists(p): os.unlink(p)
PY
# BASELINE configs[4] at full size: 16 G samples through stdin -> stdout, every output period checked
{ timeout 600 python tools/cfg5_stream.py --gsamples ${CFG5_GS:-16} --devices all
  timeout 600 python tools/cfg5_stream.py --gsamples ${CFG5_GS:-16} --devices 0
  timeout 600 python tools/cfg5_stream.py --gsamples 4 --devices all --quality; } > $O/r02_cfg5_stream_n$N.txt 2>&1
cat $O/r02_cfg5_stream_n$N.txt
tail -5 $O/r02_pytest_multi_gpu.log; tail -2 $O/r02_bench_n$N.err; cut -c1-600 $O/r02_bench_n$N.json; cat $O/r02_cli_multi_n$N.txt
