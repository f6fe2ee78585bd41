#!/usr/bin/env python
"""bench.py -- headline metric of the FM demodulation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--impl reference]

Metric (BASELINE.json): input MS/s (complex IQ samples per second), whole job over N GPUs, plus
the fraction of the HBM roofline the dominant kernel reaches (4 algorithmic bytes per input
sample: 2 B read + 4 B written per 2 samples).

A "step" = one pass of the hot path (fmd_demod_device: ONE launch of the fused unpack -> [input
lowpass] -> discriminator -> output lowpass kernel, whose leading CTAs also produce the chunk heads
op for op) over CHANNELS_PER_GPU independent channels of synthetic FM IQ that lie back to back in
HBM (chunks are independent, so a batch of channels is one call).  Default workload = BASELINE.json
configs[1] (cfg2: 192 kS/s, -L12500 -l6500 -S96000, both filter stages).  N > 1: one process per
GPU (torchrun), same per-GPU work (weak scaling), no data-path collective -- channels are
independent; NCCL is used only for the start barrier and the MAX of the timings.

`e2e` is the same metric through the host-facing C-ABI call (fmd_demod_host) with pinned HOST
buffers: H2D + kernels + D2H inside the timed region.

`--impl reference` times the reference's own CPU implementation (oracle/_ref/demodulator_ref, the
unmodified peads/demodulator CLI compiled from its sources) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import shutil
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CHUNK = 262144
CHANNELS_PER_GPU = 8            # BASELINE configs[3]: 64 channels over 8 GPUs
SAMPLES_PER_CHANNEL = 128 << 20  # 128 Mi complex samples (256 MiB in, 256 MiB out) per channel
BASE_CHUNKS = 16                # synthetic signal period per channel (tiled to full length)
E2E_SAMPLES = CHANNELS_PER_GPU * SAMPLES_PER_CHANNEL  # the same batch through the host API (2 GiB H2D + 2 GiB D2H)
ALG_BYTES_PER_SAMPLE = 4.0

FLAGS = {
    "cfg1": dict(sampleRate=96000.0, lowpassOut=12500.0),
    "cfg2": dict(sampleRate=96000.0, lowpassOut=6500.0, lowpassIn=12500.0),
    "cfg3": dict(sampleRate=48000.0, lowpassOut=15000.0),
}
CLI_FLAGS = {
    "cfg1": ["-S96000", "-l12500"],
    "cfg2": ["-S96000", "-l6500", "-L12500"],
    "cfg3": ["-S48000", "-l15000"],
}
QUALITY_CONFIGS = ["quality_cfg1", "quality_cfg2", "quality_cfg3", "quality_m1"]
WORKLOAD = {
    "cfg1": "synthetic u8 IQ FM tone 192 kS/s, -S96000 -l12500, no input lowpass",
    "cfg2": "synthetic u8 IQ FM tone 192 kS/s, -L12500 -l6500 -S96000 (input + output lowpass)",
    "cfg3": "synthetic u8 IQ FM 2.4 MS/s, -S48000 -l15000",
}


def measured_peak_gbs() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(config: str) -> float | None:
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get(config)
        if not e:
            return None
        # captured at e["samples_per_launch"]; scale linearly to this bench's launch size
        return float(e["dram_bytes_per_launch"]) * (CHANNELS_PER_GPU * SAMPLES_PER_CHANNEL / float(e["samples_per_launch"]))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int) -> None:
        self.gpu = gpu_index
        self.lines: list[str] = []
        self.proc = None

    def start(self) -> None:
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        self.proc = subprocess.Popen([exe, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu), "-lms", "100"],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self) -> None:
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        hi = [s for s, p in zip(sm, pw) if p >= 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(hi), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(pw)}


# ------------------------------------------------------------------------------------- GPU arm

def bind_to_gpu_numa_node(prop) -> str:
    """Run this rank (and therefore first-touch its pinned staging buffers) on the CPUs next to its
    GPU: the end-to-end leg is host-memory / PCIe bound, and a remote socket halves it."""
    if os.environ.get("FMD_BENCH_NO_AFFINITY"):
        return "off"
    try:
        bdf = f"{getattr(prop, 'pci_domain_id', 0):04x}:{prop.pci_bus_id:02x}:{prop.pci_device_id:02x}.0"
        base = f"/sys/bus/pci/devices/{bdf}"
        node = open(f"{base}/numa_node").read().strip()
        cpulist = open(f"{base}/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if cpus and int(node) >= 0:
            os.sched_setaffinity(0, cpus)
            return f"node {node} ({len(cpus)} cpus)"
        return f"node {node} (not bound)"
    except Exception as e:  # topology not exposed (VM): leave the scheduler alone
        return f"unavailable ({type(e).__name__})"


def gpu_arm(args) -> dict | None:
    import torch
    import torch.distributed as dist

    import demodulator_b200 as fm
    from demodulator_b200 import sharding, synth
    from demodulator_b200.api import PinnedBuffer

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (this framework has no CPU path; use --impl reference "
                         "for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(torch.cuda.get_device_properties(local))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    def barrier() -> None:
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize()

    cfgname = args.config
    sig = synth.CONFIGS[cfgname]
    n_ch = CHANNELS_PER_GPU
    my_channels = [rank * n_ch + i for i in range(n_ch)]  # global channel ids (weak scaling)
    ch_bytes = 2 * SAMPLES_PER_CHANNEL
    reps = ch_bytes // (BASE_CHUNKS * CHUNK)

    # the channels of one GPU lie back to back in one batch buffer: chunks are independent, so a
    # batch of channels (each a whole number of chunks) is ONE call / ONE launch of the hot path
    d_in_all = torch.empty(n_ch * ch_bytes, dtype=torch.uint8, device=dev)
    d_out_all = torch.empty(n_ch * ch_bytes // 4, dtype=torch.float32, device=dev)
    d_in, d_out = [], []
    for i, c in enumerate(my_channels):
        base = synth.fm_iq_u8(BASE_CHUNKS * CHUNK // 2, sig["fs_sig"], sig["df"], sig["fm"] + 50.0 * c, 100 + c)
        d_in_all[i * ch_bytes:(i + 1) * ch_bytes].view(reps, -1).copy_(torch.from_numpy(base).to(dev))
        d_in.append(d_in_all[i * ch_bytes:(i + 1) * ch_bytes])
        d_out.append(d_out_all[i * ch_bytes // 4:(i + 1) * ch_bytes // 4])

    stream = torch.cuda.Stream(device=dev)
    dem = fm.Demodulator(device=local, **FLAGS[cfgname])

    def step() -> None:
        if args.per_channel_launch:
            for i in range(n_ch):
                dem.process_device(d_in[i].data_ptr(), ch_bytes, d_out[i].data_ptr(), stream.cuda_stream)
        else:
            dem.process_device(d_in_all.data_ptr(), n_ch * ch_bytes, d_out_all.data_ptr(), stream.cuda_stream)

    with torch.cuda.stream(stream):
        for _ in range(max(args.warmup, 3)):
            step()
        stream.synchronize()

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        dem.profile_kernels(True)
        dem.kernel_time()
        l0 = dem.launches
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        stream.synchronize()
        barrier()
        dev_ms = ev0.elapsed_time(ev1)
        k_ms, k_n = dem.kernel_time()
        dem.profile_kernels(False)
        launches = dem.launches - l0

    total_ms = sharding.max_over_ranks(dev_ms, dev)
    samples_per_step_gpu = n_ch * SAMPLES_PER_CHANNEL
    value = world * samples_per_step_gpu * args.steps / (total_ms * 1e-3) / 1e6  # MS/s whole job

    # periodicity property at full size: every tile of a channel's output equals the first one
    per = BASE_CHUNKS * CHUNK // 4
    o = d_out[0].view(reps, per)
    periodic = bool(torch.equal(o[0].view(torch.int32), o[reps // 2].view(torch.int32)) and
                    torch.equal(o[0].view(torch.int32), o[reps - 1].view(torch.int32)))

    # ---- end to end through the host-facing C-ABI call, pinned host buffers ----
    e2e_bytes = 2 * E2E_SAMPLES
    pin_in, pin_out = PinnedBuffer(e2e_bytes), PinnedBuffer(e2e_bytes)
    torch.from_numpy(pin_in.u8).copy_(d_in_all[:e2e_bytes])
    dem.process_ptr(pin_in.ptr, e2e_bytes, pin_out.ptr)
    e2e_steps = max(3, min(args.steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dem.process_ptr(pin_in.ptr, e2e_bytes, pin_out.ptr)
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_s = sharding.max_over_ranks(e2e_s, dev)
    e2e_value = world * E2E_SAMPLES * e2e_steps / e2e_s / 1e6
    e2e_ok = bool(np.array_equal(pin_out.view(np.float32)[:per].view(np.uint32),
                                 d_out[0][:per].cpu().numpy().view(np.uint32)))
    clocks = sampler.stop() if rank == 0 else None

    peak, peak_src = measured_peak_gbs()
    k_avg_ms = k_ms / max(k_n, 1)
    samples_per_launch = n_ch * SAMPLES_PER_CHANNEL * args.steps / max(k_n, 1)
    achieved = ALG_BYTES_PER_SAMPLE * samples_per_launch / (k_avg_ms * 1e-3) / 1e9
    traffic = ncu_traffic(cfgname)

    cpu = cpu_baseline(cfgname, d_in[0]) if (rank == 0 and world == 1 and not args.no_cpu) else None

    pin_in.free(); pin_out.free()
    dem.close()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return None
    line = {
        "metric": "input MS/s (complex IQ samples/s), whole job",
        "value": round(value, 1),
        "unit": "MS/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": max(args.warmup, 3),
        "ms_per_step": round(total_ms / args.steps, 4),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {
            "workload": WORKLOAD[cfgname], "name": cfgname,
            "channels_per_gpu": n_ch, "samples_per_channel": SAMPLES_PER_CHANNEL,
            "samples_per_step_per_gpu": samples_per_step_gpu, "chunk_bytes": CHUNK,
            "parallelism": f"channel-parallel x{world}, no collective",
            "launches_per_step_per_gpu": n_ch if args.per_channel_launch else 1,
            "l2": "inputs larger than L2: 2 GiB in + 2 GiB out per GPU per step vs 126 MB L2 (no flush needed)",
            "e2e_samples_per_step_per_gpu": E2E_SAMPLES, "e2e_steps": e2e_steps, "cpu_affinity": numa,
        },
        "roofline": {
            "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
            "frac": round(achieved / peak, 4), "traffic": traffic,
            "kernel": "fmd_fast_kernel", "kernel_ms_avg": round(k_avg_ms, 4), "kernel_launches_timed": k_n,
            "algorithmic_bytes_per_launch": ALG_BYTES_PER_SAMPLE * samples_per_launch,
            "samples_per_launch": samples_per_launch,
            "peak_source": peak_src,
            "kernel_share_of_step": round(k_ms / max(dev_ms, 1e-9), 4),
        },
        "cpu_baseline": cpu,
        "e2e": {"value": round(e2e_value, 1), "unit": "MS/s",
                "h2d_bytes_per_step": e2e_bytes, "d2h_bytes_per_step": e2e_bytes,
                "api": "fmd_demod_host (pinned host in/out, 32 MiB slices over copy-in/compute/copy-out streams)",
                "matches_device_path": e2e_ok},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "checks": {"output_periodic_over_tiles": periodic},
    }
    return line


# --------------------------------------------------------------------------- CPU reference arm

def _ref_binary() -> str | None:
    """The unmodified reference CLI (built where /root/reference exists, travels as a prebuilt
    file); None when it is absent -- the CPU legs then time the oracle port instead."""
    from oracle import pyoracle as po
    return po.REF_BIN if os.path.exists(po.REF_BIN) and os.access(po.REF_BIN, os.X_OK) else None


def _port_worker(args) -> float:
    """One process of the fallback CPU leg: the oracle port (fm_oracle.c) on `samples` samples."""
    cfgname, samples = args
    from demodulator_b200 import synth
    from oracle import pyoracle as po
    sig = synth.CONFIGS[cfgname]
    base = synth.fm_iq_u8(BASE_CHUNKS * CHUNK // 2, sig["fs_sig"], sig["df"], sig["fm"], 100)
    raw = np.resize(base, 2 * samples)
    P = po.port()
    t0 = time.perf_counter()
    P.run(raw, po.Config(**sig["opts"]))
    return time.perf_counter() - t0


def _run_port_processes(cfgname: str, samples: int, nproc: int) -> float:
    import multiprocessing as mp
    t0 = time.perf_counter()
    with mp.get_context("spawn").Pool(nproc) as pool:
        pool.map(_port_worker, [(cfgname, samples)] * nproc)
    return time.perf_counter() - t0


def _scratch_dir() -> str:
    for d in ("/dev/shm", tempfile.gettempdir()):
        if os.path.isdir(d) and os.access(d, os.W_OK):
            return d
    return ROOT


def _run_ref_processes(binary: str, path: str, flags: list[str], nproc: int) -> float:
    t0 = time.perf_counter()
    ps = [subprocess.Popen([binary, "-i", path, "-o", "/dev/null"] + flags, stdout=subprocess.DEVNULL,
                           stderr=subprocess.DEVNULL) for _ in range(nproc)]
    for p in ps:
        p.wait()
    return time.perf_counter() - t0


def cpu_baseline(cfgname: str, d_in0=None, samples: int = 384 << 20) -> dict:
    """The reference CLI (1 reader + 1 worker thread) on a bounded sample of the same workload."""
    from demodulator_b200 import synth
    binary = _ref_binary()
    if binary is None:  # reference binary did not travel: time the plain-C port, one thread
        samples = 32 << 20
        sec = min(_port_worker((cfgname, samples)) for _ in range(2))
        return {"value": round(samples / sec / 1e6, 2), "unit": "MS/s", "cores": 1, "kind": "port",
                "sample": f"{samples >> 20} Mi samples of the same signal, oracle/fm_oracle.c in-process, best of 2",
                "host_cores_available": os.cpu_count()}
    path = os.path.join(_scratch_dir(), f"fmd_bench_{os.getpid()}.u8")
    nbytes = 2 * samples
    try:
        if d_in0 is not None and d_in0.numel() >= nbytes:
            d_in0[:nbytes].cpu().numpy().tofile(path)
        else:
            sig = synth.CONFIGS[cfgname]
            base = synth.fm_iq_u8(BASE_CHUNKS * CHUNK // 2, sig["fs_sig"], sig["df"], sig["fm"], 100)
            with open(path, "wb") as f:
                for _ in range(nbytes // base.size):
                    f.write(base.tobytes())
        _run_ref_processes(binary, path, CLI_FLAGS[cfgname], 1)  # warm the page cache / binary
        sec = min(_run_ref_processes(binary, path, CLI_FLAGS[cfgname], 1) for _ in range(2))
    finally:
        if os.path.exists(path):
            os.unlink(path)
    return {"value": round(samples / sec / 1e6, 2), "unit": "MS/s", "cores": 2, "kind": "reference",
            "sample": f"{samples >> 20} Mi samples of the same signal, file in {os.path.dirname(path)} -> /dev/null, "
                      f"unmodified reference CLI (oracle/_ref/demodulator_ref, 1 reader + 1 worker thread), best of 2",
            "host_cores_available": os.cpu_count()}


def reference_arm(args) -> dict | None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    from demodulator_b200 import synth
    cfgname = args.config
    binary = _ref_binary()
    ncpu = os.cpu_count() or 2
    if binary is None:
        return _reference_arm_port(args, cfgname, ncpu)
    nproc = max(1, ncpu // 2)                       # each reference process runs 2 threads
    steps_total = args.steps + args.warmup
    # bounded sample: aim at ~120 s for the whole run at ~30 MS/s per process
    per_proc = int(120.0 * 30e6 / max(steps_total, 1))
    per_proc = max(16 << 20, min(per_proc, 256 << 20)) // (CHUNK // 2) * (CHUNK // 2)
    path = os.path.join(_scratch_dir(), f"fmd_refarm_{os.getpid()}.u8")
    sig = synth.CONFIGS[cfgname]
    base = synth.fm_iq_u8(BASE_CHUNKS * CHUNK // 2, sig["fs_sig"], sig["df"], sig["fm"], 100)
    try:
        with open(path, "wb") as f:
            for _ in range(2 * per_proc // base.size):
                f.write(base.tobytes())
        for _ in range(args.warmup):
            _run_ref_processes(binary, path, CLI_FLAGS[cfgname], nproc)
        t = 0.0
        for _ in range(args.steps):
            t += _run_ref_processes(binary, path, CLI_FLAGS[cfgname], nproc)
    finally:
        if os.path.exists(path):
            os.unlink(path)
    value = nproc * per_proc * args.steps / t / 1e6
    sample = (f"{nproc} concurrent reference processes x {per_proc >> 20} Mi samples per step "
              f"(file in {os.path.dirname(path)} -> /dev/null), unmodified reference CLI")
    return {
        "impl": "reference",
        "metric": "input MS/s (complex IQ samples/s), whole job",
        "value": round(value, 2), "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(t / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD[cfgname], "name": cfgname, "chunk_bytes": CHUNK,
                   "parallelism": f"{nproc} processes x 2 threads on {ncpu} host cores"},
        "cpu_baseline": {"value": round(value, 2), "unit": "MS/s", "cores": min(2 * nproc, ncpu),
                         "kind": "reference", "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def _reference_arm_port(args, cfgname: str, ncpu: int) -> dict:
    """--impl reference when oracle/_ref/ did not travel: the plain-C port on every host core."""
    per_proc = max(4 << 20, min(int(120.0 * 15e6 / max(args.steps + args.warmup, 1)), 64 << 20))
    per_proc = per_proc // (CHUNK // 2) * (CHUNK // 2)
    for _ in range(args.warmup):
        _run_port_processes(cfgname, per_proc, ncpu)
    t = sum(_run_port_processes(cfgname, per_proc, ncpu) for _ in range(args.steps))
    value = ncpu * per_proc * args.steps / t / 1e6
    sample = f"{ncpu} processes x {per_proc >> 20} Mi samples per step, oracle/fm_oracle.c (reference binary absent)"
    return {
        "impl": "reference", "metric": "input MS/s (complex IQ samples/s), whole job",
        "value": round(value, 2), "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(t / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD[cfgname], "name": cfgname, "chunk_bytes": CHUNK,
                   "parallelism": f"{ncpu} processes on {ncpu} host cores"},
        "cpu_baseline": {"value": round(value, 2), "unit": "MS/s", "cores": ncpu, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(FLAGS) + QUALITY_CONFIGS,
                    help="cfg1..cfg3: the reference-compatible path (BASELINE.json configs[0..2]); quality_*: "
                         "the opt-in quality pipeline, one stream on one GPU (handed to tools/qbench.py)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--per-channel-launch", action="store_true",
                    help="one call per channel instead of one call for the GPU's whole batch of channels")
    args = ap.parse_args()
    if args.config in QUALITY_CONFIGS:
        # not a BASELINE.json workload: same JSON shape, measured by tools/qbench.py (rank 0 only)
        if int(os.environ.get("RANK", "0")) != 0 or args.impl == "reference":
            return
        cmd = [sys.executable, os.path.join(ROOT, "tools", "qbench.py"), "--config", args.config,
               "--steps", str(args.steps), "--warmup", str(args.warmup)] + (["--no-cpu"] if args.no_cpu else [])
        sys.exit(subprocess.call(cmd))
    # stdout carries exactly ONE JSON line: anything a library prints meanwhile (NCCL's version
    # banner, torchrun chatter) is sent to stderr
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    line = reference_arm(args) if args.impl == "reference" else gpu_arm(args)
    sys.stdout.flush()
    if line is not None:
        os.write(real_stdout, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
