#!/usr/bin/env python
"""bench.py -- headline metric of the FM demodulation hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config cfg2] [--impl reference]

Metric (BASELINE.json): input MS/s (complex IQ samples per second), whole job over N GPUs, plus
the fraction of the HBM roofline the dominant kernel reaches (4 algorithmic bytes per input
sample: 2 B read + 4 B written per 2 samples).

A "step" = one pass of the hot path (fmd_demod_device: ONE launch of the fused unpack -> [input
lowpass] -> discriminator -> output lowpass kernel, whose leading CTAs also produce the chunk heads
op for op) over CHANNELS_PER_GPU independent channels of synthetic FM IQ that lie back to back in
HBM (chunks are independent, so a batch of channels is one call).  Headline workload = BASELINE.json
configs[1] (cfg2: 192 kS/s, -L12500 -l6500 -S96000, both filter stages).  N > 1: one process per
GPU (torchrun), same per-GPU work (weak scaling), no data-path collective -- channels are
independent; NCCL is used only for the start barrier and the MAX of the timings.

The ONE JSON line carries, besides the contract keys:
  configs    the other BASELINE configurations as sub-legs with the same timing rules: cfg1, cfg3
             (configs[0], [2]) and cfg4 as specified (configs[3]: 8 channels x 256 Mi samples of
             2.4 MS/s material per GPU, seeds 100+channel -- 64 channels at N = 8), each with its
             own roofline fraction and parity figure
  checks     parity_max_abs: EVERY rank compares the first signal period of EVERY channel of every
             leg with the oracle (the reference itself when oracle/_ref travelled) -- max over
             channels, legs and ranks; plus the periodicity of the full-size outputs
  roofline   burst (K launches) and `sustained` (the same launch looped for > 3 s with clocks
             sampled every 20 ms under load, and a D2D copy looped the same way beside it)
  e2e        fmd_demod_host on pinned HOST buffers (H2D + kernels + D2H inside the timed region),
             with `ceiling`: the same copies WITHOUT the kernels (fmd_copy_host), all ranks at once,
             i.e. what the host side of the box can move -- and e2e as a fraction of it
  cpu_baseline  the unmodified reference CLI, generic and -march=native builds (N = 1 only)

`--impl reference` times the reference's own CPU implementation (oracle/_ref/demodulator_ref[_native],
the unmodified peads/demodulator CLI compiled from its sources) on the host cores.
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
CHANNELS_PER_GPU = 8             # BASELINE configs[3]: 64 channels over 8 GPUs
SAMPLES_PER_CHANNEL = 128 << 20  # 128 Mi complex samples (256 MiB in, 256 MiB out) per channel
CFG4_SAMPLES_PER_CHANNEL = 256 << 20  # configs[3] as SURVEY.md 8(d) words it
BASE_CHUNKS = 16                 # synthetic signal period per channel (tiled to full length)
ALG_BYTES_PER_SAMPLE = 4.0
SUSTAINED_SECONDS = 3.2

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
    "cfg4": "independent 2.4 MS/s IQ channels, -S48000 -l15000, 8 channels x 256 Mi samples per GPU "
            "(64 channels on 8 GPUs), seeds 100+channel, f_m = 1 kHz + 50 Hz x channel",
}


def config_dict(cfgname: str) -> dict:
    """The workload, in the same words for both arms (the driver compares the two dicts)."""
    return {
        "workload": WORKLOAD[cfgname], "name": cfgname, "flags": " ".join(CLI_FLAGS[cfgname]),
        "channels_per_gpu": CHANNELS_PER_GPU, "samples_per_channel": SAMPLES_PER_CHANNEL,
        "samples_per_step_per_gpu": CHANNELS_PER_GPU * SAMPLES_PER_CHANNEL, "chunk_bytes": CHUNK,
        "parallelism": "channel-parallel (independent channels, no collective), weak scaling",
        "l2": "inputs larger than L2: 2 GiB in + 2 GiB out per GPU per step vs 126 MB L2 (no flush needed)",
    }


def measured_peak_gbs() -> tuple[float, str]:
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


def ncu_traffic(config: str, samples_per_launch: float) -> float | None:
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        e = t.get(config)
        if not e:
            return None
        # captured at e["samples_per_launch"]; scale linearly to this bench's launch size
        return float(e["dram_bytes_per_launch"]) * (samples_per_launch / float(e["samples_per_launch"]))
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int, period_ms: int = 20) -> None:
        self.gpu = gpu_index
        self.period_ms = period_ms
        self.lines: list[tuple[float, str]] = []
        self.proc = None

    def start(self) -> None:
        exe = shutil.which("nvidia-smi")
        if not exe:
            return
        self.proc = subprocess.Popen([exe, f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu), "-lms", str(self.period_ms)],
                                     stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self) -> None:
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def mark(self) -> float:
        return time.perf_counter()

    def summary(self, t0: float | None = None, t1: float | None = None) -> dict:
        """clock record of the samples taken in [t0, t1] (all samples when omitted)"""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, reasons, pw = [], [], set(), []
        for ts, ln in list(self.lines):
            if (t0 is not None and ts < t0) or (t1 is not None and ts > t1):
                continue
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
        return {"sm_mhz": statistics.median(hi), "sm_mhz_min": min(hi), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "samples": len(sm), "samples_under_load": len(hi), "power_w_max": max(pw),
                "power_w_median_under_load": statistics.median([p for p in pw if p >= 0.5 * max(pw)]),
                "sample_period_ms": self.period_ms}

    def stop(self) -> None:
        if self.proc:
            time.sleep(0.05)
            self.proc.terminate()


# ------------------------------------------------------------------------------------- GPU arm

def bind_to_gpu_numa_node(prop) -> str:
    """Run this rank (and therefore first-touch its pinned staging buffers) on the CPUs next to its
    GPU when the topology is exposed (on this pool's VMs it is not: one NUMA node, numa_node = -1)."""
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
    from oracle import pyoracle as po  # the CHECKER of the parity figures below; never timed, never the product

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

    n_ch = CHANNELS_PER_GPU
    my_channels = [rank * n_ch + i for i in range(n_ch)]  # global channel ids (weak scaling)
    peak, peak_src = measured_peak_gbs()
    checker = po.best()
    steps, warmup = args.steps, max(args.warmup, 3)
    period_out = BASE_CHUNKS * CHUNK // 4

    # the channels of one GPU lie back to back in one batch buffer: chunks are independent, so a
    # batch of channels (each a whole number of chunks) is ONE call / ONE launch of the hot path.
    # Sized for the largest leg (cfg4: 8 x 256 Mi samples = 4 GiB in, 4 GiB out).
    big = CFG4_SAMPLES_PER_CHANNEL if not args.no_sublegs else SAMPLES_PER_CHANNEL
    d_in_all = torch.empty(n_ch * 2 * big, dtype=torch.uint8, device=dev)
    d_out_all = torch.empty(n_ch * big // 2, dtype=torch.float32, device=dev)
    stream = torch.cuda.Stream(device=dev)
    sampler = ClockSampler(local, 20)
    if rank == 0:
        sampler.start()

    def run_leg(leg: str, sustained: bool) -> dict:
        """One configuration: fill the channels, warm up, time `steps` launches (CUDA events on the
        launching stream, max over ranks), check every channel against the oracle."""
        cfgname = "cfg3" if leg == "cfg4" else leg
        spc = CFG4_SAMPLES_PER_CHANNEL if leg == "cfg4" else SAMPLES_PER_CHANNEL
        sig = synth.CONFIGS[cfgname]
        ch_bytes = 2 * spc
        reps = ch_bytes // (BASE_CHUNKS * CHUNK)
        nbytes = n_ch * ch_bytes
        bases = []
        for i, c in enumerate(my_channels):
            base = synth.fm_iq_u8(BASE_CHUNKS * CHUNK // 2, sig["fs_sig"], sig["df"], sig["fm"] + 50.0 * c, 100 + c)
            bases.append(base)
            d_in_all[i * ch_bytes:(i + 1) * ch_bytes].view(reps, -1).copy_(torch.from_numpy(base).to(dev))
        d_out_all[: nbytes // 4].zero_()
        dem = fm.Demodulator(device=local, **FLAGS[cfgname])

        def step() -> None:
            dem.process_device(d_in_all.data_ptr(), nbytes, d_out_all.data_ptr(), stream.cuda_stream)

        with torch.cuda.stream(stream):
            for _ in range(warmup):
                step()
            stream.synchronize()
            dem.profile_kernels(True)
            dem.kernel_time()
            l0 = dem.launches
            barrier()
            t_mark0 = sampler.mark()
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record(stream)
            for _ in range(steps):
                step()
            ev1.record(stream)
            stream.synchronize()
            barrier()
            dev_ms = ev0.elapsed_time(ev1)
            k_ms, k_n = dem.kernel_time()
            dem.profile_kernels(False)
            launches = dem.launches - l0
        total_ms = sharding.max_over_ranks(dev_ms, dev)
        samples_gpu = n_ch * spc
        value = world * samples_gpu * steps / (total_ms * 1e-3) / 1e6  # MS/s whole job
        k_avg_ms = sharding.max_over_ranks(k_ms / max(k_n, 1), dev)
        achieved = ALG_BYTES_PER_SAMPLE * samples_gpu / (k_avg_ms * 1e-3) / 1e9

        # ---- parity on this rank: first period of every channel against the oracle; periodicity
        # property at full size (every tile of a channel's output equals its first one)
        opts = po.Config(**sig["opts"])
        worst, periodic = 0.0, True
        for i in range(n_ch):
            o = d_out_all[i * ch_bytes // 4:(i + 1) * ch_bytes // 4].view(reps, period_out)
            got = o[0].cpu().numpy()
            want = checker.run(bases[i], opts)
            worst = max(worst, float(np.abs(got - want).max()))
            for r in (reps // 2, reps - 1):
                periodic = periodic and bool(torch.equal(o[0].view(torch.int32), o[r].view(torch.int32)))
        worst = sharding.max_over_ranks(worst, dev)
        periodic = sharding.max_over_ranks(0.0 if periodic else 1.0, dev) == 0.0

        res = {
            "value": round(value, 1), "unit": "MS/s", "ms_per_step": round(total_ms / steps, 4),
            "samples_per_step_per_gpu": samples_gpu, "gpu_launches": int(launches),
            "roofline": {
                "bound": "hbm", "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": ncu_traffic(cfgname, float(samples_gpu)),
                "kernel": "fmd_fast_kernel", "kernel_ms_avg": round(k_avg_ms, 4), "kernel_launches_timed": k_n,
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_SAMPLE * samples_gpu,
                "samples_per_launch": float(samples_gpu), "peak_source": peak_src,
                "kernel_share_of_step": round(k_ms / max(dev_ms, 1e-9), 4),
            },
            "parity_max_abs": worst, "parity_checker": checker.kind, "parity_channels_per_rank": n_ch,
            "output_periodic_over_tiles": periodic,
            "_marks": (t_mark0, sampler.mark()),
        }
        if sustained:
            # the same launch looped for > 3 s: what the kernel holds once clocks and power settle
            n_sus = max(steps, int(SUSTAINED_SECONDS * 1e3 / max(k_avg_ms, 1e-3)))
            barrier()
            m0 = sampler.mark()
            ms = dem.time_device(d_in_all.data_ptr(), nbytes, d_out_all.data_ptr(), stream.cuda_stream, n_sus)
            m1 = sampler.mark()
            sus_ms = sharding.max_over_ranks(ms / n_sus, dev)
            sus_ach = ALG_BYTES_PER_SAMPLE * samples_gpu / (sus_ms * 1e-3) / 1e9
            res["roofline"]["sustained"] = {
                "seconds": round(ms * 1e-3, 2), "launches": n_sus, "kernel_ms_avg": round(sus_ms, 4),
                "achieved": round(sus_ach, 1), "frac": round(sus_ach / peak, 4),
                "clocks": sampler.summary(m0 + 0.3, m1) if rank == 0 else None,
                "_marks": (m0, m1),
            }
        res["_dem"] = dem
        res["_nbytes"] = nbytes
        return res

    headline = args.config
    head = run_leg(headline, sustained=not args.no_sustained)
    dem = head.pop("_dem")
    nbytes = head.pop("_nbytes")

    # ---- end to end through the host-facing C-ABI call, pinned host buffers; then the same copies
    # without the kernels, all ranks at once: the ceiling the host side of this box sets
    e2e_bytes = nbytes
    pin_in, pin_out = PinnedBuffer(e2e_bytes), PinnedBuffer(e2e_bytes)
    torch.from_numpy(pin_in.u8).copy_(d_in_all[:e2e_bytes])
    dem.process_ptr(pin_in.ptr, e2e_bytes, pin_out.ptr)
    e2e_steps = max(3, min(steps, 5))
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dem.process_ptr(pin_in.ptr, e2e_bytes, pin_out.ptr)
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_s = sharding.max_over_ranks(e2e_s, dev)
    e2e_samples = e2e_bytes // 2
    e2e_value = world * e2e_samples * e2e_steps / e2e_s / 1e6
    e2e_ok = bool(np.array_equal(pin_out.view(np.float32)[:period_out].view(np.uint32),
                                 d_out_all[:period_out].cpu().numpy().view(np.uint32)))
    dem.copy_ptr(pin_in.ptr, e2e_bytes, pin_out.ptr)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        dem.copy_ptr(pin_in.ptr, e2e_bytes, pin_out.ptr)
    copy_s = time.perf_counter() - t0
    barrier()
    copy_s = sharding.max_over_ranks(copy_s, dev)
    ceiling_gbs = world * e2e_bytes * e2e_steps / copy_s / 1e9  # per direction (as many bytes go each way)
    e2e_gbs = world * e2e_bytes * e2e_steps / e2e_s / 1e9
    pin_in.free(); pin_out.free()
    dem.close()

    legs = {}
    if not args.no_sublegs:
        for leg in [c for c in ("cfg1", "cfg2", "cfg3") if c != headline] + ["cfg4"]:
            r = run_leg(leg, sustained=(leg == "cfg1" and not args.no_sustained))
            r.pop("_dem").close()
            r.pop("_nbytes")
            r["workload"] = WORKLOAD[leg]
            legs[leg] = r

    # ---- the denominator, sustained, in the same run: a D2D copy of 2 GiB looped for 2 s
    copy_sus = None
    if not args.no_sustained:
        n = 2 << 30
        src, dst = d_in_all[:n], d_out_all.view(torch.uint8)[:n]
        for _ in range(3):
            dst.copy_(src)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        m0 = sampler.mark()
        a.record()
        ncopy = 0
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 2.0:
            for _ in range(50):
                dst.copy_(src)
            ncopy += 50
            torch.cuda.synchronize()
        b.record()
        torch.cuda.synchronize()
        m1 = sampler.mark()
        copy_sus = {"gbs": round(2.0 * n * ncopy / (a.elapsed_time(b) * 1e-3) / 1e9, 1), "seconds": round(a.elapsed_time(b) * 1e-3, 2),
                    "how": "torch D2D copy_ of 2 GiB (read + write bytes), looped 2 s, same process, same GPU",
                    "clocks": sampler.summary(m0 + 0.3, m1) if rank == 0 else None}

    clocks = None
    if rank == 0:
        sampler.stop()
        # the record of the headline kernel under load: burst region + its sustained loop
        sus = head["roofline"].get("sustained")
        clocks = sampler.summary(head["_marks"][0], sus["_marks"][1] if sus else head["_marks"][1])
        clocks["burst_region_samples"] = sampler.summary(*head["_marks"]).get("samples", 0)

    cpu = cpu_baseline(headline) if (rank == 0 and world == 1 and not args.no_cpu) else None

    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return None

    def clean(d: dict) -> dict:
        d = {k: v for k, v in d.items() if not k.startswith("_")}
        if "roofline" in d and "sustained" in d["roofline"]:
            d["roofline"] = dict(d["roofline"])
            d["roofline"]["sustained"] = {k: v for k, v in d["roofline"]["sustained"].items() if not k.startswith("_")}
        return d

    head = clean(head)
    legs = {k: clean(v) for k, v in legs.items()}
    if copy_sus and "sustained" in head["roofline"]:
        s = head["roofline"]["sustained"]
        s["copy_same_run"] = copy_sus
        s["frac_of_same_run_copy"] = round(s["achieved"] / copy_sus["gbs"], 4)
        if "cfg1" in legs and "sustained" in legs["cfg1"]["roofline"]:
            s1 = legs["cfg1"]["roofline"]["sustained"]
            s1["frac_of_same_run_copy"] = round(s1["achieved"] / copy_sus["gbs"], 4)
    parity_all = max([head["parity_max_abs"]] + [v["parity_max_abs"] for v in legs.values()])
    line = {
        "metric": "input MS/s (complex IQ samples/s), whole job",
        "value": head["value"],
        "unit": "MS/s",
        "n_gpus": world,
        "steps": steps,
        "warmup": warmup,
        "ms_per_step": head["ms_per_step"],
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": config_dict(headline),
        "roofline": head["roofline"],
        "cpu_baseline": cpu,
        "e2e": {"value": round(e2e_value, 1), "unit": "MS/s",
                "h2d_bytes_per_step": e2e_bytes, "d2h_bytes_per_step": e2e_bytes,
                "api": "fmd_demod_host (pinned host in/out, slices of up to 64 MiB, graded at both ends, over copy-in/compute/copy-out streams)",
                "steps": e2e_steps, "samples_per_step_per_gpu": e2e_samples,
                "gbs_per_direction": round(e2e_gbs, 1),
                "ceiling": {"gbs_per_direction": round(ceiling_gbs, 1), "unit": "GB/s",
                            "how": "fmd_copy_host: the same H2D + D2H copies, same slices and streams, no kernels, "
                                   f"all {world} rank(s) at once, same run"},
                "frac_of_ceiling": round(e2e_gbs / ceiling_gbs, 4),
                "matches_device_path": e2e_ok},
        "gpu_launches": head["gpu_launches"],
        "clocks": clocks,
        "configs": {k: {kk: vv for kk, vv in v.items()} for k, v in legs.items()},
        "checks": {"parity_max_abs": parity_all, "parity_tolerance": 1e-4, "parity_checker": head["parity_checker"],
                   "parity_scope": f"first {BASE_CHUNKS} chunks of every channel of every leg on every rank, max over ranks",
                   "parity_max_abs_headline": head["parity_max_abs"],
                   "output_periodic_over_tiles": bool(head["output_periodic_over_tiles"] and
                                                      all(v["output_periodic_over_tiles"] for v in legs.values()))},
        "arm": {"impl": "b200", "cpu_affinity": numa, "launches_per_step_per_gpu": 1,
                "parallelism": f"channel-parallel x{world}, one process per GPU, no collective"},
    }
    if parity_all > 1e-4:
        line["checks"]["FAILED"] = "parity above tolerance"
    return line


# --------------------------------------------------------------------------- CPU reference arm

def _ref_binaries() -> dict:
    """The unmodified reference CLI (built where /root/reference exists, travels as prebuilt files):
    the generic build and the -march=native one the reference's own test.sh times.  The native one
    is used only if a probe run on THIS host exits 0 and agrees with the generic one."""
    from oracle import pyoracle as po
    out = {}
    gen = po.REF_BIN
    if os.path.exists(gen) and os.access(gen, os.X_OK):
        out["generic"] = gen
    nat = gen + "_native"
    if "generic" in out and os.path.exists(nat) and os.access(nat, os.X_OK):
        d = _scratch_dir()
        src, o1, o2 = (os.path.join(d, f"fmd_probe_{os.getpid()}_{k}") for k in ("in", "g", "n"))
        try:
            from demodulator_b200 import synth
            synth.config_signal("cfg2", 8).tofile(src)
            ok = True
            for b, o in ((gen, o1), (nat, o2)):
                r = subprocess.run([b, "-i", src, "-o", o, "-S96000", "-l6500", "-L12500"], stdout=subprocess.DEVNULL,
                                   stderr=subprocess.DEVNULL)
                ok = ok and r.returncode == 0
            if ok:
                a, b = np.fromfile(o1, np.float32), np.fromfile(o2, np.float32)
                n = min(a.size, b.size) - 2 * (CHUNK // 4)  # the reference's last chunk is racy
                if n > 0 and float(np.abs(a[:n] - b[:n]).max()) <= 1e-4:
                    out["native"] = nat
        finally:
            for p in (src, o1, o2):
                if os.path.exists(p):
                    os.unlink(p)
    return out


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


def _write_signal_file(cfgname: str, path: str, samples: int) -> None:
    from demodulator_b200 import synth
    sig = synth.CONFIGS[cfgname]
    base = synth.fm_iq_u8(BASE_CHUNKS * CHUNK // 2, sig["fs_sig"], sig["df"], sig["fm"], 100)
    with open(path, "wb") as f:
        for _ in range(2 * samples // base.size):
            f.write(base.tobytes())


def cpu_baseline(cfgname: str, samples: int = 256 << 20) -> dict:
    """The reference CLI (1 reader + 1 worker thread) on a bounded sample of the same workload:
    generic build and -march=native build (the one the reference's own test.sh:91 times)."""
    bins = _ref_binaries()
    if not bins:  # reference binary did not travel: time the plain-C port, one thread
        samples = 32 << 20
        sec = min(_port_worker((cfgname, samples)) for _ in range(2))
        return {"value": round(samples / sec / 1e6, 2), "unit": "MS/s", "cores": 1, "kind": "port",
                "sample": f"{samples >> 20} Mi samples of the same signal, oracle/fm_oracle.c in-process, best of 2",
                "host_cores_available": os.cpu_count()}
    path = os.path.join(_scratch_dir(), f"fmd_bench_{os.getpid()}.u8")
    res = {}
    try:
        _write_signal_file(cfgname, path, samples)
        for kind, binary in bins.items():
            _run_ref_processes(binary, path, CLI_FLAGS[cfgname], 1)  # warm the page cache / binary
            sec = min(_run_ref_processes(binary, path, CLI_FLAGS[cfgname], 1) for _ in range(2))
            res[kind] = round(samples / sec / 1e6, 2)
    finally:
        if os.path.exists(path):
            os.unlink(path)
    best = max(res, key=res.get)
    return {"value": res[best], "unit": "MS/s", "cores": 2, "kind": "reference", "build": best,
            "generic": res.get("generic"), "native": res.get("native"),
            "sample": f"{samples >> 20} Mi samples of the same signal, file in {os.path.dirname(path)} -> /dev/null, "
                      f"unmodified reference CLI (oracle/_ref/demodulator_ref: generic x86-64; _native: -march=native as "
                      f"in the reference's test.sh:91), 1 reader + 1 worker thread, best of 2 each; value = the faster build",
            "host_cores_available": os.cpu_count()}


def reference_arm(args) -> dict | None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    cfgname = args.config
    bins = _ref_binaries()
    ncpu = os.cpu_count() or 2
    if not bins:
        return _reference_arm_port(args, cfgname, ncpu)
    nproc = max(1, ncpu // 2)                       # each reference process runs 2 threads
    steps_total = args.steps + args.warmup
    # bounded sample: aim at ~120 s for the whole run at ~30 MS/s per process
    per_proc = int(120.0 * 30e6 / max(steps_total, 1))
    per_proc = max(16 << 20, min(per_proc, 256 << 20)) // (CHUNK // 2) * (CHUNK // 2)
    path = os.path.join(_scratch_dir(), f"fmd_refarm_{os.getpid()}.u8")
    try:
        _write_signal_file(cfgname, path, per_proc)
        per_proc = os.path.getsize(path) // 2
        # which build is faster on this host with every core busy?  (one untimed step each)
        speed = {}
        for kind, binary in bins.items():
            speed[kind] = nproc * per_proc / _run_ref_processes(binary, path, CLI_FLAGS[cfgname], nproc) / 1e6
        build = max(speed, key=speed.get)
        binary = bins[build]
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
              f"(file in {os.path.dirname(path)} -> /dev/null), unmodified reference CLI, {build} build")
    return {
        "impl": "reference",
        "metric": "input MS/s (complex IQ samples/s), whole job",
        "value": round(value, 2), "unit": "MS/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(t / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(cfgname),
        "cpu_baseline": {"value": round(value, 2), "unit": "MS/s", "cores": min(2 * nproc, ncpu),
                         "kind": "reference", "build": build, "sample": sample,
                         "one_untimed_step_MSps": {k: round(v, 1) for k, v in speed.items()}},
        "e2e": {"value": round(value, 2), "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "arm": {"impl": "reference", "parallelism": f"{nproc} processes x 2 threads on {ncpu} host cores",
                "samples_per_step": nproc * per_proc},
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
        "config": config_dict(cfgname),
        "cpu_baseline": {"value": round(value, 2), "unit": "MS/s", "cores": ncpu, "kind": "port", "sample": sample},
        "e2e": {"value": round(value, 2), "unit": "MS/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "arm": {"impl": "reference", "parallelism": f"{ncpu} processes on {ncpu} host cores",
                "samples_per_step": ncpu * per_proc},
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
    ap.add_argument("--no-sublegs", action="store_true", help="headline configuration only (no cfg1/cfg3/cfg4 legs)")
    ap.add_argument("--no-sustained", action="store_true", help="skip the > 3 s sustained loops")
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
