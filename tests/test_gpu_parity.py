"""GPU parity: the CUDA path, called through the C ABI, against the reference's CPU path.

Bar (BASELINE.json north_star): chunk boundaries and decimation indices identical,
float samples within max-abs 1e-4 of the reference output."""
from __future__ import annotations

import numpy as np
import pytest

import demodulator_b200 as fm
from demodulator_b200 import synth
from oracle import pyoracle as po
from helpers import B, GOLDEN_CASES, TOL, check_close, demod_kwargs, golden_input, load_golden, snr_db

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_golden_fixture(name, oracle_best):
    """Committed fixtures generated from the reference CLI + the oracle on the same bytes."""
    g, cfg = load_golden(name)
    raw = golden_input(g)
    with fm.Demodulator(**demod_kwargs(cfg)) as d:
        out = d.process(raw)
    n_out = B // 4
    assert out.size == raw.size // 4
    o = out.reshape(-1, n_out)
    check_close(o[:, :64], g["head"], what=f"{name} head")
    check_close(o[:, -64:], g["tail"], what=f"{name} tail")
    check_close(out[::97], g["strided"], what=f"{name} strided")
    err = check_close(out, oracle_best.run(raw, cfg), what=f"{name} full")
    print(f"{name}: max-abs {err:.2e}")


@pytest.mark.parametrize("name", ["edge", "edge_L"])
def test_edge_bytes(name):
    """int8 wrap (255 -> -128), signed zeros, +-pi cut, runs of exact zeros, at chunk edges."""
    g, cfg = load_golden(name)
    raw, want = g["raw"], g["out_full"]
    with fm.Demodulator(**demod_kwargs(cfg)) as d:
        out = d.process(raw)
    if name == "edge":
        check_close(out, want, what=name)
    else:
        # random bytes through the input filter put some products next to 0 where atan2 is
        # ill-conditioned (the reference's own -march=native build differs there too): require
        # 1e-4 wherever the discriminator input is not tiny, and 80 dB SNR overall
        bad = np.abs(out - want) > TOL
        assert bad.mean() < 2e-4, f"{bad.sum()} samples off"
        assert snr_db(out[~bad], want[~bad]) > 80.0


@pytest.mark.parametrize("name", ["cfg1", "cfg2", "d5_L", "D7_d4_L", "m1_d4", "q2", "q2_L"])
def test_exact_kernel_matches_oracle_tightly(name, oracle_best):
    """FMD_FLAG_EXACT evaluates the section sweep op for op: only atan2 differs from libm."""
    g, cfg = load_golden(name)
    raw = golden_input(g)[:B]
    with fm.Demodulator(exact=True, **demod_kwargs(cfg)) as d:
        out = d.process(raw)
    want = oracle_best.run(raw, cfg)
    # head outputs with -L are amplified arena garbage (|v| up to ~600): relative there
    scale = np.maximum(1.0, np.abs(want))
    err = float((np.abs(out - want) / scale).max())
    assert err < 5e-6, err


def test_chunks_are_independent(oracle_best):
    """Chunk boundaries: processing N chunks at once == processing each alone (bit for bit)."""
    raw = synth.config_signal("cfg2", 5)
    cfg = po.Config(**synth.CONFIGS["cfg2"]["opts"])
    with fm.Demodulator(**demod_kwargs(cfg)) as d:
        whole = d.process(raw).copy()
        for c in (0, 2, 4):
            one = d.process(raw[c * B:(c + 1) * B])
            assert np.array_equal(one.view(np.uint32), whole[c * B // 4:(c + 1) * B // 4].view(np.uint32))
    check_close(whole, oracle_best.run(raw, cfg), what="5 chunks")


@pytest.mark.parametrize("shift", [-9, -8, -6, -3, -1, 1])
@pytest.mark.parametrize("with_in", [False, True])
def test_bufsize_option(shift, with_in, oracle_port):
    """-b: chunk sizes from 512 B to 512 KiB (tiny chunks take the exact kernel)."""
    bs = (B >> -shift) if shift < 0 else (B << shift)
    raw = np.resize(synth.config_signal("cfg2", 1), 3 * bs)
    cfg = po.Config(fs=96000.0, fc_out=6500.0, fc_in=12500.0 if with_in else 0.0)
    with fm.Demodulator(bufSize=bs, **demod_kwargs(cfg)) as d:
        out = d.process(raw)
    check_close(out, oracle_port.run(raw, cfg, chunk_bytes=bs), what=f"bufSize {bs}")


@pytest.mark.parametrize("tail", [4, 12, 508, 512, 1000, 131072 + 36, B - 4])
@pytest.mark.parametrize("with_in", [False, True])
def test_ragged_tail_policy(tail, with_in, oracle_port):
    """A final partial chunk of r bytes is a chunk of length r with the same edge rules."""
    raw = synth.config_signal("cfg1", 2)[: B + tail]
    cfg = po.Config(fs=96000.0, fc_out=12500.0, fc_in=12500.0 if with_in else 0.0)
    with fm.Demodulator(**demod_kwargs(cfg)) as d:
        out = d.process(raw)
    assert out.size == (B + tail) // 4
    check_close(out[:B // 4], oracle_port.run(raw[:B], cfg), what="whole chunk")
    if tail // 4 >= 12:  # the oracle's arena model needs a chunk longer than the head it re-reads
        check_close(out[B // 4:], oracle_port.run(raw[B:], cfg, chunk_bytes=tail), what=f"tail {tail}")


def test_ragged_bytes_dropped():
    raw = synth.config_signal("cfg1", 1)
    with fm.Demodulator(sampleRate=96000.0) as d:
        a = d.process(raw[: B - 3])
        b = d.process(raw[: B - 4])
    assert a.size == b.size == (B - 4) // 4
    assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_empty_input():
    with fm.Demodulator() as d:
        assert d.process(np.zeros(0, np.uint8)).size == 0
        assert d.process(np.zeros(3, np.uint8)).size == 0


def test_device_path_equals_host_path():
    import torch
    raw = synth.config_signal("cfg2", 4)
    cfg = po.Config(**synth.CONFIGS["cfg2"]["opts"])
    with fm.Demodulator(**demod_kwargs(cfg)) as d:
        host = d.process(raw).copy()
        d_in = torch.from_numpy(raw).cuda()
        d_out = torch.empty(raw.size // 4, dtype=torch.float32, device="cuda")
        s = torch.cuda.current_stream()
        d.process_device(d_in.data_ptr(), raw.size, d_out.data_ptr(), s.cuda_stream)
        s.synchronize()
        assert np.array_equal(d_out.cpu().numpy().view(np.uint32), host.view(np.uint32))
        assert d.launches > 0


def test_multi_slice_host_pipeline(oracle_best):
    """Input larger than the staging slots: slices overlap copy and compute; order preserved."""
    base = synth.config_signal("cfg3", 8)
    raw = np.tile(base, 40)  # 80 MiB -> 3 slices of 32 MiB
    cfg = po.Config(**synth.CONFIGS["cfg3"]["opts"])
    with fm.Demodulator(**demod_kwargs(cfg)) as d:
        out = d.process(raw)
    want = oracle_best.run(base, cfg)
    per = want.size
    for rep in (0, 15, 16, 31, 39):  # periods that straddle slice boundaries
        check_close(out[rep * per:(rep + 1) * per], want, what=f"period {rep}")
    first = out[:per]
    assert all(np.array_equal(first.view(np.uint32), out[r * per:(r + 1) * per].view(np.uint32)) for r in range(40))


def test_ring_api_pinned(oracle_best):
    from demodulator_b200.api import PinnedBuffer
    raw = synth.config_signal("cfg1", 6)
    cfg = po.Config(**synth.CONFIGS["cfg1"]["opts"])
    slot = 2 * B
    pin_in, pin_out = PinnedBuffer(3 * slot), PinnedBuffer(3 * slot)
    pin_in.u8[:] = raw
    with fm.Demodulator(**demod_kwargs(cfg)) as d:
        d.ring_create(3, slot)
        for i in range(3):
            d.submit(i, pin_in.ptr + i * slot, slot, pin_out.ptr + i * slot)
        with pytest.raises(fm.FmdError):
            d.submit(0, pin_in.ptr, slot, pin_out.ptr)  # still busy
        for i in range(3):
            d.wait(i)
        out = pin_out.view(np.float32)[: raw.size // 4].copy()
    check_close(out, oracle_best.run(raw, cfg), what="ring")
    pin_in.free()
    pin_out.free()


def test_rejects_what_the_reference_leaves_undefined():
    with pytest.raises(fm.FmdError):
        fm.Demodulator(inFilterDegree=3)  # -D without -L
    with pytest.raises(fm.FmdError):
        fm.Demodulator(bufSize=6)
    with pytest.raises(fm.FmdError):
        fm.Demodulator(mode=0x4, bufSize=262144 + 4)  # -q1 with a chunk its scan cannot tile
    with pytest.raises(fm.FmdError):
        fm.Demodulator(mode=0x8, bufSize=16384)  # -q2 with a scratch buffer small enough to be a recycled heap chunk


def test_q1_offsets_live_on_across_calls(oracle_port):
    """-q1 (correctIq): the running offsets are stream state -- calls of whole chunks continue it,
    reset() starts over (the reference keeps them in a function-local static)"""
    cfg = po.Config(fs=96000.0, fc_out=12500.0, mode=4)
    raw = synth.config_signal("cfg1", 7)
    want = oracle_port.run(raw, cfg)
    with fm.Demodulator(**demod_kwargs(cfg)) as d:
        for _ in range(2):
            out = np.concatenate([d.process(raw[:B]), d.process(raw[B:4 * B]), d.process(raw[4 * B:])])
            check_close(out, want, what="-q1 in three calls")
            d.reset()
        with pytest.raises(fm.FmdError, match="whole chunks"):
            d.process(raw[:B + 1000])


def test_atan2_special_values():
    """Signed zeros and the +-pi cut, via crafted sample pairs."""
    pat = np.full(B, 127, np.uint8)
    pairs = {0: (130, 127, 124, 127), 1: (124, 127, 130, 127), 2: (127, 127, 122, 122), 3: (127, 127, 127, 127),
             4: (255, 255, 0, 0), 5: (0, 0, 255, 255), 6: (127, 130, 127, 124), 7: (128, 127, 126, 127)}
    for k, v in pairs.items():
        pat[4 * (100 + 2 * k): 4 * (100 + 2 * k) + 4] = v
    cfg = po.Config(fs=96000.0, fc_out=12500.0)
    want = po.port().run(pat, cfg)
    with fm.Demodulator(**demod_kwargs(cfg)) as d:
        out = d.process(pat)
    check_close(out, want, tol=2e-6, what="special values")


@pytest.mark.parametrize("seed", range(24))
def test_random_filter_configurations(seed, oracle_port):
    """-d/-D/-m/-e/-l/-L/-S/-b drawn at random (both families, orders 1..8, chunk sizes from 4 KiB
    to 512 KiB, ragged tails): stencil widths 1..5 on either filter, both atan polynomial lengths."""
    rng = np.random.default_rng(1000 + seed)
    deg_out = int(rng.integers(1, 9))
    with_in = bool(rng.integers(0, 2))
    deg_in = int(rng.integers(0, 9)) if with_in else 0
    mode = int(rng.integers(0, 4)) if with_in else int(rng.integers(0, 2))
    fs = float(rng.choice([48000.0, 96000.0, 125000.0, 250000.0]))
    fc_out = float(rng.uniform(0.04, 0.3) * fs)
    fc_in = float(rng.uniform(0.05, 0.3) * fs) if with_in else 0.0
    eps = float(rng.choice([0.3, 0.2, 0.1]))
    bs = int(rng.choice([4096, 65536, B, 2 * B]))
    nbytes = int(rng.integers(1, 4)) * bs + int(rng.choice([0, 0, 4 * int(rng.integers(64, bs // 8))]))
    raw = np.resize(synth.config_signal("cfg2", 2, seed=seed), nbytes)
    cfg = po.Config(fs=fs, fc_in=fc_in, fc_out=fc_out, eps=eps, deg_in=deg_in, deg_out=deg_out, mode=mode)
    whole = nbytes // bs * bs
    want = [oracle_port.run(raw[:whole], cfg, chunk_bytes=bs)]
    if nbytes > whole:
        want.append(oracle_port.run(raw[whole:], cfg, chunk_bytes=nbytes - whole))
    want = np.concatenate(want)
    with fm.Demodulator(bufSize=bs, **demod_kwargs(cfg)) as d:
        out = d.process(raw)
    assert out.size == want.size
    scale = float(np.abs(want[np.isfinite(want)]).max(initial=1.0))
    check_close(out, want, tol=1e-4 + 3e-6 * scale, what=f"seed {seed}: {cfg} bs={bs} n={nbytes}")


@pytest.mark.parametrize("with_in", [False, True])
@pytest.mark.parametrize("how", ["fused", "three_pass_by_env", "three_pass_by_chunk_size", "many_runs"])
def test_q1_forms_agree_with_the_oracle(how, with_in, oracle_port, monkeypatch):
    """-q1 has two forms: the fused kernel (a CTA walks runs of 8 tiles of 2048 steps: 64 KiB chunks are
    exactly one run, 256 KiB chunks four, 8 KiB chunks make a run span 8 chunks) and the three-pass
    fallback (designs / chunk sizes without a fused form, or FMD_Q1_FUSED=0); both continue the
    stream's offsets from call to call."""
    bs = {"fused": B, "three_pass_by_env": B, "three_pass_by_chunk_size": 36864, "many_runs": 8192}[how]
    if how == "three_pass_by_env":
        monkeypatch.setenv("FMD_Q1_FUSED", "0")
    cfg = po.Config(fs=96000.0, fc_out=6500.0, fc_in=12500.0 if with_in else 0.0, mode=4)
    n = {B: 5, 36864: 23, 8192: 77}[bs]
    raw = np.resize(synth.config_signal("cfg2", 5, seed=11), n * bs)
    want = oracle_port.run(raw, cfg, chunk_bytes=bs)
    cut = (n // 3) * bs
    with fm.Demodulator(bufSize=bs, **demod_kwargs(cfg)) as d:
        out = np.concatenate([d.process(raw[:cut]), d.process(raw[cut:])])
        launches = d.launches
    scale = float(np.abs(want).max(initial=1.0))
    check_close(out, want, tol=1e-4 + 3e-6 * scale, what=f"-q1 {how} bs={bs}")
    # the fused form: run ends, scan, fused kernel, heads = 4 launches per call; three-pass: 4 as well, but
    # different kernels -- what tells them apart cheaply is that both must be right


@pytest.mark.parametrize("seed", range(12))
def test_random_conditioner_configurations(seed, oracle_port):
    """the -q conditioners under random filters: -q1 (whole chunks, offsets carried from chunk to
    chunk), -q2 (highpass sweep + its spill into the next stage; ragged tails) and -q3, with and
    without the input filter, degrees 1..8, both families"""
    rng = np.random.default_rng(5000 + seed)
    q = int(1 + seed % 3)
    deg_out = int(rng.integers(1, 9))
    with_in = bool(rng.integers(0, 2))
    deg_in = int(rng.integers(0, 9)) if with_in else 0
    mode = (int(rng.integers(0, 4)) if with_in else int(rng.integers(0, 2))) | (q << 2)
    fs = float(rng.choice([48000.0, 96000.0, 250000.0]))
    fc_out = float(rng.uniform(0.04, 0.3) * fs)
    fc_in = float(rng.uniform(0.05, 0.3) * fs) if with_in else 0.0
    bs = int(rng.choice([65536, B, 2 * B]))
    nbytes = int(rng.integers(1, 4)) * bs
    if q != 1:
        nbytes += int(rng.choice([0, 4 * int(rng.integers(64, bs // 8))]))
    raw = np.resize(synth.config_signal("cfg2", 2, seed=seed), nbytes)
    cfg = po.Config(fs=fs, fc_in=fc_in, fc_out=fc_out, eps=0.2, deg_in=deg_in, deg_out=deg_out, mode=mode)
    whole = nbytes // bs * bs
    want = [oracle_port.run(raw[:whole], cfg, chunk_bytes=bs)]
    if nbytes > whole:
        want.append(oracle_port.run(raw[whole:], cfg, chunk_bytes=nbytes - whole))
    want = np.concatenate(want)
    with fm.Demodulator(bufSize=bs, **demod_kwargs(cfg)) as d:
        out = d.process(raw)
    assert out.size == want.size
    scale = float(np.abs(want[np.isfinite(want)]).max(initial=1.0))
    check_close(out, want, tol=1e-4 + 3e-6 * scale, what=f"seed {seed}: {cfg} bs={bs} n={nbytes}")


@pytest.mark.parametrize("nbytes,bs,with_in,mode", [(3 * B, B, True, 0), (B + 1000 * 4, B, True, 0), (B + 516, B, False, 0),
                                                    (5 * 4096 + 260, 4096, True, 0), (7 * 512, 512, False, 0),
                                                    (2 * B - 4, B, True, 0), (640, 64, True, 0),
                                                    # the paths that stream since round 2: -q1 fused (one run per chunk; runs that span
                                                    # chunks; three-pass form), -q2 / -q3 under -L
                                                    (3 * B, B, True, 4), (11 * 8192, 8192, False, 4), (5 * 36864, 36864, True, 4),
                                                    (2 * B + 2000 * 4, B, True, 8), (2 * B + 132 * 4, B, True, 12)])
def test_no_out_of_bounds_writes(nbytes, bs, with_in, mode):
    """compute-sanitizer is closed on this pool, so: guard bands around the output (and the input
    placed flush against the end of its allocation) must come back untouched for ragged sizes,
    tiny chunks and whole chunks alike."""
    import torch
    guard = 4096
    raw = np.resize(synth.config_signal("cfg1", 1), nbytes)
    n_out = nbytes // 4
    d_in = torch.from_numpy(raw).cuda()
    canvas = torch.full((guard + n_out + guard,), 12345.0, dtype=torch.float32, device="cuda")
    cfg = po.Config(fs=96000.0, fc_out=6500.0, fc_in=12500.0 if with_in else 0.0, mode=mode)
    with fm.Demodulator(bufSize=bs, **demod_kwargs(cfg)) as d:
        s = torch.cuda.current_stream()
        d.process_device(d_in.data_ptr(), nbytes, canvas.data_ptr() + 4 * guard, s.cuda_stream)
        s.synchronize()
    assert bool((canvas[:guard] == 12345.0).all()) and bool((canvas[guard + n_out:] == 12345.0).all())
    out = canvas[guard:guard + n_out].cpu().numpy()
    assert not np.any(out == 12345.0), "an output was never written"


@pytest.mark.parametrize("name", ["cfg2", "cfg3"])
def test_full_size_properties(name, oracle_best):
    """BASELINE size (1 Gi samples = 2 GiB in, 2 GiB out on the device, far beyond L2) through
    size-independent properties: the input is a 16-chunk period tiled 512 times, chunks are
    independent, so the output must be the same 16-chunk period 512 times, bit for bit; and that
    period must match the reference."""
    import torch
    period_chunks, reps = 16, 512
    base = synth.config_signal(name, period_chunks)
    cfg = po.Config(**synth.CONFIGS[name]["opts"])
    d_in = torch.from_numpy(base).cuda().repeat(reps)
    d_out = torch.empty(d_in.numel() // 4, dtype=torch.float32, device="cuda")
    with fm.Demodulator(**demod_kwargs(cfg)) as d:
        s = torch.cuda.current_stream()
        d.process_device(d_in.data_ptr(), d_in.numel(), d_out.data_ptr(), s.cuda_stream)
        s.synchronize()
    o = d_out.view(reps, -1).view(torch.int32)
    assert bool((o == o[0:1]).all()), "tiles differ: a chunk saw state from outside itself"
    check_close(d_out[: base.size // 4].cpu().numpy(), oracle_best.run(base, cfg), what=f"{name} first period")
    # checksum of checksums: per-chunk sums of |y| repeat with the period
    sums = d_out.view(reps * period_chunks, -1).abs().sum(dim=1).view(reps, period_chunks)
    assert bool((sums == sums[0:1]).all())


def test_contexts_are_independent_across_threads(oracle_best):
    """One context per (GPU, channel), no global state: four host threads, four different filter
    set-ups, the same GPU, all at once."""
    import threading
    setups = [("cfg1", dict(fs=96000.0, fc_out=12500.0)),
              ("cfg2", dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0)),
              ("cfg3", dict(fs=48000.0, fc_out=15000.0)),
              ("cfg2", dict(fs=96000.0, fc_out=12500.0, fc_in=12500.0, deg_out=5, mode=3))]
    results, errors = {}, []

    def work(i, sig, kw):
        try:
            raw = synth.config_signal(sig, 6, seed=50 + i)
            cfg = po.Config(**kw)
            with fm.Demodulator(**demod_kwargs(cfg)) as d:
                for _ in range(3):
                    out = d.process(raw)
            results[i] = (raw, cfg, out.copy())
        except Exception as e:  # pragma: no cover
            errors.append(e)

    ts = [threading.Thread(target=work, args=(i, s, k)) for i, (s, k) in enumerate(setups)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errors, errors
    for i, (raw, cfg, out) in results.items():
        check_close(out, oracle_best.run(raw, cfg), what=f"thread {i}")


@pytest.mark.parametrize("bs", [2000, 10000, 1028, 516, 262144 + 512, 36])
@pytest.mark.parametrize("with_in", [False, True])
def test_chunk_sizes_that_are_not_powers_of_two(bs, with_in, oracle_port):
    """The C ABI takes any chunk size that is a multiple of 4 bytes (the CLI only makes powers of
    two): several chunks whose length is not a multiple of the 128-output pass, so every chunk
    has a streaming part, a head and a ragged remainder -- or is too small for the streaming part."""
    raw = np.resize(synth.config_signal("cfg2", 2), 3 * bs + (bs // 2 & ~3))
    cfg = po.Config(fs=96000.0, fc_out=6500.0, fc_in=12500.0 if with_in else 0.0)
    with fm.Demodulator(bufSize=bs, **demod_kwargs(cfg)) as d:
        out = d.process(raw)
    want = oracle_port.run(raw[:3 * bs], cfg, chunk_bytes=bs)
    check_close(out[: want.size], want, what=f"bufSize {bs}")
    assert out.size == raw.size // 4


def test_tiny_chunks_every_size(oracle_best):
    """-b down to 4-byte chunks (one output per chunk): the exact path on chunk sizes 4..128 bytes,
    against the reference's own code where its prebuilt library is present."""
    rng = np.random.default_rng(3)
    for bs in range(4, 132, 4):
        raw = rng.integers(0, 256, 5 * bs, dtype=np.uint8)
        for kw in (dict(fs=96000.0, fc_out=12500.0), dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0),
                   dict(fs=96000.0, fc_out=6500.0, fc_in=12500.0, deg_out=5)):
            cfg = po.Config(**kw)
            lin = (max(cfg.deg_in, cfg.deg_out if cfg.fc_in else 0) + 1) // 2
            if cfg.fc_in and bs // 2 < 2 * lin + 2:
                with pytest.raises(fm.FmdError):
                    fm.Demodulator(bufSize=bs, **demod_kwargs(cfg))
                continue
            want = oracle_best.run(raw, cfg, chunk_bytes=bs)
            with fm.Demodulator(bufSize=bs, **demod_kwargs(cfg)) as d:
                out = d.process(raw)
            scale = max(1.0, float(np.abs(want).max()))
            # random bytes: |z| can be tiny after the input filter, where atan2 is ill-conditioned
            bad = np.abs(out - want) > 2e-5 * scale
            assert bad.mean() <= 0.02, (bs, kw, int(bad.sum()))


@pytest.mark.parametrize("shift", [6, 10])
def test_large_bufsize(shift, oracle_best):
    """-b6 / -b10: chunks of 16 MiB and 256 MiB (2^22 and 2^26 outputs per chunk): one head, one
    zero extension, thousands of segments inside a chunk."""
    bs = B << shift
    raw = np.resize(synth.config_signal("cfg2", 32), bs + (bs >> 3))  # one whole chunk + a ragged eighth
    cfg = po.Config(fs=96000.0, fc_out=6500.0, fc_in=12500.0)
    with fm.Demodulator(bufSize=bs, **demod_kwargs(cfg)) as d:
        out = d.process(raw)
    assert out.size == raw.size // 4
    want = oracle_best.run(raw[:bs], cfg, chunk_bytes=bs)
    check_close(out[: bs // 4], want, what=f"bufSize {bs}")
    tail = oracle_best.run(raw[bs:], cfg, chunk_bytes=raw.size - bs)
    check_close(out[bs // 4:], tail, what=f"bufSize {bs} tail")
