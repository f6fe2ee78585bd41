"""CPU, world_size 2 over gloo: the channel-parallel sharding used at N > 1 GPUs.  The data path
has no collective (channels are independent); torch.distributed only carries the start barrier and
the MAX / SUM of per-rank numbers, exactly as bench.py uses it.  The per-rank "device" work is
stood in for by the oracle port so that the partitioning can be checked end to end on CPU."""
from __future__ import annotations

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from demodulator_b200 import sharding, synth

B = synth.CHUNK_BYTES


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, n_channels: int, q) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import pyoracle as po
    P = po.port()
    cfg = po.Config(fs=48000.0, fc_out=15000.0)
    mine = sharding.assign_channels(n_channels, world, rank)
    digest = {}
    for c in mine:
        raw = synth.fm_iq_u8(B // 2, 2.4e6, 75e3, 1e3 + 50.0 * c, 100 + c)
        digest[c] = float(np.abs(P.run(raw, cfg)).sum())
    dist.barrier()
    t_max = sharding.max_over_ranks(1.0 + rank)          # the slowest rank defines the step time
    n_sum = sharding.sum_over_ranks(float(len(mine)))    # whole-job units
    gathered = [None] * world
    dist.all_gather_object(gathered, digest)
    if rank == 0:
        q.put((t_max, n_sum, gathered))
    dist.destroy_process_group()


def test_channel_sharding_world2():
    world, n_channels = 2, 5
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_channels, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    t_max, n_sum, gathered = q.get()
    assert t_max == 2.0 and n_sum == n_channels
    merged = {}
    for d in gathered:
        assert not (set(d) & set(merged)), "a channel was processed twice"
        merged.update(d)
    assert sorted(merged) == list(range(n_channels))
    # same answer as one process doing everything
    from oracle import pyoracle as po
    P = po.port()
    for c in range(n_channels):
        raw = synth.fm_iq_u8(B // 2, 2.4e6, 75e3, 1e3 + 50.0 * c, 100 + c)
        assert merged[c] == float(np.abs(P.run(raw, po.Config(fs=48000.0, fc_out=15000.0))).sum())


def test_assign_channels_partition():
    for world in (1, 2, 4, 8):
        seen = []
        for r in range(world):
            seen += sharding.assign_channels(64, world, r)
        assert sorted(seen) == list(range(64))
        assert max(len(sharding.assign_channels(64, world, r)) for r in range(world)) == 64 // world


def test_split_stream_covers_in_order():
    """One long stream: chunk-aligned batches round-robin; no overlap, no hand-off state."""
    for world, n_chunks, batch in ((2, 10, 3), (8, 122071, 128), (4, 7, 16)):
        got = []
        for r in range(world):
            got += sharding.split_stream(n_chunks, world, r, batch)
        got.sort()
        pos = 0
        for first, n in got:
            assert first == pos and n > 0
            pos += n
        assert pos == n_chunks


def test_reference_arm_runs_on_rank_0_only():
    """Under torchrun the reference arm of bench.py is rank 0's job: every other rank exits 0 without
    work and without a line (the driver reads exactly one JSON line per run)."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(_free_port()))
    p = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "1"], env=env, capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stderr[-400:]
    assert p.stdout.strip() == ""
