"""N > 1 host logic on CPU: two gloo ranks partition a batch of streams with soundkit_b200.shard, each
normalises ITS shard (the oracle stands in for the device here — this test is about the partition, the
max-over-ranks timing reduction and the per-stream seeding, not about arithmetic), and the gathered result must
equal the single-process result stream for stream."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from soundkit_b200 import shard

N_STREAMS, FRAMES = 5, 4800  # odd on purpose: ranks get 2 and 3 streams


def make_stream(stream):
    rng = np.random.default_rng(shard.stream_seed(0x5EED0000, 4, stream))
    return rng.integers(0, 256, FRAMES * 2 * 3, dtype=np.uint8)  # s24le stereo


def normalise(oracle, raw):
    return oracle.normalize_chain_mono(raw, 24, oracle.SIGNED, False, 2, 48000, 16000)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle as O
        O.build()
        lo, hi = shard.stream_range(N_STREAMS, world, rank)
        outs = {s: normalise(O, make_stream(s)) for s in range(lo, hi)}
        dist.barrier()
        # pretend rank r took (r + 1) ms: the reported step time is the slowest rank's
        slowest = shard.max_over_ranks(float(rank + 1), dist)
        counts = shard.gather_counts(sum(v.size for v in outs.values()), dist)
        gathered = [None] * world
        dist.all_gather_object(gathered, {s: v.tobytes() for s, v in outs.items()})
        if rank == 0:
            q.put((slowest, counts, gathered))
    finally:
        dist.destroy_process_group()


def test_stream_range_is_a_partition():
    for n in (0, 1, 5, 32, 32768):
        for world in (1, 2, 4, 8):
            r = [shard.stream_range(n, world, g) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[g][1] == r[g + 1][0] for g in range(world - 1))
            sizes = [hi - lo for lo, hi in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.stream_range(4, 2, 2)
    assert shard.max_over_ranks(3.5) == 3.5 and shard.gather_counts(7) == [7]


def test_two_gloo_ranks_cover_the_batch_without_a_collective_on_the_data_path(oracle):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    slowest, counts, gathered = q.get()
    assert slowest == 2.0
    merged = {}
    for part in gathered:
        assert not (set(part) & set(merged)), "a stream was processed by two ranks"
        merged.update(part)
    assert sorted(merged) == list(range(N_STREAMS))
    single = {s: normalise(oracle, make_stream(s)) for s in range(N_STREAMS)}
    assert sum(counts) == sum(v.size for v in single.values())
    for s in range(N_STREAMS):
        assert merged[s] == single[s].tobytes()
