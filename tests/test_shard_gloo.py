"""N > 1 path on CPU: two gloo ranks shard the streams of one job, render NOTHING on a GPU (there is none here) but run
the same host-side partitioning / reduction code bench.py uses, with the oracle standing in for the per-rank renderer so
that the reduced totals can be checked against a single-process run of the whole job."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import scenarios as S
from iac_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _render_count(sc, ids, n_frames):
    """output samples per channel the oracle produces for the given global stream ids"""
    total = 0
    for g in ids:
        inputs = S.synth_inputs(sc, 1, n_frames, seed=shard.stream_seed(0x1A3F, g))
        P, ramps, oramp = S.synth_params(sc, 1, n_frames, seed=0x77)
        res = S.run_oracle(sc, inputs, P, ramps, oramp)
        total += sum(c for c in res[0][0] if c > 0)
    return total


def _worker(rank, world, port, n_streams, n_frames, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = S.c5_resample()
    ids = shard.shard_streams(n_streams, rank, world)
    n = _render_count(sc, ids, n_frames)
    ms_local = 10.0 + 5.0 * rank           # rank 1 is the slow one
    dist.barrier()
    ms, tot = shard.aggregate(ms_local, n)
    q.put((rank, ids, n, ms, tot))
    dist.destroy_process_group()


def test_shards_are_a_partition():
    for world in (1, 2, 4, 8):
        for n in (1, 7, 16384):
            parts = [shard.shard_streams(n, r, world) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_two_gloo_ranks_reduce_like_one_process():
    world, n_streams, n_frames = 2, 5, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_streams, n_frames, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == [0, 2, 4] and res[1][1] == [1, 3]
    # every rank sees the same reduced values: MAX of the times, SUM of the audio
    assert all(r[3] == 15.0 for r in res)
    single = _render_count(S.c5_resample(), range(n_streams), n_frames)
    assert all(r[4] == float(single) for r in res)
    assert res[0][2] + res[1][2] == single
    v = shard.job_throughput(res[0][3], res[0][4], 48000, steps=1)
    assert np.isclose(v, single / 48000.0 / 0.015)


def test_aggregate_is_identity_without_process_group():
    assert shard.aggregate(3.5, 42) == (3.5, 42.0)
