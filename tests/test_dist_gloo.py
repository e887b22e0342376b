"""CPU, world_size 2, gloo: the N > 1 host logic -- env sharding, episode-statistic all-reduce, opponent weight
broadcast.  (The env step itself has no collective; its sharded determinism is covered on the GPU by
test_env_offset_keeps_rng_streams_per_env.)"""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from aircombat_selfplay_b200 import dist as adist


def test_env_slices_are_contiguous_and_cover():
    for total in (1, 7, 32, 4096, 65536, 65537):
        for ws in (1, 2, 3, 4, 8):
            pos = 0
            for r in range(ws):
                off, cnt = adist.env_slice(total, ws, r)
                assert off == pos and cnt >= total // ws
                pos += cnt
            assert pos == total


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, q):
    os.environ.update({"WORLD_SIZE": str(ws), "RANK": str(rank), "LOCAL_RANK": str(rank), "MASTER_ADDR": "127.0.0.1",
                       "MASTER_PORT": str(port)})
    adist.init("gloo")
    off, cnt = adist.env_slice(101, ws, rank)
    stats = adist.all_reduce_stats({"episodes": cnt, "reward_sum": float(rank + 1), "offset_check": off})
    from aircombat_selfplay_b200.controller import LowLevelController
    torch.manual_seed(rank)             # different weights per rank before the broadcast
    m = LowLevelController()
    adist.broadcast_module(m, src=0)
    checksum = float(sum(p.double().sum() for p in m.parameters()))
    q.put((rank, stats, checksum))
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(120)
def test_all_reduce_and_broadcast_world_size_2():
    ws, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=100) for _ in range(ws))
    for p in procs:
        p.join(timeout=30)
        assert p.exitcode == 0
    (r0, s0, c0), (r1, s1, c1) = res
    assert s0 == s1 and s0["episodes"] == 101 and s0["reward_sum"] == 3.0 and s0["offset_check"] == 51
    assert c0 == c1                      # rank 1 now holds rank 0's weights
