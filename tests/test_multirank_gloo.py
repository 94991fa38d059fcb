"""World-size-2 tests of the multi-rank host logic on CPU (gloo): range sharding, byte all-gather, verdict gather,
and the point-range-split MSM orchestration with the GPU context replaced by an oracle-backed stand-in (the CUDA
path of the same function is covered by tests/test_gpu_multi.py on 2 GPUs)."""
import os
import random
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as ol
from oracle import pyref as pr


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class OracleCtx:
    """test double for bulletproofs_gadgets_b200.Context: partial sums as uncompressed-but-canonical 128-byte records
    (here: the 32-byte compressed partial padded to 128 bytes), combined with the oracle's group law."""

    def __init__(self, sG, sH):
        self.sG, self.sH = sG, sH

    def msm_gens_partial_dev(self, d_sG, d_sH, n, offset):
        lo = offset
        return ol.msm_gens(self.sG[32 * lo:32 * (lo + n)], self.sH[32 * lo:32 * (lo + n)], n, offset).ljust(128, b"\0")

    def points_sum_compress(self, ext128):
        acc = bytes(32)
        for k in range(len(ext128) // 128):
            acc = ol.point_add(acc, ext128[128 * k:128 * k + 32])
        return acc


def _worker(rank, world, port, n, sG, sH, want, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bulletproofs_gadgets_b200 import parallel
    try:
        lo, hi = parallel.shard_range(n, rank, world)
        got = parallel.msm_gens_sharded(OracleCtx(sG, sH), None, None, n)
        payloads = parallel.allgather_bytes(bytes([rank]) * 7)
        verdicts = parallel.gather_verdicts([(k, k % 3 != 0) for k, _ in parallel.shard_items(list(range(11)), rank, world)], 11)
        q.put((rank, lo, hi, got == want, payloads, verdicts))
    finally:
        dist.destroy_process_group()


def test_shard_range_partitions_exactly():
    from bulletproofs_gadgets_b200 import parallel
    for n in (0, 1, 7, 64, 1000003):
        for world in (1, 2, 3, 8):
            cuts = [parallel.shard_range(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_world_size_2_gloo_sharded_msm_and_gathers():
    rnd = random.Random(12)
    n, world = 37, 2
    sG = b"".join(rnd.randrange(pr.L).to_bytes(32, "little") for _ in range(n))
    sH = b"".join(rnd.randrange(pr.L).to_bytes(32, "little") for _ in range(n))
    want = ol.msm_gens(sG, sH, n, 0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, sG, sH, want, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 19), (19, 37)]
    assert all(r[3] for r in res)                              # both ranks reconstruct the full MSM from the partials
    assert all(r[4] == [b"\x00" * 7, b"\x01" * 7] for r in res)
    assert all(r[5] == [k % 3 != 0 for k in range(11)] for r in res)
