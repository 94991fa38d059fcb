"""Multi-GPU test (needs >= 2 CUDA devices; run with `gpurun --gpus 2`): one MSM split by point range over 2 ranks with
NCCL all-gather of the 128-byte partial points, and rank-sharded verification verdicts."""
import os
import random
import socket

import pytest
import torch

import oracle_lib as ol
from oracle import pyref as pr

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, sG, sH, want, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import bulletproofs_gadgets_b200 as bpg
    from bulletproofs_gadgets_b200 import parallel
    try:
        ctx = bpg.Context(rank)
        ctx.gens_ensure(4096)
        lo, hi = parallel.shard_range(n, rank, world)
        dG, dH = ctx.dev_alloc(32 * (hi - lo)), ctx.dev_alloc(32 * (hi - lo))
        ctx.dev_upload(dG, sG[32 * lo:32 * hi])
        ctx.dev_upload(dH, sH[32 * lo:32 * hi])
        got = parallel.msm_gens_sharded(ctx, dG, dH, n, device="cuda:%d" % rank)  # torch all-gather of the partial points
        parallel.enable_comm(ctx, "cuda:%d" % rank)                              # the library's own NCCL communicator
        got2 = parallel.msm_gens_sharded(ctx, dG, dH, n, device="cuda:%d" % rank)
        parallel.disable_comm(ctx)
        q.put((rank, got == want and got2 == want))
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_msm_point_range_split_over_two_gpus():
    import torch.multiprocessing as mp
    rnd = random.Random(13)
    n, world = 3001, 2
    sG = b"".join(rnd.randrange(pr.L).to_bytes(32, "little") for _ in range(n))
    sH = b"".join(rnd.randrange(pr.L).to_bytes(32, "little") for _ in range(n))
    want = ol.msm_gens(sG, sH, n, 0)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, sG, sH, want, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]


def _shard_worker(rank, world, port, nmul, seed, want, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    import bulletproofs_gadgets_b200 as bpg
    import bulletproofs_gadgets_b200._lib as lb
    from bulletproofs_gadgets_b200 import parallel
    import circuits
    import test_gpu_r1cs as tr
    try:
        ctx = bpg.Context(rank)
        ctx.gens_ensure(4096)
        inst = circuits.chain_instance(nmul, seed)
        ok = True
        for mode in ("callback", "comm"):  # caller-supplied all-gather (bpg_ctx_set_shard) / in-library NCCL (bpg_comm_init)
            keep = parallel.enable_sharded_prover(ctx, "cuda:%d" % rank) if mode == "callback" else parallel.enable_comm(ctx, "cuda:%d" % rank)
            for flags in (lb.FLAG_FORCE_LATE_FOLD, lb.FLAG_NO_LATE_FOLD):
                got = tr.gpu_prove(ctx, inst, bytes(range(32)), flags)
                ok = ok and got == want
            if mode == "callback":
                parallel.disable_sharded_prover(ctx)
            else:
                parallel.disable_comm(ctx)
            del keep
        q.put((rank, ok))
        ctx.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("nmul", [9, 2100])
def test_one_proof_sharded_over_two_gpus(nmul):
    """bpg_ctx_set_shard: every MSM of ONE proof cut by point range over 2 ranks, partial points all-gathered by NCCL.  Both ranks
    must return exactly the oracle's proof bytes (with and without the late fold; 9 multipliers: slices of a few terms, rank 1
    sometimes empty; 2100: the 2^15-bucket path)."""
    import torch.multiprocessing as mp
    import circuits
    import test_gpu_r1cs as tr
    inst = circuits.chain_instance(nmul, 41)
    want = tr.oracle_prove(inst, 4096, bytes(range(32)))
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, nmul, 41, want, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert res == [(0, True), (1, True)]
