#!/usr/bin/env python3
"""ONE proof over the ranks of a torchrun group (in-library NCCL exchange), with the per-phase trace of rank 0.
usage: BPG_TRACE=1 python -m torch.distributed.run --nproc-per-node N tools/prove_sharded.py [blocks] [reps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    if rank != 0:
        os.environ.pop("BPG_TRACE", None)
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import bulletproofs_gadgets_b200 as bpg
    from bulletproofs_gadgets_b200 import gadgets, parallel
    blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 1022
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ctx = bpg.Context(local)
    inst = gadgets.mimc_chain_instance(blocks, ctx=ctx)
    cap = 1
    while cap < inst["n"]:
        cap *= 2
    ctx.gens_ensure(cap)
    circ = gadgets.Circuit(ctx, inst["n"], inst["m"], inst["csr"])
    FAST = bpg._lib.FLAG_FAST_BLINDING
    ext = b"\x44" * 32
    for mode in ("single", "comm"):
        if mode == "comm":
            parallel.enable_comm(ctx, "cuda:%d" % local)
        circ.prove(inst, ext, FAST)
        dist.barrier()
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        for _ in range(reps):
            proof, V = circ.prove(inst, ext, FAST)
        dt = (time.perf_counter() - t0) / reps
        if rank == 0:
            print("%s: %.2f ms per proof, %d launches" % (mode, dt * 1e3, (ctx.launch_count() - l0) // reps), flush=True)
        dist.barrier()
    parallel.disable_comm(ctx)
    circ.close()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
