"""Multi-GPU plumbing for the two ways the hot path shards.  One process per GPU;
torch.distributed (NCCL on GPUs, gloo in the CPU tests) carries only tiny payloads: 128-byte partial points,
1-byte verdicts, timing scalars.  There is no data-path collective inside a proof."""
import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """contiguous slice [lo, hi) of n items owned by `rank` (sizes differ by at most one)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_items(items, rank, world):
    """round-robin ownership of independent proofs / verifications: item k -> rank k mod world"""
    return [(k, it) for k, it in enumerate(items) if k % world == rank]


def allgather_bytes(payload, device=None):
    """all-gather equal-length byte strings; returns the list ordered by rank"""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return [bytes(payload)]
    t = torch.frombuffer(bytearray(payload), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    return [bytes(o.cpu().numpy().tobytes()) for o in outs]


def msm_gens_sharded(ctx, d_sG, d_sH, n, device=None):
    """One MSM  sum sG[i] G[i] + sH[i] H[i]  split by point range over the ranks of the default process group.
    d_sG / d_sH are this rank's device pointers to its OWN slice (shard_range(n, rank, world)) of the scalar vectors.
    Every rank returns the same 32-byte compressed result."""
    import ctypes as C
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_range(n, rank, world)
    if world > 1 and id(ctx) in _COMM_CTX:
        # the library's own communicator: partial MSM, all-gather, sum and compression are all enqueued on the context's stream
        out = C.create_string_buffer(32)
        ctx.check(ctx.lib.bpg_msm_gens_sharded_dev(ctx.h, d_sG, d_sH, hi - lo, lo, out))
        return out.raw
    if world > 1 and device is not None and str(device).startswith("cuda"):
        # device-resident exchange: the 128-byte partial never leaves HBM; one NCCL all-gather, one 32-byte read-back
        import ctypes as C
        bufs = _gather_buffers(world, device)
        ctx.msm_gens_partial_to_dev(d_sG, d_sH, hi - lo, lo, C.c_void_p(bufs[0].data_ptr()))  # complete on return
        dist.all_gather_into_tensor(bufs[1], bufs[0])
        torch.cuda.current_stream(bufs[1].device).synchronize()
        return ctx.points_sum_compress_dev(C.c_void_p(bufs[1].data_ptr()), world)
    part = ctx.msm_gens_partial_dev(d_sG, d_sH, hi - lo, lo)
    parts = allgather_bytes(part, device)
    return ctx.points_sum_compress(b"".join(parts))


_GATHER = {}


def _gather_buffers(world, device):
    key = (world, str(device))
    if key not in _GATHER:
        _GATHER[key] = (torch.zeros(128, dtype=torch.uint8, device=device), torch.zeros(128 * world, dtype=torch.uint8, device=device))
    return _GATHER[key]


def gather_verdicts(local, n_total, device=None):
    """local: list of (index, bool) owned by this rank -> full list of verdicts on every rank"""
    world = dist.get_world_size() if dist.is_initialized() else 1
    buf = bytearray(n_total)
    for k, v in local:
        buf[k] = 2 if v else 1
    if world == 1:
        return [b == 2 for b in buf]
    outs = allgather_bytes(bytes(buf), device)
    res = []
    for k in range(n_total):
        vals = {o[k] for o in outs} - {0}
        assert len(vals) == 1, "item %d owned by %d ranks" % (k, len(vals))
        res.append(vals.pop() == 2)
    return res


# ---------------------------------------------------------------- in-library NCCL exchange (include/bpg.h: bpg_comm_init)
_COMM_CTX = set()


def enable_comm(ctx, device=None):
    """Give `ctx` its own NCCL communicator over the ranks of the default process group (collective call).  From then on
    bpg_r1cs_prove on this context is ONE proof split over the ranks and msm_gens_sharded() is one stream-ordered call: the
    all-gather of the partial points is enqueued by the library on the context's stream (no host synchronisation per exchange).
    torch.distributed only carries the 128-byte NCCL unique id."""
    import ctypes as C
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return False
    uid = C.create_string_buffer(128)
    if rank == 0:
        ctx.check(ctx.lib.bpg_comm_unique_id(uid))
    uid_bytes = allgather_bytes(uid.raw, device)[0]
    ctx.check(ctx.lib.bpg_comm_init(ctx.h, rank, world, uid_bytes))
    _COMM_CTX.add(id(ctx))
    return True


def disable_comm(ctx):
    ctx.check(ctx.lib.bpg_comm_destroy(ctx.h))
    _COMM_CTX.discard(id(ctx))


# ---------------------------------------------------------------- one proof over several ranks (BASELINE configs[3])
_SHARD_CAP = 2 << 20  # bytes of partial points per exchange: the late fold sends 2 x (N / 128) outputs x 128 B = 2 MiB at N = 2^20


def enable_sharded_prover(ctx, device):
    """Split every large MSM of bpg_r1cs_prove on `ctx` by point range over the ranks of the default process group
    (include/bpg.h: bpg_ctx_set_shard).  The partial points are exchanged by ONE NCCL all-gather per MSM over torch-owned
    device buffers; every rank must then call prove() with the same arguments and gets the same proof bytes.
    Returns a handle that must stay alive as long as the context proves (it owns the buffers and the callback)."""
    import ctypes as C
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        ctx.check(ctx.lib.bpg_ctx_set_shard(ctx.h, 0, 1, None, None, 0, None, None))
        return None
    send = torch.zeros(_SHARD_CAP, dtype=torch.uint8, device=device)
    recv = torch.zeros(_SHARD_CAP * world, dtype=torch.uint8, device=device)
    stream = torch.cuda.current_stream(send.device)

    def _allgather(_user, nbytes):
        try:
            dist.all_gather_into_tensor(recv[:nbytes * world], send[:nbytes])
            stream.synchronize()
            return 0
        except Exception:  # never let an exception cross the C boundary
            return 1

    cb = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_size_t)(_allgather)
    ctx.check(ctx.lib.bpg_ctx_set_shard(ctx.h, rank, world, C.c_void_p(send.data_ptr()), C.c_void_p(recv.data_ptr()), _SHARD_CAP,
                                        C.cast(cb, C.c_void_p), None))
    return (send, recv, cb)


def disable_sharded_prover(ctx):
    """back to an unsharded prover on this context"""
    ctx.check(ctx.lib.bpg_ctx_set_shard(ctx.h, 0, 1, None, None, 0, None, None))
