#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native Bulletproofs R1CS hot path.

  python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K --warmup W

Metric (BASELINE.json): R1CS proofs/sec (and MSM Mpoints/sec in the `msm` sweep) on BASELINE configs[1]:
"merkle_tree membership with mimc_hash, depth 32, single proof" (n = 63 180 multipliers, N = 2^16, m = 4).
A proof = Pedersen commits, 3 commitment MSMs, polynomial phase, 16 IPP rounds of that circuit.  A step = K proofs, one
per concurrent prover of the GPU (independent host thread + bpg_ctx each; K is reported in `config`); the K * steps
proofs of the timed region are handed out to the free-running provers one at a time, which hides the sequential host-side
Merlin RNG of one proof (its stream shares SIMD lanes with the other provers' streams) behind the device work of the
others; `single_proof_latency_ms` is the un-overlapped figure.  Every rank runs its own provers (weak scaling, no data-path
collective).  Extras for N > 1: `msm_sharded` (one MSM split by point range) and `one_proof_2p20` (ONE large proof split
over the ranks, BASELINE configs[3]).

  value  proofs/s with the witness vectors already resident in HBM (BPG_FLAG_WITNESS_ON_DEVICE)
  e2e    proofs/s through the C ABI with HOST buffers: witness H2D, proof + commitments D2H inside the timed region
  roofline   the dominant kernel k_msm_accumulate (bucket accumulation of the fixed-base Pippenger): algorithmic bytes
             = 100 B per (term, window) pair (96 B affine-Niels table entry + 4 B sorted index) / CUDA-event time
  cpu_baseline   the C oracle (oracle/bpo.c, a restatement of dalek's algorithms) on the host cores, same circuit

The reference itself (Rust) cannot be built in this image; `--impl reference` times the oracle port.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# the concurrent provers use one CUDA stream each: give every stream its own hardware work queue (default: 8 shared)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
# (BPG_BLOCKING_SYNC=1 makes provers sleep instead of spin while they wait for the device; measured neutral at K <= cores)

_emit = print
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = "merkle_tree membership with mimc_hash, depth 32, single proof"
DEPTH = 32
GENS_CAP = 1 << 16


class ClockSampler(threading.Thread):
    """samples nvidia-smi clocks / throttle reasons during the timed region"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip().split(",")
                self.samples.append(float(out[0]))
                self.max_mhz = float(out[1])
                for nm, v in zip(names, out[2:]):
                    if v.strip().lower().startswith("active"):
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


def dist_setup(n_gpus):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local, dist


def barrier_max(dist, local, value):
    """barrier + max over ranks of a python float"""
    if dist is None:
        return value
    import torch
    t = torch.tensor([value], dtype=torch.float64, device="cuda:%d" % local)
    dist.barrier()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def oracle_prove_time(inst, threads, reps=1):
    import oracle_lib as ol
    ol.lib().bpo_set_threads(threads)
    rp, tv, tc = inst["csr"]
    tcb = inst.setdefault("_tc_bytes", tc.tobytes() if hasattr(tc, "tobytes") else tc)
    best, proof = None, None
    for _ in range(reps):
        t0 = time.perf_counter()
        proof, V = ol.r1cs_prove(inst["label"], GENS_CAP, inst["aL"], inst["aR"], inst["aO"], inst["vals"], inst["blinds"], rp, tv, tcb, bytes(32))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best, proof


def run_reference(args):
    """reference arm: the reference's CPU algorithm for the same path (oracle port; the Rust crate cannot be built here)"""
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from bulletproofs_gadgets_b200 import gadgets
    cores = os.cpu_count() or 1
    inst = gadgets.merkle_path_instance(DEPTH, trace_on_device=False)
    import oracle_lib as ol
    ol.gens(0, GENS_CAP)  # BulletproofGens::new outside the timed steps, as for the GPU arm
    for _ in range(args.warmup):
        oracle_prove_time(inst, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle_prove_time(inst, cores)
    dt = time.perf_counter() - t0
    v = args.steps / dt
    line = {"impl": "reference", "metric": "r1cs_proofs_per_sec", "value": v, "unit": "proofs/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (GF(2^255-19), Z_l)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_multipliers": inst["n"], "padded_n": GENS_CAP, "commitments": inst["m"]},
            "cpu_baseline": {"value": v, "unit": "proofs/s", "cores": cores, "kind": "port",
                             "sample": "full workload: %d complete proofs of the depth-32 circuit, OpenMP on all host cores" % args.steps},
            "e2e": {"value": v, "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(json.dumps(line))


def msm_sweep(ctx, sizes, reps=5, dist="uniform"):
    """MSM Mpoints/s over the resident generators with device-resident scalars (points = n/2 G + n/2 H).
    dist = "uniform": 252-bit scalars (SURVEY 8d MSM-uniform); "bits": scalars in {0, 1} (MSM-bits, the range-proof shape:
    half of the pairs vanish, the other half all land in bucket 1 of window 0)"""
    import numpy as np
    out = {}
    rng = np.random.default_rng(1 if dist == "uniform" else 2)
    maxn = max(sizes)
    if dist == "uniform":
        raw = rng.integers(0, 256, size=(maxn, 32), dtype=np.uint8)
        raw[:, 31] &= 0x0F  # < 2^252 < l : uniform reduced scalars
    else:
        raw = np.zeros((maxn, 32), dtype=np.uint8)
        raw[:, 0] = rng.integers(0, 2, size=maxn, dtype=np.uint8)
    d = ctx.dev_alloc(32 * maxn)
    ctx.dev_upload(d, raw.tobytes())
    import ctypes as C
    for n in sizes:
        h = n // 2
        dG, dH = d, C.c_void_p(d.value + 32 * h)
        ctx.msm_gens_dev(dG, dH, h, 0)  # warm-up (sizes exceed L2 only from 2^20 up; tables are re-gathered randomly)
        ctx.event_record(0)
        for _ in range(reps):
            ctx.msm_gens_dev(dG, dH, h, 0)
        ctx.event_record(1)
        ms = ctx.event_elapsed_ms(0, 1) / reps
        out[str(n)] = {"ms": ms, "mpoints_per_s": n / ms / 1e3}
    ctx.dev_free(d)
    return out


def msm_var_sweep(ctx, sizes, reps=3):
    """SURVEY 8d MSM-var: variable-base MSM through the host-buffer entry point bpg_msm (n compressed points + n scalars
    uploaded, decompressed, one windowed scalar multiplication per term, tree sum): the path of the verifier's own points"""
    import numpy as np
    out = {}
    rng = np.random.default_rng(3)
    maxn = max(sizes)
    G, H = ctx.gens_export(0, min(maxn, ctx.gens_capacity()))
    pts = (G + H) * (1 + maxn * 32 // max(1, len(G + H)))
    raw = rng.integers(0, 256, size=(maxn, 32), dtype=np.uint8)
    raw[:, 31] &= 0x0F
    sc = raw.tobytes()
    for n in sizes:
        ctx.msm(sc[:32 * n], pts[:32 * n])
        t0 = time.perf_counter()
        for _ in range(reps):
            ctx.msm(sc[:32 * n], pts[:32 * n])
        ms = (time.perf_counter() - t0) * 1e3 / reps
        out[str(n)] = {"ms": ms, "mpoints_per_s": n / ms / 1e3}
    return out


def msm_sharded_sweep(ctx, dist, local, rank, world, sizes, reps=5):
    """ONE MSM of n points split by point range over the ranks (parallel.msm_gens_sharded): every rank
    sums its slice of the resident generators, the 128-byte partial points are all-gathered over NCCL and added on every rank.
    Timed on the host around barrier + synchronize (the collective runs on torch's stream), max over ranks."""
    import ctypes as C
    import numpy as np
    from bulletproofs_gadgets_b200 import parallel
    out = {}
    maxn = max(sizes)
    ctx.gens_ensure(maxn // 2)
    rng = np.random.default_rng(7)
    dev = "cuda:%d" % local
    for n in sizes:
        h = n // 2
        lo, hi = parallel.shard_range(h, rank, world)
        raw = rng.integers(0, 256, size=(2 * h, 32), dtype=np.uint8)  # same stream on every rank: sG | sH
        raw[:, 31] &= 0x0F
        mine = np.concatenate([raw[lo:hi], raw[h + lo:h + hi]])
        d = ctx.dev_alloc(32 * len(mine))
        ctx.dev_upload(d, mine.tobytes())
        dG, dH = d, C.c_void_p(d.value + 32 * (hi - lo))
        first = parallel.msm_gens_sharded(ctx, dG, dH, h, dev)  # warm-up, and a cross-rank agreement check
        agree = parallel.allgather_bytes(first, dev)
        if any(a != first for a in agree):
            raise SystemExit("sharded MSM: ranks disagree on the result")
        barrier_max(dist, local, 0.0)
        t0 = time.perf_counter()
        for _ in range(reps):
            parallel.msm_gens_sharded(ctx, dG, dH, h, dev)
        ms = barrier_max(dist, local, (time.perf_counter() - t0) * 1e3 / reps)
        out[str(n)] = {"ms": ms, "mpoints_per_s": n / ms / 1e3, "points_per_rank": 2 * (hi - lo)}
        ctx.dev_free(d)
    return out


def one_large_proof(bpg, gadgets, ctx, dist, local, world):
    """BASELINE configs[3]: ONE proof of a 993 384-multiplier circuit (N = 2^20, 20 IPP rounds) on `world` GPUs -- strong scaling.
    With world > 1 every MSM of the proof is cut by point range over the ranks (bpg_ctx_set_shard, parallel.enable_sharded_prover)
    and the partial points are all-gathered by NCCL; all ranks return the same bytes (checked).  Device-side blinding, because the
    2n sequential transcript-RNG draws of the byte-exact mode (0.7 s of one host core) are the same on every rank."""
    from bulletproofs_gadgets_b200 import parallel
    inst = gadgets.mimc_chain_instance(1022, ctx=ctx)
    ctx.gens_ensure(1 << 20)
    circ = gadgets.Circuit(ctx, inst["n"], inst["m"], inst["csr"])
    keep = parallel.enable_sharded_prover(ctx, "cuda:%d" % local) if world > 1 else None
    FAST = bpg._lib.FLAG_FAST_BLINDING
    ext = b"\x44" * 32
    proof, V = circ.prove(inst, ext, FAST)  # warm-up (buffers, late-fold tables)
    if world > 1:
        same = parallel.allgather_bytes(proof, "cuda:%d" % local)
        if any(p != proof for p in same):
            raise SystemExit("sharded prover: ranks returned different proof bytes")
    if not circ.verify(inst["label"], V, proof):
        raise SystemExit("sharded prover: the verifier rejected the proof")
    barrier_max(dist, local, 0.0)
    reps = 3
    t0 = time.perf_counter()
    for _ in range(reps):
        circ.prove(inst, ext, FAST)
    ms = barrier_max(dist, local, (time.perf_counter() - t0) * 1e3 / reps)
    if world > 1:
        ctx.check(ctx.lib.bpg_ctx_set_shard(ctx.h, 0, 1, None, None, 0, None, None))
    del keep
    circ.close()
    return {"n_multipliers": inst["n"], "padded_n": 1 << 20, "gpus": world, "prove_ms_fast_blinding": ms, "proofs_per_sec": 1e3 / ms,
            "mode": "MSMs split by point range over the ranks, NCCL all-gather of the partial points" if world > 1 else "one GPU"}


class ProverLane:
    """one host thread's private context: own bpg_ctx (stream, workspace, tables), circuit copy and HBM-resident witness"""

    def __init__(self, bpg, gadgets, device, inst):
        import ctypes as C
        self.ctx = bpg.Context(device)
        self.ctx.gens_ensure(GENS_CAP)
        self.inst = inst
        self.circ = gadgets.Circuit(self.ctx, inst["n"], inst["m"], inst["csr"])
        n = inst["n"]
        self.d_w = self.ctx.dev_alloc(3 * 32 * n)
        self.ctx.dev_upload(self.d_w, inst["aL"] + inst["aR"] + inst["aO"])
        self.dev_inst = dict(inst)
        self.dev_inst["aL"], self.dev_inst["aR"], self.dev_inst["aO"] = (C.cast(C.c_void_p(self.d_w.value + 32 * n * k), C.c_char_p) for k in range(3))

    def prove(self, ext, flags, resident):
        return self.circ.prove(self.dev_inst if resident else self.inst, ext, flags)

    def close(self):
        self.circ.close()
        self.ctx.dev_free(self.d_w)
        self.ctx.close()


def run_ours(args):
    from concurrent.futures import ThreadPoolExecutor
    world, rank, local, dist = dist_setup(args.gpus)
    import bulletproofs_gadgets_b200 as bpg
    from bulletproofs_gadgets_b200 import gadgets
    cores = os.cpu_count() or 1
    per_gpu = cores / max(world, 1)
    if args.provers > 0:
        K = args.provers
    else:
        # The provers' bulk transcript-RNG draws are batched into SIMD lanes by the library (host_rng_service.h) and the waiting
        # threads sleep, so the prover count is set by latency hiding, not by the core count: a proof spends ~60-100 ms in the
        # shared RNG lanes and ~12 ms on the device, and ~300 proofs/s need ~30+ proofs in flight.
        # Measured on one B200 + 16 cores: K = 24 / 48 / 64 / 96 / 128 -> 249 / 272 / 312 / 328 / 329 proofs/s byte-exact
        # (fast blinding: 327-332, i.e. 96 provers close the gap).  With fewer than 8 cores per GPU the host is the limit and
        # more threads do not help (8 GPUs on 32 cores were measured with 64).
        K = 96 if per_gpu >= 8 else 64
    # more prover threads than cores: they sleep while they wait for the device (bpg_set_blocking_sync); the solo latency
    # measurements further down (single proof, MSM sweeps) switch back to spinning
    blocking = K * world > cores
    ctx0 = bpg.Context(local)
    ctx0.lib.bpg_set_blocking_sync(1 if blocking else 0)
    inst = gadgets.merkle_path_instance(DEPTH, seed=4 + rank, ctx=ctx0)
    lanes = [ProverLane(bpg, gadgets, local, inst) for _ in range(K)]
    ctx, circ, n = lanes[0].ctx, lanes[0].circ, inst["n"]
    ext = bytes([rank + 1]) * 32
    RES = bpg._lib.FLAG_WITNESS_ON_DEVICE

    proof, V = circ.prove(inst, ext)
    if not circ.verify(inst["label"], V, proof):
        raise SystemExit("self-check failed: the verifier rejected the benchmark proof")
    pool = ThreadPoolExecutor(max_workers=K)

    def step_batch(flags, resident):
        """one step = K independent proofs, one per host thread / context, all on this rank's GPU"""
        outs = list(pool.map(lambda ln: ln.prove(ext, flags, resident), lanes))
        return outs

    tickets = {"next": 0, "total": 0, "lock": threading.Lock()}

    def lane_run(ln, flags, resident, steps):
        # staggered start (inside the timed region): lane i begins i ms late so that the lanes' host-RNG and device phases
        # interleave from the first proof on instead of all lanes hitting the CPU, then the GPU, in lockstep.
        # The K * steps proofs of the region are handed out one at a time, so every lane stays busy until the last proofs
        # are taken and the region does not end on a few straggling lanes.
        time.sleep(0.001 * lanes.index(ln))
        while True:
            with tickets["lock"]:
                if tickets["next"] >= tickets["total"]:
                    return
                tickets["next"] += 1
            ln.prove(ext, flags, resident)

    host_cpu_ms = [0.0]

    def timed(flags, resident, steps):
        """K * steps proofs (`steps` per prover on average), the provers free-running with no barrier between steps, so one
        lane's host-side transcript RNG overlaps the other lanes' device work.  The region is bracketed by a sync of every
        context on both sides."""
        for ln in lanes:
            ln.ctx.sync()
        l0 = sum(ln.ctx.launch_count() for ln in lanes)
        tickets["next"], tickets["total"] = 0, K * steps
        ctx.event_record(2)
        c0 = os.times()
        t0 = time.perf_counter()
        list(pool.map(lambda ln: lane_run(ln, flags, resident, steps), lanes))
        for ln in lanes:
            ln.ctx.sync()
        ctx.event_record(3)
        ms_dev = ctx.event_elapsed_ms(2, 3)
        wall = (time.perf_counter() - t0) * 1e3
        c1 = os.times()
        host_cpu_ms[0] = 1e3 * ((c1.user - c0.user) + (c1.system - c0.system)) / (K * steps)  # this process: all prover threads
        return max(ms_dev, wall), sum(ln.ctx.launch_count() for ln in lanes) - l0

    # proofs from every lane are byte-identical (same transcript, same randomness): a cheap cross-context check
    outs = step_batch(0, False)
    if any(o != (proof, V) for o in outs):
        raise SystemExit("concurrent contexts produced different proof bytes")
    # untimed warm-up in exactly the shape of the timed region (free-running lanes).  At least 10 steps: the 16 host threads
    # need ~1 s of load before the OS has spread them over the cores (measured: the first second runs 35 % slower)
    warm_steps = max(args.warmup, 20)  # (10 steps = 2 s left the first timed region 10-25 % low in 2 of 8 runs)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()  # started before the warm-up: the first nvidia-smi invocations (cold NVML start) are slow and disturb the run
    timed(RES, True, warm_steps)
    sampler.samples.clear()
    ctx.prof_enable(True)
    barrier_max(dist, local, 0.0)
    ms_value, launches = timed(RES, True, args.steps)
    ms_value = barrier_max(dist, local, ms_value)
    cpu_ms_value = host_cpu_ms[0]
    nl_c, kms_c, pairs_c = ctx.prof_read()   # lane 0's launches inside the timed region: they share the GPU with the other lanes
    ctx.prof_enable(False)
    barrier_max(dist, local, 0.0)
    ms_e2e, _ = timed(0, False, args.steps)
    ms_e2e = barrier_max(dist, local, ms_e2e)
    barrier_max(dist, local, 0.0)
    ms_fast, _ = timed(RES | bpg._lib.FLAG_FAST_BLINDING, True, args.steps)
    ms_fast = barrier_max(dist, local, ms_fast)
    sampler.stop_flag = True
    if rank == 0:
        sampler.join(timeout=10)
    ctx0.lib.bpg_set_blocking_sync(0)
    # the same kernel timed alone (one prover, nothing else on the GPU): this is the figure the roofline fraction is quoted on
    barrier_max(dist, local, 0.0)
    ctx.prof_enable(True)
    ctx.event_record(4)
    for _ in range(3):
        lanes[0].prove(ext, RES | bpg._lib.FLAG_FAST_BLINDING, True)  # device-side blinding: no 45 ms host RNG gap inside the events
    ctx.event_record(5)
    solo_ms = ctx.event_elapsed_ms(4, 5)
    # A proof launches the kernel at two very different sizes: the full-size MSMs over the resident generators (commitments,
    # first IPP rounds, the late-fold materialisation: >= 1 M pairs each, > 95 % of the kernel's time) and the tiny ones over
    # the 2 x 512 materialised generators (late IPP rounds, ~16 K pairs, launch-latency bound).  The roofline is quoted on the
    # full-size launches; the small ones are listed beside it.
    per_launch = ctx.prof_read_launches()
    ctx.prof_enable(False)  # never leave a daemon thread (mid nvidia-smi call) running into interpreter shutdown
    big_cut = 0.25 * max([p for _, p in per_launch] or [0])
    big = [(m, p) for m, p in per_launch if p >= big_cut]
    small = [(m, p) for m, p in per_launch if p < big_cut]
    nl, kms, pairs = len(big), sum(m for m, _ in big), sum(p for _, p in big)
    small_launches = {"launches": len(small), "avg_launch_ms": (sum(m for m, _ in small) / len(small)) if small else None,
                      "pairs_per_launch": (sum(p for _, p in small) / len(small)) if small else None,
                      "share_of_kernel_time": (sum(m for m, _ in small) / max(sum(m for m, _ in per_launch), 1e-12)) if per_launch else None}
    nproofs = world * K * args.steps

    extras = {}
    if rank == 0 and not args.no_extras:
        # single-proof latency (one context, nothing else on the GPU), verification, fast-blinding, MSM sweep, MiMC, integer pipe
        t0 = time.perf_counter()
        for _ in range(args.steps):
            lanes[0].prove(ext, RES, True)
        extras["single_proof_latency_ms"] = 1e3 * (time.perf_counter() - t0) / args.steps
        ctx0.lib.bpg_set_blocking_sync(1 if blocking else 0)
        list(pool.map(lambda ln: ln.circ.verify(inst["label"], V, proof), lanes))  # warm-up
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(pool.map(lambda ln: ln.circ.verify(inst["label"], V, proof), lanes))
        extras["verify_per_sec"] = K * args.steps / (time.perf_counter() - t0)
        ctx0.lib.bpg_set_blocking_sync(0)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            lanes[0].prove(ext, RES | bpg._lib.FLAG_FAST_BLINDING, True)
        extras["single_proof_latency_ms_fast_blinding"] = 1e3 * (time.perf_counter() - t0) / args.steps
        extras["single_warp_latency_cycles"] = ctx.bench_latency(200)
        nh = 1 << 14
        leaves = [[os.urandom(32), os.urandom(32)] for _ in range(nh)]
        ctx.mimc_sponge_batch(leaves[:64])
        t0 = time.perf_counter()
        ctx.mimc_sponge_batch(leaves)
        extras["mimc_merkle_nodes_per_sec"] = nh / (time.perf_counter() - t0)
        sizes = [1 << k for k in range(16, 17 + 1)] if args.quick else [1 << k for k in range(16, 22 + 1)]
        if not args.quick:
            ctx.gens_ensure(1 << 21)
        extras["msm"] = msm_sweep(ctx, sizes)
        extras["msm_bits"] = msm_sweep(ctx, sizes, dist="bits")
        extras["msm_var"] = msm_var_sweep(ctx, [1 << 10, 1 << 12] if args.quick else [1 << 10, 1 << 12, 1 << 14, 1 << 16])

    if world > 1 and not args.no_extras:
        extras["msm_sharded"] = msm_sharded_sweep(ctx, dist, local, rank, world, [1 << 20] if args.quick else [1 << 20, 1 << 22])

    if not args.no_extras and not args.quick:
        extras["one_proof_2p20"] = one_large_proof(bpg, gadgets, ctx, dist, local, world)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        t_cpu, proof_cpu = oracle_prove_time(dict(inst), cores)
        proof_same, _ = circ.prove(inst, bytes(32))  # same transcript + randomness as the oracle run
        cpu = {"value": 1.0 / t_cpu, "unit": "proofs/s", "cores": cores, "kind": "port",
               "sample": "1 complete proof of the same depth-32 circuit on all host cores (oracle/bpo.c, OpenMP)",
               "proof_bytes_equal_gpu": proof_cpu == proof_same}
        if not args.quick:
            # what the reference binary does today: one thread (SURVEY 8d asks for both figures)
            t_one, proof_one = oracle_prove_time(dict(inst), 1)
            cpu["single_thread_value"] = 1.0 / t_one
            cpu["single_thread_bytes_equal"] = proof_one == proof_cpu

    if rank == 0:
        peaks = {}
        try:
            with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
                peaks = json.load(f)
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        achieved = (pairs * 100.0) / (kms * 1e-3) / 1e9 if kms > 0 else None
        # DRAM traffic of the same kernel from the committed ncu --set full capture, scaled to this run's pairs per launch
        traffic = None
        try:
            with open(os.path.join(ROOT, "profiles", "r01_accumulate_traffic.json")) as f:
                tj = json.load(f)
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) / tj["pairs_in_that_launch"] * (pairs / nl if nl else 0)
        except Exception:
            pass
        # integer view of the same kernel: one mixed addition = 7 field multiplications = 504 32x32->64 multiply-accumulates;
        # peak = the MAC32 rate a dependent fe_mul chain sustains at full occupancy on this GPU (bpg_bench_imad, measured below)
        ms_i, mac = ctx.bench_imad(400)
        imad_peak = mac / ms_i * 1e3
        imad_ach = pairs * 504.0 / (kms * 1e-3) if kms > 0 else None
        line = {"metric": "r1cs_proofs_per_sec", "value": nproofs / (ms_value * 1e-3), "unit": "proofs/s", "n_gpus": world,
                "steps": args.steps, "warmup": warm_steps, "ms_per_step": ms_value / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u32 limbs (GF(2^255-19), Z_l)", "data": "synthetic",
                "config": {"workload": WORKLOAD, "n_multipliers": n, "padded_n": GENS_CAP, "commitments": inst["m"],
                           "constraints": int(len(inst["csr"][0]) - 1), "proofs_per_step_per_gpu": K,
                           "concurrency": "%d independent provers per GPU (one host thread + one bpg_ctx each); a step is one proof per prover" % K,
                           "l2": "window tables (201 MB at 2^16 capacity, one set per GPU shared by all provers) exceed the 126 MB L2", "byte_exact": True},
                "e2e": {"value": nproofs / (ms_e2e * 1e-3), "unit": "proofs/s", "h2d_bytes_per_step": K * (3 * 32 * n + 2 * 32 * inst["m"] + 64 * n),
                        "d2h_bytes_per_step": K * (len(proof) + 32 * inst["m"])},
                "gpu_launches": int(launches),
                "proofs_per_sec_fast_blinding": nproofs / (ms_fast * 1e-3),
                "host_cores": cores, "host_cpu_ms_per_proof": cpu_ms_value,
                "roofline": {"bound": "hbm", "kernel": "k_msm_accumulate", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                             "frac": (achieved / hbm_peak) if achieved else None, "traffic": traffic, "launches": int(nl),
                             "algorithmic_bytes_per_launch": (pairs / nl * 100.0) if nl else None,
                             "avg_launch_ms": kms / nl if nl else None, "pairs_per_launch": pairs / nl if nl else None,
                             "note": "CUDA events around every full-size launch (>= 25 % of the largest pair count) of 3 proofs run alone after the timed region (kernel timed alone)",
                             "small_launches": small_launches,
                             "share_of_step": (sum(m for m, _ in per_launch) / solo_ms) if solo_ms > 0 else None,
                             "share_note": "all k_msm_accumulate launches / device time of the same 3 solo proofs (compare profiles/r01_launch_summary_final.txt)",
                             "in_timed_region": {"launches": int(nl_c), "avg_launch_ms": kms_c / nl_c if nl_c else None,
                                                 "note": "lane 0's launches while %d other provers share the GPU" % (K - 1)},
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
                "imad_roofline": {"kernel": "k_msm_accumulate", "achieved": imad_ach, "peak": imad_peak, "unit": "MAC32/s",
                                  "frac": (imad_ach / imad_peak) if imad_ach else None,
                                  "note": "the kernel is integer-multiply bound before it is HBM bound; peak = measured dependent fe_mul chain"},
                "cpu_baseline": cpu, "clocks": sampler.summary()}
        line.update(extras)
        _emit(json.dumps(line))
    pool.shutdown()
    for ln in lanes:
        ln.close()
    ctx0.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    # stdout carries exactly ONE JSON line: everything else that libraries print there (e.g. NCCL's version banner) goes to stderr
    saved = os.dup(1)
    os.dup2(2, 1)
    global _emit
    def _emit(text):
        os.write(saved, (text + "\n").encode())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--provers", type=int, default=0, help="concurrent provers (host threads / contexts) per GPU; 0 = auto")
    ap.add_argument("--quick", action="store_true", help="small MSM sweep only")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
