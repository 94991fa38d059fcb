"""Small synthetic R1CS circuits shared by the oracle tests and the GPU parity tests.

Constraint terms use the flat CSR form of include/bpg.h / oracle/bpo.h:
  term_var = kind << 29 | index   (kind: 0=L 1=R 2=O 3=V 4=One),  term_coeff = 32-byte LE scalars.
"""
import random

from oracle import pyref as pr

L = pr.L
KIND = {"L": 0, "R": 1, "O": 2, "V": 3, "1": 4}


def to_csr(constraints):
    row_ptr, tv, tc = [0], [], bytearray()
    for lc in constraints:
        for (k, i), c in lc:
            tv.append((KIND[k] << 29) | i)
            tc += pr.sc_bytes(c % L)
        row_ptr.append(len(tv))
    return row_ptr, tv, bytes(tc)


def chain_wire(cs, Vvars, seed, nmul):
    """x0 = V0 ; x_{k+1} = (x_k + c_k)(x_k + V1) ; constrain x_n - V2 = 0 (same shape as a MiMC-like chain)."""
    rnd = random.Random(seed)
    cur = [(Vvars[0], 1)]
    for _ in range(nmul):
        c = rnd.randrange(L)
        _, _, o = cs.multiply(cur + [(("1", 0), c)], cur + [(Vvars[1], 1)])
        cur = [(o, 1)]
    cs.constrain(cur + [(Vvars[2], L - 1)])


def chain_instance(nmul, seed, wrong=False):
    """-> dict(label, vals, blinds, aL, aR, aO (bytes), csr)"""
    rnd = random.Random(seed)
    v0, v1 = rnd.randrange(L), rnd.randrange(L)
    r2 = random.Random(seed + 1000)
    x = v0
    for _ in range(nmul):
        x = (x + r2.randrange(L)) * (x + v1) % L
    if wrong:
        x = (x + 1) % L
    vals = [v0, v1, x]
    blinds = [rnd.randrange(L) for _ in vals]
    t = pr.Transcript(b"chain")
    cs = pr.ConstraintSystem(t, True)
    cs.commit = lambda v, b, _cs=cs: (None, _lazy_commit(_cs, v, b))
    Vv = [cs.commit(v, b)[1] for v, b in zip(vals, blinds)]
    chain_wire(cs, Vv, seed + 1000, nmul)
    enc = lambda xs: b"".join(pr.sc_bytes(a) for a in xs)
    return dict(label=b"chain", n=nmul, vals=enc(vals), blinds=enc(blinds), aL=enc(cs.aL), aR=enc(cs.aR), aO=enc(cs.aO),
                csr=to_csr(cs.constraints), ivals=vals, iblinds=blinds, seed=seed)


def _lazy_commit(cs, v, b):
    """record the opening without the (slow, big-int) group operation."""
    cs.v.append(v)
    cs.v_blinding.append(b)
    return ("V", len(cs.v) - 1)


def random_dense_instance(n, seed, m=2):
    """n allocate_multiplier gates with random a_L,a_R and 2n sparse random constraints that the witness
    satisfies by construction (each constraint: c1*L_i + c2*R_j + c3*O_k - rhs*One = 0 is NOT satisfiable in
    general, so instead tie through committed variables): used for throughput, not soundness."""
    rnd = random.Random(seed)
    aL = [rnd.randrange(L) for _ in range(n)]
    aR = [rnd.randrange(L) for _ in range(n)]
    aO = [a * b % L for a, b in zip(aL, aR)]
    rows = []
    # constraint i: L_i + c*R_{i'} + d*O_{i''} - V_{i%m}*e ... made satisfiable by a constant term
    vals = [rnd.randrange(L) for _ in range(m)]
    for i in range(n):
        j, k = rnd.randrange(n), rnd.randrange(n)
        c, d, e = rnd.randrange(L), rnd.randrange(L), rnd.randrange(L)
        s = (aL[i] + c * aR[j] + d * aO[k] + e * vals[i % m]) % L
        rows.append([(("L", i), 1), (("R", j), c), (("O", k), d), (("V", i % m), e), (("1", 0), (-s) % L)])
    blinds = [rnd.randrange(L) for _ in range(m)]
    enc = lambda xs: b"".join(pr.sc_bytes(a) for a in xs)
    return dict(label=b"dense", n=n, vals=enc(vals), blinds=enc(blinds), aL=enc(aL), aR=enc(aR), aO=enc(aO),
                csr=to_csr(rows), ivals=vals, iblinds=blinds, seed=seed)


def host_witness(p):
    """(a_L, a_R, a_O) of an api.Prover evaluated here with big integers: Prover::eval of every multiply() in creation order
    (cs_buffer.rs:94-97, prover.rs:102-117).  Test-side checker for the device evaluation (bpg_witness_eval) and the witness
    source of the CPU-only tests."""
    n = p.num_vars
    aL, aR, aO = list(p._in_L), list(p._in_R), [0] * n
    vals = {0: aL, 1: aR, 2: aO, 3: p.v}
    tc = bytes(p._w_coeff)

    def ev(t0, t1):
        acc = 0
        for t in range(t0, t1):
            kind, idx = p._w_var[t] >> 29, p._w_var[t] & 0x1FFFFFFF
            acc += int.from_bytes(tc[32 * t:32 * t + 32], "little") * (1 if kind == 4 else vals[kind][idx])
        return acc % L

    for i in range(n):
        t0, t1, t2 = p._w_ptr[2 * i], p._w_ptr[2 * i + 1], p._w_ptr[2 * i + 2]
        if t2 > t0:
            aL[i], aR[i] = ev(t0, t1), ev(t1, t2)
        aO[i] = aL[i] * aR[i] % L
    return aL, aR, aO
