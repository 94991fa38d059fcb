"""oracle/pyref.py -- big-int CPU restatement of the hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this
module; the product (bulletproofs_gadgets_b200/) never does.

What it restates (the reference keeps this arithmetic in un-vendored crates, see
SURVEY.md section 8c: curve25519-dalek 1.x, merlin 1.x, bulletproofs fork `develop`;
/root/reference/Cargo.toml:8,10,17-20) and where the reference calls it:

  Scalar semantics        conversions.rs:18,43 (from_bits), utils.rs:12-18
  ristretto255            RFC 9496 == dalek RistrettoPoint/CompressedRistretto
  PedersenGens/commit     gadget.rs:31, commitments.rs:27,39, cs_buffer.rs:39
  BulletproofGens::new    src/bin/prover.rs:92, src/bin/verifier.rs:89
  Merlin transcript       src/bin/prover.rs:52, src/bin/verifier.rs:51
  Prover::prove           src/bin/prover.rs:93     
  InnerProductProof       inside prove/verify      
  Verifier::verify        src/bin/verifier.rs:90   
  mimc_hash / sponge      src/mimc_hash/mimc.rs:7-97, merkle_tree_gadget.rs:75-107

Parity status: MiMC/Merkle are pinned bit-exactly by the reference's own known answers
(tests/golden/mimc_kats.json); ristretto255 by RFC 9496 vectors; Merlin by its published
test vector; group law by libsodium (PyNaCl).  MSM / Pedersen / IPP / proof bytes have NO
golden bytes in the reference => "parity unpinned" at that boundary (accept/reject only).

Pure Python loops: use only for small cases (n <= ~2^10).  The C oracle (oracle/bpo.c)
is the fast checker and is itself checked against this file.
"""
import hashlib
import struct

P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493


def inv(x):
    return pow(x, P - 2, P)


D = (-121665 * inv(121666)) % P
D2 = (2 * D) % P
SQRT_M1 = pow(2, (P - 1) // 4, P)


def is_neg(x):
    return (x % P) & 1


def fabs(x):
    x %= P
    return P - x if x & 1 else x


def sqrt_ratio_m1(u, v):
    """RFC 9496 4.2 SQRT_RATIO_M1 -> (was_square, r)."""
    u %= P
    v %= P
    v3 = v * v % P * v % P
    v7 = v3 * v3 % P * v % P
    r = u * v3 % P * pow(u * v7 % P, (P - 5) // 8, P) % P
    check = v * r % P * r % P
    correct = check == u
    flipped = check == (-u) % P
    flipped_i = check == (-u * SQRT_M1) % P
    if flipped or flipped_i:
        r = r * SQRT_M1 % P
    r = fabs(r)
    return (correct or flipped), r


INVSQRT_A_MINUS_D = sqrt_ratio_m1(1, (-1 - D) % P)[1]
ONE_MINUS_D_SQ = (1 - D * D) % P
D_MINUS_ONE_SQ = (D - 1) * (D - 1) % P
# sqrt(a*d - 1) with a = -1: the ODD root
_s = sqrt_ratio_m1((-D - 1) % P, 1)[1]
SQRT_AD_MINUS_ONE = _s if _s & 1 else P - _s
assert SQRT_AD_MINUS_ONE * SQRT_AD_MINUS_ONE % P == (-D - 1) % P

# ----------------------------------------------------------------------------- Edwards
IDENT = (0, 1, 1, 0)


def pt_add(p, q):
    """add-2008-hwcd-3, complete for a=-1."""
    X1, Y1, Z1, T1 = p
    X2, Y2, Z2, T2 = q
    A = (Y1 - X1) * (Y2 - X2) % P
    B = (Y1 + X1) * (Y2 + X2) % P
    C = D2 * T1 % P * T2 % P
    Dd = 2 * Z1 * Z2 % P
    E, F, G, H = B - A, Dd - C, Dd + C, B + A
    return (E * F % P, G * H % P, F * G % P, E * H % P)


def pt_neg(p):
    X, Y, Z, T = p
    return ((-X) % P, Y, Z, (-T) % P)


def pt_dbl(p):
    return pt_add(p, p)


def pt_mul(k, p):
    r = IDENT
    q = p
    while k:
        if k & 1:
            r = pt_add(r, q)
        q = pt_dbl(q)
        k >>= 1
    return r


def pt_eq(p, q):
    """ristretto coset equality (dalek ct_eq)."""
    X1, Y1, _, _ = p
    X2, Y2, _, _ = q
    return (X1 * Y2 - Y1 * X2) % P == 0 or (X1 * X2 - Y1 * Y2) % P == 0


def ristretto_decode(b):
    """RFC 9496 4.3.1; returns None on failure."""
    s = int.from_bytes(b, "little")
    if s >= P or s & 1 or len(b) != 32:
        return None
    ss = s * s % P
    u1 = (1 - ss) % P
    u2 = (1 + ss) % P
    u2s = u2 * u2 % P
    v = (-(D * u1 % P * u1) - u2s) % P
    ok, I = sqrt_ratio_m1(1, v * u2s % P)
    Dx = I * u2 % P
    Dy = I * Dx % P * v % P
    x = fabs(2 * s * Dx % P)
    y = u1 * Dy % P
    t = x * y % P
    if (not ok) or is_neg(t) or y == 0:
        return None
    return (x, y, 1, t)


def ristretto_encode(p):
    """RFC 9496 4.3.2."""
    X, Y, Z, T = p
    u1 = (Z + Y) * (Z - Y) % P
    u2 = X * Y % P
    _, I = sqrt_ratio_m1(1, u1 * u2 % P * u2 % P)
    d1 = I * u1 % P
    d2 = I * u2 % P
    zinv = d1 * d2 % P * T % P
    if is_neg(T * zinv % P):
        X, Y = Y * SQRT_M1 % P, X * SQRT_M1 % P
        den = d1 * INVSQRT_A_MINUS_D % P
    else:
        den = d2
    if is_neg(X * zinv % P):
        Y = (-Y) % P
    s = fabs(den * (Z - Y) % P)
    return s.to_bytes(32, "little")


def elligator(t):
    """RFC 9496 4.3.4 MAP."""
    r = SQRT_M1 * t % P * t % P
    u = (r + 1) * ONE_MINUS_D_SQ % P
    v = (-1 - r * D) % P * ((r + D) % P) % P
    sq, s = sqrt_ratio_m1(u, v)
    s_prime = (-fabs(s * t % P)) % P
    if not sq:
        s = s_prime
        c = r
    else:
        c = P - 1
    N = (c * ((r - 1) % P) % P * D_MINUS_ONE_SQ - v) % P
    w0 = 2 * s * v % P
    w1 = N * SQRT_AD_MINUS_ONE % P
    w2 = (1 - s * s) % P
    w3 = (1 + s * s) % P
    return (w0 * w3 % P, w2 * w1 % P, w1 * w3 % P, w0 * w2 % P)


def from_uniform_bytes(b64):
    assert len(b64) == 64
    r1 = (int.from_bytes(b64[:32], "little") & ((1 << 255) - 1)) % P
    r2 = (int.from_bytes(b64[32:], "little") & ((1 << 255) - 1)) % P
    return pt_add(elligator(r1), elligator(r2))


BASEPOINT = ristretto_decode(bytes.fromhex(
    "e2f2ae0a6abc4e71a884a961c500515f58e30b6aa582dd8db6a65945e08d2d76"))
B_BLINDING = from_uniform_bytes(hashlib.sha3_512(ristretto_encode(BASEPOINT)).digest())


# ----------------------------------------------------------------------------- scalars
def sc_from_bits(b):
    """Scalar::from_bits: clear bit 255, NO reduction."""
    return int.from_bytes(b, "little") & ((1 << 255) - 1)


def sc_wide(b64):
    return int.from_bytes(b64, "little") % L


def sc_bytes(x):
    return int(x).to_bytes(32, "little")


def sc_inv(x):
    return pow(x % L, L - 2, L)


# ----------------------------------------------------------------------------- generators
def pedersen_commit(v, r):
    return pt_add(pt_mul(v % L, BASEPOINT), pt_mul(r % L, B_BLINDING))


def generators_chain(label, n, skip=0):
    """bulletproofs GeneratorsChain: SHAKE256("GeneratorsChain" || label) in 64-byte blocks."""
    stream = hashlib.shake_256(b"GeneratorsChain" + label).digest(64 * (skip + n))
    return [from_uniform_bytes(stream[64 * i:64 * i + 64]) for i in range(skip, skip + n)]


def bulletproof_gens(n, party=0):
    G = generators_chain(b"G" + struct.pack("<I", party), n)
    H = generators_chain(b"H" + struct.pack("<I", party), n)
    return G, H


def msm(scalars, points):
    acc = IDENT
    for s, p in zip(scalars, points):
        s %= L
        if s:
            acc = pt_add(acc, pt_mul(s, p))
    return acc


# ----------------------------------------------------------------------------- Keccak / STROBE / Merlin
_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000,
    0x000000000000808B, 0x0000000080000001, 0x8000000080008081, 0x8000000000008009,
    0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003,
    0x8000000000008002, 0x8000000000000080, 0x000000000000800A, 0x800000008000000A,
    0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61],
        [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def _rol(x, n):
    n %= 64
    return ((x << n) | (x >> (64 - n))) & _M64 if n else x


def keccak_f1600(state: bytearray):
    A = [[int.from_bytes(state[8 * (x + 5 * y):8 * (x + 5 * y) + 8], "little")
          for y in range(5)] for x in range(5)]
    for rnd in range(24):
        C = [A[x][0] ^ A[x][1] ^ A[x][2] ^ A[x][3] ^ A[x][4] for x in range(5)]
        Dd = [C[(x - 1) % 5] ^ _rol(C[(x + 1) % 5], 1) for x in range(5)]
        A = [[A[x][y] ^ Dd[x] for y in range(5)] for x in range(5)]
        Bm = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                Bm[y][(2 * x + 3 * y) % 5] = _rol(A[x][y], _ROT[x][y])
        A = [[Bm[x][y] ^ ((~Bm[(x + 1) % 5][y]) & Bm[(x + 2) % 5][y]) for y in range(5)]
             for x in range(5)]
        A[0][0] ^= _RC[rnd]
    for x in range(5):
        for y in range(5):
            state[8 * (x + 5 * y):8 * (x + 5 * y) + 8] = A[x][y].to_bytes(8, "little")


_F_I, _F_A, _F_C, _F_T, _F_M, _F_K = 1, 2, 4, 8, 16, 32
_STROBE_R = 166


class Strobe128:
    """merlin 1.x strobe.rs."""

    def __init__(self, label=None):
        if label is None:
            return
        st = bytearray(200)
        st[0:6] = bytes([1, _STROBE_R + 2, 1, 0, 1, 96])
        st[6:18] = b"STROBEv1.0.2"
        keccak_f1600(st)
        self.st, self.pos, self.pos_begin, self.cur_flags = st, 0, 0, 0
        self.meta_ad(label, False)

    def clone(self):
        c = Strobe128()
        c.st, c.pos, c.pos_begin, c.cur_flags = bytearray(self.st), self.pos, self.pos_begin, self.cur_flags
        return c

    def _run_f(self):
        self.st[self.pos] ^= self.pos_begin
        self.st[self.pos + 1] ^= 0x04
        self.st[_STROBE_R + 1] ^= 0x80
        keccak_f1600(self.st)
        self.pos = 0
        self.pos_begin = 0

    def _absorb(self, data):
        for b in data:
            self.st[self.pos] ^= b
            self.pos += 1
            if self.pos == _STROBE_R:
                self._run_f()

    def _overwrite(self, data):
        for b in data:
            self.st[self.pos] = b
            self.pos += 1
            if self.pos == _STROBE_R:
                self._run_f()

    def _squeeze(self, n):
        out = bytearray(n)
        for i in range(n):
            out[i] = self.st[self.pos]
            self.st[self.pos] = 0
            self.pos += 1
            if self.pos == _STROBE_R:
                self._run_f()
        return bytes(out)

    def _begin_op(self, flags, more):
        if more:
            assert self.cur_flags == flags
            return
        assert flags & _F_T == 0
        old = self.pos_begin
        self.pos_begin = self.pos + 1
        self.cur_flags = flags
        self._absorb(bytes([old, flags]))
        if flags & (_F_C | _F_K) and self.pos != 0:
            self._run_f()

    def meta_ad(self, data, more):
        self._begin_op(_F_M | _F_A, more)
        self._absorb(data)

    def ad(self, data, more):
        self._begin_op(_F_A, more)
        self._absorb(data)

    def prf(self, n, more):
        self._begin_op(_F_I | _F_A | _F_C, more)
        return self._squeeze(n)

    def key(self, data, more):
        self._begin_op(_F_A | _F_C, more)
        self._overwrite(data)


class TranscriptRng:
    def __init__(self, strobe):
        self.strobe = strobe

    def fill_bytes(self, n):
        self.strobe.meta_ad(struct.pack("<I", n), False)
        return self.strobe.prf(n, False)

    def scalar(self):
        """Scalar::random(rng) = 64 bytes, wide-reduced."""
        return sc_wide(self.fill_bytes(64))


class Transcript:
    def __init__(self, label: bytes):
        self.strobe = Strobe128(b"Merlin v1.0")
        self.append_message(b"dom-sep", label)

    def append_message(self, label, msg):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(struct.pack("<I", len(msg)), True)
        self.strobe.ad(msg, False)

    def append_u64(self, label, x):
        self.append_message(label, struct.pack("<Q", x))

    def challenge_bytes(self, label, n):
        self.strobe.meta_ad(label, False)
        self.strobe.meta_ad(struct.pack("<I", n), True)
        return self.strobe.prf(n, False)

    # bulletproofs TranscriptProtocol
    def append_scalar(self, label, s):
        self.append_message(label, sc_bytes(s))

    def append_point(self, label, enc):
        self.append_message(label, enc)

    def validate_and_append_point(self, label, enc):
        if enc == bytes(32):
            raise ValueError("identity point")
        self.append_message(label, enc)

    def challenge_scalar(self, label):
        return sc_wide(self.challenge_bytes(label, 64))

    def build_rng(self, witnesses, external32):
        """build_rng().rekey_with_witness_bytes("v_blinding", w)*.finalize(rng)
        where `external32` stands in for the 32 bytes drawn from thread_rng."""
        st = self.strobe.clone()
        for w in witnesses:
            st.meta_ad(b"v_blinding", False)
            st.meta_ad(struct.pack("<I", len(w)), True)
            st.key(w, False)
        st.meta_ad(b"rng", False)
        st.key(external32, False)
        return TranscriptRng(st)


# ----------------------------------------------------------------------------- MiMC (in-tree reference code)
_MIMC_CONSTS = None


def set_mimc_constants(consts):
    """consts: 486 ints (Scalar::from_bits of mimc_consts.rs:2-489), loaded by the caller
    from tests/golden/mimc_consts.json (the oracle holds no copy of reference sources)."""
    global _MIMC_CONSTS
    _MIMC_CONSTS = [c % L for c in consts]


def mimc_encrypt(p, k=0):
    """mimc.rs:7-23: 486 x  state = (state + k + c_i)^3 ; + k."""
    s = p % L
    for c in _MIMC_CONSTS:
        t = (s + k + c) % L
        s = t * t % L * t % L
    return (s + k) % L


def mimc_sponge(blocks):
    """mimc.rs:26-40 (zero key)."""
    s = 0
    for b in blocks:
        s = mimc_encrypt((s + b) % L)
    return s


def be_to_scalars(data: bytes):
    """conversions.rs:26-30 + 6-24: reverse, zero-pad to 32k, 32-byte LE chunks, from_bits."""
    le = bytes(reversed(data))
    if len(le) % 32:
        le += bytes(32 - len(le) % 32)
    return [sc_from_bits(le[i:i + 32]) for i in range(0, len(le), 32)]


def mimc_pad(blocks):
    """mimc.rs:77-97: PKCS#7 on the last (most-significant) block, or append 32 x 0x20."""
    last = sc_bytes(blocks[-1]).rstrip(b"\x00")
    if len(last) < 32:
        padn = 32 - len(last)
        padded = sc_from_bits(last + bytes([padn]) * padn)
        return blocks[:-1] + [padded]
    return blocks + [sc_from_bits(bytes([32]) * 32)]


def mimc_hash(data: bytes):
    """mimc.rs:61-75."""
    return mimc_sponge(mimc_pad(be_to_scalars(data)))


def merkle_node(l, r):
    """merkle_tree_gadget.rs:7-12,106: unpadded 2-block sponge."""
    return mimc_sponge([l, r])


def mimc_trace(blocks):
    """Witness trace of the in-circuit sponge (mimc_hash_gadget.rs:108-150): per round two
    multipliers (t,t,t^2), (t^2,t,t^3).  Returns (aL, aR, aO, digest)."""
    aL, aR, aO = [], [], []
    s = 0
    for b in blocks:
        s = (s + b) % L
        for c in _MIMC_CONSTS:
            t = (s + c) % L
            t2 = t * t % L
            t3 = t2 * t % L
            aL += [t, t2]
            aR += [t, t]
            aO += [t2, t3]
            s = t3
    return aL, aR, aO, s


# ----------------------------------------------------------------------------- R1CS
# Variables: ("L", i) ("R", i) ("O", i) ("V", i) ("1", 0).  An LC is a list of (var, coeff).
class ConstraintSystem:
    """Shared front half of bulletproofs' r1cs::Prover / Verifier."""

    def __init__(self, transcript, prover):
        self.t = transcript
        self.t.append_message(b"dom-sep", b"r1cs v1")
        self.is_prover = prover
        self.constraints = []
        self.aL, self.aR, self.aO = [], [], []
        self.v, self.v_blinding, self.V = [], [], []
        self.num_vars = 0
        self.pending = None

    def eval(self, lc):
        acc = 0
        for (k, i), c in lc:
            val = {"L": self.aL, "R": self.aR, "O": self.aO, "V": self.v}[k][i] if k != "1" else 1
            acc = (acc + val * c) % L
        return acc

    def commit(self, v, blinding):
        assert self.is_prover
        Vp = ristretto_encode(pedersen_commit(v, blinding))
        self.t.append_point(b"V", Vp)
        self.v.append(v)
        self.v_blinding.append(blinding)
        self.V.append(Vp)
        return Vp, ("V", len(self.v) - 1)

    def commit_verifier(self, Venc):
        self.t.append_point(b"V", Venc)
        self.V.append(Venc)
        return ("V", len(self.V) - 1)

    def multiply(self, left, right):
        i = self.num_vars
        self.num_vars += 1
        if self.is_prover:
            l, r = self.eval(left), self.eval(right)
            self.aL.append(l)
            self.aR.append(r)
            self.aO.append(l * r % L)
        self.constraints.append(list(left) + [(("L", i), L - 1)])
        self.constraints.append(list(right) + [(("R", i), L - 1)])
        return ("L", i), ("R", i), ("O", i)

    def allocate_multiplier(self, lr=None):
        i = self.num_vars
        self.num_vars += 1
        if self.is_prover:
            l, r = lr
            self.aL.append(l % L)
            self.aR.append(r % L)
            self.aO.append(l * r % L)
        return ("L", i), ("R", i), ("O", i)

    def allocate(self, val=None):
        if self.pending is None:
            i = self.num_vars
            self.num_vars += 1
            self.pending = i
            if self.is_prover:
                self.aL.append(val % L)
                self.aR.append(0)
                self.aO.append(0)
            return ("L", i)
        i = self.pending
        self.pending = None
        if self.is_prover:
            self.aR[i] = val % L
            self.aO[i] = self.aL[i] * self.aR[i] % L
        return ("R", i)

    def constrain(self, lc):
        self.constraints.append(list(lc))

    def flatten(self, z, n, m):
        wL, wR, wO, wV, wc = [0] * n, [0] * n, [0] * n, [0] * m, 0
        e = z
        for lc in self.constraints:
            for (k, i), c in lc:
                if k == "L":
                    wL[i] = (wL[i] + e * c) % L
                elif k == "R":
                    wR[i] = (wR[i] + e * c) % L
                elif k == "O":
                    wO[i] = (wO[i] + e * c) % L
                elif k == "V":
                    wV[i] = (wV[i] - e * c) % L
                else:
                    wc = (wc - e * c) % L
            e = e * z % L
        return wL, wR, wO, wV, wc


def _ip(a, b):
    return sum(x * y for x, y in zip(a, b)) % L


def ipp_create(t, Q, Gf, Hf, G, H, a, b):
    """InnerProductProof::create.  Returns (L_vec, R_vec, a, b)."""
    n = len(G)
    assert n & (n - 1) == 0 and n >= 1
    t.append_message(b"dom-sep", b"ipp v1")
    t.append_u64(b"n", n)
    G, H, a, b = list(G), list(H), list(a), list(b)
    Ls, Rs = [], []
    first = True
    while n != 1:
        n //= 2
        aL, aR, bL, bR = a[:n], a[n:], b[:n], b[n:]
        GL, GR, HL, HR = G[:n], G[n:], H[:n], H[n:]
        cL, cR = _ip(aL, bR), _ip(aR, bL)
        if first:
            Lp = msm([x * g for x, g in zip(aL, Gf[n:])] + [x * h for x, h in zip(bR, Hf[:n])] + [cL],
                     GR + HL + [Q])
            Rp = msm([x * g for x, g in zip(aR, Gf[:n])] + [x * h for x, h in zip(bL, Hf[n:])] + [cR],
                     GL + HR + [Q])
        else:
            Lp = msm(aL + bR + [cL], GR + HL + [Q])
            Rp = msm(aR + bL + [cR], GL + HR + [Q])
        Le, Re = ristretto_encode(Lp), ristretto_encode(Rp)
        Ls.append(Le)
        Rs.append(Re)
        t.append_point(b"L", Le)
        t.append_point(b"R", Re)
        u = t.challenge_scalar(b"u")
        ui = sc_inv(u)
        a = [(x * u + ui * y) % L for x, y in zip(aL, aR)]
        b = [(x * ui + u * y) % L for x, y in zip(bL, bR)]
        if first:
            G = [msm([ui * Gf[i], u * Gf[n + i]], [GL[i], GR[i]]) for i in range(n)]
            H = [msm([u * Hf[i], ui * Hf[n + i]], [HL[i], HR[i]]) for i in range(n)]
        else:
            G = [msm([ui, u], [GL[i], GR[i]]) for i in range(n)]
            H = [msm([u, ui], [HL[i], HR[i]]) for i in range(n)]
        first = False
    return Ls, Rs, a[0], b[0]


def ipp_verification_scalars(t, n, Ls, Rs):
    lg = len(Ls)
    if lg >= 32 or n != (1 << lg):
        raise ValueError("bad ipp size")
    t.append_message(b"dom-sep", b"ipp v1")
    t.append_u64(b"n", n)
    ch = []
    for Le, Re in zip(Ls, Rs):
        t.validate_and_append_point(b"L", Le)
        t.validate_and_append_point(b"R", Re)
        ch.append(t.challenge_scalar(b"u"))
    chi = [sc_inv(c) for c in ch]
    allinv = 1
    for c in chi:
        allinv = allinv * c % L
    usq = [c * c % L for c in ch]
    uisq = [c * c % L for c in chi]
    s = [allinv]
    for i in range(1, n):
        lg_i = i.bit_length() - 1
        s.append(s[i - (1 << lg_i)] * usq[lg - 1 - lg_i] % L)
    return usq, uisq, s


def next_pow2(n):
    return 1 if n <= 1 else 1 << (n - 1).bit_length()


def r1cs_prove(cs: ConstraintSystem, gens_capacity, external32, legacy_framing=False):
    """Prover::prove(&bp_gens); returns proof bytes + dict of parts."""
    t = cs.t
    t.append_u64(b"m", len(cs.v))
    rng = t.build_rng([sc_bytes(x) for x in cs.v_blinding], external32)
    n = n1 = len(cs.aL)
    N = next_pow2(n)
    if gens_capacity < N:
        raise ValueError("InvalidGeneratorsLength")
    G, H = bulletproof_gens(N)
    ib, ob, sb = rng.scalar(), rng.scalar(), rng.scalar()
    sL = [rng.scalar() for _ in range(n)]
    sR = [rng.scalar() for _ in range(n)]
    A_I = ristretto_encode(msm([ib] + cs.aL + cs.aR, [B_BLINDING] + G[:n] + H[:n]))
    A_O = ristretto_encode(msm([ob] + cs.aO, [B_BLINDING] + G[:n]))
    S = ristretto_encode(msm([sb] + sL + sR, [B_BLINDING] + G[:n] + H[:n]))
    t.append_point(b"A_I1", A_I)
    t.append_point(b"A_O1", A_O)
    t.append_point(b"S1", S)
    t.append_message(b"dom-sep", b"r1cs-1phase")
    Z32 = bytes(32)
    t.append_point(b"A_I2", Z32)
    t.append_point(b"A_O2", Z32)
    t.append_point(b"S2", Z32)
    y = t.challenge_scalar(b"y")
    z = t.challenge_scalar(b"z")
    wL, wR, wO, wV, _ = cs.flatten(z, n, len(cs.v))
    yinv = sc_inv(y)
    ypow = [pow(y, i, L) for i in range(N)]
    yipow = [pow(yinv, i, L) for i in range(N)]
    l1 = [(cs.aL[i] + yipow[i] * wR[i]) % L for i in range(n)]
    l2 = list(cs.aO)
    l3 = list(sL)
    r0 = [(wO[i] - ypow[i]) % L for i in range(n)]
    r1 = [(ypow[i] * cs.aR[i] + wL[i]) % L for i in range(n)]
    r3 = [ypow[i] * sR[i] % L for i in range(n)]
    t1 = _ip(l1, r0)
    t2 = (_ip(l1, r1) + _ip(l2, r0)) % L
    t3 = (_ip(l2, r1) + _ip(l3, r0)) % L
    t4 = (_ip(l1, r3) + _ip(l3, r1)) % L
    t5 = _ip(l2, r3)
    t6 = _ip(l3, r3)
    tb1, tb3, tb4, tb5, tb6 = (rng.scalar() for _ in range(5))
    Ts = [ristretto_encode(pedersen_commit(tv, tb)) for tv, tb in
          ((t1, tb1), (t3, tb3), (t4, tb4), (t5, tb5), (t6, tb6))]
    for lab, Tp in zip((b"T_1", b"T_3", b"T_4", b"T_5", b"T_6"), Ts):
        t.append_point(lab, Tp)
    u = t.challenge_scalar(b"u")
    x = t.challenge_scalar(b"x")
    tb2 = _ip(wV, cs.v_blinding)
    tx = 0
    txb = 0
    for k, (tc, tbc) in enumerate(((t1, tb1), (t2, tb2), (t3, tb3), (t4, tb4), (t5, tb5), (t6, tb6)), 1):
        tx = (tx + tc * pow(x, k, L)) % L
        txb = (txb + tbc * pow(x, k, L)) % L
    x2, x3 = x * x % L, x * x % L * x % L
    lv = [(x * l1[i] + x2 * l2[i] + x3 * l3[i]) % L for i in range(n)] + [0] * (N - n)
    rv = [(r0[i] + x * r1[i] + x3 * r3[i]) % L for i in range(n)] + [(-ypow[i]) % L for i in range(n, N)]
    eb = x * (ib + x * (ob + x * sb)) % L
    t.append_scalar(b"t_x", tx)
    t.append_scalar(b"t_x_blinding", txb)
    t.append_scalar(b"e_blinding", eb)
    w = t.challenge_scalar(b"w")
    Q = pt_mul(w, BASEPOINT)
    Gf = [1] * n1 + [u] * (N - n1)
    Hf = [yipow[i] * Gf[i] % L for i in range(N)]
    Ls, Rs, a, b = ipp_create(t, Q, Gf, Hf, G, H, lv, rv)
    parts = dict(A_I1=A_I, A_O1=A_O, S1=S, T=Ts, t_x=tx, t_x_blinding=txb, e_blinding=eb,
                 L=Ls, R=Rs, a=a, b=b, y=y, z=z, u=u, x=x, w=w)
    out = bytearray()
    if legacy_framing:
        out += A_I + A_O + S + Z32 * 3
    else:
        out += b"\x00" + A_I + A_O + S
    for Tp in Ts:
        out += Tp
    out += sc_bytes(tx) + sc_bytes(txb) + sc_bytes(eb)
    for Le, Re in zip(Ls, Rs):
        out += Le + Re
    out += sc_bytes(a) + sc_bytes(b)
    return bytes(out), parts


def proof_from_bytes(buf, legacy_framing=False):
    """R1CSProof::from_bytes; raises ValueError (FormatError)."""
    def canon(bs):
        v = int.from_bytes(bs, "little")
        if v >= L:
            raise ValueError("non-canonical scalar")
        return v
    if legacy_framing:
        body = buf
        if len(body) % 32 or len(body) < 14 * 32:
            raise ValueError("format")
        f = [body[i:i + 32] for i in range(0, len(body), 32)]
        A = f[:6]
        f = f[6:]
    else:
        if len(buf) == 0:
            raise ValueError("format")
        ver, body = buf[0], buf[1:]
        if len(body) % 32 or ver not in (0, 1):
            raise ValueError("format")
        if len(body) < (11 if ver == 0 else 14) * 32:
            raise ValueError("format")
        f = [body[i:i + 32] for i in range(0, len(body), 32)]
        if ver == 0:
            A = f[:3] + [bytes(32)] * 3
            f = f[3:]
        else:
            A = f[:6]
            f = f[6:]
    Ts, f = f[:5], f[5:]
    tx, txb, eb = canon(f[0]), canon(f[1]), canon(f[2])
    f = f[3:]
    if len(f) < 2 or (len(f) - 2) % 2:
        raise ValueError("format")
    lg = (len(f) - 2) // 2
    if lg >= 32:
        raise ValueError("format")
    Ls = [f[2 * i] for i in range(lg)]
    Rs = [f[2 * i + 1] for i in range(lg)]
    a, b = canon(f[-2]), canon(f[-1])
    return dict(A=A, T=Ts, t_x=tx, t_x_blinding=txb, e_blinding=eb, L=Ls, R=Rs, a=a, b=b)


def r1cs_verify(cs: ConstraintSystem, proof_bytes, gens_capacity, external32, legacy_framing=False):
    """Verifier::verify.  Returns True/False (VerificationError / FormatError -> False)."""
    try:
        pr = proof_from_bytes(proof_bytes, legacy_framing)
        t = cs.t
        t.append_u64(b"m", len(cs.V))
        n = n1 = cs.num_vars
        for lab, enc in zip((b"A_I1", b"A_O1", b"S1"), pr["A"][:3]):
            t.validate_and_append_point(lab, enc)
        t.append_message(b"dom-sep", b"r1cs-1phase")
        N = next_pow2(n)
        if gens_capacity < N:
            return False
        for lab, enc in zip((b"A_I2", b"A_O2", b"S2"), pr["A"][3:]):
            t.append_point(lab, enc)
        y = t.challenge_scalar(b"y")
        z = t.challenge_scalar(b"z")
        for lab, enc in zip((b"T_1", b"T_3", b"T_4", b"T_5", b"T_6"), pr["T"]):
            t.validate_and_append_point(lab, enc)
        u = t.challenge_scalar(b"u")
        x = t.challenge_scalar(b"x")
        t.append_scalar(b"t_x", pr["t_x"])
        t.append_scalar(b"t_x_blinding", pr["t_x_blinding"])
        t.append_scalar(b"e_blinding", pr["e_blinding"])
        w = t.challenge_scalar(b"w")
        wL, wR, wO, wV, wc = cs.flatten(z, n, len(cs.V))
        usq, uisq, s = ipp_verification_scalars(t, N, pr["L"], pr["R"])
    except ValueError:
        return False
    a, b = pr["a"], pr["b"]
    yinv = sc_inv(y)
    yi = [pow(yinv, i, L) for i in range(N)]
    ynwR = [wR[i] * yi[i] % L for i in range(n)] + [0] * (N - n)
    delta = _ip(ynwR[:n], wL)
    uf = [1] * n1 + [u] * (N - n1)
    wLp = wL + [0] * (N - n)
    wOp = wO + [0] * (N - n)
    g = [uf[i] * (x * ynwR[i] - a * s[i]) % L for i in range(N)]
    h = [uf[i] * (yi[i] * (x * wLp[i] + wOp[i] - b * s[N - 1 - i]) - 1) % L for i in range(N)]
    rng = t.build_rng([], external32)
    r = rng.scalar()
    xx = x * x % L
    rxx = r * xx % L
    xxx = x * xx % L
    scal = [x, xx, xxx, u * x % L, u * xx % L, u * xxx % L]
    scal += [wv * rxx % L for wv in wV]
    scal += [r * x % L, rxx * x % L, rxx * xx % L, rxx * xxx % L, rxx * xx % L * xx % L]
    scal += [(w * (pr["t_x"] - a * b) + r * (xx * (wc + delta) - pr["t_x"])) % L]
    scal += [(-pr["e_blinding"] - r * pr["t_x_blinding"]) % L]
    scal += g + h + usq + uisq
    encs = pr["A"] + cs.V + pr["T"]
    pts = []
    for e in encs:
        p = ristretto_decode(e)
        if p is None:
            return False
        pts.append(p)
    G, H = bulletproof_gens(N)
    pts += [BASEPOINT, B_BLINDING] + G + H
    for e in pr["L"] + pr["R"]:
        p = ristretto_decode(e)
        if p is None:
            return False
        pts.append(p)
    res = msm(scal, pts)
    return pt_eq(res, IDENT)
