"""Host-side mirror of the two reference gadgets that sit directly on the hot path, MimcHash256 and MerkleTree256,
plus fast (numpy) builders of the synthetic benchmark circuits BASELINE.json names.

  mimc_sponge / mimc_encryption wiring      src/mimc_hash/mimc_hash_gadget.rs:108-150
  MimcHash256.preprocess / pad / assemble   src/mimc_hash/mimc_hash_gadget.rs:15-106
  MerkleTree256.parse                        src/merkle_tree/merkle_tree_gadget.rs:75-107
  hash_witness (leaf is hashed first)        src/bin/prover.rs:160-190, 324-334

The constraint wiring is host work in the reference too (Rust); what runs on the device is the witness trace of the
MiMC chains (bpg_mimc_sponge_batch with trace) and everything inside prove / verify.
"""
import json
import os

import numpy as np

from .api import L_ORDER, ONE, Context

ROUNDS = 486
_CONSTS = None


def round_constants():
    """the 486 constants as ints (Scalar::from_bits of mimc_consts.rs, all < l)"""
    global _CONSTS
    if _CONSTS is None:
        p = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "mimc_consts.json")
        with open(p) as f:
            _CONSTS = [(int.from_bytes(bytes.fromhex(h), "little") & ((1 << 255) - 1)) % L_ORDER for h in json.load(f)]
    return _CONSTS


# ----------------------------------------------------------------------------- host big-int MiMC (front-end only)
def mimc_encrypt_int(x):
    for c in round_constants():
        t = (x + c) % L_ORDER
        x = t * t % L_ORDER * t % L_ORDER
    return x


def mimc_sponge_int(blocks):
    s = 0
    for b in blocks:
        s = mimc_encrypt_int((s + b) % L_ORDER)
    return s


def be_to_scalars(data):
    le = bytes(reversed(data))
    if len(le) % 32:
        le += bytes(32 - len(le) % 32)
    return [int.from_bytes(le[i:i + 32], "little") & ((1 << 255) - 1) for i in range(0, len(le), 32)]


def mimc_preprocess(witness_scalars):
    """MimcHash256::preprocess -> derived witnesses [padded_block, padding] or [extra_block]."""
    last = witness_scalars[-1]
    le = last.to_bytes(32, "little").rstrip(b"\x00")
    if len(le) < 32:
        padn = 32 - len(le)
        padded = int.from_bytes(le + bytes([padn]) * padn, "little") & ((1 << 255) - 1)
        return [padded, (padded - last) % L_ORDER]
    return [int.from_bytes(bytes([32]) * 32, "little") & ((1 << 255) - 1)]


def mimc_preprocess_blocks(scalars):
    """the padded block list mimc_hash absorbs (mimc.rs:61-97) for a preimage given as be_to_scalars(bytes)"""
    d = mimc_preprocess(scalars)
    return list(scalars[:-1]) + [d[0]] if len(d) == 2 else list(scalars) + [d[0]]


# ----------------------------------------------------------------------------- reference-shaped wiring (generic, slow)
def mimc_sponge_wire(cs, preimage_lcs):
    """MimcHash256::mimc_sponge over any ConstraintSystem mirror (Prover / Verifier of api.py)."""
    consts = round_constants()
    state = []
    for lc in preimage_lcs:
        p = list(state) + list(lc)
        for c in consts:
            lin = p + [(ONE, c)]
            x, _, sqr = cs.multiply(lin, lin)
            _, _, cube = cs.multiply([(sqr, 1)], [(x, 1)])
            p = [(cube, 1)]
        state = p
    return state


def mimc_hash_gadget_wire(cs, witness_vars, derived_vars, image_lc):
    """MimcHash256::assemble: honest-padding constraint, sponge, hash - image = 0."""
    commitments = list(witness_vars)
    padded = derived_vars[0]
    if len(derived_vars) == 2:
        last = commitments.pop()
        cs.constrain([(last, 1), (derived_vars[1], 1), (padded, L_ORDER - 1)])
    commitments.append(padded)
    h = mimc_sponge_wire(cs, [[(v, 1)] for v in commitments])
    cs.constrain(h + [(var, (-c) % L_ORDER) for var, c in image_lc])


def merkle_wire(cs, root_lc, pattern, witness_lcs, instance_lcs):
    """MerkleTree256::assemble for a pattern given as nested tuples of 'W' / 'I'."""
    w, i = list(witness_lcs), list(instance_lcs)

    def parse(p):
        if p == "W":
            return mimc_sponge_wire(cs, [w.pop(0)])
        if p == "I":
            return mimc_sponge_wire(cs, [i.pop(0)])
        l, r = p
        pre = []
        for side in (l, r):
            if side == "W":
                pre.append(w.pop(0))
            elif side == "I":
                pre.append(i.pop(0))
            else:
                pre.append(parse(side))
        return mimc_sponge_wire(cs, pre)

    h = parse(pattern)
    cs.constrain(h + [(var, (-c) % L_ORDER) for var, c in root_lc])


# ----------------------------------------------------------------------------- fast builders (numpy CSR)
_K = {"L": 0, "R": 1, "O": 2, "V": 3, "1": 4}


def _vid(kind, idx):
    return (_K[kind] << 29) | idx


class _Csr:
    """Row-by-row CSR emitter with a vectorised path for the 484 'inner' MiMC rounds of a block."""

    def __init__(self):
        self.var_chunks, self.coef_chunks, self.len_chunks = [], [], []
        self.nmul = 0

    def rows(self, rows):
        v = np.fromiter((t[0] for r in rows for t in r), dtype=np.uint32)
        c = np.frombuffer(b"".join(int(t[1] % L_ORDER).to_bytes(32, "little") for r in rows for t in r), dtype=np.uint8).reshape(-1, 32)
        self.var_chunks.append(v)
        self.coef_chunks.append(c)
        self.len_chunks.append(np.fromiter((len(r) for r in rows), dtype=np.uint32))

    def mimc_block(self, input_terms):
        """one absorbed block: 486 rounds, 972 multipliers; input_terms = LC of (state + block). Returns O var id."""
        consts = round_constants()
        j0 = self.nmul
        one, minus1 = 1, L_ORDER - 1
        # round 0 uses the caller's LC, the remaining rounds the previous cube
        rows = []
        lin0 = list(input_terms) + [(_vid("1", 0), consts[0])]
        rows.append(lin0 + [(_vid("L", j0), minus1)])
        rows.append(lin0 + [(_vid("R", j0), minus1)])
        rows.append([(_vid("O", j0), one), (_vid("L", j0 + 1), minus1)])
        rows.append([(_vid("L", j0), one), (_vid("R", j0 + 1), minus1)])
        self.rows(rows)
        # rounds 1..485, vectorised: per round 4 rows with 3,3,2,2 terms
        r = np.arange(1, ROUNDS, dtype=np.uint32)
        a = j0 + 2 * r          # multiplier (t, t, t^2)
        prev = a - 1            # cube of the previous round
        V = np.empty((ROUNDS - 1, 10), dtype=np.uint32)
        V[:, 0] = (2 << 29) | prev; V[:, 1] = 4 << 29; V[:, 2] = (0 << 29) | a
        V[:, 3] = (2 << 29) | prev; V[:, 4] = 4 << 29; V[:, 5] = (1 << 29) | a
        V[:, 6] = (2 << 29) | a; V[:, 7] = (0 << 29) | (a + 1)
        V[:, 8] = (0 << 29) | a; V[:, 9] = (1 << 29) | (a + 1)
        C = np.empty((ROUNDS - 1, 10, 32), dtype=np.uint8)
        one_b = np.frombuffer(one.to_bytes(32, "little"), dtype=np.uint8)
        m1_b = np.frombuffer(minus1.to_bytes(32, "little"), dtype=np.uint8)
        cb = np.frombuffer(b"".join(c.to_bytes(32, "little") for c in consts[1:]), dtype=np.uint8).reshape(-1, 32)
        for k in (0, 3, 6, 8):
            C[:, k, :] = one_b
        for k in (2, 5, 7, 9):
            C[:, k, :] = m1_b
        C[:, 1, :] = cb
        C[:, 4, :] = cb
        self.var_chunks.append(V.reshape(-1))
        self.coef_chunks.append(C.reshape(-1, 32))
        self.len_chunks.append(np.tile(np.array([3, 3, 2, 2], dtype=np.uint32), ROUNDS - 1))
        self.nmul += 2 * ROUNDS
        return _vid("O", self.nmul - 1)

    def finish(self):
        lens = np.concatenate(self.len_chunks) if self.len_chunks else np.zeros(0, np.uint32)
        row_ptr = np.zeros(len(lens) + 1, dtype=np.uint32)
        np.cumsum(lens, out=row_ptr[1:])
        tv = np.concatenate(self.var_chunks) if self.var_chunks else np.zeros(0, np.uint32)
        tc = np.concatenate(self.coef_chunks) if self.coef_chunks else np.zeros((0, 32), np.uint8)
        return row_ptr, np.ascontiguousarray(tv), np.ascontiguousarray(tc)


def _enc(xs):
    return b"".join(int(x % L_ORDER).to_bytes(32, "little") for x in xs)


def merkle_path_instance(depth, leaf=b"\x43", seed=4, ctx=None, label=b"merkle_path", trace_on_device=True):
    """BASELINE config 2: Merkle membership of `leaf` at the left-most position of a depth-`depth` MiMC tree.
    CLI semantics (prover.rs:324-334): the witness leaf is proven to hash to a committed image with MimcHash256
    (972 multipliers), sibling instances are MiMC-hashed constants, every inner node is an unpadded 2-block sponge
    (1944 multipliers).  n = 972 + 1944*depth multipliers, m = 4 commitments.
    -> dict(label, n, m, vals, blinds, aL, aR, aO, csr=(row_ptr, term_var, term_coeff) numpy, root)"""
    rng = np.random.default_rng(seed)
    w0 = be_to_scalars(leaf)
    assert len(w0) == 1
    derived = mimc_preprocess(w0)
    assert len(derived) == 2
    padded, padding = derived
    image = mimc_sponge_int([padded])
    siblings = [mimc_sponge_int([int.from_bytes(rng.bytes(32), "little") & ((1 << 255) - 1), 7]) for _ in range(depth)]
    # committed variables: V0 = leaf, V1 = image, V2 = padded block, V3 = padding
    vals = [w0[0], image, padded, padding]
    blinds = [int.from_bytes(rng.bytes(64), "little") % L_ORDER for _ in vals]
    csr = _Csr()
    csr.rows([[(_vid("V", 0), 1), (_vid("V", 3), 1), (_vid("V", 2), L_ORDER - 1)]])  # honest padding
    out = csr.mimc_block([(_vid("V", 2), 1)])
    csr.rows([[(out, 1), (_vid("V", 1), L_ORDER - 1)]])  # hash - image = 0
    sponges = [[padded]]
    cur_lc, cur_val = [(_vid("V", 1), 1)], image
    for k in range(depth):
        mid = csr.mimc_block(cur_lc)                       # absorb the running hash (left child)
        out = csr.mimc_block([(mid, 1), (_vid("1", 0), siblings[k])])  # absorb the sibling (right child, instance)
        sponges.append([cur_val, siblings[k]])
        cur_val = mimc_sponge_int([cur_val, siblings[k]])
        cur_lc = [(out, 1)]
    csr.rows([cur_lc + [(_vid("1", 0), (-cur_val) % L_ORDER)]])  # hash - root = 0
    n = csr.nmul
    # witness trace: every absorbed block contributes 972 (a_L, a_R, a_O) triples
    if trace_on_device:
        ctx = ctx or Context.default()
        _, tr = ctx.mimc_sponge_batch([[int(b).to_bytes(32, "little") for b in s] for s in sponges], trace=True)
        t = np.frombuffer(tr, dtype=np.uint8).reshape(n, 3, 32)
        aL, aR, aO = t[:, 0, :].tobytes(), t[:, 1, :].tobytes(), t[:, 2, :].tobytes()
    else:
        aL, aR, aO = _mimc_trace_host(sponges)
    return dict(label=label, n=n, m=4, vals=_enc(vals), blinds=_enc(blinds), aL=aL, aR=aR, aO=aO, csr=csr.finish(), root=cur_val)


class _Instance(dict):
    """Instance dict whose CSR coefficient array is SHARED with the other instances of the same circuit family (160 MB at 2^20
    multipliers): the only per-instance coefficient is the last one (digest / root constant), kept in "last_coeff" and written
    into the shared array whenever this instance's "csr" is read.  Single-threaded use of "csr" is assumed."""

    def __getitem__(self, key):
        v = dict.__getitem__(self, key)
        if key == "csr" and "last_coeff" in self:
            v[2][-1, :] = np.frombuffer(dict.__getitem__(self, "last_coeff"), dtype=np.uint8)
        return v


def _mimc_trace_host(sponges):
    """(a_L, a_R, a_O) bytes of the MiMC chains in gadget order, host big-int (front-end / CPU-only callers)"""
    aL, aR, aO = [], [], []
    consts = round_constants()
    for s in sponges:
        st = 0
        for b in s:
            st = (st + b) % L_ORDER
            for c in consts:
                tt = (st + c) % L_ORDER
                t2 = tt * tt % L_ORDER
                t3 = t2 * tt % L_ORDER
                aL += [tt, t2]; aR += [tt, tt]; aO += [t2, t3]
                st = t3
    return _enc(aL), _enc(aR), _enc(aO)


def mimc_chain_instances(nblocks, seeds, ctx=None, label=b"mimc_chain", trace_on_device=True):
    """Synthetic circuit of `nblocks` absorbed MiMC blocks (972 multipliers each) in one long sponge; with
    nblocks = 1022 this is the 2^20-multiplier class of BASELINE configs[3] (993 384 multipliers, N = 2^20: the size of the
    reference's ignored 512-leaf test, merkle_tree_gadget.rs:473-545).  One instance per seed: the committed first block
    V_0 (and with it the whole witness trace and the digest constant of the last constraint) differs per seed, blocks
    1.. are instance constants shared by all (seed 5's).  All traces come from ONE batched device call."""
    base = np.random.default_rng(5)
    blocks = [int.from_bytes(base.bytes(32), "little") & ((1 << 252) - 1) for _ in range(nblocks)]
    csr = _Csr()
    cur = csr.mimc_block([(_vid("V", 0), 1)])
    for b in blocks[1:]:
        cur = csr.mimc_block([(cur, 1), (_vid("1", 0), b)])
    csr.rows([[(cur, 1), (_vid("1", 0), 0)]])  # digest coefficient patched per instance below
    n = csr.nmul
    row_ptr, tv, tc = csr.finish()
    firsts, blinds = [], []
    for sd in seeds:
        rng = np.random.default_rng(sd)
        first = int.from_bytes(rng.bytes(32), "little") & ((1 << 252) - 1)
        if sd == 5:
            first = blocks[0]
        firsts.append(first)
        blinds.append(int.from_bytes(rng.bytes(64), "little") % L_ORDER)
    sponges = [[f] + blocks[1:] for f in firsts]
    if trace_on_device:
        ctx = ctx or Context.default()
        digs, tr = ctx.mimc_sponge_batch([[int(b).to_bytes(32, "little") for b in s] for s in sponges], trace=True)
        t = np.frombuffer(tr, dtype=np.uint8).reshape(len(seeds), n, 3, 32)
        traces = [(t[k, :, 0, :].tobytes(), t[k, :, 1, :].tobytes(), t[k, :, 2, :].tobytes()) for k in range(len(seeds))]
        dvals = [int.from_bytes(d, "little") for d in digs]
    else:
        traces = [_mimc_trace_host([s]) for s in sponges]
        dvals = [int.from_bytes(tr[2][-32:], "little") for tr in traces]
    out = []
    for k in range(len(seeds)):
        aL, aR, aO = traces[k]
        out.append(_Instance(label=label, n=n, m=1, vals=_enc([firsts[k]]), blinds=_enc([blinds[k]]), aL=aL, aR=aR, aO=aO,
                             csr=(row_ptr, tv, tc), root=dvals[k], last_coeff=((-dvals[k]) % L_ORDER).to_bytes(32, "little")))
    return out


def mimc_chain_instance(nblocks, seed=5, ctx=None, label=b"mimc_chain", trace_on_device=True):
    """one instance of mimc_chain_instances (seed 5: the round-1 benchmark witness)"""
    inst = mimc_chain_instances(nblocks, [seed], ctx=ctx, label=label, trace_on_device=trace_on_device)[0]
    if nblocks <= 64 and seed == 5:
        base = np.random.default_rng(5)
        blocks = [int.from_bytes(base.bytes(32), "little") & ((1 << 252) - 1) for _ in range(nblocks)]
        assert inst["root"] == mimc_sponge_int(blocks)
    return inst


# W1 of the reference's Merkle tests (merkle_tree_gadget.rs:126-131), big-endian as written there
REF_W1 = bytes.fromhex("0522a64d7b931e21760cf955a15fcc793e8a52b42a56ab03afddec8beb668749")
# root of the 512-leaf tree whose leaves are all W1 (merkle_tree_gadget.rs:476, 502)
REF_ROOT_512 = bytes.fromhex("038c137beec8e2edfb5c48cbd063f04e569139d2221a4eb7befb85aa1bf8ba40")


def merkle_tree_instances(nleaves, seeds, ctx=None, label=b"MerkleTree", trace_on_device=True):
    """The reference's largest circuit: MerkleTree256 over a complete binary tree of `nleaves` committed leaves, library API
    (test_merkle_tree_gadget_512, merkle_tree_gadget.rs:473-545: transcript "MerkleTree", 512 commitments, 511 nodes of
    1944 multipliers = 993 384 multipliers, BulletproofGens::new(1048576, 1)).  Multipliers are laid out in the order
    MerkleTree256::parse recurses (left subtree, right subtree, then the node's own 2-block sponge; merkle_tree_gadget.rs:75-107),
    leaves are consumed left to right, the last constraint is  node_root - root = 0  with the root as a constant.
    One instance per seed: seed None = the reference's test (every leaf = W1, root pinned by merkle_tree_gadget.rs:476); any
    other seed draws random 252-bit leaves.  Node values and witness traces come from ONE batched device call per tree level
    (all nodes of a level are independent: 256, 128, ... 1 two-block sponges per instance)."""
    assert nleaves >= 2 and nleaves & (nleaves - 1) == 0
    levels = nleaves.bit_length() - 1
    csr = _Csr()
    order = []  # (level, index in level) of every node in multiplier (post-)order

    def emit(level, j):
        """node j of `level` (level 0 = parents of the leaves) -> variable id of its digest"""
        if level == 0:
            left, right = [(_vid("V", 2 * j), 1)], [(_vid("V", 2 * j + 1), 1)]
        else:
            left = [(emit(level - 1, 2 * j), 1)]
            right = [(emit(level - 1, 2 * j + 1), 1)]
        mid = csr.mimc_block(left)
        out = csr.mimc_block([(mid, 1)] + right)
        order.append((level, j))
        return out

    top = emit(levels - 1, 0)
    csr.rows([[(top, 1), (_vid("1", 0), 0)]])  # root constant patched per instance
    n = csr.nmul
    row_ptr, tv, tc = csr.finish()
    K = len(seeds)
    leaves, blinds = [], []
    for sd in seeds:
        if sd is None:
            w1 = int.from_bytes(REF_W1, "big") & ((1 << 255) - 1)
            leaves.append([w1] * nleaves)
            rng = np.random.default_rng(512)
        else:
            rng = np.random.default_rng(sd)
            leaves.append([int.from_bytes(rng.bytes(32), "little") & ((1 << 252) - 1) for _ in range(nleaves)])
        blinds.append([int.from_bytes(rng.bytes(64), "little") % L_ORDER for _ in range(nleaves)])
    # level by level: inputs of level l are the digests of level l - 1
    cur = leaves
    traces, width = [], nleaves // 2
    for level in range(levels):
        pairs = [[cur[k][2 * j], cur[k][2 * j + 1]] for k in range(K) for j in range(width)]
        if trace_on_device:
            ctx = ctx or Context.default()
            digs, tr = ctx.mimc_sponge_batch([[int(a).to_bytes(32, "little"), int(b).to_bytes(32, "little")] for a, b in pairs], trace=True)
            traces.append(np.frombuffer(tr, dtype=np.uint8).reshape(K, width, 2 * 2 * ROUNDS, 3, 32))
            vals = [int.from_bytes(d, "little") for d in digs]
        else:
            rows = []
            vals = []
            for a, b in pairs:
                aL, aR, aO = _mimc_trace_host([[a, b]])
                t = np.stack([np.frombuffer(x, dtype=np.uint8).reshape(-1, 32) for x in (aL, aR, aO)], axis=1)
                rows.append(t)
                vals.append(int.from_bytes(aO[-32:], "little"))
            traces.append(np.stack(rows).reshape(K, width, 2 * 2 * ROUNDS, 3, 32))
        cur = [vals[k * width:(k + 1) * width] for k in range(K)]
        width //= 2
    pos = {}
    for p_, (lv, j) in enumerate(order):
        pos.setdefault(lv, []).append((j, p_))
    out = []
    for k in range(K):
        full = np.empty((len(order), 2 * 2 * ROUNDS, 3, 32), dtype=np.uint8)
        for lv, lst in pos.items():
            js = np.array([j for j, _ in lst])
            ps = np.array([p_ for _, p_ in lst])
            full[ps] = traces[lv][k][js]
        flat = full.reshape(n, 3, 32)
        root = cur[k][0]
        out.append(_Instance(label=label, n=n, m=nleaves, vals=_enc(leaves[k]), blinds=_enc(blinds[k]), aL=flat[:, 0, :].tobytes(),
                             aR=flat[:, 1, :].tobytes(), aO=flat[:, 2, :].tobytes(), csr=(row_ptr, tv, tc), root=root,
                             last_coeff=((-root) % L_ORDER).to_bytes(32, "little")))
    if seeds and seeds[0] is None and nleaves == 512:
        assert out[0]["root"] == int.from_bytes(REF_ROOT_512, "big"), "512-leaf root differs from merkle_tree_gadget.rs:476"
    return out


def bounds_check_batch_instance(count, nbytes=8, seed=5, label=b"bounds_batch", lo=0, hi=None, values=None):
    """BASELINE config 3: `count` BoundsCheck gadgets (min <= v <= max on `nbytes`-byte values) in ONE proof.
    Wiring per value as in bounds_check_gadget.rs:23-49 + utils.rs:5-35: commitments (v, a = v - min, b = max - v),
    constraint a + b - (max - min) = 0, and two range proofs of n = 8*nbytes bits (per bit: allocate_multiplier(1 - bit, bit),
    o = 0, a_i + b_i - 1 = 0; finally x - sum b_i 2^i = 0).  count = 4096, nbytes = 8 gives n = 2^19 multipliers,
    m = 12 288 commitments, q = 1 060 864 constraints, and a bit-valued (a_L, a_R) witness with a_O = 0."""
    nb = 8 * nbytes
    hi = (1 << nb) - 1 if hi is None else hi
    rng = np.random.default_rng(seed)
    if values is None:
        values = [int.from_bytes(rng.bytes(nbytes), "little") % (hi - lo + 1) + lo for _ in range(count)]
    m1 = L_ORDER - 1
    # ---- template of one value: (kind, index, coeff) per term, rows lengths
    kinds, idxs, coefs, lens = [], [], [], []

    def row(terms):
        for k, i, c in terms:
            kinds.append(k); idxs.append(i); coefs.append(c % L_ORDER)
        lens.append(len(terms))

    row([(3, 1, 1), (3, 2, 1), (4, 0, -(hi - lo))])
    for which in (0, 1):
        j0 = which * nb
        for i in range(nb):
            row([(2, j0 + i, 1)])
            row([(0, j0 + i, 1), (1, j0 + i, 1), (4, 0, m1)])
        row([(3, 1 + which, 1)] + [(1, j0 + i, -(1 << i)) for i in range(nb)])
    kinds = np.array(kinds, dtype=np.uint32)
    idxs = np.array(idxs, dtype=np.uint32)
    coef_b = np.frombuffer(b"".join(int(c).to_bytes(32, "little") for c in coefs), dtype=np.uint8).reshape(-1, 32)
    lens = np.array(lens, dtype=np.uint32)
    k = np.arange(count, dtype=np.uint32)[:, None]
    add = np.where(kinds < 3, 2 * nb, np.where(kinds == 3, 3, 0)).astype(np.uint32)[None, :]
    tv = ((kinds[None, :] << 29) | (idxs[None, :] + k * add)).astype(np.uint32).reshape(-1)
    tc = np.tile(coef_b, (count, 1))
    all_lens = np.tile(lens, count)
    row_ptr = np.zeros(len(all_lens) + 1, dtype=np.uint32)
    np.cumsum(all_lens, out=row_ptr[1:])
    # ---- witness
    vals, bits = [], np.zeros((count, 2 * nb), dtype=np.uint8)
    for c, v in enumerate(values):
        a, b = (v - lo) % L_ORDER, (hi - v) % L_ORDER
        vals += [v % L_ORDER, a, b]
        ab, bb = a.to_bytes(32, "little"), b.to_bytes(32, "little")  # range_proof reads the raw scalar bytes (utils.rs:12-18)
        for i in range(nb):
            bits[c, i] = (ab[i // 8] >> (i % 8)) & 1
            bits[c, nb + i] = (bb[i // 8] >> (i % 8)) & 1
    n = count * 2 * nb
    aR = np.zeros((n, 32), dtype=np.uint8)
    aR[:, 0] = bits.reshape(-1)
    aL = np.zeros((n, 32), dtype=np.uint8)
    aL[:, 0] = 1 - bits.reshape(-1)
    aO = np.zeros((n, 32), dtype=np.uint8)
    blinds = [int.from_bytes(rng.bytes(64), "little") % L_ORDER for _ in vals]
    return dict(label=label, n=n, m=3 * count, vals=_enc(vals), blinds=_enc(blinds), aL=aL.tobytes(), aR=aR.tobytes(), aO=aO.tobytes(),
                csr=(row_ptr, np.ascontiguousarray(tv), np.ascontiguousarray(tc)), values=values)


class Circuit:
    """device-resident constraint matrix (bpg_circuit) built from numpy CSR arrays"""

    def __init__(self, ctx, n, m, csr):
        import ctypes as C
        row_ptr, tv, tc = csr
        row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint32)
        tv = np.ascontiguousarray(tv, dtype=np.uint32)
        if isinstance(tc, (bytes, bytearray)):
            tcb = bytes(tc)
        else:
            tc = np.ascontiguousarray(tc, dtype=np.uint8)
            tcb = C.cast(tc.ctypes.data, C.c_char_p)  # no copy: the library reads it during the call only
        self.ctx, self.n, self.m = ctx, n, m
        h = C.c_void_p()
        ctx.check(ctx.lib.bpg_circuit_create(ctx.h, n, m, len(row_ptr) - 1, row_ptr.ctypes.data_as(C.POINTER(C.c_uint32)),
                                             tv.ctypes.data_as(C.POINTER(C.c_uint32)) if len(tv) else None, tcb, C.byref(h)))
        self.h = h

    def prove(self, inst, ext_rng32, flags=0):
        import ctypes as C
        cap = 1 + 32 * (14 + 64 + 2)
        proof, V = C.create_string_buffer(cap), C.create_string_buffer(32 * max(1, self.m))
        rc = self.ctx.lib.bpg_r1cs_prove(self.ctx.h, self.h, inst["label"], len(inst["label"]), inst["aL"], inst["aR"], inst["aO"], inst["vals"],
                                         inst["blinds"], ext_rng32, flags, V, proof, cap)
        if rc < 0:
            self.ctx.check(rc)
        return proof.raw[:rc], V.raw[:32 * self.m]

    def prefetch(self, inst, ext_rng32, flags=0):
        """start the transcript-RNG stream of a FUTURE prove(inst, ext_rng32) in the background (bpg_r1cs_prove_prefetch)"""
        self.ctx.check(self.ctx.lib.bpg_r1cs_prove_prefetch(self.ctx.h, self.h, inst["label"], len(inst["label"]), inst["vals"], inst["blinds"], ext_rng32, flags))

    def verify(self, label, V, proof, ext_rng32=None, flags=0):
        """ext_rng32 stands for the verifier's thread_rng draw: fresh secret randomness unless a test pins it"""
        import ctypes as C
        if ext_rng32 is None:
            ext_rng32 = os.urandom(32)
        acc = C.c_int(0)
        self.ctx.check(self.ctx.lib.bpg_r1cs_verify(self.ctx.h, self.h, label, len(label), V, proof, len(proof), ext_rng32, flags, C.byref(acc)))
        return bool(acc.value)

    def close(self):
        if self.h:
            self.ctx.lib.bpg_circuit_destroy(self.h)
            self.h = None
