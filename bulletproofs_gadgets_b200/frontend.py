"""Front-end driver: the reference's `.gadgets` mini-language, `.inst` / `.wtns` / `.coms` /
`.proof` files, its eight gadgets, the record/replay constraint buffer and the OR conjunction, restated on the host so
that every fixture of the reference (tests/resources/*, example.*) runs prover -> .coms/.proof -> verifier on the GPU
path without a Rust toolchain.  In the drop-in this layer stays Rust; nothing here is on the device hot path.

Mirrors (file:line under /root/reference):
  src/bin/prover.rs:47-100 (main), 102-117 (assign_buffer), 160-199 (hash_witness / hash_instance), 202-251 (OR), 253-532
  src/bin/verifier.rs:46-101, 160-454
  src/lalrpop/gadget_grammar.lalrpop, var_grammar.lalrpop, assignment_parser.rs:129-220 (.coms naming C<w>-<i>, D<g>-<s>-<i>)
  src/cs_buffer.rs (Operation log, rewind), src/or/or_conjunction.rs:4-67
  src/utils.rs:5-35 (range_proof), src/bounds_check, equality, inequality, less_than, set_membership, mimc_hash, merkle_tree
Scalar semantics follow SURVEY App. C: witness / instance values enter through Scalar::from_bits (raw, possibly >= l);
`==`, Inequality::compare and range_proof read the raw bytes; arithmetic results are reduced.
"""
import os
import random
import secrets
import re

from .api import L_ORDER, ONE, Prover, Verifier, BulletproofGens, R1CSError
from . import gadgets as _g

L = L_ORDER


# ----------------------------------------------------------------------------- scalars / linear combinations
def be_to_scalars(data):
    return _g.be_to_scalars(bytes(data))


def be_to_scalar(data):
    assert len(data) <= 32, "the given vector is longer than 32 bytes"
    return be_to_scalars(bytes(data) or b"\x00")[0]


def scalar_to_be(s):
    return int(s).to_bytes(32, "little")[::-1]


def lc_var(v):
    return [(v, 1)]


def lc_const(s):
    return [(ONE, s % L)]


def lc_neg(a):
    return [(v, (-c) % L) for v, c in a]


def lc_sub(a, b):
    return list(a) + lc_neg(b)


def lc_scale(a, s):
    return [(v, c * s % L) for v, c in a]


# ----------------------------------------------------------------------------- constraint-system buffer (cs_buffer.rs)
class _Counter:
    def __init__(self):
        self.n = 0


class Buffer:
    """ProverBuffer / VerifierBuffer: records Multiply / AllocateMultiplier / Constrain operations.  Multiplier numbers come
    from one counter shared by every buffer of a run, which is exactly what the reference obtains by initialising each
    buffer's dummy prover from the operations recorded before it (cs_buffer.rs:49-72, prover.rs:67-72, 213-218)."""

    def __init__(self, counter, is_prover):
        self.ops, self.cache, self.counter, self.is_prover = [], [], counter, is_prover

    def multiply(self, left, right):
        i = self.counter.n
        self.counter.n += 1
        self.ops.append(("mul", list(left), list(right), i))
        return ("L", i), ("R", i), ("O", i)

    def allocate_multiplier(self, assignment=None):
        if self.is_prover and assignment is None:
            raise R1CSError("MissingAssignment")
        i = self.counter.n
        self.counter.n += 1
        self.ops.append(("alloc", assignment if self.is_prover else None, None, i))
        return ("L", i), ("R", i), ("O", i)

    def allocate(self, assignment=None):
        raise R1CSError("GadgetError: call to unimplemented method allocate")

    def constrain(self, lc):
        self.ops.append(("con", list(lc), None, None))

    def rewind(self):
        self.cache.append(self.ops)
        self.ops = []


def or_conjunction(main, buf):
    """or() of or_conjunction.rs:4-38: clause multipliers are replayed, clause constraints are multiplied together over
    the Cartesian product of the clauses."""
    constraints_vec = []
    for ops in buf.cache:
        cons = []
        for op in ops:
            if op[0] == "con":
                cons.append(op[1])
            else:
                main.ops.append(op)  # already numbered by the shared counter
        constraints_vec.append(cons)
    if not constraints_vec:
        return
    combos = [[c] for c in constraints_vec[0]]
    for lst in constraints_vec[1:]:
        combos = [xs + [y] for xs in combos for y in lst]
    for cons in combos:
        prod = cons[0]
        for c in cons[1:]:
            _, _, o = main.multiply(prod, c)
            prod = lc_var(o)
        main.constrain(prod)


def assign_buffer(cs, buf):
    """replay the recorded operations into the real Prover / Verifier (prover.rs:102-117)"""
    for kind, a, b, i in buf.ops:
        if kind == "mul":
            l, _, _ = cs.multiply(a, b)
            assert l[1] == i, "multiplier numbering diverged"
        elif kind == "alloc":
            l, _, _ = cs.allocate_multiplier(a)
            assert l[1] == i, "multiplier numbering diverged"
        else:
            cs.constrain(a)


# ----------------------------------------------------------------------------- gadgets (assemble + preprocess)
def range_proof(cs, x_lc, n, x_assignment):
    """utils.rs:5-35"""
    exp2 = 1
    x = list(x_lc)
    for i in range(n):
        assign = None
        if x_assignment is not None:
            bit = (int(x_assignment) >> i) & 1
            assign = (1 - bit, bit)
        a, b, o = cs.allocate_multiplier(assign)
        cs.constrain(lc_var(o))
        cs.constrain([(a, 1), (b, 1), (ONE, L - 1)])
        x = x + [(b, (-exp2) % L)]
        exp2 = exp2 * 2 % L
    cs.constrain(x)


class BoundsCheck:
    def __init__(self, min_bytes, max_bytes):
        self.n = (len(max_bytes) * 8) & 0xFF  # `as u8`
        self.min, self.max = be_to_scalar(min_bytes), be_to_scalar(max_bytes)

    def preprocess(self, witnesses):
        v = witnesses[0]
        return [(v - self.min) % L, (self.max - v) % L]

    def assemble(self, cs, _witness_vars, derived):
        (a_val, a), (b_val, b) = derived[0], derived[1]
        cs.constrain([(a, 1), (b, 1), (ONE, (-(self.max - self.min)) % L)])
        range_proof(cs, lc_var(a), self.n, a_val)
        range_proof(cs, lc_var(b), self.n, b_val)


class Equality:
    def __init__(self, right_lcs):
        self.right = right_lcs

    def preprocess(self, _):
        return []

    def assemble(self, cs, left_vars, _derived):
        if len(self.right) != len(left_vars):
            return cs.constrain(lc_const(1))
        for r, l in zip(self.right, left_vars):
            cs.constrain(lc_sub(r, lc_var(l)))


def _raw_ge(left, right):
    """Inequality::compare (inequality_gadget.rs:103-113): big-endian comparison of the raw scalar bytes, ties -> true"""
    return int(left) >= int(right)


class Inequality:
    def __init__(self, right_lcs, right_assignment=None):
        self.right, self.right_assignment = right_lcs, right_assignment

    def preprocess(self, left_hand):
        assert self.right_assignment is not None, "missing right hand assignment"
        out, total = [], 0
        for i, left in enumerate(left_hand):
            right = self.right_assignment[i] if i < len(self.right_assignment) else 0
            delta = (left - right) % L if _raw_ge(left, right) else (right - left) % L
            out.append(delta)
            if delta == 0:
                out.append(0)
            else:
                inv = pow(delta, L - 2, L)
                out.append(inv)
                total = (total + delta * inv) % L
        out.append(pow(total, L - 2, L))
        return out

    def assemble(self, cs, left_vars, derived):
        if len(self.right) != len(left_vars):
            return cs.constrain(lc_const(0))
        total = lc_const(0)
        for i, lv in enumerate(left_vars):
            right_lc, left_lc = self.right[i], lc_var(lv)
            delta, delta_inv = derived[2 * i][1], derived[2 * i + 1][1]
            left = lc_sub(lc_sub(left_lc, right_lc), lc_var(delta))
            right = lc_sub(lc_sub(right_lc, left_lc), lc_var(delta))
            _, _, zero = cs.multiply(left, right)
            cs.constrain(lc_var(zero))
            _, _, zero_or_one = cs.multiply(lc_var(delta), lc_var(delta_inv))
            total = total + lc_var(zero_or_one)
        sum_inv = derived[-1][1]
        _, _, one = cs.multiply(total, lc_var(sum_inv))
        cs.constrain(lc_sub(lc_const(1), lc_var(one)))


class LessThan:
    def __init__(self, left_lc, left_val, right_lc, right_val):
        self.left, self.left_val, self.right, self.right_val = left_lc, left_val, right_lc, right_val

    def preprocess(self, _):
        assert self.left_val is not None and self.right_val is not None, "missing assignment"
        delta = (self.right_val - self.left_val) % L
        return [delta, 0 if delta == 0 else pow(delta, L - 2, L)]

    def assemble(self, cs, _w, derived):
        (delta_val, delta), (_, delta_inv) = derived[0], derived[1]
        n = 126
        range_proof(cs, self.left, n, self.left_val)
        range_proof(cs, self.right, n, self.right_val)
        range_proof(cs, lc_var(delta), n, delta_val)
        _, _, one = cs.multiply(lc_var(delta), lc_var(delta_inv))
        cs.constrain(lc_sub(lc_const(1), lc_var(one)))
        cs.constrain(lc_sub(lc_sub(self.right, self.left), lc_var(delta)))


class SetMembership:
    def __init__(self, value_lc, value, instance_lcs, instance_vals):
        self.value_lc, self.value, self.instance_lcs, self.instance_vals = value_lc, value, instance_lcs, instance_vals

    def preprocess(self, witnesses):
        assert self.value is not None and self.instance_vals is not None, "missing assignments"
        return [1 if int(e) == int(self.value) else 0 for e in list(witnesses) + list(self.instance_vals)]  # raw `==`

    def assemble(self, cs, witness_vars, derived):
        one_hot = []
        for _, bit in derived:
            b = lc_var(bit)
            _, _, zero = cs.multiply(lc_sub(lc_const(1), b), b)
            cs.constrain(lc_var(zero))
            one_hot.append(b)
        total = lc_const(0)
        for b in one_hot:
            total = total + b
        cs.constrain(lc_sub(lc_const(1), total))
        elements = [lc_var(w) for w in witness_vars] + list(self.instance_lcs)
        if len(one_hot) != len(elements):
            return cs.constrain(lc_const(1))
        actual = lc_const(0)
        for b, e in zip(one_hot, elements):
            _, _, p = cs.multiply(b, e)
            actual = actual + lc_var(p)
        cs.constrain(lc_sub(self.value_lc, actual))


class MimcHash256:
    def __init__(self, image_lc):
        self.image = image_lc

    def preprocess(self, witnesses):
        return _g.mimc_preprocess([int(w) for w in witnesses])

    def assemble(self, cs, witness_vars, derived):
        _g.mimc_hash_gadget_wire(cs, list(witness_vars), [d[1] for d in derived], self.image)


class MerkleTree256:
    def __init__(self, root_lc, instance_lcs, witness_lcs, pattern):
        self.root, self.instance_lcs, self.witness_lcs, self.pattern = root_lc, instance_lcs, witness_lcs, pattern

    def preprocess(self, _):
        return []

    def assemble(self, cs, _w, _d):
        _g.merkle_wire(cs, self.root, self.pattern, self.witness_lcs, self.instance_lcs)


# ----------------------------------------------------------------------------- parsing
_ASSIGN = re.compile(r"^\s*([WICD][\d]+(?:-[\d]+){0,2})\s*=\s*0[xX]([0-9a-fA-F]*)\s*$")


def parse_assignments(text):
    out = []
    for line in text.splitlines():
        if not line.strip():
            continue
        m = _ASSIGN.match(line)
        if not m:
            raise ValueError("cannot parse assignment line: %r" % line)
        out.append((m.group(1), bytes.fromhex(m.group(2))))
    return out


def parse_merkle_tree(tokens):
    """Tree rule of gadget_grammar.lalrpop:46-73 -> (instance names, witness names, pattern) in left-to-right order"""
    pos = [0]
    inst, wtns = [], []

    def node():
        t = tokens[pos[0]]
        if t == "(":
            pos[0] += 1
            l = node()
            r = node()
            assert tokens[pos[0]] == ")", "expected )"
            pos[0] += 1
            return (l, r)
        pos[0] += 1
        if t.startswith("W"):
            wtns.append(t)
            return "W"
        inst.append(t)
        return "I"

    pat = node()
    assert pos[0] == len(tokens) and pat not in ("W", "I"), "malformed MERKLE pattern"
    return inst, wtns, pat


class Assignments:
    def __init__(self):
        self.inst, self.wtns, self.coms, self.derived_cache = {}, {}, {}, []

    def instance(self, name, max32=False):
        v = self.inst[name]
        if max32:
            assert len(v) <= 32, "instance var %s is longer than 32 bytes" % name
        return v

    def witness(self, name, single=False):
        w = self.wtns[name]
        if single:
            assert len(w[0]) == 1, "witness var %s is longer than 32 bytes" % name
        return w

    def all_commitments(self, name):
        out, i = [], 0
        while "C%s-%d" % (name[1:], i) in self.coms:
            out.append(self.coms["C%s-%d" % (name[1:], i)])
            i += 1
        return out

    def commitment(self, name, i):
        return self.coms["C%s-%d" % (name[1:], i)]

    def derived(self, gadget, index, sub):
        return self.coms["D%d-%d-%d" % (gadget, sub, index)]

    def inquire_derived(self, gadget, index, sub):
        return self.coms.get("D%d-%d-%d" % (gadget, sub, index))


# ----------------------------------------------------------------------------- prover driver
class ProverRun:
    def __init__(self, label, gadgets_text, inst_text, wtns_text, test_seed=None, ctx=None):
        """Every Pedersen blinding is an independent draw from the OS CSPRNG, as the reference's
        Scalar::random(&mut thread_rng()) (commitments.rs:27,39  gadget.rs:31).  `test_seed` (tests / benchmarks ONLY)
        replaces it with a seeded, reproducible stream so proof bytes can be compared with the oracle: such
        commitments are NOT hiding."""
        self.prover = Prover.new(label, ctx=ctx)
        self._test_rng = random.Random(test_seed) if test_seed is not None else None
        self.a = Assignments()
        self.coms_names = []  # name of every committed variable, in commit order
        self.counter = _Counter()
        for name, val in parse_assignments(inst_text):
            self.a.inst[name] = val
        for name, val in parse_assignments(wtns_text):
            scalars = be_to_scalars(val or b"\x00")
            vars_ = []
            for i, s in enumerate(scalars):
                vars_.append(self._commit(s, "C%s-%d" % (name[1:], i)))
            self.a.wtns[name] = (scalars, vars_, val)
        self.lines = gadgets_text.splitlines()
        self.pos = 0
        top = Buffer(self.counter, True)
        while self.pos < len(self.lines):
            index, line = self.pos, self.lines[self.pos]
            self.pos += 1
            self._conjunction(line, top)
            self._gadget(line, top, index)
        assign_buffer(self.prover, top)

    def _commit(self, scalar, name):
        blinding = self._test_rng.randrange(L) if self._test_rng is not None else secrets.randbelow(L)
        _, var = self.prover.commit(scalar, blinding)
        self.coms_names.append(name)
        return var

    def _setup(self, gadget, witnesses, index, sub):
        """Gadget::setup (gadget.rs:18-38) + parse_derived_wtns naming"""
        derived = []
        for i, s in enumerate(gadget.preprocess(witnesses)):
            derived.append((s, self._commit(s, "D%d-%d-%d" % (index, sub, i))))
        return derived

    def _op(self, line):
        toks = line.split()
        return toks[0] if toks else ""

    def _conjunction(self, line, buf):
        if self._op(line) != "OR":
            return
        orb = Buffer(self.counter, True)
        if self.pos >= len(self.lines):
            raise ValueError("unexpected end of input")
        while self.pos < len(self.lines):
            index, ln = self.pos, self.lines[self.pos]
            self.pos += 1
            op = self._op(ln)
            if op == "]":
                break
            if op == "}":
                orb.rewind()
            else:
                self._conjunction(ln, orb)
                self._gadget(ln, orb, index)
        or_conjunction(buf, orb)

    def _hash_witness(self, buf, name, index, sub):
        """prover.rs:160-190: commit the image, prove MimcHash256(preimage) = image; .coms D<index>-<sub>-0.."""
        scalars, vars_, data = self.a.witness(name)
        image = _g.mimc_sponge_int(_g.mimc_preprocess_blocks(scalars))
        image_var = self._commit(image, "D%d-%d-0" % (index, sub))
        g = MimcHash256(lc_var(image_var))
        derived = []
        for i, s in enumerate(g.preprocess(scalars)):
            derived.append((s, self._commit(s, "D%d-%d-%d" % (index, sub, i + 1))))
        g.assemble(buf, vars_, derived)
        return image, image_var

    def _hash_instance(self, name):
        scalars = be_to_scalars(self.a.instance(name) or b"\x00")
        image = _g.mimc_sponge_int(_g.mimc_preprocess_blocks(scalars))
        return image, lc_const(image)

    def _gadget(self, line, buf, index):
        toks = line.split()
        op = toks[0] if toks else ""
        a = self.a
        if op == "BOUND":
            var, lo, hi = toks[1:4]
            w = a.witness(var, single=True)
            g = BoundsCheck(a.instance(lo, True), a.instance(hi, True))
            g.assemble(buf, w[1], self._setup(g, w[0], index, 0))
        elif op == "HASH":
            image, pre = toks[1:3]
            image_lc = lc_var(a.witness(image, single=True)[1][0]) if image[0] == "W" else lc_const(be_to_scalar(a.instance(image, True)))
            w = a.witness(pre)
            g = MimcHash256(image_lc)
            g.assemble(buf, w[1], self._setup(g, w[0], index, 0))
        elif op == "MERKLE":
            root = toks[1]
            inst, wtns, pat = parse_merkle_tree(re.findall(r"[()]|[WI]\d+", " ".join(toks[2:])))
            root_lc = lc_var(a.witness(root, single=True)[1][0]) if root[0] == "W" else lc_const(be_to_scalar(a.instance(root, True)))
            inst_lcs = [self._hash_instance(n)[1] for n in inst]
            wit_lcs = [lc_var(self._hash_witness(buf, n, index, k)[1]) for k, n in enumerate(wtns)]
            MerkleTree256(root_lc, inst_lcs, wit_lcs, pat).assemble(buf, [], [])
        elif op in ("EQUALS", "UNEQUAL"):
            left, right = toks[1:3]
            if left[0] == "I":  # grammar: (Instance, Witness) is normalised to (Witness(right), Instance(left))
                left, right = right, left
            lw = a.witness(left)
            if right[0] == "W":
                rw = a.witness(right)
                r_vals, r_lcs = rw[0], [lc_var(v) for v in rw[1]]
            else:
                r_vals = be_to_scalars(a.instance(right) or b"\x00")
                r_lcs = [lc_const(s) for s in r_vals]
            if op == "EQUALS":
                Equality(r_lcs).assemble(buf, lw[1], [])
            else:
                g = Inequality(r_lcs, r_vals)
                g.assemble(buf, lw[1], self._setup(g, lw[0], index, 0))
        elif op == "LESS_THAN":
            lw, rw = a.witness(toks[1], single=True), a.witness(toks[2], single=True)
            g = LessThan(lc_var(lw[1][0]), lw[0][0], lc_var(rw[1][0]), rw[0][0])
            g.assemble(buf, [], self._setup(g, [], index, 0))
        elif op == "SET_MEMBER":
            self._set_member(toks[1], toks[2:], buf, index)
        # "OR", "[", "{", "}", "]" and blank lines: nothing

    def _set_member(self, member, elements, buf, index):
        a = self.a
        if member[0] == "W":
            mw = a.witness(member)
            m_scalars, m_lcs = mw[0], [lc_var(v) for v in mw[1]]
        else:
            m_scalars = be_to_scalars(a.instance(member) or b"\x00")
            m_lcs = [lc_const(s) for s in m_scalars]
        m_val, m_lc = m_scalars[0], m_lcs[0]
        hashing = len(m_scalars) > 1
        w_vars, w_vals, i_lcs, i_vals = [], [], [], []
        if not hashing:
            for e in elements:
                if e[0] == "W":
                    ew = a.witness(e)
                    if len(ew[1]) == 1:
                        w_vals.append(ew[0][0]); w_vars.append(ew[1][0])
                    else:
                        hashing = True
                else:
                    es = be_to_scalars(a.instance(e) or b"\x00")
                    if len(es) == 1:
                        i_vals.append(es[0]); i_lcs.append(lc_const(es[0]))
                    else:
                        hashing = True
        if hashing:
            sub = 1
            if member[0] == "W":
                m_val, v = self._hash_witness(buf, member, index, sub)
                m_lc = lc_var(v)
                sub += 1
            else:
                m_val, m_lc = self._hash_instance(member)
            w_vars, w_vals, i_lcs, i_vals = [], [], [], []
            for e in elements:
                if e[0] == "W":
                    val, v = self._hash_witness(buf, e, index, sub)
                    sub += 1
                    w_vars.append(v); w_vals.append(val)
                else:
                    val, lc = self._hash_instance(e)
                    i_lcs.append(lc); i_vals.append(val)
        g = SetMembership(m_lc, m_val, i_lcs, i_vals)
        g.assemble(buf, w_vars, self._setup(g, w_vals, index, 0))

    def finish(self, ext_rng32=None, flags=0):
        """-> (.coms text, .proof bytes, number of constraints)"""
        n = self.prover.get_num_multiplications()
        cap = 1
        while cap < n:
            cap *= 2
        gens = BulletproofGens.new(cap, 1, ctx=self.prover.ctx)
        proof, V = self.prover.prove(gens, ext_rng32=ext_rng32, flags=flags)
        coms = "".join("%s = 0x%s\n" % (nm, c.hex()) for nm, c in zip(self.coms_names, V))
        return coms, proof, self.prover.num_constraints()


# ----------------------------------------------------------------------------- verifier driver
class VerifierRun:
    def __init__(self, label, gadgets_text, inst_text, coms_text, ctx=None):
        self.verifier = Verifier.new(label, ctx=ctx)
        self.a = Assignments()
        self.counter = _Counter()
        for name, val in parse_assignments(inst_text):
            self.a.inst[name] = val
        for name, val in parse_assignments(coms_text):
            self.a.coms[name] = self.verifier.commit(val.rjust(32, b"\0") if len(val) < 32 else val)
        self.lines = gadgets_text.splitlines()
        self.pos = 0
        top = Buffer(self.counter, False)
        while self.pos < len(self.lines):
            index, line = self.pos, self.lines[self.pos]
            self.pos += 1
            self._conjunction(line, top)
            self._gadget(line, top, index)
        assign_buffer(self.verifier, top)

    _op = ProverRun._op

    def _conjunction(self, line, buf):
        if self._op(line) != "OR":
            return
        orb = Buffer(self.counter, False)
        if self.pos >= len(self.lines):
            raise ValueError("unexpected end of input")
        while self.pos < len(self.lines):
            index, ln = self.pos, self.lines[self.pos]
            self.pos += 1
            op = self._op(ln)
            if op == "]":
                break
            if op == "}":
                orb.rewind()
            else:
                self._conjunction(ln, orb)
                self._gadget(ln, orb, index)
        or_conjunction(buf, orb)

    def _hash_witness(self, buf, name, index, sub):
        a = self.a
        pre = a.all_commitments(name)
        image = a.derived(index, 0, sub)
        d = [a.derived(index, 1, sub)]
        d2 = a.inquire_derived(index, 2, sub)
        if d2 is not None:
            d.append(d2)
        MimcHash256(lc_var(image)).assemble(buf, pre, [(None, x) for x in d])
        return image

    def _hash_instance(self, name):
        scalars = be_to_scalars(self.a.instance(name) or b"\x00")
        return lc_const(_g.mimc_sponge_int(_g.mimc_preprocess_blocks(scalars)))

    def _gadget(self, line, buf, index):
        toks = line.split()
        op = toks[0] if toks else ""
        a = self.a
        if op == "BOUND":
            var, lo, hi = toks[1:4]
            g = BoundsCheck(a.instance(lo, True), a.instance(hi, True))
            g.assemble(buf, [a.commitment(var, 0)], [(None, a.derived(index, 0, 0)), (None, a.derived(index, 1, 0))])
        elif op == "HASH":
            image, pre = toks[1:3]
            image_lc = lc_var(a.commitment(image, 0)) if image[0] == "W" else lc_const(be_to_scalar(a.instance(image, True)))
            d = [a.derived(index, 0, 0)]
            d2 = a.inquire_derived(index, 1, 0)
            if d2 is not None:
                d.append(d2)
            MimcHash256(image_lc).assemble(buf, a.all_commitments(pre), [(None, x) for x in d])
        elif op == "MERKLE":
            root = toks[1]
            inst, wtns, pat = parse_merkle_tree(re.findall(r"[()]|[WI]\d+", " ".join(toks[2:])))
            root_lc = lc_var(a.commitment(root, 0)) if root[0] == "W" else lc_const(be_to_scalar(a.instance(root, True)))
            inst_lcs = [self._hash_instance(n) for n in inst]
            wit_lcs = [lc_var(self._hash_witness(buf, n, index, k)) for k, n in enumerate(wtns)]
            MerkleTree256(root_lc, inst_lcs, wit_lcs, pat).assemble(buf, [], [])
        elif op in ("EQUALS", "UNEQUAL"):
            left, right = toks[1:3]
            if left[0] == "I":
                left, right = right, left
            lv = a.all_commitments(left)
            r_lcs = [lc_var(v) for v in a.all_commitments(right)] if right[0] == "W" else [lc_const(s) for s in be_to_scalars(a.instance(right) or b"\x00")]
            if op == "EQUALS":
                Equality(r_lcs).assemble(buf, lv, [])
            else:
                d = [(None, a.derived(index, i, 0)) for i in range(2 * len(lv) + 1)]
                Inequality(r_lcs).assemble(buf, lv, d)
        elif op == "LESS_THAN":
            l, r = a.commitment(toks[1], 0), a.commitment(toks[2], 0)
            LessThan(lc_var(l), None, lc_var(r), None).assemble(buf, [], [(None, a.derived(index, 0, 0)), (None, a.derived(index, 1, 0))])
        elif op == "SET_MEMBER":
            self._set_member(toks[1], toks[2:], buf, index)

    def _set_member(self, member, elements, buf, index):
        a = self.a
        m_lcs = [lc_var(v) for v in a.all_commitments(member)] if member[0] == "W" else [lc_const(s) for s in be_to_scalars(a.instance(member) or b"\x00")]
        m_lc = m_lcs[0]
        hashing = False
        w_vars, i_lcs = [], []
        for e in elements:
            if e[0] == "W":
                w = a.all_commitments(e)
                if len(w) == 1:
                    w_vars.append(w[0])
                else:
                    hashing = True
            else:
                es = be_to_scalars(a.instance(e) or b"\x00")
                if len(es) == 1:
                    i_lcs.append(lc_const(es[0]))
                else:
                    hashing = True
        if len(m_lcs) > 1:
            hashing = True
        derived = [(None, a.derived(index, k, 0)) for k in range(len(elements))]
        if hashing:
            sub = 1
            if member[0] == "W":
                m_lc = lc_var(self._hash_witness(buf, member, index, sub))
                sub += 1
            else:
                m_lc = self._hash_instance(member)
            w_vars, i_lcs = [], []
            for e in elements:
                if e[0] == "W":
                    w_vars.append(self._hash_witness(buf, e, index, sub))
                    sub += 1
                else:
                    i_lcs.append(self._hash_instance(e))
        SetMembership(m_lc, None, i_lcs, None).assemble(buf, w_vars, derived)

    def finish(self, proof, ext_rng32=None, flags=0):
        """-> True / False like the verifier binary prints (verifier.rs:89-100)"""
        n = self.verifier.get_num_vars()
        cap = 1
        while cap < n:
            cap *= 2
        gens = BulletproofGens.new(cap, 1, ctx=self.verifier.ctx)
        try:
            self.verifier.verify(proof, None, gens, ext_rng32=ext_rng32, flags=flags)
            return True
        except R1CSError:
            return False


# ----------------------------------------------------------------------------- file-level entry points (the two binaries)
def _read(path):
    with open(path) as f:
        return f.read()


def prover_main(stem, test_seed=None, ext_rng32=None, ctx=None, label=None):
    """`prover <stem>`: reads <stem>.gadgets/.inst/.wtns, writes <stem>.coms and <stem>.proof; returns #constraints.
    Blindings come from the OS CSPRNG unless `test_seed` is given (tests only, see ProverRun)."""
    run = ProverRun((label or stem).encode() if isinstance(label or stem, str) else (label or stem), _read(stem + ".gadgets"), _read(stem + ".inst"),
                    _read(stem + ".wtns"), test_seed=test_seed, ctx=ctx)
    coms, proof, nc = run.finish(ext_rng32=ext_rng32)
    with open(stem + ".coms", "w") as f:
        f.write(coms)
    with open(stem + ".proof", "wb") as f:
        f.write(proof)
    return nc


def verifier_main(stem, ctx=None, label=None):
    """`verifier <stem>`: reads <stem>.gadgets/.inst/.coms/.proof, returns True / False"""
    with open(stem + ".proof", "rb") as f:
        proof = f.read()
    run = VerifierRun((label or stem).encode() if isinstance(label or stem, str) else (label or stem), _read(stem + ".gadgets"), _read(stem + ".inst"),
                      _read(stem + ".coms"), ctx=ctx)
    return run.finish(proof)
