#!/usr/bin/env python3
"""Extract the reference's own known answers for the hot path into small JSON fixtures.

Run in the dev container (needs /root/reference):  python tests/golden/make_golden.py
Writes tests/golden/mimc_consts.json and tests/golden/mimc_kats.json.  These are DATA the
reference pins; no reference source code is copied.

  mimc_consts.json : the 486 MiMC round constants, 32-byte LE hex   (src/mimc_hash/mimc_consts.rs:2-489)
  mimc_kats.json   : {"hash": [[preimage_hex, digest_be_hex, source], ...],
                      "node": [[left_be_hex, right_be_hex, digest_be_hex, source], ...]}
       hash = mimc_hash(bytes) (padded sponge, src/mimc_hash/mimc.rs:61-75)
       node = unpadded 2-block sponge over from_bits(be_to_scalar(.)) (merkle_tree_gadget.rs:106)
"""
import json
import os
import re

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def read(p):
    with open(os.path.join(REF, p)) as f:
        return f.read()


def byte_lists(text):
    """every `[0x.., 0x.., ...]` / vec![..] literal in order, as bytes."""
    out = []
    for m in re.finditer(r"\[((?:\s*0x[0-9a-fA-F]{2}\s*,?)+)\s*\]", text):
        out.append(bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", m.group(1))))
    return out


def assignments(path):
    d = {}
    for line in read(path).splitlines():
        m = re.match(r"\s*([WI]\d+)\s*=\s*0x([0-9a-fA-F]*)", line)
        if m:
            d[m.group(1)] = m.group(2)
    return d


def copy_fixtures():
    """the reference's CLI fixtures (data: .gadgets/.inst/.wtns triples of tests/resources and example.*) -> tests/golden/fixtures/"""
    import shutil
    dst = os.path.join(HERE, "fixtures")
    os.makedirs(dst, exist_ok=True)
    for f in sorted(os.listdir(os.path.join(REF, "tests", "resources"))):
        shutil.copy(os.path.join(REF, "tests", "resources", f), os.path.join(dst, f))
    for ext in (".gadgets", ".inst", ".wtns"):
        shutil.copy(os.path.join(REF, "example" + ext), os.path.join(dst, "example" + ext))


def main():
    copy_fixtures()
    consts = byte_lists(read("src/mimc_hash/mimc_consts.rs"))
    assert len(consts) == 486 and all(len(c) == 32 for c in consts)
    with open(os.path.join(HERE, "mimc_consts.json"), "w") as f:
        json.dump([c.hex() for c in consts], f, indent=0)

    hashes, nodes = [], []
    # mimc.rs:104-143 : (preimage, image) x 2
    bl = byte_lists(read("src/mimc_hash/mimc.rs"))
    assert len(bl) == 4
    hashes.append([bl[0].hex(), bl[1].hex(), "src/mimc_hash/mimc.rs:106-121"])
    hashes.append([bl[2].hex(), bl[3].hex(), "src/mimc_hash/mimc.rs:126-142"])
    # mimc_hash_gadget.rs tests: 3 x (preimage, image)
    bl = byte_lists(read("src/mimc_hash/mimc_hash_gadget.rs").split("#[cfg(test)]")[1])
    assert len(bl) == 6
    for k in range(3):
        hashes.append([bl[2 * k].hex(), bl[2 * k + 1].hex(), "src/mimc_hash/mimc_hash_gadget.rs test %d" % (k + 1)])
    # tests/resources/mimc_hash.* : HASH image preimage
    a = {**assignments("tests/resources/mimc_hash.inst"), **assignments("tests/resources/mimc_hash.wtns")}
    for line in read("tests/resources/mimc_hash.gadgets").split():
        pass
    for line in read("tests/resources/mimc_hash.gadgets").splitlines():
        p = line.split()
        if len(p) == 3 and p[0] == "HASH":
            hashes.append([a[p[2]], a[p[1]], "tests/resources/mimc_hash.gadgets: " + line.strip()])
    # combine_gadgets.rs:34-68
    bl = byte_lists(read("tests/combine_gadgets.rs"))
    w2, root, leaf = bl[0], bl[1], bl[2]
    hashes.append(["43", w2.hex(), "tests/combine_gadgets.rs:31-40 (W2 = H(0x43))"])
    nodes.append([w2.hex(), leaf.hex(), root.hex(), "tests/combine_gadgets.rs:55-68"])
    # example.* : HASH W2 W1
    a = {**assignments("example.inst"), **assignments("example.wtns")}
    hashes.append([a["W1"], a["W2"], "example.gadgets: HASH W2 W1"])

    # merkle_tree_gadget.rs:126-215 : W1..W15 heap-ordered, W_p = node(W_2p, W_2p+1)
    src = read("src/merkle_tree/merkle_tree_gadget.rs")
    W = {}
    for m in re.finditer(r"const W(\d+): \[u8; 32\] = \[(.*?)\];", src, re.S):
        W[int(m.group(1))] = bytes(int(x, 16) for x in re.findall(r"0x([0-9a-fA-F]{2})", m.group(2)))
    assert sorted(W) == list(range(1, 16))
    for p in range(1, 8):
        nodes.append([W[2 * p].hex(), W[2 * p + 1].hex(), W[p].hex(),
                      "src/merkle_tree/merkle_tree_gadget.rs:126-215 W%d=node(W%d,W%d)" % (p, 2 * p, 2 * p + 1)])
    # merkle_tree_gadget.rs:476-503 : chain h1=node(W1,W1), h_{k+1}=node(h_k,h_k)
    test512 = src.split("fn test_merkle_tree_gadget_512")[1]
    chain = [c for c in byte_lists(test512) if len(c) == 32][1:10]  # [0] is the root literal
    assert len(chain) == 9
    prev = W[1]
    for k, h in enumerate(chain):
        nodes.append([prev.hex(), prev.hex(), h.hex(),
                      "src/merkle_tree/merkle_tree_gadget.rs:476-503 level %d" % (2 << k)])
        prev = h

    with open(os.path.join(HERE, "mimc_kats.json"), "w") as f:
        json.dump({"hash": hashes, "node": nodes}, f, indent=1)

    # fixtures where leaves are hashed first (prover.rs:324-334): node(H(l), H(r)) / nested
    fx = []
    for stem in ("tests/resources/merkle_tree", "example"):
        a = {**assignments(stem + ".inst"), **assignments(stem + ".wtns")}
        for line in read(stem + ".gadgets").splitlines():
            if line.startswith("MERKLE"):
                root, pat = line.split(None, 2)[1:]
                fx.append({"root": a[root], "pattern": pat.strip(),
                           "values": {k: v for k, v in a.items() if re.search(r"\b%s\b" % k, pat)},
                           "source": stem + ".gadgets: " + line.strip()})
    with open(os.path.join(HERE, "merkle_fixtures.json"), "w") as f:
        json.dump(fx, f, indent=1)
    print("hash KATs:", len(hashes), "node KATs:", len(nodes), "merkle fixtures:", len(fx))


if __name__ == "__main__":
    main()
