#!/usr/bin/env python
"""Generates tests/golden/*.json from the Python big-int oracle (oracle/hades_ref.py).

The reference is Rust and cannot be executed in this image (no rustc/cargo), so these vectors are
NOT outputs of the reference binary: they are outputs of the restatement, cross-checked in
tests/test_oracle.py against the independent survey-time vectors and against the
C restatement.  Re-run:  python tools/gen_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import hades_ref as H  # noqa: E402

P = H.P


def limbs_hex(v):
    return ["0x%016x" % l for l in H.to_mont_limbs(v)]


def case(name, vals):
    out = H.perm(vals)
    return {"name": name, "width": len(vals), "input": [hex(v) for v in vals], "output": [hex(v) for v in out],
            "input_mont_limbs": [limbs_hex(v) for v in vals], "output_mont_limbs": [limbs_hex(v) for v in out]}


def main():
    perm_cases = [
        case("readme_example_ones", [1] * 5),            # README.md:50-65
        case("hades_det_17", [17] * 5),                  # scalar.rs:62-74
        case("hades_det_19", [19] * 5),
        case("preimage_constant_5000", [5000] * 5),      # gadget.rs:225-244
        case("iota", [0, 1, 2, 3, 4]),
        case("zeros", [0] * 5),
        case("p_minus_1", [P - 1] * 5),
        case("mixed_edges", [0, 1, P - 1, H.R, H.R2]),
        case("high_bits", [P - 2, (1 << 254) + 12345, (1 << 255) % P, 0xFFFFFFFF, (1 << 64) - 1]),
        case("w3_123", [1, 2, 3]),
        case("w9_1to9", list(range(1, 10))),
    ]
    # seeded synthetic states exactly as bench/tests generate them (Montgomery limbs given directly)
    for i in range(8):
        vals = []
        for j in range(5):
            limbs = [H.synth_limb(H.SEED, i * 5 + j, l) for l in range(4)]
            vals.append(H.from_mont_limbs(limbs))
        c = case(f"synthetic_state_{i}", vals)
        perm_cases.append(c)
    merkle = [{"leaves": n, "leaf_values": "0..n-1", "root": hex(H.merkle_root(list(range(n)))),
               "root_mont_limbs": limbs_hex(H.merkle_root(list(range(n))))} for n in (1, 4, 16, 64)]
    msgs = [[], [1], [1, 2, 3], [1, 2, 3, 4], [1, 2, 3, 4, 5], list(range(1, 10)), [P - 1] * 8, [0] * 4]
    sponge = [{"message": [hex(x) for x in m], "digest": hex(H.sponge(m)), "digest_mont_limbs": limbs_hex(H.sponge(m))}
              for m in msgs]
    # sponge with domain separation: capacity word = tag (hades_ref.sponge(message, domain))
    tags = [1, 0xF, 1 << 32, (1 << 64) + 3, P - 1]
    sponge_ds = [{"message": [hex(x) for x in m], "domain": hex(t), "domain_mont_limbs": limbs_hex(t),
                  "digest": hex(H.sponge(m, t)), "digest_mont_limbs": limbs_hex(H.sponge(m, t))}
                 for t in tags for m in ([], [1, 2, 3, 4], list(range(1, 10)))]
    out = {"generator": "tools/gen_golden.py (oracle/hades_ref.py)", "modulus": hex(P),
           "ark_bin_sha256": H.ARK_BIN_SHA256, "mds_bin_sha256": H.MDS_BIN_SHA256,
           "perm": perm_cases, "merkle": merkle, "sponge": sponge, "sponge_ds": sponge_ds}
    path = os.path.join(ROOT, "tests", "golden", "hades252_kat.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
