#!/usr/bin/env python3
"""Generate tests/golden/bn254_ecntt.npz from the UNMODIFIED reference CPU backend (oracle/_ref/bn254 built with ECNTT=ON:
`bn254_ecntt`, icicle/src/ecntt.cpp:8-12 -> cpu_ecntt.cpp): inputs (projective, Z = 1, one point at infinity) and the outputs
normalised to affine by the reference's own `bn254_to_affine`, forward / inverse, coset, kNN / kNR / kRN, batch 2.

    python tools/make_golden_ecntt.py
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_icicle
import common


def main():
    r = ref_icicle.get("bn254")
    dom_log = 8
    root = r.get_root_of_unity(1 << dom_log)
    r.ntt_init_domain(root)
    n, batch = 16, 2
    aff = common.gen_g1_points("bn254", n * batch, 2024)
    proj = common.affine_to_projective_limbs(aff, 8)
    proj[3] = common.affine_to_projective_limbs(np.zeros((1, 16), dtype=np.uint32), 8)[0]   # the point at infinity (0,1,0)
    g = np.array([0x1234567, 0x89abcde, 3, 0, 0, 0, 0, 0], dtype=np.uint32)
    out = {"ntt_root": root, "dom_log": np.array([dom_log]), "input_projective": proj, "coset": g}
    for d in (0, 1):
        for o in (0, 1, 2):
            for c in (0, 1):
                y = r.ecntt(proj, n, d, coset_gen=g if c else None, batch_size=batch, ordering=o)
                out[f"d{d}_o{o}_g{c}_affine"] = np.stack([r.to_affine(p) for p in y])
    y = r.ecntt(proj, n, 0, batch_size=batch, columns_batch=True)
    out["d0_cols_affine"] = np.stack([r.to_affine(p) for p in y])
    r.ntt_release_domain()
    path = os.path.join(ROOT, "tests", "golden", "bn254_ecntt.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, len(out), "entries", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
