#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the UNMODIFIED reference CPU backend (oracle/_ref, built from /root/reference).
Run in the build container (needs oracle/_ref/<name>); the fixtures are committed so the GPU box and CPU-only runs can
check the oracle port and the CUDA path without the reference tree.  Inputs come from the reference's own generators
(scalar_t::rand_host_many, projective_t::rand_host_many) seeded by its default mt19937 state."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import ref_icicle
from icicle_b200 import utils

name = sys.argv[1] if len(sys.argv) > 1 else "bn254"
t = ref_icicle.TARGETS[name]
r = ref_icicle.get(name)
out = {}
if t["curve"]:
    n = 96
    s = r.generate_scalars(n)
    P = r.generate_affine_points(n)
    P[7] = 0  # an affine zero
    out.update(msm_scalars=s, msm_points=P)
    res = r.msm(s, P, n)
    out["msm_result_affine"] = r.to_affine(res[0])
    for bits in (1, 17, 100):
        out[f"msm_bitsize{bits}_affine"] = r.to_affine(r.msm(s, P, n, bitsize=bits)[0])
    out["msm_batch3_affine"] = np.stack([r.to_affine(x) for x in r.msm(s, P[:32], 32, batch_size=3)])
    sm = r.scalar_convert_montgomery(s, n, True)
    out["scalars_montgomery"] = sm
    out["points_montgomery"] = r.affine_convert_montgomery(P, n, True)
    if t["g2"]:
        P2 = r.generate_affine_points(24, g2=True)
        out["g2_points"] = P2
        out["g2_msm_result_affine"] = r.to_affine(r.msm(s[:24], P2, 24, g2=True)[0], g2=True)
fname = {"bn254": "bn254_fr", "bls12_381": "bls12_381_fr", "bls12_377": "bls12_377_fr", "bw6_761": "bls12_377_fq", "grumpkin": None}.get(name, name)
if fname:
    logn = 6
    root = r.get_root_of_unity(1 << (logn + 2))
    r.ntt_init_domain(root)
    x = r.generate_scalars(2 << logn)
    out.update(ntt_root=root, ntt_input=x)
    g_dom = r.get_root_of_unity_from_domain(logn + 2)
    g_arb = x[5].copy()
    for d in (0, 1):
        for o in range(4):
            out[f"ntt_d{d}_o{o}"] = r.ntt(x[: 1 << logn], 1 << logn, d, ordering=o)
        out[f"ntt_d{d}_coset_dom"] = r.ntt(x[: 1 << logn], 1 << logn, d, coset_gen=g_dom)
        out[f"ntt_d{d}_coset_arb"] = r.ntt(x[: 1 << logn], 1 << logn, d, coset_gen=g_arb)
        out[f"ntt_d{d}_batch2_cols"] = r.ntt(x, 1 << logn, d, batch_size=2, columns_batch=True)
    out["coset_arb"] = g_arb
    out["coset_dom"] = g_dom
    r.ntt_release_domain()
a, b = r.generate_scalars(50), r.generate_scalars(50)
out.update(vec_a=a, vec_b=b, vec_add=r.vec2("vector_add", a, b, 50), vec_sub=r.vec2("vector_sub", a, b, 50), vec_mul=r.vec2("vector_mul", a, b, 50))
path = os.path.join(ROOT, "tests", "golden", f"{name}.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items()})
