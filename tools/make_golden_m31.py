#!/usr/bin/env python3
"""Generate tests/golden/m31.npz from the UNMODIFIED reference CPU backend (make -C oracle ref FIELD=m31 ID=1003 HAS_NTT=0):
Mersenne-31 vec-ops (the reference's MersenneField, icicle/include/icicle/fields/stark_fields/m31.h) incl. the edge values
0, 1, p-1 and the identity Montgomery conversion.

    python tools/make_golden_m31.py
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT)
import ref_icicle

r = ref_icicle.get("m31")
p = 0x7fffffff
n = 4096
a, b = r.generate_scalars(n), r.generate_scalars(n)
edge = np.array([0, 1, p - 1, 2, p - 2, 0x40000000, 0x3fffffff, 0], dtype=np.uint32).reshape(-1, 1)
a[:8] = edge
b[:8] = edge[::-1]
b[8:16] = edge
a[8:16] = edge
out = dict(vec_a=a, vec_b=b)
for op in ("vector_add", "vector_sub", "vector_mul", "vector_accumulate"):
    out[op] = r.vec2(op, a, b, n)
out["to_montgomery"] = r.scalar_convert_montgomery(a, n, True)
out["from_montgomery"] = r.scalar_convert_montgomery(a, n, False)
out["bit_reverse"] = r.bit_reverse(a, n)
path = os.path.join(ROOT, "tests", "golden", "m31.npz")
np.savez_compressed(path, **out)
print("wrote", path, {k: v.shape for k, v in out.items()}, os.path.getsize(path))
