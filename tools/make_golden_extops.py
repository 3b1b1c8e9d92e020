#!/usr/bin/env python3
"""tests/golden/<field>_ext_ops.npz from the UNMODIFIED reference CPU backend (oracle/_ref/<field> built with EXT_FIELD=ON): every
`<field>_extension_*` vec-op the frontend exports (icicle/src/vec_ops.cpp: REGISTER_*_EXT_FIELD_BACKEND family,
icicle/include/icicle/backend/vec_ops_backend.h:297-494) on seeded inputs, incl. batch / columns_batch for the scalar-vector and
reduction ops, zero elements for inv / div, and the quartic extension's Montgomery conversion."""
import ctypes as C
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_icicle
import common

name = sys.argv[1] if len(sys.argv) > 1 else "babybear"
r = ref_icicle.get(name)
f = r.field
P = lambda a: a.ctypes.data_as(C.c_void_p)
n, batch = 48, 3
a = common.seeded_scalars(name, 4 * n * batch, 11).reshape(n * batch, 4)
b = common.seeded_scalars(name, 4 * n * batch, 12).reshape(n * batch, 4)
b[5] = 0        # a zero extension element: inverse(0) = 0
a[7, 1:] = 0    # an element of the base field embedded in the extension
s = common.seeded_scalars(name, n * batch, 13)
out = dict(a=a, b=b, s=s)


def call(sym, *args):
    rc = getattr(f, f"{name}_extension_{sym}")(*args)
    assert rc == 0, (sym, rc)


def cfg(**kw):
    return r.vec_config(**kw)


for op in ("vector_add", "vector_sub", "vector_mul", "vector_div"):
    o = np.zeros_like(a)
    c = cfg()
    call(op, P(a), P(b), C.c_uint64(n * batch), C.byref(c), P(o))
    out[op] = o
acc = a.copy()
c = cfg()
call("vector_accumulate", P(acc), P(b), C.c_uint64(n * batch), C.byref(c))
out["vector_accumulate"] = acc
o = np.zeros_like(a)
c = cfg()
call("vector_inv", P(b), C.c_uint64(n * batch), C.byref(c), P(o))
out["vector_inv"] = o
o = np.zeros_like(a)
c = cfg()
call("vector_mixed_mul", P(a), P(s), C.c_uint64(n * batch), C.byref(c), P(o))
out["vector_mixed_mul"] = o
for columns in (False, True):
    tag = "cols" if columns else "rows"
    for op in ("scalar_add_vec", "scalar_sub_vec", "scalar_mul_vec"):
        o = np.zeros_like(b)
        c = cfg(batch_size=batch, columns_batch=columns)
        call(op, P(a[:batch].copy()), P(b), C.c_uint64(n), C.byref(c), P(o))
        out[f"{op}_{tag}"] = o
    for op in ("vector_sum", "vector_product"):
        o = np.zeros((batch, 4), dtype=np.uint32)
        c = cfg(batch_size=batch, columns_batch=columns)
        call(op, P(a), C.c_uint64(n), C.byref(c), P(o))
        out[f"{op}_{tag}"] = o
for into in (True, False):
    o = np.zeros_like(a)
    c = cfg()
    call("scalar_convert_montgomery", P(a), C.c_uint64(n * batch), C.c_bool(into), C.byref(c), P(o))
    out[f"convert_montgomery_{int(into)}"] = o
o = np.zeros((32, 4), dtype=np.uint32)
c = cfg()
call("bit_reverse", P(a[:32].copy()), C.c_uint64(32), C.byref(c), P(o))
out["bit_reverse"] = o
o = np.zeros((6 * 8, 4), dtype=np.uint32)
c = cfg()
call("matrix_transpose", P(a[:48].copy()), C.c_uint32(6), C.c_uint32(8), C.byref(c), P(o))
out["matrix_transpose_6x8"] = o
o = np.zeros((10, 4), dtype=np.uint32)
c = cfg()
call("slice", P(a[:48].copy()), C.c_uint64(3), C.c_uint64(4), C.c_uint64(48), C.c_uint64(10), C.byref(c), P(o))
out["slice_3_4_10"] = o
path = os.path.join(ROOT, "tests", "golden", f"{name}_ext_ops.npz")
np.savez_compressed(path, **out)
print("wrote", path, sorted(out))
