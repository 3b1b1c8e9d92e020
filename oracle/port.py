"""TEST INFRASTRUCTURE ONLY -- Python front of the C restatement oracle/port/oracle_port.c (liboracle_port.so).

The port restates the reference's algorithms (Barrett field arithmetic, RCB projective group law, Pippenger with signed
digits, radix-2 NTT with coset/scaling) on the CPU; tests/test_oracle.py pins it against the unmodified reference CPU
backend (oracle/_ref) and the committed golden vectors.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may
import it.  Nothing under icicle_b200/ does.
"""
import ctypes as C
import json
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PARAMS = json.load(open(os.path.join(_HERE, "..", "icicle_b200", "params.json")))
MAXL = 24


class PortField(C.Structure):
    _fields_ = [("n", C.c_int), ("bits", C.c_int), ("p", C.c_uint32 * MAXL), ("m", C.c_uint32 * MAXL), ("b3", C.c_uint32 * MAXL)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle_port.so")
        if not os.path.exists(path):
            raise ImportError("oracle/liboracle_port.so not built (make -C oracle port)")
        _lib = C.CDLL(path)
    return _lib


def _limbs(v, n):
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(n)]


def make_field(name, b=0):
    f = _PARAMS["fields"][name]
    p, n, bits = int(f["p"], 16), f["limbs"], f["bits"]
    pf = PortField()
    pf.n, pf.bits = n, bits
    for i, v in enumerate(_limbs(p, n)): pf.p[i] = v
    for i, v in enumerate(_limbs((1 << (2 * bits)) // p, n)): pf.m[i] = v  # params_gen.h get_m
    for i, v in enumerate(_limbs(3 * b % p, n)): pf.b3[i] = v
    return pf, p, n


def _arr(ints, n):
    return np.array([_limbs(v, n) for v in ints], dtype=np.uint32).reshape(len(ints), n)


def _ints(arr, n):
    a = np.asarray(arr, dtype=np.uint32).reshape(-1, n)
    return [sum(int(a[i, j]) << (32 * j) for j in range(n)) for i in range(a.shape[0])]


def field_op(name, op, a_ints, b_ints):
    pf, p, n = make_field(name)
    A, B = _arr(a_ints, n), _arr(b_ints, n)
    out = np.zeros_like(A)
    lib().port_vec_op(C.byref(pf), {"add": 0, "sub": 1, "mul": 2}[op], A.ctypes.data_as(C.c_void_p), B.ctypes.data_as(C.c_void_p),
                      C.c_uint64(len(a_ints)), out.ctypes.data_as(C.c_void_p))
    return _ints(out, n)


def ntt(xs, w, p, inverse=False, coset=1, field_name=None):
    """xs: python ints in natural order; w: root of unity of order len(xs).  Returns python ints (natural order)."""
    if field_name is None:
        field_name = next(k for k, v in _PARAMS["fields"].items() if int(v["p"], 16) == p)
    pf, p, n = make_field(field_name)
    N = len(xs)
    logn = N.bit_length() - 1
    data = _arr(xs, n).copy()
    w_dir = pow(w, -1, p) if inverse else w
    g = None
    if coset != 1:
        g = pow(coset, -1, p) if inverse else coset
    W = _arr([w_dir], n)
    NI = _arr([pow(N, -1, p)], n)
    G = _arr([g], n) if g is not None else None
    lib().port_ntt(C.byref(pf), data.ctypes.data_as(C.c_void_p), C.c_int(logn), W.ctypes.data_as(C.c_void_p),
                   NI.ctypes.data_as(C.c_void_p) if inverse else None, G.ctypes.data_as(C.c_void_p) if G is not None else None,
                   C.c_int(1 if inverse else 0))
    return _ints(data, n)


def msm(curve_name, scalars, points, c=8, bitsize=0):
    """scalars: python ints; points: list of (x, y) or None (affine zero).  Returns affine (x, y) or None (group zero)."""
    cv = _PARAMS["curves"][curve_name]
    fq_name, fr_name = cv["fq"], cv["fr"]
    b = int(cv["b"], 16)
    pf, q, L = make_field(fq_name, b)
    fr = _PARAMS["fields"][fr_name]
    r, sl, sbits = int(fr["p"], 16), fr["limbs"], fr["bits"]
    S = _arr(scalars, sl)
    flat = []
    for P in points:
        flat += [0, 0] if P is None else [P[0], P[1]]
    B = _arr(flat, L)
    R = _arr([r], sl)
    out = np.zeros(3 * L, dtype=np.uint32)
    lib().port_msm(C.byref(pf), S.ctypes.data_as(C.c_void_p), C.c_int(sl), C.c_int(sbits), R.ctypes.data_as(C.c_void_p),
                   B.ctypes.data_as(C.c_void_p), C.c_int(len(scalars)), C.c_int(c), C.c_int(bitsize), out.ctypes.data_as(C.c_void_p))
    X, Y, Z = _ints(out, L)
    if Z == 0:
        return None
    zi = pow(Z, -1, q)
    return (X * zi % q, Y * zi % q)
