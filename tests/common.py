"""Shared test helpers: seeded synthetic inputs and Python-integer group/field checks (test infrastructure only)."""
import random

import numpy as np

from icicle_b200 import utils

CURVE_FIELDS = {  # curve name -> (scalar field, base field, g2?)
    "bn254": ("bn254_fr", "bn254_fq"), "bls12_381": ("bls12_381_fr", "bls12_381_fq"), "bls12_377": ("bls12_377_fr", "bls12_377_fq"),
    "bw6_761": ("bls12_377_fq", "bw6_761_fq"), "grumpkin": ("bn254_fq", "bn254_fr"),
}


def rand_field_elems(field_name, n, seed, as_ints=False):
    fp = utils.field_params(field_name)
    rng = random.Random(seed)
    vals = [rng.randrange(fp["p"]) for _ in range(n)]
    return vals if as_ints else utils.to_limbs(vals, fp["limbs"])


def seeded_scalars(field_name, n, seed):
    """n uniform-looking field elements as (n, limbs) uint32, fast enough for 2^26+: random 32-bit limbs with the top limb
    reduced below the modulus' top limb (so every value is < p).  Used identically by the GPU tests and tests/ref_worker.py."""
    fp = utils.field_params(field_name)
    L, p = fp["limbs"], fp["p"]
    rs = np.random.default_rng(seed)
    if L == 1:
        return rs.integers(0, p, size=(n, 1), dtype=np.uint32)
    out = rs.integers(0, 1 << 32, size=(n, L), dtype=np.uint32)
    top = p >> (32 * (L - 1))
    out[:, L - 1] %= np.uint32(top)
    return out


def tiled_g1_points(curve_name, n, distinct, seed):
    """n affine points = `distinct` DISTINCT curve points (gen_g1_points) tiled; (n, 2*limbs) uint32."""
    base = gen_g1_points(curve_name, min(n, distinct), seed)
    reps = (n + base.shape[0] - 1) // base.shape[0]
    return np.ascontiguousarray(np.tile(base, (reps, 1))[:n])


def fq2_projective_to_affine_ints(proj_limbs, limbs, q, nr):
    """(X, Y, Z) over Fq2 (each {c0, c1}) -> [x0, x1, y0, y1] python ints; None at infinity."""
    X0, X1, Y0, Y1, Z0, Z1 = utils.from_limbs(np.asarray(proj_limbs, dtype=np.uint32).reshape(6, limbs))
    if Z0 == 0 and Z1 == 0:
        return None
    den = pow((Z0 * Z0 - nr * Z1 * Z1) % q, -1, q)
    zi = (Z0 * den % q, (-Z1) * den % q)
    mul = lambda a, b: ((a[0] * b[0] + nr * a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)
    x, y = mul((X0, X1), zi), mul((Y0, Y1), zi)
    return [x[0], x[1], y[0], y[1]]


# ---- G1 affine arithmetic over python ints ------------------------------------------------------------------------------
def ec_add(A, B, q):
    if A is None: return B
    if B is None: return A
    x1, y1 = A; x2, y2 = B
    if x1 == x2:
        if (y1 + y2) % q == 0: return None
        lam = 3 * x1 * x1 * pow(2 * y1, -1, q) % q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, q) % q
    x3 = (lam * lam - x1 - x2) % q
    return (x3, (lam * (x1 - x3) - y1) % q)


def ec_mul(k, A, q):
    R = None
    while k:
        if k & 1: R = ec_add(R, A, q)
        A = ec_add(A, A, q)
        k >>= 1
    return R


def gen_g1_points(curve_name, n, seed, as_ints=False):
    """n DISTINCT affine points: P_0 = k0*G, P_{i+1} = P_i + D (D = k1*G).  Returns (n, 2*limbs) uint32 limbs."""
    cp = utils.curve_params(curve_name)
    fq = utils.field_params(cp["fq"])
    fr = utils.field_params(cp["fr"])
    q = fq["p"]
    rng = random.Random(seed)
    G = (cp["gx"], cp["gy"])
    P = ec_mul(rng.randrange(1, fr["p"]), G, q)
    D = ec_mul(rng.randrange(1, fr["p"]), G, q)
    pts = []
    for _ in range(n):
        pts.append(P)
        P = ec_add(P, D, q)
    if as_ints:
        return pts
    flat = []
    for (x, y) in pts:
        flat += [x, y]
    return utils.to_limbs(flat, fq["limbs"]).reshape(n, 2 * fq["limbs"])


def affine_limbs_to_ints(arr, limbs):
    v = utils.from_limbs(np.asarray(arr, dtype=np.uint32).reshape(-1, limbs))
    out = []
    for i in range(0, len(v), 2):
        out.append(None if (v[i] == 0 and v[i + 1] == 0) else (v[i], v[i + 1]))
    return out


def msm_naive_ints(scalars, points, q):
    acc = None
    for s, P in zip(scalars, points):
        if P is None or s == 0: continue
        acc = ec_add(acc, ec_mul(s, P, q), q)
    return acc


def projective_to_affine_ints(proj_limbs, limbs, q):
    X, Y, Z = utils.from_limbs(np.asarray(proj_limbs, dtype=np.uint32).reshape(3, limbs))
    if Z == 0:
        return None
    zi = pow(Z, -1, q)
    return (X * zi % q, Y * zi % q)


def is_projective_zero(proj_limbs, limbs):
    X, Y, Z = utils.from_limbs(np.asarray(proj_limbs, dtype=np.uint32).reshape(3, limbs))
    return X == 0 and Z == 0 and Y != 0


def bitrev(i, logn):
    return int(format(i, "0%db" % logn)[::-1], 2) if logn else 0


def ntt_naive_ints(x, w, p, inverse=False, coset=1):
    """Reference definition (SURVEY 8a / ntt_cpu.h:69-232): fwd out[k] = sum (x[i] g^i) w^(ik); inv out[i] = g^-i N^-1 sum x[k] w^(-ik)."""
    n = len(x)
    if not inverse:
        xs = [x[i] * pow(coset, i, p) % p for i in range(n)]
        return [sum(xs[i] * pow(w, i * k, p) for i in range(n)) % p for k in range(n)]
    wi = pow(w, -1, p)
    ninv = pow(n, -1, p)
    gi = pow(coset, -1, p)
    return [ninv * pow(gi, i, p) * sum(x[k] * pow(wi, i * k, p) for k in range(n)) % p for i in range(n)]


def affine_to_projective_limbs(aff, limbs):
    """(n, 2*limbs) affine points -> (n, 3*limbs) homogeneous projective (x, y, 1); affine zero (0,0) -> (0, 1, 0)
    (icicle/include/icicle/curves/projective.h:26-31)."""
    aff = np.asarray(aff, dtype=np.uint32).reshape(-1, 2 * limbs)
    out = np.zeros((aff.shape[0], 3 * limbs), dtype=np.uint32)
    out[:, : 2 * limbs] = aff
    out[:, 2 * limbs] = 1
    zero = ~aff.any(axis=1)
    out[zero, 2 * limbs] = 0
    out[zero, limbs] = 1
    return out
