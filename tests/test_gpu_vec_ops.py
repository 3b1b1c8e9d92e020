"""GPU parity: vec-ops through the C ABI vs (a) Python integers, (b) the reference CPU backend (oracle/_ref).
Mirrors icicle/tests/test_mod_arithmetic_api.h:55-284 (vectorVectorOps, montgomeryConversion, scalarVectorOps, bitReverse,
Slice) and test_curve_api.cpp:231-271 (MontConversion)."""
import numpy as np
import pytest

import icicle_b200 as ib
from icicle_b200 import utils
import common

pytestmark = pytest.mark.gpu

FIELDS = [(ib.Field.BN254_FR, "bn254_fr"), (ib.Field.BN254_FQ, "bn254_fq"), (ib.Field.BLS12_381_FR, "bls12_381_fr"),
          (ib.Field.BLS12_381_FQ, "bls12_381_fq"), (ib.Field.BLS12_377_FR, "bls12_377_fr"), (ib.Field.BLS12_377_FQ, "bls12_377_fq"),
          (ib.Field.BW6_761_FQ, "bw6_761_fq"), (ib.Field.STARK252, "stark252"), (ib.Field.BABYBEAR, "babybear"),
          (ib.Field.KOALABEAR, "koalabear")]


@pytest.mark.parametrize("field,name", FIELDS)
def test_vector_ops_vs_python(field, name):
    fp = utils.field_params(name)
    p, L = fp["p"], fp["limbs"]
    n = 1000 - 37  # ragged
    a = common.rand_field_elems(name, n, 1, as_ints=True)
    b = common.rand_field_elems(name, n, 2, as_ints=True)
    a[0], b[0] = 0, 0
    a[1], b[1] = p - 1, p - 1
    a[2], b[2] = p - 1, 1
    A, B = utils.to_limbs(a, L), utils.to_limbs(b, L)
    assert utils.from_limbs(ib.vector_add(field, A, B, n)) == [(x + y) % p for x, y in zip(a, b)]
    assert utils.from_limbs(ib.vector_sub(field, A, B, n)) == [(x - y) % p for x, y in zip(a, b)]
    assert utils.from_limbs(ib.vector_mul(field, A, B, n)) == [(x * y) % p for x, y in zip(a, b)]
    acc = A.copy()
    ib.vector_accumulate(field, acc, B, n)
    assert utils.from_limbs(acc) == [(x + y) % p for x, y in zip(a, b)]
    R = 1 << (32 * L)
    assert utils.from_limbs(ib.convert_montgomery(field, A, n, True)) == [x * R % p for x in a]
    assert utils.from_limbs(ib.convert_montgomery(field, A, n, False)) == [x * pow(R, -1, p) % p for x in a]


@pytest.mark.parametrize("field,name", FIELDS[:1] + FIELDS[3:4] + FIELDS[8:9])
def test_inv_div_sum_product(field, name):
    """vector_inv / vector_div / vector_sum / vector_product (cpu_vec_ops.cpp:386-403,428-490); inverse(0) = 0."""
    fp = utils.field_params(name)
    p, L = fp["p"], fp["limbs"]
    n = 1000 + 3
    a = common.rand_field_elems(name, n, 21, as_ints=True)
    b = common.rand_field_elems(name, n, 22, as_ints=True)
    a[5], a[6], b[7] = 0, 1, 0
    A, B = utils.to_limbs(a, L), utils.to_limbs(b, L)
    inv = lambda x: pow(x, -1, p) if x else 0
    assert utils.from_limbs(ib.vector_inv(field, A, n)) == [inv(x) for x in a]
    assert utils.from_limbs(ib.vector_div(field, A, B, n)) == [x * inv(y) % p for x, y in zip(a, b)]
    size, batch = 333, 3
    for columns in (False, True):
        cfg = ib.VecOpsConfig(batch_size=batch, columns_batch=columns)
        rows = [[a[(i * batch + bb) if columns else (bb * size + i)] for i in range(size)] for bb in range(batch)]
        assert utils.from_limbs(ib.vector_sum(field, A[: size * batch], size, cfg)) == [sum(r) % p for r in rows]
        prods = []
        for r in rows:
            acc = 1
            for x in r:
                acc = acc * x % p
            prods.append(acc)
        assert utils.from_limbs(ib.vector_product(field, A[: size * batch], size, cfg)) == prods


@pytest.mark.parametrize("field,name", FIELDS[:1] + FIELDS[8:9])
@pytest.mark.parametrize("columns", [False, True])
def test_scalar_vec_ops_batch(field, name, columns):
    fp = utils.field_params(name)
    p, L = fp["p"], fp["limbs"]
    size, batch = 257, 3
    s = common.rand_field_elems(name, batch, 3, as_ints=True)
    v = common.rand_field_elems(name, size * batch, 4, as_ints=True)
    S, V = utils.to_limbs(s, L), utils.to_limbs(v, L)
    cfg = lambda: ib.VecOpsConfig(batch_size=batch, columns_batch=columns)
    bidx = (lambda t: t % batch) if columns else (lambda t: t // size)
    for fn, op in ((ib.scalar_add_vec, lambda x, y: x + y), (ib.scalar_sub_vec, lambda x, y: x - y), (ib.scalar_mul_vec, lambda x, y: x * y)):
        got = utils.from_limbs(fn(field, S, V, size, cfg()))
        assert got == [op(s[bidx(t)], v[t]) % p for t in range(size * batch)]


def test_device_resident_and_async():
    import torch
    field, name = FIELDS[0]
    fp = utils.field_params(name)
    p, L = fp["p"], fp["limbs"]
    n = 4096
    a = common.rand_field_elems(name, n, 5, as_ints=True)
    b = common.rand_field_elems(name, n, 6, as_ints=True)
    dA, dB = ib.to_device(utils.to_limbs(a, L)), ib.to_device(utils.to_limbs(b, L))
    stream = torch.cuda.Stream()
    out = ib.device_empty(n * L).view(n, L)
    ib.vector_mul(field, dA, dB, n, ib.VecOpsConfig(stream=stream, is_async=True), out)
    stream.synchronize()
    assert utils.from_limbs(ib.to_host(out)) == [(x * y) % p for x, y in zip(a, b)]


@pytest.mark.parametrize("columns", [False, True])
def test_bit_reverse_transpose_slice(columns):
    field, name = FIELDS[0]
    L = 8
    logn, batch = 9, 2
    n = 1 << logn
    v = common.rand_field_elems(name, n * batch, 7, as_ints=True)
    V = utils.to_limbs(v, L)
    got = utils.from_limbs(ib.bit_reverse(field, V, n, ib.VecOpsConfig(batch_size=batch, columns_batch=columns)))
    if columns:
        exp = [v[common.bitrev(t // batch, logn) * batch + (t % batch)] for t in range(n * batch)]
    else:
        exp = [v[(t // n) * n + common.bitrev(t % n, logn)] for t in range(n * batch)]
    assert got == exp
    # in-place on device
    import torch
    d = ib.to_device(V)
    ib.bit_reverse(field, d, n, ib.VecOpsConfig(batch_size=batch, columns_batch=columns), d)
    assert utils.from_limbs(ib.to_host(d)) == exp
    if not columns:
        rows, cols = 37, 55
        m = common.rand_field_elems(name, rows * cols, 8, as_ints=True)
        T = utils.from_limbs(ib.matrix_transpose(field, utils.to_limbs(m, L), rows, cols))
        assert T == [m[r * cols + c] for c in range(cols) for r in range(rows)]
        sl = utils.from_limbs(ib.slice(field, utils.to_limbs(m, L), 3, 7, rows * cols, 100))
        assert sl == [m[3 + 7 * i] for i in range(100)]


def test_vs_reference_cpu_backend():
    ref_icicle = pytest.importorskip("ref_icicle")
    if not ref_icicle.available("bn254"):
        pytest.skip("oracle/_ref/bn254 not built")
    r = ref_icicle.get("bn254")
    n = 1 << 12
    a, b = r.generate_scalars(n), r.generate_scalars(n)
    F = ib.Field.BN254_FR
    assert np.array_equal(ib.vector_add(F, a, b, n), r.vec2("vector_add", a, b, n))
    assert np.array_equal(ib.vector_sub(F, a, b, n), r.vec2("vector_sub", a, b, n))
    assert np.array_equal(ib.vector_mul(F, a, b, n), r.vec2("vector_mul", a, b, n))
    assert np.array_equal(ib.convert_montgomery(F, a, n, True), r.scalar_convert_montgomery(a, n, True))
    assert np.array_equal(ib.convert_montgomery(F, a, n, False), r.scalar_convert_montgomery(a, n, False))
    assert np.array_equal(ib.bit_reverse(F, a, n), r.bit_reverse(a, n))
    assert np.array_equal(ib.matrix_transpose(F, a, 64, 64), r.matrix_transpose(a, 64, 64))
    pts = r.generate_affine_points(300)
    assert np.array_equal(ib.affine_convert_montgomery(ib.Curve.BN254_G1, pts, 300, True), r.affine_convert_montgomery(pts, 300, True))
    g2 = r.generate_affine_points(50, g2=True)
    assert np.array_equal(ib.affine_convert_montgomery(ib.Curve.BN254_G2, g2, 50, True), r.affine_convert_montgomery(g2, 50, True, g2=True))


def test_highest_idx_poly_eval_division():
    field, name = FIELDS[0]
    fp = utils.field_params(name)
    p, L = fp["p"], fp["limbs"]
    a = common.rand_field_elems(name, 300, 31, as_ints=True)
    a[250:] = [0] * 50
    assert list(ib.highest_non_zero_idx(field, utils.to_limbs(a, L), 300)) == [249]
    assert list(ib.highest_non_zero_idx(field, utils.to_limbs([0] * 10, L), 10)) == [-1]
    coeffs = common.rand_field_elems(name, 40, 32, as_ints=True)
    dom = common.rand_field_elems(name, 9, 33, as_ints=True)
    ev = utils.from_limbs(ib.poly_eval(field, utils.to_limbs(coeffs, L), 40, utils.to_limbs(dom, L), 9))
    assert ev == [sum(c * pow(x, i, p) for i, c in enumerate(coeffs)) % p for x in dom]
    # division: num = q*den + r with deg r < deg den
    den = common.rand_field_elems(name, 12, 34, as_ints=True)
    q_true = common.rand_field_elems(name, 20, 35, as_ints=True)
    r_true = common.rand_field_elems(name, 11, 36, as_ints=True)
    num = [0] * 31
    for i, x in enumerate(q_true):
        for j, y in enumerate(den):
            num[i + j] = (num[i + j] + x * y) % p
    for i, x in enumerate(r_true):
        num[i] = (num[i] + x) % p
    q, rr = ib.poly_division(field, utils.to_limbs(num, L), 31, utils.to_limbs(den, L), 12, 20, 31)
    assert utils.from_limbs(q) == q_true
    assert utils.from_limbs(rr)[:11] == r_true and not any(utils.from_limbs(rr)[11:])
