"""GPU parity against the committed golden vectors (tests/golden/*.npz, generated from the unmodified reference CPU backend
by tools/make_golden.py): runs on the GPU box with no reference tree and covers curves/fields whose reference build does
not travel."""
import os

import numpy as np
import pytest

import icicle_b200 as ib
from icicle_b200 import utils
import common

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    path = os.path.join(GOLD, f"{name}.npz")
    if not os.path.exists(path):
        pytest.skip(f"no golden fixture for {name}")
    return np.load(path)


def affine_of(curve_limbs, q, proj):
    return common.projective_to_affine_ints(proj, curve_limbs, q)


@pytest.mark.parametrize("name,curve,g2curve,fq,L", [("bn254", ib.Curve.BN254_G1, ib.Curve.BN254_G2, "bn254_fq", 8),
                                                     ("bls12_381", ib.Curve.BLS12_381_G1, ib.Curve.BLS12_381_G2, "bls12_381_fq", 12),
                                                     ("bls12_377", ib.Curve.BLS12_377_G1, ib.Curve.BLS12_377_G2, "bls12_377_fq", 12),
                                                     ("bw6_761", ib.Curve.BW6_761_G1, ib.Curve.BW6_761_G2, "bw6_761_fq", 24),
                                                     ("grumpkin", ib.Curve.GRUMPKIN, None, "bn254_fr", 8)])
def test_msm_golden(name, curve, g2curve, fq, L):
    """Every curve of the north-star list against outputs of ITS OWN reference build (tools/make_golden.py): G1 MSM (auto and
    forced c, bitsize, batch, Montgomery inputs), the curve Montgomery conversion, and the G2 MSM (Fq2 for bn254 / bls12-381 /
    bls12-377, Fq for bw6-761; grumpkin has no G2) -- mirrors icicle/tests/test_curve_api.cpp:275-289."""
    g = gold(name)
    q = utils.field_params(fq)["p"]
    s, P = g["msm_scalars"], g["msm_points"]
    n = s.shape[0]
    exp = common.affine_limbs_to_ints(g["msm_result_affine"], L)[0]
    for c in (0, 4, 11):
        assert affine_of(L, q, ib.msm(curve, s, P, n, ib.MSMConfig(c=c))[0]) == exp
    for bits in (1, 17, 100):
        got = ib.msm(curve, s, P, n, ib.MSMConfig(bitsize=bits))[0]
        assert affine_of(L, q, got) == common.affine_limbs_to_ints(g[f"msm_bitsize{bits}_affine"], L)[0]
    got = ib.msm(curve, s, P[:32], 32, ib.MSMConfig(batch_size=3))
    for b in range(3):
        assert affine_of(L, q, got[b]) == common.affine_limbs_to_ints(g["msm_batch3_affine"][b], L)[0]
    # Montgomery-form scalars and points (the reference's own conversions)
    got = ib.msm(curve, g["scalars_montgomery"], g["points_montgomery"], n,
                 ib.MSMConfig(are_scalars_montgomery_form=True, are_points_montgomery_form=True))
    assert affine_of(L, q, got[0]) == exp
    assert np.array_equal(ib.affine_convert_montgomery(curve, P, n, True), g["points_montgomery"])
    if g2curve is None:
        return
    P2 = g["g2_points"]
    for c in (0, 5):
        got = ib.msm(g2curve, s[:24], P2, 24, ib.MSMConfig(c=c))[0]
        if utils.curve_params(name)["g2"] == "fq":  # bw6-761: G2 lives over the base field itself
            assert affine_of(L, q, got) == common.affine_limbs_to_ints(g["g2_msm_result_affine"], L)[0]
            continue
        # G2 over Fq2: compare affine coordinates (x = X/Z, y = Y/Z with the Fq2 inversion done on integers)
        nr = utils.curve_params(name)["nonresidue"]
        X0, X1, Y0, Y1, Z0, Z1 = utils.from_limbs(got.reshape(6, L))
        den = pow((Z0 * Z0 - nr * Z1 * Z1) % q, -1, q)
        zi = (Z0 * den % q, (-Z1) * den % q)
        mul = lambda a, b: ((a[0] * b[0] + nr * a[1] * b[1]) % q, (a[0] * b[1] + a[1] * b[0]) % q)
        x, y = mul((X0, X1), zi), mul((Y0, Y1), zi)
        ex = utils.from_limbs(g["g2_msm_result_affine"].reshape(4, L))
        assert [x[0], x[1], y[0], y[1]] == ex


@pytest.mark.parametrize("name,field,fname", [("bn254", ib.Field.BN254_FR, "bn254_fr"), ("bls12_381", ib.Field.BLS12_381_FR, "bls12_381_fr"),
                                              ("babybear", ib.Field.BABYBEAR, "babybear"), ("bls12_377", ib.Field.BLS12_377_FR, "bls12_377_fr"),
                                              ("bw6_761", ib.Field.BLS12_377_FQ, "bls12_377_fq"), ("stark252", ib.Field.STARK252, "stark252"),
                                              ("goldilocks", ib.Field.GOLDILOCKS, "goldilocks")])
def test_ntt_and_vec_golden(name, field, fname):
    """NTT (all orderings, cosets, column batch, both directions) and vec-ops of every NTT field of the north-star list
    against ITS OWN reference build's outputs; bw6_761's scalar field is the 12-limb bls12_377 base field."""
    g = gold(name)
    L = utils.field_params(fname)["limbs"]
    logn = 6
    n = 1 << logn
    ib.ntt_release_domain(field)
    ib.ntt_init_domain(field, g["ntt_root"].reshape(-1))
    x = g["ntt_input"].reshape(-1, L)
    for d in (0, 1):
        for o in range(4):
            got = ib.ntt(field, x[:n], n, d, ib.NTTConfig(ordering=ib.Ordering(o)))
            assert np.array_equal(got, g[f"ntt_d{d}_o{o}"].reshape(-1, L)), (d, o)
        for kind in ("dom", "arb"):
            got = ib.ntt(field, x[:n], n, d, ib.NTTConfig(coset_gen=g[f"coset_{kind}"].reshape(-1)))
            assert np.array_equal(got, g[f"ntt_d{d}_coset_{kind}"].reshape(-1, L)), (d, kind)
        got = ib.ntt(field, x, n, d, ib.NTTConfig(batch_size=2, columns_batch=True))
        assert np.array_equal(got, g[f"ntt_d{d}_batch2_cols"].reshape(-1, L))
    ib.ntt_release_domain(field)
    a, b = g["vec_a"].reshape(-1, L), g["vec_b"].reshape(-1, L)
    assert np.array_equal(ib.vector_add(field, a, b, 50), g["vec_add"].reshape(-1, L))
    assert np.array_equal(ib.vector_sub(field, a, b, 50), g["vec_sub"].reshape(-1, L))
    assert np.array_equal(ib.vector_mul(field, a, b, 50), g["vec_mul"].reshape(-1, L))


@pytest.mark.parametrize("name,field", [("babybear", ib.Field.BABYBEAR), ("koalabear", ib.Field.KOALABEAR)])
def test_small_field_ntt_big_golden(name, field):
    """4-byte-field tile pass (csrc/ntt31.cuh) vs outputs of the unmodified reference CPU backend at 2^10 .. 2^18
    (tests/golden/<field>_ntt_big.npz, tools/make_golden_smallfield.py): bit-exact, full arrays or SHA-256 of the bytes."""
    import hashlib
    g = np.load(os.path.join(GOLD, f"{name}_ntt_big.npz"))
    p = utils.field_params(name)["p"]
    ib.ntt_release_domain(field)
    ib.ntt_init_domain(field, g["ntt_root"].reshape(-1))
    cases = sorted({(int(k.split("_")[1][1:]), int(k.split("_")[2][1:])) for k in g.files if k.startswith("sha_")})
    assert len(cases) == 9
    for logn, batch in cases:
        rs = np.random.RandomState(1000 + logn)
        x = rs.randint(0, p, size=(batch << logn, 1), dtype=np.int64).astype(np.uint32)
        for d in (0, 1):
            for c in (0, 1):
                cfg = ib.NTTConfig(batch_size=batch, coset_gen=g["coset_arb"].reshape(-1) if c else None)
                y = ib.ntt(field, x, 1 << logn, d, cfg)
                key = f"l{logn}_b{batch}_d{d}_c{c}"
                assert hashlib.sha256(np.ascontiguousarray(y, dtype=np.uint32).tobytes()).digest() == g["sha_" + key].tobytes(), key
                if "out_" + key in g.files:
                    assert np.array_equal(y, g["out_" + key]), key
    ib.ntt_release_domain(field)


@pytest.mark.parametrize("name,field", [("babybear", ib.Field.BABYBEAR), ("koalabear", ib.Field.KOALABEAR)])
def test_extension_ntt_golden(name, field):
    """Quartic-extension NTT (b200_ntt_extension) vs outputs of the unmodified reference CPU backend built with EXT_FIELD
    (`<field>_extension_ntt`; tests/golden/<field>_ext_ntt.npz, tools/make_golden_ext.py): forward / inverse, coset, row and
    columns batches, kNN and kNR, sizes 1 .. 2^16; bit-exact (full arrays or SHA-256 of the bytes)."""
    import hashlib
    g = np.load(os.path.join(GOLD, f"{name}_ext_ntt.npz"))
    p = utils.field_params(name)["p"]
    ib.ntt_release_domain(field)
    ib.ntt_init_domain(field, g["ntt_root"].reshape(-1))
    for logn, batch, col, ordering in g["cases"].tolist():
        rs = np.random.RandomState(4000 + logn)
        x = rs.randint(0, p, size=(batch << logn, 4), dtype=np.int64).astype(np.uint32)
        for d in (0, 1):
            for c in (0, 1):
                cfg = ib.NTTConfig(batch_size=batch, columns_batch=bool(col), ordering=ib.Ordering(ordering),
                                   coset_gen=g["coset_arb"].reshape(-1) if c else None)
                y = ib.ntt_extension(field, x, 1 << logn, d, cfg)
                key = f"l{logn}_b{batch}_c{col}_o{ordering}_d{d}_g{c}"
                if "out_" + key in g.files:
                    assert np.array_equal(y, g["out_" + key]), key
                assert hashlib.sha256(np.ascontiguousarray(y, dtype=np.uint32).tobytes()).digest() == g["sha_" + key].tobytes(), key
    # device-resident, in place, round trip at a size the tile pass handles in 2 passes
    n = 1 << 15
    rs = np.random.RandomState(77)
    x = rs.randint(0, p, size=(n * 2, 4), dtype=np.int64).astype(np.uint32)
    dx = ib.to_device(x)
    ib.ntt_extension(field, dx, n, 0, ib.NTTConfig(batch_size=2, are_outputs_on_device=True), dx)
    # coefficient planes: the extension transform is 4 base-field transforms over the interleaved coefficients
    planes = np.ascontiguousarray(x.reshape(2, n, 4).transpose(0, 2, 1)).reshape(-1, 1)    # [batch][4][n]
    yb = ib.ntt(field, planes, n, 0, ib.NTTConfig(batch_size=8))
    assert np.array_equal(ib.to_host(dx).reshape(2, n, 4), yb.reshape(2, 4, n).transpose(0, 2, 1))
    ib.ntt_extension(field, dx, n, 1, ib.NTTConfig(batch_size=2, are_outputs_on_device=True), dx)
    assert np.array_equal(ib.to_host(dx).reshape(-1, 4), x)
    ib.ntt_release_domain(field)


def test_m31_vec_ops_golden():
    """Mersenne-31 vec-ops (north_star's field list; vec-ops only upstream: icicle/cmake/features.cmake:7) vs the reference's
    MersenneField outputs (tests/golden/m31.npz) incl. 0, 1, p-1 and the identity Montgomery conversion; inv / div / sum /
    product against Python integers; the NTT entry point reports API_NOT_IMPLEMENTED like the reference build (no NTT feature)."""
    g = gold("m31")
    F = ib.Field.M31
    p = utils.field_params("m31")["p"]
    a, b = g["vec_a"], g["vec_b"]
    n = a.shape[0]
    assert np.array_equal(ib.vector_add(F, a, b, n), g["vector_add"])
    assert np.array_equal(ib.vector_sub(F, a, b, n), g["vector_sub"])
    assert np.array_equal(ib.vector_mul(F, a, b, n), g["vector_mul"])
    acc = ib.to_device(a)
    ib.vector_accumulate(F, acc, ib.to_device(b), n)
    assert np.array_equal(ib.to_host(acc).reshape(-1, 1), g["vector_accumulate"])
    assert np.array_equal(ib.convert_montgomery(F, a, n, True), g["to_montgomery"])
    assert np.array_equal(ib.convert_montgomery(F, a, n, False), g["from_montgomery"])
    assert np.array_equal(ib.bit_reverse(F, a, n), g["bit_reverse"])
    ai = [int(v) for v in a[:, 0]]
    nz = np.array([v if v else 5 for v in ai], dtype=np.uint32).reshape(-1, 1)
    inv = ib.vector_inv(F, nz, n)
    assert [int(v) for v in inv[:, 0]] == [pow(int(v), -1, p) for v in nz[:, 0]]
    div = ib.vector_div(F, b, nz, n)
    assert [int(v) for v in div[:, 0]] == [int(x) * pow(int(y), -1, p) % p for x, y in zip(b[:, 0], nz[:, 0])]
    assert int(ib.vector_sum(F, a, n)[0, 0]) == sum(ai) % p
    prod = 1
    for v in nz[:, 0]:
        prod = prod * int(v) % p
    assert int(ib.vector_product(F, nz, n)[0, 0]) == prod
    with pytest.raises(ib.IcicleError):
        ib.ntt_init_domain(F, np.array([p - 1], dtype=np.uint32))


def test_grumpkin_scalar_vec_ops_golden():
    """Grumpkin has no NTT upstream (features.cmake): its scalar field (= bn254 base field) vec-ops against the reference build."""
    g = gold("grumpkin")
    field, L = ib.Field.BN254_FQ, 8
    a, b = g["vec_a"].reshape(-1, L), g["vec_b"].reshape(-1, L)
    assert np.array_equal(ib.vector_add(field, a, b, 50), g["vec_add"].reshape(-1, L))
    assert np.array_equal(ib.vector_sub(field, a, b, 50), g["vec_sub"].reshape(-1, L))
    assert np.array_equal(ib.vector_mul(field, a, b, 50), g["vec_mul"].reshape(-1, L))


@pytest.mark.parametrize("name,ext,base", [("babybear", ib.Field.BABYBEAR_EXT4, ib.Field.BABYBEAR), ("koalabear", ib.Field.KOALABEAR_EXT4, ib.Field.KOALABEAR)])
def test_extension_vec_ops_golden(name, ext, base):
    """the REGISTER_*_EXT_FIELD_BACKEND family (vec_ops_backend.h:297-494) on the quartic extension of BabyBear /
    KoalaBear -- the ordinary C-ABI entry points with the extension's field id + b200_ext_mixed_mul -- bit-exact against the
    reference's own `<field>_extension_*` outputs (tests/golden/<field>_ext_ops.npz), host and device buffers."""
    g = gold(f"{name}_ext_ops")
    a, b, s = g["a"], g["b"], g["s"]
    n3 = a.shape[0]
    n, batch = n3 // 3, 3
    for dev in (False, True):
        A, B, S = (ib.to_device(a), ib.to_device(b), ib.to_device(s)) if dev else (a, b, s)
        h = (lambda x: ib.to_host(x).reshape(-1, 4)) if dev else (lambda x: x)
        cfg = lambda **kw: ib.VecOpsConfig(is_result_on_device=dev, **kw)
        assert np.array_equal(h(ib.vector_add(ext, A, B, n3, cfg())), g["vector_add"])
        assert np.array_equal(h(ib.vector_sub(ext, A, B, n3, cfg())), g["vector_sub"])
        assert np.array_equal(h(ib.vector_mul(ext, A, B, n3, cfg())), g["vector_mul"])
        assert np.array_equal(h(ib.vector_div(ext, A, B, n3, cfg())), g["vector_div"])
        assert np.array_equal(h(ib.vector_inv(ext, B, n3, cfg())), g["vector_inv"])
        assert np.array_equal(h(ib.ext_mixed_mul(ext, A, S, n3, cfg())), g["vector_mixed_mul"])
        acc = ib.to_device(a) if dev else a.copy()
        ib.vector_accumulate(ext, acc, B, n3)
        assert np.array_equal(h(acc), g["vector_accumulate"])
        for columns, tag in ((False, "rows"), (True, "cols")):
            sc = ib.to_device(a[:batch].copy()) if dev else a[:batch].copy()
            c2 = lambda: cfg(batch_size=batch, columns_batch=columns)
            assert np.array_equal(h(ib.scalar_add_vec(ext, sc, B, n, c2())), g[f"scalar_add_vec_{tag}"])
            assert np.array_equal(h(ib.scalar_sub_vec(ext, sc, B, n, c2())), g[f"scalar_sub_vec_{tag}"])
            assert np.array_equal(h(ib.scalar_mul_vec(ext, sc, B, n, c2())), g[f"scalar_mul_vec_{tag}"])
            assert np.array_equal(h(ib.vector_sum(ext, A, n, c2())), g[f"vector_sum_{tag}"])
            assert np.array_equal(h(ib.vector_product(ext, A, n, c2())), g[f"vector_product_{tag}"])
        assert np.array_equal(h(ib.convert_montgomery(ext, A, n3, True, cfg())), g["convert_montgomery_1"])
        assert np.array_equal(h(ib.convert_montgomery(ext, A, n3, False, cfg())), g["convert_montgomery_0"])
        a32 = ib.to_device(a[:32].copy()) if dev else a[:32].copy()
        a48 = ib.to_device(a[:48].copy()) if dev else a[:48].copy()
        assert np.array_equal(h(ib.bit_reverse(ext, a32, 32, cfg())), g["bit_reverse"])
        assert np.array_equal(h(ib.matrix_transpose(ext, a48, 6, 8, cfg())), g["matrix_transpose_6x8"])
        assert np.array_equal(h(ib.slice(ext, a48, 3, 4, 48, 10, cfg())), g["slice_3_4_10"])


def test_goldilocks_montgomery_conversion_and_large_ntt():
    """Goldilocks (p = 2^64 - 2^32 + 1) has no internal Montgomery domain on the device: the API-level conversion must still be
    x * 2^(+-64) mod p (fields/params_gen.h:35-50), and a 2^18 transform must round-trip and match the defining sum at a few outputs."""
    F = ib.Field.GOLDILOCKS
    fp = utils.field_params("goldilocks")
    p = fp["p"]
    x = common.seeded_scalars("goldilocks", 1000, 5)
    xi = utils.from_limbs(x)
    assert utils.from_limbs(ib.convert_montgomery(F, x, 1000, True)) == [v * (1 << 64) % p for v in xi]
    assert utils.from_limbs(ib.convert_montgomery(F, x, 1000, False)) == [v * pow(1 << 64, -1, p) % p for v in xi]
    logn = 18
    n = 1 << logn
    w = pow(fp["rou"], 1 << (fp["two_adicity"] - logn), p)
    ib.ntt_release_domain(F)
    ib.ntt_init_domain(F, utils.to_limbs([w], 2)[0])
    assert utils.from_limbs(ib.get_root_of_unity_from_domain(F, logn).reshape(1, 2))[0] == w
    v = common.seeded_scalars("goldilocks", n, 6)
    V = ib.ntt(F, v, n, ib.NTTDir.kForward)
    assert np.array_equal(ib.ntt(F, V, n, ib.NTTDir.kInverse), v)
    sp = np.zeros((n, 2), dtype=np.uint32)
    a_idx, b_idx, alpha, beta = 4321, n - 3, 0x123456789ABCDEF, 11
    sp[a_idx], sp[b_idx] = utils.to_limbs([alpha], 2)[0], utils.to_limbs([beta], 2)[0]
    S = ib.ntt(F, sp, n, ib.NTTDir.kForward)
    for k in (0, 1, 77777, n - 1):
        assert utils.from_limbs(S[k:k + 1])[0] == (alpha * pow(w, a_idx * k, p) + beta * pow(w, b_idx * k, p)) % p
    ib.ntt_release_domain(F)
