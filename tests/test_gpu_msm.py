"""GPU parity: MSM through the C ABI vs the reference CPU backend and Python-integer group arithmetic.
Mirrors icicle/tests/test_curve_api.cpp:35-171 (msm, msm_bitsize, msm_pre_compute) and
wrappers/rust/icicle-core/src/msm/tests.rs:26-300 (sizes, Montgomery scalars, zero points, batches, skewed scalars)."""
import random

import numpy as np
import pytest

import icicle_b200 as ib
from icicle_b200 import utils
import common

pytestmark = pytest.mark.gpu

G1 = [(ib.Curve.BN254_G1, "bn254"), (ib.Curve.GRUMPKIN, "grumpkin"), (ib.Curve.BLS12_381_G1, "bls12_381"),
      (ib.Curve.BLS12_377_G1, "bls12_377"), (ib.Curve.BW6_761_G1, "bw6_761")]


def _skip_if_not_built(fn):
    try:
        return fn()
    except ib.IcicleError as e:
        if e.code == 10:
            pytest.skip("curve not built into libicicle_b200.so (developer subset build)")
        raise


def check_against_ints(curve, name, scalars_int, pts_int, result_limbs, bitsize=0):
    cp = utils.curve_params(name)
    q = utils.field_params(cp["fq"])["p"]
    L = utils.field_params(cp["fq"])["limbs"]
    if bitsize:
        scalars_int = [s & ((1 << bitsize) - 1) for s in scalars_int]
    exp = common.msm_naive_ints(scalars_int, pts_int, q)
    if exp is None:
        assert common.is_projective_zero(result_limbs, L)
    else:
        assert common.projective_to_affine_ints(result_limbs, L, q) == exp


@pytest.mark.parametrize("curve,name", G1)
def test_small_vs_python(curve, name):
    cp = utils.curve_params(name)
    fr, fq = utils.field_params(cp["fr"]), utils.field_params(cp["fq"])
    n = 70
    pts_int = common.gen_g1_points(name, n, 5, as_ints=True)
    pts_int[3] = None  # affine zero (0,0) must be skipped (cpu_msm.hpp:282)
    pts_int[10] = pts_int[11]  # repeated point -> doubling path
    sc = common.rand_field_elems(cp["fr"], n, 6, as_ints=True)
    sc[0], sc[1], sc[2] = 0, 1, fr["p"] - 1
    sc[10] = sc[11]
    flat = []
    for P in pts_int:
        flat += [0, 0] if P is None else [P[0], P[1]]
    P_l = utils.to_limbs(flat, fq["limbs"]).reshape(n, -1)
    S_l = utils.to_limbs(sc, fr["limbs"])
    for c in (0, 3, 8, 13):
        res = _skip_if_not_built(lambda: ib.msm(curve, S_l, P_l, n, ib.MSMConfig(c=c)))
        check_against_ints(curve, name, sc, pts_int, res[0])
    for bitsize in (1, 10, 33, 100, fr["bits"] - 1):
        res = ib.msm(curve, S_l, P_l, n, ib.MSMConfig(bitsize=bitsize))
        check_against_ints(curve, name, sc, pts_int, res[0], bitsize=bitsize)
    # n = 1 and the all-zero MSM
    res = ib.msm(curve, S_l[1:2], P_l[1:2], 1)
    check_against_ints(curve, name, sc[1:2], pts_int[1:2], res[0])
    res = ib.msm(curve, S_l[:1], P_l[:1], 1)
    assert common.is_projective_zero(res[0], fq["limbs"])


@pytest.fixture(scope="module")
def ref():
    ref_icicle = pytest.importorskip("ref_icicle")
    if not ref_icicle.available("bn254"):
        pytest.skip("oracle/_ref/bn254 not built")
    return ref_icicle.get("bn254")


def test_bn254_vs_reference_sizes(ref):
    C = ib.Curve.BN254_G1
    for n in (1, 2, 100, (1 << 12) - 37, 1 << 14, (1 << 16) + 5):
        s = ref.generate_scalars(n)
        P = ref.generate_affine_points(n)  # 100 distinct points repeated, like the reference tests (projective.h:43-53)
        exp = ref.msm(s, P, n)
        got = ib.msm(C, s, P, n)
        assert ref.projective_eq(got[0], exp[0]), n
        # device-resident inputs/outputs + Montgomery scalars + async stream (msm/tests.rs:26-70)
        import torch
        s_m = ref.scalar_convert_montgomery(s, n, True)
        st = torch.cuda.Stream()
        cfg = ib.MSMConfig(are_scalars_montgomery_form=True, is_async=True, stream=st)
        out = ib.device_empty(24).view(1, 24)
        ib.msm(C, ib.to_device(s_m), ib.to_device(P), n, cfg, out)
        st.synchronize()
        assert ref.projective_eq(ib.to_host(out)[0], exp[0]), n


def test_bn254_distinct_points_and_window_sweep(ref):
    C = ib.Curve.BN254_G1
    n = 1 << 13
    P = common.gen_g1_points("bn254", n, 77)
    s = ref.generate_scalars(n)
    exp = ref.msm(s, P, n)
    for c in (0, 5, 10, 12, 15, 16, 18):
        got = ib.msm(C, s, P, n, ib.MSMConfig(c=c))
        assert ref.projective_eq(got[0], exp[0]), c
    # Montgomery-form points: the reference CPU backend ignores this flag in accumulation (cpu_msm.hpp:280), so the
    # documented meaning (msm.h:43-44) is checked by converting the inputs ourselves: parity unpinned upstream.
    P_m = ref.affine_convert_montgomery(P, n, True)
    got = ib.msm(C, s, P_m, n, ib.MSMConfig(are_points_montgomery_form=True))
    assert ref.projective_eq(got[0], exp[0])


def test_bn254_bitsize_sweep(ref):
    C = ib.Curve.BN254_G1
    n = (1 << 10) - 11
    s = ref.generate_scalars(n)
    P = ref.generate_affine_points(n)
    for bitsize in list(range(1, 40)) + [63, 64, 65, 127, 128, 200, 253, 254]:
        exp = ref.msm(s, P, n, bitsize=bitsize)
        got = ib.msm(C, s, P, n, ib.MSMConfig(bitsize=bitsize))
        assert ref.projective_eq(got[0], exp[0]) or (common.is_projective_zero(got[0], 8) and common.is_projective_zero(exp[0], 8)), bitsize


def test_bn254_batch_and_precompute(ref):
    C = ib.Curve.BN254_G1
    n, batch = (1 << 10) + 3, 3
    s = ref.generate_scalars(n * batch)
    P = ref.generate_affine_points(n * batch)
    # shared bases
    exp = ref.msm(s, P[:n], n, batch_size=batch, are_points_shared_in_batch=True)
    got = ib.msm(C, s, P[:n], n, ib.MSMConfig(batch_size=batch, are_points_shared_in_batch=True))
    for b in range(batch):
        assert ref.projective_eq(got[b], exp[b])
    # per-MSM bases
    exp = ref.msm(s, P, n, batch_size=batch, are_points_shared_in_batch=False)
    got = ib.msm(C, s, P, n, ib.MSMConfig(batch_size=batch, are_points_shared_in_batch=False))
    for b in range(batch):
        assert ref.projective_eq(got[b], exp[b])
    # forced chunking of the batch gives the same
    got2 = ib.msm(C, s, P, n, ib.MSMConfig(batch_size=batch, are_points_shared_in_batch=False, ext={"nof_chunks": 3}))
    for b in range(batch):  # same group elements (the projective representative may differ with the chunking)
        assert ref.projective_eq(got2[b], exp[b])
    # precompute (test_curve_api.cpp:125-171): precompute and msm must agree within a backend; result == plain MSM
    for pf, c in ((2, 0), (3, 7), (8, 4), (5, 16)):
        cfgp = ib.MSMConfig(precompute_factor=pf, c=c)
        pre = ib.msm_precompute_bases(C, P[:n], n, cfgp)
        assert pre.shape == (n * pf, 16)
        got = ib.msm(C, s, pre, n, ib.MSMConfig(precompute_factor=pf, c=c, batch_size=batch, are_points_shared_in_batch=True))
        exp = ref.msm(s, P[:n], n, batch_size=batch, are_points_shared_in_batch=True)
        for b in range(batch):
            assert ref.projective_eq(got[b], exp[b]), (pf, c, b)


def test_bn254_skewed_scalars_and_zero_points(ref):
    """msm/tests.rs:256-300: mostly 0/1 scalars with bitsize=1 and injected zero points; all-equal scalars (one huge bucket)."""
    C = ib.Curve.BN254_G1
    n = 1 << 14
    rng = random.Random(3)
    P = ref.generate_affine_points(n)
    P[::17] = 0
    sc = [rng.choice((0, 1, 1, 1)) for _ in range(n)]
    s = utils.to_limbs(sc, 8)
    exp = ref.msm(s, P, n, bitsize=1)
    got = ib.msm(C, s, P, n, ib.MSMConfig(bitsize=1))
    assert ref.projective_eq(got[0], exp[0])
    got = ib.msm(C, s, P, n, ib.MSMConfig(bitsize=1, c=2))
    assert ref.projective_eq(got[0], exp[0])
    big = rng.randrange(1 << 253)
    s = utils.to_limbs([big] * n, 8)
    exp = ref.msm(s, P, n)
    got = ib.msm(C, s, P, n)
    assert ref.projective_eq(got[0], exp[0])


def test_bn254_g2_vs_reference(ref):
    C = ib.Curve.BN254_G2
    n = 1 << 10
    s = ref.generate_scalars(n)
    P = ref.generate_affine_points(n, g2=True)
    exp = ref.msm(s, P, n, g2=True)
    got = _skip_if_not_built(lambda: ib.msm(C, s, P, n))
    assert ref.projective_eq(got[0], exp[0], g2=True)
    pre = ib.msm_precompute_bases(C, P, n, ib.MSMConfig(precompute_factor=3))
    got = ib.msm(C, s, pre, n, ib.MSMConfig(precompute_factor=3))
    assert ref.projective_eq(got[0], exp[0], g2=True)


def test_linearity_large():
    """Size-independent property at a size the CPU oracle would take minutes for: MSM(s, P) + MSM(t, P) == MSM(s + t, P)."""
    C = ib.Curve.BN254_G1
    n = 1 << 20
    fr = utils.field_params("bn254_fr")
    base = common.gen_g1_points("bn254", 1 << 10, 123)
    P = np.tile(base, (n >> 10, 1))
    rs = np.random.RandomState(5)
    s = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    t = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    s[:, 7] &= 0x0FFFFFFF
    t[:, 7] &= 0x0FFFFFFF
    st = ib.vector_add(ib.Field.BN254_FR, s, t, n)
    dP = ib.to_device(P)
    a = ib.msm(C, s, dP, n)[0]
    b = ib.msm(C, t, dP, n)[0]
    c = ib.msm(C, st, dP, n)[0]
    q = utils.field_params("bn254_fq")["p"]
    A = common.projective_to_affine_ints(a, 8, q)
    B = common.projective_to_affine_ints(b, 8, q)
    Cc = common.projective_to_affine_ints(c, 8, q)
    assert common.ec_add(A, B, q) == Cc


def test_pipelined_host_path_matches_device_path():
    """Host-pointer calls with >= 2^23 points take the chunked copy/compute pipeline (msm_chunked); the result must be the
    same group element as the device-resident single-pass path."""
    C = ib.Curve.BN254_G1
    n = 1 << 23
    base = common.gen_g1_points("bn254", 1 << 10, 321)
    P = np.tile(base, (n >> 10, 1))
    rs = np.random.RandomState(9)
    s = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    s[:, 7] &= 0x0FFFFFFF
    host = ib.msm(C, s, P, n)[0]                                  # host pointers -> pipelined
    dev = ib.msm(C, ib.to_device(s), ib.to_device(P), n)[0]       # device pointers -> single pass
    q = utils.field_params("bn254_fq")["p"]
    assert common.projective_to_affine_ints(host, 8, q) == common.projective_to_affine_ints(dev, 8, q)


def test_pipelined_host_path_vs_reference_small(ref, tuning):
    """The host-pointer pipeline (chunked H2D on a copy stream, per-chunk accumulation into one shared bucket array with the
    window size of the whole MSM, k_bucket_merge, one bucket reduction) forced at small sizes: same group element as the
    reference CPU backend for ragged chunk splits, chunk counts 1..7, forced pair levels and the degenerate inputs."""
    C = ib.Curve.BN254_G1
    tuning("msm_pipeline_min", 2)
    for n, chunks in ((2, 2), (3, 2), (1000, 3), ((1 << 12) + 17, 7), (1 << 13, 1), ((1 << 14) - 1, 5)):
        tuning("msm_pipeline_chunks", chunks)
        s = ref.generate_scalars(n)
        P = ref.generate_affine_points(n)   # 100 distinct points repeated: doublings inside buckets, across chunks too
        exp = ref.msm(s, P, n)
        for lv, c in ((0, 0), (2, 7), (1, 12), (0, 15)):
            tuning("msm_pair_levels", lv)
            got = ib.msm(C, s, P, n, ib.MSMConfig(c=c))
            assert ref.projective_eq(got[0], exp[0]), (n, chunks, lv, c)
    # the default graded schedule (1/16, 1/16, 1/8, 1/4, 1/4, rest)
    tuning("msm_pipeline_chunks", None)
    tuning("msm_pair_levels", 2)
    for n in (5, 100, (1 << 13) + 11):
        s = ref.generate_scalars(n)
        P = ref.generate_affine_points(n)
        assert ref.projective_eq(ib.msm(C, s, P, n, ib.MSMConfig(c=9))[0], ref.msm(s, P, n)[0]), n
    # bitsize=1 (one huge bucket), zero bases, Montgomery-form scalars, P + (-P) in different chunks
    tuning("msm_pipeline_chunks", 4)
    tuning("msm_pair_levels", 1)
    n = 1 << 12
    s = ref.generate_scalars(n)
    P = ref.generate_affine_points(n)
    P[5] = 0
    P[n - 1] = 0
    assert ref.projective_eq(ib.msm(C, s, P, n, ib.MSMConfig(bitsize=1))[0], ref.msm(s, P, n, bitsize=1)[0])
    q = utils.field_params("bn254_fq")["p"]
    P2 = common.gen_g1_points("bn254", n, 5)
    half = n // 2
    P2[half:] = P2[:half]
    yi = utils.from_limbs(P2[half:, 8:])
    P2[half:, 8:] = utils.to_limbs([(q - y) % q for y in yi], 8)
    s2 = ref.generate_scalars(n)
    s2[half:] = s2[:half]
    got = ib.msm(C, s2, P2, n)
    assert common.is_projective_zero(ref.msm(s2, P2, n)[0], 8) and common.is_projective_zero(got[0], 8)   # P and -P meet only in the merge


@pytest.fixture
def pair_levels_env(tuning):
    """msm_pair_levels forces the number of batched-affine pair levels (msm_pairs.cuh) regardless of size."""
    return lambda v: tuning("msm_pair_levels", v)


def test_bn254_pair_levels_vs_reference(ref, pair_levels_env):
    """The batched-affine pair tree must give the same group element as the reference for every level count, including the
    cases that need the special branches: the reference generator's 100 repeated points (doublings inside a bucket),
    P + (-P), affine zero bases, bitsize=1 (one huge bucket), all-equal scalars, odd run lengths, tiny windows."""
    C = ib.Curve.BN254_G1
    rng = random.Random(11)
    for n in (1, 2, 3, 100, (1 << 12) - 37, (1 << 14) + 5):
        s = ref.generate_scalars(n)
        P = ref.generate_affine_points(n)
        exp = ref.msm(s, P, n)
        for lv in (1, 2, 3, 5, 8):
            pair_levels_env(lv)
            for c in (0, 2, 7, 12):
                got = ib.msm(C, s, P, n, ib.MSMConfig(c=c))
                assert ref.projective_eq(got[0], exp[0]), (n, lv, c)
    # distinct points, P and -P in the same bucket, zero bases, skewed scalars
    n = 1 << 13
    P = common.gen_g1_points("bn254", n, 78)
    q = utils.field_params("bn254_fq")["p"]
    Pn = P.copy()
    yi = utils.from_limbs(P[:, 8:])
    Pn[:, 8:] = utils.to_limbs([(q - y) % q for y in yi], 8)
    P2 = np.concatenate([P[: n // 2], Pn[: n // 2]])          # second half = negatives of the first half
    s = ref.generate_scalars(n)
    s[n // 2:] = s[: n // 2]                                     # same scalars -> every bucket holds P and -P
    P2[5::31] = 0
    for Pt, st in ((P, s), (P2, s)):
        exp = ref.msm(st, Pt, n)
        for lv in (1, 3, 4):
            pair_levels_env(lv)
            for c in (0, 4, 9):
                got = ib.msm(C, st, Pt, n, ib.MSMConfig(c=c))
                if common.is_projective_zero(exp[0], 8):
                    assert common.is_projective_zero(got[0], 8), (lv, c)
                else:
                    assert ref.projective_eq(got[0], exp[0]), (lv, c)
    sc = [rng.choice((0, 1, 1, 1)) for _ in range(n)]
    s1 = utils.to_limbs(sc, 8)
    Pr = ref.generate_affine_points(n)
    Pr[::17] = 0
    exp = ref.msm(s1, Pr, n, bitsize=1)
    big = utils.to_limbs([rng.randrange(1 << 253)] * n, 8)
    exp_big = ref.msm(big, Pr, n)
    for lv in (1, 4, 8):
        pair_levels_env(lv)
        got = ib.msm(C, s1, Pr, n, ib.MSMConfig(bitsize=1))
        assert ref.projective_eq(got[0], exp[0]), lv
        got = ib.msm(C, big, Pr, n)
        assert ref.projective_eq(got[0], exp_big[0]), lv
    # batch + precompute through the levels
    nb_, batch = (1 << 10) + 3, 3
    s = ref.generate_scalars(nb_ * batch)
    P = ref.generate_affine_points(nb_ * batch)
    exp = ref.msm(s, P, nb_, batch_size=batch, are_points_shared_in_batch=False)
    pair_levels_env(2)
    got = ib.msm(C, s, P, nb_, ib.MSMConfig(batch_size=batch, are_points_shared_in_batch=False, c=6))
    for b in range(batch):
        assert ref.projective_eq(got[b], exp[b]), b
    pre = ib.msm_precompute_bases(C, P[:nb_], nb_, ib.MSMConfig(precompute_factor=3, c=7))
    got = ib.msm(C, s, pre, nb_, ib.MSMConfig(precompute_factor=3, c=7, batch_size=batch, are_points_shared_in_batch=True))
    exp = ref.msm(s, P[:nb_], nb_, batch_size=batch, are_points_shared_in_batch=True)
    for b in range(batch):
        assert ref.projective_eq(got[b], exp[b]), b


def test_bn254_g2_pair_levels(ref, pair_levels_env):
    C = ib.Curve.BN254_G2
    n = 1 << 10
    s = ref.generate_scalars(n)
    P = ref.generate_affine_points(n, g2=True)
    exp = ref.msm(s, P, n, g2=True)
    for lv in (1, 3):
        pair_levels_env(lv)
        got = _skip_if_not_built(lambda: ib.msm(C, s, P, n, ib.MSMConfig(c=5)))
        assert ref.projective_eq(got[0], exp[0], g2=True), lv


def test_pair_levels_large_matches_xyzz_only(pair_levels_env):
    """2^22 points: automatic schedule vs pair levels forced off / forced on (3 and 6 levels) -- same group element."""
    C = ib.Curve.BN254_G1
    n = 1 << 22
    base = common.gen_g1_points("bn254", 1 << 10, 55)
    dP = ib.to_device(np.tile(base, (n >> 10, 1)))
    rs = np.random.RandomState(19)
    s = rs.randint(0, 1 << 32, size=(n, 8), dtype=np.uint64).astype(np.uint32)
    s[:, 7] &= 0x0FFFFFFF
    ds = ib.to_device(s)
    q = utils.field_params("bn254_fq")["p"]
    auto = common.projective_to_affine_ints(ib.msm(C, ds, dP, n)[0], 8, q)
    pair_levels_env(0)
    plain = common.projective_to_affine_ints(ib.msm(C, ds, dP, n)[0], 8, q)
    assert auto == plain
    for lv in (3, 6):
        pair_levels_env(lv)
        deep = common.projective_to_affine_ints(ib.msm(C, ds, dP, n)[0], 8, q)
        assert deep == plain, lv
