"""GPU parity of the ECNTT (b200_ecntt, csrc/ecntt.cuh) -- the NTT over G1 points with scalar-field twiddles
(ECNttFieldImpl, icicle/include/icicle/backend/ecntt_backend.h:15-22; reference tests: icicle/tests/test_curve_api.cpp:293-400,
which compare against the CPU backend as group elements).  Checked against (i) the unmodified reference CPU backend when
oracle/_ref/bn254 travelled with the snapshot, (ii) the committed golden vectors generated from it, (iii) the defining
identity out[k] = MSM(w^(ik), P_i) through the (separately pinned) MSM on the other curves, (iv) inverse(forward) = identity."""
import os

import numpy as np
import pytest

import icicle_b200 as ib
from icicle_b200 import utils
import common

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ref():
    ref_icicle = pytest.importorskip("ref_icicle")
    if not ref_icicle.available("bn254"):
        pytest.skip("oracle/_ref/bn254 not built")
    r = ref_icicle.get("bn254")
    if not hasattr(r.curve, "bn254_ecntt"):
        pytest.skip("oracle/_ref/bn254 was built without ECNTT")
    return r


def _domain(field, fname, logn):
    fp = utils.field_params(fname)
    root = pow(fp["rou"], 1 << (fp["two_adicity"] - logn), fp["p"])
    ib.ntt_release_domain(field)
    ib.ntt_init_domain(field, utils.to_limbs([root], fp["limbs"])[0])
    return root, fp


def test_ecntt_bn254_vs_reference(ref):
    C, F = ib.Curve.BN254_G1, ib.Field.BN254_FR
    root, fp = _domain(F, "bn254_fr", 10)
    ref.ntt_init_domain(utils.to_limbs([root], 8)[0])
    g = np.array([0x7654321, 0xfedcba9, 5, 0, 0, 0, 0, 0], dtype=np.uint32)
    try:
        for n, batch in ((1, 1), (2, 1), (8, 3), (64, 1), (256, 2)):
            aff = ref.generate_affine_points(n * batch)          # 100 distinct points repeated
            if n * batch > 5:
                aff[5] = 0                                        # the point at infinity as an input
            P = common.affine_to_projective_limbs(aff, 8)
            for d in (0, 1):
                for o in (0, 1, 2, 3):
                    for coset in (None, g):
                        if n >= 64 and (o in (1, 3)) and coset is not None:
                            continue                              # keep the CPU reference time bounded
                        exp = ref.ecntt(P, n, d, coset_gen=coset, batch_size=batch, ordering=o)
                        got = ib.ecntt(C, P, n, d, ib.NTTConfig(batch_size=batch, ordering=ib.Ordering(o), coset_gen=coset))
                        for i in range(n * batch):
                            assert ref.projective_eq(got[i], exp[i]), (n, batch, d, o, coset is not None, i)
        # columns batch, device-resident in place
        n, batch = 32, 3
        P = common.affine_to_projective_limbs(ref.generate_affine_points(n * batch), 8)
        exp = ref.ecntt(P, n, 0, batch_size=batch, columns_batch=True)
        dP = ib.to_device(P)
        ib.ecntt(C, dP, n, 0, ib.NTTConfig(batch_size=batch, columns_batch=True, are_outputs_on_device=True), dP)
        got = ib.to_host(dP).reshape(n * batch, 24)
        for i in range(n * batch):
            assert ref.projective_eq(got[i], exp[i]), i
    finally:
        ref.ntt_release_domain()
        ib.ntt_release_domain(F)


def test_ecntt_bn254_golden():
    """tests/golden/bn254_ecntt.npz (tools/make_golden_ecntt.py, outputs of the unmodified reference normalised to affine by the
    reference's own to_affine): runs with no reference build on the box."""
    g = np.load(os.path.join(GOLD, "bn254_ecntt.npz"))
    C, F = ib.Curve.BN254_G1, ib.Field.BN254_FR
    q = utils.field_params("bn254_fq")["p"]
    ib.ntt_release_domain(F)
    ib.ntt_init_domain(F, g["ntt_root"].reshape(-1))
    P = g["input_projective"]
    n, batch = 16, 2

    def check(got, key):
        exp = common.affine_limbs_to_ints(g[key], 8)
        for i in range(n * batch):
            assert common.projective_to_affine_ints(got[i], 8, q) == exp[i], (key, i)

    for d in (0, 1):
        for o in (0, 1, 2):
            for c in (0, 1):
                cfg = ib.NTTConfig(batch_size=batch, ordering=ib.Ordering(o), coset_gen=g["coset"] if c else None)
                check(ib.ecntt(C, P, n, d, cfg), f"d{d}_o{o}_g{c}_affine")
    check(ib.ecntt(C, P, n, 0, ib.NTTConfig(batch_size=batch, columns_batch=True)), "d0_cols_affine")
    ib.ntt_release_domain(F)


@pytest.mark.parametrize("cname,curve,field,fr,fq", [("bls12_381", ib.Curve.BLS12_381_G1, ib.Field.BLS12_381_FR, "bls12_381_fr", "bls12_381_fq"),
                                                     ("bn254", ib.Curve.BN254_G1, ib.Field.BN254_FR, "bn254_fr", "bn254_fq")])
def test_ecntt_is_msm_with_twiddle_scalars(cname, curve, field, fr, fq):
    """Defining identity (ntt_cpu.h:69-232 with E = projective_t): out[k] = sum_i w^(ik) * P_i, evaluated with the MSM."""
    logn = 5
    n = 1 << logn
    root, fp = _domain(field, fr, logn)
    L, Lq = fp["limbs"], utils.field_params(fq)["limbs"]
    q = utils.field_params(fq)["p"]
    aff = common.gen_g1_points(cname, n, 99)
    P = common.affine_to_projective_limbs(aff, Lq)
    got = ib.ecntt(curve, P, n, 0)
    inv = ib.ecntt(curve, got, n, 1)
    for k in (0, 1, 7, n - 1):
        sc = utils.to_limbs([pow(root, i * k, fp["p"]) for i in range(n)], L)
        exp = ib.msm(curve, sc, aff, n)[0]
        assert common.projective_to_affine_ints(got[k], Lq, q) == common.projective_to_affine_ints(exp, Lq, q), k
    for i in range(n):
        assert common.projective_to_affine_ints(inv[i], Lq, q) == common.affine_limbs_to_ints(aff[i], Lq)[0], i
    ib.ntt_release_domain(field)


def test_ecntt_roundtrip_2p10_device():
    C, F = ib.Curve.BN254_G1, ib.Field.BN254_FR
    _domain(F, "bn254_fr", 10)
    n = 1 << 10
    q = utils.field_params("bn254_fq")["p"]
    aff = common.gen_g1_points("bn254", n, 5)
    dP = ib.to_device(common.affine_to_projective_limbs(aff, 8))
    mid = ib.ecntt(C, dP, n, 0, ib.NTTConfig(are_outputs_on_device=True, ordering=ib.Ordering.kNR))
    back = ib.to_host(ib.ecntt(C, mid, n, 1, ib.NTTConfig(are_outputs_on_device=True, ordering=ib.Ordering.kRN))).reshape(n, 24)
    exp = common.affine_limbs_to_ints(aff, 8)
    for i in range(0, n, 37):
        assert common.projective_to_affine_ints(back[i], 8, q) == exp[i], i
    with pytest.raises(ib.IcicleError):
        ib.ecntt(C, dP, n << 1, 0)       # larger than the domain (cpu_ntt_main.h:39-41)
    ib.ntt_release_domain(F)
