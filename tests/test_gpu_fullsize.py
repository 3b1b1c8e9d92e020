"""Parity at BASELINE.json's FULL sizes.

Against the UNMODIFIED reference CPU backend (oracle/_ref) on IDENTICAL bytes, as icicle/tests/test_curve_api.cpp:81-171 and
test_mod_arithmetic_api.h:614-695 do (MSM: equal group elements; NTT: memcmp):
  * BN254 G1 MSM 2^20, 2^22, 2^24 (device-resident and host-pointer paths) and 2^26 (host pointers, pageable -> the copier
    ring + chunk pipeline; ~2 min of CPU for the reference);
  * BN254 NTT 2^16 .. 2^24 forward and inverse, kNN and kNR (the 3-pass autosort schedule of the headline 2^24 NTT included);
  * BLS12-381 G1 and G2 MSM 2^20, BabyBear NTT 2^24 and 2^27 -- other reference builds, asked through tests/ref_worker.py
    (one reference family per process).
Beyond what the CPU finishes in minutes (MSM 2^27 / 2^28), size-independent properties: the chunked device-resident MSM must
equal the ec_sum of its oracle-checked 2^26 quarters, plus linearity."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import pytest

import icicle_b200 as ib
from icicle_b200 import utils
import common

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def ref():
    ref_icicle = pytest.importorskip("ref_icicle")
    if not ref_icicle.available("bn254"):
        pytest.skip("oracle/_ref/bn254 not present")
    return ref_icicle.get("bn254")


def _scratch_dir():
    d = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
    return tempfile.mkdtemp(prefix="b200_ref_", dir=d)


def _worker(*args):
    import ref_icicle
    family = args[1]
    if not ref_icicle.available(family):
        pytest.skip(f"oracle/_ref/{family} not present")
    subprocess.run([sys.executable, os.path.join(HERE, "ref_worker.py")] + [str(a) for a in args], check=True, timeout=1500)


def _free():
    import torch
    torch.cuda.empty_cache()
    ib.trim_scratch(0)


@pytest.mark.parametrize("logn", [20, 22, 24])
def test_bn254_msm_vs_reference(ref, logn):
    n = 1 << logn
    s = common.seeded_scalars("bn254_fr", n, 100 + logn)
    P = common.tiled_g1_points("bn254", n, 1 << 14, 200 + logn)
    exp = ref.msm(s, P, n)
    got_host = ib.msm(ib.Curve.BN254_G1, s, P, n)                      # host pointers (pageable): chunk pipeline from 2^23
    assert ref.projective_eq(got_host[0], exp[0]), ("host", logn)
    ds, dP = ib.to_device(s), ib.to_device(P)
    got_dev = ib.to_host(ib.msm(ib.Curve.BN254_G1, ds, dP, n, ib.MSMConfig(are_results_on_device=True)))
    assert ref.projective_eq(got_dev[0], exp[0]), ("device", logn)
    del ds, dP
    _free()


def test_bn254_msm_2p26_vs_reference(ref):
    """The headline size against the reference on identical bytes (the reference needs ~1-2 min of CPU at 2^26)."""
    n = 1 << 26
    s = common.seeded_scalars("bn254_fr", n, 126)
    P = common.tiled_g1_points("bn254", n, 1 << 14, 226)
    got = ib.msm(ib.Curve.BN254_G1, s, P, n)
    ds, dP = ib.to_device(s), ib.to_device(P)
    got_dev = ib.to_host(ib.msm(ib.Curve.BN254_G1, ds, dP, n, ib.MSMConfig(are_results_on_device=True)))
    del ds, dP
    _free()
    exp = ref.msm(s, P, n)
    assert ref.projective_eq(got[0], exp[0])
    assert ref.projective_eq(got_dev[0], exp[0])


def test_bn254_msm_2p27_2p28_chunked_device_path(ref):
    """Device-resident MSMs beyond 2^30 bucket entries run in point-range chunks into one bucket array (msm_chunked): the
    2^28 result must equal the sum of its four 2^26 quarters (the size checked against the reference above), the 2^27 result
    the sum of the first two, and the first quarter is compared with the reference directly."""
    import torch
    C = ib.Curve.BN254_G1
    q = utils.field_params("bn254_fq")["p"]
    aff = lambda r: common.projective_to_affine_ints(r, 8, q)
    n = 1 << 28
    quarter = n // 4
    base = ib.to_device(common.gen_g1_points("bn254", 1 << 14, 2028))
    P = base.repeat(n >> 14, 1).contiguous()
    g = torch.Generator(device="cuda")
    g.manual_seed(28)
    s = torch.randint(-2 ** 31, 2 ** 31, (n, 8), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    s[:, 7] = torch.randint(0, 0x30644E72, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    s = s.contiguous()
    parts = ib.device_empty(4 * 24).view(4, 24)
    for k in range(4):
        ib.msm(C, s[k * quarter:(k + 1) * quarter], P[k * quarter:(k + 1) * quarter], quarter, ib.MSMConfig(), parts[k:k + 1])
    full28 = aff(ib.msm(C, s, P, n)[0])
    assert full28 == aff(ib.ec_sum(C, parts, 4)[0])
    full27 = aff(ib.msm(C, s[: n // 2], P[: n // 2], n // 2)[0])
    assert full27 == aff(ib.ec_sum(C, parts[:2], 2)[0])
    # first 2^22 points of the same data against the reference
    m = 1 << 22
    hs, hP = ib.to_host(s[:m]), ib.to_host(P[:m])
    assert ref.projective_eq(ib.msm(C, s[:m], P[:m], m)[0], ref.msm(hs, hP, m)[0])
    del P, s, parts
    _free()


def test_bn254_ntt_vs_reference_up_to_2p24(ref):
    import torch
    F = ib.Field.BN254_FR
    top = 24
    root = ref.get_root_of_unity(1 << top)
    ref.ntt_release_domain()
    ref.ntt_init_domain(root)
    ib.ntt_release_domain(F)
    ib.ntt_init_domain(F, root)
    x_all = common.seeded_scalars("bn254_fr", 1 << top, 77)
    for logn in (16, 20, 22, 24):
        n = 1 << logn
        x = x_all[:n]
        dx = ib.to_device(x)
        for d in (0, 1):
            for o in (ib.Ordering.kNN, ib.Ordering.kNR):
                exp = ref.ntt(x, n, d, ordering=int(o))
                got = ib.ntt(F, x, n, d, ib.NTTConfig(ordering=o))                                  # host pointers
                assert np.array_equal(got, exp), ("host", logn, d, o)
                dy = ib.device_empty(n * 8).view(n, 8)
                ib.ntt(F, dx, n, d, ib.NTTConfig(ordering=o), dy)                                   # device resident
                assert np.array_equal(ib.to_host(dy), exp), ("device", logn, d, o)
        del dx
    ref.ntt_release_domain()
    ib.ntt_release_domain(F)
    _free()


@pytest.mark.parametrize("g2", [0, 1])
def test_bls12_381_msm_2p20_vs_reference(g2):
    """BASELINE config 4's curve at 2^20, G1 and G2, against the bls12_381 reference build (subprocess)."""
    d = _scratch_dir()
    pre = os.path.join(d, f"bls_{g2}")
    _worker("msm", "bls12_381", 20, g2, 381 + g2, pre)
    s, P, exp = np.load(pre + "_scalars.npy"), np.load(pre + "_points.npy"), np.load(pre + "_expected_affine.npy")
    n, L = 1 << 20, 12
    q = utils.field_params("bls12_381_fq")["p"]
    curve = ib.Curve.BLS12_381_G2 if g2 else ib.Curve.BLS12_381_G1
    for src in ("host", "device"):
        if src == "host":
            got = ib.msm(curve, s, P, n)[0]
        else:
            got = ib.to_host(ib.msm(curve, ib.to_device(s), ib.to_device(P), n, ib.MSMConfig(are_results_on_device=True)))[0]
        if g2:
            nr = utils.curve_params("bls12_381")["nonresidue"]
            assert common.fq2_projective_to_affine_ints(got, L, q, nr) == utils.from_limbs(exp.reshape(4, L)), src
        else:
            assert common.projective_to_affine_ints(got, L, q) == common.affine_limbs_to_ints(exp, L)[0], src
    # batch 8 with shared bases (config 4's shape) == eight single MSMs
    sb = np.concatenate([np.roll(s[: 1 << 16], k, axis=0) for k in range(8)])
    got8 = ib.msm(curve, sb, P[: 1 << 16], 1 << 16, ib.MSMConfig(batch_size=8))
    for k in (0, 5, 7):
        one = ib.msm(curve, sb[k << 16:(k + 1) << 16], P[: 1 << 16], 1 << 16)
        conv = (lambda r: common.fq2_projective_to_affine_ints(r, L, q, utils.curve_params("bls12_381")["nonresidue"])) if g2 else (lambda r: common.projective_to_affine_ints(r, L, q))
        assert conv(got8[k]) == conv(one[0]), k
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
    os.rmdir(d)
    _free()


@pytest.mark.parametrize("logn,batch", [(24, 2), (27, 1)])
def test_babybear_ntt_vs_reference_full_size(logn, batch):
    """BASELINE config 5's field at 2^24 x 2 and at its maximum size 2^27 against the babybear reference build (subprocess),
    forward kNN and inverse kNR, bit-exact on the whole array."""
    F = ib.Field.BABYBEAR
    n = 1 << logn
    d = _scratch_dir()
    # the reference needs ~2 min of CPU per 2^27 transform: forward kNN only at the maximum size, both directions at 2^24
    for direction, ordering in (((0, 0),) if logn >= 27 else ((0, 0), (1, 1))):
        pre = os.path.join(d, f"bb_{logn}_{direction}")
        _worker("ntt", "babybear", logn, batch, direction, ordering, 900 + logn, pre)
        root, exp = np.load(pre + "_root.npy"), np.load(pre + "_expected.npy")
        x = common.seeded_scalars("babybear", n * batch, 900 + logn)
        ib.ntt_release_domain(F)
        ib.ntt_init_domain(F, root)
        got = ib.ntt(F, x, n, direction, ib.NTTConfig(batch_size=batch, ordering=ib.Ordering(ordering)))
        assert np.array_equal(got, exp), (logn, direction)
        dx = ib.to_device(x)
        dy = ib.device_empty(n * batch)
        ib.ntt(F, dx, n, direction, ib.NTTConfig(batch_size=batch, ordering=ib.Ordering(ordering)), dy)
        assert np.array_equal(ib.to_host(dy).reshape(-1, 1), exp), (logn, direction, "device")
        del dx, dy
        os.remove(pre + "_expected.npy")
    ib.ntt_release_domain(F)
    for f in os.listdir(d):
        os.remove(os.path.join(d, f))
    os.rmdir(d)
    _free()


def test_msm_2p26_linearity():
    """MSM(s,P) + MSM(t,P) == MSM(s+t,P) at the headline size (device resident)."""
    import torch
    C = ib.Curve.BN254_G1
    n = 1 << 26
    base = ib.to_device(common.gen_g1_points("bn254", 1 << 12, 2026))
    P = base.repeat(n >> 12, 1).contiguous()
    mk = lambda seed: ib.to_device(common.seeded_scalars("bn254_fr", n, seed))
    s, t = mk(1), mk(2)
    st = ib.device_empty(n * 8).view(n, 8)
    ib.vector_add(ib.Field.BN254_FR, s, t, n, ib.VecOpsConfig(), st)
    q = utils.field_params("bn254_fq")["p"]
    aff = lambda r: common.projective_to_affine_ints(r, 8, q)
    A, B, AB = aff(ib.msm(C, s, P, n)[0]), aff(ib.msm(C, t, P, n)[0]), aff(ib.msm(C, st, P, n)[0])
    assert common.ec_add(A, B, q) == AB
    del P, s, t, st
    _free()
