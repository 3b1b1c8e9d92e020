"""Parity at BASELINE.json's FULL sizes through size-independent properties (the CPU oracle would need minutes to hours):
  * BN254 G1 MSM 2^26: linearity MSM(s,P) + MSM(t,P) == MSM(s+t,P), and agreement of the device path with a 4-way
    point-range split summed with ec_sum (the multi-GPU combine step on one GPU);
  * BN254 NTT 2^24: inverse(forward(x)) == x bit-exactly for kNN and kNR/kRN, linearity NTT(x+y) == NTT(x)+NTT(y), and the
    defining sum at two output indices checked with Python integers on a sparse input."""
import numpy as np
import pytest

import icicle_b200 as ib
from icicle_b200 import utils
import common

pytestmark = pytest.mark.gpu


def _rand_scalars_dev(n, seed):
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    s = torch.randint(-2 ** 31, 2 ** 31, (n, 8), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    s[:, 7] = torch.randint(0, 0x30644E72, (n,), dtype=torch.int64, device="cuda", generator=g).to(torch.int32)
    return s.contiguous()


def test_msm_2p26_linearity_and_sharded_combine():
    import torch
    C = ib.Curve.BN254_G1
    logn = 26
    n = 1 << logn
    base = ib.to_device(common.gen_g1_points("bn254", 1 << 12, 2026))
    P = base.repeat(n >> 12, 1).contiguous()
    s, t = _rand_scalars_dev(n, 1), _rand_scalars_dev(n, 2)
    st = ib.device_empty(n * 8).view(n, 8)
    ib.vector_add(ib.Field.BN254_FR, s, t, n, ib.VecOpsConfig(), st)
    q = utils.field_params("bn254_fq")["p"]
    aff = lambda r: common.projective_to_affine_ints(r, 8, q)
    A = aff(ib.msm(C, s, P, n)[0])
    B = aff(ib.msm(C, t, P, n)[0])
    AB = aff(ib.msm(C, st, P, n)[0])
    assert common.ec_add(A, B, q) == AB
    # point-range sharding (what each rank of the multi-GPU path computes) + ec_sum == the single MSM
    parts = ib.device_empty(4 * 24).view(4, 24)
    quarter = n // 4
    for k in range(4):
        ib.msm(C, s[k * quarter:(k + 1) * quarter], P[k * quarter:(k + 1) * quarter], quarter, ib.MSMConfig(), parts[k:k + 1])
    total = ib.ec_sum(C, parts, 4)
    assert aff(total[0]) == A
    del P, s, t, st
    torch.cuda.empty_cache()


def test_ntt_2p24_roundtrip_linearity_definition():
    import torch
    F = ib.Field.BN254_FR
    fp = utils.field_params("bn254_fr")
    p = fp["p"]
    logn = 24
    n = 1 << logn
    w = pow(fp["rou"], 1 << (fp["two_adicity"] - logn), p)
    ib.ntt_release_domain(F)
    ib.ntt_init_domain(F, utils.to_limbs([w], 8)[0])
    x, y = _rand_scalars_dev(n, 3), _rand_scalars_dev(n, 4)
    X = ib.device_empty(n * 8).view(n, 8)
    back = ib.device_empty(n * 8).view(n, 8)
    for fwd_o, inv_o in ((ib.Ordering.kNN, ib.Ordering.kNN), (ib.Ordering.kNR, ib.Ordering.kRN), (ib.Ordering.kNM, ib.Ordering.kMN)):
        ib.ntt(F, x, n, ib.NTTDir.kForward, ib.NTTConfig(ordering=fwd_o), X)
        ib.ntt(F, X, n, ib.NTTDir.kInverse, ib.NTTConfig(ordering=inv_o), back)
        assert torch.equal(back, x), (fwd_o, inv_o)
    # linearity
    xy = ib.device_empty(n * 8).view(n, 8)
    ib.vector_add(F, x, y, n, ib.VecOpsConfig(), xy)
    Y, XY = ib.device_empty(n * 8).view(n, 8), ib.device_empty(n * 8).view(n, 8)
    ib.ntt(F, x, n, ib.NTTDir.kForward, ib.NTTConfig(), X)
    ib.ntt(F, y, n, ib.NTTDir.kForward, ib.NTTConfig(), Y)
    ib.ntt(F, xy, n, ib.NTTDir.kForward, ib.NTTConfig(), XY)
    S = ib.device_empty(n * 8).view(n, 8)
    ib.vector_add(F, X, Y, n, ib.VecOpsConfig(), S)
    assert torch.equal(S, XY)
    # defining sum on a sparse input: x = e_a*alpha + e_b*beta  =>  X[k] = alpha*w^(a k) + beta*w^(b k)
    a_idx, b_idx, alpha, beta = 12345, n - 7, 0x1234567890ABCDEF, 3
    sp = torch.zeros((n, 8), dtype=torch.int32, device="cuda")
    sp[a_idx] = torch.from_numpy(utils.to_limbs([alpha], 8)[0].astype(np.int32)).cuda()
    sp[b_idx] = torch.from_numpy(utils.to_limbs([beta], 8)[0].astype(np.int32)).cuda()
    ib.ntt(F, sp, n, ib.NTTDir.kForward, ib.NTTConfig(), X)
    host = ib.to_host(X)
    for k in (0, 1, 999999, n - 1):
        exp = (alpha * pow(w, a_idx * k, p) + beta * pow(w, b_idx * k, p)) % p
        assert utils.from_limbs(host[k:k + 1])[0] == exp, k
    ib.ntt_release_domain(F)
    del x, y, X, Y, XY, S, xy, back, sp
    torch.cuda.empty_cache()
