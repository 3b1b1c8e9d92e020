"""Developer probe (GPU box) for compute-sanitizer --tool racecheck / memcheck: one small call of every shared-memory kernel family
(MSM with forced pair levels and the counting sort, BN254 tile NTT, BabyBear TMA-staged NTT in both schedules, reductions)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import icicle_b200 as ib
from icicle_b200 import utils
import common

ib.set_device(0)
n = 1 << 11
P = common.gen_g1_points("bn254", n, 3)
s = common.seeded_scalars("bn254_fr", n, 4)
ib.set_tuning("msm_pair_levels", 2)
a = ib.msm(ib.Curve.BN254_G1, s, P, n, ib.MSMConfig(c=8))
ib.set_tuning("msm_pair_levels", 0)
b = ib.msm(ib.Curve.BN254_G1, s, P, n, ib.MSMConfig(c=8))
q = utils.field_params("bn254_fq")["p"]
assert common.projective_to_affine_ints(a[0], 8, q) == common.projective_to_affine_ints(b[0], 8, q)
ib.set_tuning("msm_pair_levels", None)
for fname, F, L, logn in (("bn254_fr", ib.Field.BN254_FR, 8, 12), ("babybear", ib.Field.BABYBEAR, 1, 15)):
    fp = utils.field_params(fname)
    ib.ntt_init_domain(F, utils.to_limbs([pow(fp["rou"], 1 << (fp["two_adicity"] - logn), fp["p"])], L)[0])
    x = common.seeded_scalars(fname, 1 << logn, 9)
    for o in (ib.Ordering.kNN, ib.Ordering.kNR):
        y = ib.ntt(F, x, 1 << logn, 0, ib.NTTConfig(ordering=o))
        inv_o = ib.Ordering.kNN if o == ib.Ordering.kNN else ib.Ordering.kRN
        assert np.array_equal(ib.ntt(F, y, 1 << logn, 1, ib.NTTConfig(ordering=inv_o)), x)
    ib.vector_sum(F, x, 1 << logn)
    ib.ntt_release_domain(F)
print("race_probe ok")
