"""Developer probe (GPU box): timings of the widened entry points -- quartic-extension NTT, columns-batch NTT on the 4-byte
fields (transposed path vs the strided register-only schedule it replaces), ECNTT."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import icicle_b200 as ib
from icicle_b200 import utils
import common

ib.set_device(0)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


F = ib.Field.BABYBEAR
fp = utils.field_params("babybear")
p = fp["p"]
dom = 24
ib.ntt_init_domain(F, utils.to_limbs([pow(fp["rou"], 1 << (fp["two_adicity"] - dom), p)], 1)[0])
for logn, batch in ((20, 16), (24, 1), (22, 8)):
    n = 1 << logn
    x = torch.randint(0, p, (n * batch * 4,), dtype=torch.int64, device="cuda").to(torch.int32)
    y = ib.device_empty(n * batch * 4)
    ms = timeit(lambda: ib.ntt_extension(F, x, n, 0, ib.NTTConfig(batch_size=batch, is_async=True), y))
    xb = x[: n * batch * 4]
    ms_base = timeit(lambda: ib.ntt(F, xb, n, 0, ib.NTTConfig(batch_size=batch * 4, is_async=True), y))
    print(f"babybear ext4 ntt 2^{logn} x {batch}: {ms:8.3f} ms  {n*batch/ms*1e-6:7.2f} G ext-elem/s ({16*2*n*batch/ms*1e-6:7.1f} GB/s alg.)   [same data as {batch*4} base rows: {ms_base:.3f} ms]", flush=True)
for logn, cols in ((20, 64), (16, 256)):
    n = 1 << logn
    x = torch.randint(0, p, (n * cols,), dtype=torch.int64, device="cuda").to(torch.int32)
    y = ib.device_empty(n * cols)
    ms_t = timeit(lambda: ib.ntt(F, x, n, 0, ib.NTTConfig(batch_size=cols, columns_batch=True, is_async=True), y))
    ib.set_tuning("ntt_columns_strided", 1)
    ms_s = timeit(lambda: ib.ntt(F, x, n, 0, ib.NTTConfig(batch_size=cols, columns_batch=True, is_async=True), y))
    ib.set_tuning("ntt_columns_strided", None)
    ms_r = timeit(lambda: ib.ntt(F, x, n, 0, ib.NTTConfig(batch_size=cols, is_async=True), y))
    print(f"babybear columns_batch ntt 2^{logn} x {cols}: transposed {ms_t:8.3f} ms ({n*cols/ms_t*1e-6:6.2f} G elem/s) | strided {ms_s:8.3f} ms | row-major batch {ms_r:8.3f} ms", flush=True)
ib.ntt_release_domain(F)

C_, Fr = ib.Curve.BN254_G1, ib.Field.BN254_FR
fr = utils.field_params("bn254_fr")
ib.ntt_init_domain(Fr, utils.to_limbs([pow(fr["rou"], 1 << (fr["two_adicity"] - 14), fr["p"])], 8)[0])
base = common.affine_to_projective_limbs(common.gen_g1_points("bn254", 1 << 10, 3), 8)
for logn in (10, 12, 14):
    n = 1 << logn
    P = ib.to_device(np.tile(base, (n >> 10, 1)))
    out = ib.device_empty(n * 24)
    ms = timeit(lambda: ib.ecntt(C_, P, n, 0, ib.NTTConfig(is_async=True), out), reps=2)
    print(f"bn254 ecntt 2^{logn}: {ms:9.3f} ms  ({n/2*logn/ms*1e-3:7.2f} M butterflies/s)", flush=True)
ib.ntt_release_domain(Fr)
