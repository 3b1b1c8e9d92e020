"""Developer probe (GPU box): BabyBear 2^27 x 2 and BN254 2^24 forward NTT timings under tuning knobs (round 2)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import icicle_b200 as ib
from icicle_b200 import utils


def run(name, F, logn, batch, L, knobs):
    fp = utils.field_params(name)
    ib.ntt_release_domain(F)
    ib.ntt_init_domain(F, utils.to_limbs([pow(fp["rou"], 1 << (fp["two_adicity"] - logn), fp["p"])], L)[0])
    n = 1 << logn
    if L == 1:
        x = torch.randint(0, fp["p"], (n * batch,), dtype=torch.int64, device="cuda").to(torch.int32)
    else:
        x = torch.randint(-2**31, 2**31, (n * batch, L), dtype=torch.int64, device="cuda").to(torch.int32)
        x[:, L - 1] = torch.randint(0, 0x30000000, (n * batch,), dtype=torch.int64, device="cuda").to(torch.int32)
        x = x.contiguous()
    y = ib.device_empty(n * batch * L)
    for k, v in knobs.items():
        ib.set_tuning(k, v)
    for o, nm in ((ib.Ordering.kNN, "kNN"), (ib.Ordering.kNR, "kNR")):
        for d in (0, 1):
            for _ in range(3):
                ib.ntt(F, x, n, d, ib.NTTConfig(batch_size=batch, is_async=True, ordering=o), y)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                ib.ntt(F, x, n, d, ib.NTTConfig(batch_size=batch, is_async=True, ordering=o), y)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"{name} 2^{logn} x {batch} {nm} dir={d} knobs={knobs}: {ms:.3f} ms  {n * batch / ms / 1e6:.2f} G elem/s", flush=True)
    for k in knobs:
        ib.set_tuning(k, None)
    ib.ntt_release_domain(F)


run("babybear", ib.Field.BABYBEAR, 27, 2, 1, {})
run("babybear", ib.Field.BABYBEAR, 27, 2, 1, {"ntt31_tma_off": 1})
run("babybear", ib.Field.BABYBEAR, 24, 16, 1, {})
run("babybear", ib.Field.BABYBEAR, 24, 16, 1, {"ntt31_tma_off": 1})
run("bn254_fr", ib.Field.BN254_FR, 24, 1, 8, {})
run("bn254_fr", ib.Field.BN254_FR, 24, 1, 8, {"ntt_geom": 28})
