"""Developer probe (GPU box): per-stage CUDA-event times of one MSM / NTT call (b200_set_profiling)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import icicle_b200 as ib
from icicle_b200 import utils
import common

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 24
cs = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
levels = sys.argv[3].split(",") if len(sys.argv) > 3 else [None]
n = 1 << logn
base = ib.to_device(common.gen_g1_points("bn254", 1 << 12, 1))
P = base.repeat(n >> 12, 1).contiguous()
s = torch.randint(-2**31, 2**31, (n, 8), dtype=torch.int64, device="cuda").to(torch.int32)
s[:, 7] = torch.randint(0, 0x30644E72, (n,), dtype=torch.int64, device="cuda").to(torch.int32)
s = s.contiguous()
out = ib.device_empty(24).view(1, 24)
ib.set_profiling(True)
for c, lv in [(c, lv) for c in cs for lv in levels]:
    if lv is not None:
        ib.set_tuning("msm_pair_levels", int(lv))
    for rep in range(2):
        ib.msm(ib.Curve.BN254_G1, s, P, n, ib.MSMConfig(c=c, is_async=True), out)
    what, st = ib.last_profile()
    tot = sum(ms for _, ms in st)
    print(f"msm 2^{logn} levels={lv} c={c or ib.msm_choose_c(ib.Curve.BN254_G1, n)}: total {tot:.3f} ms | " + " ".join(f"{k}={v:.3f}" for k, v in st), flush=True)
