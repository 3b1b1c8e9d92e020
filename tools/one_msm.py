"""Developer probe (GPU box): `reps` device-resident BN254 G1 MSMs of 2^logn points (bench.py's synthetic inputs), for ncu."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import icicle_b200 as ib
import bench
logn = int(sys.argv[1]) if len(sys.argv) > 1 else 26
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
ib.set_device(0)
s, P = bench.synth_inputs(torch, ib, logn, 0, torch.device("cuda", 0))
out = ib.device_empty(24).view(1, 24)
for _ in range(reps):
    ib.msm(ib.Curve.BN254_G1, s, P, 1 << logn, ib.MSMConfig(is_async=True), out)
torch.cuda.synchronize()
print("done", reps)
