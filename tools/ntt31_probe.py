"""Developer probe (GPU box): a few BabyBear NTT calls for ncu (python tools/ntt31_probe.py <logn> <batch> <reps>)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import icicle_b200 as ib
from icicle_b200 import utils
logn, batch, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
fp = utils.field_params("babybear")
F = ib.Field.BABYBEAR
ib.ntt_init_domain(F, utils.to_limbs([pow(fp["rou"], 1 << (fp["two_adicity"] - logn), fp["p"])], 1)[0])
n = 1 << logn
x = torch.randint(0, fp["p"], (n * batch,), dtype=torch.int64, device="cuda").to(torch.int32).contiguous()
y = ib.device_empty(n * batch)
for _ in range(reps):
    ib.ntt(F, x, n, ib.NTTDir.kForward, ib.NTTConfig(batch_size=batch, is_async=True), y)
torch.cuda.synchronize()
