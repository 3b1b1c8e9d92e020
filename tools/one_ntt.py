"""Developer probe (GPU box): a few device-resident forward NTTs for ncu: `bn254 24 1` or `babybear 27 2` (field, logn, batch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import icicle_b200 as ib
from icicle_b200 import utils
name, logn, batch = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
F = {"bn254": ib.Field.BN254_FR, "babybear": ib.Field.BABYBEAR}[name]
fname = {"bn254": "bn254_fr", "babybear": "babybear"}[name]
fp = utils.field_params(fname)
L = fp["limbs"]
ib.set_device(0)
ib.ntt_init_domain(F, utils.to_limbs([pow(fp["rou"], 1 << (fp["two_adicity"] - logn), fp["p"])], L)[0])
n = 1 << logn
if L == 1:
    x = torch.randint(0, fp["p"], (n * batch,), dtype=torch.int64, device="cuda").to(torch.int32)
else:
    x = torch.randint(-2**31, 2**31, (n * batch, L), dtype=torch.int64, device="cuda").to(torch.int32)
    x[:, L - 1] = torch.randint(0, 0x30000000, (n * batch,), dtype=torch.int64, device="cuda").to(torch.int32)
    x = x.contiguous()
y = ib.device_empty(n * batch * L)
for _ in range(2):
    ib.ntt(F, x, n, 0, ib.NTTConfig(batch_size=batch, is_async=True), y)
torch.cuda.synchronize()
print("done")
