"""Batch-sharded multi-GPU run of BASELINE configs 4 and 5 (scaled): one process per GPU (torchrun), no data-path collective.
  config 4: BLS12-381 G1 MSM, batch = 8 with shared bases  -> each rank computes batch/W of the MSMs
  config 5: BabyBear NTT batch                              -> each rank transforms batch/W rows
Rank 0 gathers the per-rank results (NCCL all_gather of the small MSM results; NTT outputs stay sharded) and checks them
against the same batch computed on one GPU.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/multi_gpu_batch.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
import icicle_b200 as ib
from icicle_b200 import utils
from icicle_b200 import shard_range as shard_batch
import common

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); ib.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)

def timed(fn):
    if world > 1: dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1: dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

# ---- config 4 (scaled): BLS12-381 G1 MSM 2^20, batch 8, shared bases ----
C, logn, batch = ib.Curve.BLS12_381_G1, 20, 8
n = 1 << logn
base = ib.to_device(common.gen_g1_points("bls12_381", 1 << 10, 5), dev)
P = base.repeat(n >> 10, 1).contiguous()
g = torch.Generator(device=dev); g.manual_seed(11)                      # identical scalars on every rank
S = torch.randint(-2**31, 2**31, (batch * n, 8), dtype=torch.int64, device=dev, generator=g).to(torch.int32)
S[:, 7] &= 0x3FFFFFFF
S = S.contiguous()
lo, hi = shard_batch(batch, world, rank)
mine = ib.device_empty((hi - lo) * 36, dev).view(hi - lo, 36)
run = lambda: ib.msm(C, S[lo * n:hi * n], P, n, ib.MSMConfig(batch_size=hi - lo, is_async=True), mine)
run(); t_msm = timed(run)
if world > 1:
    assert batch % world == 0
    allres = ib.device_empty(batch * 36, dev).view(batch, 36)
    dist.all_gather_into_tensor(allres.view(-1), mine.view(-1))
else:
    allres = mine
if rank == 0:
    full = ib.msm(C, S, P, n, ib.MSMConfig(batch_size=batch))
    q = utils.field_params("bls12_381_fq")["p"]
    got = ib.to_host(allres)
    for b in range(batch):
        assert common.projective_to_affine_ints(got[b], 12, q) == common.projective_to_affine_ints(full[b], 12, q), b
    print(f"[config 4 scaled] BLS12-381 G1 MSM 2^{logn} x batch {batch} on {world} GPU(s): {t_msm:.2f} ms  {batch * n / t_msm / 1e3:.1f} Mpts/s  (sharded results == single-GPU results)")

# ---- config 5 (scaled): BabyBear NTT 2^24, batch 16 ----
F, fp = ib.Field.BABYBEAR, utils.field_params("babybear")
logn, batch = 24, 16
n = 1 << logn
ib.ntt_init_domain(F, utils.to_limbs([pow(fp["rou"], 1 << (fp["two_adicity"] - logn), fp["p"])], 1)[0])
g.manual_seed(12)
X = torch.randint(0, fp["p"], (batch * n,), dtype=torch.int64, device=dev, generator=g).to(torch.int32).contiguous()
lo, hi = shard_batch(batch, world, rank)
Y = ib.device_empty((hi - lo) * n, dev)
run = lambda: ib.ntt(F, X[lo * n:hi * n], n, ib.NTTDir.kForward, ib.NTTConfig(batch_size=hi - lo, is_async=True), Y)
run(); t_ntt = timed(run)
# parity of this rank's rows against a single full-batch call on the same GPU (inputs are identical on every rank)
Yfull = ib.device_empty(batch * n, dev)
ib.ntt(F, X, n, ib.NTTDir.kForward, ib.NTTConfig(batch_size=batch), Yfull)
assert torch.equal(Y, Yfull[lo * n:hi * n])
ok = torch.tensor([1], device=dev)
if world > 1: dist.all_reduce(ok, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"[config 5 scaled] BabyBear NTT 2^{logn} x batch {batch} on {world} GPU(s): {t_ntt:.2f} ms  {batch * n / t_ntt / 1e6:.1f} Gelem/s  (every rank's rows == single-call rows: {bool(ok.item())})")
ib.ntt_release_domain(F)
if world > 1: dist.destroy_process_group()
