"""Round-1 measurements for the BASELINE.json configs that are parity cases rather than the headline bench line
(run on the GPU box; one GPU).  Prints a table; copied to profiles/ as evidence."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import icicle_b200 as ib
from icicle_b200 import utils
import common


def ev_time(fn, reps=3, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


def rand_dev(n, limbs, top_mask):
    t = torch.randint(-2**31, 2**31, (n, limbs), dtype=torch.int64, device="cuda").to(torch.int32)
    t[:, limbs - 1] &= top_mask
    return t.contiguous()


ib.set_device(0)
print("== config 4 (scaled to one GPU): BLS12-381 G1 / G2 MSM, batch 1 and batch 8 with shared bases ==")
for curve, name, g2 in ((ib.Curve.BLS12_381_G1, "bls12_381 G1", False), (ib.Curve.BLS12_381_G2, "bls12_381 G2", True)):
    L = 12
    if not g2:
        base = ib.to_device(common.gen_g1_points("bls12_381", 1 << 10, 3))
    else:
        g = np.load(os.path.join(ROOT, "tests", "golden", "bls12_381.npz"))
        base = ib.to_device(np.tile(g["g2_points"], (43, 1))[:1024])
    for logn, batch in ((20, 1), (22, 1), (20, 8)):
        n = 1 << logn
        P = base.repeat(n >> 10, 1).contiguous()
        s = rand_dev(n * batch, 8, 0x3FFFFFFF)
        out = ib.device_empty(batch * (72 if g2 else 36)).view(batch, -1)
        t = ev_time(lambda: ib.msm(curve, s, P, n, ib.MSMConfig(batch_size=batch, is_async=True), out), reps=2, warm=1)
        print(f"msm {name:13s} 2^{logn} batch {batch}: {t:9.3f} ms  {n * batch / t / 1e3:9.2f} Mpts/s  (c={ib.msm_choose_c(curve, n)})", flush=True)
        del P, s

print("== config 3: BN254 NTT forward/inverse, mixed radix (tiles) vs Radix2 key, device-resident ==")
fp = utils.field_params("bn254_fr")
F = ib.Field.BN254_FR
ib.ntt_init_domain(F, utils.to_limbs([pow(fp["rou"], 1 << (fp["two_adicity"] - 24), fp["p"])], 8)[0])
for logn in (12, 16, 20, 24):
    n = 1 << logn
    batch = max(1, (1 << 24) >> logn)
    x = rand_dev(n * batch, 8, 0x0FFFFFFF)
    y = ib.device_empty(n * batch * 8)
    for alg, nm in ((0, "mixed"), (1, "radix2")):
        if alg == 1 and logn > 20: continue
        for d in (ib.NTTDir.kForward, ib.NTTDir.kInverse):
            t = ev_time(lambda: ib.ntt(F, x, n, d, ib.NTTConfig(batch_size=batch, is_async=True, ext={"ntt_algorithm": alg}), y))
            print(f"ntt bn254 2^{logn} x{batch:5d} {nm:6s} {d.name:9s}: {t:8.3f} ms  {n * batch / t / 1e6:7.3f} Gelem/s", flush=True)
ib.ntt_release_domain(F)

print("== CPU reference (oracle/_ref, all host threads) for the same NTT ==")
try:
    import ref_icicle
    r = ref_icicle.get("bn254")
    for logn in (16, 20, 22):
        n = 1 << logn
        root = r.get_root_of_unity(n)
        r.ntt_release_domain(); r.ntt_init_domain(root)
        x = r.generate_scalars(n)
        r.ntt(x, n, 0)
        t0 = time.perf_counter(); r.ntt(x, n, 0); t = time.perf_counter() - t0
        print(f"cpu ntt bn254 2^{logn}: {t * 1e3:9.2f} ms  {n / t / 1e6:8.2f} Melem/s  ({os.cpu_count()} cores)", flush=True)
    r.ntt_release_domain()
except Exception as e:
    print("reference unavailable:", e)
