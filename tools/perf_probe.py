"""Developer probe (GPU box): time MSM and NTT at a few sizes with CUDA events; prints one line per case."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import icicle_b200 as ib
from icicle_b200 import utils
import common


def ev_time(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sorted(ts)[len(ts) // 2]


def rand_scalars_dev(n, top_mask=0x0FFFFFFF):
    t = torch.randint(-2**31, 2**31, (n, 8), dtype=torch.int64, device="cuda").to(torch.int32)
    t[:, 7] &= top_mask
    return t.contiguous()


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    max_log = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    if which in ("all", "msm"):
        base = ib.to_device(common.gen_g1_points("bn254", 1 << 12, 1))
        for logn in range(16, max_log + 1, 2):
            n = 1 << logn
            P = base.repeat(n >> 12, 1).contiguous()
            s = rand_scalars_dev(n)
            out = ib.device_empty(24).view(1, 24)
            cs = [0] if logn < 20 else [0, 14, 16, 18, 20, 22]
            for c in cs:
                cfg = lambda: ib.MSMConfig(c=c, is_async=True)
                try:
                    best, med = ev_time(lambda: ib.msm(ib.Curve.BN254_G1, s, P, n, cfg(), out), reps=3)
                except Exception as e:
                    print(f"msm 2^{logn} c={c}: {e}"); continue
                cc = c or ib.msm_choose_c(ib.Curve.BN254_G1, n)
                print(f"msm bn254 2^{logn} c={cc:2d}{'(auto)' if not c else '      '} best {best:9.3f} ms  med {med:9.3f} ms  {n / best / 1e3:9.2f} Mpts/s", flush=True)
            del P, s
    if which in ("all", "ntt"):
        fp = utils.field_params("bn254_fr")
        dom = min(max_log, 26)
        ib.ntt_init_domain(ib.Field.BN254_FR, utils.to_limbs([pow(fp["rou"], 1 << (fp["two_adicity"] - dom), fp["p"])], 8)[0])
        for logn in range(12, dom + 1, 2):
            n = 1 << logn
            batch = max(1, (1 << 22) >> logn)
            x = rand_scalars_dev(n * batch)
            y = ib.device_empty(n * batch * 8)
            for ordering in (ib.Ordering.kNN, ib.Ordering.kNR):
                for d in (ib.NTTDir.kForward, ib.NTTDir.kInverse):
                    cfg = lambda: ib.NTTConfig(batch_size=batch, ordering=ordering, is_async=True)
                    best, med = ev_time(lambda: ib.ntt(ib.Field.BN254_FR, x, n, d, cfg(), y), reps=5)
                    print(f"ntt bn254 2^{logn} x{batch:5d} {ordering.name} {d.name:9s} best {best:8.3f} ms  {n * batch / best / 1e6:8.3f} Gelem/s  {n*batch*64/best/1e6:8.1f} GB/s(alg)", flush=True)
        ib.ntt_release_domain(ib.Field.BN254_FR)
    if which in ("all", "bb"):
        fp = utils.field_params("babybear")
        F = ib.Field.BABYBEAR
        dom = min(max_log, 27)
        ib.ntt_init_domain(F, utils.to_limbs([pow(fp["rou"], 1 << (fp["two_adicity"] - dom), fp["p"])], 1)[0])
        for logn in range(16, dom + 1, 2 if dom % 2 == 0 else 1):
            if logn not in (16, 20, 24, dom): continue
            n = 1 << logn
            batch = max(1, (1 << 28) >> logn)
            x = torch.randint(0, 0x78000001, (n * batch,), dtype=torch.int64, device="cuda").to(torch.int32).contiguous()
            y = ib.device_empty(n * batch)
            for ordering in (ib.Ordering.kNN, ib.Ordering.kNR):
                cfg = lambda: ib.NTTConfig(batch_size=batch, ordering=ordering, is_async=True)
                best, med = ev_time(lambda: ib.ntt(F, x, n, ib.NTTDir.kForward, cfg(), y), reps=5)
                print(f"ntt babybear 2^{logn} x{batch:5d} {ordering.name} fwd best {best:8.3f} ms  {n * batch / best / 1e6:8.3f} Gelem/s  {n*batch*8/best/1e6:8.1f} GB/s(alg)", flush=True)
        ib.ntt_release_domain(F)
    if which in ("all", "vec"):
        n = 1 << 24
        a, b = rand_scalars_dev(n), rand_scalars_dev(n)
        o = ib.device_empty(n * 8).view(n, 8)
        for nm, fn in (("add", ib.vector_add), ("mul", ib.vector_mul)):
            best, med = ev_time(lambda: fn(ib.Field.BN254_FR, a, b, n, ib.VecOpsConfig(is_async=True), o), reps=5)
            print(f"vec {nm} bn254 2^24 best {best:8.3f} ms  {n*96/best/1e6:8.1f} GB/s", flush=True)


if __name__ == "__main__":
    main()
