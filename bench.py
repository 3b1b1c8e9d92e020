#!/usr/bin/env python3
"""bench.py -- BN254 G1 MSM points/s @ 2^26 (headline) and BN254 NTT elements/s @ 2^24 (secondary), per BASELINE.json.

    python bench.py --gpus N --steps K --warmup W [--impl b200|reference] [--logn 26] [--ntt-logn 24]

A step = one pass of the hot path over one batch of synthetic input: one MSM over 2^logn (scalar, point) pairs per GPU.
  value       whole-job points/s with inputs resident in HBM (device-timed, CUDA events on the launching stream, max over ranks)
  e2e         the same metric through the reference-facing C ABI with HOST (pinned) buffers: H2D of scalars+points and
              D2H of the result inside the timed region
  roofline    dominant kernel (k_accumulate): algorithmic bytes (96 B/point) / its CUDA-event duration vs the
              measured HBM peak (MEASURED_PEAKS.json); the kernel is integer-multiply bound, see imad_frac
  cpu_baseline the UNMODIFIED reference CPU backend (oracle/_ref, built from /root/reference sources) on a bounded sample
N > 1 (torchrun): every rank runs the MSM over its own 2^logn-point shard of one (N * 2^logn)-point MSM; the partial results
are combined with ONE NCCL all-gather (96 B per rank) + the ec_sum kernel -> "scaling": "weak".
`--impl reference` times the reference's own CPU implementation (all host threads) on a bounded sample of the same
workload and prints the same JSON line with "impl": "reference".
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

ALG_BYTES_PER_POINT = 96  # |scalar_t| + |affine_t| for BN254 G1
ALG_BYTES_PER_NTT_ELEM = 64  # one read + one write of a 32-byte element
IMAD_WIDE_PEAK = 9.26e12  # measured on this pool's B200 (tools/imad_bench.cu, profiles/r1_imad_microbench.txt): IMAD.WIDE.U32.X thread-instr/s
IMAD_WIDE_PER_MADD = 10 * 140  # 8M+2S Montgomery products per mixed add x 140 IMAD(.WIDE) per 8-limb product (cuobjdump)
# dram__bytes_read.sum + dram__bytes_write.sum of the bucket-accumulation stage of ONE MSM at the headline config (2^26 points,
# c = 20, 5 pair levels): all k_pair_prefix / k_pair_apply / k_inv_* launches + k_accumulate, from the ncu launch list
# profiles/r1_ncu_launches_msm_2p26.txt (218.4 GB read + 89.4 GB write).  It is ~48x the algorithmic 6.44 GB: level 0 gathers
# every point once per window in each of its two passes and every level writes its halved list (planar scratch); the stage is
# bound by the random-sector rate of HBM at level 0 and by IMAD.WIDE issue above it, not by bytes.
NCU_TRAFFIC_BYTES = {(26, 20): 307.8e9}
# dram bytes of ONE forward BN254 NTT of 2^24 (3 k_ntt_tile passes; the first also reads the 0.5 GB twiddle table), from
# profiles/r1_ncu_ntt_passes.txt; algorithmic = 1.07 GB
NCU_NTT_TRAFFIC_BYTES = {24: 3.60e9}


def hbm_peak():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], 0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            p = [x.strip() for x in l.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                mx = max(mx, float(p[1]))
            except ValueError:
                continue
            for nm, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def pinned_array(ib, shape):
    n = int(np.prod(shape))
    ptr = C.c_void_p()
    ib.capi.check(ib.capi.lib.b200_host_alloc_pinned(C.byref(ptr), n * 4), "host_alloc_pinned")
    buf = (C.c_uint32 * n).from_address(ptr.value)
    return np.frombuffer(buf, dtype=np.uint32).reshape(shape), ptr


def synth_inputs(torch, ib, logn, seed, device):
    """2^logn uniform scalars (top limb below the modulus' top limb -> always < p) and 2^logn points = 2^16 DISTINCT curve
    points tiled (upstream's generator repeats only 100 points)."""
    import common
    n = 1 << logn
    distinct = min(n, 1 << 16)
    base = common.gen_g1_points("bn254", distinct, 1000 + seed)
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    s = torch.randint(-2 ** 31, 2 ** 31, (n, 8), dtype=torch.int64, device=device, generator=g).to(torch.int32)
    top = torch.randint(0, 0x30644E72, (n,), dtype=torch.int64, device=device, generator=g).to(torch.int32)
    s[:, 7] = top
    pts = ib.to_device(base).repeat(n // distinct, 1).contiguous()
    return s.contiguous(), pts


def run_reference_arm(args, rank, out_fd):
    """--impl reference: the reference's own CPU MSM (all host threads) on a bounded sample of the workload."""
    if rank != 0:
        return
    import ref_icicle
    r = ref_icicle.get("bn254")
    cores = os.cpu_count()
    sample_log = args.ref_sample_logn
    n = 1 << sample_log
    s = r.generate_scalars(n)
    P = r.generate_affine_points(n)
    for _ in range(max(1, min(args.warmup, 1))):
        r.msm(s, P, n)
    ts = []
    for _ in range(args.steps):
        t0 = time.perf_counter()
        r.msm(s, P, n)
        ts.append(time.perf_counter() - t0)
    t = sum(ts) / len(ts)
    val = n / t
    line = {
        "impl": "reference", "metric": "bn254_g1_msm_points_per_s", "value": val, "unit": "points/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32 limbs (254-bit modular integers)",
        "data": "synthetic (reference generators: uniform scalars, 100 distinct points repeated)",
        "config": {"workload": f"BN254 G1 MSM 2^{args.logn} (bounded sample: 2^{sample_log} points per step on the host cores)"},
        "cpu_baseline": {"value": val, "unit": "points/s", "cores": cores, "kind": "reference",
                         "sample": f"icicle CPU backend (oracle/_ref, g++ -O3, Taskflow stand-in) MSM of 2^{sample_log} points, mean of {args.steps}"},
        "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out_fd, line)


def cpu_baseline(budget_s=20.0):
    """Reference CPU backend on a bounded sample sized to ~10-20 s of CPU work."""
    try:
        import ref_icicle
        r = ref_icicle.get("bn254")
    except Exception as e:  # noqa
        return {"value": None, "unit": "points/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}
    logn = 16
    n = 1 << logn
    s, P = r.generate_scalars(n), r.generate_affine_points(n)
    r.msm(s, P, n)
    t0 = time.perf_counter(); r.msm(s, P, n); t16 = time.perf_counter() - t0
    # grow while the predicted time stays inside the budget (CPU MSM cost is ~linear in n)
    target = logn
    while target < 22 and t16 * (1 << (target + 1 - 16)) * 2 < budget_s:
        target += 1
    n = 1 << target
    s, P = r.generate_scalars(n), r.generate_affine_points(n)
    t0 = time.perf_counter(); r.msm(s, P, n); t1 = time.perf_counter() - t0
    t0 = time.perf_counter(); r.msm(s, P, n); t2 = time.perf_counter() - t0
    t = min(t1, t2)
    return {"value": n / t, "unit": "points/s", "cores": os.cpu_count(), "kind": "reference",
            "sample": f"icicle CPU backend (oracle/_ref built from /root/reference sources, g++ -O3, Taskflow stand-in), BN254 G1 MSM 2^{target} points, best of 2 ({t:.2f} s)"}


def main():
    # exactly ONE line on stdout (the JSON): library chatter (e.g. "NCCL version ..." under torchrun) goes to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        _main(real_stdout)
    finally:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)


def emit(fd, line):
    os.write(fd, (json.dumps(line) + "\n").encode())


def _main(out_fd):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--logn", type=int, default=26)
    ap.add_argument("--ntt-logn", type=int, default=24)
    ap.add_argument("--c", type=int, default=0)
    ap.add_argument("--ref-sample-logn", type=int, default=20)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ntt", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank, out_fd)
        return

    import torch
    import torch.distributed as dist
    import icicle_b200 as ib  # raises ImportError if the native library is missing: no fallback

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (the product has no CPU path)")
    torch.cuda.set_device(local_rank)
    ib.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    n = 1 << args.logn
    CURVE = ib.Curve.BN254_G1
    scalars, points = synth_inputs(torch, ib, args.logn, seed=rank, device=dev)
    result = ib.device_empty(24, dev).view(1, 24)
    gathered = ib.device_empty(24 * world, dev).view(world, 24) if world > 1 else None
    total_out = ib.device_empty(24, dev).view(1, 24)
    c_used = args.c or ib.msm_choose_c(CURVE, n)

    def step_device():
        ib.msm(CURVE, scalars, points, n, ib.MSMConfig(c=args.c, is_async=True), result)
        if world > 1:
            # the single exchange step of a point-sharded MSM: all-gather 96 B per rank over NVLink, then one EC-sum kernel
            dist.all_gather_into_tensor(gathered.view(-1), result.view(-1))
            ib.ec_sum(CURVE, gathered, world, ib.VecOpsConfig(is_async=True), total_out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = ib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device()
    e1.record()
    barrier()
    launches = ib.launch_count() - launches0
    ms_total = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    value = world * n / (ms_step * 1e-3)

    # ---- roofline of the dominant kernel: CUDA events around k_accumulate inside the call ----------------------------------
    ib.set_profiling(True)
    acc_ms, stage_sum = [], {}
    for _ in range(3):
        ib.msm(CURVE, scalars, points, n, ib.MSMConfig(c=args.c, is_async=True), result)
        _, stages = ib.last_profile()
        for nm, ms in stages:
            stage_sum[nm] = stage_sum.get(nm, 0.0) + ms / 3
        d = dict(stages)
        acc_ms.append(d.get("accumulate", float("nan")) + d.get("pair_levels", 0.0))
    ib.set_profiling(False)
    acc = sum(acc_ms) / len(acc_ms)
    peak, peak_kind = hbm_peak()
    achieved = ALG_BYTES_PER_POINT * n / (acc * 1e-3) / 1e9
    nwin = (254 + 1 + c_used - 1) // c_used
    levels = ib.msm_pair_levels(CURVE, n, args.c)
    # products per bucket entry: a pair level turns 2 entries into 1 with ~6 Montgomery products (batched-affine add); what is
    # left after L levels goes through the 10-product mixed XYZZ add
    prod_per_entry = sum(6.0 / (2 << l) for l in range(levels)) + 10.0 / (1 << levels)
    roofline = {"bound": "hbm", "kernel": "bucket accumulation: k_pair_prefix + k_pair_apply x %d levels + k_accumulate <Fp<bn254_fq>>" % levels,
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "pair_levels": levels,
                "traffic": NCU_TRAFFIC_BYTES.get((args.logn, c_used)), "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})", "kernel_ms": acc,
                "stage_ms": {k: round(v, 3) for k, v in stage_sum.items()},
                "imad_frac": (n * nwin * prod_per_entry * 140 / (acc * 1e-3)) / IMAD_WIDE_PEAK,
                "note": "integer-multiply bound: imad_frac = (N * windows * products/entry * 140 IMAD.WIDE) / stage time vs the measured 9.26e12 IMAD.WIDE/s"}

    # ---- e2e: host (pinned) buffers through the C ABI; H2D + D2H inside the timed region -----------------------------------
    e2e = None
    if not args.no_e2e:
        h_s, p1 = pinned_array(ib, (n, 8))
        h_p, p2 = pinned_array(ib, (n, 16))
        ib.capi.check(ib.capi.lib.b200_copy_to_host(h_s.ctypes.data, scalars.data_ptr(), h_s.nbytes, None, 0), "d2h")
        ib.capi.check(ib.capi.lib.b200_copy_to_host(h_p.ctypes.data, points.data_ptr(), h_p.nbytes, None, 0), "d2h")
        h_res = np.zeros((1, 24), dtype=np.uint32)

        def step_host():
            ib.msm(CURVE, h_s, h_p, n, ib.MSMConfig(c=args.c), h_res)  # host in, host out: blocking call
            if world > 1:
                r_dev = ib.to_device(h_res, dev)
                dist.all_gather_into_tensor(gathered.view(-1), r_dev.view(-1))
                ib.ec_sum(CURVE, gathered, world, ib.VecOpsConfig(), total_out)

        step_host()
        barrier()
        k2 = max(2, min(args.steps, 3))
        t0 = time.perf_counter()
        e0.record()
        for _ in range(k2):
            step_host()
        e1.record()
        barrier()
        ms2 = e0.elapsed_time(e1)
        t2 = torch.tensor([ms2], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t2, op=dist.ReduceOp.MAX)
        ms2 = float(t2.item()) / k2
        e2e = {"value": world * n / (ms2 * 1e-3), "unit": "points/s", "h2d_bytes_per_step": ALG_BYTES_PER_POINT * n, "d2h_bytes_per_step": 96,
               "ms_per_step": ms2, "steps": k2}
        ib.capi.lib.b200_host_free_pinned(p1)
        ib.capi.lib.b200_host_free_pinned(p2)
        del h_s, h_p

    # ---- secondary metric: BN254 NTT elements/s @ 2^ntt_logn (N = 1 only) ----------------------------------------------------
    ntt = None
    if not args.no_ntt and world == 1:
        del points
        torch.cuda.empty_cache()
        from icicle_b200 import utils
        fp = utils.field_params("bn254_fr")
        F = ib.Field.BN254_FR
        nl = args.ntt_logn
        ib.ntt_init_domain(F, utils.to_limbs([pow(fp["rou"], 1 << (fp["two_adicity"] - nl), fp["p"])], 8)[0])
        x = scalars[: 1 << nl].contiguous() if nl <= args.logn else synth_inputs(torch, ib, nl, 7, dev)[0]
        y = ib.device_empty((1 << nl) * 8, dev)
        out = {}
        for nm, d in (("forward", ib.NTTDir.kForward), ("inverse", ib.NTTDir.kInverse)):
            for _ in range(3):
                ib.ntt(F, x, 1 << nl, d, ib.NTTConfig(is_async=True), y)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(args.steps):
                ib.ntt(F, x, 1 << nl, d, ib.NTTConfig(is_async=True), y)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / args.steps
            out[nm] = {"elements_per_s": (1 << nl) / (ms * 1e-3), "ms": ms}
        a = ALG_BYTES_PER_NTT_ELEM * (1 << nl) / (out["forward"]["ms"] * 1e-3) / 1e9
        ntt = {"metric": "bn254_ntt_elements_per_s", "logn": nl, "ordering": "kNN", **out,
               "roofline": {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": NCU_NTT_TRAFFIC_BYTES.get(nl),
                            "kernel": "k_ntt_tile<Fp<bn254_fr>> x 3 passes (IMAD.WIDE bound: sm 62-69 %, dram 9-12 %, profiles/r1_ncu_ntt_passes.txt)"}}
        ib.ntt_release_domain(F)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()

    if rank == 0:
        line = {
            "metric": "bn254_g1_msm_points_per_s", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (254-bit modular integers, Montgomery arithmetic in IMAD.WIDE chains)",
            "data": "synthetic: uniform scalars < p; 2^16 distinct BN254 G1 points tiled to N; device-resident for `value`, pinned host for `e2e`",
            "config": {"workload": f"BN254 G1 MSM 2^{args.logn} per GPU (BASELINE configs[1]), precompute_factor 1, window c={c_used}",
                       "l2": "inputs (6 GiB at 2^26) exceed the 126 MB L2, no flush needed", "multi_gpu": "point-sharded; one NCCL all-gather of 96 B partials + ec_sum kernel"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "secondary": ntt,
        }
        emit(out_fd, line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
