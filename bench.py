#!/usr/bin/env python3
"""bench.py -- BN254 G1 MSM points/s @ 2^26 (headline) and BN254 NTT elements/s @ 2^24 (secondary), per BASELINE.json.

    python bench.py --gpus N --steps K --warmup W [--impl b200|reference] [--logn 26] [--ntt-logn 24] [--no-configs]

A step = one pass of the hot path over one batch of synthetic input: one MSM over 2^logn (scalar, point) pairs per GPU.
  value        whole-job points/s with inputs resident in HBM (device-timed, CUDA events on the launching stream, max over ranks)
  e2e          the same metric through the reference-facing PLUGIN call with HOST buffers, H2D of scalars+points and D2H of
               the result inside the timed region.  e2e.value is what an unmodified ICICLE caller sees: the reference
               frontend's own `bn254_msm` (oracle/_ref, Device{"CUDA"}) -> dispatcher -> registration shim -> C ABI, with
               PAGEABLE host vectors (numpy arrays; what Rust/Go/C++ callers pass).  e2e.variants also lists the C-ABI call
               with pageable and with pinned buffers.
  roofline     dominant stage (bucket accumulation): algorithmic bytes (96 B/point) / its CUDA-event duration vs the
               measured HBM peak (MEASURED_PEAKS.json); the stage is integer-multiply bound, see imad_frac
  cpu_baseline the UNMODIFIED reference CPU backend (oracle/_ref: built from /root/reference sources with g++ -O3 and a Taskflow
               STAND-IN thread pool -- upstream prefers clang + real Taskflow) on a bounded sample; the SAME bytes go through the
               GPU and the results are compared -> parity_checked
  configs      BASELINE configs 4 and 5 (BLS12-381 G1+G2 MSM 2^24 batch 8 with shared bases; BabyBear NTT 2^27 batch 128) sharded
               by batch index over the N GPUs, and (N > 1) strong scaling of one 2^26 MSM split by point range
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
# dram__bytes_read.sum + dram__bytes_write.sum of the bucket-accumulation stage of ONE MSM at the headline config (2^26 points,
# c = 20, 5 pair levels): all k_pair_prefix / k_pair_apply / k_inv_* / scan launches + k_accumulate, from the ncu launch list
# profiles/r2_ncu_launches_msm_2p26.txt.  It is ~70x the algorithmic 6.44 GB: level 0 gathers every point once per window in each
# of its two passes -- and since round 2 no longer lowers the device-wide L2 fetch granularity behind the caller's back, every
# random 32-byte gather pulls a 128-byte line (round 1 with the limit at 32 B: 307.8 GB, same run time) -- and every level writes
# its halved list (planar scratch); the stage is bound by the random-sector rate of HBM at level 0 and by IMAD.WIDE issue above it.
NCU_TRAFFIC_BYTES = {(26, 20): 446.8e9}  # round 2 (profiles/r2_ncu_launches_msm_2p26.txt); 307.8e9 with the opt-in l2_fetch_granularity = 32
# dram bytes of ONE forward BN254 NTT of 2^24 (3 k_ntt_tile passes), from profiles/; algorithmic = 1.07 GB
NCU_NTT_TRAFFIC_BYTES = {24: 4.02e9}  # 3 passes: 1.46 + 0.54 + 0.54 GB read, 3 x 0.49 GB written (profiles/r2_ncu_launches_ntt_bn254_2p24.txt)
CPU_NOTE = "icicle CPU backend (oracle/_ref built from /root/reference sources with g++ -O3; Taskflow STAND-IN thread pool, not upstream's clang + Taskflow 3.8; auto window size; `cpu_quota` = CPUs the container's cgroup grants, the reference starts `cores` threads regardless)"


def cpu_quota():
    """CPUs the container really grants (cgroup v2 cpu.max); the GPU boxes of this pool show 128 cores under a 16-CPU quota."""
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        return None if q == "max" else round(int(q) / int(per), 2)
    except Exception:
        return None


def bind_to_gpu_numa_node(torch, local_rank):
    """CPU affinity per rank (round-1 verdict item 8): run this process -- and therefore the host vectors it first-touches and the
    library's copier threads, which inherit the mask -- on the cores of the NUMA node its GPU hangs off
    (/sys/bus/pci/devices/<bus id>/local_cpulist).  torchrun leaves ranks floating over both sockets; a rank whose pageable
    vectors live on the other socket reads them over UPI and the e2e number collapses (profiles/r2_e2e_numa.txt)."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        txt = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return txt
    except Exception:
        pass
    return None


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


def rand_scalars_dev(torch, n, seed, device, top=0x30644E72, limbs=8):
    """n uniform scalars as (n, limbs) int32 on the device; top limb below the modulus' top limb -> always < p."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    s = torch.randint(-2 ** 31, 2 ** 31, (n, limbs), dtype=torch.int64, device=device, generator=g).to(torch.int32)
    s[:, limbs - 1] = torch.randint(0, top, (n,), dtype=torch.int64, device=device, generator=g).to(torch.int32)
    return s.contiguous()


def synth_inputs(torch, ib, logn, seed, device):
    """2^logn uniform scalars and 2^logn points = 2^16 DISTINCT curve points tiled (upstream's generator
    repeats only 100 points)."""
    import common
    n = 1 << logn
    distinct = min(n, 1 << 16)
    base = common.gen_g1_points("bn254", distinct, 1000 + seed)
    pts = ib.to_device(base, device).repeat(n // distinct, 1).contiguous()
    return rand_scalars_dev(torch, n, seed, device), pts


def load_reference(name="bn254"):
    try:
        import ref_icicle
        if not ref_icicle.available(name):
            return None
        return ref_icicle.get(name)
    except Exception:
        return None


def load_frontend_cuda(ref, local_rank):
    """Make the reference frontend dispatch to our backend DSOs (build/backend/bn254) -- the drop-in path of tests/test_gpu_dropin.py."""
    bdir = os.path.join(ROOT, "build", "backend", "bn254")
    if ref is None or not os.path.exists(os.path.join(bdir, "libicicle_backend_cuda_device.so")):
        return False
    try:
        if "CUDA" not in ref.registered_devices():
            ref.load_backend(bdir)
        if "CUDA" not in ref.registered_devices():
            return False
        ref.set_device("CUDA", local_rank)
        ref.set_device("CPU", 0)
        return True
    except Exception:
        return False


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
        "cpu_baseline": {"value": val, "unit": "points/s", "cores": cores, "cpu_quota": cpu_quota(), "kind": "reference",
                         "sample": f"{CPU_NOTE}; MSM of 2^{sample_log} points, mean of {args.steps}; per-step times {[round(x, 3) for x in ts]} s"},
        "e2e": {"value": val, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(out_fd, line)


def cpu_baseline_and_parity(ib, ref, budget_s=20.0):
    """Reference CPU backend on a bounded sample sized to ~10-20 s of CPU work; the SAME bytes then go through the GPU (host
    pointers -> the plugin-facing path) and the two results are compared as group elements."""
    import common
    if ref is None:
        return {"value": None, "unit": "points/s", "cores": os.cpu_count(), "kind": "reference", "sample": "unavailable: oracle/_ref/bn254 not present"}, None
    ref.set_device("CPU", 0)
    n16 = 1 << 16
    s, P = common.seeded_scalars("bn254_fr", n16, 16), common.tiled_g1_points("bn254", n16, 1 << 14, 17)
    ref.msm(s, P, n16)
    t0 = time.perf_counter(); ref.msm(s, P, n16); t16 = time.perf_counter() - t0
    target = 16
    while target < 22 and t16 * (1 << (target + 1 - 16)) * 2 < budget_s:  # CPU MSM cost is ~linear in n
        target += 1
    n = 1 << target
    s, P = common.seeded_scalars("bn254_fr", n, 1600 + target), common.tiled_g1_points("bn254", n, 1 << 14, 1700 + target)
    t0 = time.perf_counter(); exp = ref.msm(s, P, n); t1 = time.perf_counter() - t0
    t0 = time.perf_counter(); ref.msm(s, P, n); t2 = time.perf_counter() - t0
    t = min(t1, t2)
    got = ib.msm(ib.Curve.BN254_G1, s, P, n)
    ok = bool(ref.projective_eq(got[0], exp[0]))
    cpu = {"value": n / t, "unit": "points/s", "cores": os.cpu_count(), "cpu_quota": cpu_quota(), "kind": "reference",
           "sample": f"{CPU_NOTE}; BN254 G1 MSM 2^{target} points (seeded scalars, 2^14 distinct points tiled), best of 2 ({t1:.2f} s, {t2:.2f} s)"}
    return cpu, {"what": "GPU (host-pointer path) == reference CPU backend on identical bytes, projective_eq", "logn": target, "ok": ok}


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


def timed(torch, dist, world, dev, fn, steps, warmup=1):
    """barrier + sync, `steps` calls of fn bracketed by CUDA events on the current stream, barrier + sync; max over ranks -> ms/step"""
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / steps


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
    ap.add_argument("--no-configs", action="store_true", help="skip BASELINE configs 4/5 and the strong-scaling line")
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
    import common

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (the product has no CPU path)")
    numa_cpus = None if os.environ.get("B200_BENCH_NO_NUMA_BIND") else bind_to_gpu_numa_node(torch, local_rank)
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

    def combine(res):
        # the single exchange step of a point-sharded MSM: all-gather 96 B per rank over NVLink, then one EC-sum kernel
        dist.all_gather_into_tensor(gathered.view(-1), res.view(-1))
        ib.ec_sum(CURVE, gathered, world, ib.VecOpsConfig(is_async=True), total_out)

    def step_device():
        ib.msm(CURVE, scalars, points, n, ib.MSMConfig(c=args.c, is_async=True), result)
        if world > 1:
            combine(result)

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

    # ---- roofline of the dominant stage: CUDA events around the bucket accumulation inside the call ----------------------------
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

    ref = load_reference("bn254") if rank == 0 or not args.no_e2e else None
    have_frontend = (not args.no_e2e) and load_frontend_cuda(ref, local_rank)

    # ---- e2e: host buffers through the plugin call; H2D + D2H inside the timed region --------------------------------------------
    e2e = None
    if not args.no_e2e:
        h_s = ib.to_host(scalars)                      # pageable numpy arrays: what an unmodified caller holds
        h_p = ib.to_host(points)
        h_res = np.zeros((1, 24), dtype=np.uint32)
        k2 = max(2, min(args.steps, 3))
        variants = {}

        def finish_multi():
            if world > 1:
                combine(ib.to_device(h_res, dev))
                torch.cuda.synchronize()

        def run_variant(name, call):
            def one():
                call()
                finish_multi()
            ms = timed(torch, dist, world, dev, one, k2, warmup=1)
            variants[name] = {"value": world * n / (ms * 1e-3), "ms_per_step": ms}

        p_s, p1 = pinned_array(ib, (n, 8))
        p_p, p2 = pinned_array(ib, (n, 16))
        p_s[:] = h_s
        p_p[:] = h_p

        def frontend(s_arr, p_arr):
            ref.set_device("CUDA", local_rank)
            h_res[:] = ref.msm(s_arr, p_arr, n, c=args.c)
        if have_frontend:
            run_variant("frontend_pinned", lambda: frontend(p_s, p_p))
            run_variant("frontend_pageable", lambda: frontend(h_s, h_p))
            ref.set_device("CPU", 0)
        run_variant("cabi_pinned", lambda: ib.msm(CURVE, p_s, p_p, n, ib.MSMConfig(c=args.c), h_res))    # host in, host out: blocking call
        run_variant("cabi_pageable", lambda: ib.msm(CURVE, h_s, h_p, n, ib.MSMConfig(c=args.c), h_res))
        ib.capi.lib.b200_host_free_pinned(p1)
        ib.capi.lib.b200_host_free_pinned(p2)
        head = "frontend_pinned" if "frontend_pinned" in variants else "cabi_pinned"
        e2e = {"value": variants[head]["value"], "unit": "points/s", "h2d_bytes_per_step": ALG_BYTES_PER_POINT * n, "d2h_bytes_per_step": 96,
               "ms_per_step": variants[head]["ms_per_step"], "steps": k2, "path": head,
               "path_note": ("the PLUGIN call: unmodified reference frontend bn254_msm with Device{CUDA} -> dispatcher -> registration shim -> C ABI, host vectors in "
                             "PINNED memory (the bench contract); `variants` lists the same call with PAGEABLE vectors (what Rust/Go/C++ callers usually hold: moved "
                             "through the backend's pinned copier ring) and the bare C-ABI calls.  Pageable vectors are bound by host memory traffic shared by all "
                             "ranks at N > 1 (profiles/r2_e2e_numa.txt)"
                             if head == "frontend_pinned" else "C ABI b200_msm with pinned host vectors (reference frontend build not present)"),
               "variants": variants}
        del h_s, h_p, p_s, p_p

    # ---- the other BASELINE configs, sharded by batch index / point range over the ranks ----------------------------------------
    configs = None
    if not args.no_configs:
        del points
        torch.cuda.empty_cache()
        ib.trim_scratch(0)
        configs = run_configs(torch, dist, ib, common, args, rank, world, dev, peak, scalars, combine if world > 1 else None)

    # ---- secondary metric: BN254 NTT elements/s @ 2^ntt_logn (N = 1 only) ----------------------------------------------------
    ntt = None
    if not args.no_ntt and world == 1:
        points = None
        torch.cuda.empty_cache()
        ntt = run_ntt_secondary(torch, ib, common, args, dev, peak, scalars, ref if rank == 0 else None, have_frontend)

    cpu, parity = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu, parity = cpu_baseline_and_parity(ib, ref)

    if rank == 0:
        line = {
            "metric": "bn254_g1_msm_points_per_s", "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32 limbs (254-bit modular integers, Montgomery arithmetic in IMAD.WIDE chains)",
            "data": "synthetic: uniform scalars < p; 2^16 distinct BN254 G1 points tiled to N; device-resident for `value`, host (pageable / pinned) for `e2e`",
            "config": {"workload": f"BN254 G1 MSM 2^{args.logn} per GPU (BASELINE configs[1]), precompute_factor 1, window c={c_used}",
                       "l2": "inputs (6 GiB at 2^26) exceed the 126 MB L2, no flush needed", "multi_gpu": "point-sharded; one NCCL all-gather of 96 B partials + ec_sum kernel",
                       "cpu_affinity": f"each rank bound to its GPU's NUMA node (cpus {numa_cpus})" if numa_cpus else "not bound"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "e2e": e2e, "cpu_baseline": cpu, "parity_checked": parity,
            "secondary": ntt, "configs": configs,
        }
        emit(out_fd, line)
    if world > 1:
        dist.destroy_process_group()


def run_ntt_secondary(torch, ib, common, args, dev, peak, scalars, ref, have_frontend):
    from icicle_b200 import utils
    fp = utils.field_params("bn254_fr")
    F = ib.Field.BN254_FR
    nl = args.ntt_logn
    nn = 1 << nl
    root = utils.to_limbs([pow(fp["rou"], 1 << (fp["two_adicity"] - nl), fp["p"])], 8)[0]
    ib.ntt_release_domain(F)
    ib.ntt_init_domain(F, root)
    x = scalars[:nn].contiguous() if nl <= args.logn else rand_scalars_dev(torch, nn, 7, dev)
    y = ib.device_empty(nn * 8, dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    for nm, d in (("forward", ib.NTTDir.kForward), ("inverse", ib.NTTDir.kInverse)):
        for _ in range(3):
            ib.ntt(F, x, nn, d, ib.NTTConfig(is_async=True), y)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(args.steps):
            ib.ntt(F, x, nn, d, ib.NTTConfig(is_async=True), y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        out[nm] = {"elements_per_s": nn / (ms * 1e-3), "ms": ms}
    a = ALG_BYTES_PER_NTT_ELEM * nn / (out["forward"]["ms"] * 1e-3) / 1e9
    ntt = {"metric": "bn254_ntt_elements_per_s", "logn": nl, "ordering": "kNN", **out,
           "roofline": {"bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": a / peak, "traffic": NCU_NTT_TRAFFIC_BYTES.get(nl),
                        "kernel": "k_ntt_tile<Fp<bn254_fr>, 2, 8> x 3 passes (radix-4 rounds, 3 CTAs/SM; IMAD.WIDE bound)"}}
    # e2e: host vectors in and out through the plugin call (H2D + D2H of 32 B/element each way inside the timed region)
    hx = ib.to_host(x)
    hy = np.zeros_like(hx)   # caller-owned output vector, already touched (first-touch page faults are not the backend's)
    variants = {}

    def t_host(fn, k=3):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / k * 1e3
    ms = t_host(lambda: ib.ntt(F, hx, nn, ib.NTTDir.kForward, None, hy))
    variants["cabi_pageable"] = {"value": nn / (ms * 1e-3), "ms_per_step": ms}
    if have_frontend and ref is not None:
        ref.set_device("CUDA", dev.index or 0)
        ref.ntt_release_domain()
        ref.ntt_init_domain(root)
        ms = t_host(lambda: ref.ntt(hx, nn, 0, out=hy))
        variants["frontend_pageable"] = {"value": nn / (ms * 1e-3), "ms_per_step": ms}
        ref.ntt_release_domain()           # releases the CUDA device's (= our) domain: restore it for the rest of the run
        ref.set_device("CPU", 0)
        ib.ntt_init_domain(F, root)
    head = "frontend_pageable" if "frontend_pageable" in variants else "cabi_pageable"
    ntt["e2e"] = {"value": variants[head]["value"], "unit": "elements/s", "h2d_bytes_per_step": 32 * nn, "d2h_bytes_per_step": 32 * nn,
                  "ms_per_step": variants[head]["ms_per_step"], "path": head, "variants": variants}
    # CPU baseline on a bounded sample + parity on identical bytes
    if ref is not None and not args.no_cpu_baseline:
        sl = min(nl, 22)
        sn = 1 << sl
        ref.set_device("CPU", 0)
        ref.ntt_release_domain()
        ref.ntt_init_domain(root)
        hs = hx[:sn]
        ref.ntt(hs, sn, 0)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); exp = ref.ntt(hs, sn, 0); ts.append(time.perf_counter() - t0)
        got = ib.ntt(F, hs, sn, ib.NTTDir.kForward)
        back = ib.ntt(F, exp, sn, ib.NTTDir.kInverse)
        ref.ntt_release_domain()
        ntt["cpu_baseline"] = {"value": sn / min(ts), "unit": "elements/s", "cores": os.cpu_count(), "cpu_quota": cpu_quota(), "kind": "reference",
                               "sample": f"{CPU_NOTE}; BN254 forward NTT 2^{sl}, best of 3 ({min(ts):.3f} s)"}
        ntt["parity_checked"] = {"what": "GPU forward NTT == reference CPU backend (memcmp) and GPU inverse of the reference output == input", "logn": sl,
                                 "ok": bool(np.array_equal(got, exp) and np.array_equal(back, hs))}
    ib.ntt_release_domain(F)
    return ntt


def run_configs(torch, dist, ib, common, args, rank, world, dev, peak, scalars, combine):
    """BASELINE configs 4 and 5 sharded over the ranks (weak in nothing: the TOTAL work is fixed = "strong"), plus strong scaling
    of one 2^26 BN254 MSM by point range.  Every line: aggregate rate, ms per pass (max over ranks), HBM roofline fraction of the
    algorithmic bytes, and a parity flag for the check that ran in this process."""
    from icicle_b200 import utils
    out = []
    GOLD = os.path.join(ROOT, "tests", "golden")
    steps = 2

    def alltrue(flag):
        t = torch.tensor([1 if flag else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    # ---- config 4: BLS12-381 G1 and G2 MSM 2^24, batch 8, shared bases: batch index partitioned, no exchange -------------------
    try:
        g = np.load(os.path.join(GOLD, "bls12_381.npz"))
        logn, batch = 24, 8
        n = 1 << logn
        lo, hi = ib.shard_range(batch, world, rank)
        mine = hi - lo
        for nm, curve, src, alg in (("G1", ib.Curve.BLS12_381_G1, g["msm_points"], 32 * batch + 96), ("G2", ib.Curve.BLS12_381_G2, g["g2_points"], 32 * batch + 192)):
            if mine == 0:
                continue
            base = np.ascontiguousarray(src[[i for i in range(src.shape[0]) if src[i].any()]])  # the reference's own points (no zero point)
            reps = (1 << 12) // base.shape[0] + 1
            tile = ib.to_device(np.tile(base, (reps, 1))[: 1 << 12], dev)
            P = tile.repeat(n >> 12, 1).contiguous()
            s = rand_scalars_dev(torch, n * mine, 400 + rank, dev, top=0x73EDA753)
            res = ib.device_empty(mine * ib.projective_limbs(curve), dev).view(mine, -1)
            cfg = lambda c=0: ib.MSMConfig(batch_size=mine, are_points_shared_in_batch=True, is_async=True, c=c)
            ms = timed(torch, dist, world, dev, lambda: ib.msm(curve, s, P, n, cfg(), res), steps, warmup=1)
            # parity that fits a bench: the first MSM of this rank's share recomputed alone with another window size must be the same group element
            alt = ib.msm(curve, s[:n], P, n, ib.MSMConfig(c=14))
            same = _same_point(ib, utils, curve, ib.to_host(res[0:1])[0], alt[0]) if hasattr(alt, "shape") else False
            rate = batch * n / (ms * 1e-3)
            gbs = alg * n / (ms * 1e-3) / 1e9 / world
            out.append({"config": f"BLS12-381 {nm} MSM 2^{logn} x batch {batch}, shared bases (BASELINE configs[3]), {mine} MSMs per GPU", "n_gpus": world,
                        "value": rate, "unit": "points/s", "ms_per_pass": ms, "scaling": "strong",
                        "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                                     "note": f"algorithmic bytes = (batch*32 + {alg - 32 * batch}) B per point index, per GPU; integer-multiply bound"},
                        "parity": {"what": "batched result[0] == the same MSM alone with c=14 (group elements); reference parity at 2^20 in tests/test_gpu_fullsize.py", "ok": alltrue(same)}})
            del P, s, res
            torch.cuda.empty_cache()
            ib.trim_scratch(0)
    except Exception as e:  # noqa
        out.append({"config": "BLS12-381 G1+G2 MSM 2^24 x batch 8", "error": repr(e)})

    # ---- config 5: BabyBear NTT 2^27, batch 128 (2-adicity 27: the survey's substitute for "2^28 x 64"), rows partitioned -------
    try:
        F = ib.Field.BABYBEAR
        fp = utils.field_params("babybear")
        logn, batch = 27, 128
        n = 1 << logn
        lo, hi = ib.shard_range(batch, world, rank)
        mine = hi - lo
        ib.ntt_release_domain(F)
        ib.ntt_init_domain(F, utils.to_limbs([pow(fp["rou"], 1 << (fp["two_adicity"] - logn), fp["p"])], 1)[0])
        chunk = min(mine, 16)   # rows per call: 16 x 512 MiB in + out
        x = torch.randint(0, fp["p"], (chunk * n,), dtype=torch.int64, device=dev).to(torch.int32).contiguous()
        y = ib.device_empty(chunk * n, dev)

        def one_pass():
            for _ in range(0, mine, chunk):
                ib.ntt(F, x, n, ib.NTTDir.kForward, ib.NTTConfig(batch_size=chunk, is_async=True), y)
        ms = timed(torch, dist, world, dev, one_pass, steps, warmup=1)
        back = ib.device_empty(chunk * n, dev)
        ib.ntt(F, y, n, ib.NTTDir.kInverse, ib.NTTConfig(batch_size=chunk), back)
        ok = bool(torch.equal(back.view(-1), x.view(-1)))
        rate = batch * n / (ms * 1e-3)
        gbs = 8.0 * mine * n / (ms * 1e-3) / 1e9
        out.append({"config": f"BabyBear NTT 2^{logn} x batch {batch} forward kNN (BASELINE configs[4] with the survey's size substitution), {mine} rows per GPU in calls of {chunk}",
                    "n_gpus": world, "value": rate, "unit": "elements/s", "ms_per_pass": ms, "scaling": "strong",
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak, "note": "algorithmic bytes = 8 B per element per transform, per GPU"},
                    "parity": {"what": "inverse(forward(x)) == x bit-exactly on this rank's rows; reference parity at 2^24 x 2 and 2^27 in tests/test_gpu_fullsize.py", "ok": alltrue(ok)}})
        del x, y, back
        ib.ntt_release_domain(F)
        torch.cuda.empty_cache()
        ib.trim_scratch(0)
    except Exception as e:  # noqa
        out.append({"config": "BabyBear NTT 2^27 x batch 128", "error": repr(e)})

    # ---- ONE BN254 NTT spanning the GPUs: phase 1 -> NCCL all-to-all over NVLink -> phase 2 (csrc/ntt.cu b200_ntt_dist_phase1/2) ----
    if world > 1 and (world & (world - 1)) == 0:
        try:
            F = ib.Field.BN254_FR
            fp = utils.field_params("bn254_fr")
            n_log = 26
            a_log = (n_log + 1) // 2
            b_log = n_log - a_log
            A, B = 1 << a_log, 1 << b_log
            ib.ntt_release_domain(F)
            ib.ntt_init_domain(F, utils.to_limbs([pow(fp["rou"], 1 << (fp["two_adicity"] - n_log), fp["p"])], 8)[0])
            slab0 = rand_scalars_dev(torch, (A * B) // world, 900 + rank, dev)          # this rank's column slab [A][B/world]
            slab = slab0.clone()
            recv = torch.empty_like(slab)
            outb = torch.empty_like(slab)
            a2a_ms = []

            def one(direction=0, al=a_log, bl=b_log, src=None):
                buf = slab if src is None else src
                ib.capi.check(ib.capi.lib.b200_ntt_dist_phase1(int(F), buf.data_ptr(), al, bl, world, rank, direction, None), "dist_phase1")
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                dist.all_to_all_single(recv.view(-1), buf.view(-1))
                e1.record()
                ib.capi.check(ib.capi.lib.b200_ntt_dist_phase2(int(F), recv.data_ptr(), outb.data_ptr(), al, bl, world, rank, direction, None), "dist_phase2")
                a2a_ms.append((e0, e1))

            def fwd():
                slab.copy_(slab0)
                one()
            ms = timed(torch, dist, world, dev, fwd, 3, warmup=2)
            ex = [a.elapsed_time(b) for a, b in a2a_ms[-3:]]
            a2a = sum(ex) / len(ex)
            # parity: the inverse transform (dimensions swapped) of the result must give the input back, bit for bit
            res = outb.clone()
            one(1, b_log, a_log, res)
            torch.cuda.synchronize()
            ok = bool(torch.equal(outb.view(-1), slab0.view(-1)))
            sent = (A * B // world) * 32 * (world - 1) / world                              # bytes this rank puts on NVLink per transform
            out.append({"config": f"ONE BN254 NTT of 2^{n_log} spanning {world} GPUs (column slabs; 4-step with a single NCCL all_to_all_single over NVLink)",
                        "n_gpus": world, "value": (A * B) / (ms * 1e-3), "unit": "elements/s", "ms_per_pass": ms, "scaling": "strong",
                        "all_to_all_ms": a2a, "nvlink_gbs_per_gpu_each_way": sent / (a2a * 1e-3) / 1e9,
                        "roofline": {"bound": "hbm", "achieved": 64.0 * (A * B // world) / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                     "frac": 64.0 * (A * B // world) / (ms * 1e-3) / 1e9 / peak},
                        "parity": {"what": "distributed inverse(distributed forward(x)) == x bit-exactly; distributed == single-device transform in tests/test_gpu_multi.py", "ok": alltrue(ok)}})
            del slab, slab0, recv, outb, res
            ib.ntt_release_domain(F)
            torch.cuda.empty_cache()
            ib.trim_scratch(0)
        except Exception as e:  # noqa
            out.append({"config": "ONE BN254 NTT spanning the GPUs", "error": repr(e)})

    # ---- strong scaling of ONE 2^26 BN254 MSM: the point range is partitioned, partials combined by all-gather + ec_sum ----------
    if world > 1:
        try:
            CURVE = ib.Curve.BN254_G1
            n = 1 << args.logn
            lo, hi = ib.shard_range(n, world, rank)
            m = hi - lo
            base = ib.to_device(common.gen_g1_points("bn254", 1 << 12, 4242), dev)
            P = base.repeat((m >> 12) + 1, 1)[:m].contiguous()
            s = scalars[:m].contiguous()
            res = ib.device_empty(24, dev).view(1, 24)

            def one():
                ib.msm(CURVE, s, P, m, ib.MSMConfig(is_async=True), res)
                combine(res)
            ms = timed(torch, dist, world, dev, one, 3, warmup=2)
            out.append({"config": f"BN254 G1 MSM 2^{args.logn} split by point range over {world} GPUs ({m} points per GPU) + NCCL all-gather of 96 B partials + ec_sum",
                        "n_gpus": world, "value": n / (ms * 1e-3), "unit": "points/s", "ms_per_pass": ms, "scaling": "strong",
                        "roofline": {"bound": "hbm", "achieved": 96.0 * m / (ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": 96.0 * m / (ms * 1e-3) / 1e9 / peak}})
            del P, s
        except Exception as e:  # noqa
            out.append({"config": "BN254 G1 MSM 2^26 strong scaling", "error": repr(e)})
    return out


def _same_point(ib, utils, curve, a, b):
    """projective equality by cross-multiplication on python integers (G1) / Fq2 pairs (G2)"""
    L = ib.projective_limbs(curve) // 3
    if curve in (ib.Curve.BLS12_381_G2,):
        Lq = L // 2
        q = utils.field_params("bls12_381_fq")["p"]
        nr = utils.curve_params("bls12_381")["nonresidue"]
        A = utils.from_limbs(np.asarray(a, dtype=np.uint32).reshape(6, Lq))
        B = utils.from_limbs(np.asarray(b, dtype=np.uint32).reshape(6, Lq))
        mul = lambda u, v: ((u[0] * v[0] + nr * u[1] * v[1]) % q, (u[0] * v[1] + u[1] * v[0]) % q)
        (ax, ay, az), (bx, by, bz) = ((A[0], A[1]), (A[2], A[3]), (A[4], A[5])), ((B[0], B[1]), (B[2], B[3]), (B[4], B[5]))
        return mul(ax, bz) == mul(bx, az) and mul(ay, bz) == mul(by, az)
    q = utils.field_params("bls12_381_fq")["p"]
    ax, ay, az = utils.from_limbs(np.asarray(a, dtype=np.uint32).reshape(3, L))
    bx, by, bz = utils.from_limbs(np.asarray(b, dtype=np.uint32).reshape(3, L))
    return (ax * bz - bx * az) % q == 0 and (ay * bz - by * az) % q == 0


if __name__ == "__main__":
    main()
