"""Developer probe (GPU box): host-pointer (pinned) BN254 G1 MSM through the C ABI for several pipeline chunk counts
(B200_MSM_PIPELINE_CHUNKS), wall-clock per blocking call + the per-stage CUDA-event profile of one call."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import icicle_b200 as ib
import bench

logn = int(sys.argv[1]) if len(sys.argv) > 1 else 26
chunk_list = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0, 4, 8]
n = 1 << logn
dev = torch.device("cuda", 0)
ib.set_device(0)
C_ = ib.Curve.BN254_G1
scalars, points = bench.synth_inputs(torch, ib, logn, 0, dev)
h_s, p1 = bench.pinned_array(ib, (n, 8))
h_p, p2 = bench.pinned_array(ib, (n, 16))
ib.capi.check(ib.capi.lib.b200_copy_to_host(h_s.ctypes.data, scalars.data_ptr(), h_s.nbytes, None, 0), "d2h")
ib.capi.check(ib.capi.lib.b200_copy_to_host(h_p.ctypes.data, points.data_ptr(), h_p.nbytes, None, 0), "d2h")
res_dev = ib.device_empty(24).view(1, 24)
ib.msm(C_, scalars, points, n, ib.MSMConfig(is_async=True), res_dev)
torch.cuda.synchronize()
ref = res_dev.cpu().numpy()
del scalars, points
torch.cuda.empty_cache()
h_res = np.zeros((1, 24), dtype=np.uint32)
for ch in chunk_list:
    if ch > 0:
        ib.set_tuning("msm_pipeline_chunks", ch)      # k equal chunks
    else:
        ib.set_tuning("msm_pipeline_chunks", None)      # 0 = the default graded schedule
    ib.msm(C_, h_s, h_p, n, ib.MSMConfig(), h_res)
    ts = []
    for _ in range(3):
        t0 = time.perf_counter(); ib.msm(C_, h_s, h_p, n, ib.MSMConfig(), h_res); ts.append(time.perf_counter() - t0)
    same = bool((h_res == ref).all())
    ib.set_profiling(True)
    ib.msm(C_, h_s, h_p, n, ib.MSMConfig(), h_res)
    what, st = ib.last_profile()
    ib.set_profiling(False)
    agg = {}
    for k, v in st:
        agg[k] = agg.get(k, 0.0) + v
    print(f"e2e 2^{logn} chunks={ch}: best {min(ts)*1e3:.2f} ms  med {sorted(ts)[1]*1e3:.2f} ms  same_repr_as_device_path={same} | {what} " +
          " ".join(f"{k}={v:.2f}" for k, v in agg.items()), flush=True)
