"""CPU suite: host logic -- the C ABI library loads and exports every symbol include/icicle_b200.h declares, argument
validation works without a GPU, the generated field constants agree with the reference headers, the device field
arithmetic (ff.cuh) is exercised through its host emulation, and the N>1 sharding logic runs under gloo (world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cabi_exports_every_declared_symbol():
    import icicle_b200 as ib
    hdr = open(os.path.join(ROOT, "include", "icicle_b200.h")).read()
    declared = set(re.findall(r"B200_API\s+[\w\s\*]+?\b(b200_\w+)\s*\(", hdr))
    assert len(declared) >= 40
    out = subprocess.run(["nm", "-D", "--defined-only", ib.capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (b200_\w+)", out))
    assert declared <= exported, declared - exported
    assert declared <= set(ib.capi.SYMBOLS), declared - set(ib.capi.SYMBOLS)


def test_argument_validation_without_gpu():
    import icicle_b200 as ib
    lib = ib.capi.lib
    cfg = ib.capi.MsmConfigC()
    lib.b200_msm_default_config(C.byref(cfg))
    assert (cfg.precompute_factor, cfg.batch_size, cfg.are_points_shared_in_batch, cfg.c, cfg.bitsize) == (1, 1, 1, 0, 0)  # msm.h:60-78
    assert lib.b200_msm(0, None, None, 10, C.byref(cfg), None) == 3  # INVALID_POINTER
    assert lib.b200_msm(99, 1, 1, 10, C.byref(cfg), 1) == 11  # INVALID_ARGUMENT (unknown curve)
    n = ib.capi.NttConfigC()
    lib.b200_ntt_default_config(C.byref(n))
    assert (n.batch_size, n.ordering, n.columns_batch) == (1, 0, 0)  # ntt.h:73-86
    assert lib.b200_ntt(0, None, 8, 0, C.byref(n), None) == 3
    assert ib.field_limbs(ib.Field.BLS12_381_FQ) == 12 and ib.affine_limbs(ib.Curve.BN254_G2) == 32 and ib.projective_limbs(ib.Curve.BW6_761_G1) == 72
    assert ib.scalar_field(ib.Curve.GRUMPKIN) == ib.Field.BN254_FQ
    # window heuristic is monotone in the problem size
    cs = [ib.msm_choose_c(ib.Curve.BN254_G1, 1 << k) for k in (10, 14, 18, 22, 26)]
    assert cs == sorted(cs) and 4 <= cs[0] and cs[-1] <= 24


def test_ff_host_emulation_matches_python_ints(tmp_path):
    """Compiles icicle_b200/csrc/ff.cuh for the host (carry flag emulated) and checks add/sub/mul/Montgomery conversion for
    every supported field -- the exact code the kernels run."""
    exe = str(tmp_path / "ff_emul")
    src = os.path.join(ROOT, "tests", "emul", "ff_emul_test.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-x", "c++", src, "-o", exe], check=True)
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    import check_ff
    n, m, bad = check_ff.run(exe, n_random=150, seed=3)
    assert n == m and not bad, bad[:2]


def test_generated_params_match_reference_headers():
    ref = os.environ.get("ICICLE_REF", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "icicle")):
        pytest.skip("reference tree not present (GPU box)")
    import json
    before = json.load(open(os.path.join(ROOT, "icicle_b200", "params.json")))
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import importlib
    gp = importlib.import_module("gen_params")
    for f in gp.FIELDS:
        assert int(before["fields"][f["name"]]["p"], 16) == f["p"]
        assert int(before["fields"][f["name"]]["R2"], 16) == pow(1 << (32 * f["limbs"]), 2, f["p"])
    for c in gp.CURVES:
        assert int(before["curves"][c["name"]]["gx"], 16) == c["gx"]


WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "oracle")); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch, torch.distributed as dist, numpy as np
import common, port
from icicle_b200 import utils
from icicle_b200 import shard_range
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, world = dist.get_rank(), dist.get_world_size()
n = 64
pts = common.gen_g1_points("bn254", n, 9, as_ints=True)
sc = common.rand_field_elems("bn254_fr", n, 10, as_ints=True)
lo, hi = shard_range(n, world, rank)
part = port.msm("bn254", sc[lo:hi], pts[lo:hi], c=6)       # this rank's partial result (CPU oracle stands in for the GPU)
buf = torch.tensor(np.array(utils.to_limbs([part[0], part[1]], 8)).astype(np.int64).reshape(-1))
gathered = [torch.zeros_like(buf) for _ in range(world)]
dist.all_gather(gathered, buf)                              # the single exchange step of the sharded MSM
q = utils.field_params("bn254_fq")["p"]
acc = None
for g in gathered:
    x, y = utils.from_limbs(g.numpy().astype(np.uint32).reshape(2, 8))
    acc = common.ec_add(acc, (x, y), q)
assert acc == common.msm_naive_ints(sc, pts, q), "sharded MSM != full MSM"
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_point_sharded_msm_two_ranks_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port_no = str(29500 + os.getpid() % 1000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port_no, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_pipeline_schedule(tuning):
    """Chunk schedule of the host-pointer MSM pipeline (msm_impl.cuh pipeline_schedule; pure host logic): covers every point
    exactly once, starts with a small chunk (exposed H2D time) and ends with quarters (long bucket runs), honours the
    equal-chunks override and the test override that allows tiny chunks."""
    import ctypes as C
    from icicle_b200 import capi
    def sched(n):
        buf = (C.c_uint32 * 32)()
        k = capi.lib.b200_msm_pipeline_schedule(n, buf, 32)
        return [buf[i] for i in range(k)]
    tuning("msm_pipeline_chunks", None)
    tuning("msm_pipeline_min", None)
    assert sched(1 << 26) == [1 << 22, 1 << 22, 1 << 23, 1 << 24, 1 << 24, 1 << 24]
    for n in ((1 << 23), (1 << 23) + 12345, (1 << 25) - 1, (1 << 27) + 7, (1 << 30)):
        s = sched(n)
        assert sum(s) == n and 1 <= len(s) <= 6 and all(c >= 1 << 20 for c in s), (n, s)
        assert s[0] <= max(n >> 4, 1 << 20)
    assert sched(1 << 20) == [1 << 20]
    tuning("msm_pipeline_chunks", 4)
    assert sched(1 << 26) == [1 << 24] * 4
    s = sched((1 << 24) + 3)
    assert sum(s) == (1 << 24) + 3 and len(s) == 4
    tuning("msm_pipeline_min", 2)
    tuning("msm_pipeline_chunks", 7)
    s = sched(4113)
    assert sum(s) == 4113 and len(s) == 7
    tuning("msm_pipeline_chunks", None)
    for n in (1, 2, 5, 100, 8203):
        s = sched(n)
        assert sum(s) == n and all(c > 0 for c in s), (n, s)
    assert capi.lib.b200_msm_pipeline_schedule(0, (C.c_uint32 * 4)(), 4) == -1


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's own CPU MSM on a bounded sample): exactly one JSON line on stdout with the
    contract's keys; needs no GPU."""
    import json
    ref_icicle = pytest.importorskip("ref_icicle")
    if not ref_icicle.available("bn254"):
        pytest.skip("oracle/_ref/bn254 not built")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--ref-sample-logn", "10"],
                         stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300, check=True).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "points/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0


def test_ec_host_emulation_matches_python_ints(tmp_path):
    """Compiles icicle_b200/csrc/ec.cuh for the host and checks the XYZZ group law the MSM / ECNTT kernels run (mixed add, full
    add on general representatives, doubling, chained adds, P + (-P), affine zero / infinity on either side) against
    Python-integer curve arithmetic on BN254, BLS12-381, Grumpkin and the two G2 groups over Fq2 (ext.cuh) -- the exact device code, carry flag emulated."""
    exe = str(tmp_path / "ec_emul")
    src = os.path.join(ROOT, "tests", "emul", "ec_emul_test.cpp")
    subprocess.run(["g++", "-std=c++17", "-O1", "-x", "c++", src, "-o", exe], check=True)
    sys.path.insert(0, os.path.join(ROOT, "tests", "emul"))
    import check_ec
    n, m, bad = check_ec.run(exe, n_random=40, seed=7)
    assert n == m and not bad, bad[:2]


def test_tuning_table_roundtrip():
    """b200_set_tuning / b200_get_tuning (developer knobs: read once from B200_* at load, never via getenv on the hot path): unknown
    names are rejected, negative values unset, values survive a round trip; pure host logic."""
    import icicle_b200 as ib
    from icicle_b200 import capi
    old = ib.get_tuning("msm_pair_levels")
    ib.set_tuning("msm_pair_levels", 3)
    assert ib.get_tuning("msm_pair_levels") == 3
    ib.set_tuning("msm_pair_levels", None)
    assert ib.get_tuning("msm_pair_levels") == -1
    ib.set_tuning("msm_pair_levels", old)
    with pytest.raises(ib.IcicleError):
        ib.set_tuning("no_such_knob", 1)
    assert capi.lib.b200_get_tuning(b"no_such_knob") == -1
    for lo, hi in ((0, 10), (3, 8), (10, 3)):
        assert ib.shard_range(hi, 4, 0)[0] == 0


DIST_NTT_WORKER = r'''
import os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "oracle")); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch, torch.distributed as dist, numpy as np
import common, port
from icicle_b200 import utils, shard_range
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank, G = dist.get_rank(), dist.get_world_size()
fp = utils.field_params("babybear"); p = fp["p"]
a_log, b_log = 3, 4
A, B, N = 1 << a_log, 1 << b_log, 1 << (a_log + b_log)
w = pow(fp["rou"], 1 << (fp["two_adicity"] - (a_log + b_log)), p)
x = common.rand_field_elems("babybear", N, 77, as_ints=True)
for inverse in (False, True):
    ww = pow(w, -1, p) if inverse else w
    # column slab of the A x B row-major view of the natural-order array (b200_ntt_dist_phase1's input layout)
    lo, hi = shard_range(B, G, rank)
    cols = hi - lo
    slab = [[x[r * B + c] for c in range(lo, hi)] for r in range(A)]
    # phase 1: A-point NTTs down the local columns (the CPU oracle stands in for the GPU), then w_N^(col * k) on element (k, col)
    wa = pow(w, B, p)
    for c in range(cols):
        colv = port.ntt([slab[r][c] for r in range(A)], wa, p, inverse=inverse)
        for k in range(A):
            slab[k][c] = colv[k] * pow(ww, (lo + c) * k, p) % p
    # the ONE exchange: block s (rows of rank s) goes to rank s; every rank receives its blocks in source-rank order
    rows_lo, rows_hi = shard_range(A, G, rank)
    send = torch.tensor([[slab[k][c] for c in range(cols)] for k in range(A)], dtype=torch.int64)
    gathered = [torch.zeros_like(send) for _ in range(G)]
    dist.all_gather(gathered, send)               # gloo has no all_to_all_single on CPU: gather everything, keep my row block of each source
    rows = [[0] * B for _ in range(rows_hi - rows_lo)]
    for src in range(G):
        slo, shi = shard_range(B, G, src)
        blk = gathered[src][rows_lo:rows_hi]
        for k in range(rows_hi - rows_lo):
            for c in range(shi - slo):
                rows[k][slo + c] = int(blk[k][c])
    # phase 2: B-point NTTs along the rows -> X[kb * A + k]; my output is the column slab [B][A/G] of the B x A view of the result
    wb = pow(w, A, p)
    out = [port.ntt(rows[k], wb, p, inverse=inverse) for k in range(rows_hi - rows_lo)]
    full = port.ntt(x, w, p, inverse=inverse)
    for kb in range(B):
        for k in range(rows_hi - rows_lo):
            assert out[k][kb] == full[kb * A + rows_lo + k], (inverse, kb, k)
dist.barrier(); dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_distributed_ntt_decomposition_two_ranks_gloo(tmp_path):
    """The data movement and index arithmetic of ONE transform spanning ranks (b200_ntt_dist_phase1 -> all-to-all -> phase2:
    column slabs of the A x B view in, column slabs of the B x A view of the natural-order result out, forward and inverse), run on
    two gloo ranks with the CPU oracle standing in for the two local GPU phases; the GPU phases themselves are tested against the
    single-device transform in tests/test_gpu_multi.py."""
    script = tmp_path / "dist_ntt_worker.py"
    script.write_text(DIST_NTT_WORKER)
    port_no = str(30500 + os.getpid() % 1000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port_no, str(r)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs


def test_multi_gpu_entry_points_fail_loudly_without_a_device():
    """No CUDA device (this container): the orchestrators must return the reference's error codes, not fall back to anything."""
    import numpy as np
    import icicle_b200 as ib
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    s = np.zeros((4, 8), dtype=np.uint32)
    P = np.zeros((4, 16), dtype=np.uint32)
    with pytest.raises(ib.IcicleError):
        ib.msm_multi_gpu(ib.Curve.BN254_G1, s, P, 4, n_devices=2)
    with pytest.raises(ib.IcicleError):
        ib.ntt_multi_gpu(ib.Field.BN254_FR, s, 4, ib.NTTDir.kForward, n_devices=2)
    with pytest.raises(ib.IcicleError):
        ib.msm(ib.Curve.BN254_G1, s, P, 4)
