"""The drop-in proof: the UNMODIFIED reference frontend (oracle/_ref: libicicle_device/field/curve built from
/root/reference sources) dlopens our backend DSOs (build/backend/<curve>/libicicle_backend_cuda_*.so) through its own
icicle_load_backend, and `bn254_msm` / `bn254_ntt` / `bn254_vector_*` dispatch to the B200 kernels when the active device
is "CUDA" -- exactly how the reference's own differential tests run (icicle/tests/test_base.h:28-48: main device = first
non-CPU device, reference device = "CPU")."""
import ctypes as C
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BACKEND_DIR = os.path.join(ROOT, "build", "backend", "bn254")


@pytest.fixture(scope="module")
def r():
    ref_icicle = pytest.importorskip("ref_icicle")
    if not ref_icicle.available("bn254") or not os.path.exists(os.path.join(BACKEND_DIR, "libicicle_backend_cuda_device.so")):
        pytest.skip("reference frontend or backend DSOs not built")
    ref = ref_icicle.get("bn254")
    if "CUDA" not in ref.registered_devices():
        assert ref.load_backend(BACKEND_DIR) == 0
    assert ref.registered_devices()[:2] == ["CUDA", "CPU"] or set(ref.registered_devices()) == {"CUDA", "CPU"}
    yield ref
    ref.set_device("CPU", 0)


def test_device_api(r):
    """icicle/tests/test_device_api.cpp:17-189 equivalents through the reference runtime's C API."""
    d = r.dev
    r.set_device("CUDA", 0)
    with pytest.raises(Exception):
        r.set_device("CUDA", 1000)  # INVALID_DEVICE (test_device_api.cpp:183-189)
    r.set_device("CUDA", 0)
    n = 1 << 16
    src = np.arange(n, dtype=np.uint32)
    dst = np.zeros_like(src)
    ptr = C.c_void_p()
    d.icicle_malloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
    assert d.icicle_malloc(C.byref(ptr), src.nbytes) == 0
    d.icicle_copy_to_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    d.icicle_copy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    assert d.icicle_copy_to_device(ptr, src.ctypes.data, src.nbytes) == 0
    assert d.icicle_copy_to_host(dst.ctypes.data, ptr, src.nbytes) == 0
    assert np.array_equal(src, dst)
    d.icicle_memset.argtypes = [C.c_void_p, C.c_int, C.c_size_t]
    assert d.icicle_memset(ptr, 0, src.nbytes) == 0
    assert d.icicle_copy_to_host(dst.ctypes.data, ptr, src.nbytes) == 0
    assert not dst.any()
    stream = C.c_void_p()
    d.icicle_create_stream.argtypes = [C.POINTER(C.c_void_p)]
    assert d.icicle_create_stream(C.byref(stream)) == 0
    d.icicle_stream_synchronize.argtypes = [C.c_void_p]
    assert d.icicle_stream_synchronize(stream) == 0
    d.icicle_destroy_stream.argtypes = [C.c_void_p]
    assert d.icicle_destroy_stream(stream) == 0
    total, free = C.c_size_t(), C.c_size_t()
    d.icicle_get_available_memory.argtypes = [C.POINTER(C.c_size_t), C.POINTER(C.c_size_t)]
    assert d.icicle_get_available_memory(C.byref(total), C.byref(free)) == 0 and total.value > (100 << 30)
    d.icicle_free.argtypes = [C.c_void_p]
    assert d.icicle_free(ptr) == 0


def test_msm_main_vs_ref_device(r):
    rng = random.Random(1)
    for n in (1, (1 << 12) - rng.randrange(60), 1 << 15):
        r.set_device("CPU", 0)
        s, P = r.generate_scalars(n), r.generate_affine_points(n)
        exp = r.msm(s, P, n)
        r.set_device("CUDA", 0)
        got = r.msm(s, P, n)
        assert r.projective_eq(got[0], exp[0])
    # msm_bitsize (test_curve_api.cpp:81-123), a few sizes; batch + precompute (test_curve_api.cpp:125-171)
    n = 1 << 10
    for bitsize in (1, 7, 64, 200):
        r.set_device("CPU", 0)
        exp = r.msm(s[:n], P[:n], n, bitsize=bitsize)
        r.set_device("CUDA", 0)
        got = r.msm(s[:n], P[:n], n, bitsize=bitsize)
        assert r.projective_eq(got[0], exp[0])
    batch, pf = 3, 4
    r.set_device("CPU", 0)
    exp = r.msm(s[: n * batch], P[:n], n, batch_size=batch)
    r.set_device("CUDA", 0)
    pre = r.msm_precompute_bases(P[:n], n, precompute_factor=pf)
    got = r.msm(s[: n * batch], pre, n, batch_size=batch, precompute_factor=pf)
    for b in range(batch):
        assert r.projective_eq(got[b], exp[b])
    # G2
    r.set_device("CPU", 0)
    P2 = r.generate_affine_points(n, g2=True)
    exp = r.msm(s[:n], P2, n, g2=True)
    r.set_device("CUDA", 0)
    got = r.msm(s[:n], P2, n, g2=True)
    assert r.projective_eq(got[0], exp[0], g2=True)


def test_ntt_and_vec_ops_main_vs_ref_device(r):
    logn = 13
    n = 1 << logn
    r.set_device("CPU", 0)
    root = r.get_root_of_unity(1 << (logn + 2))
    r.ntt_init_domain(root)
    x = r.generate_scalars(n * 2)
    r.set_device("CUDA", 0)
    r.ntt_init_domain(root)
    assert np.array_equal(r.get_root_of_unity_from_domain(logn), (r.set_device("CPU", 0), r.get_root_of_unity_from_domain(logn))[1])
    for direction in (0, 1):
        for ordering in range(6):
            for columns in (False, True):
                r.set_device("CPU", 0)
                exp = r.ntt(x, n, direction, batch_size=2, columns_batch=columns, ordering=ordering if ordering < 4 else 0)
                r.set_device("CUDA", 0)
                got = r.ntt(x, n, direction, batch_size=2, columns_batch=columns, ordering=ordering)
                if ordering < 4:
                    assert np.array_equal(got, exp), (direction, ordering, columns)
    for op in ("vector_add", "vector_sub", "vector_mul", "vector_accumulate"):
        r.set_device("CPU", 0)
        exp = r.vec2(op, x[:n], x[n:], n)
        r.set_device("CUDA", 0)
        assert np.array_equal(r.vec2(op, x[:n], x[n:], n), exp), op
    r.set_device("CUDA", 0)
    r.ntt_release_domain()
    r.set_device("CPU", 0)
    r.ntt_release_domain()


def test_polynomial_api_on_device(r):
    """the reference's Polynomial class (device-agnostic default backend) running on our device through
    the registered factory + ntt + vec-ops; checked against the CPU device (icicle/tests/test_polynomial_api.cpp flows)."""
    f = r.field
    logn = 10
    r.set_device("CPU", 0)
    root = r.get_root_of_unity(1 << (logn + 3))
    a_c, b_c = r.generate_scalars(1 << logn), r.generate_scalars(1 << (logn - 1))
    dom = r.generate_scalars(16)
    f.bn254_polynomial_create_from_coefficients.restype = C.c_void_p
    f.bn254_polynomial_create_from_coefficients.argtypes = [C.c_void_p, C.c_size_t]
    for fn in ("multiply", "add", "subtract", "quotient", "remainder"):
        getattr(f, f"bn254_polynomial_{fn}").restype = C.c_void_p
        getattr(f, f"bn254_polynomial_{fn}").argtypes = [C.c_void_p, C.c_void_p]
    f.bn254_polynomial_degree.restype = C.c_int64
    f.bn254_polynomial_degree.argtypes = [C.c_void_p]
    f.bn254_polynomial_copy_coeffs_range.restype = C.c_int64
    f.bn254_polynomial_copy_coeffs_range.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    f.bn254_polynomial_evaluate_on_domain.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    f.bn254_polynomial_delete.argtypes = [C.c_void_p]

    def run(device):
        r.set_device(device, 0)
        r.ntt_init_domain(root)
        A = f.bn254_polynomial_create_from_coefficients(a_c.ctypes.data, a_c.shape[0])
        B = f.bn254_polynomial_create_from_coefficients(b_c.ctypes.data, b_c.shape[0])
        out = {}
        for name in ("multiply", "add", "subtract", "quotient", "remainder"):
            P = getattr(f, f"bn254_polynomial_{name}")(A, B)
            deg = f.bn254_polynomial_degree(P)
            coeffs = np.zeros((deg + 1, 8), dtype=np.uint32)
            f.bn254_polynomial_copy_coeffs_range(P, coeffs.ctypes.data, 0, deg)
            out[name] = (deg, coeffs)
            if name == "multiply":
                ev = np.zeros((16, 8), dtype=np.uint32)
                f.bn254_polynomial_evaluate_on_domain(P, dom.ctypes.data, 16, ev.ctypes.data)
                out["evals"] = (16, ev)
            f.bn254_polynomial_delete(P)
        f.bn254_polynomial_delete(A)
        f.bn254_polynomial_delete(B)
        r.ntt_release_domain()
        return out

    exp = run("CPU")
    got = run("CUDA")
    for k in exp:
        assert exp[k][0] == got[k][0], k
        assert np.array_equal(exp[k][1], got[k][1]), k
    r.set_device("CPU", 0)


def test_ecntt_main_vs_ref_device(r):
    """`bn254_ecntt` through the unmodified frontend: REGISTER_ECNTT_BACKEND("CUDA", ...) in our curve DSO vs the CPU device
    (icicle/tests/test_curve_api.cpp:293-400 compares the same way, as group elements)."""
    import common
    if not hasattr(r.curve, "bn254_ecntt"):
        pytest.skip("reference built without ECNTT")
    logn = 6
    n = 1 << logn
    r.set_device("CPU", 0)
    root = r.get_root_of_unity(1 << (logn + 1))
    r.ntt_init_domain(root)
    P = common.affine_to_projective_limbs(r.generate_affine_points(n), 8)
    r.set_device("CUDA", 0)
    r.ntt_init_domain(root)
    for direction in (0, 1):
        for ordering in (0, 1, 2):
            r.set_device("CPU", 0)
            exp = r.ecntt(P, n, direction, ordering=ordering)
            r.set_device("CUDA", 0)
            got = r.ecntt(P, n, direction, ordering=ordering)
            for i in range(n):
                assert r.projective_eq(got[i], exp[i]), (direction, ordering, i)
    r.set_device("CUDA", 0)
    r.ntt_release_domain()
    r.set_device("CPU", 0)
    r.ntt_release_domain()


def test_precompute_device_output_default_config(r):
    """the Rust wrapper's precompute_bases always hands a DeviceSlice output with the user's config unchanged,
    so are_results_on_device / are_points_on_device stay false for DEVICE buffers (wrappers/rust/icicle-core/src/msm/mod.rs).
    The backend must look at the pointers, not only at the flags: device in / device out with a default config, then the
    table is used by an MSM straight from device memory."""
    n, pf = 300, 4
    r.set_device("CPU", 0)
    s, P = r.generate_scalars(n), r.generate_affine_points(n)
    exp = r.msm(s, P, n)
    r.set_device("CUDA", 0)
    d_in, d_out = r.malloc(P.nbytes), r.malloc(P.nbytes * pf)
    r.copy_to_device(d_in, P)
    r.msm_precompute_bases_raw(d_in, n, d_out, precompute_factor=pf)          # every *_on_device flag left false
    table = np.zeros((n * pf, 16), dtype=np.uint32)
    r.copy_to_host(table, d_out)
    assert np.array_equal(table, r.msm_precompute_bases(P, n, precompute_factor=pf))   # host/host path of the same backend
    got = r.msm(s, table, n, precompute_factor=pf)
    assert r.projective_eq(got[0], exp[0])
    r.free(d_in)
    r.free(d_out)
    r.set_device("CPU", 0)


def test_multi_gpu_extension_key_through_frontend(r):
    """The opt-in ConfigExtension key "multi_gpu"  set by an unmodified caller: bn254_msm / bn254_ntt with host
    vectors shard over that many devices inside the backend (one host thread per device); unknown to the CPU backend, ignored
    there.  With one GPU in the box the request is clamped by the orchestrator's device check -> INVALID_DEVICE is NOT what a
    caller wants, so the test asks for min(2, device count) devices and also covers the single-device fall-through."""
    import icicle_b200 as ib
    k = min(2, ib.get_device_count())
    n = (1 << 12) + 7
    r.set_device("CPU", 0)
    s, P = r.generate_scalars(n * 2), r.generate_affine_points(n)
    exp = r.msm(s, P, n, batch_size=2)
    logn = 10
    root = r.get_root_of_unity(1 << logn)
    x = r.generate_scalars(3 << logn)
    r.ntt_init_domain(root)
    exp_ntt = r.ntt(x, 1 << logn, 0, batch_size=3)
    r.ntt_release_domain()
    r.set_device("CUDA", 0)
    ext = r.config_extension(multi_gpu=k)
    got = r.msm(s, P, n, batch_size=2, ext=ext)
    assert r.projective_eq(got[0], exp[0]) and r.projective_eq(got[1], exp[1])
    got1 = r.msm(s[:n], P, n, ext=ext)                                          # one MSM: point-range split + ec_sum
    assert r.projective_eq(got1[0], exp[0])
    r.ntt_init_domain(root)
    assert np.array_equal(r.ntt(x, 1 << logn, 0, batch_size=3, ext=ext), exp_ntt)
    r.ntt_release_domain()
    r.set_device("CPU", 0)


def test_unaligned_device_pointers(r):
    """storage<N> promises only 4-byte alignment (icicle/include/icicle/math/storage.h:4-9); a caller may hand a device pointer
    at a 4-byte offset inside an icicle_malloc'd buffer.  The kernels use 128-bit accesses, so the backend stages such buffers
    through aligned scratch instead of faulting."""
    import icicle_b200 as ib
    n = 1 << 10
    r.set_device("CPU", 0)
    a, b = r.generate_scalars(n), r.generate_scalars(n)
    exp = r.vec2("vector_mul", a, b, n)
    ib.set_device(0)
    da = ib.device_empty(n * 8 + 1)
    db = ib.device_empty(n * 8 + 1)
    do = ib.device_empty(n * 8 + 1)
    ib.capi.check(ib.capi.lib.b200_copy_to_device(da.data_ptr() + 4, a.ctypes.data, a.nbytes, None, 0), "h2d")
    ib.capi.check(ib.capi.lib.b200_copy_to_device(db.data_ptr() + 4, b.ctypes.data, b.nbytes, None, 0), "h2d")
    got = ib.vector_mul(ib.Field.BN254_FR, da[1:].view(n, 8), db[1:].view(n, 8), n, ib.VecOpsConfig(), do[1:].view(n, 8))
    assert np.array_equal(ib.to_host(got), exp)


@pytest.mark.parametrize("family", ["bls12_381", "bls12_377", "bw6_761", "grumpkin", "babybear", "koalabear", "stark252", "goldilocks", "m31"])
def test_dropin_other_reference_builds(family):
    """Row 15 of the round-1 verdict ("only the bn254 DSO set is exercised"): every other reference build of the north-star list loads
    ITS backend DSOs (build/backend/<family>) in its own process and compares the hot-path symbols between Device{"CUDA"} and
    Device{"CPU"} -- MSM (+G2, batch, bitsize, precompute), NTT (orderings, batch rows/columns, coset, extension NTT), vec-ops,
    the quartic-extension vec-op family, Montgomery conversions (tests/dropin_worker.py)."""
    import subprocess
    import sys
    ref_icicle = pytest.importorskip("ref_icicle")
    if not ref_icicle.available(family) or not os.path.exists(os.path.join(ROOT, "build", "backend", family, "libicicle_backend_cuda_device.so")):
        pytest.skip(f"reference build or backend DSOs for {family} not present")
    p = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "dropin_worker.py"), family], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-1500:] + p.stderr[-3000:]
