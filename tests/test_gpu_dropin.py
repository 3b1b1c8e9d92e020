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
