"""TEST INFRASTRUCTURE -- the drop-in proof for ONE reference build in its own process (one family per process, oracle/ref_icicle.py):
the unmodified frontend `oracle/_ref/<family>` loads `build/backend/<family>/libicicle_backend_cuda_*.so` through its own
icicle_load_backend and every hot-path C symbol it exports is compared between Device{"CUDA"} (our kernels) and Device{"CPU"}
(the reference), the reference's own differential-test method (icicle/tests/test_base.h:28-48, test_curve_api.cpp:275-289,
test_mod_arithmetic_api.h).  usage: python tests/dropin_worker.py <family>; exit code 0 = every comparison passed."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import ref_icicle  # noqa: E402
import common  # noqa: E402


def main(family):
    r = ref_icicle.get(family)
    t = ref_icicle.TARGETS[family]
    bdir = os.path.join(ROOT, "build", "backend", family)
    assert r.load_backend(bdir) == 0
    assert "CUDA" in r.registered_devices(), r.registered_devices()
    both = lambda fn: ((r.set_device("CPU", 0), fn())[1], (r.set_device("CUDA", 0), fn())[1])
    checks = 0
    f = r.field
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    has_ntt = hasattr(f, f"{family}_ntt")
    # ---- vec-ops on the scalar field -------------------------------------------------------------------------------------
    n = 1000
    a, b = r.generate_scalars(n), r.generate_scalars(n)
    for op in ("vector_add", "vector_sub", "vector_mul", "vector_accumulate"):
        cpu, gpu = both(lambda: r.vec2(op, a, b, n))
        assert np.array_equal(cpu, gpu), op
        checks += 1
    for into in (True, False):
        cpu, gpu = both(lambda: r.scalar_convert_montgomery(a, n, into))
        assert np.array_equal(cpu, gpu), ("convert_montgomery", into)
        checks += 1
    # ---- NTT -------------------------------------------------------------------------------------------------------------
    if has_ntt:
        logn, batch = 11, 3
        root = r.get_root_of_unity(1 << logn)
        x = r.generate_scalars(batch << logn)
        for dev in ("CPU", "CUDA"):
            r.set_device(dev, 0)
            r.ntt_init_domain(root)
        for d in (0, 1):
            for o in (0, 1, 2, 3):
                for cols in (False, True):
                    cpu, gpu = both(lambda: r.ntt(x, 1 << logn, d, batch_size=batch, columns_batch=cols, ordering=o))
                    assert np.array_equal(cpu, gpu), ("ntt", d, o, cols)
                    checks += 1
        g = x[3].copy()
        cpu, gpu = both(lambda: r.ntt(x[: 1 << logn], 1 << logn, 0, coset_gen=g))
        assert np.array_equal(cpu, gpu), "ntt coset"
        checks += 1
        if hasattr(f, f"{family}_extension_ntt"):
            xe = r.generate_scalars(4 << logn).reshape(1 << logn, -1)
            for d in (0, 1):
                cpu, gpu = both(lambda: r.extension_ntt(xe, 1 << logn, d))
                assert np.array_equal(cpu, gpu), ("extension_ntt", d)
                checks += 1
        for dev in ("CPU", "CUDA"):
            r.set_device(dev, 0)
            r.ntt_release_domain()
    # ---- quartic-extension vec-ops (EXT_FIELD builds) -----------------------------------------------------------------------
    if hasattr(f, f"{family}_extension_vector_mul") and t["s"] == 1:
        m = 300
        ea = r.generate_scalars(4 * m).reshape(m, 4)
        eb = r.generate_scalars(4 * m).reshape(m, 4)
        eb[7] = 0
        sc = r.generate_scalars(m)

        def ext2(sym, x, y, out_like):
            o = np.zeros_like(out_like)
            cfg = r.vec_config()
            rc = getattr(f, f"{family}_extension_{sym}")(P(x), P(y), C.c_uint64(m), C.byref(cfg), P(o))
            assert rc == 0, (sym, rc)
            return o

        def ext1(sym, x, rows):
            o = np.zeros((rows, 4), dtype=np.uint32)
            cfg = r.vec_config()
            rc = getattr(f, f"{family}_extension_{sym}")(P(x), C.c_uint64(m), C.byref(cfg), P(o))
            assert rc == 0, (sym, rc)
            return o
        for sym in ("vector_add", "vector_sub", "vector_mul", "vector_div"):
            cpu, gpu = both(lambda: ext2(sym, ea, eb, ea))
            assert np.array_equal(cpu, gpu), ("extension", sym)
            checks += 1
        cpu, gpu = both(lambda: ext2("vector_mixed_mul", ea, sc, ea))
        assert np.array_equal(cpu, gpu), "extension mixed_mul"
        for sym, rows in (("vector_inv", m), ("vector_sum", 1), ("vector_product", 1)):
            cpu, gpu = both(lambda: ext1(sym, eb if sym == "vector_inv" else ea, rows))
            assert np.array_equal(cpu, gpu), ("extension", sym)
            checks += 1
    # ---- curve: MSM (+G2), precompute, Montgomery conversion ---------------------------------------------------------------------
    if t["curve"]:
        n = (1 << 10) + 3
        s = r.generate_scalars(n * 2)
        for g2 in ((False, True) if t["g2"] else (False,)):
            pts = r.generate_affine_points(n, g2=g2)
            pts[5] = 0
            cpu, gpu = both(lambda: r.msm(s[:n], pts, n, g2=g2))
            assert r.projective_eq(cpu[0], gpu[0], g2=g2), ("msm", g2)
            cpu, gpu = both(lambda: r.msm(s, pts, n, g2=g2, batch_size=2, bitsize=61))
            assert all(r.projective_eq(cpu[k], gpu[k], g2=g2) for k in range(2)), ("msm batch bitsize", g2)
            r.set_device("CUDA", 0)
            pre = r.msm_precompute_bases(pts, n, g2=g2, precompute_factor=3)
            got = r.msm(s[:n], pre, n, g2=g2, precompute_factor=3)
            r.set_device("CPU", 0)
            assert r.projective_eq(r.msm(s[:n], pts, n, g2=g2)[0], got[0], g2=g2), ("precompute", g2)
            cpu, gpu = both(lambda: r.affine_convert_montgomery(pts, n, True, g2=g2))
            assert np.array_equal(cpu, gpu), ("affine montgomery", g2)
            checks += 4
    r.set_device("CPU", 0)
    print(f"dropin_worker {family}: {checks} comparisons ok")


if __name__ == "__main__":
    main(sys.argv[1])
