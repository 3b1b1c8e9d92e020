"""TEST INFRASTRUCTURE ONLY -- ctypes driver for the UNMODIFIED reference (ingonyama-zk/icicle) CPU backend built by
oracle/Makefile into oracle/_ref/<curve|field>/.  It is the parity oracle and the CPU baseline (cpu_baseline.kind =
"reference"): tests/, __graft_entry__.smoke() and bench.py's reference/cpu_baseline legs are the only importers.
Nothing in icicle_b200/ may import this module.

The C entry points and config structs bound here are the reference's own FFI surface (what its Rust/Go wrappers bind):
  <prefix>_msm / _g2_msm / _msm_precompute_bases          icicle/src/msm.cpp:12-16,28-32,45-49
  <prefix>_ntt / _ntt_init_domain / _ntt_release_domain / _get_root_of_unity[_from_domain]   icicle/src/ntt.cpp:11-15,26-30,41-46,55-79
  <prefix>_vector_{add,sub,mul,accumulate}, _scalar_{add,sub,mul}_vec, _scalar_convert_montgomery, _bit_reverse, _slice
                                                           icicle/src/vec_ops.cpp:71,120,169,...,362,404,440,472
  <prefix>_matrix_transpose                                icicle/src/matrix_ops.cpp:75-79
  <prefix>_{affine,projective}_convert_montgomery          icicle/src/curves/montgomery_conversion.cpp:13-17,47-51
  <prefix>_generate_random / _generate_affine_points / _projective_eq / _to_affine      icicle/src/{fields,curves}/ffi_extern.cpp
  icicle_set_device / icicle_load_backend ...              icicle/src/runtime.cpp
  MSMConfig (msm.h:21-53), NTTConfig<S> (ntt.h:52-64), NTTInitDomainConfig (ntt.h:92-96), VecOpsConfig (vec_ops.h:19-44),
  Device (device.h:14-17)
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(_HERE, "_ref")

# name -> (is_curve, scalar limbs, base-field limbs, has g2, g2 coordinate limbs)
TARGETS = {
    "bn254": dict(curve=True, s=8, q=8, g2=True, g2q=16),
    "bls12_381": dict(curve=True, s=8, q=12, g2=True, g2q=24),
    "bls12_377": dict(curve=True, s=8, q=12, g2=True, g2q=24),
    "bw6_761": dict(curve=True, s=12, q=24, g2=True, g2q=24),
    "grumpkin": dict(curve=True, s=8, q=8, g2=False, g2q=0),
    "babybear": dict(curve=False, s=1),
    "koalabear": dict(curve=False, s=1),
    "stark252": dict(curve=False, s=8),
    "m31": dict(curve=False, s=1),
    "goldilocks": dict(curve=False, s=2),
}


def available(name):
    d = os.path.join(REF_DIR, name)
    return os.path.exists(os.path.join(d, "libicicle_device.so")) and os.path.exists(os.path.join(d, f"libicicle_field_{name}.so"))


class Device(C.Structure):
    _fields_ = [("type", C.c_char * 64), ("id", C.c_int)]


class MSMConfig(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("precompute_factor", C.c_int), ("c", C.c_int), ("bitsize", C.c_int), ("batch_size", C.c_int),
                ("are_points_shared_in_batch", C.c_bool), ("are_scalars_on_device", C.c_bool), ("are_scalars_montgomery_form", C.c_bool),
                ("are_points_on_device", C.c_bool), ("are_points_montgomery_form", C.c_bool), ("are_results_on_device", C.c_bool),
                ("is_async", C.c_bool), ("ext", C.c_void_p)]


class VecOpsConfig(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("is_a_on_device", C.c_bool), ("is_b_on_device", C.c_bool), ("is_result_on_device", C.c_bool),
                ("is_async", C.c_bool), ("batch_size", C.c_int), ("columns_batch", C.c_bool), ("ext", C.c_void_p)]


class NTTInitDomainConfig(C.Structure):
    _fields_ = [("stream", C.c_void_p), ("is_async", C.c_bool), ("ext", C.c_void_p)]


def make_ntt_config_type(limbs):
    class NTTConfig(C.Structure):
        _fields_ = [("stream", C.c_void_p), ("coset_gen", C.c_uint32 * limbs), ("batch_size", C.c_int), ("columns_batch", C.c_bool),
                    ("ordering", C.c_int), ("are_inputs_on_device", C.c_bool), ("are_outputs_on_device", C.c_bool), ("is_async", C.c_bool),
                    ("ext", C.c_void_p)]
    return NTTConfig


class RefError(RuntimeError):
    pass


def _chk(code, what):
    if code != 0:
        raise RefError(f"reference {what} failed with eIcicleError {code}")


class Ref:
    """One loaded reference build (one curve or field), CPU device selected."""

    def __init__(self, name):
        if not available(name):
            raise ImportError(f"reference build oracle/_ref/{name} not found (make -C oracle ref ...)")
        self.name = name
        self.t = TARGETS[name]
        d = os.path.join(REF_DIR, name)
        mode = C.RTLD_GLOBAL
        self.dev = C.CDLL(os.path.join(d, "libicicle_device.so"), mode=mode)
        self.field = C.CDLL(os.path.join(d, f"libicicle_field_{name}.so"), mode=mode)
        self.curve = C.CDLL(os.path.join(d, f"libicicle_curve_{name}.so"), mode=mode) if self.t["curve"] else None
        self.NTTConfig = make_ntt_config_type(self.t["s"])
        cpu = Device(b"CPU", 0)
        self.dev.icicle_set_device.argtypes = [C.POINTER(Device)]
        _chk(self.dev.icicle_set_device(C.byref(cpu)), "icicle_set_device(CPU)")

    def set_device(self, dtype="CPU", idx=0):
        d = Device(dtype.encode(), idx)
        _chk(self.dev.icicle_set_device(C.byref(d)), f"icicle_set_device({dtype})")

    def load_backend(self, path, recursive=True):
        self.dev.icicle_load_backend.argtypes = [C.c_char_p, C.c_bool]
        return self.dev.icicle_load_backend(path.encode(), recursive)

    def registered_devices(self):
        buf = C.create_string_buffer(256)
        self.dev.icicle_get_registered_devices.argtypes = [C.c_char_p, C.c_size_t]
        _chk(self.dev.icicle_get_registered_devices(buf, 256), "get_registered_devices")
        return buf.value.decode().split(",")

    def config_extension(self, **ints):
        """ConfigExtension* with integer keys (icicle/src/config_extension.cpp:7-15); pass it as `ext=` of a config.  Leaked on
        purpose (tests only)."""
        self.dev.create_config_extension.restype = C.c_void_p
        self.dev.config_extension_set_int.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
        ext = self.dev.create_config_extension()
        for k, v in ints.items():
            self.dev.config_extension_set_int(ext, k.encode(), int(v))
        return ext

    def malloc(self, nbytes):
        """icicle_malloc on the active device (icicle/src/runtime.cpp)."""
        ptr = C.c_void_p()
        self.dev.icicle_malloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        _chk(self.dev.icicle_malloc(C.byref(ptr), nbytes), "icicle_malloc")
        return ptr

    def free(self, ptr):
        self.dev.icicle_free.argtypes = [C.c_void_p]
        _chk(self.dev.icicle_free(ptr), "icicle_free")

    def copy_to_device(self, dptr, host):
        self.dev.icicle_copy_to_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _chk(self.dev.icicle_copy_to_device(dptr, host.ctypes.data, host.nbytes), "icicle_copy_to_device")

    def copy_to_host(self, host, dptr):
        self.dev.icicle_copy_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        _chk(self.dev.icicle_copy_to_host(host.ctypes.data, dptr, host.nbytes), "icicle_copy_to_host")

    def msm_precompute_bases_raw(self, bases_ptr, n, out_ptr, g2=False, **cfgkw):
        """<curve>_msm_precompute_bases with caller-supplied raw pointers (device or host) and config flags as given."""
        cfg = self.msm_config(**cfgkw)
        fn = self._f(self.curve, ("g2_" if g2 else "") + "msm_precompute_bases")
        fn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _chk(fn(bases_ptr, C.c_int(n), C.byref(cfg), out_ptr), "msm_precompute_bases")

    # ---- helpers -------------------------------------------------------------------------------------------------------
    def _f(self, lib, sym):
        return getattr(lib, f"{self.name}_{sym}")

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def msm_config(self, **kw):
        cfg = MSMConfig(None, 1, 0, 0, 1, True, False, False, False, False, False, False, None)
        for k, v in kw.items():
            setattr(cfg, k, v)
        return cfg

    def vec_config(self, **kw):
        cfg = VecOpsConfig(None, False, False, False, False, 1, False, None)
        for k, v in kw.items():
            setattr(cfg, k, v)
        return cfg

    def ntt_config(self, coset_gen=None, **kw):
        cfg = self.NTTConfig()
        cfg.stream = None
        one = [1] + [0] * (self.t["s"] - 1)
        cg = list(np.asarray(coset_gen, dtype=np.uint32).reshape(-1)) if coset_gen is not None else one
        for i, v in enumerate(cg):
            cfg.coset_gen[i] = int(v)
        cfg.batch_size, cfg.columns_batch, cfg.ordering = 1, False, 0
        cfg.are_inputs_on_device = cfg.are_outputs_on_device = cfg.is_async = False
        cfg.ext = None
        for k, v in kw.items():
            setattr(cfg, k, v)
        return cfg

    # ---- generators ------------------------------------------------------------------------------------------------------
    def generate_scalars(self, n):
        out = np.zeros((n, self.t["s"]), dtype=np.uint32)
        self._f(self.field, "generate_random")(self._p(out), C.c_int(n))
        return out

    def generate_affine_points(self, n, g2=False):
        limbs = 2 * (self.t["g2q"] if g2 else self.t["q"])
        out = np.zeros((n, limbs), dtype=np.uint32)
        self._f(self.curve, ("g2_" if g2 else "") + "generate_affine_points")(self._p(out), C.c_int(n))
        return out

    # ---- MSM -------------------------------------------------------------------------------------------------------------
    def msm(self, scalars, bases, msm_size, g2=False, **cfgkw):
        cfg = self.msm_config(**cfgkw)
        limbs = 3 * (self.t["g2q"] if g2 else self.t["q"])
        scalars = np.ascontiguousarray(scalars, dtype=np.uint32)
        bases = np.ascontiguousarray(bases, dtype=np.uint32)
        out = np.zeros((cfg.batch_size, limbs), dtype=np.uint32)
        fn = self._f(self.curve, ("g2_" if g2 else "") + "msm")
        _chk(fn(self._p(scalars), self._p(bases), C.c_int(msm_size), C.byref(cfg), self._p(out)), "msm")
        return out

    def msm_precompute_bases(self, bases, n, g2=False, **cfgkw):
        cfg = self.msm_config(**cfgkw)
        limbs = 2 * (self.t["g2q"] if g2 else self.t["q"])
        bases = np.ascontiguousarray(bases, dtype=np.uint32)
        out = np.zeros((n * cfg.precompute_factor, limbs), dtype=np.uint32)
        fn = self._f(self.curve, ("g2_" if g2 else "") + "msm_precompute_bases")
        _chk(fn(self._p(bases), C.c_int(n), C.byref(cfg), self._p(out)), "msm_precompute_bases")
        return out

    def projective_eq(self, p1, p2, g2=False):
        fn = self._f(self.curve, ("g2_" if g2 else "") + "projective_eq")
        fn.restype = C.c_bool
        a = np.ascontiguousarray(p1, dtype=np.uint32)
        b = np.ascontiguousarray(p2, dtype=np.uint32)
        return bool(fn(self._p(a), self._p(b)))

    def to_affine(self, p, g2=False):
        q = self.t["g2q"] if g2 else self.t["q"]
        a = np.ascontiguousarray(p, dtype=np.uint32).reshape(3 * q)
        out = np.zeros(2 * q, dtype=np.uint32)
        self._f(self.curve, ("g2_" if g2 else "") + "to_affine")(self._p(a), self._p(out))
        return out

    def affine_convert_montgomery(self, pts, n, is_into, g2=False):
        cfg = self.vec_config()
        pts = np.ascontiguousarray(pts, dtype=np.uint32)
        out = np.zeros_like(pts)
        fn = self._f(self.curve, ("g2_" if g2 else "") + "affine_convert_montgomery")
        _chk(fn(self._p(pts), C.c_uint64(n), C.c_bool(is_into), C.byref(cfg), self._p(out)), "affine_convert_montgomery")
        return out

    def projective_convert_montgomery(self, pts, n, is_into, g2=False):
        cfg = self.vec_config()
        pts = np.ascontiguousarray(pts, dtype=np.uint32)
        out = np.zeros_like(pts)
        fn = self._f(self.curve, ("g2_" if g2 else "") + "projective_convert_montgomery")
        _chk(fn(self._p(pts), C.c_uint64(n), C.c_bool(is_into), C.byref(cfg), self._p(out)), "projective_convert_montgomery")
        return out

    # ---- NTT -------------------------------------------------------------------------------------------------------------
    def get_root_of_unity(self, max_size):
        out = np.zeros(self.t["s"], dtype=np.uint32)
        _chk(self._f(self.field, "get_root_of_unity")(C.c_uint64(max_size), self._p(out)), "get_root_of_unity")
        return out

    def ntt_init_domain(self, root):
        cfg = NTTInitDomainConfig(None, False, None)
        r = np.ascontiguousarray(root, dtype=np.uint32)
        _chk(self._f(self.field, "ntt_init_domain")(self._p(r), C.byref(cfg)), "ntt_init_domain")

    def ntt_release_domain(self):
        _chk(self._f(self.field, "ntt_release_domain")(), "ntt_release_domain")

    def get_root_of_unity_from_domain(self, logn):
        out = np.zeros(self.t["s"], dtype=np.uint32)
        _chk(self._f(self.field, "get_root_of_unity_from_domain")(C.c_uint64(logn), self._p(out)), "get_root_of_unity_from_domain")
        return out

    def ntt(self, inp, size, direction, coset_gen=None, out=None, **cfgkw):
        cfg = self.ntt_config(coset_gen, **cfgkw)
        inp = np.ascontiguousarray(inp, dtype=np.uint32)
        if out is None:
            out = np.zeros_like(inp)
        _chk(self._f(self.field, "ntt")(self._p(inp), C.c_int(size), C.c_int(direction), C.byref(cfg), self._p(out)), "ntt")
        return out

    def ecntt(self, inp, size, direction, coset_gen=None, **cfgkw):
        """<curve>_ecntt (icicle/src/ecntt.cpp:8-12): NTT over G1 projective points with scalar-field twiddles."""
        cfg = self.ntt_config(coset_gen, **cfgkw)
        inp = np.ascontiguousarray(inp, dtype=np.uint32)
        out = np.zeros_like(inp)
        _chk(self._f(self.curve, "ecntt")(self._p(inp), C.c_int(size), C.c_int(direction), C.byref(cfg), self._p(out)), "ecntt")
        return out

    def extension_ntt(self, inp, size, direction, coset_gen=None, **cfgkw):
        """<field>_extension_ntt (icicle/src/ntt.cpp:90-95): elements are quartic extension elements (4 base limbs rows)."""
        cfg = self.ntt_config(coset_gen, **cfgkw)
        inp = np.ascontiguousarray(inp, dtype=np.uint32)
        out = np.zeros_like(inp)
        _chk(self._f(self.field, "extension_ntt")(self._p(inp), C.c_int(size), C.c_int(direction), C.byref(cfg), self._p(out)), "extension_ntt")
        return out

    # ---- vec ops ---------------------------------------------------------------------------------------------------------
    def vec2(self, op, a, b, size, **cfgkw):
        cfg = self.vec_config(**cfgkw)
        a = np.ascontiguousarray(a, dtype=np.uint32)
        b = np.ascontiguousarray(b, dtype=np.uint32)
        if op == "vector_accumulate":
            a = a.copy()
            _chk(self._f(self.field, op)(self._p(a), self._p(b), C.c_uint64(size), C.byref(cfg)), op)
            return a
        out = np.zeros_like(b)
        _chk(self._f(self.field, op)(self._p(a), self._p(b), C.c_uint64(size), C.byref(cfg), self._p(out)), op)
        return out

    def scalar_convert_montgomery(self, a, size, is_into, **cfgkw):
        cfg = self.vec_config(**cfgkw)
        a = np.ascontiguousarray(a, dtype=np.uint32)
        out = np.zeros_like(a)
        _chk(self._f(self.field, "scalar_convert_montgomery")(self._p(a), C.c_uint64(size), C.c_bool(is_into), C.byref(cfg), self._p(out)), "convert_montgomery")
        return out

    def bit_reverse(self, a, size, **cfgkw):
        cfg = self.vec_config(**cfgkw)
        a = np.ascontiguousarray(a, dtype=np.uint32)
        out = np.zeros_like(a)
        _chk(self._f(self.field, "bit_reverse")(self._p(a), C.c_uint64(size), C.byref(cfg), self._p(out)), "bit_reverse")
        return out

    def matrix_transpose(self, a, rows, cols, **cfgkw):
        cfg = self.vec_config(**cfgkw)
        a = np.ascontiguousarray(a, dtype=np.uint32)
        out = np.zeros_like(a)
        _chk(self._f(self.field, "matrix_transpose")(self._p(a), C.c_uint32(rows), C.c_uint32(cols), C.byref(cfg), self._p(out)), "matrix_transpose")
        return out

    def slice(self, a, offset, stride, size_in, size_out, **cfgkw):
        cfg = self.vec_config(**cfgkw)
        a = np.ascontiguousarray(a, dtype=np.uint32)
        out = np.zeros((size_out * cfg.batch_size, a.shape[-1]), dtype=np.uint32)
        _chk(self._f(self.field, "slice")(self._p(a), C.c_uint64(offset), C.c_uint64(stride), C.c_uint64(size_in), C.c_uint64(size_out),
                                          C.byref(cfg), self._p(out)), "slice")
        return out


_cache = {}


def get(name):
    """Load (once per process) the reference build for `name`.  NOTE: different curve builds each carry their own
    libicicle_device.so; load only one curve family per process to keep their dispatch tables apart."""
    if name not in _cache:
        _cache[name] = Ref(name)
    return _cache[name]
