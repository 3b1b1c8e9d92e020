"""Host-side mirror of the reference's MSM / NTT / vec-ops frontend (icicle/include/icicle/{msm,ntt,vec_ops}.h) over the
C ABI.  Same names, argument meaning and error behaviour as the reference API so that tests read like the reference's own
(icicle/tests/test_curve_api.cpp, test_mod_arithmetic_api.h).  Host data = numpy uint32 limb arrays; device data =
torch CUDA tensors (dtype int32/uint8/...; only the pointer is used) -- torch is plumbing for device memory, not compute.

Nothing here computes: every function forwards to libicicle_b200.so and raises IcicleError on failure.
"""
import copy
import ctypes as C
import enum

import numpy as np

from . import capi
from .capi import IcicleError, MsmConfigC, NttConfigC, VecOpsConfigC, lib, check


class Field(enum.IntEnum):
    BN254_FR = 0
    BN254_FQ = 1
    BLS12_381_FR = 2
    BLS12_381_FQ = 3
    BLS12_377_FR = 4
    BLS12_377_FQ = 5
    BW6_761_FQ = 6
    STARK252 = 7
    BABYBEAR = 8
    KOALABEAR = 9
    M31 = 10
    GOLDILOCKS = 11
    BABYBEAR_EXT4 = 12   # babybear::extension_t (quartic extension): vec-ops; the extension NTT is ntt_extension(Field.BABYBEAR, ...)
    KOALABEAR_EXT4 = 13


FIELD_NAMES = {Field.BN254_FR: "bn254_fr", Field.BN254_FQ: "bn254_fq", Field.BLS12_381_FR: "bls12_381_fr",
               Field.BLS12_381_FQ: "bls12_381_fq", Field.BLS12_377_FR: "bls12_377_fr", Field.BLS12_377_FQ: "bls12_377_fq",
               Field.BW6_761_FQ: "bw6_761_fq", Field.STARK252: "stark252", Field.BABYBEAR: "babybear", Field.KOALABEAR: "koalabear", Field.M31: "m31",
               Field.GOLDILOCKS: "goldilocks"}


class Curve(enum.IntEnum):
    BN254_G1 = 0
    BN254_G2 = 1
    BLS12_381_G1 = 2
    BLS12_381_G2 = 3
    BLS12_377_G1 = 4
    BLS12_377_G2 = 5
    BW6_761_G1 = 6
    BW6_761_G2 = 7
    GRUMPKIN = 8


class NTTDir(enum.IntEnum):  # icicle/include/icicle/ntt.h:23-26
    kForward = 0
    kInverse = 1


class Ordering(enum.IntEnum):  # icicle/include/icicle/ntt.h:37-44
    kNN = 0
    kNR = 1
    kRN = 2
    kRR = 3
    kNM = 4
    kMN = 5


class VecOp(enum.IntEnum):
    ADD = 0
    SUB = 1
    MUL = 2
    ACCUMULATE = 3
    SCALAR_ADD_VEC = 4
    SCALAR_SUB_VEC = 5
    SCALAR_MUL_VEC = 6


def field_limbs(field):
    return lib.b200_field_bytes(int(field)) // 4


def scalar_field(curve):
    return Field(lib.b200_curve_scalar_field(int(curve)))


def affine_limbs(curve):
    return lib.b200_curve_affine_bytes(int(curve)) // 4


def projective_limbs(curve):
    return lib.b200_curve_projective_bytes(int(curve)) // 4


# ---- buffers ----------------------------------------------------------------------------------------------------------
def _is_torch(x):
    return type(x).__module__.startswith("torch")


def is_on_device(x):
    return _is_torch(x) and x.is_cuda


def _ptr(x):
    """(pointer, on_device, keepalive)"""
    if _is_torch(x):
        if not x.is_contiguous():
            raise ValueError("device/host tensors must be contiguous")
        return x.data_ptr(), bool(x.is_cuda), x
    a = np.ascontiguousarray(x)
    return a.ctypes.data, False, a


def _out_ptr(x):
    """(pointer, on_device, keepalive) of an OUTPUT buffer: never copied -- a non-contiguous or non-32-bit array would make the
    kernel write into a temporary and leave the caller's array untouched, so it is rejected instead."""
    if _is_torch(x):
        return _ptr(x)
    if not isinstance(x, np.ndarray) or not x.flags.c_contiguous or not x.flags.writeable or x.dtype.itemsize != 4:
        raise ValueError("output buffers must be writeable C-contiguous numpy arrays of a 32-bit dtype (uint32 limbs)")
    return x.ctypes.data, False, x


def _stream_handle(stream):
    if stream is None:
        return None
    if hasattr(stream, "cuda_stream"):
        return stream.cuda_stream
    return int(stream)


def device_empty(n_limbs_total, device=None):
    import torch
    return torch.empty(int(n_limbs_total), dtype=torch.int32, device=device or "cuda")


def to_device(host_array, device=None, stream=None):
    import torch
    a = np.ascontiguousarray(host_array, dtype=np.uint32)
    t = torch.empty(a.size, dtype=torch.int32, device=device or "cuda")
    check(lib.b200_copy_to_device(t.data_ptr(), a.ctypes.data, a.nbytes, _stream_handle(stream), 0), "copy_to_device")
    return t.view(*a.shape)


def to_host(dev_tensor, shape=None):
    n = dev_tensor.numel() * dev_tensor.element_size() // 4
    out = np.empty(n, dtype=np.uint32)
    check(lib.b200_copy_to_host(out.ctypes.data, dev_tensor.data_ptr(), out.nbytes, None, 0), "copy_to_host")
    return out.reshape(shape if shape is not None else tuple(dev_tensor.shape))


def set_device(device_id):
    check(lib.b200_set_device(int(device_id)), "set_device")


def get_device_count():
    n = C.c_int(0)
    check(lib.b200_get_device_count(C.byref(n)), "get_device_count")
    return n.value


# ---- MSM --------------------------------------------------------------------------------------------------------------
class MSMConfig:
    """icicle::MSMConfig (icicle/include/icicle/msm.h:21-53); defaults of default_msm_config() (msm.h:60-78)."""

    def __init__(self, **kw):
        self.stream = None
        self.precompute_factor = 1
        self.c = 0
        self.bitsize = 0
        self.batch_size = 1
        self.are_points_shared_in_batch = True
        self.are_scalars_on_device = False
        self.are_scalars_montgomery_form = False
        self.are_points_on_device = False
        self.are_points_montgomery_form = False
        self.are_results_on_device = False
        self.is_async = False
        self.ext = {}  # backend extension keys: large_bucket_factor, nof_chunks, is_big_triangle (backend/msm_config.h:10-17)
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError(f"MSMConfig has no field {k}")
            setattr(self, k, v)

    def _c(self):
        c = MsmConfigC()
        lib.b200_msm_default_config(C.byref(c))
        c.stream = _stream_handle(self.stream)
        for k in ("precompute_factor", "c", "bitsize", "batch_size"):
            setattr(c, k, int(getattr(self, k)))
        for k in ("are_points_shared_in_batch", "are_scalars_on_device", "are_scalars_montgomery_form", "are_points_on_device",
                  "are_points_montgomery_form", "are_results_on_device", "is_async"):
            setattr(c, k, 1 if getattr(self, k) else 0)
        c.ext_large_bucket_factor = int(self.ext.get("large_bucket_factor", 0))
        c.ext_nof_chunks = int(self.ext.get("nof_chunks", 0))
        c.ext_is_big_triangle = int(bool(self.ext.get("is_big_triangle", False)))
        return c


def default_msm_config():
    return MSMConfig()


def msm(curve, scalars, bases, msm_size, config=None, results=None):
    """icicle::msm (icicle/include/icicle/msm.h:93-94 -> src/msm.cpp:12-16).  Returns `results`:
    (batch, 3*coord_limbs) uint32 homogeneous projective points in standard form."""
    cfg = copy.copy(config) if config else MSMConfig()
    sp, s_dev, _ks = _ptr(scalars)
    bp, b_dev, _kb = _ptr(bases)
    cfg.are_scalars_on_device = s_dev
    cfg.are_points_on_device = b_dev
    if results is None:
        if cfg.are_results_on_device:
            results = device_empty(cfg.batch_size * projective_limbs(curve)).view(cfg.batch_size, -1)
        else:
            results = np.zeros((cfg.batch_size, projective_limbs(curve)), dtype=np.uint32)
    rp, r_dev, _kr = _out_ptr(results)
    cfg.are_results_on_device = r_dev
    c = cfg._c()
    check(lib.b200_msm(int(curve), sp, bp, int(msm_size), C.byref(c), rp), "msm")
    return results


def msm_precompute_bases(curve, bases, nof_bases, config, output=None):
    """icicle::msm_precompute_bases (msm.h:106-107)."""
    cfg = copy.copy(config)
    bp, b_dev, _kb = _ptr(bases)
    cfg.are_points_on_device = b_dev
    if output is None:
        n = nof_bases * cfg.precompute_factor
        output = (device_empty(n * affine_limbs(curve)).view(n, -1) if cfg.are_results_on_device
                  else np.zeros((n, affine_limbs(curve)), dtype=np.uint32))
    op, o_dev, _ko = _out_ptr(output)
    cfg.are_results_on_device = o_dev
    c = cfg._c()
    check(lib.b200_msm_precompute_bases(int(curve), bp, int(nof_bases), C.byref(c), op), "msm_precompute_bases")
    return output


def ec_sum(curve, points, n, config=None, output=None):
    """Sum of n projective points (multi-GPU partial-result combine; see b200_ec_sum in include/icicle_b200.h)."""
    cfg = config or VecOpsConfig()
    fn, ap, op, c, output = _unary("b200_ec_sum", curve, points, 1, projective_limbs(curve), cfg, output)
    check(fn(int(curve), ap, int(n), C.byref(c), op), "ec_sum")
    return output


def msm_choose_c(curve, msm_size, config=None):
    c = (config or MSMConfig())._c()
    return lib.b200_msm_choose_c(int(curve), int(msm_size), C.byref(c))


def msm_pair_levels(curve, msm_size, c=0, config=None):
    """Number of batched-affine pair levels the MSM schedule will run for this size (planning query)."""
    cfg = (config or MSMConfig(c=c))._c()
    return lib.b200_msm_pair_levels(int(curve), int(msm_size), C.byref(cfg))


# ---- multi-GPU: one host thread per device inside the backend, or one process per GPU with the same shard arithmetic ---
def shard_range(total, parts, index):
    """[lo, hi) of `total` units for part `index` of `parts` (b200_shard_range: sizes differ by at most one, earlier parts
    take the extras) -- the split b200_msm_multi_gpu / b200_ntt_multi_gpu use per device and bench.py uses per rank."""
    b, c = C.c_uint64(0), C.c_uint64(0)
    lib.b200_shard_range(int(total), int(parts), int(index), C.byref(b), C.byref(c))
    return b.value, b.value + c.value


def _device_ids(n_devices, device_ids):
    if device_ids is None:
        return int(n_devices), None
    arr = (C.c_int * len(device_ids))(*[int(d) for d in device_ids])
    return len(device_ids), arr


def msm_multi_gpu(curve, scalars, bases, msm_size, config=None, results=None, n_devices=0, device_ids=None):
    """b200_msm_multi_gpu: host-resident scalars / bases / results sharded over the devices (batch index, or point range for a
    single MSM); what the shim calls when MSMConfig.ext carries "multi_gpu"."""
    cfg = copy.copy(config) if config else MSMConfig()
    sp, s_dev, _ks = _ptr(scalars)
    bp, b_dev, _kb = _ptr(bases)
    if results is None:
        results = np.zeros((cfg.batch_size, projective_limbs(curve)), dtype=np.uint32)
    rp, r_dev, _kr = _out_ptr(results)
    cfg.are_scalars_on_device, cfg.are_points_on_device, cfg.are_results_on_device = s_dev, b_dev, r_dev
    n, ids = _device_ids(n_devices, device_ids)
    c = cfg._c()
    check(lib.b200_msm_multi_gpu(int(curve), sp, bp, int(msm_size), C.byref(c), rp, n, ids), "msm_multi_gpu")
    return results


def ntt_multi_gpu(field, input, size, direction, config=None, output=None, n_devices=0, device_ids=None):
    """b200_ntt_multi_gpu: a row batch of NTTs sharded by batch index over the devices (host-resident data)."""
    cfg = copy.copy(config) if config else NTTConfig()
    ip, i_dev, _ki = _ptr(input)
    if output is None:
        output = np.zeros((size * cfg.batch_size, field_limbs(field)), dtype=np.uint32)
    op, o_dev, _ko = _out_ptr(output)
    cfg.are_inputs_on_device, cfg.are_outputs_on_device = i_dev, o_dev
    n, ids = _device_ids(n_devices, device_ids)
    c = cfg._c()
    check(lib.b200_ntt_multi_gpu(int(field), ip, int(size), int(direction), C.byref(c), op, n, ids), "ntt_multi_gpu")
    return output


# ---- NTT --------------------------------------------------------------------------------------------------------------
class NTTConfig:
    """icicle::NTTConfig<S> (icicle/include/icicle/ntt.h:52-64); defaults of default_ntt_config() (ntt.h:73-86)."""

    def __init__(self, **kw):
        self.stream = None
        self.coset_gen = None  # (limbs,) uint32 standard form, None = one
        self.batch_size = 1
        self.columns_batch = False
        self.ordering = Ordering.kNN
        self.are_inputs_on_device = False
        self.are_outputs_on_device = False
        self.is_async = False
        self.ext = {}  # ntt_algorithm (0 auto / 1 radix2 / 2 mixed radix), fast_twiddles (backend/ntt_config.h:7-18)
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError(f"NTTConfig has no field {k}")
            setattr(self, k, v)

    def _c(self):
        c = NttConfigC()
        lib.b200_ntt_default_config(C.byref(c))
        c.stream = _stream_handle(self.stream)
        self._coset_keep = None
        if self.coset_gen is not None:
            self._coset_keep = np.ascontiguousarray(self.coset_gen, dtype=np.uint32)
            c.coset_gen = self._coset_keep.ctypes.data
        c.batch_size = int(self.batch_size)
        c.columns_batch = 1 if self.columns_batch else 0
        c.are_inputs_on_device = 1 if self.are_inputs_on_device else 0
        c.are_outputs_on_device = 1 if self.are_outputs_on_device else 0
        c.is_async = 1 if self.is_async else 0
        c.ordering = int(self.ordering)
        c.ext_ntt_algorithm = int(self.ext.get("ntt_algorithm", 0))
        c.ext_fast_twiddles = int(bool(self.ext.get("fast_twiddles", False)))
        return c


def default_ntt_config():
    return NTTConfig()


def ntt_init_domain(field, primitive_root, stream=None):
    """icicle::ntt_init_domain (icicle/include/icicle/ntt.h:117 -> src/ntt.cpp:26-30)."""
    r = np.ascontiguousarray(primitive_root, dtype=np.uint32)
    check(lib.b200_ntt_init_domain(int(field), r.ctypes.data, _stream_handle(stream)), "ntt_init_domain")


def ntt_release_domain(field):
    check(lib.b200_ntt_release_domain(int(field)), "ntt_release_domain")


def get_root_of_unity_from_domain(field, logn):
    out = np.zeros(field_limbs(field), dtype=np.uint32)
    check(lib.b200_ntt_get_root_of_unity_from_domain(int(field), int(logn), out.ctypes.data), "get_root_of_unity_from_domain")
    return out


def ntt(field, input, size, direction, config=None, output=None):
    """icicle::ntt (icicle/include/icicle/ntt.h:108 -> src/ntt.cpp:11-15)."""
    cfg = copy.copy(config) if config else NTTConfig()
    ip, i_dev, _ki = _ptr(input)
    cfg.are_inputs_on_device = i_dev
    if output is None:
        n = size * cfg.batch_size * field_limbs(field)
        output = device_empty(n) if cfg.are_outputs_on_device else np.zeros((size * cfg.batch_size, field_limbs(field)), dtype=np.uint32)
    op, o_dev, _ko = _out_ptr(output)
    cfg.are_outputs_on_device = o_dev
    c = cfg._c()
    check(lib.b200_ntt(int(field), ip, int(size), int(direction), C.byref(c), op), "ntt")
    return output


EXTENSION_DEGREE = 4  # quartic extension of BabyBear / KoalaBear (fields/stark_fields/babybear.h:88-93)


def ntt_extension(field, input, size, direction, config=None, output=None):
    """icicle::ntt over extension_t (icicle/include/icicle/ntt.h:108 -> src/ntt.cpp:90-103, `<field>_extension_ntt`): `size`
    quartic extension elements per transform, each 4 base-field coefficients; twiddles, coset generator and domain are the
    scalar field's."""
    cfg = copy.copy(config) if config else NTTConfig()
    ip, i_dev, _ki = _ptr(input)
    cfg.are_inputs_on_device = i_dev
    w = field_limbs(field) * EXTENSION_DEGREE
    if output is None:
        n = size * cfg.batch_size * w
        output = device_empty(n) if cfg.are_outputs_on_device else np.zeros((size * cfg.batch_size, w), dtype=np.uint32)
    op, o_dev, _ko = _out_ptr(output)
    cfg.are_outputs_on_device = o_dev
    c = cfg._c()
    check(lib.b200_ntt_extension(int(field), ip, int(size), int(direction), C.byref(c), op), "ntt_extension")
    return output


def ecntt(curve, input, size, direction, config=None, output=None):
    """icicle::ntt over projective_t (icicle/include/icicle/ntt.h:108 -> src/ecntt.cpp:8-18, `<curve>_ecntt`): NTT of `size` G1
    points (homogeneous projective, standard form) with the scalar field's twiddles; the scalar field's domain must be
    initialised (ntt_init_domain on the curve's scalar field)."""
    cfg = copy.copy(config) if config else NTTConfig()
    ip, i_dev, _ki = _ptr(input)
    cfg.are_inputs_on_device = i_dev
    w = projective_limbs(curve)
    if output is None:
        n = size * cfg.batch_size * w
        output = device_empty(n) if cfg.are_outputs_on_device else np.zeros((size * cfg.batch_size, w), dtype=np.uint32)
    op, o_dev, _ko = _out_ptr(output)
    cfg.are_outputs_on_device = o_dev
    c = cfg._c()
    check(lib.b200_ecntt(int(curve), ip, int(size), int(direction), C.byref(c), op), "ecntt")
    return output


# ---- vec ops ----------------------------------------------------------------------------------------------------------
class VecOpsConfig:
    """icicle::VecOpsConfig (icicle/include/icicle/vec_ops.h:19-44)."""

    def __init__(self, **kw):
        self.stream = None
        self.is_a_on_device = False
        self.is_b_on_device = False
        self.is_result_on_device = False
        self.is_async = False
        self.batch_size = 1
        self.columns_batch = False
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError(f"VecOpsConfig has no field {k}")
            setattr(self, k, v)

    def _c(self):
        c = VecOpsConfigC()
        lib.b200_vec_ops_default_config(C.byref(c))
        c.stream = _stream_handle(self.stream)
        c.is_a_on_device = 1 if self.is_a_on_device else 0
        c.is_b_on_device = 1 if self.is_b_on_device else 0
        c.is_result_on_device = 1 if self.is_result_on_device else 0
        c.is_async = 1 if self.is_async else 0
        c.batch_size = int(self.batch_size)
        c.columns_batch = 1 if self.columns_batch else 0
        return c


def _out_like(field, n_elems, on_device):
    if on_device:
        return device_empty(n_elems * field_limbs(field)).view(n_elems, -1)
    return np.zeros((n_elems, field_limbs(field)), dtype=np.uint32)


def _vec2(field, op, a, b, size, config, output):
    cfg = copy.copy(config) if config else VecOpsConfig()
    ap, a_dev, _ka = _ptr(a)
    bp, b_dev, _kb = _ptr(b)
    cfg.is_a_on_device, cfg.is_b_on_device = a_dev, b_dev
    if op == VecOp.ACCUMULATE:
        ap, a_dev, _ka = _out_ptr(a)  # a is updated in place
        output, op_ptr = a, ap
    else:
        if output is None:
            output = _out_like(field, size * cfg.batch_size, cfg.is_result_on_device)
        op_ptr, o_dev, _ko = _out_ptr(output)
        cfg.is_result_on_device = o_dev
    c = cfg._c()
    check(lib.b200_vec_op(int(field), int(op), ap, bp, int(size), C.byref(c), op_ptr), f"vec_op({VecOp(op).name})")
    return output


def vector_add(field, a, b, size, config=None, output=None):
    return _vec2(field, VecOp.ADD, a, b, size, config, output)


def vector_sub(field, a, b, size, config=None, output=None):
    return _vec2(field, VecOp.SUB, a, b, size, config, output)


def vector_mul(field, a, b, size, config=None, output=None):
    return _vec2(field, VecOp.MUL, a, b, size, config, output)


def vector_accumulate(field, a, b, size, config=None):
    return _vec2(field, VecOp.ACCUMULATE, a, b, size, config, None)


def scalar_add_vec(field, scalar_a, b, size, config=None, output=None):
    return _vec2(field, VecOp.SCALAR_ADD_VEC, scalar_a, b, size, config, output)


def scalar_sub_vec(field, scalar_a, b, size, config=None, output=None):
    return _vec2(field, VecOp.SCALAR_SUB_VEC, scalar_a, b, size, config, output)


def scalar_mul_vec(field, scalar_a, b, size, config=None, output=None):
    return _vec2(field, VecOp.SCALAR_MUL_VEC, scalar_a, b, size, config, output)


def ext_mixed_mul(ext_field, a, b, size, config=None, output=None):
    """extension_vector_mixed_mul (icicle/src/vec_ops.cpp:198-210): a[] in the quartic extension, b[] in the base field."""
    cfg = copy.copy(config) if config else VecOpsConfig()
    ap, a_dev, _ka = _ptr(a)
    bp, b_dev, _kb = _ptr(b)
    cfg.is_a_on_device, cfg.is_b_on_device = a_dev, b_dev
    if output is None:
        output = _out_like(ext_field, size * cfg.batch_size, cfg.is_result_on_device)
    op, o_dev, _ko = _out_ptr(output)
    cfg.is_result_on_device = o_dev
    c = cfg._c()
    check(lib.b200_ext_mixed_mul(int(ext_field), ap, bp, int(size), C.byref(c), op), "ext_mixed_mul")
    return output


def _unary(fn_name, field_or_curve, a, n_out_elems, limbs, config, output, *extra):
    cfg = copy.copy(config) if config else VecOpsConfig()
    ap, a_dev, _ka = _ptr(a)
    cfg.is_a_on_device = a_dev
    if output is None:
        output = (device_empty(n_out_elems * limbs).view(n_out_elems, -1) if cfg.is_result_on_device
                  else np.zeros((n_out_elems, limbs), dtype=np.uint32))
    op, o_dev, _ko = _out_ptr(output)
    cfg.is_result_on_device = o_dev
    c = cfg._c()
    fn = getattr(lib, fn_name)
    return fn, ap, op, c, output


def vector_inv(field, a, size, config=None, output=None):
    cfg = config or VecOpsConfig()
    fn, ap, op, c, output = _unary("b200_vector_inv", field, a, size * cfg.batch_size, field_limbs(field), cfg, output)
    check(fn(int(field), ap, int(size), C.byref(c), op), "vector_inv")
    return output


def vector_div(field, a, b, size, config=None, output=None):
    cfg = copy.copy(config) if config else VecOpsConfig()
    ap, a_dev, _ka = _ptr(a)
    bp, b_dev, _kb = _ptr(b)
    cfg.is_a_on_device, cfg.is_b_on_device = a_dev, b_dev
    if output is None:
        output = _out_like(field, size * cfg.batch_size, cfg.is_result_on_device)
    op, o_dev, _ko = _out_ptr(output)
    cfg.is_result_on_device = o_dev
    c = cfg._c()
    check(lib.b200_vector_div(int(field), ap, bp, int(size), C.byref(c), op), "vector_div")
    return output


def vector_sum(field, a, size, config=None, output=None):
    cfg = config or VecOpsConfig()
    fn, ap, op, c, output = _unary("b200_vector_sum", field, a, cfg.batch_size, field_limbs(field), cfg, output)
    check(fn(int(field), ap, int(size), C.byref(c), op), "vector_sum")
    return output


def vector_product(field, a, size, config=None, output=None):
    cfg = config or VecOpsConfig()
    fn, ap, op, c, output = _unary("b200_vector_product", field, a, cfg.batch_size, field_limbs(field), cfg, output)
    check(fn(int(field), ap, int(size), C.byref(c), op), "vector_product")
    return output


def highest_non_zero_idx(field, a, size, config=None):
    cfg = copy.copy(config) if config else VecOpsConfig()
    ap, a_dev, _ka = _ptr(a)
    cfg.is_a_on_device = a_dev
    cfg.is_result_on_device = False
    out = np.zeros(cfg.batch_size, dtype=np.int64)
    c = cfg._c()
    check(lib.b200_highest_non_zero_idx(int(field), ap, int(size), C.byref(c), out.ctypes.data), "highest_non_zero_idx")
    return out


def poly_eval(field, coeffs, coeffs_size, domain, domain_size, config=None, output=None):
    cfg = copy.copy(config) if config else VecOpsConfig()
    cp, c_dev, _kc = _ptr(coeffs)
    dp, d_dev, _kd = _ptr(domain)
    cfg.is_a_on_device, cfg.is_b_on_device = c_dev, d_dev
    if output is None:
        output = _out_like(field, domain_size * cfg.batch_size, cfg.is_result_on_device)
    op, o_dev, _ko = _out_ptr(output)
    cfg.is_result_on_device = o_dev
    c = cfg._c()
    check(lib.b200_poly_eval(int(field), cp, int(coeffs_size), dp, int(domain_size), C.byref(c), op), "poly_eval")
    return output


def poly_division(field, numerator, numerator_size, denominator, denominator_size, q_size, r_size, config=None):
    """Returns (q, r) host arrays (host inputs) -- device variant through the C ABI directly."""
    cfg = copy.copy(config) if config else VecOpsConfig()
    n_p, n_dev, _kn = _ptr(numerator)
    d_p, d_dev, _kd = _ptr(denominator)
    cfg.is_a_on_device, cfg.is_b_on_device, cfg.is_result_on_device = n_dev, d_dev, False
    q = np.zeros((q_size * cfg.batch_size, field_limbs(field)), dtype=np.uint32)
    r = np.zeros((r_size * cfg.batch_size, field_limbs(field)), dtype=np.uint32)
    c = cfg._c()
    check(lib.b200_poly_division(int(field), n_p, int(numerator_size), d_p, int(denominator_size), C.byref(c), q.ctypes.data, int(q_size),
                                 r.ctypes.data, int(r_size)), "poly_division")
    return q, r


def convert_montgomery(field, a, size, is_into, config=None, output=None):
    cfg = config or VecOpsConfig()
    fn, ap, op, c, output = _unary("b200_convert_montgomery", field, a, size * cfg.batch_size, field_limbs(field), cfg, output)
    check(fn(int(field), ap, int(size), 1 if is_into else 0, C.byref(c), op), "convert_montgomery")
    return output


def bit_reverse(field, a, size, config=None, output=None):
    cfg = config or VecOpsConfig()
    fn, ap, op, c, output = _unary("b200_bit_reverse", field, a, size * cfg.batch_size, field_limbs(field), cfg, output)
    check(fn(int(field), ap, int(size), C.byref(c), op), "bit_reverse")
    return output


def matrix_transpose(field, a, rows, cols, config=None, output=None):
    cfg = config or VecOpsConfig()
    fn, ap, op, c, output = _unary("b200_matrix_transpose", field, a, rows * cols * cfg.batch_size, field_limbs(field), cfg, output)
    check(fn(int(field), ap, int(rows), int(cols), C.byref(c), op), "matrix_transpose")
    return output


def slice(field, a, offset, stride, size_in, size_out, config=None, output=None):
    cfg = config or VecOpsConfig()
    fn, ap, op, c, output = _unary("b200_slice", field, a, size_out * cfg.batch_size, field_limbs(field), cfg, output)
    check(fn(int(field), ap, int(offset), int(stride), int(size_in), int(size_out), C.byref(c), op), "slice")
    return output


def affine_convert_montgomery(curve, a, n, is_into, config=None, output=None):
    cfg = config or VecOpsConfig()
    fn, ap, op, c, output = _unary("b200_affine_convert_montgomery", curve, a, n, affine_limbs(curve), cfg, output)
    check(fn(int(curve), ap, int(n), 1 if is_into else 0, C.byref(c), op), "affine_convert_montgomery")
    return output


def projective_convert_montgomery(curve, a, n, is_into, config=None, output=None):
    cfg = config or VecOpsConfig()
    fn, ap, op, c, output = _unary("b200_projective_convert_montgomery", curve, a, n, projective_limbs(curve), cfg, output)
    check(fn(int(curve), ap, int(n), 1 if is_into else 0, C.byref(c), op), "projective_convert_montgomery")
    return output


# ---- developer knobs / memory -------------------------------------------------------------------------------------------
def set_tuning(name, value):
    """b200_set_tuning: developer / test knob (see include/icicle_b200.h); value None or < 0 restores the built-in policy."""
    check(lib.b200_set_tuning(name.encode(), -1 if value is None else int(value)), f"set_tuning({name})")


def get_tuning(name):
    return int(lib.b200_get_tuning(name.encode()))


def trim_scratch(keep_bytes=0):
    """Return the library's retained scratch (private stream-ordered pool of the current device) to the driver."""
    check(lib.b200_trim_scratch(int(keep_bytes)), "trim_scratch")


# ---- instrumentation ----------------------------------------------------------------------------------------------------
def launch_count():
    """Kernels of libicicle_b200.so launched so far in this process."""
    return int(lib.b200_get_launch_count())


def set_profiling(on):
    lib.b200_set_profiling(1 if on else 0)


def last_profile():
    """(what, [(stage, ms), ...]) of the last call made while profiling was on."""
    names = C.create_string_buffer(2048)
    ms = (C.c_float * 64)()
    k = lib.b200_get_last_profile(names, 2048, ms, 64)
    parts = names.value.decode().split(",")
    return parts[0], [(parts[1 + i], float(ms[i])) for i in range(k)]
