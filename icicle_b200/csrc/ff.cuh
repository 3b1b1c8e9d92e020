// Prime-field arithmetic for the sm_100a MSM / NTT / vec-ops kernels.
//
// Replaces (on device) what the reference computes on the host with Barrett reduction:
//   icicle/include/icicle/math/modular_arithmetic.h:354-406 (add/sub/mul), :583-597 (Montgomery conversion, neg)
//   icicle/include/icicle/math/host_math.h:209-238,437-470 (multiply_raw, Barrett)
// The reference's device branch includes headers that are not in its tree (modular_arithmetic.h:4-7), so nothing here comes
// from the reference: values are N little-endian u32 limbs (the reference's `storage<N>` layout, math/storage.h:36-48);
// multiplication is Montgomery (R = 2^(32N), the same R the reference uses for its *_montgomery_form flags,
// fields/params_gen.h:35-50) written as even/odd-column carry chains so that every `mad.lo.cc/madc.hi.cc` pair
// becomes one IMAD.WIDE.U32.X in SASS (B200 has no 64-bit integer multiplier; IMAD.WIDE.U32 is the widest).
// CREDIT: the even/odd-column Montgomery multiplier (mad_n_redc / cmad_n / madc_n_rshift below, their structure and names)
// is the public algorithm of Supranational's sppark (ff/mont_t.cuh, Apache-2.0), restated here for this code base's Fp<>
// type and host emulation; the dedicated squaring (sqr_impl) is derived from it.
//
// All inputs/outputs of add/sub/mul are fully reduced, i.e. in [0, p).  Kernels keep data in the reference's canonical
// standard form at the API boundary and use the identity  mont_mul(x, y*R) = x*y  to avoid conversions where possible.
//
// Every primitive also has a host emulation (carry flag in a thread_local) so tests/ can exercise exactly this code
// on the CPU-only build box.  The host emulation is test-only; no product entry point ever calls it.
#pragma once
#include <cstdint>
#include <type_traits>
#include "params_gen.cuh"

#ifdef __CUDACC__
  #define B200_HD __host__ __device__ __forceinline__
  #define B200_D __device__ __forceinline__
#else
  #define B200_HD inline
  #define B200_D inline
#endif

namespace b200 {

// ------------------------------------------------------------------------------------------------------------------
// carry-chain primitives
// ------------------------------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define B200_ASM asm volatile
B200_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; B200_ASM("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; B200_ASM("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; B200_ASM("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; B200_ASM("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; B200_ASM("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
B200_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; B200_ASM("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
// (lo,hi) = a*b
B200_D void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b)
{
  B200_ASM("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) += a*b ; CF out          (starts a chain)
B200_D void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b)
{
  B200_ASM("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) += a*b + CF ; CF out     (continues a chain)
B200_D void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b)
{
  B200_ASM("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
// (lo,hi) = a*b + (clo,chi) + CF ; CF out
B200_D void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi)
{
  B200_ASM("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
           : "=r"(lo), "=r"(hi)
           : "r"(a), "r"(b), "r"(clo), "r"(chi));
}
// (lo,hi) = a*b + CF ; no CF out   (ends a chain; the caller guarantees no overflow)
B200_D void madc_wide_last(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b)
{
  B200_ASM("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;" : "=r"(lo), "=r"(hi) : "r"(a), "r"(b));
}
#else
// ---- host emulation (tests only) ----
inline uint32_t& emu_cf()
{
  static thread_local uint32_t cf = 0;
  return cf;
}
inline uint32_t emu_add(uint32_t a, uint32_t b, uint32_t cin, bool set)
{
  uint64_t t = (uint64_t)a + b + cin;
  if (set) emu_cf() = (uint32_t)(t >> 32);
  return (uint32_t)t;
}
inline uint32_t emu_sub(uint32_t a, uint32_t b, uint32_t bin, bool set)
{
  uint64_t t = (uint64_t)a - b - bin;
  if (set) emu_cf() = (uint32_t)((t >> 32) & 1); // CF holds the BORROW after sub.cc (PTX: CC.CF = borrow)
  return (uint32_t)t;
}
inline uint32_t add_cc(uint32_t a, uint32_t b) { return emu_add(a, b, 0, true); }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { return emu_add(a, b, emu_cf(), true); }
inline uint32_t addc(uint32_t a, uint32_t b) { return emu_add(a, b, emu_cf(), false); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, 0, true); }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { return emu_sub(a, b, emu_cf(), true); }
inline uint32_t subc(uint32_t a, uint32_t b) { return emu_sub(a, b, emu_cf(), false); }
inline void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b)
{
  uint64_t t = (uint64_t)a * b;
  lo = (uint32_t)t;
  hi = (uint32_t)(t >> 32);
}
inline void emu_mad(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi, uint32_t cin, bool set)
{
  unsigned __int128 t = (unsigned __int128)a * b + (((uint64_t)chi << 32) | clo) + cin;
  lo = (uint32_t)t;
  hi = (uint32_t)(t >> 32);
  if (set) emu_cf() = (uint32_t)(t >> 64);
}
inline void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { emu_mad(lo, hi, a, b, lo, hi, 0, true); }
inline void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { emu_mad(lo, hi, a, b, lo, hi, emu_cf(), true); }
inline void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi)
{
  emu_mad(lo, hi, a, b, clo, chi, emu_cf(), true);
}
inline void madc_wide_last(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { emu_mad(lo, hi, a, b, 0, 0, emu_cf(), false); }
#endif

// ------------------------------------------------------------------------------------------------------------------
// Fp<P>: element of the prime field described by P (one of b200::params::*), N = P::N limbs.
// ------------------------------------------------------------------------------------------------------------------
template <class P_>
struct Fp {
  typedef P_ P;
  static constexpr int N = P::N;
  static constexpr int BYTES = 4 * N;
  uint32_t v[N];

  static B200_HD Fp zero()
  {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = 0;
    return r;
  }
  // Montgomery form of 1
  static B200_HD Fp one()
  {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::r(i);
    return r;
  }
  static B200_HD Fp r2()
  {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::r2(i);
    return r;
  }
  // standard-form 1 (== Montgomery form of R^-1)
  static B200_HD Fp raw_one()
  {
    Fp r = zero();
    r.v[0] = 1;
    return r;
  }
  static B200_HD Fp modulus()
  {
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = P::p(i);
    return r;
  }

  B200_HD bool is_zero() const
  {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < N; i++) t |= v[i];
    return t == 0;
  }
  friend B200_HD bool operator==(const Fp& a, const Fp& b)
  {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < N; i++) t |= a.v[i] ^ b.v[i];
    return t == 0;
  }
  friend B200_HD bool operator!=(const Fp& a, const Fp& b) { return !(a == b); }

  // r = (a >= p) ? a - p : a   for a < 2p (no carry out of the top limb: every supported P has SPARE_BITS >= 1)
  static B200_HD Fp reduce_once(const Fp& a)
  {
    if constexpr (N == 1) {
      Fp r1 = a;
      if (r1.v[0] >= P::p(0)) r1.v[0] -= P::p(0);
      return r1;
    }
    Fp t;
    t.v[0] = sub_cc(a.v[0], P::p(0));
#pragma unroll
    for (int i = 1; i < N; i++) t.v[i] = subc_cc(a.v[i], P::p(i));
    uint32_t borrow_mask = subc(0, 0); // 0xffffffff if a < p
    Fp r;
#pragma unroll
    for (int i = 0; i < N; i++) r.v[i] = (a.v[i] & borrow_mask) | (t.v[i] & ~borrow_mask);
    return r;
  }

  friend B200_HD Fp operator+(const Fp& a, const Fp& b)
  {
    Fp s;
    if constexpr (N == 1) {
      s.v[0] = a.v[0] + b.v[0]; // < 2^32 because p < 2^31
      if (s.v[0] >= P::p(0)) s.v[0] -= P::p(0);
      return s;
    } else {
      s.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
      for (int i = 1; i < N - 1; i++) s.v[i] = addc_cc(a.v[i], b.v[i]);
      s.v[N - 1] = addc(a.v[N - 1], b.v[N - 1]);
      return reduce_once(s);
    }
  }

  friend B200_HD Fp operator-(const Fp& a, const Fp& b)
  {
    Fp d;
    if constexpr (N == 1) {
      d.v[0] = a.v[0] - b.v[0];
      if (a.v[0] < b.v[0]) d.v[0] += P::p(0);
      return d;
    } else {
      d.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
      for (int i = 1; i < N; i++) d.v[i] = subc_cc(a.v[i], b.v[i]);
      uint32_t borrow_mask = subc(0, 0); // 0xffffffff if a < b
      Fp r;
      r.v[0] = add_cc(d.v[0], P::p(0) & borrow_mask);
#pragma unroll
      for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(d.v[i], P::p(i) & borrow_mask);
      r.v[N - 1] = addc(d.v[N - 1], P::p(N - 1) & borrow_mask);
      return r;
    }
  }

  B200_HD Fp neg() const
  {
    if (is_zero()) return *this;
    return modulus_minus(*this);
  }
  // p - a (a in (0,p])
  static B200_HD Fp modulus_minus(const Fp& a)
  {
    Fp r;
    if constexpr (N == 1) {
      r.v[0] = P::p(0) - a.v[0];
      return r;
    } else {
      r.v[0] = sub_cc(P::p(0), a.v[0]);
#pragma unroll
      for (int i = 1; i < N - 1; i++) r.v[i] = subc_cc(P::p(i), a.v[i]);
      r.v[N - 1] = subc(P::p(N - 1), a.v[N - 1]);
      return r;
    }
  }
  B200_HD Fp dbl() const { return *this + *this; }

  // ---- Montgomery multiplication: a*b*R^-1 mod p, fully reduced -------------------------------------------------
  // Even/odd column accumulators: ev[k] holds column k, od[k] holds column k+1.  One outer iteration adds a*b_i and
  // m*p to both, which zeroes column 0; the one-column right shift is then free: the arrays swap roles (old `od` is the
  // new `ev`), and the old `ev` is re-aligned by two limbs inside the next multiply-accumulate chain.
  // Per iteration: 2N IMAD.WIDE.U32(.X) + 1 IMAD (m) + 3 IADD3(.X); total 2N^2 + O(N).
  static B200_HD void madc_n_rshift(uint32_t* od, const uint32_t* a_odd, uint32_t bi)
  {
    // (od'[j], od'[j+1]) = a_odd[j]*bi + (od[j+2], od[j+3]) + CF, j = 0,2,..,N-4 ; top pair takes no addend.
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) madc_wide_cc(od[j], od[j + 1], a_odd[j], bi, od[j + 2], od[j + 3]);
    madc_wide_last(od[N - 2], od[N - 1], a_odd[N - 2], bi);
  }
  static B200_HD void cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi)
  {
    mad_wide_cc(acc[0], acc[1], a[0], bi);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_wide_cc(acc[j], acc[j + 1], a[j], bi);
  }
  static B200_HD void cmad_p_even(uint32_t* acc, uint32_t m)
  {
    mad_wide_cc(acc[0], acc[1], P::p(0), m);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_wide_cc(acc[j], acc[j + 1], P::p(j), m);
  }
  static B200_HD void cmad_p_odd(uint32_t* acc, uint32_t m)
  {
    mad_wide_cc(acc[0], acc[1], P::p(1), m);
#pragma unroll
    for (int j = 2; j < N; j += 2) madc_wide_cc(acc[j], acc[j + 1], P::p(j + 1), m);
  }
  static B200_HD void mad_n_redc(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t bi, bool first)
  {
    if (first) {
#pragma unroll
      for (int j = 0; j < N; j += 2) {
        mul_wide(od[j], od[j + 1], a[j + 1], bi);
        mul_wide(ev[j], ev[j + 1], a[j], bi);
      }
    } else {
      ev[0] = add_cc(ev[0], od[1]);
      madc_n_rshift(od, a + 1, bi);
      cmad_n(ev, a, bi);
      od[N - 1] = addc(od[N - 1], 0);
    }
    uint32_t m = ev[0] * P::NP0;
    cmad_p_odd(od, m);
    cmad_p_even(ev, m);
    od[N - 1] = addc(od[N - 1], 0);
  }

  friend B200_HD Fp operator*(const Fp& a, const Fp& b) { return mont_mul(a, b); }

  static B200_HD Fp mont_mul(const Fp& a, const Fp& b)
  {
    if constexpr (N == 1) {
      uint64_t t = (uint64_t)a.v[0] * b.v[0];
      uint32_t m = (uint32_t)t * P::NP0;
      t += (uint64_t)m * P::p(0); // low word becomes 0; a*b < 2^62 and m*p < 2^63 so no overflow
      Fp r;
      r.v[0] = (uint32_t)(t >> 32);
      if (r.v[0] >= P::p(0)) r.v[0] -= P::p(0);
      return r;
    } else {
      static_assert(N % 2 == 0, "multi-limb fields must have an even limb count");
      uint32_t ev[N], od[N];
#pragma unroll
      for (int i = 0; i < N; i += 2) {
        mad_n_redc(ev, od, a.v, b.v[i], i == 0);
        mad_n_redc(od, ev, a.v, b.v[i + 1], false);
      }
      // the last call used `od` as the even-column array (its column 0 is now zero) and `ev` as the odd-column one:
      // result = (od >> 32) + ev
      Fp r;
      r.v[0] = add_cc(od[1], ev[0]);
#pragma unroll
      for (int i = 1; i < N - 1; i++) r.v[i] = addc_cc(od[i + 1], ev[i]);
      r.v[N - 1] = addc(ev[N - 1], 0);
      return reduce_once(r);
    }
  }
  static B200_HD Fp sqr(const Fp& a) { return mont_mul(a, a); }

  // standard form <-> Montgomery form (reference: to_montgomery / from_montgomery, modular_arithmetic.h:583-585)
  B200_HD Fp to_mont() const { return mont_mul(*this, r2()); }
  B200_HD Fp from_mont() const { return mont_mul(*this, raw_one()); }
};

// 128-bit vectorised global memory access for N % 4 == 0, 32-bit otherwise.  The pointer must be 16-byte aligned: the host
// side guarantees it (stage_in / stage_out in common.cuh copy caller DEVICE buffers that are only 4-byte aligned -- all
// that storage<N> promises, math/storage.h:4-9 -- through aligned scratch; staging areas are cudaMalloc-aligned).
template <class F>
B200_D F load_fp(const uint32_t* p)
{
  F r;
#ifdef __CUDA_ARCH__
  if constexpr (F::N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < F::N; i += 4) {
      uint4 t = *reinterpret_cast<const uint4*>(p + i);
      r.v[i] = t.x; r.v[i + 1] = t.y; r.v[i + 2] = t.z; r.v[i + 3] = t.w;
    }
    return r;
  }
#endif
#pragma unroll
  for (int i = 0; i < F::N; i++) r.v[i] = p[i];
  return r;
}
template <class F>
B200_D void store_fp(uint32_t* p, const F& a)
{
#ifdef __CUDA_ARCH__
  if constexpr (F::N % 4 == 0) {
#pragma unroll
    for (int i = 0; i < F::N; i += 4) *reinterpret_cast<uint4*>(p + i) = make_uint4(a.v[i], a.v[i + 1], a.v[i + 2], a.v[i + 3]);
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < F::N; i++) p[i] = a.v[i];
}

} // namespace b200
