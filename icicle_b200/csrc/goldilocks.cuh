// Goldilocks, p = 2^64 - 2^32 + 1: the reference's special-form 2-limb field (icicle/include/icicle/fields/stark_fields/
// goldilocks.h:278 modulus, :328-329 rou / omegas_count; arithmetic notes :13-19: no slack bit -- additions carry out of the
// 64-bit word -- and a reduction that uses 2^64 = 2^32 - 1, 2^96 = -1 (math/goldilocks_host_math.h)).
// Fp<params::goldilocks> is a FULL specialisation with the Fp<> interface.  Internally there is no Montgomery domain: the
// product is the plain modular product (one 64x64 multiply + the special reduction, cheaper than a Montgomery step), so
// one() == raw_one() == r2() == 1 and to_mont()/from_mont() are the identity; the API-level Montgomery conversion (R = 2^64,
// fields/params_gen.h:35-50) is a multiplication by a constant in b200_convert_montgomery.
#pragma once
#include "ff.cuh"

namespace b200 {
namespace params {
struct goldilocks {
  static constexpr int N = 2;
  static constexpr int BITS = 64;
  static constexpr int TWO_ADICITY = 32;
  static constexpr int SPARE_BITS = 0;
  static constexpr uint32_t NP0 = 0xffffffffu; // unused (no Montgomery reduction)
  static __host__ __device__ constexpr uint32_t p(int i) { constexpr uint32_t a[2] = {0x00000001u, 0xffffffffu}; return a[i]; }
  static __host__ __device__ constexpr uint32_t rou(int i) { constexpr uint32_t a[2] = {0xda58878cu, 0x185629dcu}; return a[i]; } // goldilocks.h:328
  static constexpr bool HAS_ROU = true;
  static constexpr uint64_t MONT_R = 0x00000000ffffffffull;     // 2^64 mod p
  static constexpr uint64_t MONT_R_INV = 0xfffffffe00000001ull; // 2^-64 mod p = 2^128 mod p
};
} // namespace params

template <>
struct Fp<params::goldilocks> {
  typedef params::goldilocks P;
  static constexpr int N = 2;
  static constexpr int BYTES = 8;
  static constexpr uint64_t MOD = 0xffffffff00000001ull;
  static constexpr uint64_t EPS = 0x00000000ffffffffull; // 2^64 mod p
  uint32_t v[2];

  B200_HD uint64_t u64() const { return ((uint64_t)v[1] << 32) | v[0]; }
  static B200_HD Fp from_u64(uint64_t x) { Fp r; r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32); return r; }
  static B200_HD Fp zero() { return from_u64(0); }
  static B200_HD Fp one() { return from_u64(1); }
  static B200_HD Fp r2() { return from_u64(1); }
  static B200_HD Fp raw_one() { return from_u64(1); }
  static B200_HD Fp modulus() { return from_u64(MOD); }
  B200_HD bool is_zero() const { return (v[0] | v[1]) == 0; }
  friend B200_HD bool operator==(const Fp& a, const Fp& b) { return a.v[0] == b.v[0] && a.v[1] == b.v[1]; }
  friend B200_HD bool operator!=(const Fp& a, const Fp& b) { return !(a == b); }

  friend B200_HD Fp operator+(const Fp& a, const Fp& b)
  {
    const uint64_t x = a.u64(), y = b.u64();
    uint64_t s = x + y;
    if (s < x || s >= MOD) s -= MOD; // carry out of 64 bits (s + 2^64 - p = s + EPS wraps to the same value) or s in [p, 2^64)
    return from_u64(s);
  }
  friend B200_HD Fp operator-(const Fp& a, const Fp& b)
  {
    const uint64_t x = a.u64(), y = b.u64();
    uint64_t d = x - y;
    if (x < y) d += MOD;
    return from_u64(d);
  }
  B200_HD Fp neg() const { return is_zero() ? *this : from_u64(MOD - u64()); }
  B200_HD Fp dbl() const { return *this + *this; }

  // x = hi * 2^64 + lo  ->  x mod p, using 2^64 = 2^32 - 1 and 2^96 = -1 (mod p)
  static B200_HD uint64_t reduce128(uint64_t hi, uint64_t lo)
  {
    const uint64_t hi_hi = hi >> 32, hi_lo = hi & EPS;
    uint64_t t = lo - hi_hi;
    if (lo < hi_hi) t -= EPS;                 // borrow: + p = - EPS (mod 2^64)
    const uint64_t m = hi_lo * EPS;           // < 2^64
    uint64_t r = t + m;
    if (r < t) r += EPS;                      // carry: - p = + EPS (mod 2^64); cannot carry again
    if (r >= MOD) r -= MOD;
    return r;
  }
  static B200_HD Fp mont_mul(const Fp& a, const Fp& b)
  {
    const uint64_t x = a.u64(), y = b.u64();
#ifdef __CUDA_ARCH__
    const uint64_t lo = x * y, hi = __umul64hi(x, y);
#else
    const unsigned __int128 w = (unsigned __int128)x * y;
    const uint64_t lo = (uint64_t)w, hi = (uint64_t)(w >> 64);
#endif
    return from_u64(reduce128(hi, lo));
  }
  friend B200_HD Fp operator*(const Fp& a, const Fp& b) { return mont_mul(a, b); }
  static B200_HD Fp sqr(const Fp& a) { return mont_mul(a, a); }
  B200_HD Fp to_mont() const { return *this; }
  B200_HD Fp from_mont() const { return *this; }
};

} // namespace b200
