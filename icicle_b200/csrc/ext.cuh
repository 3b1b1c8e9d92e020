// Quadratic extension Fq2 = Fq[u]/(u^2 - nr) with the Fp<> interface, for G2 MSM.
// Layout {c0 = real, c1 = imaginary} and the sign/size of the non-residue follow the reference:
//   icicle/include/icicle/fields/complex_extension.h:42-43 (members), :202-228 (Karatsuba product), :179-189 (nonresidue)
//   nonresidue constants: fields/snark_fields/{bn254,bls12_381,bls12_377}_base.h (fq_config::nonresidue*)
#pragma once
#include "ff.cuh"

namespace b200 {

template <class P_>
struct Fp2 {
  typedef P_ P;
  typedef Fp<P_> B;
  static constexpr int N = 2 * P::N; // limbs per element
  static constexpr int BYTES = 4 * N;
  B c0, c1;

  static B200_HD Fp2 zero() { return {B::zero(), B::zero()}; }
  static B200_HD Fp2 one() { return {B::one(), B::zero()}; }
  B200_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  friend B200_HD bool operator==(const Fp2& a, const Fp2& b) { return a.c0 == b.c0 && a.c1 == b.c1; }
  friend B200_HD bool operator!=(const Fp2& a, const Fp2& b) { return !(a == b); }
  friend B200_HD Fp2 operator+(const Fp2& a, const Fp2& b) { return {a.c0 + b.c0, a.c1 + b.c1}; }
  friend B200_HD Fp2 operator-(const Fp2& a, const Fp2& b) { return {a.c0 - b.c0, a.c1 - b.c1}; }
  B200_HD Fp2 neg() const { return {c0.neg(), c1.neg()}; }
  B200_HD Fp2 dbl() const { return {c0.dbl(), c1.dbl()}; }

  // x * (+-NONRESIDUE) for the small non-residues of the supported towers (1 or 5)
  static B200_HD B mul_nr(const B& x)
  {
    B t = x;
    if constexpr (P::NONRESIDUE == 1) {
      // t = x
    } else if constexpr (P::NONRESIDUE == 5) {
      B x2 = x.dbl();
      t = x2.dbl() + x;
    } else {
      B acc = x;
      for (uint32_t i = 1; i < P::NONRESIDUE; i++) acc = acc + x;
      t = acc;
    }
    return P::NONRESIDUE_IS_NEG ? t.neg() : t;
  }

  friend B200_HD Fp2 operator*(const Fp2& a, const Fp2& b)
  {
    B re = a.c0 * b.c0;
    B im = a.c1 * b.c1;
    B s = (a.c0 + a.c1) * (b.c0 + b.c1);
    return {re + mul_nr(im), s - re - im};
  }
  static B200_HD Fp2 sqr(const Fp2& a)
  {
    // (a0 + a1)(a0 + nr*a1) = a0^2 + nr*a1^2 + (1 + nr) a0 a1
    B ab = a.c0 * a.c1;
    B t = (a.c0 + a.c1) * (a.c0 + mul_nr(a.c1));
    return {t - ab - mul_nr(ab), ab.dbl()};
  }
  B200_HD Fp2 to_mont() const { return {c0.to_mont(), c1.to_mont()}; }
  B200_HD Fp2 from_mont() const { return {c0.from_mont(), c1.from_mont()}; }
};

template <class F>
struct is_ext : std::false_type {};
template <class P>
struct is_ext<Fp2<P>> : std::true_type {};

// generic load/store for Fp and Fp2
template <class P>
B200_D Fp2<P> load_el(const uint32_t* p, Fp2<P>*)
{
  return {load_fp<Fp<P>>(p), load_fp<Fp<P>>(p + P::N)};
}
template <class P>
B200_D Fp<P> load_el(const uint32_t* p, Fp<P>*)
{
  return load_fp<Fp<P>>(p);
}
template <class F>
B200_D F load_el(const uint32_t* p)
{
  return load_el(p, (F*)nullptr);
}
// Gather loads.  Random 32-byte reads are bound by the number of load REQUESTS as much as by sectors on B200: one 256-bit load
// (sm_100+: LDG.E.ENL2.256) per 32-byte coordinate instead of two 128-bit ones raises the measured random-gather rate from
// 30 to 48 G gathers/s (tools/gather_bench.cu, profiles/r1_gather_qualifiers.txt).  `wide` = the base pointer is 32-byte
// aligned (our own staging copies always are; a caller's device buffer is checked on the host).  Read-only data (ld.global.nc).
template <class P>
B200_D Fp<P> load_fp_gather(const uint32_t* p, bool wide)
{
#ifdef __CUDA_ARCH__
  if constexpr (P::N % 8 == 0) {
    if (wide) {
      Fp<P> r;
#pragma unroll
      for (int i = 0; i < P::N; i += 8)
        asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r.v[i]), "=r"(r.v[i + 1]), "=r"(r.v[i + 2]), "=r"(r.v[i + 3]), "=r"(r.v[i + 4]), "=r"(r.v[i + 5]), "=r"(r.v[i + 6]), "=r"(r.v[i + 7])
                     : "l"(p + i));
      return r;
    }
  }
#endif
  return load_fp<Fp<P>>(p);
}
template <class P>
B200_D Fp2<P> load_el_gather(const uint32_t* p, bool wide, Fp2<P>*)
{
  return {load_fp_gather<P>(p, wide), load_fp_gather<P>(p + P::N, wide)};
}
template <class P>
B200_D Fp<P> load_el_gather(const uint32_t* p, bool wide, Fp<P>*)
{
  return load_fp_gather<P>(p, wide);
}
template <class F>
B200_D F load_el_gather(const uint32_t* p, bool wide)
{
  return load_el_gather(p, wide, (F*)nullptr);
}

template <class P>
B200_D void store_el(uint32_t* p, const Fp2<P>& a)
{
  store_fp(p, a.c0);
  store_fp(p + P::N, a.c1);
}
template <class P>
B200_D void store_el(uint32_t* p, const Fp<P>& a)
{
  store_fp(p, a);
}

} // namespace b200
