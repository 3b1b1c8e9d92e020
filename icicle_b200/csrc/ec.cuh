// Short-Weierstrass (a = 0) group law for the MSM kernels.
//
// Boundary types follow the reference exactly:
//   Affine<F>      {x, y}, zero == (0,0)                       icicle/include/icicle/curves/affine.h:11-39
//   Projective<F>  homogeneous {X:Y:Z}, zero == (0,1,0)        icicle/include/icicle/curves/projective.h:23-31
// Internally buckets are kept in extended-Jacobian XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity is
// ZZ == 0) because the mixed add costs 8M+2S against 11M+2*(x3b) for the reference's complete projective formulas
// (projective.h:147-188).  Results are converted to the reference's homogeneous representative at the very end; the
// reference compares points by cross-multiplication (projective.h:228-231) so the representative itself is free.
// All coordinates are in Montgomery form inside the kernels.
// F is a field type with the Fp<> interface (ff.cuh); Fp2<> (ext.cuh) satisfies it for G2.
#pragma once
#include "ff.cuh"

namespace b200 {

template <class F>
struct Affine {
  F x, y;
  B200_HD bool is_zero() const { return x.is_zero() && y.is_zero(); }
  static B200_HD Affine zero() { return {F::zero(), F::zero()}; }
  B200_HD Affine neg() const { return {x, y.neg()}; }
};

template <class F>
struct Projective {
  F x, y, z;
  static B200_HD Projective zero() { return {F::zero(), F::one(), F::zero()}; } // (0,1,0) with 1 in Montgomery form
};

template <class F>
struct XYZZ {
  F x, y, zz, zzz;

  static B200_HD XYZZ inf() { return {F::zero(), F::zero(), F::zero(), F::zero()}; }
  B200_HD bool is_inf() const { return zz.is_zero(); }
  static B200_HD XYZZ from_affine(const Affine<F>& p)
  {
    if (p.is_zero()) return inf();
    return {p.x, p.y, F::one(), F::one()};
  }
  B200_HD XYZZ neg() const { return {x, y.neg(), zz, zzz}; }

  // 2*(x,y) for an affine, non-zero point  (EFD "mdbl-2008-s-1", a = 0): 3M + 3S... written as 2S + 4M + adds
  static B200_HD XYZZ dbl_affine(const Affine<F>& p)
  {
    if (p.y.is_zero()) return inf(); // order-2 point (cannot happen on the prime-order curves we target)
    F U = p.y.dbl();
    F V = F::sqr(U);
    F W = U * V;
    F S = p.x * V;
    F X2 = F::sqr(p.x);
    F M = X2.dbl() + X2;
    F X3 = F::sqr(M) - S.dbl();
    F Y3 = M * (S - X3) - W * p.y;
    return {X3, Y3, V, W};
  }

  // 2*this  (EFD "dbl-2008-s-1", a = 0)
  B200_HD XYZZ dbl_impl() const
  {
    if (is_inf() || y.is_zero()) return inf();
    F U = y.dbl();
    F V = F::sqr(U);
    F W = U * V;
    F S = x * V;
    F X2 = F::sqr(x);
    F M = X2.dbl() + X2;
    F X3 = F::sqr(M) - S.dbl();
    F Y3 = M * (S - X3) - W * y;
    return {X3, Y3, V * zz, W * zzz};
  }

  // this += p (affine, Montgomery coordinates).  EFD "madd-2008-s": 8M + 2S.
  // Replaces the reference's mixed add in the bucket hot loop (cpu_msm.hpp:296-304 -> projective.h:147-188).
  B200_HD void add_affine_impl(const Affine<F>& p)
  {
    if (p.is_zero()) return; // reference skips zero bases: cpu_msm.hpp:282
    if (is_inf()) {
      x = p.x; y = p.y; zz = F::one(); zzz = F::one();
      return;
    }
    F U2 = p.x * zz;
    F S2 = p.y * zzz;
    F P = U2 - x;
    F R = S2 - y;
    if (P.is_zero()) {
      if (R.is_zero()) *this = dbl_affine(p);
      else *this = inf();
      return;
    }
    F PP = F::sqr(P);
    F PPP = P * PP;
    F Q = x * PP;
    F X3 = F::sqr(R) - PPP - Q.dbl();
    F Y3 = R * (Q - X3) - y * PPP;
    x = X3; y = Y3;
    zz = zz * PP;
    zzz = zzz * PPP;
  }

  // this += o.  EFD "add-2008-s": 12M + 2S.
  B200_HD void add_impl(const XYZZ& o)
  {
    if (o.is_inf()) return;
    if (is_inf()) { *this = o; return; }
    F U1 = x * o.zz;
    F U2 = o.x * zz;
    F S1 = y * o.zzz;
    F S2 = o.y * zzz;
    F P = U2 - U1;
    F R = S2 - S1;
    if (P.is_zero()) {
      if (R.is_zero()) *this = dbl_impl();
      else *this = inf();
      return;
    }
    F PP = F::sqr(P);
    F PPP = P * PP;
    F Q = U1 * PP;
    F X3 = F::sqr(R) - PPP - Q.dbl();
    F Y3 = R * (Q - X3) - S1 * PPP;
    x = X3; y = Y3;
    zz = zz * o.zz * PP;
    zzz = zzz * o.zzz * PPP;
  }


  // For fields wider than 256 bits (and for Fq2) one group operation is 5-15 thousand SASS instructions; inlining it at
  // every call site makes kernels that ptxas needs tens of minutes for and that thrash the instruction cache.  Those
  // instantiations call the operation out of line (the operands then live in local memory across the call, a few hundred
  // bytes against thousands of multiply instructions); the 256-bit G1 hot path stays fully inlined.
  static constexpr bool kOutOfLine = (F::BYTES > 32);
#ifdef __CUDACC__
  __device__ __noinline__ void add_affine_ool(const Affine<F>& p) { add_affine_impl(p); }
  __device__ __noinline__ void add_ool(const XYZZ& o) { add_impl(o); }
  __device__ __noinline__ XYZZ dbl_ool() const { return dbl_impl(); }
#endif
  B200_HD void add_affine(const Affine<F>& p)
  {
#ifdef __CUDA_ARCH__
    if constexpr (kOutOfLine) { add_affine_ool(p); return; }
#endif
    add_affine_impl(p);
  }
  B200_HD void add(const XYZZ& o)
  {
#ifdef __CUDA_ARCH__
    if constexpr (kOutOfLine) { add_ool(o); return; }
#endif
    add_impl(o);
  }
  B200_HD XYZZ dbl() const
  {
#ifdef __CUDA_ARCH__
    if constexpr (kOutOfLine) return dbl_ool();
#endif
    return dbl_impl();
  }

  // homogeneous projective representative (X*ZZZ : Y*ZZ : ZZ*ZZZ), Montgomery coordinates; infinity -> (0,1,0)
  B200_HD Projective<F> to_projective() const
  {
    if (is_inf()) return Projective<F>::zero();
    return {x * zzz, y * zz, zz * zzz};
  }
  static B200_HD XYZZ from_projective(const Projective<F>& p)
  {
    if (p.z.is_zero()) return inf();
    // x = X/Z = X*Z/Z^2, y = Y/Z = Y*Z^2/Z^3
    F z2 = F::sqr(p.z);
    return {p.x * p.z, p.y * z2, z2, z2 * p.z};
  }
};

} // namespace b200
