// Quartic extension Fp[x]/(x^4 - nr) of the 4-byte fields (BabyBear: nr = 11, KoalaBear: nr = 3) with the Fp<> interface, so the
// generic vec-op kernels of vec_ops.cu instantiate for it unchanged.
// Layout {c0, c1, c2, c3} and the multiplication / inversion rules follow the reference's QuarticExtensionField
// (icicle/include/icicle/fields/quartic_extension.h:45-53 members, :182-199 product, :246-283 inverse); nonresidue constants:
// fields/stark_fields/babybear.h:84-87, koalabear.h:74-77.  "Montgomery form" is coefficient-wise (quartic_extension.h:78-88).
#pragma once
#include "ff.cuh"

namespace b200 {

template <class P_>
struct Ext4 {
  typedef P_ P;
  typedef Fp<P_> B;
  static_assert(P::N == 1, "quartic extension of a single-limb field");
  static_assert(!P::NONRESIDUE_IS_NEG, "x^4 = +nr");
  static constexpr int N = 4;
  static constexpr int BYTES = 16;
  uint32_t v[4];

  B200_HD B c(int i) const { B r; r.v[0] = v[i]; return r; }
  static B200_HD Ext4 make(const B& a, const B& b, const B& cc, const B& d) { Ext4 r; r.v[0] = a.v[0]; r.v[1] = b.v[0]; r.v[2] = cc.v[0]; r.v[3] = d.v[0]; return r; }
  static B200_HD Ext4 zero() { return make(B::zero(), B::zero(), B::zero(), B::zero()); }
  static B200_HD Ext4 one() { return make(B::one(), B::zero(), B::zero(), B::zero()); }          // Montgomery form of 1
  static B200_HD Ext4 r2() { return make(B::r2(), B::zero(), B::zero(), B::zero()); }
  static B200_HD Ext4 raw_one() { return make(B::raw_one(), B::zero(), B::zero(), B::zero()); }
  B200_HD bool is_zero() const { return (v[0] | v[1] | v[2] | v[3]) == 0; }
  friend B200_HD bool operator==(const Ext4& a, const Ext4& b) { return a.v[0] == b.v[0] && a.v[1] == b.v[1] && a.v[2] == b.v[2] && a.v[3] == b.v[3]; }
  friend B200_HD Ext4 operator+(const Ext4& a, const Ext4& b) { return make(a.c(0) + b.c(0), a.c(1) + b.c(1), a.c(2) + b.c(2), a.c(3) + b.c(3)); }
  friend B200_HD Ext4 operator-(const Ext4& a, const Ext4& b) { return make(a.c(0) - b.c(0), a.c(1) - b.c(1), a.c(2) - b.c(2), a.c(3) - b.c(3)); }

  static B200_HD B mul_nr(const B& x) // x * nr for the small non-residues (3, 11)
  {
    B acc = x;
#pragma unroll
    for (uint32_t i = 1; i < P::NONRESIDUE; i++) acc = acc + x;
    return acc;
  }
  // coefficient products are Montgomery products: (a*b)/R coefficient-wise, exactly like Fp<>::operator*
  friend B200_HD Ext4 operator*(const Ext4& a, const Ext4& b)
  {
    const B a0 = a.c(0), a1 = a.c(1), a2 = a.c(2), a3 = a.c(3), b0 = b.c(0), b1 = b.c(1), b2 = b.c(2), b3 = b.c(3);
    return make(a0 * b0 + mul_nr(a1 * b3 + a2 * b2 + a3 * b1), a0 * b1 + a1 * b0 + mul_nr(a2 * b3 + a3 * b2),
                a0 * b2 + a1 * b1 + a2 * b0 + mul_nr(a3 * b3), a0 * b3 + a1 * b2 + a2 * b1 + a3 * b0);
  }
  B200_HD Ext4 scale(const B& s) const { return make(c(0) * s, c(1) * s, c(2) * s, c(3) * s); }
  B200_HD Ext4 to_mont() const { return scale(B::r2()); }
  B200_HD Ext4 from_mont() const { return scale(B::raw_one()); }
};

// base-field Fermat inverse in the Montgomery domain (0 -> 0)
template <class P>
B200_HD Fp<P> ext4_base_inv(const Fp<P>& a_m)
{
  Fp<P> r = Fp<P>::one();
  const uint32_t e = P::p(0) - 2;
  for (int i = 31; i >= 0; i--) {
    r = r * r;
    if ((e >> i) & 1) r = r * a_m;
  }
  return r;
}

// inverse in the Montgomery domain, reference formula (quartic_extension.h:246-283, non-negative non-residue branch); 0 -> 0
template <class P>
B200_HD Ext4<P> fermat_inv_mont(const Ext4<P>& xs)
{
  typedef Fp<P> B;
  typedef Ext4<P> E;
  const B c0 = xs.c(0), c1 = xs.c(1), c2 = xs.c(2), c3 = xs.c(3);
  B x0 = c0 * c0 - E::mul_nr(c1 * (c3 + c3) - c2 * c2);
  B x2 = c0 * (c2 + c2) - c1 * c1 - E::mul_nr(c3 * c3);
  const B x = x0 * x0 - E::mul_nr(x2 * x2);
  const B xi = ext4_base_inv(x);
  x0 = x0 * xi;
  x2 = x2 * xi;
  return E::make(c0 * x0 - E::mul_nr(c2 * x2), E::mul_nr(c3 * x2) - c1 * x0, c2 * x0 - c0 * x2, c1 * x2 - c3 * x0);
}

} // namespace b200
