// TEST-ONLY host emulation harness: compiles icicle_b200/csrc/ec.cuh (+ ff.cuh, ext.cuh) with g++ (carry flag emulated) and
// runs the exact group-law code of the MSM / ECNTT kernels -- XYZZ mixed add, full add, doubling, projective round trip -- on
// affine points given in hex on stdin; the driver (check_ec.py) compares with Python-integer curve arithmetic.
//   line:  "<curve> x1 y1 x2 y2 z"   (standard form; (0,0) = the affine zero; z = a non-zero scaling for the projective input)
//   prints 5 results as "X Y Z" (homogeneous projective, standard form):
//     P1 (+) P2 with the mixed add; P1 + P2 with the full add on rescaled-through-projective operands; 2*P1;
//     (P1 + P2) + P2 (full add chained on a general XYZZ accumulator); P1 + (-P1)
#include <cstdio>
#include <string>
#include <iostream>
#define __host__
#define __device__
#include "../../icicle_b200/csrc/ext.cuh"
#include "../../icicle_b200/csrc/ec.cuh"
using namespace b200;

template <class F> static F parse(const std::string& h)
{
  F r = F::zero();
  int n = (int)h.size();
  for (int i = 0; i < n; i++) {
    char c = h[n - 1 - i];
    uint32_t d = (c >= '0' && c <= '9') ? c - '0' : (c - 'a' + 10);
    if (i / 8 < F::N) r.v[i / 8] |= d << (4 * (i % 8));
  }
  return r;
}
template <class P> static void print_el(const Fp<P>& a)
{
  Fp<P> s = a.from_mont();
  for (int i = P::N - 1; i >= 0; i--) printf("%08x", s.v[i]);
  printf(" ");
}
template <class P> static void print_el(const Fp2<P>& a) // "re im" as two words
{
  print_el(a.c0);
  print_el(a.c1);
}
template <class P> static Fp2<P> parse2(const std::string& re, const std::string& im) { return {parse<Fp<P>>(re), parse<Fp<P>>(im)}; }
template <class F> static void print_pt(const XYZZ<F>& q)
{
  Projective<F> p = q.to_projective();
  print_el(p.x); print_el(p.y); print_el(p.z);
}
template <class F> static void run_pts(Affine<F> p1, Affine<F> p2, F z);
template <class F> static void run(const std::string& x1, const std::string& y1, const std::string& x2, const std::string& y2, const std::string& zs)
{
  run_pts<F>({parse<F>(x1).to_mont(), parse<F>(y1).to_mont()}, {parse<F>(x2).to_mont(), parse<F>(y2).to_mont()}, parse<F>(zs).to_mont());
}
template <class F> static void run_pts(Affine<F> p1, Affine<F> p2, F z)
{
  // mixed add
  XYZZ<F> a = XYZZ<F>::from_affine(p1);
  a.add_affine(p2);
  print_pt(a);
  // full add of two general representatives: route both operands through projective (x*z : y*z : z)
  Projective<F> j1 = p1.is_zero() ? Projective<F>::zero() : Projective<F>{p1.x * z, p1.y * z, z};
  Projective<F> j2 = p2.is_zero() ? Projective<F>::zero() : Projective<F>{p2.x * z * z, p2.y * z * z, z * z};
  XYZZ<F> b = XYZZ<F>::from_projective(j1);
  b.add(XYZZ<F>::from_projective(j2));
  print_pt(b);
  // doubling of a general representative
  print_pt(XYZZ<F>::from_projective(j1).dbl());
  // chained
  XYZZ<F> c = b;
  c.add(XYZZ<F>::from_projective(j2));
  print_pt(c);
  // P1 + (-P1)
  XYZZ<F> d = XYZZ<F>::from_projective(j1);
  d.add(XYZZ<F>::from_affine(p1).neg());
  print_pt(d);
  printf("\n");
}
int main()
{
  std::string c, x1, y1, x2, y2, z;
  while (std::cin >> c) {
    if (c == "bn254_g2" || c == "bls12_381_g2") { // G2 over Fq2: every coordinate is "re im"
      std::string w[10];
      for (auto& t : w) std::cin >> t;
      if (c == "bn254_g2") {
        typedef params::bn254_fq P;
        run_pts<Fp2<P>>({parse2<P>(w[0], w[1]).to_mont(), parse2<P>(w[2], w[3]).to_mont()}, {parse2<P>(w[4], w[5]).to_mont(), parse2<P>(w[6], w[7]).to_mont()},
                        parse2<P>(w[8], w[9]).to_mont());
      } else {
        typedef params::bls12_381_fq P;
        run_pts<Fp2<P>>({parse2<P>(w[0], w[1]).to_mont(), parse2<P>(w[2], w[3]).to_mont()}, {parse2<P>(w[4], w[5]).to_mont(), parse2<P>(w[6], w[7]).to_mont()},
                        parse2<P>(w[8], w[9]).to_mont());
      }
      continue;
    }
    std::cin >> x1 >> y1 >> x2 >> y2 >> z;
    if (c == "bn254") run<Fp<params::bn254_fq>>(x1, y1, x2, y2, z);
    else if (c == "bls12_381") run<Fp<params::bls12_381_fq>>(x1, y1, x2, y2, z);
    else if (c == "grumpkin") run<Fp<params::bn254_fr>>(x1, y1, x2, y2, z);
    else printf("unknown\n");
  }
  return 0;
}
