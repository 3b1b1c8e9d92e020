// TEST-ONLY host emulation harness: compiles icicle_b200/csrc/ff.cuh with g++ (carry flag emulated) and checks the
// Montgomery multiply / add / sub against __int128-free schoolbook big-int arithmetic done here with Python-provided
// vectors on stdin:  lines "field a b" (hex) -> prints "a+b a-b a*b(mont) to_mont(a) from_mont(a)" in hex.
#include <cstdio>
#include <cstring>
#include <string>
#include <iostream>
#define __host__
#define __device__
#include "../../icicle_b200/csrc/ff.cuh"
#include "../../icicle_b200/csrc/goldilocks.cuh"
#include "../../icicle_b200/csrc/ext4.cuh"
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
template <class F> static void print(const F& a)
{
  for (int i = F::N - 1; i >= 0; i--) printf("%08x", a.v[i]);
  printf(" ");
}
template <class F> static void run(const std::string& sa, const std::string& sb)
{
  F a = parse<F>(sa), b = parse<F>(sb);
  print(a + b); print(a - b); print(a * b); print(a.to_mont()); print(a.from_mont()); print(a.neg());
  printf("\n");
}
// quartic extension: a+b, a-b, a*b (coefficient products are Montgomery products: true product / R), to_mont, from_mont, and the
// true inverse computed in the Montgomery domain (fermat_inv_mont) -- 0 -> 0
template <class E> static void run_ext(const std::string& sa, const std::string& sb)
{
  E a = parse<E>(sa), b = parse<E>(sb);
  print(a + b); print(a - b); print(a * b); print(a.to_mont()); print(a.from_mont()); print(fermat_inv_mont(a.to_mont()).from_mont());
  printf("\n");
}
int main()
{
  std::string f, a, b;
  while (std::cin >> f >> a >> b) {
    if (f == "bn254_fr") run<Fp<params::bn254_fr>>(a, b);
    else if (f == "bn254_fq") run<Fp<params::bn254_fq>>(a, b);
    else if (f == "bls12_381_fr") run<Fp<params::bls12_381_fr>>(a, b);
    else if (f == "bls12_381_fq") run<Fp<params::bls12_381_fq>>(a, b);
    else if (f == "bls12_377_fr") run<Fp<params::bls12_377_fr>>(a, b);
    else if (f == "bls12_377_fq") run<Fp<params::bls12_377_fq>>(a, b);
    else if (f == "bw6_761_fq") run<Fp<params::bw6_761_fq>>(a, b);
    else if (f == "stark252") run<Fp<params::stark252>>(a, b);
    else if (f == "babybear") run<Fp<params::babybear>>(a, b);
    else if (f == "koalabear") run<Fp<params::koalabear>>(a, b);
    else if (f == "m31") run<Fp<params::m31>>(a, b);
    else if (f == "goldilocks") run<Fp<params::goldilocks>>(a, b);
    else if (f == "ext4_babybear") run_ext<Ext4<params::babybear>>(a, b);
    else if (f == "ext4_koalabear") run_ext<Ext4<params::koalabear>>(a, b);
    else { printf("unknown\n"); }
  }
  return 0;
}
