// Batched-affine pair levels for the MSM bucket accumulation (sm_100a).
//
// The bucket accumulation of the reference (worker_run_phase1, icicle/backend/cpu/src/curve/cpu_msm.hpp:259-314) adds every
// point of a bucket into one running projective sum.  A mixed XYZZ add costs 8M+2S; an AFFINE add costs 1 inversion + 2M+1S,
// and Montgomery's trick turns n inversions into 1 inversion + 3(n-1) products.  The GPU schedule here therefore first
// halves the bucket-sorted entry list a few times with affine adds whose inversions are shared across the WHOLE GRID:
//
//   level l list : affine points sorted by bucket, with offsets off_l[k] (level 0 = the radix-sorted (key, point-index|sign)
//                  entries, points gathered from the caller's bases)
//   k_pair_prefix: pair slot q covers entries 2q, 2q+1 (a real pair when both lie in one bucket run); thread t of a block
//                  visits slots base + j*128 + t so all streams are coalesced; it forms the denominators (x2-x1, or 2y for
//                  a doubling) and stores the exclusive running product of its own slots; 1 product / pair
//   k_inv_up/top/down: batch inversion of the per-thread totals by a 64-ary product tree (3 products per thread, one
//                  Fermat inversion at the root)
//   k_pair_apply : visits the same slots backwards, peels each pair's inverse off the thread's inverted total (2 products)
//                  and finishes the add (lambda, x3, y3: 2M+1S); unpaired entries are copied through; 5 products / pair
//
// so one pair add costs ~6 products instead of 10, at the price of ~360 B of HBM traffic per pair (still far below the
// HBM roof).  After L levels the list is 2^L times shorter and goes through the XYZZ slice accumulation (k_accumulate,
// direct mode).  Results are the same group elements as the XYZZ-only path; doubling (equal points in one bucket, which the
// reference's own test generator produces all the time: curves/projective.h:37-53), P + (-P) and affine zero (0,0) inputs
// (skipped by the reference, cpu_msm.hpp:282) are handled explicitly.
#pragma once
#include "common.cuh"

namespace b200 { namespace msm {

constexpr uint32_t PAIR_SIGN_BIT = 0x80000000u;

template <class F>
struct base_fp {
  typedef F type;
};
template <class P>
struct base_fp<Fp2<P>> {
  typedef Fp<P> type;
};

// Fermat inversion in the base field / its quadratic extension (set-up / precompute only)
template <class P>
__device__ Fp<P> inv_fp(const Fp<P>& a)
{
  typedef Fp<P> B;
  uint32_t e[B::N];
#pragma unroll
  for (int i = 0; i < B::N; i++) e[i] = P::p(i);
  { // e = p - 2 with borrow propagation (several moduli end in ...00000001)
    uint32_t borrow = 2;
    for (int i = 0; i < (int)(sizeof(e) / sizeof(e[0])) && borrow; i++) {
      uint32_t before = e[i];
      e[i] = before - borrow;
      borrow = (before < borrow) ? 1u : 0u;
    }
  }
  B r = B::one();
  for (int i = B::N * 32 - 1; i >= 0; i--) {
    r = r * r;
    if ((e[i / 32] >> (i % 32)) & 1) r = r * a;
  }
  return r;
}
template <class P>
__device__ Fp<P> inv_el(const Fp<P>& a)
{
  return inv_fp(a);
}
template <class P>
__device__ Fp2<P> inv_el(const Fp2<P>& a)
{
  // 1/(a0 + a1 u) = (a0 - a1 u) / (a0^2 - nr a1^2)
  typedef Fp<P> B;
  B n = B::sqr(a.c0) - Fp2<P>::mul_nr(B::sqr(a.c1));
  B ni = inv_fp(n);
  return {a.c0 * ni, (a.c1 * ni).neg()};
}


// out-of-line product for wide fields (keeps ptxas time and code size bounded, see ec.cuh)
template <class F>
__device__ __noinline__ F fmul_ool(const F& a, const F& b) { return a * b; }
template <class F>
B200_D F fmul(const F& a, const F& b)
{
  if constexpr (F::BYTES > 32) return fmul_ool<F>(a, b);
  else return a * b;
}

// limb access that works for Fp<> and Fp2<> alike (the index is a compile-time constant after unrolling)
template <class P>
B200_D uint32_t& limb(Fp<P>& a, int i) { return a.v[i]; }
template <class P>
B200_D uint32_t limb(const Fp<P>& a, int i) { return a.v[i]; }
template <class P>
B200_D uint32_t& limb(Fp2<P>& a, int i) { return i < P::N ? a.c0.v[i] : a.c1.v[i - P::N]; }
template <class P>
B200_D uint32_t limb(const Fp2<P>& a, int i) { return i < P::N ? a.c0.v[i] : a.c1.v[i - P::N]; }

// Planar (structure-of-arrays) storage of field elements: 16-byte group g of element `idx` lives at base[g*stride + idx],
// so consecutive lanes touching consecutive elements issue fully coalesced 128-bit accesses.  Used for the private
// scratch of the pair levels (running products, intermediate level points); the API-facing arrays stay AoS.
template <class F>
B200_D F load_planar(const uint4* __restrict__ base, uint64_t stride, uint64_t idx)
{
  static_assert(F::N % 4 == 0, "planar storage needs whole 16-byte groups");
  F r;
#pragma unroll
  for (int g = 0; g < F::N / 4; g++) {
    const uint4 t = base[(uint64_t)g * stride + idx];
    limb(r, 4 * g) = t.x; limb(r, 4 * g + 1) = t.y; limb(r, 4 * g + 2) = t.z; limb(r, 4 * g + 3) = t.w;
  }
  return r;
}
template <class F>
B200_D void store_planar(uint4* __restrict__ base, uint64_t stride, uint64_t idx, const F& a)
{
#pragma unroll
  for (int g = 0; g < F::N / 4; g++)
    base[(uint64_t)g * stride + idx] = make_uint4(limb(a, 4 * g), limb(a, 4 * g + 1), limb(a, 4 * g + 2), limb(a, 4 * g + 3));
}

enum PairKind : int { PK_ADD = 0, PK_DOUBLE = 1, PK_FIRST = 2, PK_SECOND = 3, PK_ZERO = 4 };

// Denominator of the affine sum p1 + p2 and what kind of sum it is.  Points are (x, y) with the sign already applied.
template <class F>
B200_D F pair_denominator(const F& x1, const F& y1, const F& x2, const F& y2, int& kind)
{
  const bool z1 = x1.is_zero() && y1.is_zero();
  const bool z2 = x2.is_zero() && y2.is_zero();
  if (z1) { kind = PK_SECOND; return F::one(); }
  if (z2) { kind = PK_FIRST; return F::one(); }
  if (x1 == x2) {
    if (y1 == y2 && !y1.is_zero()) { kind = PK_DOUBLE; return y1.dbl(); }
    kind = PK_ZERO;
    return F::one();
  }
  kind = PK_ADD;
  return x2 - x1;
}

// ---------------------------------------------------------------------------------------------------------------------
// off[k] = first entry with key >= k, k = 0..nb (off[nb] = number of non-sentinel entries).  Thread per bucket.
// ---------------------------------------------------------------------------------------------------------------------
static __global__ void __launch_bounds__(256) k_bounds(const uint32_t* __restrict__ keys, uint32_t n_ent, uint32_t nb, uint32_t* __restrict__ off)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nb) return;
  uint32_t lo = 0, hi = n_ent;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (keys[mid] < k) lo = mid + 1; else hi = mid;
  }
  off[k] = lo;
}

// Pairing is by ABSOLUTE position: pair slot q covers entries 2q and 2q+1 of the level list and is a real pair when both
// belong to the same bucket run; otherwise the two entries pass through on their own.  A run [rb, re) therefore shrinks to
//   (rb odd: its first entry is alone) + (whole slots inside) + (re odd: its last entry is alone)
// entries, laid out in that order.  cnt[k] is that length (cnt[nb] = 0); its exclusive scan is the next offset table.
static __global__ void __launch_bounds__(256) k_pair_counts(const uint32_t* __restrict__ off, uint32_t nb, uint32_t* __restrict__ cnt)
{
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nb) return;
  uint32_t c = 0;
  if (k < nb) {
    const uint32_t rb = off[k], re = off[k + 1];
    if (re > rb) {
      const uint32_t a = rb + (rb & 1u), b = re - (re & 1u);
      c = ((b - a) >> 1) + (rb & 1u) + (re & 1u);
    }
  }
  cnt[k] = c;
}

// One level's input.  GATHER (level 0): entry e is (keys[e], vals[e] = point index | sign), points AoS in `pts`.
// Otherwise entry e is the e-th point of a planar level buffer (x groups, then y groups; plane stride `cap`).
template <class F, bool GATHER>
struct PairSrc {
  static constexpr int N = F::N, AW = 2 * F::N;
  const uint32_t* keys;
  const uint32_t* vals;
  const uint32_t* pts;
  uint64_t cap;
  bool wide; // GATHER: pts is 32-byte aligned -> 256-bit gather loads (ext.cuh load_el_gather)
  B200_D void load_x(uint32_t e, F& x, uint32_t& ref) const
  {
    if constexpr (GATHER) {
      ref = vals[e];
      x = load_el<F>(pts + (uint64_t)(ref & ~PAIR_SIGN_BIT) * AW);
    } else {
      ref = e;
      x = load_planar<F>(reinterpret_cast<const uint4*>(pts), cap, e);
    }
  }
  // --- split accessors for the software pipeline: the (key, value) words are prefetched two slots ahead, the coordinates one
  // slot ahead, and nothing touches a loaded register before the slot is consumed (a sign flip inside the fetch would make
  // the warp wait for the gather right there: ncu, profiles/r1_ncu_pair_kernels_2p26.txt)
  B200_D uint32_t ref_of(uint32_t e, uint32_t v) const { return GATHER ? v : e; }
  B200_D F load_x_raw(uint32_t ref) const
  {
    if constexpr (GATHER) return load_el_gather<F>(pts + (uint64_t)(ref & ~PAIR_SIGN_BIT) * AW, wide);
    else return load_planar<F>(reinterpret_cast<const uint4*>(pts), cap, ref);
  }
  B200_D F load_y_raw(uint32_t ref) const
  {
    if constexpr (GATHER) return load_el_gather<F>(pts + (uint64_t)(ref & ~PAIR_SIGN_BIT) * AW + N, wide);
    else return load_planar<F>(reinterpret_cast<const uint4*>(pts) + (uint64_t)(N / 4) * cap, cap, ref);
  }
  B200_D F signed_y(uint32_t ref, const F& y) const
  {
    if constexpr (GATHER) return (ref & PAIR_SIGN_BIT) ? y.neg() : y;
    else return y;
  }
  B200_D F load_y(uint32_t ref) const // sign applied
  {
    if constexpr (GATHER) {
      F y = load_el<F>(pts + (uint64_t)(ref & ~PAIR_SIGN_BIT) * AW + N);
      return (ref & PAIR_SIGN_BIT) ? y.neg() : y;
    } else {
      return load_planar<F>(reinterpret_cast<const uint4*>(pts) + (uint64_t)(N / 4) * cap, cap, ref);
    }
  }
};

constexpr int PAIR_THREADS = 128;

// ---------------------------------------------------------------------------------------------------------------------
// Pass A.  Block b owns the J*128 consecutive pair slots starting at b*J*128; thread t visits slots b*J*128 + j*128 + t,
// j = 0..J-1 (so every access of a warp is to consecutive slots), multiplies the denominators of its real pairs into a
// running product, stores the exclusive running product per pair (planar) and its total in totals[thread].
// ---------------------------------------------------------------------------------------------------------------------
template <class F, bool GATHER>
__global__ void __launch_bounds__(PAIR_THREADS) k_pair_prefix(
  PairSrc<F, GATHER> src, const uint32_t* __restrict__ off, uint32_t nb, uint32_t J, uint4* __restrict__ pbuf, uint64_t pcap,
  uint32_t* __restrict__ totals)
{
  constexpr int N = F::N;
  const uint32_t n = off[nb];
  const uint64_t tid = (uint64_t)blockIdx.x * PAIR_THREADS + threadIdx.x;
  uint64_t q = (uint64_t)blockIdx.x * J * PAIR_THREADS + threadIdx.x;
  F run = F::one();
  // software pipeline, two deep: the (key, value) words of slot j+2 and the x coordinates of slot j+1 are in flight while the
  // product of slot j runs (level 0 chases keys -> point index -> point: three dependent DRAM accesses per slot otherwise)
  struct Idx {
    uint2 kk, vv;
    bool ok;
  };
  auto fetch_idx = [&](uint64_t qq, Idx& ix) {
    ix.ok = (2 * qq + 1 < n);
    ix.kk = make_uint2(0u, 1u);
    ix.vv = make_uint2(0u, 0u);
    if (ix.ok) {
      ix.kk = *reinterpret_cast<const uint2*>(src.keys + 2 * qq);
      if constexpr (GATHER) ix.vv = *reinterpret_cast<const uint2*>(src.vals + 2 * qq);
    }
  };
  bool have = false, nhave = false;
  F x1, x2, nx1, nx2;
  uint32_t r1 = 0, r2 = 0, nr1 = 0, nr2 = 0;
  auto fetch_pts = [&](uint64_t qq, const Idx& ix, bool& hv, F& a, F& b, uint32_t& ra, uint32_t& rb_) {
    hv = ix.ok && ix.kk.x == ix.kk.y;
    if (hv) {
      ra = src.ref_of((uint32_t)(2 * qq), ix.vv.x);
      rb_ = src.ref_of((uint32_t)(2 * qq + 1), ix.vv.y);
      a = src.load_x_raw(ra);
      b = src.load_x_raw(rb_);
    }
  };
  Idx i0, i1, i2;
  fetch_idx(q, i0);
  i1.ok = false; i1.kk = make_uint2(0u, 1u); i1.vv = make_uint2(0u, 0u);
  if (J > 1) fetch_idx(q + PAIR_THREADS, i1);
  fetch_pts(q, i0, have, x1, x2, r1, r2);
  for (uint32_t j = 0; j < J; j++) {
    i2.ok = false; i2.kk = make_uint2(0u, 1u); i2.vv = make_uint2(0u, 0u);
    if (j + 2 < J) fetch_idx(q + 2 * PAIR_THREADS, i2);
    nhave = false;
    if (j + 1 < J) fetch_pts(q + PAIR_THREADS, i1, nhave, nx1, nx2, nr1, nr2);
    if (have) {
      F d;
      if (x1 == x2 || x1.is_zero() || x2.is_zero()) { // rare: the y coordinates decide what kind of sum this is
        const F y1 = src.load_y(r1), y2 = src.load_y(r2);
        int kind;
        d = pair_denominator(x1, y1, x2, y2, kind);
      } else {
        d = x2 - x1;
      }
      store_planar<F>(pbuf, pcap, q, run);
      run = fmul(run, d);
    }
    have = nhave; x1 = nx1; x2 = nx2; r1 = nr1; r2 = nr2;
    i1 = i2;
    q += PAIR_THREADS;
  }
  store_el(totals + tid * N, run);
}

// ---------------------------------------------------------------------------------------------------------------------
// Batch inversion of an array (no zero elements: pass A substitutes 1 for degenerate pairs) by a G-ary product tree.
// ---------------------------------------------------------------------------------------------------------------------
template <class F>
__global__ void __launch_bounds__(128) k_inv_up(const uint32_t* __restrict__ elems, uint32_t n, uint32_t G, uint32_t* __restrict__ prefix,
                                                uint32_t* __restrict__ totals, uint32_t n_threads)
{
  constexpr int N = F::N;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_threads) return;
  const uint64_t lo = (uint64_t)i * G;
  const uint64_t hi = (lo + G < n) ? lo + G : n;
  F run = F::one();
  for (uint64_t j = lo; j < hi; j++) {
    store_el(prefix + j * N, run);
    run = fmul(run, load_el<F>(elems + j * N));
  }
  store_el(totals + (uint64_t)i * N, run);
}

template <class F>
__global__ void __launch_bounds__(128) k_inv_down(uint32_t* __restrict__ elems, uint32_t n, uint32_t G, const uint32_t* __restrict__ prefix,
                                                  const uint32_t* __restrict__ totals_inv, uint32_t n_threads)
{
  constexpr int N = F::N;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_threads) return;
  const uint64_t lo = (uint64_t)i * G;
  const uint64_t hi = (lo + G < n) ? lo + G : n;
  F inv = load_el<F>(totals_inv + (uint64_t)i * N);
  for (uint64_t j = hi; j-- > lo;) {
    const F el = load_el<F>(elems + j * N);
    const F pre = load_el<F>(prefix + j * N);
    store_el(elems + j * N, fmul(inv, pre));
    inv = fmul(inv, el);
  }
}

// root of the tree: one thread inverts n (<= G) elements in place with the serial form of the same trick
template <class F>
__global__ void k_inv_top(uint32_t* __restrict__ elems, uint32_t n, uint32_t* __restrict__ prefix)
{
  constexpr int N = F::N;
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  F run = F::one();
  for (uint32_t j = 0; j < n; j++) {
    store_el(prefix + (uint64_t)j * N, run);
    run = fmul(run, load_el<F>(elems + (uint64_t)j * N));
  }
  F inv = inv_el(run);
  for (uint32_t j = n; j-- > 0;) {
    const F el = load_el<F>(elems + (uint64_t)j * N);
    const F pre = load_el<F>(prefix + (uint64_t)j * N);
    store_el(elems + (uint64_t)j * N, fmul(inv, pre));
    inv = fmul(inv, el);
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// Pass C.  Same slot ownership as pass A, visited backwards: peel each pair's inverse denominator off the thread's inverted
// total, finish the affine add and write the next level (entries that are not part of a real pair are copied through).
// Output position: see k_pair_counts.  OUT_PLANAR selects the layout of the next level buffer (the last level is written
// AoS for the XYZZ accumulation, whose threads walk private slices).
// ---------------------------------------------------------------------------------------------------------------------
template <class F, bool OUT_PLANAR>
B200_D void pair_store_point(uint32_t* __restrict__ out_pts, uint64_t ocap, uint32_t slot, const F& x, const F& y)
{
  if constexpr (OUT_PLANAR) {
    store_planar<F>(reinterpret_cast<uint4*>(out_pts), ocap, slot, x);
    store_planar<F>(reinterpret_cast<uint4*>(out_pts) + (uint64_t)(F::N / 4) * ocap, ocap, slot, y);
  } else {
    store_el(out_pts + (uint64_t)slot * 2 * F::N, x);
    store_el(out_pts + (uint64_t)slot * 2 * F::N + F::N, y);
  }
}

template <class F, bool GATHER, bool OUT_PLANAR>
__global__ void __launch_bounds__(PAIR_THREADS, (F::BYTES <= 32) ? 4 : 1) k_pair_apply(
  PairSrc<F, GATHER> src, const uint32_t* __restrict__ off, const uint32_t* __restrict__ off_next, uint32_t nb, uint32_t J,
  const uint4* __restrict__ pbuf, uint64_t pcap, const uint32_t* __restrict__ totals_inv, uint32_t* __restrict__ out_pts, uint64_t ocap,
  uint32_t* __restrict__ out_keys)
{
  constexpr int N = F::N;
  const uint32_t n = off[nb];
  const uint64_t tid = (uint64_t)blockIdx.x * PAIR_THREADS + threadIdx.x;
  const uint64_t q0 = (uint64_t)blockIdx.x * J * PAIR_THREADS + threadIdx.x;
  if (2 * q0 >= n) return;
  F inv = load_el<F>(totals_inv + tid * N);

  struct In {
    uint32_t k1, k2, r1, r2;
    uint32_t o1, o2; // offset-table words the output position needs (real pair: off[k1], off_next[k1]; else off_next[k1 + 1], off_next[k2])
    int state; // 0 = nothing, 1 = one entry only, 2 = two entries of different buckets, 3 = real pair
    F x1, y1, x2, y2, pre; // y as stored: the sign of a gathered entry is applied when the slot is consumed
  };
  // two-deep software pipeline (see k_pair_prefix): index words two slots ahead, coordinates one slot ahead
  struct Idx {
    uint2 kk, vv;
    int cnt; // entries of the slot that exist: 0, 1 or 2
  };
  auto fetch_idx = [&](uint64_t qq, Idx& ix) {
    const uint64_t e0 = 2 * qq;
    ix.cnt = (e0 + 1 < n) ? 2 : ((e0 < n) ? 1 : 0);
    ix.kk = make_uint2(0u, 0u);
    ix.vv = make_uint2(0u, 0u);
    if (ix.cnt == 2) {
      ix.kk = *reinterpret_cast<const uint2*>(src.keys + e0);
      if constexpr (GATHER) ix.vv = *reinterpret_cast<const uint2*>(src.vals + e0);
    } else if (ix.cnt == 1) {
      ix.kk.x = src.keys[e0];
      if constexpr (GATHER) ix.vv.x = src.vals[e0];
    }
  };
  auto fetch_pts = [&](uint64_t qq, const Idx& ix, In& in) {
    in.state = (ix.cnt == 2) ? ((ix.kk.x == ix.kk.y) ? 3 : 2) : ix.cnt;
    if (in.state == 0) return;
    in.k1 = ix.kk.x; in.k2 = ix.kk.y;
    in.r1 = src.ref_of((uint32_t)(2 * qq), ix.vv.x);
    in.x1 = src.load_x_raw(in.r1);
    in.y1 = src.load_y_raw(in.r1);
    if (in.state >= 2) {
      in.r2 = src.ref_of((uint32_t)(2 * qq + 1), ix.vv.y);
      in.x2 = src.load_x_raw(in.r2);
      in.y2 = src.load_y_raw(in.r2);
    }
    if (in.state == 3) {
      in.pre = load_planar<F>(pbuf, pcap, qq);
      in.o1 = off[in.k1];
      in.o2 = off_next[in.k1];
    } else {
      in.o1 = off_next[in.k1 + 1];
      in.o2 = (in.state == 2) ? off_next[in.k2] : 0u;
    }
  };

  In cur, nxt;
  Idx i0, i1, i2;
  uint64_t q = q0 + (uint64_t)(J - 1) * PAIR_THREADS;
  fetch_idx(q, i0);
  i1.cnt = 0; i1.kk = make_uint2(0u, 0u); i1.vv = make_uint2(0u, 0u);
  if (J > 1) fetch_idx(q - PAIR_THREADS, i1);
  fetch_pts(q, i0, cur);
  for (uint32_t j = J; j-- > 0;) {
    i2.cnt = 0; i2.kk = make_uint2(0u, 0u); i2.vv = make_uint2(0u, 0u);
    if (j > 1) fetch_idx(q - 2 * PAIR_THREADS, i2);
    nxt.state = 0;
    if (j > 0) fetch_pts(q - PAIR_THREADS, i1, nxt);
    if (cur.state != 0) {
      cur.y1 = src.signed_y(cur.r1, cur.y1);
      if (cur.state >= 2) cur.y2 = src.signed_y(cur.r2, cur.y2);
    }
    if (cur.state == 3) {
      int kind;
      const F d = pair_denominator(cur.x1, cur.y1, cur.x2, cur.y2, kind);
      const F dinv = fmul(inv, cur.pre);
      inv = fmul(inv, d);
      F x3, y3;
      if (kind == PK_ADD || kind == PK_DOUBLE) {
        F num;
        if (kind == PK_ADD) {
          num = cur.y2 - cur.y1;
        } else {
          const F xx = fmul(cur.x1, cur.x1);
          num = xx.dbl() + xx;
        }
        const F lam = fmul(num, dinv);
        x3 = fmul(lam, lam) - cur.x1 - cur.x2; // x2 == x1 for a doubling
        y3 = fmul(lam, cur.x1 - x3) - cur.y1;
      } else if (kind == PK_FIRST) {
        x3 = cur.x1; y3 = cur.y1;
      } else if (kind == PK_SECOND) {
        x3 = cur.x2; y3 = cur.y2;
      } else {
        x3 = F::zero(); y3 = F::zero();
      }
      const uint32_t rb = cur.o1;
      const uint32_t a = rb + (rb & 1u);
      const uint32_t slot = cur.o2 + (rb & 1u) + (((uint32_t)(2 * q) - a) >> 1);
      pair_store_point<F, OUT_PLANAR>(out_pts, ocap, slot, x3, y3);
      out_keys[slot] = cur.k1;
    } else if (cur.state != 0) {
      // entry 2q is the last of its run; entry 2q+1 (if any) is the first of the next non-empty run
      const uint32_t s1 = cur.o1 - 1;
      pair_store_point<F, OUT_PLANAR>(out_pts, ocap, s1, cur.x1, cur.y1);
      out_keys[s1] = cur.k1;
      if (cur.state == 2) {
        const uint32_t s2 = cur.o2;
        pair_store_point<F, OUT_PLANAR>(out_pts, ocap, s2, cur.x2, cur.y2);
        out_keys[s2] = cur.k2;
      }
    }
    cur = nxt;
    i1 = i2;
    q -= PAIR_THREADS;
  }
}

}} // namespace b200::msm
