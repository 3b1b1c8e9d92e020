// Shared-memory tile pass for the 4-byte (31-bit) NTT fields: BabyBear, KoalaBear.  Included by ntt.cu (needs PassParams
// and load_twiddle from there).
//
// These are the one family where the transform is genuinely HBM-sized work (8 B of traffic per element and pass against
// ~80 instructions), so the pass is organised around memory behaviour rather than around the register file:
//   * one CTA (256 threads) owns a tile of 2^S strided rows x 32 contiguous columns, i.e. every global access of a warp is
//     one full 128-byte line (B200 fills L2 with whole lines even for a single 32-byte sector, so narrower row segments would
//     multiply the read traffic); a 2^S x 33-word padded tile keeps both row-wise and column-wise shared-memory accesses
//     bank-conflict free;
//   * ~68 KB of shared memory and <= 64 registers per thread leave three CTAs per SM, so one CTA's loads, another's
//     butterflies and a third's stores overlap (the generic tile kernel runs one 512-thread CTA per SM for these fields);
//   * the sub-transform is a pure 2^S-point decimation-in-time NTT on rows stored bit-reversed at load time (free: it is
//     only an address), radix-8 rounds with the lane = column, so every twiddle of a round is warp-uniform and comes from a
//     2^(S-1)-entry table in shared memory; the column-dependent "four-step" factor w_L^(l*k) is applied once per element in
//     the last round, by power iteration per lane (one table look-up per lane and pass);
//   * the pass writes in the autosort layout of ntt.cu (digit written in natural frequency order just above the digits
//     transformed so far), choosing lane = frequency for the first pass and lane = column afterwards so that stores are
//     128-byte runs too.  Natural-order input (kNN; k_ntt31) or the in-place schedule with bit-reversed output (kNR / kNM;
//     k_ntt31_inplace, natural rows + DIF rounds), any batch, coset and inverse scaling folded in.
// Reference semantics: icicle/backend/cpu/include/ntt_cpu.h:69-232 (see ntt.cu header); results are canonical field
// elements, hence bit-identical to the reference whatever the schedule.
#pragma once
// (included inside ntt.cu's anonymous namespace, after PassParams / load_twiddle)

constexpr int NTT31_THREADS = 256;
constexpr int NTT31_ROWPAD = 33;

// Branch-free arithmetic for p < 2^31 on canonical values in [0, p): a conditional +-p is the unsigned minimum of the two
// candidates (the wrong one wraps past 2^32 - p > p), so an add or sub is 3 ALU instructions and the Montgomery product is
// 2 IMAD.WIDE + 1 IMAD + 2 ALU.  Same values as Fp<P>::operator+,-,* (ff.cuh); only the instruction selection differs.
template <class F>
struct A31 {
  static constexpr uint32_t P = F::P::p(0);
  static constexpr uint32_t NP0 = F::P::NP0;
  static __device__ __forceinline__ uint32_t add(uint32_t a, uint32_t b) { const uint32_t s = a + b; return min(s, s - P); }
  static __device__ __forceinline__ uint32_t sub(uint32_t a, uint32_t b) { const uint32_t d = a - b; return min(d, d + P); }
  static __device__ __forceinline__ uint32_t mul(uint32_t a, uint32_t b) // a * b * 2^-32 mod p (b usually a Montgomery-form twiddle)
  {
    uint64_t t = (uint64_t)a * b;
    const uint32_t m = (uint32_t)t * NP0;
    t += (uint64_t)m * P; // low word becomes 0; < 2^62 + 2^63, no overflow
    const uint32_t r = (uint32_t)(t >> 32);
    return min(r, r - P);
  }
};

// FIRST: the round starts at stage 0 (a == 0, blow == 0), where the twiddles with jj == 0 are all 1: 7 of its 12 products vanish.
template <class F, int Q, bool FIRST>
__device__ __forceinline__ void ntt31_dit_group(F (&e)[8], uint32_t a, uint32_t blow, const uint32_t* __restrict__ twsm, uint32_t S)
{
  // Q decimation-in-time stages t = a .. a+Q-1 on the 2^Q rows base + (j << a); blow = base mod 2^a.
  // stage t pairs rows differing in bit t with twiddle w_{2^(t+1)}^(row mod 2^t) = twsm[(row mod 2^t) << (S-1-t)]
#pragma unroll
  for (int i = 0; i < Q; i++) {
    const uint32_t t = a + i;
#pragma unroll
    for (int jj = 0; jj < (1 << i); jj++) {
      const bool trivial = FIRST && jj == 0;
      F w;
      if (!trivial) w.v[0] = twsm[(blow + ((uint32_t)jj << a)) << (S - 1 - t)];
#pragma unroll
      for (int up = 0; up < (1 << (Q - 1 - i)); up++) {
        const int j0 = (up << (i + 1)) | jj, j1 = j0 | (1 << i);
        const uint32_t u = e[j0].v[0], v = trivial ? e[j1].v[0] : A31<F>::mul(e[j1].v[0], w.v[0]);
        e[j0].v[0] = A31<F>::add(u, v);
        e[j1].v[0] = A31<F>::sub(u, v);
      }
    }
  }
}

template <class F>
__device__ __forceinline__ F ntt31_pow(F b, uint32_t ex)
{
  F r = F::one();
  while (ex) {
    if (ex & 1u) r = r * b;
    b = b * b;
    ex >>= 1;
  }
  return r;
}

// Q decimation-in-frequency stages t = a+Q-1 .. a on the rows base + (j << a) (in-place schedule: natural rows in, the
// frequency k ends up at row rev_S(k), which is where that schedule wants it).  LASTR: the round ends at stage 0 (a == 0), where
// the twiddles with jj == 0 are 1.
template <class F, int Q, bool LASTR>
__device__ __forceinline__ void ntt31_dif_group(F (&e)[8], uint32_t a, uint32_t blow, const uint32_t* __restrict__ twsm, uint32_t S)
{
#pragma unroll
  for (int i = Q - 1; i >= 0; i--) {
    const uint32_t t = a + i;
#pragma unroll
    for (int jj = 0; jj < (1 << i); jj++) {
      const bool trivial = LASTR && jj == 0;
      F w;
      if (!trivial) w.v[0] = twsm[(blow + ((uint32_t)jj << a)) << (S - 1 - t)];
#pragma unroll
      for (int up = 0; up < (1 << (Q - 1 - i)); up++) {
        const int j0 = (up << (i + 1)) | jj, j1 = j0 | (1 << i);
        const uint32_t u = e[j0].v[0], v = e[j1].v[0];
        e[j0].v[0] = A31<F>::add(u, v);
        const uint32_t d = A31<F>::sub(u, v);
        e[j1].v[0] = trivial ? d : A31<F>::mul(d, w.v[0]);
      }
    }
  }
}

// DIF round over the whole tile.  LASTR (a == 0): rows r = (b << Q) + j hold frequency k = (rev_Q(j) << (S-Q)) | rev_(S-Q)(b);
// the inter-pass factor g1^k = T_b' * G^rev_Q(j) with b' = rev(b), G = g1^(2^(S-Q)); warps walk b' so that T advances by g1^8.
template <class F, int Q, bool LASTR, int PAD>
__device__ __forceinline__ void ntt31_dif_round(uint32_t* __restrict__ tile, const uint32_t* __restrict__ twsm, uint32_t S, uint32_t a, uint32_t lane,
                                                uint32_t warp, bool interpass, F g1)
{
  constexpr int NW = NTT31_THREADS / 32;
  const uint32_t ngroups = 1u << (S - Q);
  F G = F::one(), T = F::one(), g8 = F::one();
  if (LASTR && interpass) {
    G = g1;
    for (uint32_t i = 0; i < S - Q; i++) G = G * G;
    T = ntt31_pow(g1, warp);
    g8 = ntt31_pow(g1, NW);
  }
  for (uint32_t gi = warp; gi < ngroups; gi += NW) {
    const uint32_t g = LASTR ? (__brev(gi) >> (32 - (S - Q))) : gi; // LASTR: gi is b' = rev(b)
    const uint32_t blow = g & ((1u << a) - 1);
    const uint32_t base = ((g >> a) << (a + Q)) | blow;
    F e[8];
#pragma unroll
    for (int j = 0; j < (1 << Q); j++) e[j].v[0] = tile[(base + ((uint32_t)j << a)) * PAD + lane];
    ntt31_dif_group<F, Q, LASTR>(e, a, blow, twsm, S);
    if (LASTR && interpass) {
      F pw[1 << Q];
      F t = T;
#pragma unroll
      for (int j = 0; j < (1 << Q); j++) {
        pw[j] = t;
        if (j + 1 < (1 << Q)) t.v[0] = A31<F>::mul(t.v[0], G.v[0]);
      }
#pragma unroll
      for (int j = 0; j < (1 << Q); j++) {
        int rj = 0;
#pragma unroll
        for (int b = 0; b < Q; b++) rj |= ((j >> b) & 1) << (Q - 1 - b);
        e[j].v[0] = A31<F>::mul(e[j].v[0], pw[rj].v[0]);
      }
      T.v[0] = A31<F>::mul(T.v[0], g8.v[0]);
    }
#pragma unroll
    for (int j = 0; j < (1 << Q); j++) tile[(base + ((uint32_t)j << a)) * PAD + lane] = e[j].v[0];
  }
}

// One round of Q stages over the whole tile.  LAST: apply the inter-pass twiddle (per lane l, per row k) before storing.
template <class F, int Q, bool LAST, bool FIRST, int PAD>
__device__ __forceinline__ void ntt31_round(uint32_t* __restrict__ tile, const uint32_t* __restrict__ twsm, uint32_t S, uint32_t a, uint32_t lane,
                                            uint32_t warp, bool interpass, F g1)
{
  constexpr int NW = NTT31_THREADS / 32;
  const uint32_t ngroups = 1u << (S - Q);
  // LAST (a == S - Q): group g has base g; row k = g + (j << a) gets g1^k = T_g * G^j with G = g1^(2^a), T_g = g1^g
  F G = F::one(), T = F::one(), g8 = F::one();
  if (LAST && interpass) {
    G = g1;
    for (uint32_t i = 0; i < a; i++) G = G * G;
    T = ntt31_pow(g1, warp);
    g8 = ntt31_pow(g1, NW);
  }
  for (uint32_t g = warp; g < ngroups; g += NW) {
    const uint32_t blow = g & ((1u << a) - 1);
    const uint32_t base = ((g >> a) << (a + Q)) | blow;
    F e[8];
#pragma unroll
    for (int j = 0; j < (1 << Q); j++) e[j].v[0] = tile[(base + ((uint32_t)j << a)) * PAD + lane];
    ntt31_dit_group<F, Q, FIRST>(e, a, blow, twsm, S);
    if (LAST && interpass) {
      F t = T;
#pragma unroll
      for (int j = 0; j < (1 << Q); j++) {
        e[j].v[0] = A31<F>::mul(e[j].v[0], t.v[0]);
        if (j + 1 < (1 << Q)) t.v[0] = A31<F>::mul(t.v[0], G.v[0]);
      }
      T.v[0] = A31<F>::mul(T.v[0], g8.v[0]);
    }
#pragma unroll
    for (int j = 0; j < (1 << Q); j++) tile[(base + ((uint32_t)j << a)) * PAD + lane] = e[j].v[0];
  }
}

// ---- TMA (bulk asynchronous copy) helpers: one 128-byte tile row per cp.async.bulk, completion on an mbarrier (loads) or a
// bulk group (stores).  SASS: UBLKCP.  The rows of a tile are 2^rsh elements apart in HBM, so each row is its own 1-D bulk copy.
constexpr int NTT31_TMAPAD = 36; // words per shared-memory row: 144 B keeps every row 16-byte aligned for the bulk copies
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity)
{
  asm volatile(
    "{\n"
    ".reg .pred p;\n"
    "WAIT_%=:\n"
    "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
    "@p bra DONE_%=;\n"
    "bra WAIT_%=;\n"
    "DONE_%=:\n"
    "}\n" ::"r"(smem_u32(bar)),
    "r"(parity)
    : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)), "l"(gmem_src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes)
{
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit_and_wait_reads()
{
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// TMA = true: tile rows staged by bulk copies (row stride NTT31_TMAPAD); TMA = false: the LDG/STS path (row stride 33), kept
// for buffers that are not 16-byte aligned and as the comparison point (knob ntt31_tma_off = 1 selects it).
template <class F, bool TMA>
__global__ void __launch_bounds__(NTT31_THREADS, 3) k_ntt31(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, PassParams p, uint32_t S)
{
  static_assert(F::N == 1, "k_ntt31 is the 4-byte-field pass");
  constexpr int PAD = TMA ? NTT31_TMAPAD : NTT31_ROWPAD;
  extern __shared__ __align__(16) uint32_t sm[];
  uint32_t* tile = sm;
  uint32_t* twsm = sm + ((size_t)PAD << S);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(twsm + ((size_t)1 << (S - 1))); // 8-byte aligned: both terms are even word counts
  constexpr int NW = NTT31_THREADS / 32;
  const uint32_t T = threadIdx.x, lane = T & 31, warp = T >> 5;
  const uint32_t n_log = p.n_log;
  const uint32_t rsh = n_log - S;           // the digit transformed by a pass sits in the top S bits of the position
  const uint64_t rmask = (1ull << rsh) - 1; // everything below it: [untransformed (lo bits)][transformed so far (done bits)]
  const uint64_t ntt_mask = (1ull << n_log) - 1;
  const uint64_t dom_mask = (1ull << p.dom_log) - 1;
  const uint64_t col0 = (uint64_t)blockIdx.x * 32;
  const uint64_t colg = col0 + lane;
  const uint64_t hi_part = (colg >> rsh) << n_log; // batch index (the tile never straddles two transforms: rsh >= 5)
  const uint64_t lowfull = colg & rmask;
  const uint32_t nrows = 1u << S;
  const bool tma_load = TMA && !p.in_mul;

  if (TMA && T == 0) mbar_init(mbar, 1);
  // ---- twiddle table of the 2^S-point sub-transform ----
  for (uint32_t j = T; j < (1u << (S - 1)); j += NTT31_THREADS) {
    uint64_t ex = (uint64_t)j << (p.dom_log - S);
    if (p.inverse) ex = (0 - ex) & dom_mask;
    twsm[j] = load_twiddle<F>(p.tw, ex).v[0];
  }
  // ---- load: row m of the tile goes to shared-memory row rev_S(m) ----
  if (tma_load) {
    __syncthreads(); // the mbarrier is initialised
    if (T == 0) mbar_expect_tx(mbar, nrows * 128u);
    const uint32_t* __restrict__ row0 = src + (((col0 >> rsh) << n_log) | (col0 & rmask)); // column 0 of the tile: 128-byte aligned
    for (uint32_t m = T; m < nrows; m += NTT31_THREADS) bulk_g2s(tile + (__brev(m) >> (32 - S)) * PAD, row0 + ((uint64_t)m << rsh), 128u, mbar);
  } else {
    constexpr int LD = 8; // independent 128-byte row loads in flight per warp
    const uint32_t* __restrict__ srow = src + (hi_part | lowfull);
    for (uint32_t m0 = warp; m0 < nrows; m0 += NW * LD) {
      uint32_t v[LD];
#pragma unroll
      for (int u = 0; u < LD; u++) {
        const uint32_t m = m0 + u * NW;
        v[u] = (m < nrows) ? srow[m << rsh] : 0u; // (m << rsh) < 2^n_log: a 32-bit offset from this lane's row-0 element
      }
      if (p.in_mul) {
#pragma unroll
        for (int u = 0; u < LD; u++) {
          const uint32_t m = m0 + u * NW;
          if (m < nrows) v[u] = A31<F>::mul(v[u], p.in_mul[(m << rsh) | (uint32_t)lowfull]);
        }
      }
#pragma unroll
      for (int u = 0; u < LD; u++) {
        const uint32_t m = m0 + u * NW;
        if (m < nrows) tile[(__brev(m) >> (32 - S)) * PAD + lane] = v[u];
      }
    }
  }
  // inter-pass factor of this lane: g1 = w_L^l, L = 2^(lo+S), l = untransformed index below the digit
  const bool interpass = (p.lo > 0);
  F g1 = F::one();
  if (interpass) {
    const uint64_t l = lowfull >> p.done;
    uint64_t ex = (l << (p.dom_log - (p.lo + S))) & dom_mask;
    if (p.inverse) ex = (0 - ex) & dom_mask;
    g1 = load_twiddle<F>(p.tw, ex);
  }
  if (tma_load) mbar_wait(mbar, 0);
  __syncthreads();

  // ---- rounds of <= 3 DIT stages ----
  uint32_t a = 0;
  while (a < S) {
    const uint32_t q = (S - a >= 3) ? 3 : (S - a);
    const bool last = (a + q == S);
    if (last) { // S >= 5: the last round is never the first
      if (q == 3) ntt31_round<F, 3, true, false, PAD>(tile, twsm, S, a, lane, warp, interpass, g1);
      else if (q == 2) ntt31_round<F, 2, true, false, PAD>(tile, twsm, S, a, lane, warp, interpass, g1);
      else ntt31_round<F, 1, true, false, PAD>(tile, twsm, S, a, lane, warp, interpass, g1);
    } else if (a == 0) {
      ntt31_round<F, 3, false, true, PAD>(tile, twsm, S, a, lane, warp, false, g1);
    } else {
      ntt31_round<F, 3, false, false, PAD>(tile, twsm, S, a, lane, warp, false, g1);
    }
    a += q;
    __syncthreads();
  }

  // ---- store in the autosort layout: idx = batch | (untransformed << (S+done)) | (k << done) | (transformed so far) ----
  const uint32_t done = p.done;
  F scale = F::one();
  const bool has_scale = p.last && !p.out_mul && p.out_scale;
  if (has_scale) scale.v[0] = p.out_scale[0];
  if (done == 0) {
    // first pass: column c of the tile is the contiguous run [(lowfull0 + c) << S, +2^S) -> lanes walk the frequency k
    const uint64_t base0 = ((col0 >> rsh) << n_log) | ((col0 & rmask) << S);
    if constexpr (TMA) {
      // a warp stores 4 columns x 8 consecutive frequencies (four full 32-byte sectors); with the 36-word row stride the 32 reads
      // hit 32 different banks: bank = (4k + c) mod 32 with k mod 8 = lane & 7 and c = 4*cq + (lane >> 3)
      const uint32_t ntask = 8u * (nrows >> 3);
      for (uint32_t task = warp; task < ntask; task += NW) {
        const uint32_t c = ((task & 7) << 2) | (lane >> 3), k = ((task >> 3) << 3) | (lane & 7);
        const uint64_t idx = base0 + ((uint64_t)c << S) + k;
        uint32_t v = tile[k * PAD + c];
        if (p.last) {
          if (p.out_mul) v = A31<F>::mul(v, p.out_mul[idx & ntt_mask]);
          else if (has_scale) v = A31<F>::mul(v, scale.v[0]);
        }
        dst[idx] = v;
      }
    } else {
      uint32_t* __restrict__ drow = dst + base0 + lane;
      const uint32_t ntask = 32u * (nrows >> 5);
      for (uint32_t task = warp; task < ntask; task += NW) {
        const uint32_t c = task & 31, k = ((task >> 5) << 5) | lane;
        const uint32_t off = (c << S) + ((task >> 5) << 5);
        uint32_t v = tile[k * PAD + c];
        if (p.last) {
          if (p.out_mul) v = A31<F>::mul(v, p.out_mul[(base0 + off + lane) & ntt_mask]);
          else if (has_scale) v = A31<F>::mul(v, scale.v[0]);
        }
        drow[off] = v;
      }
    }
  } else {
    // later passes (done >= 5): the 32 columns of a row stay adjacent -> one 128-byte run per row
    const uint64_t lowfull0 = col0 & rmask;
    const uint64_t base1_0 = ((col0 >> rsh) << n_log) | ((lowfull0 >> done) << (S + done)) | (lowfull0 & ((1ull << done) - 1));
    const bool plain = !(p.last && (p.out_mul || has_scale));
    if (TMA && plain) {
      fence_async_proxy(); // the rounds' generic-proxy writes must be visible to the bulk copies
      __syncthreads();
      for (uint32_t k = T; k < nrows; k += NTT31_THREADS) bulk_s2g(dst + base1_0 + ((uint64_t)k << done), tile + k * PAD, 128u);
      bulk_commit_and_wait_reads(); // shared memory must stay alive until the copies have read it
    } else {
      uint32_t* __restrict__ drow = dst + base1_0 + lane;
      for (uint32_t k = warp; k < nrows; k += NW) {
        uint32_t v = tile[k * PAD + lane];
        if (p.last) {
          if (p.out_mul) v = A31<F>::mul(v, p.out_mul[((base1_0 + lane) | ((uint64_t)k << done)) & ntt_mask]);
          else if (has_scale) v = A31<F>::mul(v, scale.v[0]);
        }
        drow[(uint64_t)k << done] = v;
      }
    }
  }
}

// In-place schedule (output bit-reversed: kNR / kNM, ntt.cu): the pass transforms the digit at bits [lo, lo+S) where it
// lies.  lo >= 5: the 32 tile columns are the 32 lowest position bits (lane = column).  lo == 0 (the last pass): the tile is one
// contiguous run of 32 * 2^S elements (lane = row).  Natural rows + DIF rounds leave frequency k at row rev_S(k), as the schedule wants.
// TMA = true (passes with lo >= 5 only, no input multiplier): rows staged in and out by bulk copies, row stride NTT31_TMAPAD.
template <class F, bool TMA>
__global__ void __launch_bounds__(NTT31_THREADS, 3) k_ntt31_inplace(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, PassParams p, uint32_t S)
{
  static_assert(F::N == 1, "k_ntt31_inplace is the 4-byte-field pass");
  constexpr int PAD = TMA ? NTT31_TMAPAD : NTT31_ROWPAD;
  extern __shared__ __align__(16) uint32_t sm[];
  uint32_t* tile = sm;
  uint32_t* twsm = sm + ((size_t)PAD << S);
  uint64_t* mbar = reinterpret_cast<uint64_t*>(twsm + ((size_t)1 << (S - 1)));
  constexpr int NW = NTT31_THREADS / 32;
  const uint32_t T = threadIdx.x, lane = T & 31, warp = T >> 5;
  const uint32_t n_log = p.n_log, lo = p.lo;
  const uint64_t ntt_mask = (1ull << n_log) - 1;
  const uint64_t dom_mask = (1ull << p.dom_log) - 1;
  const uint32_t rev_shift = 64 - n_log;
  const uint64_t col0 = (uint64_t)blockIdx.x * 32;
  const uint32_t nrows = 1u << S;
  const bool by_col = (lo != 0);
  const uint64_t colg = col0 + lane;
  const uint64_t lomask = (1ull << lo) - 1;
  const uint64_t col_base = by_col ? (((colg >> lo) << (lo + S)) | (colg & lomask)) : 0; // position of (row 0, this lane's column)

  for (uint32_t j = T; j < (1u << (S - 1)); j += NTT31_THREADS) {
    uint64_t ex = (uint64_t)j << (p.dom_log - S);
    if (p.inverse) ex = (0 - ex) & dom_mask;
    twsm[j] = load_twiddle<F>(p.tw, ex).v[0];
  }
  constexpr int LD = 8;
  if constexpr (TMA) { // by_col, no in_mul (launch_ntt31): row m = 128 contiguous bytes at col_base(lane 0) + (m << lo)
    if (T == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (T == 0) mbar_expect_tx(mbar, nrows * 128u);
    const uint64_t cb0 = ((col0 >> lo) << (lo + S)) | (col0 & lomask);
    for (uint32_t m = T; m < nrows; m += NTT31_THREADS) bulk_g2s(tile + m * PAD, src + cb0 + ((uint64_t)m << lo), 128u, mbar);
  } else if (by_col) {
    for (uint32_t m0 = warp; m0 < nrows; m0 += NW * LD) {
      uint32_t v[LD];
#pragma unroll
      for (int u = 0; u < LD; u++) {
        const uint32_t m = m0 + u * NW;
        v[u] = (m < nrows) ? src[col_base + (m << lo)] : 0u;
      }
      if (p.in_mul) {
#pragma unroll
        for (int u = 0; u < LD; u++) {
          const uint32_t m = m0 + u * NW;
          if (m < nrows) v[u] = A31<F>::mul(v[u], p.in_mul[(col_base + (m << lo)) & ntt_mask]);
        }
      }
#pragma unroll
      for (int u = 0; u < LD; u++) {
        const uint32_t m = m0 + u * NW;
        if (m < nrows) tile[m * PAD + lane] = v[u];
      }
    }
  } else {
    const uint32_t ntask = 32u * (nrows >> 5); // (column, 32-row block)
    for (uint32_t t0 = warp; t0 < ntask; t0 += NW * LD) {
      uint32_t v[LD];
#pragma unroll
      for (int u = 0; u < LD; u++) {
        const uint32_t task = t0 + u * NW;
        const uint32_t c = task & 31, m = ((task >> 5) << 5) | lane;
        const uint64_t pos = ((col0 + c) << S) | m;
        v[u] = (task < ntask) ? src[pos] : 0u;
        if (task < ntask && p.in_mul) v[u] = A31<F>::mul(v[u], p.in_mul[pos & ntt_mask]);
      }
#pragma unroll
      for (int u = 0; u < LD; u++) {
        const uint32_t task = t0 + u * NW;
        const uint32_t c = task & 31, m = ((task >> 5) << 5) | lane;
        if (task < ntask) tile[m * PAD + c] = v[u];
      }
    }
  }
  const bool interpass = (lo > 0);
  F g1 = F::one();
  if (interpass) {
    const uint64_t l = colg & lomask;
    uint64_t ex = (l << (p.dom_log - (lo + S))) & dom_mask;
    if (p.inverse) ex = (0 - ex) & dom_mask;
    g1 = load_twiddle<F>(p.tw, ex);
  }
  if constexpr (TMA) mbar_wait(mbar, 0);
  __syncthreads();

  // ---- DIF rounds, top stages first; the last round covers stages [0, q0) with q0 = S mod 3 (or 3) ----
  {
    const uint32_t q0 = (S % 3 == 0) ? 3 : (S % 3);
    uint32_t a = S;
    while (a > q0) {
      a -= 3;
      ntt31_dif_round<F, 3, false, PAD>(tile, twsm, S, a, lane, warp, false, g1);
      __syncthreads();
    }
    if (q0 == 3) ntt31_dif_round<F, 3, true, PAD>(tile, twsm, S, 0, lane, warp, interpass, g1);
    else if (q0 == 2) ntt31_dif_round<F, 2, true, PAD>(tile, twsm, S, 0, lane, warp, interpass, g1);
    else ntt31_dif_round<F, 1, true, PAD>(tile, twsm, S, 0, lane, warp, interpass, g1);
    __syncthreads();
  }

  // ---- store in place; the last pass applies N^-1 / the inverse-coset table (indexed by the logical output index) ----
  F scale = F::one();
  const bool has_scale = p.last && !p.out_mul && p.out_scale;
  if (has_scale) scale.v[0] = p.out_scale[0];
  if constexpr (TMA) { // never the last pass (lo >= 5): plain rows out
    const uint64_t cb0 = ((col0 >> lo) << (lo + S)) | (col0 & lomask);
    fence_async_proxy();
    __syncthreads();
    for (uint32_t m = T; m < nrows; m += NTT31_THREADS) bulk_s2g(dst + cb0 + ((uint64_t)m << lo), tile + m * PAD, 128u);
    bulk_commit_and_wait_reads();
  } else if (by_col) {
    for (uint32_t m = warp; m < nrows; m += NW) {
      const uint64_t pos = col_base + (m << lo);
      uint32_t v = tile[m * PAD + lane];
      if (p.last) {
        if (p.out_mul) v = A31<F>::mul(v, p.out_mul[__brevll(pos & ntt_mask) >> rev_shift]);
        else if (has_scale) v = A31<F>::mul(v, scale.v[0]);
      }
      dst[pos] = v;
    }
  } else {
    const uint32_t ntask = 32u * (nrows >> 5);
    for (uint32_t task = warp; task < ntask; task += NW) {
      const uint32_t c = task & 31, m = ((task >> 5) << 5) | lane;
      const uint64_t pos = ((col0 + c) << S) | m;
      uint32_t v = tile[m * PAD + c];
      if (p.last) {
        if (p.out_mul) v = A31<F>::mul(v, p.out_mul[__brevll(pos & ntt_mask) >> rev_shift]);
        else if (has_scale) v = A31<F>::mul(v, scale.v[0]);
      }
      dst[pos] = v;
    }
  }
}

template <class F>
int launch_ntt31(const uint32_t* src, uint32_t* dst, const PassParams& p, int S, cudaStream_t s)
{
  if constexpr (F::N == 1) {
    const uint64_t total = ((uint64_t)1 << p.n_log) * p.batch;
    const uint64_t total_cols = total >> S;
    const uint64_t blocks = total_cols / 32; // n_log - S >= 5: always a whole number of 32-column tiles
    const size_t smem = (((size_t)NTT31_ROWPAD << S) + ((size_t)1 << (S - 1))) * 4;
    if (p.rot) {
      // bulk-copy (TMA) staging needs 16-byte aligned rows: guaranteed for cudaMalloc'd / staged buffers, checked for the rest
      const bool tma = tune(T_NTT31_TMA_OFF) <= 0 && ((((uintptr_t)src) | ((uintptr_t)dst)) & 15u) == 0;
      if (tma) {
        const size_t smem_t = (((size_t)NTT31_TMAPAD << S) + ((size_t)1 << (S - 1))) * 4 + 16;
        B200_CUDA_TRY(cudaFuncSetAttribute(k_ntt31<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t), B200_UNKNOWN_ERROR);
        k_ntt31<F, true><<<(unsigned)blocks, NTT31_THREADS, smem_t, s>>>(src, dst, p, (uint32_t)S); B200_LAUNCHED(1);
      } else {
        B200_CUDA_TRY(cudaFuncSetAttribute(k_ntt31<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), B200_UNKNOWN_ERROR);
        k_ntt31<F, false><<<(unsigned)blocks, NTT31_THREADS, smem, s>>>(src, dst, p, (uint32_t)S); B200_LAUNCHED(1);
      }
    } else {
      const bool tma = tune(T_NTT31_TMA_OFF) <= 0 && ((((uintptr_t)src) | ((uintptr_t)dst)) & 15u) == 0 && p.lo >= 5 && !p.in_mul;
      if (tma) {
        const size_t smem_t = (((size_t)NTT31_TMAPAD << S) + ((size_t)1 << (S - 1))) * 4 + 16;
        B200_CUDA_TRY(cudaFuncSetAttribute(k_ntt31_inplace<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t), B200_UNKNOWN_ERROR);
        k_ntt31_inplace<F, true><<<(unsigned)blocks, NTT31_THREADS, smem_t, s>>>(src, dst, p, (uint32_t)S); B200_LAUNCHED(1);
      } else {
        B200_CUDA_TRY(cudaFuncSetAttribute(k_ntt31_inplace<F, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), B200_UNKNOWN_ERROR);
        k_ntt31_inplace<F, false><<<(unsigned)blocks, NTT31_THREADS, smem, s>>>(src, dst, p, (uint32_t)S); B200_LAUNCHED(1);
      }
    }
    B200_CUDA_TRY(cudaGetLastError(), B200_UNKNOWN_ERROR);
    return B200_SUCCESS;
  } else {
    return B200_UNKNOWN_ERROR;
  }
}

