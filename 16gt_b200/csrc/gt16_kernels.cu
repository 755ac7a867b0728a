// gt16_kernels.cu -- sm_100a kernels of the 16GT calling hot path.
//
//   k_build_table / k_build_ycsq   error tables (replaces the ~70 s CPU prefill,
//                                  likelihood_cache.cpp:5-54)
//   k_build_decisions              exact memoised decision bitmaps, built by the same evaluator
//   k_stream_tma / k_stream        HBM-bound pass over the counters: decode + classify + ballot -> flag bitmap
//   k_eval_mask                    exact 16-genotype evaluation of the flagged sites, straight from the bitmap
//   k_mask_count/scan/scatter      ordered compaction of what is left (ballot / prefix sums)
//   k_candidates                   full record (likelihoods, posterior, windows, indel patterns)
//   k_bias                         SB/BQB Fisher tests of the records
//   k_op_*                         the reference's operator seam on HBM-resident AoS buffers
//
// No tensor cores: the path is byte/integer decode plus table lookups (HBM roofline).
#include "gt16_kernels.cuh"

namespace gt16 {

// ------------------------------------------------------------------------------ tables
__global__ void __launch_bounds__(256) k_build_table(double *__restrict__ out, float err) {
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  int i = blockIdx.y;
  if (j < kFisherN) out[(size_t) i * kFisherN + j] = table_entry(i, j, err);
}
void launch_build_table(double *out, float err, cudaStream_t s) {
  dim3 grid((kFisherN + 255) / 256, kFisherN);
  k_build_table<<<grid, 256, 0, s>>>(out, err);
}

__global__ void k_build_ycsq(double *__restrict__ out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < kYcsqN) out[i] = chisqr1(i / 10.0);                       // likelihood_cache.cpp:50-53
}
void launch_build_ycsq(double *out, cudaStream_t s) { k_build_ycsq<<<(kYcsqN + 255) / 256, 256, 0, s>>>(out); }

// A key (Wr, Wp, Wq1, Wq3) -- weights of the reference allele, its transition partner and the
// two transversion alleles, no indel evidence -- is "certainly hom-ref" when the exact evaluator,
// run on the representative labelling ref = A, picks hom-ref with a 2x margin over every other
// posterior.  The margin makes the verdict independent of the allele labelling (which only
// permutes the summation order of pD) and of tie-breaking, so bit 0 can never disagree with
// the per-site evaluation.  Everything else (bit 1) is evaluated exactly per site.
__device__ __forceinline__ bool key_needs_exact(const EvalParams &P, uint32_t wr, uint32_t wp, uint32_t wq1,
                                                uint32_t wq3) {
  uint32_t W[6] = {wr, wq1, wp, wq3, 0, 0};
  double lh[16];
  genotype_likelihood16(P.lh, W, 0, lh);
  Posterior po;
  select_genotype(0, lh, P.prior, P.g2i, po);
  if (po.pD == 0.0) return false;                                   // hom-ref by definition (:351-361)
  bool certain = (po.genotype == 0) && (po.bestD > 2.0 * po.secondD);
  return !certain;
}

__global__ void __launch_bounds__(256) k_build_decisions(EvalParams P, uint32_t *s1, uint32_t *s2p, uint32_t *s2q,
                                                         uint32_t *g4) {
  const uint64_t nS1 = kS1Bits, nS2 = (uint64_t) kS2R * kS2A, nG4 = (uint64_t) kG4R * kG4A * kG4A * kG4A;
  uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  bool bit = false;
  uint32_t *dst = nullptr;
  uint64_t idx = 0;
  if (t < nS1) { bit = key_needs_exact(P, (uint32_t) t, 0, 0, 0); dst = s1; idx = t; }
  else if (t < nS1 + nS2) { idx = t - nS1; bit = key_needs_exact(P, (uint32_t) (idx % kS2R), (uint32_t) (idx / kS2R), 0, 0); dst = s2p; }
  else if (t < nS1 + 2 * nS2) { idx = t - nS1 - nS2; bit = key_needs_exact(P, (uint32_t) (idx % kS2R), 0, (uint32_t) (idx / kS2R), 0); dst = s2q; }
  else if (t < nS1 + 2 * nS2 + nG4) {
    idx = t - nS1 - 2 * nS2;
    uint32_t r = (uint32_t) (idx % kG4R), rest = (uint32_t) (idx / kG4R);
    uint32_t p = rest % kG4A; rest /= kG4A;
    uint32_t q1 = rest % kG4A, q3 = rest / kG4A;
    bit = key_needs_exact(P, r, p, q1, q3); dst = g4;
  }
  // all section sizes are multiples of 32, so a warp never straddles two tables
  uint32_t word = __ballot_sync(0xffffffffu, bit);
  if (dst && (threadIdx.x & 31) == 0) dst[idx >> 5] = word;
}
void launch_build_decisions(const EvalParams &P, uint32_t *s1, uint32_t *s2p, uint32_t *s2q, uint32_t *g4,
                            cudaStream_t s) {
  uint64_t total = (uint64_t) kS1Bits + 2ull * kS2R * kS2A + (uint64_t) kG4R * kG4A * kG4A * kG4A;
  k_build_decisions<<<(unsigned) ((total + 255) / 256), 256, 0, s>>>(P, s1, s2p, s2q, g4);
}

// ------------------------------------------------------------------------------ streaming pass
__device__ __forceinline__ bool tab_bit(const uint32_t *__restrict__ t, uint32_t i) {
  return (__ldg(t + (i >> 5)) >> (i & 31)) & 1u;
}

// role-indexed weights -> does this site need exact evaluation?
__device__ __forceinline__ bool lookup_needs_exact(const DecisionTables &D, uint32_t wr, uint32_t wp, uint32_t wq1,
                                                   uint32_t wq3) {
  if ((wp | wq1 | wq3) == 0) return wr < (uint32_t) kS1Bits ? tab_bit(D.s1, wr) : true;
  if ((wq1 | wq3) == 0) return (wr < (uint32_t) kS2R && wp < (uint32_t) kS2A) ? tab_bit(D.s2p, wp * kS2R + wr) : true;
  if (wp == 0 && (wq1 == 0 || wq3 == 0)) {
    uint32_t q = wq1 | wq3;
    return (wr < (uint32_t) kS2R && q < (uint32_t) kS2A) ? tab_bit(D.s2q, q * kS2R + wr) : true;
  }
  if (wr < (uint32_t) kG4R && wp < (uint32_t) kG4A && wq1 < (uint32_t) kG4A && wq3 < (uint32_t) kG4A)
    return tab_bit(D.g4, ((wq3 * kG4A + wq1) * kG4A + wp) * kG4R + wr);
  return true;
}

// ---- TMA bulk-copy + mbarrier primitives (PTX ISA 8.x; SASS: UBLKCP / SYNCS) ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
#ifndef GT16_WAIT_HINT
#define GT16_WAIT_HINT 1000      // ns; tools/probes: polling with __nanosleep 0.1291 ms, hint 1 us 0.1280, hint 10 ms 0.1284
#endif
// try_wait with a suspend-time hint: the warp is parked by the hardware until the phase completes (or the hint runs out)
// instead of polling -- waiting warps must not take issue slots from working ones
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t) GT16_WAIT_HINT) : "memory");
    if (!ok && ++spins > (1u << 22)) __trap();     // a lost bulk copy must not hang the device
  } while (!ok);
}

// Streaming classification of a dense, single-region range of positions [pBeg, pEnd) -- the HBM-bound
// pass of the path.  Persistent CTAs (one per SM); each walks tiles of kTile positions that an elected
// thread stages into shared memory with three 1-D TMA bulk copies per tile (counters + one look-ahead
// entry, packed reference, invalid bitmap) through an mbarrier ring, so ~200 KB per SM is in flight
// independent of the warps' progress.
//
// The integer ALU pipe and the issue slots are what bound this kernel once the loads are asynchronous
// (every LOP3/SHF/ISETP occupies the pipe for two cycles per warp), so the row code does ONE thing: it
// decides whether a site is *settled* -- certainly not a candidate, by a handful of mask tests on its
// counter word (and, in the table-driven variant, one byte lookup) -- and parks the index of every other
// site (~1.5 %) in a per-warp shared-memory queue.  The flag bitmap is written as zeros per tile visit;
// whenever 32 sites are parked the whole warp settles them, one per lane: the counter, its successor and the
// reference base are read again from global memory (they have just come through L2), the general role-indexed
// bitmap test runs, and the verdicts are OR-ed into the bitmap.  One ballot per 32-site row, no store.
//   settled: no coverage; or the reference allele in slot 1 with slot 2 dead (table-driven: and the (Wr) table
//            says hom-ref); shortcut variant: ... or slot 2 holding one allele of weight <= bound
//   parked : overflow-type counter, successor with indel records (-> flagged), a non-reference allele in a
//            compact counter (-> bitmap test), invalid positions never
// Requires pBeg % 128 == 0 (16-byte alignment of all three sources) and arrays padded by one tile.
//
// kShortcut (chosen at launch when the decision bitmaps allow it, see stream_shortcut_bound): the bitmap (Wr) is all
// zero -- a position where only the reference allele was seen is never a candidate -- and every row
// (Wr < 1024, alt <= altBound >> 16) of the two-allele bitmaps is zero.  Then "reference allele in slot 1, plus
// at most one more allele of weight <= bound in slot 2" is settled with mask tests and no table read; that
// is ~99 % of all positions of a 30x snapshot (a sequencing error adds one read of weight <= 4).
#ifndef GT16_ROW_V
#define GT16_ROW_V 1
#endif
#ifndef GT16_STREAM_CTAS
#define GT16_STREAM_CTAS 1      // resident CTAs per SM (each with its own ring)
#endif
template <bool kShortcut, int kRows>
struct StreamSmemT {
  static constexpr int kTile = kStreamThreads * kRows;     // positions per tile
#ifdef GT16_STREAM_STAGES
  static constexpr int kStages = GT16_STREAM_STAGES;
#else
  static constexpr int kStages = kRows >= 8 ? 3 : (kShortcut ? 5 : 4);
#endif
  uint64_t counters[kStages][kTile + 2];
  uint32_t ref[kStages][kTile / 16];
  uint32_t bm[kStages][kTile / 32];
  uint8_t s1[kShortcut ? 16 : kS1Bits];                  // byte per weight: 1 = needs exact evaluation (all zero when kShortcut)
  uint32_t queue[kStreamThreads / 32][64];               // parked sites (index relative to pBeg)
  uint64_t full[kStages];                                // TMA bytes have landed
  uint32_t released[kStages];                            // warps that are done reading the stage
};
static_assert(sizeof(StreamSmemT<false, GT16_STREAM_ROWS>) <= 227 * 1024 && sizeof(StreamSmemT<true, GT16_STREAM_ROWS>) <= 227 * 1024, "shared memory budget");
static_assert((size_t) kStreamThreads * GT16_STREAM_ROWS + 64 <= kStreamPadSites, "arrays are padded by one tile");

// general role-indexed test of one compact (non-overflow) counter; same verdict as lookup_needs_exact
template <bool kShortcut>
__device__ __forceinline__ bool compact_needs_exact(const DecisionTables &D, const uint8_t *__restrict__ s1,
                                                    uint32_t x, uint32_t y, uint32_t ref) {
  uint32_t v1 = (y & 0xFFFFu) ? (x & 0x3FFFu) : 0u, v2 = (y >> 16) ? ((x >> 16) & 0x3FFFu) : 0u;
  const uint32_t r1 = ((x >> 14) & 3u) ^ ref, r2 = (x >> 30) ^ ref;
  if (r1 == r2 && (y >> 16)) v1 = 0;                       // the second slot overwrites the first
  const uint32_t wr = (r1 == 0 ? v1 : 0u) | (r2 == 0 ? v2 : 0u);
  const uint32_t a1 = r1 ? v1 : 0u, a2 = r2 ? v2 : 0u, alt = a1 | a2, ar = a1 ? r1 : r2;
  if (alt == 0) return kShortcut ? false : (s1 ? s1[wr] != 0 : tab_bit(D.s1, wr));   // kShortcut: the (Wr) bitmap is all zero
  if ((a1 && a2) || wr >= (uint32_t) kS2R || alt >= (uint32_t) kS2A) return true;
  return tab_bit(ar == 2 ? D.s2p : D.s2q, alt * kS2R + wr);
}

template <bool kShortcut, int kRows>
__global__ void __launch_bounds__(kStreamThreads, GT16_STREAM_CTAS)
k_stream_tma(const uint64_t *__restrict__ counters, const uint32_t *__restrict__ packedRef,
             const uint32_t *__restrict__ invalid, uint32_t pBeg, uint32_t pEnd, DecisionTables D,
             uint32_t altBound, uint32_t *__restrict__ mask) {
  extern __shared__ __align__(128) unsigned char smemRaw[];
  using Smem = StreamSmemT<kShortcut, kRows>;
  constexpr int kStages = Smem::kStages;
  constexpr uint32_t kTile = Smem::kTile;
  Smem &S = *reinterpret_cast<Smem *>(smemRaw);
  const uint32_t nTiles = (pEnd - pBeg + kTile - 1) / kTile;
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t kBytesC = sizeof(uint64_t) * (kTile + 2), kBytesR = kTile / 4, kBytesB = kTile / 8;
  static_assert(kBytesC % 16 == 0, "bulk copy sizes must be multiples of 16 bytes");
  const uint32_t txBytes = kBytesC + kBytesR + (invalid ? kBytesB : 0);

  auto issue = [&](uint32_t stage, uint32_t tile) {        // elected thread only
    const uint32_t pT = pBeg + tile * kTile;
    mbar_expect_tx(&S.full[stage], txBytes);
    bulk_g2s(S.counters[stage], counters + pT, kBytesC, &S.full[stage]);
    bulk_g2s(S.ref[stage], packedRef + (pT >> 4), kBytesR, &S.full[stage]);
    if (invalid) bulk_g2s(S.bm[stage], invalid + (pT >> 5), kBytesB, &S.full[stage]);
  };

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&S.full[s], 1); S.released[s] = 0; }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (!kShortcut)
    for (uint32_t k = tid; k < (uint32_t) kS1Bits / 4; k += kStreamThreads) {     // expand the s1 bitmap to bytes
      const uint32_t bits = (__ldg(D.s1 + (k >> 3)) >> ((k & 7) << 2)) & 0xFu;
      reinterpret_cast<uint32_t *>(S.s1)[k] = (bits & 1u) | ((bits & 2u) << 7) | ((bits & 4u) << 14) | ((bits & 8u) << 21);
    }
  if (!invalid)
    for (uint32_t k = tid; k < kStages * (kTile / 32); k += kStreamThreads) (&S.bm[0][0])[k] = 0;
  __syncthreads();
  if (tid == 0)
    for (uint32_t s = 0; s < (uint32_t) kStages; ++s) {
      uint32_t t = blockIdx.x + s * gridDim.x;
      if (t < nTiles) issue(s, t);
    }

  // per-warp ring of parked sites
  uint32_t *q = S.queue[warp];
  uint32_t qhead = 0, qtail = 0;
  auto drain = [&](uint32_t count) {                        // lanes < count settle one parked site each
    if (lane < count) {
      const uint32_t b = q[(qhead + lane) & 63u], p = pBeg + b;
      const uint2 c = __ldg(reinterpret_cast<const uint2 *>(counters + p));
      const uint32_t n = __ldg(reinterpret_cast<const uint32_t *>(counters + p + 1));      // successor, low word
      const uint32_t ref = (__ldg(packedRef + (p >> 4)) >> (30u - ((p & 15u) << 1))) & 3u;
      // an overflow-type counter, or a successor that is overflow-type AND carries indel records: exact evaluator.  The
      // successor test is ONE compare: with m = n & 0xE000E000, m >= 0xC000E000 holds for the three true patterns (C000E000,
      // E000C000, E000E000) and otherwise only for a compact successor whose slot 2 holds base T with weight >= 8192 --
      // such a site is flagged needlessly and settled by the exact evaluator.
      const bool special = ((c.x & 0xC000C000u) == 0xC000C000u) | ((n & 0xE000E000u) >= 0xC000E000u);
      if (special || compact_needs_exact<kShortcut>(D, S.s1, c.x, c.y, ref)) atomicOr(mask + (b >> 5), 1u << (b & 31));
    }
    qhead += count;
  };
  // the lane's reference base sits in bits (31 - 2l, 30 - 2l) of its packed word, l = lane % 16: rotated to bits 15:14 it
  // lines up with the base of counter slot 1
  const uint32_t rsh = (48u - ((lane & 15u) << 1)) & 31u, ltMask = (1u << lane) - 1u;
  const uint32_t site0 = warp * (32 * kRows) + lane;        // this lane's site in row 0 of a tile
  const bool deadBit = kShortcut ? false : S.s1[0] != 0;    // verdict for a position without coverage

  uint32_t stage = 0, parity = 0;
  for (uint32_t tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    mbar_wait(&S.full[stage], parity);
    const uint32_t pT = pBeg + tile * kTile;
    const uint32_t lim = (pEnd - pT) < kTile ? pEnd - pT : kTile;
    // the warp's words of the flag bitmap start out as zeros; parked sites set their bits later (same warp)
    if (lane < (uint32_t) kRows && (warp * kRows + lane) * 32u < lim) mask[((pT - pBeg) >> 5) + warp * kRows + lane] = 0u;
#pragma unroll
    for (int h = 0; h < kRows / 4; ++h) {                      // four rows at a time (registers)
      const uint32_t siteH = site0 + 128u * h;
      const uint2 *sc = reinterpret_cast<const uint2 *>(S.counters[stage]) + siteH;
      const uint32_t *sr = S.ref[stage] + (siteH >> 4), *sb = S.bm[stage] + warp * kRows + 4 * h;
      uint2 c[4]; uint32_t nlo[4], rw[4], bw[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) { c[r] = sc[32 * r]; nlo[r] = sc[32 * r + 1].x; rw[r] = sr[2 * r]; bw[r] = sb[r]; }
      // masked: some position of these rows is invalid, or (partial) the range ends inside this tile -- decided
      // once per visit, so that the common rows carry no validity test at all
      auto row = [&](int r, bool partial, bool masked) {
        const uint32_t site = siteH + 32 * r;
        const uint32_t x = c[r].x, y = c[r].y, n = nlo[r];
        const uint32_t rot = __funnelshift_r(rw[r], rw[r], rsh);
        const bool isRef1 = ((x ^ rot) & 0xC000u) == 0;
        const bool dead = y == 0;
        const bool special = ((x & 0xC000C000u) == 0xC000C000u) | ((n & 0xE000E000u) >= 0xC000E000u);
        bool settled;
        if (kShortcut) {
#if GT16_ROW_V == 1
          // shifts by constants are IMADs (fma pipe); the logic ops they replace sit on the alu pipe, which is the busy one
          const bool live1 = (y << 16) != 0;
          const bool oneAllele = y < 0x10000u;                   // slot 2 dead
          const bool tinyAlt = ((x << 2) <= ((altBound << 2) | 0x3FFFFu)) & ((x & 0x3C00u) == 0);
#else
          const bool live1 = (y & 0xFFFFu) != 0;
          const bool oneAllele = y < 0x10000u;                   // slot 2 dead
          const bool tinyAlt = ((x & 0x3FFF0000u) <= altBound) & ((x & 0x3C00u) == 0);   // slot-2 weight <= bound, slot-1 weight < 1024
#endif
          settled = dead | (isRef1 & live1 & (oneAllele | tinyAlt));                     // certainly hom-ref
        } else {
          const bool oneSlot = (y - 1u) < 0xFFFFu;               // slot 1 live, slot 2 dead
          const bool bit = S.s1[x & 0x3FFFu] != 0;
          settled = (dead & !deadBit) | (oneSlot & isRef1 & !bit);
        }
        bool park = special | !settled;
        if (masked) {                                            // invalid positions / positions beyond the range
          bool valid = (bw[r] & (0x80000000u >> lane)) == 0;     // bitmap word of the row, MSB first
          if (partial) valid &= site < lim;
          park &= valid;
        }
        const uint32_t cw = __ballot_sync(0xffffffffu, park);
        if (cw) {
          if ((cw >> lane) & 1u) q[(qtail + __popc(cw & ltMask)) & 63u] = pT - pBeg + site;
          qtail += __popc(cw);
          __syncwarp();                                         // entries (and the zeroed mask words) are visible to the warp
          if (qtail - qhead >= 32u) { drain(32); __syncwarp(); }
        }
      };
      if (lim == kTile) {                                        // all tiles but the last: no per-site range test
        if ((bw[0] | bw[1] | bw[2] | bw[3]) == 0) {
#pragma unroll
          for (int r = 0; r < 4; ++r) row(r, false, false);
        } else {
#pragma unroll
          for (int r = 0; r < 4; ++r) row(r, false, true);
        }
      } else {
#pragma unroll
        for (int r = 0; r < 4; ++r) row(r, true, true);
      }
    }
    __syncwarp();
    // This warp no longer reads the stage.  The LAST warp to say so refills it (no producer warp, nobody waits
    // for a release): the stage's next tile is kStages rounds ahead.
    if (lane == 0) {
      __threadfence_block();
      if (atomicAdd(&S.released[stage], 1u) == kStreamThreads / 32 - 1) {
        S.released[stage] = 0;
        __threadfence_block();
        const uint32_t nt = tile + kStages * gridDim.x;
        if (nt < nTiles) issue(stage, nt);
      }
    }
    if (++stage == (uint32_t) kStages) { stage = 0; parity ^= 1u; }
  }
  __syncwarp();
  drain(qtail - qhead);
}

static cudaError_t configure_stream_runs();     // (defined behind k_stream_runs)
// function attributes are per device: called once per context, with its device current
cudaError_t configure_kernels() {
  constexpr int R = GT16_STREAM_ROWS;
  cudaError_t e = cudaFuncSetAttribute(k_stream_tma<false, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(StreamSmemT<false, R>));
  if (e != cudaSuccess) return e;
  e = configure_stream_runs();
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_stream_tma<true, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(StreamSmemT<true, R>));
}

// what the decision bitmaps (host copies) allow the streaming kernel to skip: returns the largest alt weight A
// such that the (Wr) bitmap is all zero and rows alt = 1..A of both two-allele bitmaps are all zero (0 = none)
uint32_t stream_shortcut_bound(const uint32_t *s1, const uint32_t *s2p, const uint32_t *s2q) {
  for (size_t k = 0; k < kS1Words; ++k) if (s1[k]) return 0;
  uint32_t a = 0;
  for (uint32_t alt = 0; alt < (uint32_t) kS2A; ++alt) {       // row alt = 0 is the (Wr, 0) key again
    bool zero = true;
    for (uint32_t w = 0; w < (uint32_t) kS2R / 32 && zero; ++w) zero = !s2p[alt * (kS2R / 32) + w] && !s2q[alt * (kS2R / 32) + w];
    if (!zero) break;
    a = alt;
  }
  return a;
}

void launch_stream_tma(const SnapView &V, uint32_t pBeg, uint32_t pEnd, const DecisionTables &D, uint32_t shortcutAlt,
                       uint32_t *mask, int numSMs, cudaStream_t s) {
  if (pEnd <= pBeg) return;
  constexpr int R = GT16_STREAM_ROWS;
  constexpr uint32_t tileSites = kStreamThreads * R;
  uint32_t nTiles = (pEnd - pBeg + tileSites - 1) / tileSites;
  const uint32_t ctas = (uint32_t) numSMs * GT16_STREAM_CTAS;
  unsigned grid = nTiles < ctas ? nTiles : ctas;
  if (shortcutAlt)
    k_stream_tma<true, R><<<grid, kStreamThreads, sizeof(StreamSmemT<true, R>), s>>>(V.counters, V.packedRef, V.invalid, pBeg, pEnd, D, shortcutAlt << 16, mask);
  else
    k_stream_tma<false, R><<<grid, kStreamThreads, sizeof(StreamSmemT<false, R>), s>>>(V.counters, V.packedRef, V.invalid, pBeg, pEnd, D, 0, mask);
}

// Table-driven classification of a range of PROCESSED sites [first, first + n) of a region list (exome targets: a few
// hundred positions each; also a dense range whose start is not tile-aligned).  A CTA takes 1,024 consecutive processed
// sites; the regions they fall into (three or four for exome targets) are found through the per-tile cursor and staged
// in shared memory once, so a thread maps a site to its position with a short scan over shared memory instead of a
// binary search over the whole list; every thread then issues the loads of its four sites together (counter,
// successor, reference word, invalid word: coalesced within a region), classifies them with the same tests the parked
// sites of k_stream_tma go through, and the warp's ballots are the words of the flag bitmap.
constexpr int kRegThreads = 256, kRegRows = 4, kRegTile = kRegThreads * kRegRows, kRegTable = 64;
static_assert(kRegTile % kCursorTile == 0, "a CTA's sites start on a cursor tile");
__global__ void __launch_bounds__(kRegThreads) k_stream_regions(SnapView V, DevSiteMap M, uint64_t first, uint32_t n,
                                                                DecisionTables D, uint32_t altBound, uint32_t *__restrict__ mask,
                                                                uint32_t *errWord) {
  __shared__ uint64_t sFirst[kRegTable + 1];
  __shared__ uint32_t sStart[kRegTable];
  // parked sites (everything the mask tests below do not settle): position, both words of the counter,
  __shared__ uint32_t qCount, qPos[kRegTile], qX[kRegTile], qY[kRegTile];
  __shared__ uint16_t qSite[kRegTile];                             // site within the tile | reference base << 12
  const uint32_t tid = threadIdx.x, lane = tid & 31;
  if (tid == 0) qCount = 0;
  const uint64_t i0 = first + (uint64_t) blockIdx.x * kRegTile;
  const uint32_t done = blockIdx.x * (uint32_t) kRegTile, cnt = n - done < (uint32_t) kRegTile ? n - done : (uint32_t) kRegTile;
  // regions that hold sites i0 .. i0 + cnt - 1: [r0, rHi]
  uint32_t r0 = 0, rHi = 0;
  if (M.m.tileReg) { r0 = M.m.tileReg[i0 / kCursorTile]; rHi = M.m.tileReg[(i0 + cnt - 1) / kCursorTile + 1]; }
  else if (M.m.nRegions > 1) rHi = M.m.nRegions - 1;
  const uint32_t nr = rHi - r0 + 1;
  const bool staged = nr <= (uint32_t) kRegTable;
  if (staged) {
    if (tid < nr) { sFirst[tid] = M.m.regFirst[r0 + tid]; sStart[tid] = M.m.regStart[r0 + tid]; }
    if (tid == nr) sFirst[nr] = M.m.regFirst[r0 + nr];
  }
  __syncthreads();
  uint32_t p[kRegRows], r = 0;
  bool in[kRegRows];
#pragma unroll
  for (int k = 0; k < kRegRows; ++k) {
    const uint32_t s = tid + kRegThreads * k;
    in[k] = s < cnt;
    p[k] = 0;
    if (in[k]) {
      const uint64_t i = i0 + s;
      if (staged) {
        while (r + 1 < nr && sFirst[r + 1] <= i) ++r;          // the thread's sites ascend: the scan never goes back
        p[k] = sStart[r] + (uint32_t) (i - sFirst[r]);
      } else {
        p[k] = M.m.tileReg ? pos_of_cursor(M.m, i) : pos_of(M.m, i);
      }
    }
  }
  uint2 c[kRegRows]; uint32_t nx[kRegRows], rw[kRegRows], iw[kRegRows];
#pragma unroll
  for (int k = 0; k < kRegRows; ++k) {
    c[k] = make_uint2(0u, 0u); nx[k] = 0; rw[k] = 0; iw[k] = 0;
    if (in[k]) {
      c[k] = __ldg(reinterpret_cast<const uint2 *>(V.counters + p[k]));
      nx[k] = __ldg(reinterpret_cast<const uint32_t *>(V.counters + (p[k] + 1 < V.L ? p[k] + 1 : V.L - 1)));
      rw[k] = (p[k] >> 4) < V.refWords ? __ldg(V.packedRef + (p[k] >> 4)) : 0u;
      if (V.invalid) iw[k] = __ldg(V.invalid + (p[k] >> 5));
    }
  }
#pragma unroll
  for (int k = 0; k < kRegRows; ++k) {
    bool flag = false;
    if (in[k] && !((iw[k] >> (31 - (p[k] & 31))) & 1u)) {
      const uint32_t ref = (rw[k] >> (30u - ((p[k] & 15u) << 1))) & 3u;
      const uint32_t x = c[k].x, y = c[k].y;
      if ((nx[k] & 0xE000E000u) >= 0xC000E000u) flag = true;              // the successor carries indel records: exact evaluator
      else {
        // the mask tests of k_stream_tma's rows: no coverage, or the reference allele in slot 1 with at most one more
        // allele of weight <= bound (altBound != 0: the decision bitmaps say such a site is never a candidate)
        const bool isRef1 = ((x >> 14) & 3u) == ref, live1 = (y & 0xFFFFu) != 0;
        const bool tinyAlt = ((x & 0x3FFF0000u) <= altBound) & ((x & 0x3C00u) == 0);
        const bool compact = (x & 0xC000C000u) != 0xC000C000u;
        const bool settled = altBound != 0 && compact && (y == 0 || (isRef1 & live1 & ((y < 0x10000u) | tinyAlt)));
        if (!settled) {
          // everything else -- a non-reference allele, an overflow-type counter (three or more alleles, or deep coverage: a
          // few per cent of the sites at exome depth) -- is parked and settled below by the whole CTA, one site per
          // thread, instead of by a divergent branch here
          const uint32_t q = atomicAdd(&qCount, 1u);
          qPos[q] = p[k]; qX[q] = x; qY[q] = y; qSite[q] = (uint16_t) ((tid + kRegThreads * k) | (ref << 12));
        }
      }
    }
    const uint32_t word = __ballot_sync(0xffffffffu, flag);
    const uint32_t s0 = kRegThreads * k + (tid & ~31u);          // the warp's first site of this row
    if (lane == 0 && s0 < cnt) mask[(done + s0) >> 5] = word;
  }
  __syncthreads();
  for (uint32_t q = tid; q < qCount; q += kRegThreads) {         // the general role-indexed test
    const uint32_t ref = qSite[q] >> 12, site = qSite[q] & 0xFFFu;
    if ((qX[q] & 0xC000C000u) != 0xC000C000u) {                    // compact counter: its two slots
      if (compact_needs_exact<false>(D, nullptr, qX[q], qY[q], ref)) atomicOr(mask + ((done + site) >> 5), 1u << (site & 31u));
      continue;
    }
    const uint8_t *rec = ovf_record(V, ovf_region(V, qPos[q]), qY[q]);     // overflow-type: the record's weights
    if (!rec) { atomicOr(errWord, kErrBadSnapshot); continue; }
    const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(rec));           // W[4], F[4]
    const uint2 hi = __ldg(reinterpret_cast<const uint2 *>(rec + 16));      // R[4]
    const uint32_t Wv[4] = {lo.x & 0xFFFFu, lo.x >> 16, lo.y & 0xFFFFu, lo.y >> 16};
    const uint32_t Fv[4] = {lo.z & 0xFFFFu, lo.z >> 16, lo.w & 0xFFFFu, lo.w >> 16};
    const uint32_t Rv[4] = {hi.x & 0xFFFFu, hi.x >> 16, hi.y & 0xFFFFu, hi.y >> 16};
    uint32_t wr[4] = {0u, 0u, 0u, 0u};                                      // weights indexed by role = base ^ ref
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const uint32_t w = (Fv[b] | Rv[b]) ? Wv[b] : 0u, role = (uint32_t) b ^ ref;
      if (role == 0) wr[0] = w; else if (role == 1) wr[1] = w; else if (role == 2) wr[2] = w; else wr[3] = w;
    }
    if (lookup_needs_exact(D, wr[0], wr[2], wr[1], wr[3])) atomicOr(mask + ((done + site) >> 5), 1u << (site & 31u));
  }
}

// exact inline mode (no decision tables): one thread per site, the full evaluator
__global__ void __launch_bounds__(256) k_stream_exact(SnapView V, DevSiteMap M, uint64_t first, uint32_t n, EvalParams P,
                                                      uint32_t *__restrict__ mask, uint32_t *errWord) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool flag = false;
  if (i < n) {
    const uint32_t p = (M.m.nRegions > 1 && M.m.tileReg) ? pos_of_cursor(M.m, first + i) : site_pos(M, first + i);
    uint32_t e = 0;
    flag = site_is_candidate(V, P, p, e);
    if (e) atomicOr(errWord, e);
  }
  uint32_t word = __ballot_sync(0xffffffffu, flag);
  if ((threadIdx.x & 31) == 0 && i < n) mask[i >> 5] = word;
}

// ---- region lists through the TMA ring ----------------------------------------------------------------------------
// The streaming kernel for the targets of an exome (and any other list of regions) when the decision bitmaps allow the
// shortcut tests: persistent CTAs, tiles of kRunTile PROCESSED sites.  A tile is a handful of runs of consecutive
// positions -- the parts of the (seven or eight) regions it covers -- so a producer warp (one per stage), one lane per run, works out
// where each run lies (per-tile cursor + the region arrays), issues its bulk copies (counters incl. one look-ahead
// entry, reference words, invalid-bitmap words: sources rounded out to 16 bytes) into the stage's slabs and writes the
// run table {first site, first position, where its data starts in the slabs} beside them; the sixteen consumer warps
// then do what k_stream_tma's rows do (settle by mask tests, park the rest, one ballot per row, flag words written as
// zeros), addressing shared memory through the run table.  A tile with more runs than the table holds (a list of tiny
// regions) is not staged: its sites are read straight from global memory.
constexpr int kRunTile = 2048, kRunWarps = 16, kRunStages = 4, kRunThreads = (kRunWarps + kRunStages) * 32, kRunParts = 16;
constexpr int kRunCSlab = kRunTile + 4 * kRunParts, kRunRSlab = kRunTile / 16 + 8 * kRunParts, kRunBSlab = kRunTile / 32 + 8 * kRunParts;
struct RunPart { uint32_t siteFirst, pA, cOff, rOff, bOff, wA, bA, pad; };
struct RunSmem {
  uint64_t counters[kRunStages][kRunCSlab];
  uint32_t ref[kRunStages][kRunRSlab];
  uint32_t bm[kRunStages][kRunBSlab];
  RunPart part[kRunStages][kRunParts];
  uint32_t nParts[kRunStages];                           // 0: the tile is not staged (too many runs)
  uint2 queue[kRunWarps][64];                            // parked sites: {bit index, position}
  uint64_t full[kRunStages];                             // the tile's bytes have landed (producer's expect-tx + arrive)
  uint64_t empty[kRunStages];                            // every consumer warp has left the stage (kRunWarps arrivals)
};
static_assert(sizeof(RunSmem) <= 100 * 1024, "two CTAs per SM");
static_assert(kRunTile % kCursorTile == 0 && (kRunCSlab % 2) == 0 && (kRunRSlab % 4) == 0 && (kRunBSlab % 4) == 0, "slab alignment");

__global__ void __launch_bounds__(kRunThreads, 2)
k_stream_runs(SnapView V, DevSiteMap M, uint64_t first, uint32_t n, DecisionTables D, uint32_t altBound,
              uint32_t *__restrict__ mask, uint32_t *errWord) {
  extern __shared__ __align__(128) unsigned char smemRaw[];
  RunSmem &S = *reinterpret_cast<RunSmem *>(smemRaw);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t nTiles = (n + kRunTile - 1) / kRunTile;
  if (tid == 0) {
    for (int k = 0; k < kRunStages; ++k) { mbar_init(&S.full[k], 1); mbar_init(&S.empty[k], kRunWarps); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp >= kRunWarps) {
    // ---------------- producer warps, one per stage (working out a tile's runs is a chain of dependent loads: four of them
    // overlap theirs): lane k stages run k of the tile
    const uint32_t stage = warp - kRunWarps;
    uint32_t round = 0;
    for (uint32_t tile = blockIdx.x + stage * gridDim.x; tile < nTiles; tile += kRunStages * gridDim.x, ++round) {
      // the stage's previous tile has been read by all consumer warps (a parked wait, not a poll: the producers must not take
      // issue slots from the consumers)
      if (round) mbar_wait(&S.empty[stage], (round - 1) & 1u);
      const uint64_t i0 = first + (uint64_t) tile * kRunTile;
      const uint32_t done = tile * (uint32_t) kRunTile, cnt = n - done < (uint32_t) kRunTile ? n - done : (uint32_t) kRunTile;
      const uint32_t r0 = M.m.tileReg[i0 / kCursorTile], rHi = M.m.tileReg[(i0 + cnt - 1) / kCursorTile + 1];
      const uint32_t nr = rHi - r0 + 1;
      uint32_t bytes = 0;
      if (nr <= (uint32_t) kRunParts) {
        // this lane's run: the part of region r0 + lane inside [i0, i0 + cnt)
        uint32_t nSites = 0, pA = 0, sFirst = cnt;
        if (lane < nr) {
          const uint64_t a = M.m.regFirst[r0 + lane], b = M.m.regFirst[r0 + lane + 1];
          const uint64_t sA = a > i0 ? a : i0, sB = b < i0 + cnt ? b : i0 + cnt;
          if (sB > sA) { nSites = (uint32_t) (sB - sA); sFirst = (uint32_t) (sA - i0); pA = M.m.regStart[r0 + lane] + (uint32_t) (sA - a); }
        }
        // sources rounded out to 16 bytes: counters from an even entry (+ the look-ahead entry), words in fours
        const uint32_t e0 = pA & ~1u, entries = nSites ? ((pA - e0) + nSites + 2u) & ~1u : 0u;
        const uint32_t wA = (pA >> 4) & ~3u, words = nSites ? ((((pA + nSites - 1u) >> 4) + 1u - wA) + 3u) & ~3u : 0u;
        const uint32_t bA = (pA >> 5) & ~3u, bwords = (nSites && V.invalid) ? ((((pA + nSites - 1u) >> 5) + 1u - bA) + 3u) & ~3u : 0u;
        uint32_t cAt = entries, rAt = words, bAt = bwords;    // exclusive prefix sums over the lanes: where the run's data goes
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t c1 = __shfl_up_sync(0xffffffffu, cAt, o), r1 = __shfl_up_sync(0xffffffffu, rAt, o), b1 = __shfl_up_sync(0xffffffffu, bAt, o);
          if ((int) lane >= o) { cAt += c1; rAt += r1; bAt += b1; }
        }
        cAt -= entries; rAt -= words; bAt -= bwords;
        if (lane < (uint32_t) kRunParts) {
          RunPart P; P.siteFirst = sFirst; P.pA = pA; P.cOff = cAt + (pA - e0); P.rOff = rAt; P.bOff = bAt; P.wA = wA; P.bA = bA; P.pad = 0;
          S.part[stage][lane] = P;
        }
        bytes = entries * 8u + words * 4u + bwords * 4u;
        const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
        if (lane == 0) S.nParts[stage] = nr;
        __syncwarp();
        if (lane == 0) mbar_expect_tx(&S.full[stage], total);               // arrives, too: the phase completes when the bytes have landed
        __syncwarp();
        if (nSites) {
          bulk_g2s(&S.counters[stage][cAt], V.counters + e0, entries * 8u, &S.full[stage]);
          bulk_g2s(&S.ref[stage][rAt], V.packedRef + wA, words * 4u, &S.full[stage]);
          if (bwords) bulk_g2s(&S.bm[stage][bAt], V.invalid + bA, bwords * 4u, &S.full[stage]);
        }
      } else {
        if (lane == 0) { S.nParts[stage] = 0; mbar_expect_tx(&S.full[stage], 0); }
      }
    }
    return;
  }

  // ---------------- consumer warps
  uint2 *q = S.queue[warp];
  uint32_t qhead = 0, qtail = 0;
  auto drain = [&](uint32_t count) {                        // lanes < count settle one parked site each (as in k_stream_tma)
    if (lane < count) {
      const uint2 e = q[(qhead + lane) & 63u];
      const uint32_t p = e.y;
      const uint2 c = __ldg(reinterpret_cast<const uint2 *>(V.counters + p));
      const uint32_t nx = __ldg(reinterpret_cast<const uint32_t *>(V.counters + p + 1));
      const uint32_t ref = (p >> 4) < V.refWords ? (__ldg(V.packedRef + (p >> 4)) >> (30u - ((p & 15u) << 1))) & 3u : 0u;
      bool flag;
      if ((nx & 0xE000E000u) >= 0xC000E000u) flag = true;
      else if ((c.x & 0xC000C000u) != 0xC000C000u) flag = compact_needs_exact<true>(D, nullptr, c.x, c.y, ref);
      else {
        flag = false;
        const uint8_t *rec = ovf_record(V, ovf_region(V, p), c.y);
        if (!rec) atomicOr(errWord, kErrBadSnapshot);
        else {
          const uint4 lo = __ldg(reinterpret_cast<const uint4 *>(rec));
          const uint2 hi = __ldg(reinterpret_cast<const uint2 *>(rec + 16));
          const uint32_t Wv[4] = {lo.x & 0xFFFFu, lo.x >> 16, lo.y & 0xFFFFu, lo.y >> 16};
          const uint32_t Fv[4] = {lo.z & 0xFFFFu, lo.z >> 16, lo.w & 0xFFFFu, lo.w >> 16};
          const uint32_t Rv[4] = {hi.x & 0xFFFFu, hi.x >> 16, hi.y & 0xFFFFu, hi.y >> 16};
          uint32_t wr[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int b = 0; b < 4; ++b) {
            const uint32_t w = (Fv[b] | Rv[b]) ? Wv[b] : 0u, role = (uint32_t) b ^ ref;
            if (role == 0) wr[0] = w; else if (role == 1) wr[1] = w; else if (role == 2) wr[2] = w; else wr[3] = w;
          }
          flag = lookup_needs_exact(D, wr[0], wr[2], wr[1], wr[3]);
        }
      }
      if (flag) atomicOr(mask + (e.x >> 5), 1u << (e.x & 31));
    }
    qhead += count;
  };
  const uint32_t ltMask = (1u << lane) - 1u;
  constexpr int kRows = kRunTile / (kRunWarps * 32);        // 32-site rows per warp and tile
  uint32_t stage = 0, parity = 0;
  for (uint32_t tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    mbar_wait(&S.full[stage], parity);
    const uint64_t i0 = first + (uint64_t) tile * kRunTile;
    const uint32_t done = tile * (uint32_t) kRunTile, cnt = n - done < (uint32_t) kRunTile ? n - done : (uint32_t) kRunTile;
    const uint32_t nParts = S.nParts[stage];
    if (lane < (uint32_t) kRows && (warp * kRows + lane) * 32u < cnt) mask[(done >> 5) + warp * kRows + lane] = 0u;
    uint32_t k = 0;                                         // run of the lane's current site (its sites ascend)
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
      const uint32_t s = warp * (32 * kRows) + 32 * r + lane;
      bool park = false;
      uint32_t p = 0;
      if (s < cnt) {
        uint32_t x, y, nx, ref;
        bool invalidBit;
        if (nParts) {
          while (k + 1 < nParts && S.part[stage][k + 1].siteFirst <= s) ++k;
          const RunPart P = S.part[stage][k];
          const uint32_t d = s - P.siteFirst;
          p = P.pA + d;
          const uint2 c = reinterpret_cast<const uint2 *>(S.counters[stage])[P.cOff + d];
          x = c.x; y = c.y;
          nx = reinterpret_cast<const uint2 *>(S.counters[stage])[P.cOff + d + 1].x;
          ref = (S.ref[stage][P.rOff + ((p >> 4) - P.wA)] >> (30u - ((p & 15u) << 1))) & 3u;
          invalidBit = V.invalid ? (S.bm[stage][P.bOff + ((p >> 5) - P.bA)] >> (31u - (p & 31u))) & 1u : false;
        } else {                                           // tile of many tiny regions: straight from global memory
          p = pos_of_cursor(M.m, i0 + s);
          const uint2 c = __ldg(reinterpret_cast<const uint2 *>(V.counters + p));
          x = c.x; y = c.y;
          nx = __ldg(reinterpret_cast<const uint32_t *>(V.counters + p + 1));
          ref = (p >> 4) < V.refWords ? (__ldg(V.packedRef + (p >> 4)) >> (30u - ((p & 15u) << 1))) & 3u : 0u;
          invalidBit = V.invalid ? (__ldg(V.invalid + (p >> 5)) >> (31u - (p & 31u))) & 1u : false;
        }
        const bool isRef1 = ((x >> 14) & 3u) == ref, live1 = (y & 0xFFFFu) != 0;
        const bool tinyAlt = ((x & 0x3FFF0000u) <= altBound) & ((x & 0x3C00u) == 0);
        const bool special = ((x & 0xC000C000u) == 0xC000C000u) | ((nx & 0xE000E000u) >= 0xC000E000u);
        const bool settled = (y == 0) | (isRef1 & live1 & ((y < 0x10000u) | tinyAlt));
        park = (special | !settled) & !invalidBit;
      }
      const uint32_t cw = __ballot_sync(0xffffffffu, park);
      if (cw) {
        if ((cw >> lane) & 1u) q[(qtail + __popc(cw & ltMask)) & 63u] = make_uint2(done + s, p);
        qtail += __popc(cw);
        __syncwarp();
        if (qtail - qhead >= 32u) { drain(32); __syncwarp(); }
      }
    }
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&S.empty[stage])) : "memory");
    if (++stage == (uint32_t) kRunStages) { stage = 0; parity ^= 1u; }
  }
  __syncwarp();
  drain(qtail - qhead);
}

static cudaError_t configure_stream_runs() {
  return cudaFuncSetAttribute(k_stream_runs, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(RunSmem));
}

void launch_stream(const SnapView &V, const DevSiteMap &M, uint64_t first, uint32_t n, bool useTables,
                   const DecisionTables &D, uint32_t shortcutAlt, const EvalParams &P, uint32_t *mask, uint32_t *errWord,
                   cudaStream_t s) {
  if (!n) return;
  if (useTables && shortcutAlt && M.m.tileReg && first % kRunTile == 0 && !getenv("GT16_NO_RUNS")) {
    // a region list, the shortcut tests allowed: runs of positions through the TMA ring, two persistent CTAs per SM
    const uint32_t nTiles = (n + kRunTile - 1) / kRunTile;
    static int numSMs = 0;
    if (!numSMs) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&numSMs, cudaDevAttrMultiProcessorCount, dev); }
    unsigned grid = nTiles < (uint32_t) numSMs * 2u ? nTiles : (unsigned) numSMs * 2u;
    k_stream_runs<<<grid, kRunThreads, sizeof(RunSmem), s>>>(V, M, first, n, D, shortcutAlt << 16, mask, errWord);
  } else if (useTables) k_stream_regions<<<(n + kRegTile - 1) / kRegTile, kRegThreads, 0, s>>>(V, M, first, n, D, shortcutAlt << 16, mask, errWord);
  else k_stream_exact<<<(n + 255) / 256, 256, 0, s>>>(V, M, first, n, P, mask, errWord);
}

// ---- sparse snapshots (mode 2, SnapshotHandler.cpp:178-236: {u32 position; SnpCounter} records of 12 bytes) ----------
// counters[pos] = counter for every record whose position lies in [lo, hi); positions must ascend strictly.
__global__ void __launch_bounds__(256) k_scatter_sparse(const uint32_t *__restrict__ rec, uint64_t n, uint64_t *counters,
                                                        uint32_t L, uint32_t lo, uint32_t hi, uint32_t prevPos, uint32_t *errWord) {
  const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const uint32_t pos = rec[3 * t];
  const uint32_t before = t ? rec[3 * (t - 1)] : prevPos;
  if (pos >= L || (pos <= before && !(t == 0 && prevPos == 0xFFFFFFFFu))) { atomicOr(errWord, kErrBadSnapshot); return; }
  if (pos >= lo && pos < hi) counters[pos] = (uint64_t) rec[3 * t + 1] | ((uint64_t) rec[3 * t + 2] << 32);
}
void launch_scatter_sparse(const void *rec, uint64_t n, uint64_t *counters, uint32_t L, uint32_t lo, uint32_t hi,
                           uint32_t prevPos, uint32_t *errWord, cudaStream_t s) {
  if (n) k_scatter_sparse<<<(unsigned) ((n + 255) / 256), 256, 0, s>>>((const uint32_t *) rec, n, counters, L, lo, hi, prevPos, errWord);
}
// zero counters of regions [rA, rB] of the list, widened by `margin` positions on either side and clamped to [lo, hi)
__global__ void __launch_bounds__(256) k_zero_spans(const uint32_t *__restrict__ regStart, const uint64_t *__restrict__ regFirst,
                                                    uint32_t rA, uint64_t *counters, uint32_t lo, uint32_t hi, uint32_t margin) {
  const uint32_t r = rA + blockIdx.x;
  const uint64_t a0 = regStart[r], len = regFirst[r + 1] - regFirst[r];
  const uint64_t a = a0 > margin ? a0 - margin : 0, b = a0 + len + margin;
  const uint64_t x = a < lo ? lo : a, y = b > hi ? hi : b;
  for (uint64_t q = x + threadIdx.x; q < y; q += blockDim.x) counters[q] = 0ull;
}
void launch_zero_spans(const uint32_t *regStart, const uint64_t *regFirst, uint32_t rA, uint32_t rB, uint64_t *counters,
                       uint32_t lo, uint32_t hi, uint32_t margin, cudaStream_t s) {
  if (rB >= rA) k_zero_spans<<<rB - rA + 1, 256, 0, s>>>(regStart, regFirst, rA, counters, lo, hi, margin);
}

// Exact predicate for the flagged sites.  Lane phase: every lane decodes one listed site (overflow
// gather, indel lists of position+1) and, when the site carries no indel evidence, asks the decision
// bitmaps (the test the streaming kernel could not do for overflow-type sites).  Warp phase: the sites
// that still need the full evaluator are taken two at a time; each half-warp computes the 16 genotype
// likelihoods of one site in parallel (one genotype per lane), one lane runs the posterior / argmax.
// A site that is NOT a candidate has its bit cleared in the site mask, so that one more ordered
// compaction of the mask yields the candidate list.
constexpr int kEvalThreads = 128;
#ifndef GT16_EVAL_BLOCKS
#define GT16_EVAL_BLOCKS 8      // resident blocks per SM (register budget 65536 / (128 x blocks)); tools/probes sweeps it: 6 -> 8 = -5 us, 10 = +13 us
#endif
constexpr int kEvalQueue = 64;                       // per-warp ring of flagged site indices
#ifndef GT16_EVAL_CHUNK
#define GT16_EVAL_CHUNK 8
#endif
constexpr int kEvalSlotChunk = GT16_EVAL_CHUNK;      // side-buffer entries a warp reserves at a time (unused ones are marked)

// Exact evaluation of up to 32 flagged sites, one per lane (site index i relative to `base`, lanes with !active idle).
// Lane phase: decode + the general bitmap test; warp phase: the sites that still need the full evaluator are taken
// two at a time, each half-warp computing the 16 genotype likelihoods of one site (one genotype per lane).  A site
// that is NOT a candidate has its bit cleared in the site mask.
__device__ __forceinline__ void eval_group(const SnapView &V, const DevSiteMap &M, uint64_t base, uint32_t i, bool active,
                                           bool useTables, const DecisionTables &D, const EvalParams &P,
                                           uint32_t *__restrict__ siteMask, uint32_t *errWord,
                                           double (*lhS)[16], uint32_t (*wpS)[6],
                                           EvalOut *__restrict__ part, uint32_t *count, uint32_t cap,
                                           uint32_t &slotNext, uint32_t &slotEnd) {
  const uint32_t lane = threadIdx.x & 31, half = lane >> 4, g = lane & 15;
  {
    uint32_t e = 0;
    bool cand = false, full = false;
    uint32_t W[6] = {0, 0, 0, 0, 0, 0};
    int ref = 0; uint32_t validPlus = 0, delLen = 0;
    if (active) {
      const uint32_t p = site_pos(M, base + i);
      BaseCounts bc;
      decode_bases(V, p, bc);
      if (bc.bad) e |= kErrBadSnapshot;
      else if (bc.valid) {
        IndelInfo ii; PatRef d0, d1;
        decode_indel_refs(V, p, ii, d0, d1);
        if (ii.bad) e |= kErrBadSnapshot;
        else {
          ref = (int) ref_base(V, p);
          W[0] = bc.W[0]; W[1] = bc.W[1]; W[2] = bc.W[2]; W[3] = bc.W[3]; W[4] = ii.W4; W[5] = ii.W5;
          validPlus = ii.isValidPlus;
          delLen = d0.type == 'D' ? d0.len : 0u;
          full = true;
          if (useTables && (ii.W4 | ii.W5) == 0) {
            const uint32_t wr = ref == 0 ? W[0] : ref == 1 ? W[1] : ref == 2 ? W[2] : W[3];
            const uint32_t wp = ref == 0 ? W[2] : ref == 1 ? W[3] : ref == 2 ? W[0] : W[1];
            const uint32_t wq1 = ref == 0 ? W[1] : ref == 1 ? W[0] : ref == 2 ? W[3] : W[2];
            const uint32_t wq3 = ref == 0 ? W[3] : ref == 1 ? W[2] : ref == 2 ? W[1] : W[0];
            full = lookup_needs_exact(D, wr, wp, wq1, wq3);
          }
        }
      }
    }
    uint32_t todo = __ballot_sync(0xffffffffu, full);
    while (todo) {
      // the two lowest pending lanes -> half-warps 0 and 1
      const int src0 = __ffs(todo) - 1;
      uint32_t rest = todo & (todo - 1);
      const int src1 = rest ? __ffs(rest) - 1 : src0;
      rest = rest ? rest & (rest - 1) : 0;
      const int src = half ? src1 : src0;
      uint32_t Ws[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) Ws[k] = __shfl_sync(0xffffffffu, W[k], src);
      const int refS = __shfl_sync(0xffffffffu, ref, src);
      const uint32_t vpS = __shfl_sync(0xffffffffu, validPlus, src);
      const uint32_t siteS = __shfl_sync(0xffffffffu, i, src), delS = __shfl_sync(0xffffffffu, delLen, src);
      // penalised weights (penalised_weights(), FisherExactTest.cpp:251-266) into shared memory: the likelihood
      // classes index them by allele
      if (g < 6) {
        uint32_t v = g == 0 ? Ws[0] : g == 1 ? Ws[1] : g == 2 ? Ws[2] : g == 3 ? Ws[3] : g == 4 ? Ws[4] : Ws[5];
        const uint32_t indelCount = Ws[4] + Ws[5];
        if ((int) g == refS && indelCount) {
          double tmp = (double) indelCount * (1.375f) + 0.5f;
          tmp = v - tmp;
          if (tmp <= 0) tmp = 0;
          v = (uint32_t) tmp;
        }
        wpS[half][g] = v;
      }
      __syncwarp();
      int wt = 0;
#pragma unroll
      for (int k = 0; k < 6; ++k) wt += (int) wpS[half][k];
      lhS[half][g] = genotype_lh_one(P.lh, wpS[half], wt, (int) g);
      __syncwarp();
      // selectGenotype (FisherExactTest.cpp:317-509) across the half-warp: the sum in the reference's order in every
      // lane, ONE posterior division per lane, the winner (lowest index holding the maximum) by shuffles.  pD == 0
      // makes the reference call hom-ref (:351-361): that lane alone gets a non-zero score.
      double pD = lhS[half][0];
#pragma unroll
      for (int k = 1; k < 16; ++k) pD = pD + lhS[half][k];
      const int homRef = hom_ref_genotype(refS);
      double f0 = pD == 0.0 ? ((int) g == homRef ? 1.0 : 0.0) : lhS[half][g] / pD * P.prior[(refS << 6) | P.g2i[g]];
      const double fMine = f0;
      int genotype = (int) g;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, f0, o);
        const int gw = __shfl_xor_sync(0xffffffffu, genotype, o);
        if (w > f0 || (w == f0 && gw < genotype)) { f0 = w; genotype = gw; }
      }
      const int res = (genotype != homRef ? 1 : 0) | ((genotype > 9 && vpS == 0) ? 2 : 0);   // bit0 candidate, bit1 genotype error (SNP_Caller.cpp:235-239)
      // A candidate (pD != 0 then: pD == 0 is hom-ref by definition) leaves what has just been computed to the record
      // and Fisher-test kernels: likelihoods, posteriors, best / second (the runner-up is the largest of the other
      // fifteen, argmax_posteriors()), genotype, indel kind and the deletion length the neighbour tests skip.
      const bool emit = (res & 1) && (half == 0 || src1 != src0);
      if (__any_sync(0xffffffffu, emit)) {
        double f1 = (int) g == genotype ? -1.0 : fMine;
#pragma unroll
        for (int o = 8; o > 0; o >>= 1) {
          const double w = __shfl_xor_sync(0xffffffffu, f1, o);
          if (w > f1) f1 = w;
        }
        // entries of the side buffer are handed out from a chunk the warp reserves with ONE atomic (the reply of a
        // contended global atomic takes a microsecond: per candidate it was a fifth of this kernel's stall samples)
        const uint32_t emits = __ballot_sync(0xffffffffu, emit && g == 0);          // bit 0 and / or bit 16
        if (slotEnd - slotNext < (uint32_t) __popc(emits)) {
          if (slotNext < slotEnd && lane == 0 && slotNext < cap) part[slotNext].site = 0xFFFFFFFFu;   // at most one entry is left: unused
          uint32_t base = 0;
          if (lane == 0) base = atomicAdd(count, (uint32_t) kEvalSlotChunk);
          slotNext = __shfl_sync(0xffffffffu, base, 0);
          slotEnd = slotNext + (uint32_t) kEvalSlotChunk;
        }
        const uint32_t slot = slotNext + ((half && (emits & 1u)) ? 1u : 0u);
        slotNext += (uint32_t) __popc(emits);
        if (emit) {
          if (slot < cap) {
            EvalOut &eo = part[slot];
            eo.lh[g] = lhS[half][g];
            eo.flh[g] = fMine;
            if (g == 0) {
              const bool indel = genotype > 9 && vpS != 0;
              eo.pD = pD; eo.bestD = f0; eo.secondD = f1; eo.site = siteS;
              eo.genotype = (uint8_t) genotype; eo.isIndel = indel ? (uint8_t) (1 + vpS) : (uint8_t) 0;
              eo.nextBp = indel ? (uint8_t) delS : (uint8_t) 0; eo.pad = 0;
            }
          } else if (g == 0) {
            e |= kErrRecordOverflow;
          }
        }
      }
      const int res0 = __shfl_sync(0xffffffffu, res, 0), res1 = __shfl_sync(0xffffffffu, res, 16);
      if ((int) lane == src0) { cand = res0 & 1; if (res0 & 2) e |= kErrGenotype; }
      if ((int) lane == src1 && src1 != src0) { cand = res1 & 1; if (res1 & 2) e |= kErrGenotype; }
      todo = rest;
      __syncwarp();
    }
    if (e) atomicOr(errWord, e);
    if (active && !cand) atomicAnd(siteMask + (i >> 5), ~(1u << (i & 31)));
  }
}

// Exact evaluation of every site flagged in mask words [wordBegin, wordBegin + nWords) of the call's site mask (bit k of
// word w = site 32 w + k relative to `base`).  No compaction pass and no host round trip in front of it: a warp walks
// chunks of 32 mask words, parks the set bits in a shared-memory ring and evaluates them 32 at a time (flag density is
// ~0.4 %, so one chunk of 1,024 sites yields ~4 sites).  *flagged += number of sites evaluated.
__global__ void __launch_bounds__(kEvalThreads, GT16_EVAL_BLOCKS) k_eval_mask(SnapView V, DevSiteMap M, uint64_t base, uint32_t wordBegin,
                                                            uint32_t nWords, bool useTables, DecisionTables D, EvalParams P,
                                                            uint32_t *__restrict__ siteMask, uint32_t *errWord,
                                                            uint32_t *flagged, EvalOut *__restrict__ part, uint32_t *count,
                                                            uint32_t cap) {
  __shared__ double lhS[kEvalThreads / 32][2][16];
  __shared__ uint32_t wpS[kEvalThreads / 32][2][6];
  __shared__ uint32_t ringS[kEvalThreads / 32][kEvalQueue];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t *ring = ringS[warp];
  const uint32_t warpsTotal = gridDim.x * (kEvalThreads / 32), nChunks = (nWords + 31) / 32;
  uint32_t head = 0, tail = 0, seen = 0;
  uint32_t slotNext = 0, slotEnd = 0;                   // the warp's reserved entries of part[]: [slotNext, slotEnd)
  // plain loads: the words are cleared by atomics of this very launch, but a word is read by exactly one lane and only
  // bits that lane has queued are ever cleared -- so the word of the warp's NEXT chunk may be fetched one chunk ahead
  uint32_t c = blockIdx.x * (kEvalThreads / 32) + warp;
  uint32_t bitsAhead = (c < nChunks && c * 32 + lane < nWords) ? siteMask[wordBegin + c * 32 + lane] : 0u;
  for (; c < nChunks; c += warpsTotal) {
    const uint32_t wi = c * 32 + lane;
    uint32_t bits = bitsAhead;
    {
      const uint32_t cn = c + warpsTotal;
      bitsAhead = (cn < nChunks && cn * 32 + lane < nWords) ? siteMask[wordBegin + cn * 32 + lane] : 0u;
    }
    seen += __popc(bits);
    uint32_t any = __ballot_sync(0xffffffffu, bits != 0);
    while (any) {                                       // one bit per lane and round
      const bool have = bits != 0;
      if (have) {
        const uint32_t b = __ffs(bits) - 1;
        bits &= bits - 1;
        ring[(tail + __popc(any & ((1u << lane) - 1u))) & (kEvalQueue - 1)] = (wordBegin + wi) * 32 + b;
      }
      tail += __popc(any);
      __syncwarp();
      if (tail - head >= 32u) {
        eval_group(V, M, base, ring[(head + lane) & (kEvalQueue - 1)], true, useTables, D, P, siteMask, errWord, lhS[warp], wpS[warp], part, count, cap, slotNext, slotEnd);
        head += 32;
        __syncwarp();
      }
      any = __ballot_sync(0xffffffffu, bits != 0);
    }
  }
  if (tail != head) {
    const bool active = lane < tail - head;
    eval_group(V, M, base, active ? ring[(head + lane) & (kEvalQueue - 1)] : 0u, active, useTables, D, P, siteMask, errWord, lhS[warp], wpS[warp], part, count, cap, slotNext, slotEnd);
  }
  for (uint32_t k = slotNext + lane; k < slotEnd; k += 32) if (k < cap) part[k].site = 0xFFFFFFFFu;   // reserved, not used
  seen = __reduce_add_sync(0xffffffffu, seen);
  if (lane == 0 && seen) atomicAdd(flagged, seen);
}
void launch_eval_mask(const SnapView &V, const DevSiteMap &M, uint64_t base, uint32_t firstSite, uint32_t nSites,
                      bool useTables, const DecisionTables &D, const EvalParams &P, uint32_t *siteMask,
                      uint32_t *errWord, uint32_t *flagged, EvalOut *part, uint32_t *count, uint32_t cap, cudaStream_t s) {
  if (!nSites) return;
  const uint32_t nWords = (nSites + 31) / 32, nChunks = (nWords + 31) / 32;
  unsigned grid = (nChunks + (kEvalThreads / 32) - 1) / (kEvalThreads / 32);
  if (grid > 148 * GT16_EVAL_BLOCKS) grid = 148 * GT16_EVAL_BLOCKS;      // one resident wave (launch bounds)
  k_eval_mask<<<grid, kEvalThreads, 0, s>>>(V, M, base, firstSite / 32, nWords, useTables, D, P, siteMask, errWord, flagged, part, count, cap);
}

// ------------------------------------------------------------------------------ ordered compaction
// blockCounts[b] = set bits of the block's kScanBlock words; subCounts[b * 8 + k] = set bits of its k-th run of 32 words
__global__ void __launch_bounds__(kScanBlock) k_mask_count(const uint32_t *__restrict__ mask, uint32_t nWords,
                                                          uint32_t *__restrict__ blockCounts, uint32_t *__restrict__ subCounts) {
  __shared__ uint32_t warpSum[kScanBlock / 32];
  uint32_t w = blockIdx.x * kScanBlock + threadIdx.x;
  uint32_t c = w < nWords ? __popc(mask[w]) : 0;
  c = __reduce_add_sync(0xffffffffu, c);
  if ((threadIdx.x & 31) == 0) { warpSum[threadIdx.x >> 5] = c; subCounts[blockIdx.x * (kScanBlock / 32) + (threadIdx.x >> 5)] = c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int k = 0; k < kScanBlock / 32; ++k) t += warpSum[k];
    blockCounts[blockIdx.x] = t;
  }
}
void launch_mask_count(const uint32_t *mask, uint32_t nItems, uint32_t *blockCounts, uint32_t *subCounts, cudaStream_t s) {
  uint32_t nb = scan_blocks(nItems);
  if (nb) k_mask_count<<<nb, kScanBlock, 0, s>>>(mask, (nItems + 31) / 32, blockCounts, subCounts);
}

// Exclusive scan of the block counts in two levels, so that a whole-genome call (380 K blocks) costs what a chromosome
// costs: k_scan_chunks -- one CTA per chunk of 1024 counts, exclusive scan in place + the chunk's total --, then
// k_scan_top -- one CTA scans the (at most 512) chunk totals.  A block's offset is counts[b] + chunkOffsets[b >> 10].
__device__ __forceinline__ uint32_t cta_exclusive_scan_1024(uint32_t v, uint32_t *warpTot, uint32_t *total) {
  uint32_t x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
  if ((threadIdx.x & 31) == 31) warpTot[threadIdx.x >> 5] = x;
  __syncthreads();
  if (threadIdx.x < 32) {
    uint32_t w = warpTot[threadIdx.x], z = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, z, o); if (threadIdx.x >= o) z += y; }
    warpTot[threadIdx.x] = z - w;   // exclusive warp offsets
    if (threadIdx.x == 31) *total = z;
  }
  __syncthreads();
  return warpTot[threadIdx.x >> 5] + x - v;
}
__global__ void __launch_bounds__(1024) k_scan_chunks(uint32_t *counts, uint32_t n, uint32_t *chunkTotals) {
  __shared__ uint32_t warpTot[32];
  __shared__ uint32_t tot;
  const uint32_t i = blockIdx.x * 1024u + threadIdx.x;
  const uint32_t v = i < n ? counts[i] : 0u;
  const uint32_t excl = cta_exclusive_scan_1024(v, warpTot, &tot);
  if (i < n) counts[i] = excl;
  if (threadIdx.x == 0) chunkTotals[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(1024) k_scan_top(uint32_t *chunkTotals, uint32_t nChunks, uint32_t *total) {
  __shared__ uint32_t warpTot[32];
  __shared__ uint32_t tot;
  const uint32_t v = threadIdx.x < nChunks ? chunkTotals[threadIdx.x] : 0u;
  const uint32_t excl = cta_exclusive_scan_1024(v, warpTot, &tot);
  if (threadIdx.x < nChunks) chunkTotals[threadIdx.x] = excl;
  if (threadIdx.x == 0) *total = tot;
}
void launch_scan_counts(uint32_t *blockCounts, uint32_t nBlocks, uint32_t *chunkOffsets, uint32_t *total, cudaStream_t s) {
  const uint32_t nChunks = (nBlocks + 1023) / 1024;                 // <= 512 for any 32-bit site count
  k_scan_chunks<<<nChunks, 1024, 0, s>>>(blockCounts, nBlocks, chunkOffsets);
  k_scan_top<<<1, 1024, 0, s>>>(chunkOffsets, nChunks, total);
}

// Place of every candidate in the ordered output: rank = number of set mask bits before its site, from the scanned block
// counts plus the (at most kScanBlock) words of its own block.  One warp per side-buffer entry; the count is read on the
// device.  list[rank] = entry: the record kernel walks the ranks, so entries a warp of the evaluator reserved and did not
// use cost nothing there, and neighbouring warps write neighbouring records.
__global__ void __launch_bounds__(256) k_rank(const uint32_t *__restrict__ mask, uint32_t firstWord,
                                              const uint32_t *__restrict__ blockOffsets,
                                              const uint32_t *__restrict__ chunkOffsets,
                                              const uint32_t *__restrict__ subCounts, const EvalOut *__restrict__ part,
                                              const uint32_t *__restrict__ count, uint32_t cap,
                                              uint32_t *__restrict__ rankOf, uint32_t *__restrict__ list) {
  const uint32_t n = *count < cap ? *count : cap;
  const uint32_t lane = threadIdx.x & 31, warpsTotal = gridDim.x * (blockDim.x >> 5);
  for (uint32_t j = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); j < n; j += warpsTotal) {
    const uint32_t site = part[j].site;
    if (site == 0xFFFFFFFFu) continue;                                  // an entry a warp reserved and did not use
    // set bits before the site's word inside its block: whole runs of 32 words from the sub-counts (lanes 0..7), the
    // words of its own run one per lane
    const uint32_t w = (site >> 5) - firstWord, run = w / 32, w0 = run * 32;
    uint32_t c = 0;
    if (lane < (run & (kScanBlock / 32 - 1))) c = subCounts[(run & ~(uint32_t) (kScanBlock / 32 - 1)) + lane];
    if (w0 + lane < w) c += __popc(mask[firstWord + w0 + lane]);
    c = __reduce_add_sync(0xffffffffu, c);
    if (lane == 0) {
      const uint32_t blk = w / kScanBlock;
      const uint32_t r = chunkOffsets[blk >> 10] + blockOffsets[blk] + c + __popc(mask[firstWord + w] & ((1u << (site & 31)) - 1u));
      rankOf[j] = r;
      if (r < cap) list[r] = j;        // (more candidates than the buffers hold: the call is repeated with larger ones)
    }
  }
}
void launch_rank(const uint32_t *siteMask, uint32_t firstSite, const uint32_t *blockOffsets, const uint32_t *chunkOffsets,
                 const uint32_t *subCounts, const EvalOut *part, const uint32_t *count, uint32_t cap, uint32_t *rankOf,
                 uint32_t *list, cudaStream_t s) {
  unsigned grid = (cap + 7) / 8;
  if (grid > 148 * 8) grid = 148 * 8;
  k_rank<<<grid, 256, 0, s>>>(siteMask, firstSite / 32, blockOffsets, chunkOffsets, subCounts, part, count, cap, rankOf, list);
}

// ------------------------------------------------------------------------------ candidates
// One WARP per called site (computeMeta + outputPossibleSNPs, SNP_Caller.cpp:101-393).  The work of a
// record is a few hundred dependent gathers (two 101-entry windows of decoded counters) plus up to 24
// serial float Fisher recurrences; spread over the lanes it becomes: one genotype likelihood per lane
// (16), four window entries per lane and side, and warp reductions for the window sums; the Fisher
// tests of the record are done by k_bias.  Every leaf routine is the one the serial build_called_site()
// uses, so the bits are the same.
constexpr int kCandThreads = 128;
#ifndef GT16_CAND_BLOCKS
#define GT16_CAND_BLOCKS 8      // resident blocks per SM (register budget)
#endif
#ifndef GT16_CAND_GRID
#define GT16_CAND_GRID GT16_CAND_BLOCKS      // blocks per SM that are launched
#endif
#ifndef GT16_BIAS_GRID
#define GT16_BIAS_GRID 4        // blocks of the Fisher-test kernel per SM
#endif
// what the window sums need of one position (SNP_Caller.cpp:276-367): strand count, weighted depth, flags
struct WinLite { uint32_t sc, w, fl; };                      // fl: bit0 valid, bit1 C-or-G seen, bit2 G seen

// decode_bases() reduced to those sums (same conditions: a slot is live iff its strand counts are not both zero,
// an overflow record's base is counted iff F|R != 0; a record index out of range leaves a valid, empty entry).
// Overflow records are 32-byte aligned in device memory (alloc_snapshot), read with two vector loads.
__device__ __forceinline__ WinLite win_decode(const SnapView &V, uint32_t p, uint64_t c) {
  WinLite e; e.sc = 0; e.w = 0; e.fl = 1;
  const uint32_t lo = (uint32_t) c, hi = (uint32_t) (c >> 32);
  if (!is_overflow_type(c)) {
    // branch-free: a dead slot has both strand counts zero, so the four count bytes can be summed as they are; its
    // weight and base fields are masked out (they need not be zero).  base in {C, G} <=> bit0 ^ bit1, base == G <=> 10b.
    const uint32_t live = ((hi & 0xFFFFu) ? 0xFFFFu : 0u) | ((hi >> 16) ? 0xFFFF0000u : 0u);
    const uint32_t f = lo & live;
    e.sc = __dp4a(hi, 0x01010101u, 0u);
    e.w = (f & 0x3FFFu) + ((f >> 16) & 0x3FFFu);
    const uint32_t cg = (f ^ (f >> 1)) & 0x40004000u;            // bit 14 / bit 30: slot holds C or G
    const uint32_t gg = (f & ~(f << 1)) & 0x80008000u;           // bit 15 / bit 31: slot holds G (bit1 set, bit0 clear)
    e.fl = 1u | (cg ? 2u : 0u) | (gg ? 4u : 0u);
  } else {
    const uint8_t *r = ovf_record(V, ovf_region(V, p), hi);
    if (r) {
      const uint4 a = __ldg(reinterpret_cast<const uint4 *>(r));           // W[4], F[4] (u16 each)
      const uint2 b = __ldg(reinterpret_cast<const uint2 *>(r + 16));      // R[4]
      const uint32_t F[4] = {a.z & 0xFFFFu, a.z >> 16, a.w & 0xFFFFu, a.w >> 16};
      const uint32_t R[4] = {b.x & 0xFFFFu, b.x >> 16, b.y & 0xFFFFu, b.y >> 16};
      const uint32_t W[4] = {a.x & 0xFFFFu, a.x >> 16, a.y & 0xFFFFu, a.y >> 16};
#pragma unroll
      for (int i = 0; i < 4; ++i)
        if (F[i] | R[i]) { e.sc += F[i] + R[i]; e.w += W[i]; if (i == 1 || i == 2) e.fl |= 2u; if (i == 2) e.fl |= 4u; }
    }
  }
  return e;
}

__device__ __forceinline__ uint32_t region_of(const SiteMap &M, uint64_t i) {      // last region with regFirst <= i
  uint32_t lo = 0, hi = M.nRegions;
  if (M.tileReg) { lo = M.tileReg[i / kCursorTile]; hi = M.tileReg[i / kCursorTile + 1] + 1; }   // the cursor confines the search to the site's tile
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (M.regFirst[mid] <= i) lo = mid; else hi = mid; }
  return lo;
}

__device__ __forceinline__ int warp_sum(int v) { return __reduce_add_sync(0xffffffffu, v); }   // one REDUX.SUM

// Sums of one 100-entry window of the batch's mwInfo array starting at entry k0 (SNP_Caller.cpp:286-314), by a whole
// warp: lane l owns the four CONSECUTIVE entries 4 l .. 4 l + 3 (entries 0..100 are needed: the 101st only as the
// neighbour of the 100th), so the |strand-count difference| of neighbours is an in-lane affair but for one shuffle, the
// sums are five REDUX and nothing goes through shared memory.  The positions are resolved first, then all the lane's loads
// are issued together (invalid words + counters: eight independent gathers in flight), then decoded by win_decode.
// gflags (per lane, not reduced): bit s = entry 4 l + s is valid and saw a G.
struct WinSums { int strand, delta, depth, gc, valid; uint32_t gflags; };
__device__ __forceinline__ WinSums window_sums(const SnapView &V, const DevSiteMap &M, const BatchGeom &B, uint32_t k0,
                                               uint32_t kLo, uint32_t kHi, bool multi, uint32_t lane) {
  uint32_t pp[4], inv = 0; bool ex[4];                        // inv: bit 31 - s = entry 4 l + s sits on an invalid position
  uint32_t iw[4] = {0u, 0u, 0u, 0u}; uint64_t cc[4];
  if (!multi && k0 >= kLo && k0 + 104u <= kHi) {
    // the common case -- one region, window away from the batch's ends: 101 consecutive positions, all of them there;
    // the lane's four invalid bits come out of two bitmap words with one funnel shift
    const uint32_t p0 = B.amb0 - (uint32_t) kWindow + k0 + 4u * lane;
#pragma unroll
    for (int it = 0; it < 4; ++it) { pp[it] = p0 + it; ex[it] = 4u * lane + it <= (uint32_t) kWindow; }
    if (V.invalid && ex[0]) inv = __funnelshift_l(__ldg(V.invalid + (p0 >> 5) + 1), __ldg(V.invalid + (p0 >> 5)), p0 & 31u);
#pragma unroll
    for (int it = 0; it < 4; ++it)
      cc[it] = ex[it] ? __ldg(reinterpret_cast<const unsigned long long *>(V.counters) + pp[it]) : 0ull;
  } else {
    uint32_t reg = 0;                                         // region of the first in-batch entry (exome): a lower bound
    if (multi) reg = region_of(M.m, B.first + (k0 >= (uint32_t) kWindow ? k0 - kWindow : 0u));
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const uint32_t t = 4u * lane + it, k = k0 + t;
      ex[it] = t <= (uint32_t) kWindow && k >= kLo && k < kHi;
      pp[it] = B.amb0 - (uint32_t) kWindow + k;
      if (multi && ex[it] && k >= (uint32_t) kWindow) {
        if (k < (uint32_t) kWindow + B.n) {
          const uint64_t q = B.first + (k - (uint32_t) kWindow);
          while (reg + 1 < M.m.nRegions && M.m.regFirst[reg + 1] <= q) ++reg;
          pp[it] = M.m.regStart[reg] + (uint32_t) (q - M.m.regFirst[reg]);
        } else {
          pp[it] = B.ambLast + 1u + (k - (uint32_t) kWindow - B.n);
        }
      }
      ex[it] = ex[it] && pp[it] < V.L;
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      iw[it] = (ex[it] && V.invalid) ? __ldg(V.invalid + (pp[it] >> 5)) : 0u;
      cc[it] = ex[it] ? __ldg(reinterpret_cast<const unsigned long long *>(V.counters) + pp[it]) : 0ull;
    }
#pragma unroll
    for (int it = 0; it < 4; ++it) inv |= ((iw[it] >> (31 - (pp[it] & 31))) & 1u) << (31 - it);
  }
  WinLite e[4];
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    e[it].sc = e[it].w = e[it].fl = 0;
    if (ex[it] && !((inv >> (31 - it)) & 1u)) e[it] = win_decode(V, pp[it], cc[it]);
  }
  // the neighbour of the lane's last entry is the next lane's first one (strand counts stay below 2^31)
  const uint32_t nextLane = __shfl_down_sync(0xffffffffu, e[0].sc | (e[0].fl << 31), 1);
  int a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0;
  WinSums r; r.gflags = 0;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const uint32_t nsc = it < 3 ? e[it < 3 ? it + 1 : 3].sc : (nextLane & 0x7FFFFFFFu);
    const bool nvalid = it < 3 ? (e[it < 3 ? it + 1 : 3].fl & 1u) != 0 : (nextLane >> 31) != 0;
    if ((e[it].fl & 5u) == 5u) r.gflags |= 1u << it;
    if ((e[it].fl & 1u) && 4u * lane + it < (uint32_t) kWindow) {
      const uint32_t sc = e[it].sc;
      a0 += (int) sc; a1 += (int) e[it].w; a2 += (int) ((e[it].fl >> 1) & 1u); a3 += 1;
      if (nvalid) a4 += (int) (sc > nsc ? sc - nsc : nsc - sc);
    }
  }
  r.strand = warp_sum(a0); r.depth = warp_sum(a1); r.gc = warp_sum(a2) & 0xFF;      // gcSum / validSum are unsigned char
  r.valid = warp_sum(a3) & 0xFF; r.delta = warp_sum(a4);
  return r;
}

// rare paths of the record kernel, kept out of line so that the loop the common record walks stays small
__device__ __noinline__ uint32_t left_g_near_batch_start(const SnapView &V, const DevSiteMap &M, const BatchGeom &B, uint32_t idx,
                                                         uint32_t lane) {
  bool gbit = false;
  if (lane < 3) { WinEntry e = window_entry(V, M.m, B, (uint64_t) idx - 3 + lane); gbit = e.valid && e.g; }
  return __popc(__ballot_sync(0xffffffffu, gbit));
}
__device__ __noinline__ void write_indel_patterns(const SnapView &V, gt16_called_site &o, const PatRef &d0, const PatRef &d1,
                                                  uint32_t isIndel, uint32_t lane) {
  for (int k = 0; k < 2; ++k) {                                         // indel pattern bytes, two per lane
    gt16_indel &d = o.indel[k];
    const PatRef &pr = k ? d1 : d0;
    const bool have = k == 0 || isIndel == 3;
    const uint32_t nb = ((uint32_t) pr.len + 3) / 4;
    for (uint32_t bIdx = lane; bIdx < 64; bIdx += 32) {
      uint8_t v = 0;
      if (have && bIdx < nb) {
        if (pr.type == 'I') v = pr.seq[bIdx];
        else for (uint32_t q = bIdx * 4; q < bIdx * 4 + 4 && q < pr.len; ++q) v |= (uint8_t) (ref_base(V, pr.delPos + q) << ((q & 3) << 1));
      }
      d.seq[bIdx] = v;
    }
    if (lane == 0) { d.type = have ? pr.type : (uint8_t) '*'; d.len = have ? pr.len : 0; d.hq = have ? pr.hq : 0; d.lq = have ? pr.lq : 0; }
  }
}

// ---- Fisher tests of the records (strand bias / base-quality bias, SNPFunctions.cpp:137-159) -------------------
// SB and BQB for the four bases of the site itself (SNP rows) or of the three sites {index-1, index+nextBp, index}
// (indel rows, SNP_Caller.cpp:248-274).
// A test is a pure function of its three counts, and a genome repeats the same (F, R) and (W, F+R) pairs
// endlessly, so finished tests are memoised in a direct-mapped cache in global memory (L2-resident, owned by the
// context, kept across calls): {key, result} in one aligned 8-byte word.  A hit returns the very value a walk
// would produce; a miss runs the walk and stores it.  Keys that do not fit 32 bits (W >= 32768 or F+R >= 65536)
// are simply not cached.
__device__ __forceinline__ uint32_t bias_key(bool bqb, uint32_t W, uint32_t F, uint32_t R) {
  if (bqb) { const uint32_t S = F + R; return (W < 32768u && S < 65536u && !(W == 32767u && S == 65535u)) ? (0x80000000u | (W << 16) | S) : 0xFFFFFFFFu; }
  return (F < 32768u && R < 65536u) ? ((F << 16) | R) : 0xFFFFFFFFu;
}
__device__ __forceinline__ uint32_t bias_slot(uint32_t key, uint32_t mask) { return ((key * 0x9E3779B1u) ^ (key >> 15)) & mask; }

// A test the cache does not hold goes to the walk kernel (k_bias) through a list in global memory.
struct BiasList { BiasTask *tasks; uint32_t *count; uint32_t cap; };

// the list is full (it holds one test per record of the bound, far more than a warmed-up cache ever misses): the lane
// walks the test itself
__device__ __noinline__ float bias_inline(const LnFact &G, bool bqb, uint32_t W, uint32_t F, uint32_t R) {
  LnFact L; L.tab = G.tab; L.size = G.size; L.far = G.tab; L.farSize = G.size;
  return bqb ? base_quality_bias(L, W, F, R) : strand_bias(L, F, R);
}

__global__ void __launch_bounds__(kCandThreads, GT16_CAND_BLOCKS) k_candidates(SnapView V, DevSiteMap M, uint64_t base,
                                                             const EvalOut *__restrict__ part,
                                                             const uint32_t *__restrict__ entryOf,
                                                             const uint32_t *__restrict__ count, uint32_t cap,
                                                             LnFact G, uint2 *__restrict__ cache, uint32_t cacheMask, BiasList missed,
                                                             float sbEmpty, float bqbEmpty,
                                                             gt16_called_site *__restrict__ out, uint32_t *errWord) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warpsTotal = gridDim.x * (kCandThreads / 32);
  // count[0]: side-buffer entries the evaluator reserved, count[1]: candidates (the scan's total).  More entries than the
  // buffers hold: some candidates were dropped, the call is going to be repeated with larger buffers.
  if (count[0] > cap) return;
  const uint32_t n = count[1];
  for (uint32_t rk = blockIdx.x * (kCandThreads / 32) + (threadIdx.x >> 5); rk < n; rk += warpsTotal) {
    const EvalOut &eo = part[entryOf[rk]];
    const uint64_t i = base + eo.site;
    const uint32_t p = site_pos(M, i);
    gt16_called_site &o = out[rk];
    uint32_t err = 0;
    // -- the likelihoods, posteriors and the call itself come from the exact evaluator (k_eval_mask), which computed them
    //    to decide that this site is a candidate
    if (lane < 16) { o.lh[lane] = eo.lh[lane]; o.flh[lane] = eo.flh[lane]; }
    const int genotype = eo.genotype;
    const int ref = (int) ref_base(V, p);
    IndelInfo ii; PatRef d0, d1;
    decode_indel_refs(V, p, ii, d0, d1);
    if (ii.bad) err |= kErrBadSnapshot;
    const uint8_t isIndel = eo.isIndel, nextBp = eo.nextBp;             // SNP_Caller.cpp:228-239 (decided by the evaluator)
    if (genotype > 9 && !isIndel) err |= kErrGenotype;
    BatchGeom B;
    B.first = (i / kBatch) * (uint64_t) kBatch;
    const uint64_t left = M.totalSites - B.first;
    B.n = left < kBatch ? (uint32_t) left : kBatch;
    B.amb0 = site_pos(M, B.first);
    B.ambLast = site_pos(M, B.first + B.n - 1);
    const uint32_t idx = (uint32_t) (i - B.first);
    // -- counts and Fisher tests: lane 4 k + b holds base b of the k-th site whose strand / base-quality bias the row
    //    carries -- the site itself (SNP rows, k = 0) or {index-1, index+nextBp, index} (indel rows) :248-274 -- decoded
    //    one base per lane (decode_bases); the site's own counts go into the record, and every lane looks its two
    //    tests up in the memo cache: a hit is written at once, a miss is left to the walk kernel.
    {
      const uint32_t k = lane >> 2, b = lane & 3u;
      bool have = false; uint64_t si = i;
      if (lane < 12) {
        if (isIndel) {
          if (k == 0) { have = idx != 0; si = i - 1; }
          else if (k == 1) { have = idx + nextBp < B.n; si = i + nextBp; }
          else have = true;
        } else have = k == 0;
      }
      const bool own = have && si == i;
      uint32_t w = 0, f = 0, r = 0;
      bool bad = false;
      if (have) {
        const uint32_t ps = own ? p : site_pos(M, si);
        if (!is_invalid(V, ps)) {
          const uint64_t c = V.counters[ps];
          const uint32_t lo = (uint32_t) c, hi = (uint32_t) (c >> 32);
          if (!is_overflow_type(c)) {
            if ((hi & 0xFFFFu) && ((lo >> 14) & 3u) == b) { w = lo & 0x3FFFu; f = hi & 0xFFu; r = (hi >> 8) & 0xFFu; }
            if ((hi >> 16) && (lo >> 30) == b) { w = (lo >> 16) & 0x3FFFu; f = (hi >> 16) & 0xFFu; r = hi >> 24; }   // second slot wins
          } else {
            const uint8_t *rec = ovf_record(V, ovf_region(V, ps), hi);
            if (!rec) bad = own;                              // a neighbour with a bad index counts as empty, like decode_bases
            else {
              f = rd16(rec + 8 + 2 * b); r = rd16(rec + 16 + 2 * b);
              if (f | r) w = rd16(rec + 2 * b);
            }
          }
        }
      }
      if (__any_sync(0xffffffffu, bad)) err |= kErrBadSnapshot;
      if (own) { o.W[b] = w; o.F[b] = f; o.R[b] = r; }
      if (lane == 12) { o.W[4] = ii.W4; o.F[4] = ii.F4; o.R[4] = ii.R4; }
      if (lane == 13) { o.W[5] = ii.W5; o.F[5] = ii.F5; o.R[5] = ii.R5; }
      if (lane < 12) {
        float sbv = 0.f, bqv = 0.f;
        bool needS = false, needB = false;
        if (have) {
          sbv = sbEmpty; bqv = bqbEmpty;
          if (w | f | r) {
            const uint32_t kS = bias_key(false, w, f, r), kB = bias_key(true, w, f, r);
            unsigned long long eS = ~0ull, eB = ~0ull;
            if (kS != 0xFFFFFFFFu) eS = *reinterpret_cast<const volatile unsigned long long *>(cache + bias_slot(kS, cacheMask));
            if (kB != 0xFFFFFFFFu) eB = *reinterpret_cast<const volatile unsigned long long *>(cache + bias_slot(kB, cacheMask));
            needS = !(kS != 0xFFFFFFFFu && (uint32_t) eS == kS);
            needB = !(kB != 0xFFFFFFFFu && (uint32_t) eB == kB);
            sbv = __uint_as_float((uint32_t) (eS >> 32)); bqv = __uint_as_float((uint32_t) (eB >> 32));
          }
        }
        if (needS | needB) {                                  // rare once the cache is warm
          const uint32_t dst = rk * 12u + lane;
          const uint32_t want = (needS ? 1u : 0u) + (needB ? 1u : 0u);
          const uint32_t at = atomicAdd(missed.count, want);
          if (at + want <= missed.cap) {
            uint32_t q = at;
            if (needS) { BiasTask t; t.W = w; t.F = f; t.R = r; t.dst = dst; missed.tasks[q++] = t; }
            if (needB) { BiasTask t; t.W = w; t.F = f; t.R = r; t.dst = dst | 0x80000000u; missed.tasks[q] = t; }
          } else {
            atomicSub(missed.count, want);
            if (needS) { sbv = bias_inline(G, false, w, f, r); needS = false; }
            if (needB) { bqv = bias_inline(G, true, w, f, r); needB = false; }
          }
        }
        if (!needS) o.sb[k][b] = sbv;
        if (!needB) o.bqb[k][b] = bqv;
      }
    }


    // -- windows (reference's un-offset mwInfo indexing)                                 :276-367
    // entry k of the batch's mwInfo array (k < n + 200; the cases of window_entry()): the left margin [0, 100) holds
    // the positions before the batch's first site -- filled only when amb0 > 100 (VariantCaller.cpp:713-729) --, then
    // the n sites of the batch, then 99 positions after its last site as far as the genome goes (:731-747).  In a
    // single-region call all three are one run of consecutive positions.
    const bool haveLeft = idx >= (uint32_t) kWindow;
    const bool multi = M.m.nRegions > 1;
    const uint32_t kLo = B.amb0 > (uint32_t) kWindow ? 0u : (uint32_t) kWindow;
    const uint32_t past = V.L - 1u - B.ambLast;
    const uint32_t kHi = (uint32_t) kWindow + B.n + (past < (uint32_t) kWindow - 1u ? past : (uint32_t) kWindow - 1u);
    WinSums wl, wr;
    wl.strand = wl.delta = wl.depth = wl.gc = wl.valid = 0; wl.gflags = 0;
    wr = wl;
#pragma unroll 1
    for (int side = haveLeft ? 0 : 1; side < 2; ++side) {              // one copy of the window code (instruction cache)
      const WinSums w = window_sums(V, M, B, side ? idx + nextBp : idx - (uint32_t) kWindow, kLo, kHi, multi, lane);
      if (side) wr = w; else wl = w;
    }
    // G counts of the three sites before / after (:355-367): mwInfo entries idx-3..idx-1 and idx+nextBp..+2, i.e.
    // entries 97..99 of the left window (lane 24, slots 1..3) and 0..2 of the right one (lane 0, slots 0..2)
    uint32_t lg = __popc(__shfl_sync(0xffffffffu, wl.gflags, 24) & 0xEu);
    const uint32_t rg = __popc(__shfl_sync(0xffffffffu, wr.gflags, 0) & 0x7u);
    if (!haveLeft && idx >= 3) lg = left_g_near_batch_start(V, M, B, idx, lane);   // the first 100 sites of a batch: no left window
    const uint32_t run = polyrun(V, p);

    // -- record
    static_assert(offsetof(gt16_called_site, indel) % 4 == 0 && sizeof(gt16_indel) == 70, "indel patterns are written as 35 words");
    if (!isIndel) {
      // two empty patterns {'*', 0, 0, 0, zeros}: byte 0 and byte 70 of the 140 hold '*'
      uint32_t *iw = reinterpret_cast<uint32_t *>(o.indel);
      iw[lane] = lane == 0 ? 0x2Au : (lane == 17 ? 0x2A0000u : 0u);
      if (lane < 3) iw[32 + lane] = 0u;
    } else {
      write_indel_patterns(V, o, d0, d1, isIndel, lane);
    }
    // window features (:316-353): ten float divisions, spread over the lanes instead of lined up on one.  Round 1, lanes
    // 0..6: rpb, lgc, avgSC[0], avgD[0], rgc, avgSC[1], avgD[1] -- (int + 0.0f) / int, IEEE division, the same bits on
    // any lane; round 2, lanes 0..2: asc, lq, rq.  rpb/lgc/rgc and asc/lq/rq are six consecutive floats of the record.
    {
      const bool right = lane >= 4;
      const int sel = right ? (int) lane - 3 : (int) lane;             // sel 1..3: gc, strand, depth
      int num = right ? wr.gc : wl.gc, den = right ? wr.valid : wl.valid;
      if (sel == 2) num = right ? wr.strand : wl.strand;
      if (sel == 3) num = right ? wr.depth : wl.depth;
      if (lane == 0) { num = wl.delta - wr.delta; den = wl.delta + wr.delta; }
      float q1 = 0.0f;
      if (lane < 7 && den != 0) q1 = (num + 0.0f) / den;
      const float avgSC0 = __shfl_sync(0xffffffffu, q1, 2), avgD0 = __shfl_sync(0xffffffffu, q1, 3);
      const float avgSC1 = __shfl_sync(0xffffffffu, q1, 5), avgD1 = __shfl_sync(0xffffffffu, q1, 6);
      const bool assigned = fabsf(avgSC0 + avgSC1) > FLT_EPSILON;
      float a = avgSC0 - avgSC1, bden = avgSC0 + avgSC1;
      bool ok = assigned;
      if (lane == 1) { a = avgD0; bden = avgSC0; ok = assigned && fabsf(avgSC0) > FLT_EPSILON; }
      if (lane == 2) { a = avgD1; bden = avgSC1; ok = assigned && fabsf(avgSC1) > FLT_EPSILON; }
      float q2 = 0.0f;
      if (lane < 3 && ok) q2 = a / bden;
      float *feat = &o.rpb;                                             // rpb, lgc, rgc, asc, lq, rq
      if (lane == 0) { feat[0] = q1; feat[3] = q2; }
      if (lane == 1) { feat[1] = q1; feat[4] = q2; }
      if (lane == 2) feat[5] = q2;
      if (lane == 4) feat[2] = q1;
      if (lane == 0) {
        o.amb = p; o.ref = (uint8_t) ref; o.genotype = (uint8_t) genotype; o.is_indel = isIndel;
        o.rq_assigned = assigned ? 1 : 0;
        o.pD = eo.pD; o.bestD = eo.bestD; o.secondD = eo.secondD;
        o.lg = (uint8_t) lg; o.rg = (uint8_t) rg;
        o.polyrun = (uint16_t) run;
        o.total_hq = ii.hqTotal; o.total_lq = ii.lqTotal;
        if (err) atomicOr(errWord, err);
      }
    }
  }
}
void launch_candidates(const SnapView &V, const DevSiteMap &M, uint64_t base, const EvalOut *part, const uint32_t *entryOf,
                       const uint32_t *count, uint32_t cap, const EvalParams &P, uint2 *cache, uint32_t cacheMask,
                       BiasTask *missed, uint32_t *missedCount, uint32_t missedCap, gt16_called_site *out, uint32_t *errWord,
                       int numSMs, cudaStream_t s) {
  // persistent grid: the number of records is only known on the device; 8 blocks of 4 warps per SM (launch bounds), and a
  // small call (cap is a bound on its records) does not need more warps than records
  unsigned grid = (cap + (kCandThreads / 32) - 1) / (kCandThreads / 32);
  const unsigned full = (unsigned) numSMs * GT16_CAND_GRID;
  if (grid > full) grid = full;
  BiasList ml; ml.tasks = missed; ml.count = missedCount; ml.cap = missedCap;
  // the tests of a base that was never seen (all counts zero) need lnfactC(0) only: evaluated here, by the same routines,
  // with the host libm's lgammaf(1) the device table starts with
  static const float lg1 = lgammaf(1.0f);
  LnFact H; H.tab = &lg1; H.size = 1; H.far = nullptr; H.farSize = 0;
  const float sbEmpty = strand_bias(H, 0, 0), bqbEmpty = base_quality_bias(H, 0, 0, 0);
  k_candidates<<<grid, kCandThreads, 0, s>>>(V, M, base, part, entryOf, count, cap, P.lnfact, cache, cacheMask, ml, sbEmpty, bqbEmpty, out, errWord);
}

// ------------------------------------------------------------------------------ Fisher tests of the candidates
// Strand bias and base-quality bias (computeStrandBias / computeBaseQualityBias, SNPFunctions.cpp:137-159)
// of every called site: SB and BQB for the four bases of the site itself (SNP rows) or of the three sites
// {index-1, index+nextBp, index} (indel rows, SNP_Caller.cpp:248-274).
//
// A test (fisher2tailC, SNPFunctions.cpp:94-135) is a serial float recurrence -- a left and a right walk
// over the hypergeometric support, a few to several hundred steps -- that cannot be split or reordered
// without changing its rounding, and most (site, base) slots are empty (a base that was never seen).  A
// slot-per-lane mapping leaves ~80 % of the lanes idle, so the kernel runs the walks as per-lane state
// machines: a warp decodes 32 slots at a time and queues the non-empty tests in shared memory; every lane
// owns one test at a time and advances it by one step per iteration of a warp-uniform loop (both walks
// share one step body: the left walk's single factor is multiplied by an exact 1.0f), and a lane that
// finishes takes the next queued test at the next refill point.  lnfactC(n) = lgammaf(n+1) for
// n < kLnSmem is staged in shared memory.
constexpr int kBiasThreads = 256;
constexpr int kLnSmem = 4096;
constexpr int kBiasQueue = 128;
constexpr int kBiasStepsPerRefill = 12;

struct FisherWalk {                               // state of one fisher2tailC evaluation
  int n1_, n_1, n, x, lo, hi, n11;
  float pox, poa, poa_e, two;
  uint32_t dst, key;                              // key: cache key of the running test (0xFFFFFFFF = not cacheable)
  int phase;                                      // 0 idle, 1 left walk, 2 right walk
  bool stopAtZero;
};

// fisher2tailC(n11, n1_, n_1, n) up to the start of the left walk (same expressions, same order)
__device__ __forceinline__ void walk_begin(const LnFact &L, FisherWalk &w, int n11, int n1_, int n_1, int n, uint32_t dst) {
  n += n11 + n1_ + n_1;
  n1_ += n11;
  n_1 += n11;
  const float poa = expf_glibc(lnbicoC(L, n1_, n11) + lnbicoC(L, n - n1_, n_1 - n11) - lnbicoC(L, n, n_1));
  w.n11 = n11; w.n1_ = n1_; w.n_1 = n_1; w.n = n; w.dst = dst;
  w.poa = poa; w.pox = poa; w.two = 0; w.two += poa; w.poa_e = poa * 1.000000001f;
  w.stopAtZero = n <= 46340;
  w.lo = (n1_ + n_1 - n > 0) ? n1_ + n_1 - n : 0;
  w.hi = (n1_ < n_1) ? n1_ : n_1;
  w.x = n11;
  w.phase = 1;
  if (!(w.x > w.lo)) { w.phase = 2; w.x = n11 + 1; w.pox = poa; if (!(w.x <= w.hi)) w.phase = 3; }
}
// one step of whichever walk the lane is in; phase 3 = finished (result in w.two)
__device__ __forceinline__ void walk_step(FisherWalk &w) {
  if (w.phase != 1 && w.phase != 2) return;
  const bool left = w.phase == 1;
  const int x = w.x;
  const int a = left ? (int) ((uint32_t) x * (uint32_t) (w.n + x - w.n1_ - w.n_1)) : w.n1_ - (x - 1);
  const int b = left ? 1 : w.n_1 - (x - 1);
  const int den = left ? (int) ((uint32_t) (w.n1_ - x + 1) * (uint32_t) (w.n_1 - x + 1))
                       : (int) ((uint32_t) (w.n + x - w.n1_ - w.n_1) * (uint32_t) x);
  float pox = w.pox * (float) a;
  pox = pox * (float) b;                          // left walk: times 1.0f, exact
  pox = pox / (float) den;
  w.pox = pox;
  w.two += pox * (pox < w.poa_e);
  const bool zero = w.stopAtZero && pox == 0.0f;
  if (left) {
    w.x = x - 1;
    if (zero || !(w.x > w.lo)) { w.phase = 2; w.x = w.n11 + 1; w.pox = w.poa; if (!(w.x <= w.hi)) w.phase = 3; }
  } else {
    w.x = x + 1;
    if (zero || !(w.x <= w.hi)) w.phase = 3;
  }
}

__global__ void __launch_bounds__(kBiasThreads) k_bias(BiasList missed, LnFact G, uint2 *__restrict__ cache, uint32_t cacheMask,
                                                       gt16_called_site *__restrict__ out) {
  const uint32_t n = *missed.count < missed.cap ? *missed.count : missed.cap;
  if (blockIdx.x * (kBiasThreads / 32) * 32u >= n) return;          // persistent grid sized for the bound, not the count
  __shared__ float lnS[kLnSmem];
  __shared__ BiasTask queue[kBiasThreads / 32][kBiasQueue];
  const uint32_t nS = G.size < (uint32_t) kLnSmem ? G.size : (uint32_t) kLnSmem;
  for (uint32_t k = threadIdx.x; k < nS; k += kBiasThreads) lnS[k] = G.tab[k];
  __syncthreads();
  LnFact L;
  L.tab = lnS; L.size = nS; L.far = G.tab; L.farSize = G.size;
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5, ltMask = (1u << lane) - 1u;
  BiasTask *q = queue[warp];
  uint32_t qhead = 0, qtail = 0;                   // ring indices (monotonic), entries live in q[i % kBiasQueue]
  FisherWalk w;
  w.phase = 0;

  auto store_result = [&]() {                      // min(1, two) into the record (computeStrandBias / ...BaseQualityBias)
    const uint32_t d = w.dst & 0x7FFFFFFFu;
    gt16_called_site &o = out[d / 12];
    const uint32_t k = (d % 12) >> 2, b = d & 3;
    const float v = fminf(1.0f, w.two);
    if (w.dst >> 31) o.bqb[k][b] = v; else o.sb[k][b] = v;
    if (w.key != 0xFFFFFFFFu) {                      // memoise (8-byte store: key and value land together)
      *reinterpret_cast<volatile unsigned long long *>(cache + bias_slot(w.key, cacheMask)) =
          (unsigned long long) w.key | ((unsigned long long) __float_as_uint(v) << 32);
    }
  };
  auto refill = [&]() {                            // finished lanes store; idle lanes take queued tests
    if (w.phase == 3) { store_result(); w.phase = 0; }
    const uint32_t idle = __ballot_sync(0xffffffffu, w.phase == 0);
    const uint32_t avail = qtail - qhead;
    if (idle && avail) {
      const uint32_t rank = __popc(idle & ltMask);
      if (w.phase == 0 && rank < avail) {
        const BiasTask t = q[(qhead + rank) % kBiasQueue];
        w.key = bias_key(t.dst >> 31, t.W, t.F, t.R);
        // the same test may have been finished by another lane since the record kernel looked (a cold cache lists every
        // test of the call, most of them many times over): look again before walking
        unsigned long long e = ~0ull;
        if (w.key != 0xFFFFFFFFu) e = *reinterpret_cast<const volatile unsigned long long *>(cache + bias_slot(w.key, cacheMask));
        if (w.key != 0xFFFFFFFFu && (uint32_t) e == w.key) {
          const uint32_t d = t.dst & 0x7FFFFFFFu;
          gt16_called_site &o = out[d / 12];
          const float v = __uint_as_float((uint32_t) (e >> 32));
          if (t.dst >> 31) o.bqb[(d % 12) >> 2][d & 3] = v; else o.sb[(d % 12) >> 2][d & 3] = v;
        } else if (t.dst >> 31) { const int S = (int) ((t.F + t.R) * 4); walk_begin(L, w, (int) t.W, S, S, S, t.dst); }      // getBaseQualityBias
        else { const int S = (int) (t.F + t.R); walk_begin(L, w, (int) t.F, (int) t.R, (int) (S * 0.5), (int) (S * 0.5), t.dst); }  // getStrandBiasC
      }
      const uint32_t taken = __popc(idle) < avail ? __popc(idle) : avail;
      qhead += taken;
    }
    __syncwarp();
  };
  auto push = [&](bool want, const BiasTask &t) {
    const uint32_t bal = __ballot_sync(0xffffffffu, want);
    if (want) q[(qtail + __popc(bal & ltMask)) % kBiasQueue] = t;
    qtail += __popc(bal);
    __syncwarp();
  };

  const uint32_t warpsTotal = gridDim.x * (kBiasThreads / 32);
  for (uint32_t chunk = blockIdx.x * (kBiasThreads / 32) + warp; chunk * 32u < n; chunk += warpsTotal) {
    const uint32_t t = chunk * 32u + lane;
    BiasTask task; task.W = task.F = task.R = task.dst = 0;
    if (t < n) task = missed.tasks[t];
    push(t < n, task);
    // keep room for the next chunk's tests
    while (qtail - qhead > (uint32_t) (kBiasQueue - 32)) {
      refill();
#pragma unroll 1
      for (int it = 0; it < kBiasStepsPerRefill; ++it) walk_step(w);
    }
  }
  while (true) {                                    // drain
    refill();
    if (!__any_sync(0xffffffffu, w.phase != 0)) break;
#pragma unroll 1
    for (int it = 0; it < kBiasStepsPerRefill; ++it) walk_step(w);
  }
}
void launch_bias(BiasTask *missed, const uint32_t *missedCount, uint32_t missedCap, const EvalParams &P, uint2 *cache,
                 uint32_t cacheMask, gt16_called_site *out, int numSMs, cudaStream_t s) {
  // the walks are long dependent chains: what matters is how many run at once, so spread the chunks over as many
  // resident warps as the device holds (4 blocks of 8 warps per SM); blocks beyond the device-side count leave at once
  const uint32_t chunks = (missedCap + 31) / 32, perBlock = kBiasThreads / 32;
  unsigned grid = (chunks + perBlock - 1) / perBlock;
  if (grid > (unsigned) numSMs * GT16_BIAS_GRID) grid = (unsigned) numSMs * GT16_BIAS_GRID;
  if (grid < 1) grid = 1;
  BiasList ml; ml.tasks = missed; ml.count = const_cast<uint32_t *>(missedCount); ml.cap = missedCap;
  k_bias<<<grid, kBiasThreads, 0, s>>>(ml, P.lnfact, cache, cacheMask, out);
}

// ------------------------------------------------------------------------------ operator seam
__global__ void __launch_bounds__(128) k_op_likelihood(EvalParams P, const uint8_t *__restrict__ w, size_t wStride,
                                                       const uint8_t *__restrict__ ref, size_t refStride,
                                                       double *__restrict__ lh, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t W[6];
  const uint32_t *src = (const uint32_t *) (w + (size_t) i * wStride);
  for (int k = 0; k < 6; ++k) W[k] = src[k];
  double v[16];
  genotype_likelihood16(P.lh, W, (int) (int8_t) ref[(size_t) i * refStride], v);
  for (int g = 0; g < 16; ++g) lh[(size_t) i * 16 + g] = v[g];
}
void launch_op_likelihood(const EvalParams &P, const uint8_t *w, size_t wStride, const uint8_t *ref,
                          size_t refStride, double *lh, uint32_t n, cudaStream_t s) {
  if (n) k_op_likelihood<<<(n + 127) / 128, 128, 0, s>>>(P, w, wStride, ref, refStride, lh, n);
}

__global__ void __launch_bounds__(128) k_op_genotype(EvalParams P, const uint8_t *__restrict__ ref,
                                                     const double *__restrict__ lh, uint8_t *__restrict__ info,
                                                     size_t infoStride, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v[16];
  for (int g = 0; g < 16; ++g) v[g] = lh[(size_t) i * 16 + g];
  Posterior po;
  select_genotype((int) ref[i], v, P.prior, P.g2i, po);
  uint8_t *d = info + (size_t) i * infoStride;          // SnpCallingInfo, SNP_Meta.h:115-125
  *(double *) (d + 0) = po.pD;
  *(int *) (d + 8) = po.genotype;
  for (int g = 0; g < 16; ++g) *(double *) (d + 16 + 8 * g) = po.flh[g];
  *(double *) (d + 144) = po.bestD;
  *(double *) (d + 152) = po.secondD;
}
void launch_op_genotype(const EvalParams &P, const uint8_t *ref, const double *lh, uint8_t *info,
                        size_t infoStride, uint32_t n, cudaStream_t s) {
  if (n) k_op_genotype<<<(n + 127) / 128, 128, 0, s>>>(P, ref, lh, info, infoStride, n);
}

__global__ void __launch_bounds__(128) k_op_bias(EvalParams P, const uint8_t *__restrict__ w,
                                                 const uint8_t *__restrict__ f, const uint8_t *__restrict__ r,
                                                 size_t stride, uint32_t n, float *__restrict__ sb,
                                                 float *__restrict__ bqb) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;     // one thread per (site, base)
  if (t >= n * 4) return;
  uint32_t i = t >> 2, b = t & 3;
  uint32_t Wv = ((const uint32_t *) (w + (size_t) i * stride))[b];
  uint32_t Fv = ((const uint32_t *) (f + (size_t) i * stride))[b];
  uint32_t Rv = ((const uint32_t *) (r + (size_t) i * stride))[b];
  sb[t] = strand_bias(P.lnfact, Fv, Rv);
  bqb[t] = base_quality_bias(P.lnfact, Wv, Fv, Rv);
}
void launch_op_bias(const EvalParams &P, const uint8_t *w, const uint8_t *f, const uint8_t *r, size_t stride,
                    uint32_t n, float *sb, float *bqb, cudaStream_t s) {
  if (n) k_op_bias<<<(n * 4 + 127) / 128, 128, 0, s>>>(P, w, f, r, stride, n, sb, bqb);
}

}  // namespace gt16
