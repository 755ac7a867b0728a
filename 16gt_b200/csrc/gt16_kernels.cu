// gt16_kernels.cu -- sm_100a kernels of the 16GT calling hot path.
//
//   k_build_table / k_build_ycsq   error tables (replaces the ~70 s CPU prefill,
//                                  likelihood_cache.cpp:5-54)
//   k_build_decisions              exact memoised decision bitmaps, built by the same evaluator
//   k_stream                       HBM-bound pass over the counters: decode + classify + ballot
//   k_mask_count/scan/scatter      ordered compaction (ballot / prefix sums)
//   k_eval_list                    exact 16-genotype evaluation of the flagged sites
//   k_candidates                   full record (likelihoods, posterior, SB/BQB Fisher, windows)
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

template <bool kUseTables>
__global__ void __launch_bounds__(256) k_stream(SnapView V, DevSiteMap M, uint64_t first, uint32_t n,
                                                DecisionTables D, EvalParams P, uint32_t *__restrict__ mask,
                                                uint32_t *errWord) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  bool flag = false;
  if (i < n) {
    uint32_t p = site_pos(M, first + i);
    if (!kUseTables) {
      uint32_t e = 0;
      flag = site_is_candidate(V, P, p, e);
      if (e) atomicOr(errWord, e);
    } else if (!is_invalid(V, p)) {
      uint64_t c = __ldg(V.counters + p);
      uint32_t q = p + 1 < V.L ? p + 1 : V.L - 1;
      uint64_t cn = __ldg(V.counters + q);
      uint32_t ref = ref_base(V, p);
      bool nextIndel = is_overflow_type(cn) && (((uint32_t) cn | (uint32_t) (cn >> 16)) & 0x2000u);
      if (nextIndel) {
        flag = true;                                   // W[4] / W[5] may be non-zero: exact path
      } else {
        uint32_t wr[4] = {0, 0, 0, 0};                 // weights indexed by role = base ^ ref
        bool ok = true;
        if (!is_overflow_type(c)) {
          uint32_t w1 = (uint32_t) c & 0xFFFF, w2 = (uint32_t) (c >> 16) & 0xFFFF;
          if ((uint32_t) (c >> 32) & 0xFFFFu) wr[(w1 >> 14) ^ ref] = w1 & 0x3FFF;
          if ((uint32_t) (c >> 48) & 0xFFFFu) wr[(w2 >> 14) ^ ref] = w2 & 0x3FFF;
        } else {
          const uint8_t *r = ovf_record(V, ovf_region(V, p), (uint32_t) (c >> 32));
          if (!r) { atomicOr(errWord, kErrBadSnapshot); ok = false; }
          else {
            const uint4 lo = __ldg((const uint4 *) r);           // W[4], F[4]
            const uint2 hi = __ldg((const uint2 *) (r + 16));    // R[4]
            uint32_t Wv[4] = {lo.x & 0xFFFF, lo.x >> 16, lo.y & 0xFFFF, lo.y >> 16};
            uint32_t Fv[4] = {lo.z & 0xFFFF, lo.z >> 16, lo.w & 0xFFFF, lo.w >> 16};
            uint32_t Rv[4] = {hi.x & 0xFFFF, hi.x >> 16, hi.y & 0xFFFF, hi.y >> 16};
#pragma unroll
            for (int b = 0; b < 4; ++b) if (Fv[b] | Rv[b]) wr[b ^ ref] = Wv[b];
          }
        }
        if (ok) flag = lookup_needs_exact(D, wr[0], wr[2], wr[1], wr[3]);
      }
    }
  }
  uint32_t word = __ballot_sync(0xffffffffu, flag);
  if ((threadIdx.x & 31) == 0 && i < n) mask[i >> 5] = word;
}

void launch_stream(const SnapView &V, const DevSiteMap &M, uint64_t first, uint32_t n, bool useTables,
                   const DecisionTables &D, const EvalParams &P, uint32_t *mask, uint32_t *errWord,
                   cudaStream_t s) {
  if (!n) return;
  unsigned grid = (n + 255) / 256;
  if (useTables) k_stream<true><<<grid, 256, 0, s>>>(V, M, first, n, D, P, mask, errWord);
  else k_stream<false><<<grid, 256, 0, s>>>(V, M, first, n, D, P, mask, errWord);
}

__global__ void __launch_bounds__(128) k_eval_list(SnapView V, DevSiteMap M, uint64_t base,
                                                   const uint32_t *__restrict__ list, uint32_t n, EvalParams P,
                                                   uint32_t *__restrict__ mask, uint32_t *errWord) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  bool flag = false;
  if (j < n) {
    uint32_t e = 0;
    flag = site_is_candidate(V, P, site_pos(M, base + list[j]), e);
    if (e) atomicOr(errWord, e);
  }
  uint32_t word = __ballot_sync(0xffffffffu, flag);
  if ((threadIdx.x & 31) == 0 && j < n) mask[j >> 5] = word;
}
void launch_eval_list(const SnapView &V, const DevSiteMap &M, uint64_t base, const uint32_t *list, uint32_t n,
                      const EvalParams &P, uint32_t *mask, uint32_t *errWord, cudaStream_t s) {
  if (!n) return;
  k_eval_list<<<(n + 127) / 128, 128, 0, s>>>(V, M, base, list, n, P, mask, errWord);
}

// ------------------------------------------------------------------------------ ordered compaction
__global__ void __launch_bounds__(kScanBlock) k_mask_count(const uint32_t *__restrict__ mask, uint32_t nWords,
                                                          uint32_t *__restrict__ blockCounts) {
  __shared__ uint32_t warpSum[kScanBlock / 32];
  uint32_t w = blockIdx.x * kScanBlock + threadIdx.x;
  uint32_t c = w < nWords ? __popc(mask[w]) : 0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0) warpSum[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int k = 0; k < kScanBlock / 32; ++k) t += warpSum[k];
    blockCounts[blockIdx.x] = t;
  }
}
void launch_mask_count(const uint32_t *mask, uint32_t nItems, uint32_t *blockCounts, cudaStream_t s) {
  uint32_t nb = scan_blocks(nItems);
  if (nb) k_mask_count<<<nb, kScanBlock, 0, s>>>(mask, (nItems + 31) / 32, blockCounts);
}

// in-place exclusive scan of up to a few 10^5 block counts by one CTA; total -> *total
__global__ void __launch_bounds__(1024) k_scan_counts(uint32_t *counts, uint32_t n, uint32_t *total) {
  __shared__ uint32_t warpTot[32];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < n; base += 1024) {
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < n ? counts[i] : 0, x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
    if ((threadIdx.x & 31) == 31) warpTot[threadIdx.x >> 5] = x;
    __syncthreads();
    if (threadIdx.x < 32) {
      uint32_t w = warpTot[threadIdx.x], z = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, z, o); if (threadIdx.x >= o) z += y; }
      warpTot[threadIdx.x] = z - w;   // exclusive warp offsets
    }
    __syncthreads();
    uint32_t excl = carry + warpTot[threadIdx.x >> 5] + x - v;
    if (i < n) counts[i] = excl;
    __syncthreads();
    if (threadIdx.x == 1023) carry = excl + v;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
void launch_scan_counts(uint32_t *blockCounts, uint32_t nBlocks, uint32_t *total, cudaStream_t s) {
  k_scan_counts<<<1, 1024, 0, s>>>(blockCounts, nBlocks, total);
}

__global__ void __launch_bounds__(kScanBlock) k_mask_scatter(const uint32_t *__restrict__ mask, uint32_t nWords,
                                                            const uint32_t *__restrict__ blockOffsets,
                                                            const uint32_t *__restrict__ values, uint32_t offset,
                                                            uint32_t *__restrict__ out) {
  __shared__ uint32_t warpTot[kScanBlock / 32];
  uint32_t w = blockIdx.x * kScanBlock + threadIdx.x;
  uint32_t bits = w < nWords ? mask[w] : 0;
  uint32_t c = __popc(bits), x = c;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { uint32_t y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
  if ((threadIdx.x & 31) == 31) warpTot[threadIdx.x >> 5] = x;
  __syncthreads();
  uint32_t pre = 0;
  for (uint32_t k = 0; k < (threadIdx.x >> 5); ++k) pre += warpTot[k];
  uint32_t dst = blockOffsets[blockIdx.x] + pre + x - c;
  while (bits) {
    uint32_t b = __ffs(bits) - 1;
    bits &= bits - 1;
    uint32_t item = w * 32 + b;
    out[dst++] = values ? values[item] : item + offset;
  }
}
void launch_mask_scatter(const uint32_t *mask, uint32_t nItems, const uint32_t *blockOffsets,
                         const uint32_t *values, uint32_t offset, uint32_t *out, cudaStream_t s) {
  uint32_t nb = scan_blocks(nItems);
  if (nb) k_mask_scatter<<<nb, kScanBlock, 0, s>>>(mask, (nItems + 31) / 32, blockOffsets, values, offset, out);
}

// ------------------------------------------------------------------------------ candidates
__global__ void __launch_bounds__(64) k_candidates(SnapView V, DevSiteMap M, uint64_t base,
                                                   const uint32_t *__restrict__ list, uint32_t n, EvalParams P,
                                                   gt16_called_site *__restrict__ out, uint32_t *errWord) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  uint32_t e = 0;
  build_called_site(V, M, base + list[j], P, out[j], e);
  if (e) atomicOr(errWord, e);
}
void launch_candidates(const SnapView &V, const DevSiteMap &M, uint64_t base, const uint32_t *list, uint32_t n,
                       const EvalParams &P, gt16_called_site *out, uint32_t *errWord, cudaStream_t s) {
  if (!n) return;
  k_candidates<<<(n + 63) / 64, 64, 0, s>>>(V, M, base, list, n, P, out, errWord);
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
