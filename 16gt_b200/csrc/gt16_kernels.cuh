// gt16_kernels.cuh -- launch-side declarations shared by gt16_kernels.cu and gt16_capi.cu
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "gt16_core.cuh"
#include "../../include/gt16.h"

namespace gt16 {

// decision-table geometry (see DESIGN.md "exact memoised fast path")
constexpr int kS1Bits = 16384;                 // (Wr,0,0,0), Wr < 16384 (14-bit compact weight)
constexpr int kS2R = 1024, kS2A = 64;          // (Wr,Wp,0,0) and (Wr,0,Wq,0): Wr < 1024, alt < 64
constexpr int kG4R = 1024, kG4A = 32;          // (Wr,Wp,Wq1,Wq3): Wr < 1024, each alt < 32
constexpr size_t kS1Words = kS1Bits / 32;
constexpr size_t kS2Words = (size_t) kS2R * kS2A / 32;
constexpr size_t kG4Words = (size_t) kG4R * kG4A * kG4A * kG4A / 32;

struct DecisionTables {
  const uint32_t *s1;    // bit = 1: needs exact evaluation
  const uint32_t *s2p;   // index Wp * kS2R + Wr
  const uint32_t *s2q;   // index Wq * kS2R + Wr
  const uint32_t *g4;    // index ((Wq3 * kG4A + Wq1) * kG4A + Wp) * kG4R + Wr
};

// TMA-staged streaming kernel geometry
#ifndef GT16_STREAM_THREADS
#define GT16_STREAM_THREADS 1024
#endif
#ifndef GT16_STREAM_ROWS
#define GT16_STREAM_ROWS 8             // 32-site rows per warp and tile visit: tiles of 8,192 positions (64 KB of counters), 3 stages
#endif
constexpr int kStreamThreads = GT16_STREAM_THREADS;
// positions by which every streamed array must be over-allocated so that full tiles can be copied
constexpr size_t kStreamPadSites = 8192 + 64;

cudaError_t configure_kernels();   // per-device kernel attributes (dynamic shared memory opt-in)
void launch_build_table(double *out, float err, cudaStream_t s);
void launch_build_ycsq(double *out, cudaStream_t s);
void launch_build_decisions(const EvalParams &P, uint32_t *s1, uint32_t *s2p, uint32_t *s2q, uint32_t *g4,
                            cudaStream_t s);

// streaming classification of processed sites [first, first+n): bit i of mask = site first+i must be
// evaluated exactly (useTables) / is a candidate (exact inline mode)
// (shortcutAlt: stream_shortcut_bound() of the bitmaps, as for launch_stream_tma)
void launch_stream(const SnapView &V, const DevSiteMap &M, uint64_t first, uint32_t n, bool useTables,
                   const DecisionTables &D, uint32_t shortcutAlt, const EvalParams &P, uint32_t *mask, uint32_t *errWord,
                   cudaStream_t s);
// sparse snapshots: scatter n {u32 position; u64 counter} records (12 bytes each, strictly ascending positions; prevPos =
// position of the record before rec[0], 0xFFFFFFFF if none) into the position-indexed counter array, positions in
// [lo, hi) only; a record out of order or beyond L raises kErrBadSnapshot
void launch_scatter_sparse(const void *rec, uint64_t n, uint64_t *counters, uint32_t L, uint32_t lo, uint32_t hi,
                           uint32_t prevPos, uint32_t *errWord, cudaStream_t s);
// counters of regions [rA, rB] (device copies of the plan), widened by `margin` positions, clamped to [lo, hi) := 0
void launch_zero_spans(const uint32_t *regStart, const uint64_t *regFirst, uint32_t rA, uint32_t rB, uint64_t *counters,
                       uint32_t lo, uint32_t hi, uint32_t margin, cudaStream_t s);
// same contract for a dense single-region range of positions [pBeg,pEnd), pBeg % 128 == 0, arrays padded
// by kStreamPadSites: persistent CTAs, TMA bulk copies into a shared-memory ring
// shortcutAlt: stream_shortcut_bound() of the bitmaps in D (0 = take the table-driven path for every site)
void launch_stream_tma(const SnapView &V, uint32_t pBeg, uint32_t pEnd, const DecisionTables &D, uint32_t shortcutAlt,
                       uint32_t *mask, int numSMs, cudaStream_t s);
uint32_t stream_shortcut_bound(const uint32_t *s1, const uint32_t *s2p, const uint32_t *s2q);
// What the exact evaluator knows about a site it declares a candidate (one per called site, in the order the warps get
// there): the 16 likelihoods and posteriors it has just computed -- so that the record kernel does not evaluate them a
// second time -- and what the Fisher-test kernel needs to find the site's neighbours.
struct EvalOut {
  double lh[16];      // GenotypeLikelihood
  double flh[16];     // SnpCallingInfo.flh
  double pD, bestD, secondD;
  uint32_t site;      // processed index relative to the call's first site
  uint8_t genotype;
  uint8_t isIndel;    // 0, 2 or 3 (SNP_Caller.cpp:228-234)
  uint8_t nextBp;     // length of the deepest pattern when it is a deletion (:254-265)
  uint8_t pad;
};
static_assert(sizeof(EvalOut) == 288, "EvalOut layout");

// exact predicate for every site flagged in bits [firstSite, firstSite + nSites) of siteMask (firstSite % 32 == 0; bit i =
// site base + i): clears the bit unless the site is a candidate; *flagged += number of sites evaluated.  Every candidate
// takes the next free entry of part[] (*count is the number taken; entries beyond `cap` are counted, not written, and
// kErrRecordOverflow is raised).
void launch_eval_mask(const SnapView &V, const DevSiteMap &M, uint64_t base, uint32_t firstSite, uint32_t nSites,
                      bool useTables, const DecisionTables &D, const EvalParams &P, uint32_t *siteMask,
                      uint32_t *errWord, uint32_t *flagged, EvalOut *part, uint32_t *count, uint32_t cap, cudaStream_t s);
// rankOf[j] = number of set mask bits before part[j].site among sites [firstSite, ...) (its place in the ordered output),
// list[rank] = j; entries with site 0xFFFFFFFF (reserved, unused) are skipped; blockOffsets = scanned block counts of the
// same mask range
void launch_rank(const uint32_t *siteMask, uint32_t firstSite, const uint32_t *blockOffsets, const uint32_t *chunkOffsets,
                 const uint32_t *subCounts, const EvalOut *part, const uint32_t *count, uint32_t cap, uint32_t *rankOf,
                 uint32_t *list, cudaStream_t s);

// ordered compaction of a bit mask over nItems items
constexpr int kScanBlock = 256;                       // mask words per block
inline uint32_t scan_blocks(uint32_t nItems) { return ((nItems + 31) / 32 + kScanBlock - 1) / kScanBlock; }
// subCounts: 8 per block (set bits of each run of 32 words), so that a rank needs one run of words, not a block's worth
void launch_mask_count(const uint32_t *mask, uint32_t nItems, uint32_t *blockCounts, uint32_t *subCounts, cudaStream_t s);
// exclusive scan in two levels: blockCounts[b] becomes the offset within its chunk of 1024 blocks, chunkOffsets[c]
// (room for (nBlocks + 1023) / 1024 words) the offset of the chunk; *total = number of set bits
void launch_scan_counts(uint32_t *blockCounts, uint32_t nBlocks, uint32_t *chunkOffsets, uint32_t *total, cudaStream_t s);
// One Fisher test (strand bias or base-quality bias) that the memo cache did not hold: three counts and where the result
// goes (dst = record * 12 + k * 4 + base, bit 31 = base-quality test).
struct BiasTask { uint32_t W, F, R, dst; };

// full records of the candidates: out[r] from part[entryOf[r]], r < count[1] (persistent grid; count[0] = side-buffer entries
// reserved, count[1] = candidates, both read on the device; nothing is done when count[0] > cap) -- counts, window features, indel patterns and the strand-bias / base-quality-bias tests the memo cache holds
// (cache: (cacheMask + 1) 8-byte {key, result} entries, all bytes 0xFF when empty); tests it does not hold are appended to
// missed[0, missedCap) (*missedCount, zeroed by the caller) for launch_bias, and walked in place when that list is full
void launch_candidates(const SnapView &V, const DevSiteMap &M, uint64_t base, const EvalOut *part, const uint32_t *entryOf,
                       const uint32_t *count, uint32_t cap, const EvalParams &P, uint2 *cache, uint32_t cacheMask,
                       BiasTask *missed, uint32_t *missedCount, uint32_t missedCap, gt16_called_site *out, uint32_t *errWord,
                       int numSMs, cudaStream_t s);

// the Fisher-test walks of missed[0, min(*missedCount, missedCap)): results into out[dst / 12].sb / .bqb and into the cache
void launch_bias(BiasTask *missed, const uint32_t *missedCount, uint32_t missedCap, const EvalParams &P, uint2 *cache,
                 uint32_t cacheMask, gt16_called_site *out, int numSMs, cudaStream_t s);

// operator seam kernels
void launch_op_likelihood(const EvalParams &P, const uint8_t *w, size_t wStride, const uint8_t *ref,
                          size_t refStride, double *lh, uint32_t n, cudaStream_t s);
void launch_op_genotype(const EvalParams &P, const uint8_t *ref, const double *lh, uint8_t *info,
                        size_t infoStride, uint32_t n, cudaStream_t s);
void launch_op_bias(const EvalParams &P, const uint8_t *w, const uint8_t *f, const uint8_t *r, size_t stride,
                    uint32_t n, float *sb, float *bqb, cudaStream_t s);

}  // namespace gt16
