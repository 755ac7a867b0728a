// gt16_pileup.cu -- the step in front of the calling path: bam2snapshot's counter construction
// on sm_100a.  Aligned reads in, the snapshot's arrays out -- SnpCounter[L], invalid bitmap, per-region
// SnpOverflowCounter arrays, pattern pool -- in the encoding the caller reads (coreStruct.h:10-67).
//
// The reference (updateSnpCounterForReads / updateSnpCounter, SNP.cpp:438-891; the overflow file and
// updateSnpCounterForOverflow / ...ForInsert / ...ForDelete, :893-1688) folds the reads into the counters one
// at a time with compare-and-swap, so the bytes it writes depend on the order of the reads: which base sits
// in slot 1, which arrayIndex an overflow record gets, where a pattern list lies in the pool.  What a snapshot
// MEANS does not: per position the weight / forward / reverse totals of the four bases, whether the position
// is compact or overflow-type (three alleles seen, a strand count that would reach 255, an indel recorded
// there), whether it is invalid (a u16 total saturates, a pattern list outgrows 4 KB), and the insertion /
// deletion patterns with their hq / lq counts in order of first appearance.  That content is a histogram, and
// is built here as one:
//
//   k_pileup   one CTA per tile of 2,048 genome positions.  The reads that reach into the tile (the reads are
//              sorted by start; two binary searches) are walked by the CTA's warps -- the CIGAR walk of
//              SNP.cpp:518-741 with its head / tail trims and break conditions, lanes spread over the bases of
//              a match run -- into a shared-memory histogram (12 u32 per position, shared-memory atomics, no
//              traffic to HBM) and a shared-memory list of indel events.  The events are sorted (bitonic, in
//              shared memory) by position, kind and read number, folded into pattern lists in order of first
//              appearance, and every position is then encoded once: 8 coalesced bytes per position, an
//              overflow record and pool bytes where needed (global atomics only to reserve those).
//
// No CPU path: gt16_build_snapshot needs a context.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <functional>
#include <memory>
#include <string>
#include <thread>
#include <vector>
#include "gt16_kernels.cuh"
#include "gt16_pileup.cuh"

namespace gt16 {

constexpr int kPileTile = 2048, kPileThreads = 512, kPileEvents = 2048;
constexpr uint32_t kErrPileEvents = 8u, kErrPileOvf = 16u, kErrPilePool = 32u;

struct PileEvent {              // one insertion / deletion a read reports
  uint64_t key;                 // (position in tile << 1 | kind) << 32 | read number   (kind 0 insertion, 1 deletion)
  uint64_t seq;                 // offset of the inserted bases in the base array (insertions)
  uint32_t rep;                 // first event of the same position / kind with the same pattern
  uint16_t hq, lq;              // counts (accumulated on the rep)
  uint8_t len, isHq, pad[6];
};
static_assert(sizeof(PileEvent) == 32, "PileEvent layout");

struct PileSmem {
  uint32_t hist[kPileTile][12];           // W[4], F[4], R[4] per position
  PileEvent ev[kPileEvents];
  uint32_t nEvents;
  uint32_t evFlags[kPileTile / 32];       // bit 31 - (p & 31): position has an indel event
};
static_assert(sizeof(PileSmem) <= 227 * 1024, "shared memory budget");

struct PileArgs {
  // reads, sorted by start position
  const uint32_t *start, *length, *order;   // order: number of the read in the input (first appearance of a pattern)
  const uint64_t *seqOff, *cigarOff;
  const uint32_t *nCigar, *cigar;           // BAM encoding: length << 4 | op (0 M, 1 I, 2 D, 3 N, 4 S, 5 H, 6 P, 7 =, 8 X)
  const uint8_t *bases;                     // base (bits 1:0, A C G T) | weight << 2, in reference orientation
  uint64_t nReads;
  uint32_t maxSpan;                         // longest reference span of a read
  uint32_t L, trim, indelThreshold;
  // snapshot under construction
  uint64_t *counters;
  uint32_t *invalid;
  uint8_t *ovf;                             // nRegions arrays of ovfCap records of 32 bytes
  uint32_t *ovfCount;                       // per region
  uint32_t ovfCap, nRegions, ovfLen;
  double ovfInv;
  uint8_t *pool;
  uint32_t *poolPtr;
  uint32_t poolCap;
  uint32_t *errWord;
  // Tiles whose indel events do not fit the shared-memory list (deep amplicons: tens of thousands of reads over an indel)
  // are left out by the first launch -- it notes them and how many events they have -- and done by a second launch of the
  // same kernel, one CTA per noted tile, with the event list in global memory.
  uint32_t *heavyTiles, *heavyEvents, *heavyCount;    // first launch: tiles noted, their event counts, how many
  uint32_t heavyCap;
  const uint32_t *tileList;                           // second launch: the tiles to do (null in the first)
  PileEvent *heavyList;                               // ... their event lists, tile k at heavyOff[k]
  const uint64_t *heavyOff;
};

__device__ __forceinline__ uint32_t pile_region(const PileArgs &A, uint32_t p) {   // findRegionByPosition, coreStruct.h:141-154
  if (A.ovfLen == 0) return A.nRegions - 1;
  uint32_t r = (uint32_t) ((double) p * A.ovfInv);
  if ((uint64_t) r * A.ovfLen > p) --r;
  else if ((uint64_t) (r + 1) * A.ovfLen <= p) ++r;
  return r < A.nRegions - 1 ? r : A.nRegions - 1;
}

// one read into the tile's histogram and event list (whole warp; control flow is warp-uniform)
__device__ void pile_read(const PileArgs &A, PileSmem &S, PileEvent *ev, uint32_t evCap, uint64_t r, uint32_t tileLo, uint32_t lane) {
  const uint32_t readLength = A.length[r], nOps = A.nCigar[r];
  const uint32_t *ops = A.cigar + A.cigarOff[r];
  const uint8_t *sq = A.bases + A.seqOff[r];
  const uint32_t trimHead = A.trim, tailFrom = readLength - A.trim;      // unsigned, like the reference (:564)
  const uint64_t tileEnd64 = (uint64_t) tileLo + kPileTile;               // (the last tile of a genome close to 2^32 positions)
  const uint32_t tileEnd = tileEnd64 < A.L ? (uint32_t) tileEnd64 : A.L;
  uint32_t gpos = A.start[r], rpos = 0, trimmedHead = 0;
  uint32_t insLen = 0, insStart = 0, delLen = 0, delPos = 0;
  bool delActive = false;
  auto emit = [&](uint32_t pos, uint32_t kind, uint32_t len, bool hq, uint64_t seq) {   // lane 0
    if (pos < tileLo || pos >= tileEnd) return;
    const uint32_t at = atomicAdd(&S.nEvents, 1u);
    if (at >= evCap) return;                                   // counted: the tile goes to the second launch
    PileEvent e;
    e.key = ((uint64_t) (((pos - tileLo) << 1) | kind) << 32) | (A.order[r] & 0x7FFFFFFFu);
    e.seq = seq; e.rep = at; e.hq = 0; e.lq = 0; e.len = (uint8_t) len; e.isHq = hq ? 1 : 0;
    for (int k = 0; k < 6; ++k) e.pad[k] = 0;
    ev[at] = e;
  };
  // An indel reaches the snapshot only if an applied match base follows it in the read: the reference writes a read's
  // overflow bytes up to its last match base (lastMPos, :735-741).  Indels closed by another indel wait here for one.
  struct Held { uint32_t pos, len; uint64_t seq; bool del, hq; } held[4];
  uint32_t nHeld = 0;
  auto hold = [&](uint32_t pos, bool del, uint32_t len, bool hq, uint64_t seq) {
    if (nHeld < 4) { held[nHeld].pos = pos; held[nHeld].del = del; held[nHeld].len = len; held[nHeld].hq = hq; held[nHeld].seq = seq; ++nHeld; }
  };
  auto flush_insertion = [&](uint32_t pos) {       // readFilesForOverflow :961-971: average weight of the inserted bases
    if (lane == 0 && insLen) {
      float total = 0.0f;
      for (uint32_t k = 0; k < insLen; ++k) total += (float) (sq[insStart + k] >> 2);
      const uint32_t avg = (uint32_t) (unsigned char) (total / (float) (insLen & 0xFFu));
      hold(pos, false, insLen & 0xFFu, avg > A.indelThreshold, A.seqOff[r] + insStart);
    }
    insLen = 0;
  };
  auto flush_deletion = [&](uint32_t weight) {     // :973-978: the weight is that of the base that follows
    if (lane == 0 && delActive) hold(delPos, true, delLen & 0xFFu, weight > A.indelThreshold, 0);
    delActive = false; delLen = 0;
  };
  auto commit = [&]() {                            // a match base is being applied: what was held is on record
    for (uint32_t k = 0; k < nHeld; ++k) emit(held[k].pos, held[k].del ? 1u : 0u, held[k].len, held[k].hq, held[k].seq);
    nHeld = 0;
  };
  for (uint32_t i = 0; i < nOps; ++i) {
    uint32_t range = ops[i] >> 4;
    const uint32_t op = ops[i] & 15u;
    if (op == 0 || op == 7 || op == 8) {           // match / mismatch                                   :529-590
      if (rpos < trimHead) {
        const uint32_t room = trimHead - trimmedHead, t = range < room ? range : room;
        gpos += t; rpos += t; range -= t; trimmedHead += t;
      }
      if (range == 0) continue;
      if (rpos >= tailFrom) return;                // tail trim: the read ends here
      uint32_t n = tailFrom - rpos;
      const bool tail = n < range;
      if (!tail) n = range;
      // the first base of the run closes what is pending: the insertion lands on its position, the deletion takes its weight
      flush_insertion(gpos);
      flush_deletion((uint32_t) (sq[rpos] >> 2));
      commit();
      // does the run touch the tile?
      if (gpos < tileEnd && gpos + n > tileLo) {
        const uint32_t k0 = gpos < tileLo ? tileLo - gpos : 0u;
        uint32_t k1 = n;
        if (gpos + n > tileEnd) k1 = tileEnd - gpos;
        const uint32_t strandOff = A.order[r] >> 31 ? 4u : 8u;                  // F (forward read) or R: see below
        for (uint32_t k = k0 + lane; k < k1; k += 32) {
          const uint32_t p = gpos + k;
          if (p >= A.L) break;
          const uint32_t b = sq[rpos + k];
          uint32_t *h = S.hist[p - tileLo];
          if (b >> 2) atomicAdd(h + (b & 3u), b >> 2);
          atomicAdd(h + strandOff + (b & 3u), 1u);
        }
      }
      if (tail) return;
      gpos += n; rpos += n;
    } else if (op == 1 || op == 4) {               // insertion / soft clip                              :591-680
      if (rpos < trimHead) { rpos += range; trimmedHead += range; continue; }
      if (rpos + range >= readLength) return;
      if (op == 1) {
        if (delActive) flush_deletion((uint32_t) (sq[rpos] >> 2));             // "deletion before insertion" (:983-989)
        if (!insLen) insStart = rpos;
        insLen += range;
      }
      rpos += range;
    } else if (op == 2) {                          // deletion                                           :681-722
      if (rpos < trimHead) { gpos += range; continue; }
      if (rpos + 1 >= readLength) return;
      flush_insertion(gpos);                       // :1007-1016
      if (!delActive) { delActive = true; delPos = gpos; delLen = 0; }
      delLen += range;
      gpos += range;
    }                                              // N, H, P: not handled by the reference's walk either
  }
}

__global__ void __launch_bounds__(kPileThreads, 1) k_pileup(PileArgs A) {
  extern __shared__ __align__(16) unsigned char smemRaw[];
  PileSmem &S = *reinterpret_cast<PileSmem *>(smemRaw);
  const uint32_t tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t tile = A.tileList ? A.tileList[blockIdx.x] : blockIdx.x;
  const uint32_t tileLo = tile * (uint32_t) kPileTile;
  PileEvent *ev = A.tileList ? A.heavyList + A.heavyOff[blockIdx.x] : S.ev;
  const uint32_t evCap = A.tileList ? (uint32_t) (A.heavyOff[blockIdx.x + 1] - A.heavyOff[blockIdx.x]) : (uint32_t) kPileEvents;
  const uint32_t tileHi = (uint64_t) tileLo + kPileTile < A.L ? tileLo + (uint32_t) kPileTile : A.L;
  // reads with start in (tileLo - maxSpan, tileHi)
  uint64_t rLo, rHi;
  {
    const uint32_t from = tileLo > A.maxSpan ? tileLo - A.maxSpan : 0u;
    uint64_t lo = 0, hi = A.nReads;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (A.start[mid] < from) lo = mid + 1; else hi = mid; }
    rLo = lo; hi = A.nReads;
    while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (A.start[mid] < tileHi) lo = mid + 1; else hi = mid; }
    rHi = lo;
  }
  if (rLo == rHi) {                                 // nothing reaches the tile: zero counters, valid positions
    for (uint32_t p = tileLo + tid; p < tileHi; p += kPileThreads) A.counters[p] = 0ull;
    if (tid < (uint32_t) kPileTile / 32 && tileLo + tid * 32u < A.L) A.invalid[(tileLo >> 5) + tid] = 0u;
    return;
  }
  for (uint32_t k = tid; k < (uint32_t) kPileTile * 12; k += kPileThreads) (&S.hist[0][0])[k] = 0u;
  if (tid < (uint32_t) kPileTile / 32) S.evFlags[tid] = 0u;
  if (tid == 0) S.nEvents = 0;
  __syncthreads();
  for (uint64_t r = rLo + warp; r < rHi; r += kPileThreads / 32) pile_read(A, S, ev, evCap, r, tileLo, lane);
  __syncthreads();
  if (S.nEvents > evCap) {                          // more events than the list holds: noted for the second launch
    if (tid == 0) {
      if (A.tileList) atomicOr(A.errWord, kErrPileEvents);            // cannot happen: the list was sized by the first launch's count
      else {
        const uint32_t k = atomicAdd(A.heavyCount, 1u);
        if (k < A.heavyCap) { A.heavyTiles[k] = tile; A.heavyEvents[k] = S.nEvents; }
      }
    }
    return;
  }

  // ---- indel events: sort by (position, kind, read number), fold equal patterns onto their first appearance
  const uint32_t nEv = S.nEvents;
  if (nEv > 1) {
    uint32_t n2 = 2; while (n2 < nEv) n2 <<= 1;
    for (uint32_t k = nEv + tid; k < n2; k += kPileThreads) ev[k].key = ~0ull;          // padding sorts last
    __syncthreads();
    for (uint32_t size = 2; size <= n2; size <<= 1)
      for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
        for (uint32_t t = tid; t < n2 / 2; t += kPileThreads) {
          const uint32_t i = 2 * t - (t & (stride - 1)), j = i + stride;
          const bool up = (i & size) == 0;
          if ((ev[i].key > ev[j].key) == up) { const PileEvent x = ev[i]; ev[i] = ev[j]; ev[j] = x; }
        }
        __syncthreads();
      }
  }
  // group leaders: the first event of each (position, kind)
  for (uint32_t g = tid; g < nEv; g += kPileThreads) {
    const uint32_t gk = (uint32_t) (ev[g].key >> 32);
    if (g > 0 && (uint32_t) (ev[g - 1].key >> 32) == gk) continue;
    for (uint32_t e = g; e < nEv && (uint32_t) (ev[e].key >> 32) == gk; ++e) {
      uint32_t rep = e;
      for (uint32_t f = g; f < e; ++f) {
        if (ev[f].rep != f || ev[f].len != ev[e].len) continue;
        bool same = true;
        if (!(gk & 1u))                                                                  // insertions: same bases
          for (uint32_t k = 0; k < ev[e].len && same; ++k) same = ((A.bases[ev[f].seq + k] ^ A.bases[ev[e].seq + k]) & 3u) == 0;
        if (same) { rep = f; break; }
      }
      ev[e].rep = rep;
      PileEvent &R = ev[rep];
      if (ev[e].isHq) { if (R.hq < 0xFFFFu) ++R.hq; } else { if (R.lq < 0xFFFFu) ++R.lq; }
    }
    const uint32_t p = gk >> 1;
    atomicOr(&S.evFlags[p >> 5], 0x80000000u >> (p & 31));
  }
  __syncthreads();

  // ---- encode: one thread per position
  for (uint32_t pr = tid; pr < (uint32_t) kPileTile; pr += kPileThreads) {
    const uint32_t p = tileLo + pr;
    bool invalidBit = false;
    uint64_t counter = 0;
    if (p < A.L) {
      const uint32_t *h = S.hist[pr];
      uint32_t W[4], F[4], R[4], alleles = 0;
      bool deep = false;
      for (int b = 0; b < 4; ++b) {
        W[b] = h[b]; F[b] = h[4 + b]; R[b] = h[8 + b];
        if (F[b] | R[b]) ++alleles;
        if (F[b] >= 255u || R[b] >= 255u) deep = true;                    // updateSnpCounter :852-861: a compact strand count stops at 254
        if (F[b] >= 0xFFFCu || R[b] >= 0xFFFCu || W[b] >= 0xFFFFu) invalidBit = true;   // :783-813, :1123-1135
      }
      const bool hasIndel = (S.evFlags[pr >> 5] << (pr & 31)) >> 31;
      if (!(alleles >= 3 || deep || hasIndel)) {
        // compact: the allele with more reads in slot 1 (where a sequential producer meets the reference allele first)
        int s1 = -1, s2 = -1;
        for (int b = 0; b < 4; ++b) if (F[b] | R[b]) { if (s1 < 0) s1 = b; else s2 = b; }
        if (s2 >= 0 && F[s2] + R[s2] > F[s1] + R[s1]) { const int t = s1; s1 = s2; s2 = t; }
        uint32_t lo = 0, hi = 0;
        if (s1 >= 0) { lo |= ((uint32_t) s1 << 14) | W[s1]; hi |= F[s1] | (R[s1] << 8); }
        if (s2 >= 0) { lo |= (((uint32_t) s2 << 14) | W[s2]) << 16; hi |= (F[s2] << 16) | (R[s2] << 24); }
        counter = (uint64_t) lo | ((uint64_t) hi << 32);
      } else {
        uint32_t w1 = 0xC000u, w2 = 0xC000u;
        uint32_t rec[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        rec[0] = (W[0] & 0xFFFFu) | (W[1] << 16); rec[1] = (W[2] & 0xFFFFu) | (W[3] << 16);
        rec[2] = (F[0] & 0xFFFFu) | (F[1] << 16); rec[3] = (F[2] & 0xFFFFu) | (F[3] << 16);
        rec[4] = (R[0] & 0xFFFFu) | (R[1] << 16); rec[5] = (R[2] & 0xFFFFu) | (R[3] << 16);
        if (hasIndel) {
          // the position's events are contiguous in the sorted list: find them (binary search on the group key)
          for (uint32_t kind = 0; kind < 2; ++kind) {
            const uint32_t gk = (pr << 1) | kind;
            uint32_t lo = 0, hi = nEv;
            while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if ((uint32_t) (ev[mid].key >> 32) < gk) lo = mid + 1; else hi = mid; }
            uint32_t g = lo, nPat = 0, bytes = 0, one = 0;
            for (uint32_t e = g; e < nEv && (uint32_t) (ev[e].key >> 32) == gk; ++e)
              if (ev[e].rep == e) {
                if (!nPat) one = e;
                ++nPat;
                bytes += kind ? 5u : 1u + ((uint32_t) ev[e].len + 3) / 4 + 4u;
                if (ev[e].hq >= 0xFFFFu || ev[e].lq >= 0xFFFFu) invalidBit = true;    // :1343-1372
              }
            if (!nPat) continue;
            if (bytes >= 4096u) { invalidBit = true; continue; }                          // :1408-1419, :1445-1456
            const PileEvent &E = ev[one];
            const bool inlineOk = nPat == 1 && E.hq < 0xFFu && E.lq < 0xFFu && (kind || E.len <= 12);
            uint32_t &w = kind ? w2 : w1;
            w |= 0x2000u;
            if (inlineOk && !kind) {                                                      // :1219-1229
              w |= ((uint32_t) E.len << 8) | E.hq;
              uint32_t u = (uint32_t) E.lq << 24;
              for (uint32_t k = 0; k < E.len; ++k) u |= (uint32_t) (A.bases[E.seq + k] & 3u) << ((k >> 2) * 8 + ((k & 3u) << 1));
              rec[6] = u;
            } else if (inlineOk) {                                                        // :1507-1513
              rec[7] = (uint32_t) E.len | ((uint32_t) E.hq << 8) | ((uint32_t) E.lq << 16);
            } else {
              const uint32_t at = atomicAdd(A.poolPtr, (bytes + 3u) & ~3u);
              if (at + bytes > A.poolCap) { atomicOr(A.errWord, kErrPilePool); continue; }
              uint8_t *dst = A.pool + at;
              uint32_t q = 0;
              for (uint32_t e = g; e < nEv && (uint32_t) (ev[e].key >> 32) == gk; ++e) {
                if (ev[e].rep != e) continue;
                const PileEvent &X = ev[e];
                dst[q++] = X.len;
                if (!kind) {
                  const uint32_t nb = ((uint32_t) X.len + 3) / 4;
                  for (uint32_t k = 0; k < nb; ++k) dst[q + k] = 0;
                  for (uint32_t k = 0; k < X.len; ++k) dst[q + (k >> 2)] |= (uint8_t) ((A.bases[X.seq + k] & 3u) << ((k & 3u) << 1));
                  q += nb;
                }
                dst[q++] = (uint8_t) (X.hq & 0xFF); dst[q++] = (uint8_t) (X.hq >> 8);
                dst[q++] = (uint8_t) (X.lq & 0xFF); dst[q++] = (uint8_t) (X.lq >> 8);
              }
              w |= 0x1000u | bytes;
              rec[kind ? 7 : 6] = at;
            }
          }
        }
        const uint32_t region = pile_region(A, p);
        const uint32_t idx = atomicAdd(A.ovfCount + region, 1u);
        if (idx >= A.ovfCap) atomicOr(A.errWord, kErrPileOvf);
        else {
          uint4 *dst = reinterpret_cast<uint4 *>(A.ovf + ((size_t) region * A.ovfCap + idx) * 32);
          dst[0] = make_uint4(rec[0], rec[1], rec[2], rec[3]);
          dst[1] = make_uint4(rec[4], rec[5], rec[6], rec[7]);
        }
        counter = (uint64_t) (w1 | (w2 << 16)) | ((uint64_t) idx << 32);
      }
      A.counters[p] = counter;
    }
    const uint32_t word = __ballot_sync(0xffffffffu, invalidBit);
    if (lane == 0 && p < A.L) A.invalid[p >> 5] = __brev(word);          // bit 31 - (p & 31), SNP.cpp:790
  }
}

}  // namespace gt16

using namespace gt16;

// ------------------------------------------------------------------------------------ C ABI
struct gt16_ctx;
extern "C" const char *gt16_last_error(const gt16_ctx *ctx);
int gt16_pileup_set_error(gt16_ctx *ctx, int code, const char *msg);          // gt16_capi.cu
int gt16_pileup_device(gt16_ctx *ctx);                                        // gt16_capi.cu: cudaSetDevice(ctx's device)

namespace {
template <typename T>
struct Dev {
  T *p = nullptr;
  ~Dev() { if (p) cudaFree(p); }
  cudaError_t alloc(size_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
  cudaError_t put(const T *src, size_t n) { cudaError_t e = alloc(n); return e != cudaSuccess || !n ? e : cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice); }
};
}  // namespace

#define PCU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { (void) cudaGetLastError(); \
    return gt16_pileup_set_error(ctx, GT16_ECUDA, (std::string(#call) + " failed: " + cudaGetErrorString(e_)).c_str()); } } while (0)

extern "C" int gt16_build_snapshot(gt16_ctx *ctx, const gt16_reads *reads, uint32_t text_length, const gt16_pileup_params *params,
                                   gt16_built_snapshot *out) {
  if (!ctx || !reads || !params || !out || !text_length || !params->n_overflow_regions) return GT16_EINVAL;
  int rc = gt16_pileup_device(ctx);
  if (rc) return rc;
  memset(out, 0, sizeof(*out));
  auto clock0 = std::chrono::steady_clock::now();
  auto ms_since = [](std::chrono::steady_clock::time_point t0) { return std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
  const uint64_t n = reads->n_reads;
  const uint32_t L = text_length;
  if (n >= 0x7FFFFFFFull) return gt16_pileup_set_error(ctx, GT16_EINVAL, "too many reads for one call");
  unsigned nThreads = std::thread::hardware_concurrency();
  nThreads = std::max(1u, std::min(16u, nThreads ? nThreads : 4u));
  auto parallel = [&](uint64_t count, const std::function<void(unsigned, uint64_t, uint64_t)> &f) {    // f(slice, lo, hi)
    const unsigned nt = count < (1u << 16) ? 1u : nThreads;
    if (nt == 1) { f(0, 0, count); return; }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t]() { f(t, count * t / nt, count * (t + 1) / nt); });
    for (auto &x : th) x.join();
  };
  // the kernel indexes bases[] and cigar[] by what the read records say: they are checked here, once, on the host
  {
    std::vector<uint64_t> firstBad(nThreads, UINT64_MAX);
    parallel(n, [&](unsigned t, uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; ++i) {
        bool ok = reads->start[i] < L && reads->n_cigar[i] >= 1 && reads->length[i] >= 1 &&
                  reads->cigar_off[i] <= reads->n_cigar_ops && reads->n_cigar[i] <= reads->n_cigar_ops - reads->cigar_off[i] &&
                  reads->seq_off[i] <= reads->n_bases && reads->length[i] <= reads->n_bases - reads->seq_off[i];
        uint64_t query = 0;
        for (uint32_t c = 0; ok && c < reads->n_cigar[i]; ++c) {
          const uint32_t v = reads->cigar[reads->cigar_off[i] + c], op = v & 15u;
          if (op > 8) ok = false;
          if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) query += v >> 4;
        }
        if (!ok || query > reads->length[i]) { firstBad[t] = i; return; }
      }
    });
    const uint64_t bad = *std::min_element(firstBad.begin(), firstBad.end());
    if (bad != UINT64_MAX) {
      char msg[160];
      snprintf(msg, sizeof(msg), "read %llu: start, CIGAR or sequence range outside the arrays (or a CIGAR longer than the read)", (unsigned long long) bad);
      return gt16_pileup_set_error(ctx, GT16_EINVAL, msg);
    }
  }
  // ---- sort the reads by start (stable: the input order decides which pattern appears first), pack what the kernel reads
  std::vector<uint32_t> order(n);
  for (uint64_t i = 0; i < n; ++i) order[i] = (uint32_t) i;
  bool sorted = true;
  for (uint64_t i = 1; i < n && sorted; ++i) sorted = reads->start[i - 1] <= reads->start[i];
  if (!sorted) {                                  // coordinate-sorted input never gets here; several files side by side do
    std::vector<uint32_t> other(n);
    std::vector<uint64_t> count(65537);
    for (int pass = 0; pass < 2; ++pass) {        // stable radix sort on the two halves of the start position
      const int shift = 16 * pass;
      std::fill(count.begin(), count.end(), 0);
      for (uint64_t k = 0; k < n; ++k) count[((reads->start[order[k]] >> shift) & 0xFFFFu) + 1]++;
      for (int d = 0; d < 65536; ++d) count[d + 1] += count[d];
      for (uint64_t k = 0; k < n; ++k) other[count[(reads->start[order[k]] >> shift) & 0xFFFFu]++] = order[k];
      order.swap(other);
    }
  }
  std::vector<uint32_t> hStart(n), hLen(n), hOrder(n), hNCig(n);
  std::vector<uint64_t> hSeq(n), hCig(n);
  uint32_t maxSpan = 1;
  {
    std::vector<uint32_t> spanOf(nThreads, 1);
    parallel(n, [&](unsigned t, uint64_t lo, uint64_t hi) {
      uint32_t mine = 1;
      for (uint64_t k = lo; k < hi; ++k) {
        const uint64_t i = order[k];
        hStart[k] = reads->start[i]; hLen[k] = reads->length[i]; hNCig[k] = reads->n_cigar[i];
        hSeq[k] = reads->seq_off[i]; hCig[k] = reads->cigar_off[i];
        // bit 31: the read lies on the forward strand (its bases count as posStrandCount)
        hOrder[k] = (uint32_t) i | (reads->reverse[i] ? 0u : 0x80000000u);
        uint64_t span = 0;
        for (uint32_t c = 0; c < reads->n_cigar[i]; ++c) {
          const uint32_t op = reads->cigar[reads->cigar_off[i] + c] & 15u, len = reads->cigar[reads->cigar_off[i] + c] >> 4;
          if (op == 0 || op == 2 || op == 7 || op == 8) span += len;
        }
        mine = (uint32_t) std::max<uint64_t>(mine, std::min<uint64_t>(span + 1, 0x7FFFFFFFull));
      }
      spanOf[t] = mine;
    });
    maxSpan = *std::max_element(spanOf.begin(), spanOf.end());
  }
  // base | weight << 2 per base (weightMap of the ini, bam2snapshot.cpp:352-363; quality correction :478-481)
  const uint64_t nBases = reads->n_bases;
  std::unique_ptr<uint8_t[]> hBases(new uint8_t[nBases ? nBases : 1]);       // not zeroed: every byte is written below
  {
    uint8_t lut[256];                                // quality byte -> weight << 2 (0xFF = absent = negative = 0)
    for (int q = 0; q < 256; ++q) lut[q] = (uint8_t) (params->weight_map[q >= 128 ? 0 : q > 40 ? 40 : q] << 2);
    uint8_t *dst = hBases.get();
    parallel(nBases, [&](unsigned, uint64_t lo, uint64_t hi) {
      for (uint64_t k = lo; k < hi; ++k) dst[k] = (uint8_t) ((reads->bases[k] & 3u) | lut[reads->quals[k]]);
    });
  }
  const float msPack = ms_since(clock0);
  clock0 = std::chrono::steady_clock::now();
  Dev<uint32_t> dStart, dLen, dOrder, dNCig, dCigar, dInvalid, dOvfCount, dPoolPtr, dErr;
  Dev<uint64_t> dSeq, dCig, dCounters;
  Dev<uint8_t> dBases, dOvf, dPool;
  PCU(dStart.put(hStart.data(), n)); PCU(dLen.put(hLen.data(), n)); PCU(dOrder.put(hOrder.data(), n));
  PCU(dNCig.put(hNCig.data(), n)); PCU(dSeq.put(hSeq.data(), n)); PCU(dCig.put(hCig.data(), n));
  PCU(dCigar.put(reads->cigar, reads->n_cigar_ops)); PCU(dBases.put(hBases.get(), nBases));
  const uint32_t nReg = params->n_overflow_regions, regLen = L / nReg;
  const size_t words = ((size_t) L + 31) >> 5;
  PCU(dCounters.alloc((size_t) L + 1)); PCU(dInvalid.alloc(words + 1));
  PCU(dOvfCount.alloc(nReg)); PCU(dPoolPtr.alloc(1)); PCU(dErr.alloc(1));
  PCU(cudaFuncSetAttribute(k_pileup, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sizeof(PileSmem)));
  // overflow arrays and the pool are sized by a guess and grown when the kernel reports that they were too small
  uint32_t ovfCap = std::max<uint32_t>(4096u, (uint32_t) std::min<uint64_t>(((uint64_t) std::max(regLen, L - regLen * (nReg - 1))) / 8 + 4096, 0x7FFFFFF0ull));
  uint32_t poolCap = 16u << 20;
  std::vector<uint32_t> hCount(nReg);
  uint32_t hPool = 0;
  const float msH2d = ms_since(clock0);
  float msKernel = 0.f;
  cudaEvent_t e0, e1;
  PCU(cudaEventCreate(&e0)); PCU(cudaEventCreate(&e1));
  struct EvGuard { cudaEvent_t a, b; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); } } evGuard{e0, e1};
  for (int attempt = 0;; ++attempt) {
    if (dOvf.p) { cudaFree(dOvf.p); dOvf.p = nullptr; }
    if (dPool.p) { cudaFree(dPool.p); dPool.p = nullptr; }
    PCU(dOvf.alloc((size_t) nReg * ovfCap * 32)); PCU(dPool.alloc(poolCap));
    PCU(cudaMemset(dOvfCount.p, 0, nReg * 4)); PCU(cudaMemset(dPoolPtr.p, 0, 4)); PCU(cudaMemset(dErr.p, 0, 4));
    PileArgs A;
    A.start = dStart.p; A.length = dLen.p; A.order = dOrder.p; A.seqOff = dSeq.p; A.cigarOff = dCig.p; A.nCigar = dNCig.p;
    A.cigar = dCigar.p; A.bases = dBases.p; A.nReads = n; A.maxSpan = maxSpan; A.L = L; A.trim = params->trim_size;
    A.indelThreshold = params->indel_weight_threshold;
    A.counters = dCounters.p; A.invalid = dInvalid.p; A.ovf = dOvf.p; A.ovfCount = dOvfCount.p; A.ovfCap = ovfCap;
    A.nRegions = nReg; A.ovfLen = regLen; A.ovfInv = regLen ? 1.0 / (double) regLen : 0.0;
    A.pool = dPool.p; A.poolPtr = dPoolPtr.p; A.poolCap = poolCap; A.errWord = dErr.p;
    const unsigned tiles = (unsigned) (((uint64_t) L + kPileTile - 1) / kPileTile);
    Dev<uint32_t> dHeavyTiles, dHeavyEvents, dHeavyCount;
    PCU(dHeavyTiles.alloc(tiles)); PCU(dHeavyEvents.alloc(tiles)); PCU(dHeavyCount.alloc(1));
    PCU(cudaMemset(dHeavyCount.p, 0, 4));
    A.heavyTiles = dHeavyTiles.p; A.heavyEvents = dHeavyEvents.p; A.heavyCount = dHeavyCount.p; A.heavyCap = tiles;
    A.tileList = nullptr; A.heavyList = nullptr; A.heavyOff = nullptr;
    PCU(cudaEventRecord(e0, 0));
    k_pileup<<<tiles, kPileThreads, sizeof(PileSmem)>>>(A);
    PCU(cudaGetLastError());
    PCU(cudaEventRecord(e1, 0));
    PCU(cudaDeviceSynchronize());
    PCU(cudaEventElapsedTime(&msKernel, e0, e1));
    // tiles whose events did not fit the shared-memory list: again, with lists in global memory sized by their counts
    uint32_t nHeavy = 0;
    PCU(cudaMemcpy(&nHeavy, dHeavyCount.p, 4, cudaMemcpyDeviceToHost));
    if (nHeavy) {
      std::vector<uint32_t> hv(nHeavy);
      std::vector<uint64_t> off(nHeavy + 1, 0);
      PCU(cudaMemcpy(hv.data(), dHeavyEvents.p, (size_t) nHeavy * 4, cudaMemcpyDeviceToHost));
      for (uint32_t k = 0; k < nHeavy; ++k) { uint64_t c = 2; while (c < hv[k]) c <<= 1; off[k + 1] = off[k] + c; }   // the sort pads to a power of two
      Dev<PileEvent> dList;
      Dev<uint64_t> dOff;
      PCU(dList.alloc((size_t) off[nHeavy])); PCU(dOff.put(off.data(), off.size()));
      A.tileList = dHeavyTiles.p; A.heavyList = dList.p; A.heavyOff = dOff.p;
      float ms2 = 0.f;
      PCU(cudaEventRecord(e0, 0));
      k_pileup<<<nHeavy, kPileThreads, sizeof(PileSmem)>>>(A);
      PCU(cudaGetLastError());
      PCU(cudaEventRecord(e1, 0));
      PCU(cudaDeviceSynchronize());
      PCU(cudaEventElapsedTime(&ms2, e0, e1));
      msKernel += ms2;
    }
    out->heavy_tiles = nHeavy;
    uint32_t err = 0;
    PCU(cudaMemcpy(&err, dErr.p, 4, cudaMemcpyDeviceToHost));
    PCU(cudaMemcpy(hCount.data(), dOvfCount.p, nReg * 4, cudaMemcpyDeviceToHost));
    PCU(cudaMemcpy(&hPool, dPoolPtr.p, 4, cudaMemcpyDeviceToHost));
    if (err & kErrPileEvents) return gt16_pileup_set_error(ctx, GT16_ECUDA, "indel event list of a tile overflowed in the second launch");
    if (!(err & (kErrPileOvf | kErrPilePool))) break;
    if (attempt >= 4) return gt16_pileup_set_error(ctx, GT16_ECUDA, "overflow arrays / pattern pool could not be sized");
    if (err & kErrPileOvf) ovfCap = *std::max_element(hCount.begin(), hCount.end()) + 4096;
    if (err & kErrPilePool) poolCap = std::max(poolCap * 2, hPool + (1u << 20));
  }
  // ---- results to the host
  clock0 = std::chrono::steady_clock::now();
  uint64_t *counters = (uint64_t *) malloc(((size_t) L + 2) * 8);
  uint32_t *invalid = (uint32_t *) malloc((words + 1) * 4);
  uint8_t *pool = (uint8_t *) malloc((size_t) hPool + 16);
  size_t nOvf = 0;
  for (uint32_t r = 0; r < nReg; ++r) nOvf += hCount[r];
  uint8_t *ovf = (uint8_t *) malloc(nOvf * 32 + 32);
  uint32_t *ovfSize = (uint32_t *) malloc(nReg * 4);
  const void **ovfPtr = (const void **) malloc(nReg * sizeof(void *));
  if (!counters || !invalid || !pool || !ovf || !ovfSize || !ovfPtr) {
    free(counters); free(invalid); free(pool); free(ovf); free(ovfSize); free(ovfPtr);
    return gt16_pileup_set_error(ctx, GT16_EIO, "Memory Allocation Failed!");
  }
  cudaError_t e = cudaMemcpy(counters, dCounters.p, (size_t) L * 8, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(invalid, dInvalid.p, words * 4, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && hPool) e = cudaMemcpy(pool, dPool.p, hPool, cudaMemcpyDeviceToHost);
  size_t off = 0;
  for (uint32_t r = 0; r < nReg && e == cudaSuccess; ++r) {
    ovfSize[r] = hCount[r]; ovfPtr[r] = ovf + off;
    if (hCount[r]) e = cudaMemcpy(ovf + off, dOvf.p + (size_t) r * ovfCap * 32, (size_t) hCount[r] * 32, cudaMemcpyDeviceToHost);
    off += (size_t) hCount[r] * 32;
  }
  if (e != cudaSuccess) {
    free(counters); free(invalid); free(pool); free(ovf); free(ovfSize); free(ovfPtr);
    (void) cudaGetLastError();
    return gt16_pileup_set_error(ctx, GT16_ECUDA, cudaGetErrorString(e));
  }
  counters[L] = counters[L - 1];
  // the invalid bits beyond L in the last word are zero (the kernel's ballot masks them)
  out->view.counters = counters; out->view.text_length = L; out->view.invalid_bitmap = invalid;
  out->view.n_overflow_regions = nReg; out->view.overflow = ovfPtr; out->view.overflow_size = ovfSize;
  out->view.pool = pool; out->view.pool_bytes = hPool;
  out->overflow_storage = ovf;
  out->ms_pack = msPack; out->ms_h2d = msH2d; out->ms_kernel = msKernel; out->ms_d2h = ms_since(clock0);
  out->aligned_bases = nBases;
  return GT16_OK;
}

extern "C" void gt16_built_snapshot_free(gt16_built_snapshot *s) {
  if (!s) return;
  free((void *) s->view.counters); free((void *) s->view.invalid_bitmap); free((void *) s->view.pool);
  free((void *) s->view.overflow); free((void *) s->view.overflow_size); free(s->overflow_storage);
  memset(s, 0, sizeof(*s));
}
