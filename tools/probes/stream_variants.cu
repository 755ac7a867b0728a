// stream_variants.cu -- DEV TOOL (not product): the streaming-classification kernel as a template with
// geometry / structure knobs, run on synthetic counters with the statistics of a 30x snapshot, to find
// which structure gets closest to the HBM roofline.  The winning structure is what gt16_kernels.cu ships.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o stream_variants stream_variants.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t) __cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
template <int kSleep>
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  for (;;) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (ok) break;
    if (kSleep) __nanosleep(kSleep);
    if (++spins > (1u << 22)) __trap();
  }
}

constexpr int kS1 = 16384, kS2Words = 2048;
struct QEntry { uint32_t x, y, bit, ref; };

__device__ __forceinline__ bool general_test(const uint32_t *s2, const uint8_t *s1, uint32_t x, uint32_t y, uint32_t ref) {
  uint32_t v1 = (y & 0xFFFFu) ? (x & 0x3FFFu) : 0u, v2 = (y >> 16) ? ((x >> 16) & 0x3FFFu) : 0u;
  const uint32_t r1 = ((x >> 14) & 3u) ^ ref, r2 = (x >> 30) ^ ref;
  if (r1 == r2 && (y >> 16)) v1 = 0;
  const uint32_t wr = (r1 == 0 ? v1 : 0u) | (r2 == 0 ? v2 : 0u);
  const uint32_t a1 = r1 ? v1 : 0u, a2 = r2 ? v2 : 0u, alt = a1 | a2, ar = a1 ? r1 : r2;
  if (alt == 0) return s1[wr] != 0;
  if ((a1 && a2) || wr >= 1024u || alt >= 64u) return true;
  const uint32_t idx = (ar == 2 ? 0u : 65536u) + alt * 1024u + wr;
  return (s2[idx >> 5] >> (idx & 31)) & 1u;
}

// MODE 0: full (simple / special / complex with per-warp queue)   1: complex sites are just flagged (no queue)
//      2: no per-site work at all (wait, touch, release)            3: complex handled inline by every lane (no queue)
// OPT bits: 1 = specialise full tiles (no per-site range test)   2 = successor flags from ballots (no look-ahead load)
//           4 = 8-byte queue entries, pushes without branches, drains once per tile   8 = one vector mask store per tile
template <int THREADS, int TILE, int STAGES, int MODE, int SLEEP, int MINBLOCKS, int OPT>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
k_stream(const uint64_t *__restrict__ counters, const uint32_t *__restrict__ packedRef, const uint32_t *__restrict__ invalid,
         uint32_t pBeg, uint32_t pEnd, const uint8_t *__restrict__ gs1, const uint32_t *__restrict__ gs2, uint32_t *__restrict__ mask) {
  constexpr int WARPS = THREADS / 32, ROWS = TILE / THREADS;
  extern __shared__ __align__(128) unsigned char raw[];
  uint64_t *sCnt = reinterpret_cast<uint64_t *>(raw);                                   // [STAGES][TILE + 2]
  uint32_t *sRef = reinterpret_cast<uint32_t *>(sCnt + (size_t) STAGES * (TILE + 2));   // [STAGES][TILE / 16]
  uint32_t *sBm = sRef + STAGES * (TILE / 16);                                          // [STAGES][TILE / 32]
  uint8_t *s1 = reinterpret_cast<uint8_t *>(sBm + STAGES * (TILE / 32));
  uint32_t *s2 = reinterpret_cast<uint32_t *>(s1 + kS1);
  QEntry *queue = reinterpret_cast<QEntry *>(s2 + 2 * kS2Words);                         // [WARPS][64] or uint2 [WARPS][QCAP]
  constexpr uint32_t QCAP = 32 * ROWS + 32;
  uint64_t *full = reinterpret_cast<uint64_t *>(reinterpret_cast<unsigned char *>(queue) + (MODE == 0 ? ((OPT & 4) ? WARPS * QCAP * 8 : WARPS * 64 * 16) : MODE == 4 ? WARPS * 128 * 8 : 0));
  uint32_t *released = reinterpret_cast<uint32_t *>(full + STAGES);
  const uint32_t nTiles = (pEnd - pBeg + TILE - 1) / TILE, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t kBytesC = 8 * (TILE + 2), kBytesR = TILE / 4, kBytesB = TILE / 8;
  auto issue = [&](uint32_t stage, uint32_t tile) {
    const uint32_t pT = pBeg + tile * TILE;
    mbar_expect_tx(&full[stage], kBytesC + kBytesR + kBytesB);
    bulk_g2s(sCnt + (size_t) stage * (TILE + 2), counters + pT, kBytesC, &full[stage]);
    bulk_g2s(sRef + stage * (TILE / 16), packedRef + (pT >> 4), kBytesR, &full[stage]);
    bulk_g2s(sBm + stage * (TILE / 32), invalid + (pT >> 5), kBytesB, &full[stage]);
  };
  if (tid == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); released[s] = 0; } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  for (uint32_t k = tid; k < kS1 / 4; k += THREADS) reinterpret_cast<uint32_t *>(s1)[k] = reinterpret_cast<const uint32_t *>(gs1)[k];
  for (uint32_t k = tid; k < 2 * kS2Words; k += THREADS) s2[k] = gs2[k];
  __syncthreads();
  if (tid == 0) for (uint32_t s = 0; s < (uint32_t) STAGES; ++s) { uint32_t t = blockIdx.x + s * gridDim.x; if (t < nTiles) issue(s, t); }
  QEntry *q = queue + warp * 64;
  uint2 *q8 = reinterpret_cast<uint2 *>(queue) + warp * QCAP;
  uint2 *stk = reinterpret_cast<uint2 *>(queue) + warp * 128 + lane;     // MODE 4: lane-private stack, slot d at stk[32 * d]
  uint32_t depth = 0;
  uint32_t qhead = 0, qtail = 0;
  auto wrapq = [](uint32_t p) { return p >= QCAP ? p - QCAP : p; };
  auto drain = [&](uint32_t count) {
    if (OPT & 4) {
      if (lane < count) {
        const uint2 e = q8[wrapq(qhead + lane)];
        const uint32_t bit = e.y & 0x0FFFFFFFu;
        const uint32_t y = ((e.y >> 28) & 1u) | (((e.y >> 29) & 1u) << 16);
        if (general_test(s2, s1, e.x, y, e.y >> 30)) atomicOr(mask + (bit >> 5), 1u << (bit & 31));
      }
      qhead = wrapq(qhead + count); qtail -= count;          // in this mode qtail holds the LENGTH
    } else {
      if (lane < count) { const QEntry e = q[(qhead + lane) & 63u]; if (general_test(s2, s1, e.x, e.y, e.ref)) atomicOr(mask + (e.bit >> 5), 1u << (e.bit & 31)); }
      qhead += count;
    }
  };
  const uint32_t sh = 30u - ((lane & 15u) << 1), laneBit = 0x80000000u >> lane, ltMask = (1u << lane) - 1u;
  const uint32_t site0 = warp * (32 * ROWS) + lane, refIdx0 = site0 >> 4, bmIdx0 = warp * ROWS;
  const bool deadBit = s1[0] != 0;
  uint32_t k = 0, acc = 0;
  for (uint32_t tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++k) {
    const uint32_t stage = k % STAGES, parity = (k / STAGES) & 1;
    mbar_wait<SLEEP>(&full[stage], parity);
    const uint2 *sc = reinterpret_cast<const uint2 *>(sCnt + (size_t) stage * (TILE + 2)) + site0;
    const uint32_t *sr = sRef + stage * (TILE / 16) + refIdx0, *sb = sBm + stage * (TILE / 32) + bmIdx0;
    const uint32_t pT = pBeg + tile * TILE;
    const uint32_t lim = (pEnd - pT) < (uint32_t) TILE ? pEnd - pT : (uint32_t) TILE;
    if (MODE == 2) {
#pragma unroll
      for (int r = 0; r < ROWS; ++r) acc ^= sc[32 * r].x;
    } else {
      uint2 c[ROWS]; uint32_t nlo[ROWS], rw[ROWS], bw[ROWS];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) { c[r] = sc[32 * r]; if (!(OPT & 2)) nlo[r] = sc[32 * r + 1].x; rw[r] = sr[2 * r]; bw[r] = sb[r]; }
      uint32_t ind[ROWS + 1];
      if (OPT & 2) {
        const uint32_t xAfter = sc[32 * ROWS - lane].x;
#pragma unroll
        for (int r = 0; r < ROWS; ++r) ind[r] = __ballot_sync(0xffffffffu, ((c[r].x & 0xC000C000u) == 0xC000C000u) & ((c[r].x & 0x20002000u) != 0));
        ind[ROWS] = (((xAfter & 0xC000C000u) == 0xC000C000u) & ((xAfter & 0x20002000u) != 0)) ? 1u : 0u;
      }
      uint32_t *mw = mask + ((pT - pBeg) >> 5) + bmIdx0;
      const bool fullTile = (OPT & 1) && lim == (uint32_t) TILE;
      uint32_t words[ROWS];
      auto row = [&](int r, bool checkLim) {
        const uint32_t site = site0 + 32 * r;
        const uint32_t x = c[r].x, y = c[r].y;
        const uint32_t refx = rw[r] >> sh;
        const bool isRef1 = (((x >> 14) ^ refx) & 3u) == 0;
        const bool oneSlot = (y - 1u) < 0xFFFFu;
        const bool dead = y == 0;
        bool special = (x & 0xC000C000u) == 0xC000C000u;
        if (OPT & 2) special |= (__funnelshift_r(ind[r], ind[r + 1], 1) & (1u << lane)) != 0;
        else { const uint32_t n = nlo[r]; special |= ((n & 0xC000C000u) == 0xC000C000u) & ((n & 0x20002000u) != 0); }
        bool valid = (bw[r] & laneBit) == 0;
        if (checkLim) valid &= site < lim;
        const bool simple1 = oneSlot & isRef1;
        const bool bit = s1[x & 0x3FFFu] != 0;
        bool flag = valid & (special | (simple1 & bit) | (dead & deadBit));
        const bool cplx = valid & !special & !simple1 & !dead;
        if (MODE == 1) flag |= cplx;
        if (MODE == 4) {
          if (cplx) {
            if (depth < 4) { stk[32 * depth] = make_uint2(x, (pT - pBeg + site) | ((y & 0xFFFFu) ? 1u << 28 : 0u) | ((y >> 16) ? 1u << 29 : 0u) | (refx << 30)); ++depth; }
            else flag = true;                                   // stack full: let the exact evaluator decide
          }
        }
        if (MODE == 3) { if (cplx) flag = general_test(s2, s1, x, y, refx & 3u); }
        const uint32_t word = (OPT & 16) ? __reduce_or_sync(0xffffffffu, flag ? (1u << lane) : 0u) : __ballot_sync(0xffffffffu, flag);
        words[r] = word;
        if (!(OPT & 8)) { if (lane == 0 && (!checkLim || site < lim)) mw[r] = word; }
        if (MODE == 0) {
          const uint32_t cw = __ballot_sync(0xffffffffu, cplx);
          if (OPT & 4) {
            if (cplx) q8[wrapq(wrapq(qhead + qtail) + __popc(cw & ltMask))] = make_uint2(x, (pT - pBeg + site) | ((y & 0xFFFFu) ? 1u << 28 : 0u) | ((y >> 16) ? 1u << 29 : 0u) | (refx << 30));
            qtail += __popc(cw);
          } else if (cw) {
            if (cplx) { QEntry e; e.x = x; e.y = y; e.bit = pT - pBeg + site; e.ref = refx & 3u; q[(qtail + __popc(cw & ltMask)) & 63u] = e; }
            qtail += __popc(cw);
            __syncwarp();
            if (qtail - qhead >= 32u) { drain(32); __syncwarp(); }
          }
        }
      };
      if (fullTile) {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) row(r, false);
      } else {
#pragma unroll
        for (int r = 0; r < ROWS; ++r) row(r, true);
      }
      if (OPT & 8) {
        if (lane == 0) {
          if (fullTile || warp * (32 * ROWS) + 32 * ROWS <= lim) {
            if (ROWS == 4) *reinterpret_cast<uint4 *>(mw) = make_uint4(words[0], words[1], words[2], words[3]);
            else { for (int r = 0; r < ROWS; ++r) mw[r] = words[r]; }
          } else { for (int r = 0; r < ROWS; ++r) if (warp * (32 * ROWS) + 32 * r < lim) mw[r] = words[r]; }
        }
      }
      if (MODE == 0 && (OPT & 4)) { __syncwarp(); while (qtail >= 32u) { drain(32); __syncwarp(); } }
      if (MODE == 4) {
        __syncwarp();                                           // mask words of this tile are written
        if (depth) {                                            // one parked site per lane and tile
          --depth;
          const uint2 e = stk[32 * depth];
          const uint32_t bit = e.y & 0x0FFFFFFFu, y = ((e.y >> 28) & 1u) | (((e.y >> 29) & 1u) << 16);
          if (general_test(s2, s1, e.x, y, e.y >> 30)) atomicOr(mask + (bit >> 5), 1u << (bit & 31));
        }
      }
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      if (atomicAdd(&released[stage], 1u) == WARPS - 1) {
        released[stage] = 0;
        __threadfence_block();
        const uint32_t nt = tile + STAGES * gridDim.x;
        if (nt < nTiles) issue(stage, nt);
      }
    }
  }
  __syncwarp();
  if (MODE == 0) drain((OPT & 4) ? qtail : qtail - qhead);
  if (MODE == 4) {
    while (depth) {
      --depth;
      const uint2 e = stk[32 * depth];
      const uint32_t bit = e.y & 0x0FFFFFFFu, y = ((e.y >> 28) & 1u) | (((e.y >> 29) & 1u) << 16);
      if (general_test(s2, s1, e.x, y, e.y >> 30)) atomicOr(mask + (bit >> 5), 1u << (bit & 31));
    }
  }
  if (acc == 0x12345u) mask[0] = acc;
}

// LAYOUT B: every lane owns 4 CONSECUTIVE sites (two 16-byte shared loads, one reference word, one bitmap word),
// flag nibbles are merged into mask words with three shuffles, complex sites are pushed with a shared-memory
// atomic (no ballots at all).  QMODE 0: complex sites are only flagged (fast path alone)  1: queue via ATOMS
template <int THREADS, int TILE, int STAGES, int QMODE, int SLEEP>
__global__ void __launch_bounds__(THREADS, 1)
k_stream4(const uint64_t *__restrict__ counters, const uint32_t *__restrict__ packedRef, const uint32_t *__restrict__ invalid,
          uint32_t pBeg, uint32_t pEnd, const uint8_t *__restrict__ gs1, const uint32_t *__restrict__ gs2, uint32_t *__restrict__ mask) {
  constexpr int WARPS = THREADS / 32, GROUPS = TILE / (THREADS * 4);       // 4-site groups per lane and tile
  constexpr uint32_t QCAP = 256;
  extern __shared__ __align__(128) unsigned char raw[];
  uint64_t *sCnt = reinterpret_cast<uint64_t *>(raw);
  uint32_t *sRef = reinterpret_cast<uint32_t *>(sCnt + (size_t) STAGES * (TILE + 2));
  uint32_t *sBm = sRef + STAGES * (TILE / 16);
  uint8_t *s1 = reinterpret_cast<uint8_t *>(sBm + STAGES * (TILE / 32));
  uint32_t *s2 = reinterpret_cast<uint32_t *>(s1 + kS1);
  uint2 *queue = reinterpret_cast<uint2 *>(s2 + 2 * kS2Words);                            // [WARPS][QCAP]
  uint64_t *full = reinterpret_cast<uint64_t *>(queue + (QMODE ? WARPS * QCAP : 0));
  uint32_t *released = reinterpret_cast<uint32_t *>(full + STAGES);
  uint32_t *qcount = released + STAGES;                                                    // [WARPS] entries ever pushed
  const uint32_t nTiles = (pEnd - pBeg + TILE - 1) / TILE, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t kBytesC = 8 * (TILE + 2), kBytesR = TILE / 4, kBytesB = TILE / 8;
  auto issue = [&](uint32_t stage, uint32_t tile) {
    const uint32_t pT = pBeg + tile * TILE;
    mbar_expect_tx(&full[stage], kBytesC + kBytesR + kBytesB);
    bulk_g2s(sCnt + (size_t) stage * (TILE + 2), counters + pT, kBytesC, &full[stage]);
    bulk_g2s(sRef + stage * (TILE / 16), packedRef + (pT >> 4), kBytesR, &full[stage]);
    bulk_g2s(sBm + stage * (TILE / 32), invalid + (pT >> 5), kBytesB, &full[stage]);
  };
  if (tid == 0) { for (int s = 0; s < STAGES; ++s) { mbar_init(&full[s], 1); released[s] = 0; } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (tid < WARPS) qcount[tid] = 0;
  for (uint32_t k = tid; k < kS1 / 4; k += THREADS) reinterpret_cast<uint32_t *>(s1)[k] = reinterpret_cast<const uint32_t *>(gs1)[k];
  for (uint32_t k = tid; k < 2 * kS2Words; k += THREADS) s2[k] = gs2[k];
  __syncthreads();
  if (tid == 0) for (uint32_t s = 0; s < (uint32_t) STAGES; ++s) { uint32_t t = blockIdx.x + s * gridDim.x; if (t < nTiles) issue(s, t); }
  uint2 *q = queue + warp * QCAP;
  uint32_t qhead = 0;
  auto drain = [&](uint32_t count) {
    if (lane < count) {
      const uint2 e = q[(qhead + lane) & (QCAP - 1)];
      const uint32_t bit = e.y & 0x0FFFFFFFu, y = ((e.y >> 28) & 1u) | (((e.y >> 29) & 1u) << 16);
      if (general_test(s2, s1, e.x, y, e.y >> 30)) atomicOr(mask + (bit >> 5), 1u << (bit & 31));
    }
    qhead += count;
  };
  const uint32_t shb = 30u - ((lane & 3u) << 3);            // shift of this lane's first site in its reference word
  const uint32_t bmb = 0x80000000u >> ((lane & 7u) << 2);   // bitmap bit of this lane's first site
  const bool deadBit = s1[0] != 0;
  uint32_t k = 0;
  for (uint32_t tile = blockIdx.x; tile < nTiles; tile += gridDim.x, ++k) {
    const uint32_t stage = k % STAGES, parity = (k / STAGES) & 1;
    mbar_wait<SLEEP>(&full[stage], parity);
    const uint32_t pT = pBeg + tile * TILE;
    const uint32_t lim = (pEnd - pT) < (uint32_t) TILE ? pEnd - pT : (uint32_t) TILE;
#pragma unroll
    for (int g = 0; g < GROUPS; ++g) {
      const uint32_t base = (warp * GROUPS + g) * 128 + 4 * lane;                          // first of this lane's 4 sites
      const uint64_t *sc = sCnt + (size_t) stage * (TILE + 2) + base;
      const uint4 c01 = *reinterpret_cast<const uint4 *>(sc), c23 = *reinterpret_cast<const uint4 *>(sc + 2);
      const uint32_t xn = reinterpret_cast<const uint2 *>(sc + 4)->x;
      const uint32_t rwd = sRef[stage * (TILE / 16) + (base >> 4)], bwd = sBm[stage * (TILE / 32) + (base >> 5)];
      const uint32_t xs[5] = {c01.x, c01.z, c23.x, c23.z, xn}, ys[4] = {c01.y, c01.w, c23.y, c23.w};
      uint32_t nib = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t x = xs[j], y = ys[j], n = xs[j + 1];
        const uint32_t refx = rwd >> (shb - 2 * j);
        const bool isRef1 = (((x >> 14) ^ refx) & 3u) == 0;
        const bool oneSlot = (y - 1u) < 0xFFFFu;
        const bool dead = y == 0;
        const bool special = ((x & 0xC000C000u) == 0xC000C000u) | (((n & 0xC000C000u) == 0xC000C000u) & ((n & 0x20002000u) != 0));
        const bool valid = (base + j < lim) & ((bwd & (bmb >> j)) == 0);
        const bool simple1 = oneSlot & isRef1;
        const bool bit = s1[x & 0x3FFFu] != 0;
        bool flag = valid & (special | (simple1 & bit) | (dead & deadBit));
        const bool cplx = valid & !special & !simple1 & !dead;
        if (QMODE == 0) flag |= cplx;
        else if (cplx) {
          const uint32_t pos = atomicAdd(&qcount[warp], 1u);
          q[pos & (QCAP - 1)] = make_uint2(x, (pT - pBeg + base + j) | ((y & 0xFFFFu) ? 1u << 28 : 0u) | ((y >> 16) ? 1u << 29 : 0u) | (refx << 30));
        }
        nib |= flag ? (1u << j) : 0u;
      }
      uint32_t v = nib << ((lane & 7u) << 2);
      v |= __shfl_xor_sync(0xffffffffu, v, 1);
      v |= __shfl_xor_sync(0xffffffffu, v, 2);
      v |= __shfl_xor_sync(0xffffffffu, v, 4);
      if ((lane & 7u) == 0 && base < lim) mask[((pT - pBeg + base) >> 5)] = v;
    }
    __syncwarp();
    if (lane == 0) {
      __threadfence_block();
      if (atomicAdd(&released[stage], 1u) == WARPS - 1) {
        released[stage] = 0;
        __threadfence_block();
        const uint32_t nt = tile + STAGES * gridDim.x;
        if (nt < nTiles) issue(stage, nt);
      }
    }
    if (QMODE) {
      const uint32_t qt = *reinterpret_cast<volatile uint32_t *>(&qcount[warp]);
      while (qt - qhead >= 32u) { drain(32); __syncwarp(); }
    }
  }
  __syncwarp();
  if (QMODE) drain(*reinterpret_cast<volatile uint32_t *>(&qcount[warp]) - qhead);
}

template <int THREADS, int TILE, int STAGES, int MODE, int OPT>
size_t smem_bytes() {
  return (size_t) STAGES * (TILE + 2) * 8 + STAGES * (TILE / 16) * 4 + STAGES * (TILE / 32) * 4 + kS1 + 2 * kS2Words * 4 +
         (MODE == 0 ? ((OPT & 4) ? (THREADS / 32) * (32 * (TILE / THREADS) + 32) * 8 : (THREADS / 32) * 64 * sizeof(QEntry)) : MODE == 4 ? (THREADS / 32) * 128 * 8 : 0) + STAGES * 8 + STAGES * 4 + 64;
}

static uint64_t rng_state = 88172645463325252ull;
static inline uint64_t rnd() { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return rng_state; }

int main() {
  const uint32_t L = 64444160;                        // multiple of 128
  const size_t pad = 16384;
  std::vector<uint64_t> h((size_t) L + pad, 0);
  std::vector<uint32_t> ref(((size_t) L + pad) / 16 + 16, 0), inv(((size_t) L + pad) / 32 + 16, 0);
  for (size_t w = 0; w < ref.size(); ++w) ref[w] = (uint32_t) rnd();
  for (uint32_t p = 0; p < L; ++p) {
    const uint32_t rb = (ref[p >> 4] >> (30 - ((p & 15) << 1))) & 3;
    const uint32_t u = (uint32_t) (rnd() % 10000);
    const uint32_t depth = 20 + (uint32_t) (rnd() % 21), pos = depth / 2, neg = depth - pos, w = 4 * depth - (uint32_t) (rnd() % 8);
    uint64_t c;
    if (u < 30) c = 0xC000ull | (0xC000ull << 16) | ((uint64_t) (rnd() % 1000) << 32) | ((rnd() & 1) ? 0x2000ull : 0);   // overflow type
    else if (u < 890) {                                  // ref + one error allele
      const uint32_t ab = (rb + 1 + (uint32_t) (rnd() % 3)) & 3;
      c = (uint64_t) ((rb << 14) | w) | ((uint64_t) ((ab << 14) | 4) << 16) | ((uint64_t) pos << 32) | ((uint64_t) neg << 40) | (1ull << 48);
    } else if (u < 900) c = 0;                            // no coverage
    else c = (uint64_t) ((rb << 14) | w) | ((uint64_t) pos << 32) | ((uint64_t) neg << 40);
    h[p] = c;
  }
  std::vector<uint8_t> s1(kS1, 0);
  for (int i = 1; i < 12; ++i) s1[i] = 1;
  std::vector<uint32_t> s2(2 * kS2Words, 0);
  for (auto &v : s2) v = (uint32_t) (rnd() & rnd() & rnd() & rnd());
  uint64_t *dC; uint32_t *dR, *dI, *dS2, *dM; uint8_t *dS1;
  cudaMalloc(&dC, h.size() * 8); cudaMalloc(&dR, ref.size() * 4); cudaMalloc(&dI, inv.size() * 4); cudaMalloc(&dS1, kS1);
  cudaMalloc(&dS2, s2.size() * 4); cudaMalloc(&dM, (size_t) L / 8 + 4096);
  cudaMemcpy(dC, h.data(), h.size() * 8, cudaMemcpyHostToDevice); cudaMemcpy(dR, ref.data(), ref.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dI, inv.data(), inv.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(dS1, s1.data(), kS1, cudaMemcpyHostToDevice);
  cudaMemcpy(dS2, s2.data(), s2.size() * 4, cudaMemcpyHostToDevice);
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  const double bytes = (double) L * 8.375;
  std::vector<uint32_t> refMask((size_t) L / 32), gotMask((size_t) L / 32);
  bool haveRef = false;
#define RUN(THREADS, TILE, STAGES, MODE, SLEEP, MINB, CTAS, OPT)                                                         \
  do {                                                                                                              \
    auto kern = k_stream<THREADS, TILE, STAGES, MODE, SLEEP, MINB, OPT>;                                                 \
    size_t sm = smem_bytes<THREADS, TILE, STAGES, MODE, OPT>();                                                          \
    if (sm * CTAS > 227 * 1024) { printf("thr %4d tile %5d st %2d mode %d sleep %3d ctas %d: smem %zu too big\n", THREADS, TILE, STAGES, MODE, SLEEP, CTAS, sm); break; } \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);                              \
    cudaMemset(dM, 0, (size_t) L / 8);                                                                              \
    for (int w = 0; w < 2; ++w) kern<<<148 * CTAS, THREADS, sm>>>(dC, dR, dI, 0, L, dS1, dS2, dM);                    \
    cudaEventRecord(a);                                                                                             \
    for (int r = 0; r < 10; ++r) kern<<<148 * CTAS, THREADS, sm>>>(dC, dR, dI, 0, L, dS1, dS2, dM);                   \
    cudaEventRecord(b); cudaEventSynchronize(b);                                                                    \
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;                                                            \
    cudaError_t e = cudaGetLastError();                                                                             \
    const char *chk = "";                                                                                           \
    if ((MODE == 0 || MODE == 3 || MODE == 4) && e == cudaSuccess) {                                                             \
      cudaMemcpy(gotMask.data(), dM, (size_t) L / 8, cudaMemcpyDeviceToHost);                                       \
      if (!haveRef) { refMask = gotMask; haveRef = true; chk = "(reference mask)"; }                               \
      else chk = (refMask == gotMask) ? "mask ok" : "MASK DIFFERS";                                                 \
    }                                                                                                               \
    printf("opt %2d thr %4d tile %5d st %2d mode %d sleep %3d ctas %d: %7.3f ms %7.1f GB/s %4.1f%% %s %s\n", OPT, THREADS, TILE, STAGES, MODE, SLEEP, \
           CTAS, ms, bytes / (ms * 1e-3) / 1e9, 100 * bytes / (ms * 1e-3) / 1e9 / 6541.5, chk, e == cudaSuccess ? "" : cudaGetErrorString(e)); \
    if (e != cudaSuccess) return 1;                                                                                 \
  } while (0)
  // shipped geometry first (its mask is the reference for the others)
  RUN(1024, 4096, 4, 0, 40, 1, 1, 1);
  RUN(1024, 4096, 4, 1, 40, 1, 1, 1);
  RUN(1024, 4096, 4, 4, 40, 1, 1, 1);
  RUN(1024, 4096, 4, 4, 40, 1, 1, 9);
  RUN(512, 4096, 5, 4, 40, 1, 1, 9);
  RUN(1024, 4096, 5, 4, 40, 1, 1, 9);
#define RUN4(THREADS, TILE, STAGES, QMODE, SLEEP)                                                                   \
  do {                                                                                                              \
    auto kern = k_stream4<THREADS, TILE, STAGES, QMODE, SLEEP>;                                                     \
    size_t sm = (size_t) STAGES * (TILE + 2) * 8 + STAGES * (TILE / 16) * 4 + STAGES * (TILE / 32) * 4 + kS1 + 2 * kS2Words * 4 + \
                (QMODE ? (THREADS / 32) * 256 * 8 : 0) + STAGES * 12 + (THREADS / 32) * 4 + 64;                     \
    if (sm > 227 * 1024) { printf("layoutB thr %d tile %d st %d q %d: smem %zu too big\n", THREADS, TILE, STAGES, QMODE, sm); break; } \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) sm);                              \
    cudaMemset(dM, 0, (size_t) L / 8);                                                                              \
    for (int w = 0; w < 2; ++w) kern<<<148, THREADS, sm>>>(dC, dR, dI, 0, L, dS1, dS2, dM);                           \
    cudaEventRecord(a);                                                                                             \
    for (int r = 0; r < 10; ++r) kern<<<148, THREADS, sm>>>(dC, dR, dI, 0, L, dS1, dS2, dM);                          \
    cudaEventRecord(b); cudaEventSynchronize(b);                                                                    \
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;                                                            \
    cudaError_t e = cudaGetLastError();                                                                             \
    const char *chk = "";                                                                                           \
    if (QMODE && e == cudaSuccess) { cudaMemcpy(gotMask.data(), dM, (size_t) L / 8, cudaMemcpyDeviceToHost); chk = (refMask == gotMask) ? "mask ok" : "MASK DIFFERS"; } \
    printf("layoutB thr %4d tile %5d st %2d queue %d sleep %3d: %7.3f ms %7.1f GB/s %4.1f%% %s %s\n", THREADS, TILE, STAGES, QMODE, SLEEP, ms, \
           bytes / (ms * 1e-3) / 1e9, 100 * bytes / (ms * 1e-3) / 1e9 / 6541.5, chk, e == cudaSuccess ? "" : cudaGetErrorString(e)); \
    if (e != cudaSuccess) return 1;                                                                                 \
  } while (0)
  return 0;
}
