// gt16_synth -- synthetic reference index + bam2snapshot-format snapshot generator.
//
// TEST / BENCH INFRASTRUCTURE (not part of the calling hot path).  It writes exactly the
// on-disk formats the reference tools produce, so that BOTH the reference caller
// (oracle/_ref/ref_caller) and this repository's snapshotSnpcaller can consume them:
//
//   <index>.pac/.ann/.amb/.tra   2BWT-lib packed reference (formats: readIndex.cpp:24-60,96-115;
//                                 verified byte-for-byte against HSPParseFASTAToPacked in tests)
//   <out>.snapshot                5-subfile container of SnapshotHandler.cpp:68-236
//                                 (counter encoding of coreStruct.h:10-67 / SNP.cpp:745-1688)
//   region.bin                    u32 n; {u32 start,end}[n]  (buildRegionList.cpp:221-222)
//
// The data model is "per-read" so the counters look like what bam2snapshot accumulates:
// every position draws a Poisson depth, every read a strand, a base-quality weight 0..4
// (thresholds 10/13/17/20 of snapshotSnpcaller.ini:15-18) and a base (reference, planted
// variant allele, or sequencing error).  A counter is promoted to the overflow array under
// the same conditions as the producer (third allele, strand count reaching 255, or an
// insertion/deletion record, SNP.cpp:842-845,862,1166-1188).
//
// Everything is a pure function of (seed, position): output does not depend on --threads.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <string>
#include <vector>
#include <thread>
#include <algorithm>

typedef uint32_t u32;
typedef uint64_t u64;
typedef uint16_t u16;
typedef uint8_t u8;

static const u32 BATCH = 8388608u;  // VariantCaller.cpp:526-528

struct OvfRec {  // SnpOverflowCounter, coreStruct.h:34-67 (32 bytes)
  u16 W[4], F[4], R[4];
  u8 ins[4];  // insertSeq[3], lqCount | u32 ptr
  u8 del[4];  // deletionLength, hqCount, lqCount, pad | u32 ptr
};
static_assert(sizeof(OvfRec) == 32, "overflow record must be 32 bytes");

struct Contig { std::string name; u32 len; };
struct Region { u32 start, end; };

struct Params {
  std::vector<Contig> contigs;
  double depth = 30.0;
  u64 seed = 1;
  double snpRate = 1e-3, indelRate = 1.25e-4, errRate = 3e-3, lowQualFrac = 0.10;
  double longIndelFrac = 0.03, multiPatFrac = 0.15;
  u32 ovfRegions = 16;          // bam2snapshot.ini NumOfCpuThreads
  u32 nGaps = 0, gapLen = 100000;
  double hotspotRate = 0.0;     // per-position probability of a depth hotspot start (config 5)
  u32 hotspotLen = 40;
  double hotspotDepth = 1500.0;
  double invalidRate = 0.0;     // extra invalid bits
  bool exome = false;
  u32 nTargets = 0, targetLen = 300;
  int threads = 8;
  std::string outIndex, outSnapshot, outRegion, outFasta;
};

// ---------------------------------------------------------------- RNG (stateless mix)
static inline u64 mix64(u64 x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
struct Rng {
  u64 s;
  Rng(u64 seed, u64 stream, u64 pos) { s = mix64(seed ^ mix64(stream * 0x632BE59BD9B4E019ull + pos)); }
  inline u64 next() { s += 0x9E3779B97F4A7C15ull; u64 z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
                      z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
  inline double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
  inline u32 below(u32 n) { return (u32) ((next() >> 32) * (u64) n >> 32); }
};

static u32 poisson(Rng &r, double lambda) {
  if (lambda <= 0) return 0;
  if (lambda > 60) {  // normal approximation is plenty for a synthetic depth profile
    double u1 = r.uni() + 1e-300, u2 = r.uni();
    double z = sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
    double v = lambda + sqrt(lambda) * z + 0.5;
    return v < 0 ? 0 : (u32) v;
  }
  double L = exp(-lambda), p = 1.0;
  u32 k = 0;
  do { ++k; p *= r.uni(); } while (p > L);
  return k - 1;
}

// ---------------------------------------------------------------- variant plan
struct Variant {
  u8 kind;     // 0 none, 1 SNP, 2 insertion, 3 deletion (anchored at p, record lives at counter p+1)
  u8 alt;      // SNP alt base
  u8 len;      // indel length
  u8 hom;      // 1 = homozygous
  u8 second;   // 1 = add a second (noise) pattern at the same counter
  u8 secondLen;
  u8 secondKind;  // 2 / 3
};

struct Genome {
  u32 L = 0;
  std::vector<u8> base;          // 1 byte per base (0..3)
  std::vector<u32> contigStart;  // per contig
  std::vector<std::pair<u32, u32> > gaps;
};

static inline bool in_gap(const Genome &g, u32 p) {
  for (size_t i = 0; i < g.gaps.size(); ++i) if (p >= g.gaps[i].first && p < g.gaps[i].second) return true;
  return false;
}

static Variant variant_at(const Params &P, const Genome &G, u32 p) {
  Variant v; memset(&v, 0, sizeof(v));
  Rng r(P.seed, 0x56415231, p);
  double u = r.uni();
  if (u < P.snpRate) {
    v.kind = 1; v.alt = (G.base[p] + 1 + r.below(3)) & 3; v.hom = r.below(3) == 0;
    return v;
  }
  if (u < P.snpRate + P.indelRate) {
    v.kind = r.below(2) ? 2 : 3;
    v.len = (r.uni() < P.longIndelFrac) ? (u8) (13 + r.below(18)) : (u8) (1 + std::min(r.below(12), r.below(12)));
    v.hom = r.below(3) == 0;
    if (r.uni() < P.multiPatFrac) {
      v.second = 1; v.secondKind = r.below(2) ? 2 : 3;
      v.secondLen = (u8) (1 + r.below(v.secondKind == v.kind && v.len == 1 ? 6 : 8));
      if (v.secondKind == v.kind && v.secondLen == v.len) v.secondLen = v.len + 1;
    }
    // keep clear of: contig ends, overflow-region starts (reference indexes the wrong region's
    // array there, VariantCaller.cpp:47,150,268), and batch tails (window loops run past the
    // mwInfo allocation for deletions near a batch end, SNP_Caller.cpp:301).
    u32 q = p + 1, span = (u32) std::max(v.len, v.secondLen) + 2;
    if ((u64) q + span + 300 >= G.L) { v.kind = 0; return v; }
    u32 regLen = G.L / P.ovfRegions;
    if (regLen && q % regLen == 0 && q / regLen < P.ovfRegions) { v.kind = 0; return v; }
    if ((p % BATCH) + 600 > BATCH) { v.kind = 0; return v; }
    return v;
  }
  return v;
}

// ---------------------------------------------------------------- snapshot in memory
struct Snapshot {
  u32 L = 0;
  std::vector<u64> counters;               // SnpCounter as little-endian 8-byte words
  std::vector<u32> invalid;
  std::vector<std::vector<OvfRec> > ovf;   // per overflow region
  std::vector<u8> pool;
};

static inline u32 region_of(u32 p, u32 L, u32 n) {  // findRegionByPosition, coreStruct.h:141-154
  u32 i = 0;
  for (; i + 1 < n; ++i) if (p < (L / n) * (i + 1)) return i;
  return i;
}
static inline u64 pack_compact(u16 w1, u16 w2, u8 p1, u8 n1, u8 p2, u8 n2) {
  return (u64) w1 | ((u64) w2 << 16) | ((u64) p1 << 32) | ((u64) n1 << 40) | ((u64) p2 << 48) | ((u64) n2 << 56);
}
static inline u64 pack_overflow(u16 w1, u16 w2, u32 idx) { return (u64) w1 | ((u64) w2 << 16) | ((u64) idx << 32); }

struct Pattern { u8 kind; u8 len; u8 seq[8]; u16 hq, lq; };

struct ChunkOut {  // produced per worker chunk, merged in order afterwards
  std::vector<std::pair<u32, OvfRec> > ovf;  // (position, record); ptr fields hold chunk-local pool offsets
  std::vector<u8> poolBytes;
  std::vector<std::pair<u32, u8> > poolFix;  // (index into ovf, 0 = insertion ptr / 1 = deletion ptr)
};

static double coverage_at(const Params &P, const Genome &G, const std::vector<Region> &targets, u32 p,
                          size_t &tcur) {
  if (in_gap(G, p)) return 0.0;
  if (!P.exome) return 1.0;
  while (tcur < targets.size() && p >= targets[tcur].end + 200) ++tcur;
  if (tcur >= targets.size()) return 0.0;
  const Region &t = targets[tcur];
  if (p >= t.start && p < t.end) return 1.0;
  if (p < t.start) { u32 d = t.start - p; return d >= 150 ? 0.0 : 1.0 - d / 150.0; }
  u32 d = p - t.end + 1; return d >= 150 ? 0.0 : 1.0 - d / 150.0;
}

static void gen_chunk(const Params &P, const Genome &G, const std::vector<Region> &targets, u32 a, u32 b,
                      Snapshot &S, ChunkOut &out) {
  // carry state: an active planted deletion covering positions [delFrom, delTo) with carrier fraction
  u32 delTo = 0; double delFrac = 0;
  Variant pending; memset(&pending, 0, sizeof(pending)); u32 pendingAnchor = 0xFFFFFFFFu;
  u32 hotTo = 0;
  size_t tcur = 0;
  if (P.exome) { while (tcur < targets.size() && targets[tcur].end + 200 <= (a > 400 ? a - 400 : 0)) ++tcur; }
  u32 lookback = a > 64 ? a - 64 : 0;
  for (u32 p = lookback; p < b; ++p) {
    bool emit = p >= a;
    // --- planning (also run over the look-back window so the carried state is right)
    Variant v; memset(&v, 0, sizeof(v));
    bool covered = p < delTo;
    if (!covered && pendingAnchor != p - 1) v = variant_at(P, G, p);
    if (P.hotspotRate > 0) { Rng hr(P.seed, 0x484F54, p); if (hr.uni() < P.hotspotRate) hotTo = p + P.hotspotLen; }
    Variant here = (pendingAnchor + 1 == p) ? pending : Variant();
    if (pendingAnchor + 1 != p) memset(&here, 0, sizeof(here));
    if (v.kind >= 2) { pending = v; pendingAnchor = p; }
    if (here.kind == 3) { delTo = p + here.len; delFrac = here.hom ? 0.97 : 0.5; covered = true; }
    if (!emit) continue;

    double cov = coverage_at(P, G, targets, p, tcur);
    Rng r(P.seed, 0x53495445, p);
    double lam = P.depth * cov;
    bool hot = p < hotTo && cov > 0;
    if (hot) lam = P.hotspotDepth;
    u32 d = poisson(r, lam);
    if (d > 60000) d = 60000;
    u8 ref = G.base[p];
    u32 W[4] = {0, 0, 0, 0}, F[4] = {0, 0, 0, 0}, Rv[4] = {0, 0, 0, 0};
    u8 order[4]; int norder = 0;
    double altFrac = v.kind == 1 ? (v.hom ? 0.985 : 0.5) : 0.0;
    u32 insHq = 0, insLq = 0, delHq = 0, delLq = 0, ins2 = 0, del2 = 0;
    for (u32 k = 0; k < d; ++k) {
      u64 x = r.next();
      u32 strand = x & 1;
      double uq = ((x >> 1) & 0xFFFFF) * (1.0 / 1048576.0);
      u32 w = uq < P.lowQualFrac ? (u32) ((x >> 21) & 3) : 4;
      double ub = ((x >> 24) & 0xFFFFF) * (1.0 / 1048576.0);
      double uv = ((x >> 44) & 0xFFFFF) * (1.0 / 1048576.0);
      // indel evidence carried by this read is recorded at counter p (anchor p-1)
      if (here.kind == 2 && uv < (here.hom ? 0.95 : 0.47)) { if (w > 1) ++insHq; else ++insLq; }
      if (here.kind == 3 && uv < delFrac) { if (w > 1) ++delHq; else ++delLq; }
      else if (here.second && uv > 0.93) { if (here.secondKind == 2) ++ins2; else ++del2; }
      if (covered && uv < delFrac) continue;  // read carries the deletion: no base here
      u8 bse = ref;
      if (ub < P.errRate) bse = (ref + 1 + ((x >> 40) % 3)) & 3;
      else if (v.kind == 1 && uv < altFrac) bse = v.alt;
      if (!(F[bse] | Rv[bse])) order[norder++] = bse;
      W[bse] += w; if (strand) ++F[bse]; else ++Rv[bse];
    }
    // --- indel patterns recorded at this counter
    std::vector<Pattern> insP, delP;
    if (here.kind == 2 && insHq + insLq) {
      Pattern q; memset(&q, 0, sizeof(q)); q.kind = 'I'; q.len = here.len; q.hq = (u16) insHq; q.lq = (u16) insLq;
      Rng sr(P.seed, 0x494E53, p);
      for (u32 i = 0; i < q.len; ++i) q.seq[i >> 2] |= (u8) (sr.below(4) << ((i & 3) << 1));
      insP.push_back(q);
    }
    if (here.kind == 3 && delHq + delLq) {
      Pattern q; memset(&q, 0, sizeof(q)); q.kind = 'D'; q.len = here.len; q.hq = (u16) delHq; q.lq = (u16) delLq;
      delP.push_back(q);
    }
    if (here.second && (ins2 | del2) && (insP.size() + delP.size())) {
      Pattern q; memset(&q, 0, sizeof(q)); q.len = here.secondLen;
      if (here.secondKind == 2 && ins2) {
        q.kind = 'I'; q.hq = (u16) ins2; q.lq = (u16) (ins2 > 2 ? 1 : 0);
        Rng sr(P.seed, 0x494E54, p);
        for (u32 i = 0; i < q.len; ++i) q.seq[i >> 2] |= (u8) (sr.below(4) << ((i & 3) << 1));
        insP.push_back(q);
      } else if (here.secondKind == 3 && del2) {
        q.kind = 'D'; q.hq = (u16) del2; q.lq = 0; delP.push_back(q);
      }
    }
    // --- invalid positions (weighted count saturating the 14-bit compact field, SNP.cpp:873-884)
    bool invalid = false;
    if (P.invalidRate > 0) { Rng ir(P.seed, 0x494E56, p); invalid = ir.uni() < P.invalidRate; }
    for (int k = 0; k < 4; ++k) if (W[k] >= 0xFFF0u || F[k] >= 0xFFF0u || Rv[k] >= 0xFFF0u) invalid = true;
    if (invalid && !P.exome) S.invalid[p >> 5] |= 1u << (31 - (p & 31));

    u32 regLen = G.L / P.ovfRegions;
    bool regionStart = regLen && p != 0 && p % regLen == 0 && p / regLen < P.ovfRegions;
    if (regionStart) {
      // the reference looks this counter up in the PREVIOUS region's overflow array when it
      // decodes position p-1 (VariantCaller.cpp:47,129-132); keep it compact so that read is defined
      norder = std::min(norder, 2);
      for (int k = 0; k < 4; ++k) { F[k] = std::min(F[k], 254u); Rv[k] = std::min(Rv[k], 254u); W[k] = std::min(W[k], 0x3FFEu); }
      insP.clear(); delP.clear();
    }
    bool needOvf = norder > 2 || !insP.empty() || !delP.empty();
    for (int k = 0; k < 4; ++k) if (F[k] >= 255 || Rv[k] >= 255 || W[k] >= 0x3FFF) needOvf = true;
    if (!needOvf) {
      u16 w1 = 0, w2 = 0; u8 p1 = 0, n1 = 0, p2 = 0, n2 = 0;
      if (norder >= 1) { u8 bs = order[0]; w1 = (u16) ((bs << 14) | (W[bs] & 0x3FFF)); p1 = (u8) F[bs]; n1 = (u8) Rv[bs]; }
      if (norder >= 2) { u8 bs = order[1]; w2 = (u16) ((bs << 14) | (W[bs] & 0x3FFF)); p2 = (u8) F[bs]; n2 = (u8) Rv[bs]; }
      // the producer can leave slot 1 empty and slot 2 live only transiently; both orders of
      // (ref, alt) occur naturally through `order`
      S.counters[p] = pack_compact(w1, w2, p1, n1, p2, n2);
      continue;
    }
    OvfRec rec; memset(&rec, 0, sizeof(rec));
    for (int k = 0; k < 4; ++k) { rec.W[k] = (u16) std::min(W[k], 0xFFFFu); rec.F[k] = (u16) F[k]; rec.R[k] = (u16) Rv[k]; }
    u16 w1 = 0xC000, w2 = 0xC000;
    if (!insP.empty()) {
      w1 |= 0x2000;
      if (insP.size() == 1 && insP[0].len <= 12 && insP[0].hq < 255 && insP[0].lq < 255) {
        w1 |= (u16) ((insP[0].len << 8) + insP[0].hq);          // SNP.cpp:1198-1206
        rec.ins[0] = insP[0].seq[0]; rec.ins[1] = insP[0].seq[1]; rec.ins[2] = insP[0].seq[2];
        rec.ins[3] = (u8) insP[0].lq;
      } else {
        w1 |= 0x1000;
        u32 off = (u32) out.poolBytes.size(), n = 0;
        for (size_t i = 0; i < insP.size(); ++i) {              // SNP.cpp:1222-1231 record layout
          out.poolBytes.push_back(insP[i].len); ++n;
          for (u32 k = 0; k < (u32) (insP[i].len + 3) / 4; ++k) { out.poolBytes.push_back(insP[i].seq[k]); ++n; }
          out.poolBytes.push_back(insP[i].hq & 0xff); out.poolBytes.push_back(insP[i].hq >> 8);
          out.poolBytes.push_back(insP[i].lq & 0xff); out.poolBytes.push_back(insP[i].lq >> 8); n += 4;
        }
        w1 |= (u16) (n & 0xFFF);
        memcpy(rec.ins, &off, 4);
        out.poolFix.push_back(std::make_pair((u32) out.ovf.size(), (u8) 0));
        while (out.poolBytes.size() & 15) out.poolBytes.push_back(0);
      }
    }
    if (!delP.empty()) {
      w2 |= 0x2000;
      if (delP.size() == 1 && delP[0].hq < 255 && delP[0].lq < 255) {
        rec.del[0] = delP[0].len; rec.del[1] = (u8) delP[0].hq; rec.del[2] = (u8) delP[0].lq;   // SNP.cpp:1517-1519
      } else {
        w2 |= 0x1000;
        u32 off = (u32) out.poolBytes.size(), n = 0;
        for (size_t i = 0; i < delP.size(); ++i) {              // SNP.cpp:1544-1558 record layout
          out.poolBytes.push_back(delP[i].len);
          out.poolBytes.push_back(delP[i].hq & 0xff); out.poolBytes.push_back(delP[i].hq >> 8);
          out.poolBytes.push_back(delP[i].lq & 0xff); out.poolBytes.push_back(delP[i].lq >> 8); n += 5;
        }
        w2 |= (u16) (n & 0xFFF);
        memcpy(rec.del, &off, 4);
        out.poolFix.push_back(std::make_pair((u32) out.ovf.size(), (u8) 1));
        while (out.poolBytes.size() & 15) out.poolBytes.push_back(0);
      }
    }
    S.counters[p] = pack_overflow(w1, w2, 0);  // array index patched at merge time
    out.ovf.push_back(std::make_pair(p, rec));
  }
}

// ---------------------------------------------------------------- file writers
static void write_index(const Params &P, const Genome &G) {
  const std::string &pre = P.outIndex;
  FILE *f = fopen((pre + ".ann").c_str(), "w");
  fprintf(f, "%u %d %u\n", G.L, (int) P.contigs.size(), 0u);
  for (size_t i = 0; i < P.contigs.size(); ++i) {
    fprintf(f, "%u %s\n", 0u, P.contigs[i].name.c_str());
    fprintf(f, "%u %u %d\n", G.contigStart[i], P.contigs[i].len, 0);
  }
  fclose(f);
  f = fopen((pre + ".amb").c_str(), "w");
  fprintf(f, "%u %d %d\n", G.L, (int) P.contigs.size(), 0);
  fclose(f);
  // .tra: header, grid (one line per 2^18-bp cell: index of the last translate entry that starts
  // at or before the end of the cell), translate entries, then actual (start, length) lines
  f = fopen((pre + ".tra").c_str(), "w");
  u32 nGrid = (u32) (((u64) G.L + (1u << 18) - 1) >> 18);
  fprintf(f, "%u %d %d %u\n", G.L, (int) P.contigs.size(), 0, nGrid);
  size_t j = 0;
  for (u32 c = 0; c < nGrid; ++c) {
    u64 cellEnd = ((u64) (c + 1) << 18) - 1;
    if (cellEnd >= G.L) cellEnd = G.L - 1;
    while (j + 1 < P.contigs.size() && G.contigStart[j + 1] <= cellEnd) ++j;
    fprintf(f, "%u\n", (u32) j);
  }
  for (size_t i = 0; i < P.contigs.size(); ++i)
    fprintf(f, "%u %hu %u\n", G.contigStart[i], (unsigned short) (i + 1), (u32) (G.contigStart[i] - 1u));
  for (size_t i = 0; i < P.contigs.size(); ++i) fprintf(f, "%u %u\n", G.contigStart[i], P.contigs[i].len);
  fclose(f);
  // .pac: 4 bases per byte, first base in bits 7..6; last byte = L % 4 (an all-zero byte precedes it
  // when L % 4 == 0)
  f = fopen((pre + ".pac").c_str(), "wb");
  std::vector<u8> buf(((u64) G.L + 3) / 4 + 2, 0);
  for (u32 p = 0; p < G.L; ++p) buf[p >> 2] |= (u8) (G.base[p] << (6 - ((p & 3) << 1)));
  size_t n = G.L / 4 + 1;  // full bytes + (partial | zero) byte
  buf[n] = (u8) (G.L % 4);
  fwrite(buf.data(), 1, n + 1, f);
  fclose(f);
  if (!P.outFasta.empty()) {
    f = fopen(P.outFasta.c_str(), "w");
    static const char sym[] = "ACGT";
    for (size_t i = 0; i < P.contigs.size(); ++i) {
      fprintf(f, ">%s\n", P.contigs[i].name.c_str());
      std::string line;
      for (u32 k = 0; k < P.contigs[i].len; ++k) {
        line.push_back(sym[G.base[G.contigStart[i] + k]]);
        if (line.size() == 60) { fprintf(f, "%s\n", line.c_str()); line.clear(); }
      }
      if (!line.empty()) fprintf(f, "%s\n", line.c_str());
    }
    fclose(f);
  }
}

static void write_snapshot(const Params &P, const Snapshot &S, const std::vector<Region> &regions) {
  const size_t page = 4096;
  FILE *f = fopen((P.outSnapshot + ".snapshot").c_str(), "wb");
  if (!f) { fprintf(stderr, "cannot write %s.snapshot\n", P.outSnapshot.c_str()); exit(1); }
  // sparse position list, SnapshotHandler.cpp:19-38
  std::vector<u32> sparsePos;
  if (P.exome) {
    std::vector<u8> mark(S.L, 0);
    for (size_t i = 0; i < regions.size(); ++i) {
      u32 s = regions[i].start > 100 ? regions[i].start - 100 : 0;
      u32 e = regions[i].end < S.L - 100 ? regions[i].end + 100 : S.L - 1;
      for (u32 j = s; j <= e; ++j) mark[j] = 1;
    }
    for (u32 p = 0; p < S.L; ++p) if (mark[p]) sparsePos.push_back(p);
  }
  size_t poolCap = (S.pool.size() + page) / page * page;
  size_t len[5], start[5];
  len[0] = 4 + 8 + poolCap;
  len[1] = 1 + 4 + (P.exome ? sparsePos.size() * 12 : (size_t) S.L * 8);
  len[2] = 4; for (size_t r = 0; r < S.ovf.size(); ++r) len[2] += 8 + S.ovf[r].size() * 32;
  len[3] = P.exome ? 0 : ((size_t) (S.L + 31) >> 5) * 4;
  len[4] = 0;
  size_t cur = page;
  for (int i = 0; i < 5; ++i) { start[i] = cur; cur += (len[i] + page - 1) / page * page; }
  std::vector<u8> hdr(page, 0);
  int pg = (int) page; memcpy(&hdr[0], &pg, 4);
  for (int i = 0; i < 5; ++i) { memcpy(&hdr[4 + i * 16], &start[i], 8); memcpy(&hdr[12 + i * 16], &len[i], 8); }
  fwrite(hdr.data(), 1, page, f);
  // subfile 0: pool
  fseek(f, (long) start[0], SEEK_SET);
  u32 curPtr = (u32) S.pool.size();
  fwrite(&curPtr, 4, 1, f); fwrite(&poolCap, 8, 1, f);
  std::vector<u8> pool(poolCap, 0); if (!S.pool.empty()) memcpy(pool.data(), S.pool.data(), S.pool.size());
  fwrite(pool.data(), 1, poolCap, f);
  // subfile 1: counters
  fseek(f, (long) start[1], SEEK_SET);
  char mode = P.exome ? 2 : 1;
  fwrite(&mode, 1, 1, f); fwrite(&S.L, 4, 1, f);
  if (!P.exome) fwrite(S.counters.data(), 8, S.L, f);
  else for (size_t i = 0; i < sparsePos.size(); ++i) { fwrite(&sparsePos[i], 4, 1, f); fwrite(&S.counters[sparsePos[i]], 8, 1, f); }
  // subfile 2: overflow arrays
  fseek(f, (long) start[2], SEEK_SET);
  u32 nr = (u32) S.ovf.size(); fwrite(&nr, 4, 1, f);
  for (u32 r = 0; r < nr; ++r) {
    u32 sz = (u32) S.ovf[r].size(), lim = sz + 16;
    fwrite(&sz, 4, 1, f); fwrite(&lim, 4, 1, f);
    if (sz) fwrite(S.ovf[r].data(), 32, sz, f);
  }
  // subfile 3: invalid bitmap
  if (!P.exome) { fseek(f, (long) start[3], SEEK_SET); fwrite(S.invalid.data(), 4, (S.L + 31) >> 5, f); }
  // pad the file to its nominal end so every subfile is fully addressable
  fseek(f, 0, SEEK_END);
  long endNow = ftell(f);
  if ((size_t) endNow < cur) { std::vector<u8> z(cur - endNow, 0); fwrite(z.data(), 1, z.size(), f); }
  fclose(f);
}

// ---------------------------------------------------------------- main
static void usage() {
  fprintf(stderr,
    "gt16_synth --index <prefix> --snapshot <prefix> [--fasta f.fa] [--region-out region.bin]\n"
    "           --contigs name:len[,name:len...] | --genome-len L [--n-contigs K]\n"
    "           [--depth 30] [--seed 1] [--snp-rate 1e-3] [--indel-rate 1.25e-4] [--err-rate 3e-3]\n"
    "           [--ovf-regions 16] [--gaps n:len] [--hotspot-rate r --hotspot-depth d --hotspot-len n]\n"
    "           [--invalid-rate r] [--exome nTargets:targetLen] [--multi-pattern-frac f] [--long-indel-frac f]\n"
    "           [--threads 8]\n");
  exit(2);
}

int main(int argc, char **argv) {
  Params P;
  u64 genomeLen = 0; u32 nContigs = 1;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    auto val = [&]() -> const char * { if (i + 1 >= argc) usage(); return argv[++i]; };
    if (a == "--index") P.outIndex = val();
    else if (a == "--snapshot") P.outSnapshot = val();
    else if (a == "--fasta") P.outFasta = val();
    else if (a == "--region-out") P.outRegion = val();
    else if (a == "--contigs") {
      std::string s = val(); size_t pos = 0;
      while (pos < s.size()) {
        size_t c = s.find(',', pos); if (c == std::string::npos) c = s.size();
        std::string item = s.substr(pos, c - pos); size_t k = item.find(':');
        if (k == std::string::npos) usage();
        Contig ct; ct.name = item.substr(0, k); ct.len = (u32) strtoul(item.c_str() + k + 1, NULL, 10);
        P.contigs.push_back(ct); pos = c + 1;
      }
    }
    else if (a == "--genome-len") genomeLen = strtoull(val(), NULL, 10);
    else if (a == "--n-contigs") nContigs = (u32) atoi(val());
    else if (a == "--depth") P.depth = atof(val());
    else if (a == "--seed") P.seed = strtoull(val(), NULL, 10);
    else if (a == "--snp-rate") P.snpRate = atof(val());
    else if (a == "--indel-rate") P.indelRate = atof(val());
    else if (a == "--err-rate") P.errRate = atof(val());
    else if (a == "--ovf-regions") P.ovfRegions = (u32) atoi(val());
    else if (a == "--gaps") { std::string s = val(); size_t k = s.find(':'); P.nGaps = (u32) atoi(s.c_str()); if (k != std::string::npos) P.gapLen = (u32) atoi(s.c_str() + k + 1); }
    else if (a == "--hotspot-rate") P.hotspotRate = atof(val());
    else if (a == "--hotspot-depth") P.hotspotDepth = atof(val());
    else if (a == "--hotspot-len") P.hotspotLen = (u32) atoi(val());
    else if (a == "--invalid-rate") P.invalidRate = atof(val());
    else if (a == "--multi-pattern-frac") P.multiPatFrac = atof(val());
    else if (a == "--long-indel-frac") P.longIndelFrac = atof(val());
    else if (a == "--exome") { std::string s = val(); size_t k = s.find(':'); P.exome = true; P.nTargets = (u32) atoi(s.c_str()); if (k != std::string::npos) P.targetLen = (u32) atoi(s.c_str() + k + 1); }
    else if (a == "--threads") P.threads = atoi(val());
    else usage();
  }
  if (P.outIndex.empty() || P.outSnapshot.empty()) usage();
  if (P.contigs.empty()) {
    if (!genomeLen) usage();
    // GRCh38-like decreasing sizes: weights 1/(i+4)
    double tot = 0; for (u32 i = 0; i < nContigs; ++i) tot += 1.0 / (i + 4);
    u64 used = 0;
    for (u32 i = 0; i < nContigs; ++i) {
      Contig c; char nm[32]; snprintf(nm, sizeof(nm), "chr%u", i + 1); c.name = nm;
      c.len = (i + 1 == nContigs) ? (u32) (genomeLen - used) : (u32) (genomeLen * (1.0 / (i + 4)) / tot);
      used += c.len; P.contigs.push_back(c);
    }
  }
  if (P.ovfRegions < 1) P.ovfRegions = 1;
  if (P.threads < 1) P.threads = 1;

  Genome G;
  u64 tot = 0;
  for (size_t i = 0; i < P.contigs.size(); ++i) { G.contigStart.push_back((u32) tot); tot += P.contigs[i].len; }
  if (tot >= 0xFFFFFF00ull) { fprintf(stderr, "genome too long for 32-bit coordinates\n"); return 1; }
  G.L = (u32) tot;
  G.base.resize(G.L);
  {  // reference bases: uniform with occasional homopolymer / low-complexity stretches
    std::vector<std::thread> th;
    u32 per = (G.L + P.threads - 1) / P.threads;
    for (int t = 0; t < P.threads; ++t) th.emplace_back([&, t]() {
      u32 a = std::min<u64>((u64) t * per, G.L), b = std::min<u64>((u64) (t + 1) * per, G.L);
      for (u32 blk = a / 64; blk <= (b ? (b - 1) / 64 : 0); ++blk) {
        Rng r(P.seed, 0x524546, blk);
        u64 bits = r.next(), bits2 = r.next();
        bool run = (r.next() & 63) == 0;  // ~1.5 % of 64-bp blocks carry a homopolymer run
        u32 runAt = (u32) (r.next() % 48), runLen = 4 + (u32) (r.next() % 14); u8 runBase = (u8) (r.next() & 3);
        for (u32 k = 0; k < 64; ++k) {
          u64 p = (u64) blk * 64 + k; if (p < a || p >= b) continue;
          u8 bs = (u8) (((k < 32 ? bits >> (2 * k) : bits2 >> (2 * (k - 32)))) & 3);
          if (run && k >= runAt && k < runAt + runLen) bs = runBase;
          G.base[p] = bs;
        }
      }
    });
    for (auto &x : th) x.join();
  }
  for (u32 gidx = 0; gidx < P.nGaps; ++gidx) {
    Rng r(P.seed, 0x474150, gidx);
    u32 s = (u32) (r.uni() * (G.L > P.gapLen + 1000 ? G.L - P.gapLen - 1000 : 1)) + 500;
    G.gaps.push_back(std::make_pair(s, std::min<u64>((u64) s + P.gapLen, G.L)));
  }
  std::vector<Region> targets;
  if (P.exome) {
    // evenly spread, jittered, non-overlapping targets; merged/sorted like RegionIndexBuilder output
    u64 stride = G.L / (P.nTargets + 1);
    if (stride < P.targetLen + 400) { fprintf(stderr, "too many targets for this genome\n"); return 1; }
    for (u32 i = 0; i < P.nTargets; ++i) {
      Rng r(P.seed, 0x544152, i);
      u32 s = (u32) ((u64) (i + 1) * stride - stride / 2 + r.below((u32) (stride - P.targetLen - 400)));
      u32 len = P.targetLen / 2 + r.below(P.targetLen);
      Region t; t.start = s; t.end = std::min<u64>((u64) s + len, G.L - 1);
      if (!targets.empty() && t.start <= targets.back().end + 1) continue;
      targets.push_back(t);
    }
    if (!P.outRegion.empty()) {
      FILE *f = fopen(P.outRegion.c_str(), "wb");
      u32 n = (u32) targets.size(); fwrite(&n, 4, 1, f); fwrite(targets.data(), 8, n, f); fclose(f);
    }
  }

  Snapshot S;
  S.L = G.L;
  S.counters.assign(G.L, 0);
  S.invalid.assign(((size_t) G.L + 31) >> 5, 0);
  S.ovf.resize(P.ovfRegions);
  // chunk boundaries are multiples of 32 so that invalid-bitmap words are owned by one worker
  u32 chunk = 1u << 20;
  u32 nChunks = (u32) (((u64) G.L + chunk - 1) / chunk);
  std::vector<ChunkOut> outs(nChunks);
  {
    std::vector<std::thread> th;
    for (int t = 0; t < P.threads; ++t) th.emplace_back([&, t]() {
      for (u32 c = (u32) t; c < nChunks; c += (u32) P.threads) {
        u32 a = c * chunk, b = (u32) std::min<u64>((u64) a + chunk, G.L);
        if (P.exome) {  // skip chunks with no coverage at all
          bool any = false;
          for (size_t i = 0; i < targets.size() && !any; ++i)
            if (targets[i].end + 400 > a && targets[i].start < (u64) b + 400) any = true;
          if (!any) continue;
        }
        gen_chunk(P, G, targets, a, b, S, outs[c]);
      }
    });
    for (auto &x : th) x.join();
  }
  // ordered merge: assign overflow array indices per region, relocate pool pointers
  u64 nOvf = 0;
  for (u32 c = 0; c < nChunks; ++c) {
    ChunkOut &o = outs[c];
    u32 poolBase = (u32) S.pool.size();
    S.pool.insert(S.pool.end(), o.poolBytes.begin(), o.poolBytes.end());
    for (size_t k = 0; k < o.poolFix.size(); ++k) {
      OvfRec &rec = o.ovf[o.poolFix[k].first].second;
      u8 *fld = o.poolFix[k].second ? rec.del : rec.ins;
      u32 off; memcpy(&off, fld, 4); off += poolBase; memcpy(fld, &off, 4);
    }
    for (size_t k = 0; k < o.ovf.size(); ++k) {
      u32 p = o.ovf[k].first, reg = region_of(p, G.L, P.ovfRegions);
      u32 idx = (u32) S.ovf[reg].size();
      S.ovf[reg].push_back(o.ovf[k].second);
      S.counters[p] = (S.counters[p] & 0xFFFFFFFFull) | ((u64) idx << 32);
      ++nOvf;
    }
    std::vector<std::pair<u32, OvfRec> >().swap(o.ovf);
    std::vector<u8>().swap(o.poolBytes);
  }
  write_index(P, G);
  write_snapshot(P, S, targets);
  fprintf(stderr, "gt16_synth: L=%u contigs=%zu overflow=%llu (%.3f%%) pool=%zu B%s\n", G.L, P.contigs.size(),
          (unsigned long long) nOvf, 100.0 * nOvf / G.L, S.pool.size(), P.exome ? " [sparse]" : "");
  return 0;
}
