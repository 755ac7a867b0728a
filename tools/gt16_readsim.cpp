// gt16_readsim -- deterministic diploid read simulator: FASTA + SAM for the reference's OWN producer.
//
// TEST / BENCH INPUT GENERATOR.  tools/gt16_synth writes snapshots directly (this repository's reading of the
// encoding); this tool writes what comes BEFORE the snapshot -- a reference FASTA and aligned reads -- so that the
// snapshot itself is written by the reference's unmodified bam2snapshot (oracle/_ref/ref_bam2snapshot:
// bam2snapshot.cpp:591-716, SnapshotHandler.cpp:178-236, promotion rules SNP.cpp:745-891) and the index files by
// the reference's packer (oracle/_ref/ref_mkindex).  BASELINE.json configs[0]: "5 Mb synthetic reference, 30x
// diploid snapshot with planted SNPs/indels".
//
// A pure function of its arguments (own PRNG, no libc rand): the same command line gives the same bytes anywhere.
//
//   gt16_readsim --out PREFIX --contigs chrA:3000000,chrB:2000000 [--depth 30] [--read-len 100:150]
//                [--snp-rate 1e-3] [--indel-rate 1.25e-4] [--long-indel-frac 0.05] [--softclip-frac 0.02]
//                [--dup-frac 0.005] [--compound-frac 0.1] [--hotspots N:LEN:DEPTH[,N:LEN:DEPTH]] [--seed 1]
// writes PREFIX.fa, PREFIX.sam, PREFIX.truth.tsv
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>

namespace {

struct Rng {                                       // xoshiro256**, seeded by splitmix64
  uint64_t s[4];
  explicit Rng(uint64_t seed) {
    for (int i = 0; i < 4; ++i) { seed += 0x9E3779B97F4A7C15ull; uint64_t z = seed; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; s[i] = z ^ (z >> 31); }
  }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    const uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  double uni() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
  uint32_t below(uint32_t n) { return (uint32_t) (((next() >> 32) * (uint64_t) n) >> 32); }
};

struct Allele {
  char type = 0;       // 0 = reference, 'S', 'I', 'D'
  uint8_t alt = 0;     // SNP: alternative base 0..3
  std::string ins;     // inserted bases (ACGT)
  uint32_t len = 0;    // deletion length
};
struct Variant {
  uint32_t pos;        // 0-based position in the contig: SNP at pos; insertion AFTER pos; deletion of [pos+1, pos+1+len)
  Allele a[2];         // what each haplotype carries there (het, hom, or compound het: two different indels)
};

struct Contig { std::string name; uint32_t len; std::string seq; std::vector<Variant> vars; };
struct Hotspot { uint32_t contig, start, len; double depth; };

const char kBases[] = "ACGT";

struct Out {
  FILE *f; std::vector<char> buf;
  explicit Out(const std::string &path) : f(fopen(path.c_str(), "wb")) { if (!f) { perror(path.c_str()); exit(1); } buf.reserve(1 << 22); }
  void put(const char *p, size_t n) { if (buf.size() + n > (1u << 22)) flush(); buf.insert(buf.end(), p, p + n); }
  void puts_(const std::string &s) { put(s.data(), s.size()); }
  void flush() { if (!buf.empty()) { fwrite(buf.data(), 1, buf.size(), f); buf.clear(); } }
  ~Out() { flush(); fclose(f); }
};

// base-quality mix: ~90 % above Q20, the rest spread over the four weight classes (Weight0..3 = 10/13/17/20)
int draw_quality(Rng &r) {
  const double u = r.uni();
  if (u < 0.90) return 21 + (int) r.below(20);
  if (u < 0.94) return 18 + (int) r.below(3);
  if (u < 0.97) return 14 + (int) r.below(4);
  if (u < 0.99) return 11 + (int) r.below(3);
  return 2 + (int) r.below(9);
}

}  // namespace

int main(int argc, char **argv) {
  std::string out, contigSpec;
  double depth = 30, snpRate = 1e-3, indelRate = 1.25e-4, longFrac = 0.05, clipFrac = 0.02, dupFrac = 0.005, compoundFrac = 0.1;
  int lenLo = 100, lenHi = 150;
  uint64_t seed = 1;
  struct HotSpec { int n; uint32_t len; double depth; };
  std::vector<HotSpec> hotSpecs;
  for (int i = 1; i + 1 < argc; i += 2) {
    const std::string k = argv[i]; const char *v = argv[i + 1];
    if (k == "--out") out = v;
    else if (k == "--contigs") contigSpec = v;
    else if (k == "--depth") depth = atof(v);
    else if (k == "--read-len") { if (sscanf(v, "%d:%d", &lenLo, &lenHi) != 2) lenHi = lenLo = atoi(v); }
    else if (k == "--snp-rate") snpRate = atof(v);
    else if (k == "--indel-rate") indelRate = atof(v);
    else if (k == "--long-indel-frac") longFrac = atof(v);
    else if (k == "--softclip-frac") clipFrac = atof(v);
    else if (k == "--dup-frac") dupFrac = atof(v);
    else if (k == "--seed") seed = strtoull(v, NULL, 10);
    else if (k == "--compound-frac") compoundFrac = atof(v);
    else if (k == "--hotspots") {
      for (const char *q = v; q && *q;) {
        unsigned a = 0, b = 0; double c = 0;
        if (sscanf(q, "%u:%u:%lf", &a, &b, &c) == 3) hotSpecs.push_back(HotSpec{(int) a, b, c});
        q = strchr(q, ','); if (q) ++q;
      }
    }
    else { fprintf(stderr, "gt16_readsim: unknown option %s\n", k.c_str()); return 2; }
  }
  if (out.empty() || contigSpec.empty() || lenLo < 30 || lenHi < lenLo || lenHi > 250) {
    fprintf(stderr, "usage: gt16_readsim --out PREFIX --contigs name:len[,name:len...] [options]\n");
    return 2;
  }
  std::vector<Contig> contigs;
  for (size_t p = 0; p < contigSpec.size();) {
    size_t q = contigSpec.find(',', p); if (q == std::string::npos) q = contigSpec.size();
    const std::string item = contigSpec.substr(p, q - p);
    const size_t c = item.find(':');
    Contig ct; ct.name = item.substr(0, c); ct.len = (uint32_t) strtoul(item.c_str() + c + 1, NULL, 10);
    contigs.push_back(ct);
    p = q + 1;
  }
  Rng rg(seed);
  // ---- genome + FASTA
  {
    Out fa(out + ".fa");
    for (Contig &ct : contigs) {
      ct.seq.resize(ct.len);
      for (uint32_t i = 0; i < ct.len; ++i) ct.seq[i] = kBases[rg.next() >> 62];
      fa.puts_(">" + ct.name + "\n");
      for (uint32_t i = 0; i < ct.len; i += 60) { fa.put(ct.seq.data() + i, std::min<uint32_t>(60, ct.len - i)); fa.put("\n", 1); }
    }
  }
  // ---- planted variants (kept >= 45 bp apart so that a read never meets two events within one CIGAR operation)
  Rng rv(seed * 7919 + 1);
  {
    Out truth(out + ".truth.tsv");
    const double rate = snpRate + indelRate;
    for (Contig &ct : contigs) {
      uint32_t pos = 60;
      while (rate > 0) {
        const double gap = -log(1.0 - rv.uni()) / rate;
        if (gap > 4e9) break;
        pos += 45 + (uint32_t) gap;
        if (pos + 100 >= ct.len) break;
        Variant v; v.pos = pos;
        const int ref = (int) (strchr(kBases, ct.seq[pos]) - kBases);
        auto draw_indel = [&](Allele &al) {
          uint32_t len = 1 + rv.below(12);
          if (rv.uni() < longFrac) len = 13 + rv.below(18);         // 13..30: pooled patterns (inline patterns hold <= 12 bases)
          if (rv.uni() < 0.5) { al.type = 'I'; al.ins.clear(); for (uint32_t k = 0; k < len; ++k) al.ins.push_back(kBases[rv.next() >> 62]); }
          else { al.type = 'D'; al.len = len; }
        };
        Allele al;
        const bool snp = rv.uni() < snpRate / rate;
        if (snp) { al.type = 'S'; al.alt = (uint8_t) ((ref + 1 + rv.below(3)) & 3); }
        else draw_indel(al);
        const double z = rv.uni();
        if (z < 1.0 / 3) { v.a[0] = al; v.a[1] = al; }                 // het : hom = 2 : 1
        else v.a[z < 2.0 / 3 ? 0 : 1] = al;
        if (rv.uni() < compoundFrac) {                                // compound het: the other haplotype carries a different allele
          Allele &other = v.a[v.a[0].type ? 1 : 0];
          if (z < 1.0 / 3) other = Allele();
          if (snp) { other.type = 'S'; other.alt = (uint8_t) ((al.alt + 1 + rv.below(2)) & 3); if ((int) other.alt == ref) other.alt = (uint8_t) ((other.alt + 1) & 3); if (other.alt == al.alt) other = al; }
          else draw_indel(other);
        }
        ct.vars.push_back(v);
        char line[512];
        for (int h = 0; h < 2; ++h) {
          const Allele &x = v.a[h];
          if (!x.type) continue;
          snprintf(line, sizeof(line), "%s\t%u\thap%d\t%c\t%s\t%u\n", ct.name.c_str(), pos + 1, h, x.type,
                   x.type == 'S' ? std::string(1, kBases[x.alt]).c_str() : (x.type == 'I' ? x.ins.c_str() : "-"), x.len);
          truth.puts_(line);
        }
      }
    }
  }
  // ---- depth hotspots (per-strand counts >= 255: overflow promotion, SNP.cpp:862; very deep ones: invalid bit :873-884)
  std::vector<Hotspot> hots;
  for (const HotSpec &sp : hotSpecs)
    for (int h = 0; h < sp.n; ++h) {
      Hotspot hs; hs.contig = rv.below((uint32_t) contigs.size());
      const uint32_t L = contigs[hs.contig].len;
      if (L < sp.len + 1000) continue;
      hs.start = 400 + rv.below(L - sp.len - 800); hs.len = sp.len; hs.depth = sp.depth;
      hots.push_back(hs);
    }
  // ---- reads
  Out sam(out + ".sam");
  sam.puts_("@HD\tVN:1.0\tSO:coordinate\n");
  for (const Contig &ct : contigs) { char l[256]; snprintf(l, sizeof(l), "@SQ\tSN:%s\tLN:%u\n", ct.name.c_str(), ct.len); sam.puts_(l); }
  Rng rr(seed * 104729 + 3);
  const double meanLen = 0.5 * (lenLo + lenHi);
  uint64_t readId = 0, nReads = 0;
  std::string seq, qual, cigar, line;
  for (size_t ci = 0; ci < contigs.size(); ++ci) {
    const Contig &ct = contigs[ci];
    size_t vLo = 0;                                   // first variant with pos + reach >= start
    for (uint32_t start = 0; start + lenHi + 40 < ct.len; ++start) {
      double lam = depth / meanLen;
      for (const Hotspot &h : hots) if (h.contig == ci && start + meanLen / 2 >= h.start && start + meanLen / 2 < h.start + h.len) lam = h.depth / meanLen;
      uint32_t k = (uint32_t) lam;
      if (rr.uni() < lam - k) ++k;
      while (vLo < ct.vars.size() && ct.vars[vLo].pos + 40 < start) ++vLo;
      for (; k > 0; --k) {
        const int hap = (int) (rr.next() >> 63), strand = (int) (rr.next() >> 63);
        const int want = lenLo + (int) rr.below((uint32_t) (lenHi - lenLo + 1));
        seq.clear(); qual.clear(); cigar.clear();
        // walk the haplotype from `start`
        uint32_t g = start; int run = 0; bool bad = false;
        size_t vi = vLo;
        while (vi < ct.vars.size() && ct.vars[vi].pos < start) {         // a read must not start inside a deletion it carries
          const Allele &al = ct.vars[vi].a[hap];
          if (al.type == 'D' && start <= ct.vars[vi].pos + al.len) bad = true;
          ++vi;
        }
        if (bad) continue;
        auto flushM = [&]() { if (run) { cigar += std::to_string(run) + "M"; run = 0; } };
        while ((int) seq.size() < want && g < ct.len) {
          char b = ct.seq[g];
          const Allele *v = (vi < ct.vars.size() && ct.vars[vi].pos == g) ? &ct.vars[vi].a[hap] : nullptr;
          if (v && v->type == 'S') b = kBases[v->alt];
          seq.push_back(b); ++run; ++g;
          if (v) {
            ++vi;
            if (v->type == 'I' || v->type == 'D') {
              const int room = want - (int) seq.size();
              if (v->type == 'I') {
                if (room < (int) v->ins.size() + 8) break;               // keep >= 8 aligned bases after the event
                flushM(); seq += v->ins; cigar += std::to_string(v->ins.size()) + "I";
              } else {
                if (room < 8 || g + v->len + 8 >= ct.len) break;
                flushM(); cigar += std::to_string(v->len) + "D"; g += v->len;
              }
            }
          }
        }
        flushM();
        if ((int) seq.size() < 40 || cigar.empty() || cigar.back() != 'M') continue;
        uint32_t pos1 = start + 1;
        // sequencing errors + qualities
        const size_t n = seq.size();
        qual.resize(n);
        for (size_t t = 0; t < n; ++t) {
          const int q = draw_quality(rr);
          qual[t] = (char) (33 + q);
          if (rr.uni() < pow(10.0, -q / 10.0)) { const int cur = (int) (strchr(kBases, seq[t]) - kBases); seq[t] = kBases[(cur + 1 + rr.below(3)) & 3]; }
        }
        // soft clips: the first / last aligned bases become clipped (POS moves to the first aligned base)
        if (rr.uni() < clipFrac) {
          const int c = 3 + (int) rr.below(10);
          const size_t mEnd = cigar.find('M');
          const int m0 = atoi(cigar.c_str());
          if (rr.next() >> 63) {
            if (mEnd != std::string::npos && m0 > c + 12) {
              cigar = std::to_string(c) + "S" + std::to_string(m0 - c) + cigar.substr(mEnd);
              for (int t = 0; t < c; ++t) seq[t] = kBases[rr.next() >> 62];
              pos1 += c;
            }
          } else {
            size_t q = cigar.size() - 1, p = q;                           // last operation is M
            while (p > 0 && cigar[p - 1] >= '0' && cigar[p - 1] <= '9') --p;
            const int ml = atoi(cigar.c_str() + p);
            if (ml > c + 12) {
              cigar = cigar.substr(0, p) + std::to_string(ml - c) + "M" + std::to_string(c) + "S";
              for (int t = 0; t < c; ++t) seq[n - 1 - t] = kBases[rr.next() >> 62];
            }
          }
        }
        int flag = strand ? 16 : 0;
        if (rr.uni() < dupFrac) flag |= 0x400;                             // duplicates are skipped by the producer (bam2snapshot.cpp:172)
        line.clear();
        line += "r"; line += std::to_string(readId++); line += "\t"; line += std::to_string(flag); line += "\t"; line += ct.name;
        line += "\t"; line += std::to_string(pos1); line += "\t60\t"; line += cigar; line += "\t*\t0\t0\t"; line += seq; line += "\t"; line += qual; line += "\n";
        sam.puts_(line);
        ++nReads;
      }
    }
  }
  fprintf(stderr, "gt16_readsim: %zu contigs, %llu reads\n", contigs.size(), (unsigned long long) nReads);
  return 0;
}
