// gt16_sam.cpp -- host side of the producer: SAM text -> aligned reads as bam2snapshot hands them to
// updateSnpCounterForReads (bam2snapshot.cpp:79-206), and the snapshot container writer (writeSnpInfoSnapshot,
// SnapshotHandler.cpp:83-236).  Plain C++; the counter construction itself is gt16_pileup.cu.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <unistd.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <algorithm>
#include <atomic>
#include <future>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <zlib.h>
#include <vector>
#include "../../include/gt16.h"

namespace {
void set_err(char *err, size_t n, const char *fmt, ...) {
  if (!err || !n) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err, n, fmt, ap);
  va_end(ap);
}
}  // namespace

// vectors whose resize() leaves new elements unset (every one is written right after): growing a 150 MB array of bases
// does not zero it first
template <class T>
struct DefaultInit : std::allocator<T> {
  template <class U> struct rebind { using other = DefaultInit<U>; };
  template <class U> void construct(U *p) noexcept { ::new ((void *) p) U; }
  template <class U, class... A> void construct(U *p, A &&... a) { ::new ((void *) p) U(std::forward<A>(a)...); }
};
template <class T> using Vec = std::vector<T, DefaultInit<T> >;

struct gt16_sam {
  Vec<uint32_t> start, length, nCigar, cigar;
  Vec<uint8_t> reverse, bases, quals;
  Vec<uint64_t> seqOff, cigarOff;
  uint64_t total = 0;
};

extern "C" {

void gt16_pileup_default_params(gt16_pileup_params *p) {
  if (!p) return;
  memset(p, 0, sizeof(*p));
  // bam2snapshot.cpp:352-363 with Weight0..3 = 10 13 17 20
  const int bound[4] = {10, 13, 17, 20};
  for (int q = 0; q < 41; ++q) p->weight_map[q] = q <= bound[0] ? 0 : q <= bound[1] ? 1 : q <= bound[2] ? 2 : q <= bound[3] ? 3 : 4;
  p->trim_size = 5;
  p->indel_weight_threshold = 1;      // IndelQualityThreshold 13 >= Weight1 (:364-368)
  p->n_overflow_regions = 16;
}

// SEQ character -> the reference's 2-bit code through BAM's 4-bit code (bam_nt16_table, then bamcode2snpcode,
// bam2snapshot.h:52-58): A C G T keep their codes, N becomes G (code 15 -> 2), every other IUPAC letter A
static inline uint8_t snp_code(char c) {
  switch (c) {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    case 'N': case 'n': return 2;
    default: break;
  }
  // ambiguity codes: bam_nt16_table gives M=3 R=5 S=6 V=7 W=9 Y=10 H=11 K=12 D=13 B=14 (all -> 0), anything else 15 (-> 2)
  switch (c) {
    case 'M': case 'R': case 'S': case 'V': case 'W': case 'Y': case 'H': case 'K': case 'D': case 'B':
    case 'm': case 'r': case 's': case 'v': case 'w': case 'y': case 'h': case 'k': case 'd': case 'b':
    case '=': return 0;
    default: return 2;
  }
}

// one alignment line (NUL-terminated, tabs still in place) into S; false = malformed line
static bool parse_alignment(char *line, const gt16_index *ix, uint32_t L, uint32_t mapqThreshold,
                            std::unordered_map<std::string, int> &ids, std::vector<uint32_t> &ops, gt16_sam *S) {
  // QNAME FLAG RNAME POS MAPQ CIGAR RNEXT PNEXT TLEN SEQ QUAL
  char *fld[11];
  int nf = 0;
  char *p = line;
  while (nf < 11) {
    fld[nf++] = p;
    char *t = strchr(p, '\t');
    if (!t) break;
    *t = 0; p = t + 1;
  }
  if (nf < 11) return false;
  S->total++;
  const unsigned flag = (unsigned) strtoul(fld[1], nullptr, 10);
  const long pos1 = strtol(fld[3], nullptr, 10);
  const unsigned mapq = (unsigned) strtoul(fld[4], nullptr, 10);
  if (mapq < mapqThreshold) return true;                                 // bam2snapshot.cpp:84-87
  if (fld[2][0] == '*' || pos1 <= 0) return true;                        // :91-94 (tid < 0 or pos < 0)
  if (flag & 0x400) return true;                                         // :171-174
  const char *seq = fld[9], *qual = fld[10];
  const size_t lseq = (seq[0] == '*' && seq[1] == 0) ? 0 : strlen(seq);
  if (!lseq) return true;                                                // readLength <= 0 (:181)
  // CIGAR: the reference string is rebuilt from the BAM operations (:132-141); l_qseq must equal M + I + S (:181)
  ops.clear();
  size_t fromCigar = 0;
  bool bad = fld[5][0] == '*' || fld[5][0] == 0;
  for (const char *c = fld[5]; *c && !bad;) {
    char *e;
    const unsigned long len = strtoul(c, &e, 10);
    if (e == c || !*e) { bad = true; break; }
    int op;
    switch (*e) {
      case 'M': op = 0; break; case 'I': op = 1; break; case 'D': op = 2; break; case 'N': op = 3; break;
      case 'S': op = 4; break; case 'H': op = 5; break; case 'P': op = 6; break; case '=': op = 7; break;
      case 'X': op = 8; break; default: op = -1;
    }
    if (op < 0 || len >= (1ul << 28)) { bad = true; break; }
    if (op == 0 || op == 1 || op == 4) fromCigar += len;                 // getLenFromCigar (:55-77) counts M, I and S
    ops.push_back((uint32_t) (len << 4) | (uint32_t) op);
    c = e + 1;
  }
  if (bad || fromCigar != lseq) return true;                             // :181-184, :186-188 (empty CIGAR)
  int id;
  {
    auto it = ids.find(fld[2]);
    if (it == ids.end()) { id = gt16_index_sequence_id(ix, fld[2]); ids.emplace(fld[2], id); }
    else id = it->second;
  }
  uint32_t amb;
  if (id < 1 || gt16_index_locate_id(ix, id, (uint32_t) pos1, &amb) != 0 || amb >= L) return true;   // :186
  const bool haveQual = !(qual[0] == '*' && qual[1] == 0) && strlen(qual) == lseq;
  S->start.push_back(amb);
  S->reverse.push_back((flag & 16) ? 1 : 0);
  S->length.push_back((uint32_t) lseq);
  S->seqOff.push_back(S->bases.size());
  S->cigarOff.push_back(S->cigar.size());
  S->nCigar.push_back((uint32_t) ops.size());
  S->cigar.insert(S->cigar.end(), ops.begin(), ops.end());
  const size_t at = S->bases.size();
  S->bases.resize(at + lseq); S->quals.resize(at + lseq);
  for (size_t k = 0; k < lseq; ++k) {
    S->bases[at + k] = snp_code(seq[k]);
    S->quals[at + k] = haveQual ? (uint8_t) (qual[k] - 33) : (uint8_t) 0xFF;
  }
  return true;
}

// Lists parsed by different threads, joined in order: every list is copied to its place in the whole by some thread.
static void join_parts(std::vector<gt16_sam> &part, gt16_sam *S) {
  const size_t np = part.size();
  if (np == 1) { *S = std::move(part[0]); return; }
  std::vector<size_t> r0(np + 1, 0), b0(np + 1, 0), c0(np + 1, 0);
  for (size_t t = 0; t < np; ++t) {
    r0[t + 1] = r0[t] + part[t].start.size(); b0[t + 1] = b0[t] + part[t].bases.size(); c0[t + 1] = c0[t] + part[t].cigar.size();
    S->total += part[t].total;
  }
  const size_t nr = r0[np], nb = b0[np], nc = c0[np];
  S->start.resize(nr); S->reverse.resize(nr); S->length.resize(nr); S->seqOff.resize(nr); S->cigarOff.resize(nr); S->nCigar.resize(nr);
  S->bases.resize(nb); S->quals.resize(nb); S->cigar.resize(nc);
  unsigned nt = std::thread::hardware_concurrency();
  nt = (unsigned) std::max<size_t>(1, std::min<size_t>(std::min(16u, nt ? nt : 4u), np));
  std::atomic<size_t> next(0);
  auto work = [&]() {
    for (;;) {
      const size_t t = next.fetch_add(1);
      if (t >= np) break;
      gt16_sam &P = part[t];
      const size_t n = P.start.size();
      if (n) {
        memcpy(&S->start[r0[t]], P.start.data(), n * 4); memcpy(&S->reverse[r0[t]], P.reverse.data(), n);
        memcpy(&S->length[r0[t]], P.length.data(), n * 4); memcpy(&S->nCigar[r0[t]], P.nCigar.data(), n * 4);
        for (size_t i = 0; i < n; ++i) { S->seqOff[r0[t] + i] = b0[t] + P.seqOff[i]; S->cigarOff[r0[t] + i] = c0[t] + P.cigarOff[i]; }
      }
      if (!P.bases.empty()) { memcpy(&S->bases[b0[t]], P.bases.data(), P.bases.size()); memcpy(&S->quals[b0[t]], P.quals.data(), P.quals.size()); }
      if (!P.cigar.empty()) memcpy(&S->cigar[c0[t]], P.cigar.data(), P.cigar.size() * 4);
      P = gt16_sam();
    }
  };
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back(work);
  for (auto &x : th) x.join();
}

// The file is mapped and cut into slices at line ends; worker threads parse their slices into lists of their own, which
// are then joined in file order (the order of the reads decides which of two indel patterns is listed first).
int gt16_sam_read(const char *path, const gt16_index *ix, uint32_t mapqThreshold, gt16_sam **out, char *err, size_t errLen) {
  if (!path || !ix || !out) return GT16_EINVAL;
  int fd = open(path, O_RDONLY);
  if (fd < 0) { set_err(err, errLen, "[ERROR] Failed accessing %s.", path); return GT16_EIO; }
  struct stat st;
  if (fstat(fd, &st) != 0) { close(fd); set_err(err, errLen, "[ERROR] Failed accessing %s.", path); return GT16_EIO; }
  const size_t size = (size_t) st.st_size;
  gt16_sam *S = new gt16_sam();
  if (!size) { close(fd); *out = S; return GT16_OK; }
  const char *data = (const char *) mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
  close(fd);
  if (data == MAP_FAILED) { delete S; set_err(err, errLen, "[ERROR] Failed accessing %s.", path); return GT16_EIO; }
  const uint32_t L = gt16_index_length(ix);
  unsigned nt = std::thread::hardware_concurrency();
  nt = std::max(1u, std::min(16u, nt ? nt : 4u));
  if (size < (4u << 20)) nt = 1;
  std::vector<size_t> cut(nt + 1, size);
  cut[0] = 0;
  for (unsigned t = 1; t < nt; ++t) {
    size_t c = size / nt * t;
    const char *nl = (const char *) memchr(data + c, '\n', size - c);
    cut[t] = nl ? (size_t) (nl - data) + 1 : size;
  }
  std::vector<gt16_sam> part(nt);
  std::vector<char> ok(nt, 1);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t]() {
    std::unordered_map<std::string, int> ids;          // RNAME -> chromosome id (chrIdMap, bam2snapshot.cpp:157-158)
    std::vector<uint32_t> ops;
    std::string line;
    gt16_sam &P = part[t];
    {                                                  // about half of a SAM file's bytes are SEQ and QUAL
      const size_t slice = cut[t + 1] - cut[t];
      P.bases.reserve(slice / 2); P.quals.reserve(slice / 2);
      for (auto *v : {&P.start, &P.length, &P.nCigar}) v->reserve(slice / 160);
    }
    for (size_t a = cut[t]; a < cut[t + 1];) {
      const char *nl = (const char *) memchr(data + a, '\n', cut[t + 1] - a);
      size_t b = nl ? (size_t) (nl - data) : cut[t + 1];
      size_t e = b;
      while (e > a && (data[e - 1] == '\r')) --e;
      if (e > a && data[a] != '@') {
        line.assign(data + a, e - a);
        if (!parse_alignment(&line[0], ix, L, mapqThreshold, ids, ops, &P)) { ok[t] = 0; return; }
      }
      a = b + 1;
    }
  });
  for (auto &x : th) x.join();
  munmap((void *) data, size);
  for (unsigned t = 0; t < nt; ++t) if (!ok[t]) { delete S; set_err(err, errLen, "%s: alignment line with fewer than 11 fields", path); return GT16_EDATA; }
  join_parts(part, S);
  *out = S;
  return GT16_OK;
}

// ---- BAM (the reference reads it through samtools' bgzf / bam_read1: bam2snapshot.cpp:159-163) --------------------
// BGZF is a series of gzip members; zlib inflates them one after the other.  The decompressed stream: "BAM\1", header
// text, reference names, then the alignment records bamToAlignment looks at (SAM specification, section 4.2).
namespace {
// A BGZF file is a series of independent gzip members of at most 64 KB, each with its total size in a "BC" extra field:
// the members of a batch (up to 256 MB of output) are inflated by all the host's threads at once while the batch before
// it is being parsed.  A gzip stream without the fields (a BAM written by something else than bgzf) goes through one
// streaming inflate instead.
struct BgzfBatch {
  Vec<uint8_t> data;
  bool corruptBehind = false;      // the stream is damaged right behind these bytes
};
struct BgzfBlock { size_t off, size, out; uint32_t isize; };

struct BgzfReader {
  bool eof = false, bad = false, open = false;
  // streaming
  FILE *f = nullptr;
  z_stream z;
  std::vector<uint8_t> in;
  // member by member, in parallel
  const uint8_t *map = nullptr;
  size_t mapSize = 0, nextOff = 0;
  bool parallel = false, scanBad = false;
  BgzfBatch cur;
  size_t curPos = 0;
  std::future<BgzfBatch> ahead;
  size_t batchBytes = (size_t) 256 << 20;              // GT16_BGZF_BATCH: test hook (small batches: records straddle them)

  // total size of the member at `o` when it carries a BC field and lies inside the file, else 0
  static size_t member_size(const uint8_t *m, size_t size, size_t o) {
    if (o + 18 > size || m[o] != 0x1f || m[o + 1] != 0x8b || m[o + 2] != 8 || !(m[o + 3] & 4)) return 0;
    const size_t xlen = (size_t) m[o + 10] | ((size_t) m[o + 11] << 8);
    if (o + 12 + xlen > size) return 0;
    for (size_t s = o + 12; s + 4 <= o + 12 + xlen;) {
      const size_t slen = (size_t) m[s + 2] | ((size_t) m[s + 3] << 8);
      if (m[s] == 'B' && m[s + 1] == 'C' && slen == 2 && s + 6 <= o + 12 + xlen) {
        const size_t total = ((size_t) m[s + 4] | ((size_t) m[s + 5] << 8)) + 1;
        return total >= 12 + xlen + 8 && o + total <= size ? total : 0;
      }
      s += 4 + slen;
    }
    return 0;
  }
  // the members of the next batch, from nextOff on
  std::vector<BgzfBlock> scan() {
    std::vector<BgzfBlock> blocks;
    size_t out = 0;
    while (nextOff < mapSize && out < batchBytes) {
      const size_t total = member_size(map, mapSize, nextOff);
      uint32_t isize = 0;
      if (total) memcpy(&isize, map + nextOff + total - 4, 4);
      if (!total || isize > 65536u) { scanBad = true; break; }
      blocks.push_back(BgzfBlock{nextOff, total, out, isize});
      out += isize;
      nextOff += total;
    }
    return blocks;
  }
  static BgzfBatch inflate_batch(const uint8_t *m, std::vector<BgzfBlock> blocks, bool corruptBehind) {
    BgzfBatch B;
    B.corruptBehind = corruptBehind;
    const size_t total = blocks.empty() ? 0 : blocks.back().out + blocks.back().isize;
    B.data.resize(total);
    unsigned nt = std::thread::hardware_concurrency();
    nt = std::max(1u, std::min(32u, nt ? nt : 4u));
    if (blocks.size() < 64) nt = 1;
    std::atomic<size_t> next(0), firstBad(blocks.size());
    auto work = [&]() {
      z_stream zs;
      memset(&zs, 0, sizeof(zs));
      if (inflateInit2(&zs, 15 + 16) != Z_OK) { firstBad = 0; return; }
      for (;;) {
        const size_t k0 = next.fetch_add(16);
        if (k0 >= blocks.size()) break;
        for (size_t k = k0; k < std::min(blocks.size(), k0 + 16); ++k) {
          const BgzfBlock &b = blocks[k];
          zs.next_in = (Bytef *) (m + b.off); zs.avail_in = (uInt) b.size;
          Bytef none;                                              // an empty member (the end-of-file marker) still needs a place to write to
          zs.next_out = b.isize ? B.data.data() + b.out : &none; zs.avail_out = b.isize;
          const int rc = inflate(&zs, Z_FINISH);                   // the gzip wrapper checks CRC32 and ISIZE
          if (rc != Z_STREAM_END || zs.avail_out != 0) { size_t seen = firstBad.load(); while (k < seen && !firstBad.compare_exchange_weak(seen, k)) {} }
          inflateReset(&zs);
        }
      }
      inflateEnd(&zs);
    };
    if (nt == 1) work();
    else {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < nt; ++t) th.emplace_back(work);
      for (auto &x : th) x.join();
    }
    if (firstBad.load() < blocks.size()) {                          // bytes up to the first damaged member are good
      B.data.resize(blocks[firstBad.load()].out);
      B.corruptBehind = true;
    }
    return B;
  }
  void launch_ahead() {
    if (scanBad || nextOff >= mapSize) return;
    std::vector<BgzfBlock> blocks = scan();
    if (blocks.empty() && !scanBad) return;
    ahead = std::async(std::launch::async, inflate_batch, map, std::move(blocks), scanBad);
  }

  bool start(const char *path) {
    int fd = ::open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) == 0 && st.st_size > 0) {
      void *m = mmap(nullptr, (size_t) st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
      if (m != MAP_FAILED) {
        map = (const uint8_t *) m; mapSize = (size_t) st.st_size;
        if (member_size(map, mapSize, 0)) {
          ::close(fd);
          madvise(m, mapSize, MADV_SEQUENTIAL);
          parallel = true; open = true;
          if (const char *e = getenv("GT16_BGZF_BATCH")) { const long v = atol(e); if (v >= 1) batchBytes = (size_t) v; }
          launch_ahead();
          return true;
        }
        munmap(m, mapSize); map = nullptr;
      }
    }
    ::close(fd);
    f = fopen(path, "rb");
    if (!f) return false;
    in.resize(1u << 20);
    memset(&z, 0, sizeof(z));
    if (inflateInit2(&z, 15 + 16) != Z_OK) return false;      // gzip wrapper
    open = true;
    return true;
  }
  // (parallel mode) the unread rest of the current batch, else the whole next batch; false at the end of the file or
  // at a damaged member (bad set).  The bytes stay where they are until the next call.
  bool span(const uint8_t **p, size_t *n) {
    while (curPos == cur.data.size()) {
      if (cur.corruptBehind) { bad = true; return false; }
      if (!ahead.valid()) return false;
      cur = ahead.get();
      curPos = 0;
      launch_ahead();
    }
    *p = cur.data.data() + curPos; *n = cur.data.size() - curPos;
    curPos = cur.data.size();
    return true;
  }
  // exactly n bytes, false at a clean end of file (with none read) or on error (bad set)
  bool get(void *dst, size_t n) {
    if (parallel) {
      size_t done = 0;
      while (done < n) {
        if (curPos == cur.data.size()) {
          if (cur.corruptBehind) { bad = true; return false; }
          if (!ahead.valid()) { if (done) bad = true; return false; }
          cur = ahead.get();
          curPos = 0;
          launch_ahead();
          continue;
        }
        const size_t k = std::min(n - done, cur.data.size() - curPos);
        memcpy((char *) dst + done, cur.data.data() + curPos, k);
        done += k; curPos += k;
      }
      return true;
    }
    z.next_out = (Bytef *) dst;
    z.avail_out = (uInt) n;
    while (z.avail_out) {
      if (!z.avail_in && !eof) {
        const size_t k = fread(in.data(), 1, in.size(), f);
        if (k < in.size()) eof = true;
        z.next_in = in.data(); z.avail_in = (uInt) k;
      }
      if (!z.avail_in && eof) { if (z.avail_out != n) bad = true; return false; }
      const int rc = inflate(&z, Z_NO_FLUSH);
      if (rc == Z_STREAM_END) { if (inflateReset(&z) != Z_OK) { bad = true; return false; } }
      else if (rc != Z_OK && rc != Z_BUF_ERROR) { bad = true; return false; }
    }
    return true;
  }
  ~BgzfReader() {
    if (ahead.valid()) ahead.wait();
    if (parallel) { if (map) munmap((void *) map, mapSize); }
    else if (open) inflateEnd(&z);
    if (f) fclose(f);
  }
};
}  // namespace

// One alignment record (without its block_size field) into S the way bamToAlignment / updateSnapshotForFile see it;
// false when the record's own sizes do not fit into it.
struct BamRecordParser {
  const gt16_index *ix;
  uint32_t L, mapqThreshold;
  const int *chrId;                  // reference id of the BAM -> chromosome id of the index (chrIdMap, bam2snapshot.cpp:157-158), -1: none
  int32_t nRef;
  bool operator()(const uint8_t *rec, size_t size, gt16_sam *S) const {
    static const uint8_t code[16] = {0, 0, 1, 0, 2, 0, 0, 0, 3, 0, 0, 0, 0, 0, 0, 2};   // bamcode2snpcode, bam2snapshot.h:52-58
    S->total++;
    int32_t tid, pos, lSeq;
    memcpy(&tid, &rec[0], 4); memcpy(&pos, &rec[4], 4);
    const uint32_t lName = rec[8], mapq = rec[9];
    uint16_t nCig, flag;
    memcpy(&nCig, &rec[12], 2); memcpy(&flag, &rec[14], 2); memcpy(&lSeq, &rec[16], 4);
    const size_t need = 32 + (size_t) lName + (size_t) nCig * 4 + ((size_t) (lSeq > 0 ? lSeq : 0) + 1) / 2 + (size_t) (lSeq > 0 ? lSeq : 0);
    if (need > size) return false;
    if (mapq < mapqThreshold) return true;                                        // bam2snapshot.cpp:84-87
    if (tid < 0 || tid >= nRef || pos < 0 || pos == INT32_MAX) return true;       // :91-94
    if (flag & 0x400) return true;                                                // :171-174
    if (lSeq <= 0) return true;                                                   // :181
    const uint8_t *cig = &rec[32 + lName], *seq = cig + (size_t) nCig * 4, *qual = seq + ((size_t) lSeq + 1) / 2;
    size_t fromCigar = 0;
    for (uint32_t c = 0; c < nCig; ++c) {
      uint32_t v; memcpy(&v, cig + 4 * c, 4);
      const uint32_t op = v & 15u;
      if (op == 0 || op == 1 || op == 4) fromCigar += v >> 4;                     // getLenFromCigar (:55-77)
    }
    if (!nCig || fromCigar != (size_t) lSeq) return true;                         // :181-188
    uint32_t amb;
    if (chrId[tid] < 1 || gt16_index_locate_id(ix, chrId[tid], (uint32_t) pos + 1, &amb) != 0 || amb >= L) return true;
    S->start.push_back(amb);
    S->reverse.push_back((flag & 16) ? 1 : 0);
    S->length.push_back((uint32_t) lSeq);
    S->seqOff.push_back(S->bases.size());
    S->cigarOff.push_back(S->cigar.size());
    S->nCigar.push_back(nCig);
    const size_t c0 = S->cigar.size();
    S->cigar.resize(c0 + nCig);
    memcpy(S->cigar.data() + c0, cig, (size_t) nCig * 4);
    const size_t at = S->bases.size();
    S->bases.resize(at + (size_t) lSeq); S->quals.resize(at + (size_t) lSeq);
    uint8_t *bo = S->bases.data() + at;
    for (int32_t k = 0; k + 1 < lSeq; k += 2) { const uint8_t b = seq[k >> 1]; bo[k] = code[b >> 4]; bo[k + 1] = code[b & 15u]; }   // bam1_seqi
    if (lSeq & 1) bo[lSeq - 1] = code[seq[(lSeq - 1) >> 1] >> 4];
    memcpy(S->quals.data() + at, qual, (size_t) lSeq);
    return true;
  }
};

int gt16_bam_read(const char *path, const gt16_index *ix, uint32_t mapqThreshold, gt16_sam **out, char *err, size_t errLen) {
  if (!path || !ix || !out) return GT16_EINVAL;
  BgzfReader R;
  if (!R.start(path)) { set_err(err, errLen, "[ERROR] Failed accessing %s.", path); return GT16_EIO; }
  auto fail = [&](const char *what) { set_err(err, errLen, "%s: %s", path, what); return GT16_EDATA; };
  char magic[4];
  int32_t lText = 0, nRef = 0;
  if (!R.get(magic, 4) || memcmp(magic, "BAM\1", 4) != 0) return fail("not a BAM file");
  if (!R.get(&lText, 4) || lText < 0) return fail("truncated BAM header");
  // sizes in a damaged file may say anything: nothing is allocated ahead of the bytes that are really there
  {
    std::vector<char> text(1u << 16);
    for (size_t left = (size_t) lText; left;) {
      const size_t k = std::min(left, text.size());
      if (!R.get(text.data(), k)) return fail("truncated BAM header");
      left -= k;
    }
  }
  if (!R.get(&nRef, 4) || nRef < 0) return fail("truncated BAM header");
  std::vector<int> chrId;                                       // 0: not looked up yet (chrIdMap, bam2snapshot.cpp:157-158)
  std::vector<std::string> names;
  for (int32_t r = 0; r < nRef; ++r) {
    int32_t lName = 0, lRef = 0;
    if (!R.get(&lName, 4) || lName <= 0 || lName > (1 << 16)) return fail("bad reference name in the BAM header");
    std::vector<char> nm((size_t) lName);
    if (!R.get(nm.data(), (size_t) lName) || !R.get(&lRef, 4)) return fail("truncated BAM header");
    names.push_back(std::string(nm.data(), strnlen(nm.data(), (size_t) lName)));
    chrId.push_back(0);
  }
  for (int32_t r = 0; r < nRef; ++r) { const int id = gt16_index_sequence_id(ix, names[r].c_str()); chrId[r] = id < 1 ? -1 : id; }
  const BamRecordParser parse{ix, gt16_index_length(ix), mapqThreshold, chrId.data(), nRef};
  gt16_sam *S = new gt16_sam();
  if (!R.parallel) {                                            // one streaming inflate: record after record
    std::vector<uint8_t> rec;
    for (;;) {
      int32_t blockSize = 0;
      if (!R.get(&blockSize, 4)) { if (R.bad) { delete S; return fail("corrupt BGZF stream"); } break; }
      if (blockSize < 32) { delete S; return fail("bad alignment record"); }
      rec.clear();
      for (size_t left = (size_t) blockSize; left;) {           // in steps: a damaged size does not reserve 2 GB
        const size_t k = std::min<size_t>(left, 1u << 20), at = rec.size();
        rec.resize(at + k);
        if (!R.get(rec.data() + at, k)) { delete S; return fail("truncated alignment record"); }
        left -= k;
      }
      if (!parse(rec.data(), rec.size(), S)) { delete S; return fail("bad alignment record"); }
    }
    *out = S;
    return GT16_OK;
  }
  // BGZF: what a batch of members inflates to is cut at record boundaries (a hop from size field to size field) and the
  // records are parsed by several threads into lists of their own, joined in file order at the end; a record that
  // straddles two batches is put together in `carry`
  std::vector<gt16_sam> parts;
  Vec<uint8_t> carry;
  std::vector<size_t> starts;
  unsigned nt = std::thread::hardware_concurrency();
  nt = std::max(1u, std::min(16u, nt ? nt : 4u));
  const uint8_t *p = nullptr;
  size_t n = 0;
  bool malformed = false;
  while (!malformed && R.span(&p, &n)) {
    size_t pos = 0;
    if (!carry.empty()) {                                       // finish the record the last batch ended in
      while (carry.size() < 4 && pos < n) carry.push_back(p[pos++]);
      if (carry.size() < 4) continue;
      int32_t blockSize;
      memcpy(&blockSize, carry.data(), 4);
      if (blockSize < 32) { malformed = true; break; }
      const size_t want = 4 + (size_t) blockSize, take = std::min(want - carry.size(), n - pos);
      carry.insert(carry.end(), p + pos, p + pos + take);
      pos += take;
      if (carry.size() < want) continue;
      parts.emplace_back();
      if (!parse(carry.data() + 4, (size_t) blockSize, &parts.back())) { malformed = true; break; }
      carry.clear();
    }
    starts.clear();
    while (pos + 4 <= n) {
      int32_t blockSize;
      memcpy(&blockSize, p + pos, 4);
      if (blockSize < 32) { malformed = true; break; }
      if (pos + 4 + (size_t) blockSize > n) break;
      starts.push_back(pos);
      pos += 4 + (size_t) blockSize;
    }
    if (malformed) break;
    carry.assign(p + pos, p + n);
    const unsigned use = starts.size() < 4096 ? 1u : nt;
    const size_t first = parts.size();
    parts.resize(first + use);
    std::vector<char> ok(use, 1);
    auto work = [&](unsigned t) {
      const size_t lo = starts.size() * t / use, hi = starts.size() * (t + 1) / use;
      gt16_sam &P = parts[first + t];
      if (hi > lo) {
        const size_t bytes = (hi < starts.size() ? starts[hi] : pos) - starts[lo];
        P.bases.reserve(bytes * 2 / 3); P.quals.reserve(bytes * 2 / 3);
        for (auto *v : {&P.start, &P.length, &P.nCigar}) v->reserve(hi - lo);
      }
      for (size_t k = lo; k < hi; ++k) {
        int32_t blockSize;
        memcpy(&blockSize, p + starts[k], 4);
        if (!parse(p + starts[k] + 4, (size_t) blockSize, &P)) { ok[t] = 0; return; }
      }
    };
    if (use == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < use; ++t) th.emplace_back(work, t);
      for (auto &x : th) x.join();
    }
    for (char c : ok) if (!c) malformed = true;
  }
  if (malformed) { delete S; return fail("bad alignment record"); }
  if (R.bad) { delete S; return fail("corrupt BGZF stream"); }
  if (!carry.empty()) { delete S; return fail("truncated alignment record"); }
  if (!parts.empty()) join_parts(parts, S);
  *out = S;
  return GT16_OK;
}

void gt16_sam_reads(const gt16_sam *S, gt16_reads *r) {
  memset(r, 0, sizeof(*r));
  r->n_reads = S->start.size();
  r->start = S->start.data(); r->reverse = S->reverse.data(); r->length = S->length.data();
  r->seq_off = S->seqOff.data(); r->cigar_off = S->cigarOff.data(); r->n_cigar = S->nCigar.data();
  r->n_bases = S->bases.size(); r->n_cigar_ops = S->cigar.size();
  r->bases = S->bases.data(); r->quals = S->quals.data(); r->cigar = S->cigar.data();
}
uint64_t gt16_sam_total_alignments(const gt16_sam *S) { return S ? S->total : 0; }
void gt16_sam_free(gt16_sam *S) { delete S; }

// ---- writeSnpInfoSnapshot (SnapshotHandler.cpp:83-236) ---------------------------------------------------------
int gt16_snapshot_write(const char *path, const gt16_snapshot_view *v, char *err, size_t errLen) {
  if (!path || !v || !v->counters || !v->text_length) return GT16_EINVAL;
  const uint32_t L = v->text_length;
  const size_t page = (size_t) sysconf(_SC_PAGE_SIZE);
  // populated positions (isLowPopulation, bam2snapshot.cpp:527-561): overflow-type, or any non-zero field
  size_t populated = 0;
  {
    unsigned nt = std::thread::hardware_concurrency();
    nt = L < (1u << 24) ? 1u : std::max(1u, std::min(16u, nt ? nt : 4u));
    std::vector<size_t> cnt(nt, 0);
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t]() {
      size_t c = 0;
      for (uint64_t p = (uint64_t) L * t / nt, e = (uint64_t) L * (t + 1) / nt; p < e; ++p) c += v->counters[p] != 0;
      cnt[t] = c;
    });
    for (auto &x : th) x.join();
    for (size_t c : cnt) populated += c;
  }
  const bool sparse = (double) populated / L < 0.1;
  std::vector<uint32_t> keep;                       // sparse: positions within 100 of a populated one (writeHeader :103-122)
  if (sparse) {
    uint64_t next = 0;                              // first position not yet kept
    for (uint32_t p = 0; p < L; ++p) if (v->counters[p]) {
      const uint64_t s = p > 100 ? p - 100 : 0, e = p < L - 100 ? (uint64_t) p + 100 : (uint64_t) L - 1;
      for (uint64_t j = s > next ? s : next; j <= e; ++j) keep.push_back((uint32_t) j);
      if (e + 1 > next) next = e + 1;
    }
  }
  const size_t poolCap = ((size_t) v->pool_bytes + page) / page * page;
  size_t len[5], start[5];
  len[0] = 4 + 8 + poolCap;
  len[1] = 1 + 4 + (sparse ? keep.size() * 12 : (size_t) L * 8);
  len[2] = 4;
  for (uint32_t r = 0; r < v->n_overflow_regions; ++r) len[2] += 8 + (size_t) v->overflow_size[r] * 32;
  len[3] = sparse ? 0 : (((size_t) L + 31) >> 5) * 4;
  len[4] = 0;                                       // directional soft-clip counters: not used by the caller
  size_t cur = page;
  for (int i = 0; i < 5; ++i) { start[i] = cur; cur += (len[i] + page - 1) / page * page; }
  FILE *f = fopen(path, "wb");
  if (!f) { set_err(err, errLen, "cannot write %s", path); return GT16_EIO; }
  bool ok = true;
  std::vector<uint8_t> hdr(page, 0);
  const int pg = (int) page;
  memcpy(&hdr[0], &pg, 4);
  for (int i = 0; i < 5; ++i) { memcpy(&hdr[4 + i * 16], &start[i], 8); memcpy(&hdr[12 + i * 16], &len[i], 8); }
  ok = ok && fwrite(hdr.data(), 1, page, f) == page;
  // subfile 0: pool
  ok = ok && fseeko(f, (off_t) start[0], SEEK_SET) == 0;
  const uint32_t curPtr = v->pool_bytes;
  const uint64_t cap64 = poolCap;
  ok = ok && fwrite(&curPtr, 4, 1, f) == 1 && fwrite(&cap64, 8, 1, f) == 1;
  if (curPtr) ok = ok && fwrite(v->pool, 1, curPtr, f) == curPtr;
  { std::vector<uint8_t> z(poolCap - curPtr, 0); if (!z.empty()) ok = ok && fwrite(z.data(), 1, z.size(), f) == z.size(); }
  // subfile 1: counters
  ok = ok && fseeko(f, (off_t) start[1], SEEK_SET) == 0;
  const char mode = sparse ? 2 : 1;
  ok = ok && fwrite(&mode, 1, 1, f) == 1 && fwrite(&L, 4, 1, f) == 1;
  if (!sparse) ok = ok && fwrite(v->counters, 8, L, f) == L;
  else {
    std::vector<uint8_t> buf;
    buf.reserve((size_t) 12 << 16);
    for (size_t i = 0; i < keep.size(); ++i) {
      const uint32_t p = keep[i];
      const uint64_t c = v->counters[p];
      const size_t at = buf.size();
      buf.resize(at + 12);
      memcpy(&buf[at], &p, 4); memcpy(&buf[at + 4], &c, 8);
      if (buf.size() >= ((size_t) 12 << 16) || i + 1 == keep.size()) { ok = ok && fwrite(buf.data(), 1, buf.size(), f) == buf.size(); buf.clear(); }
    }
  }
  // subfile 2: overflow arrays
  ok = ok && fseeko(f, (off_t) start[2], SEEK_SET) == 0;
  const uint32_t nr = v->n_overflow_regions;
  ok = ok && fwrite(&nr, 4, 1, f) == 1;
  for (uint32_t r = 0; r < nr; ++r) {
    const uint32_t sz = v->overflow_size[r], lim = sz + 16;
    ok = ok && fwrite(&sz, 4, 1, f) == 1 && fwrite(&lim, 4, 1, f) == 1;
    if (sz) ok = ok && fwrite(v->overflow[r], 32, sz, f) == sz;
  }
  // subfile 3: invalid bitmap
  if (!sparse) {
    ok = ok && fseeko(f, (off_t) start[3], SEEK_SET) == 0;
    const size_t w = ((size_t) L + 31) >> 5;
    if (v->invalid_bitmap) ok = ok && fwrite(v->invalid_bitmap, 4, w, f) == w;
    else { std::vector<uint32_t> z(w, 0); ok = ok && fwrite(z.data(), 4, w, f) == w; }
  }
  // every subfile is addressable up to its page-rounded end
  ok = ok && fseeko(f, 0, SEEK_END) == 0;
  const off_t endNow = ftello(f);
  if (ok && (size_t) endNow < cur) { std::vector<uint8_t> z(cur - (size_t) endNow, 0); ok = fwrite(z.data(), 1, z.size(), f) == z.size(); }
  if (fclose(f) != 0) ok = false;
  if (!ok) { set_err(err, errLen, "error writing %s", path); return GT16_EIO; }
  return GT16_OK;
}

}  // extern "C"
