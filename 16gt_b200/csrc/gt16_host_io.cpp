// gt16_host_io.cpp -- host-side formats on either side of the hot path: snapshot container
// reader, packed-reference index reader, coordinate translation and the .txt row printer.
// Plain C++ in the reference's own style; no CUDA here except optional page-locking of the
// big buffers (so cudaMemcpyAsync from them is truly asynchronous).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <stdarg.h>
#include <math.h>
#include <float.h>
#include <cmath>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <algorithm>
#include <mutex>
#include <condition_variable>
#include <string>
#include <thread>
#include <vector>
#include <cuda_runtime_api.h>
#include "../../include/gt16.h"

namespace {

void set_err(char *err, size_t n, const char *fmt, ...) {
  if (!err || !n) return;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err, n, fmt, ap);
  va_end(ap);
}

// big host buffer: page-locked on request, or a read-only private mapping of a byte range of a file
struct HostBuf {
  void *p = nullptr;
  size_t bytes = 0;
  bool pinned = false;
  void *mapBase = nullptr;                 // mmap()ed range that holds p (map_file)
  size_t mapBytes = 0;
  bool alloc(size_t n, bool pin) {
    bytes = n ? n : 1;
    if (pin && cudaHostAlloc(&p, bytes, cudaHostAllocDefault) == cudaSuccess) { pinned = true; return true; }
    (void) cudaGetLastError();
    pinned = false;
    p = malloc(bytes);
    return p != nullptr;
  }
  // p = the n bytes at offset `off` of the file, straight from the page cache: no copy into anonymous memory.
  // (The library only ever copies out of the counter array; a page-locked ring stands between it and the DMA.)
  bool map_file(const char *path, uint64_t off, size_t n) {
    const long page = sysconf(_SC_PAGESIZE);
    const uint64_t lo = off & ~(uint64_t) (page - 1);
    int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    struct stat st;
    if (fstat(fd, &st) != 0 || (uint64_t) st.st_size < off + n) { close(fd); return false; }
    void *m = mmap(nullptr, (size_t) (off - lo) + n, PROT_READ, MAP_PRIVATE, fd, (off_t) lo);
    close(fd);
    if (m == MAP_FAILED) return false;
    madvise(m, (size_t) (off - lo) + n, MADV_SEQUENTIAL);
    mapBase = m; mapBytes = (size_t) (off - lo) + n;
    p = (char *) m + (off - lo);
    bytes = n;
    pinned = false;
    return true;
  }
  void release() {
    if (!p) return;
    if (mapBase) munmap(mapBase, mapBytes);
    else if (pinned) cudaFreeHost(p);
    else free(p);
    p = nullptr; mapBase = nullptr;
  }
};

bool read_at(FILE *f, uint64_t off, void *dst, size_t n) {
  if (fseeko(f, (off_t) off, SEEK_SET) != 0) return false;
  return fread(dst, 1, n, f) == n;
}

unsigned io_threads() {
  unsigned n = std::thread::hardware_concurrency();
  return std::max(1u, std::min(16u, n ? n : 4u));
}

// big contiguous read: the range is cut into slices that worker threads pread() concurrently
// (snapshot ingest is row (f)1 of SURVEY 8: it bounds the drop-in once the kernels are fast)
bool pread_parallel(const char *path, uint64_t off, void *dst, size_t n) {
  if (n < (8u << 20)) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return false;
    size_t done = 0;
    while (done < n) { ssize_t k = pread(fd, (char *) dst + done, n - done, (off_t) (off + done)); if (k <= 0) break; done += (size_t) k; }
    close(fd);
    return done == n;
  }
  const unsigned nt = io_threads();
  const size_t slice = ((n + nt - 1) / nt + 4095) & ~(size_t) 4095;
  std::vector<char> ok(nt, 1);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back([&, t]() {
    size_t a = (size_t) t * slice, b = std::min(n, a + slice);
    if (a >= b) return;
    int fd = open(path, O_RDONLY);
    if (fd < 0) { ok[t] = 0; return; }
    while (a < b) { ssize_t k = pread(fd, (char *) dst + a, b - a, (off_t) (off + a)); if (k <= 0) { ok[t] = 0; break; } a += (size_t) k; }
    close(fd);
  });
  for (auto &x : th) x.join();
  for (char c : ok) if (!c) return false;
  return true;
}

}  // namespace

struct gt16_host_snapshot {
  uint32_t textLength = 0;
  uint32_t posLo = 0, posHi = 0;           // counters.p holds positions [posLo, posHi) (gt16_snapshot_read_range)
  uint32_t nRegions = 0;
  HostBuf counters, invalid, pool, ovfAll, sparse;   // sparse: the {position, counter} records of a mode-2 snapshot kept as they are
  uint64_t nSparse = 0;
  bool hasInvalid = false;
  uint32_t poolBytes = 0;
  std::vector<const void *> ovfPtr;
  std::vector<uint32_t> ovfSize;
};

struct gt16_index {
  uint32_t length = 0;
  std::vector<uint32_t> packed;            // 16 bases per word, MSB first (HSP.packedDNA)
  std::vector<std::string> names;          // annotation[i].text
  struct Tr { uint32_t startPos; uint16_t chrID; uint32_t correction; };
  std::vector<Tr> translate;               // HSP.translate
  // gt16_index_locate_id is called once per read by the SAM / BAM readers: when the table is laid out the way the index
  // builder writes it (a chromosome's entries side by side, in ascending order) the entry is found by bisection
  std::vector<uint32_t> chrLo, chrHi;      // translate[chrLo[id] .. chrHi[id]) belong to chromosome id
  std::vector<uint32_t> chrBegin;          // chromosome position (1-based) at which entry k begins: startPos - correction
  bool bisect = false;
  bool pinned = false;                     // packed[] is page-locked (gt16_index_pin)
};

extern "C" {

// ---- snapshot container (SnapshotHandler.cpp:238-347) -----------------------------------
int gt16_snapshot_read(const char *path, int pinned, gt16_host_snapshot **out, char *err, size_t errLen) {
  return gt16_snapshot_read_range(path, pinned, 0, 0xFFFFFFFFu, out, err, errLen);
}

int gt16_snapshot_read_range(const char *path, int pinned, uint32_t pos_lo, uint32_t pos_hi, gt16_host_snapshot **out,
                             char *err, size_t errLen) {
  return gt16_snapshot_read_ex(path, pinned ? GT16_READ_PINNED : 0, pos_lo, pos_hi, out, err, errLen);
}

int gt16_snapshot_read_ex(const char *path, int flags, uint32_t pos_lo, uint32_t pos_hi, gt16_host_snapshot **out,
                          char *err, size_t errLen) {
  const int pinned = (flags & GT16_READ_PINNED) ? 1 : 0;
  const bool keepSparse = (flags & GT16_READ_KEEP_SPARSE) != 0;
  if (!path || !out) return GT16_EINVAL;
  FILE *f = fopen(path, "rb");
  if (!f) { set_err(err, errLen, "Error opening %s", path); return GT16_EIO; }
  gt16_host_snapshot *hs = new gt16_host_snapshot();
  int rc = GT16_OK;
  uint8_t hdr[4 + 5 * 16];
  uint64_t start[5], len[5];
  uint64_t fileSize = 0;                    // sizes read from a damaged file are checked against it before anything is allocated
  { struct stat st; if (fstat(fileno(f), &st) == 0) fileSize = (uint64_t) st.st_size; }
  do {
    if (!read_at(f, 0, hdr, sizeof(hdr))) { set_err(err, errLen, "%s: short header", path); rc = GT16_EIO; break; }
    for (int i = 0; i < 5; ++i) { memcpy(&start[i], hdr + 4 + 16 * i, 8); memcpy(&len[i], hdr + 12 + 16 * i, 8); }
    // subfile 1: counters (dense mode 1, sparse mode 2)
    uint8_t head[5];
    if (!read_at(f, start[1], head, 5)) { set_err(err, errLen, "%s: no counter subfile", path); rc = GT16_EIO; break; }
    char mode = (char) head[0];
    memcpy(&hs->textLength, head + 1, 4);
    const uint32_t L = hs->textLength;
    if (!L) { set_err(err, errLen, "%s: empty snapshot", path); rc = GT16_EDATA; break; }
    if ((mode & 1) && (uint64_t) L * 8 > fileSize) { set_err(err, errLen, "%s: truncated counters", path); rc = GT16_EIO; break; }
    // the window of counters kept in memory (a shard of a sharded genome reads and page-locks its own part only)
    const uint32_t wLo = pos_lo < L ? pos_lo : 0, wHi = pos_hi > L || pos_hi <= wLo ? L : pos_hi;
    const size_t W = (size_t) wHi - wLo;
    hs->posLo = wLo; hs->posHi = wHi;
    // Dense counters of a pageable snapshot are mapped, not read: the only consumer copies them chunk by chunk into
    // page-locked memory on their way to the GPU, so a private copy would just be one more pass over 8 bytes per
    // site of host memory (the slowest resource of the drop-in binary on a 3.1 Gb genome).
    const bool mapped = (mode & 1) && !pinned && !getenv("GT16_NO_MMAP") && hs->counters.map_file(path, start[1] + 5 + (uint64_t) wLo * 8, W * 8);
    const bool asRecords = !(mode & 1) && (mode & 2) && keepSparse;
    if (!mapped && !asRecords && !hs->counters.alloc(W * 8 + 16, pinned != 0)) { set_err(err, errLen, "Memory Allocation Failed!"); rc = GT16_EIO; break; }
    if (mapped) {
    } else if (asRecords) {
      // the records stay as the producer wrote them (12 bytes each, ascending positions): they are expanded in HBM, not
      // here -- a whole-genome array of zeros for an exome's worth of counters would cost 8 B per genome position of host
      // memory and of host-link traffic
      if (len[1] < 5 || (len[1] - 5) % 12 || len[1] > fileSize) { set_err(err, errLen, "%s: bad sparse counter subfile", path); rc = GT16_EDATA; break; }
      hs->nSparse = (len[1] - 5) / 12;
      const size_t bytes = (size_t) hs->nSparse * 12;
      const bool mappedRecords = !pinned && !getenv("GT16_NO_MMAP") && bytes && hs->sparse.map_file(path, start[1] + 5, bytes);
      if (!mappedRecords) {
        if (!hs->sparse.alloc(bytes + 16, pinned != 0)) { set_err(err, errLen, "Memory Allocation Failed!"); rc = GT16_EIO; break; }
        if (bytes && !pread_parallel(path, start[1] + 5, hs->sparse.p, bytes)) { set_err(err, errLen, "%s: truncated counters", path); rc = GT16_EIO; break; }
      }
      if (fseeko(f, (off_t) (start[1] + len[1]), SEEK_SET) != 0) { set_err(err, errLen, "%s: bad sparse counter subfile", path); rc = GT16_EIO; break; }
    } else if (mode & 1) {
      if (!pread_parallel(path, start[1] + 5 + (uint64_t) wLo * 8, hs->counters.p, W * 8)) { set_err(err, errLen, "%s: truncated counters", path); rc = GT16_EIO; break; }
      memset((char *) hs->counters.p + W * 8, 0, 16);
    } else if (mode & 2) {
      memset(hs->counters.p, 0, W * 8 + 16);
      uint64_t n = len[1] / 12;
      std::vector<uint8_t> buf((size_t) 12 << 16);
      uint64_t done = 0;
      bool ok = true;
      while (done < n && ok) {
        size_t chunk = (size_t) ((n - done) < (1u << 16) ? (n - done) : (1u << 16));
        if (fread(buf.data(), 12, chunk, f) != chunk) { ok = false; break; }
        for (size_t k = 0; k < chunk; ++k) {
          uint32_t pos; memcpy(&pos, &buf[k * 12], 4);
          if (pos >= L) { ok = false; break; }
          if (pos < wLo || pos >= wHi) continue;
          memcpy((char *) hs->counters.p + (size_t) (pos - wLo) * 8, &buf[k * 12 + 4], 8);
        }
        done += chunk;
      }
      if (!ok) { set_err(err, errLen, "%s: bad sparse counter subfile", path); rc = GT16_EDATA; break; }
    } else { set_err(err, errLen, "%s: unknown counter mode %d", path, (int) mode); rc = GT16_EDATA; break; }
    // subfile 2: overflow arrays
    if (!read_at(f, start[2], &hs->nRegions, 4) || hs->nRegions == 0 || hs->nRegions > 4096) {
      set_err(err, errLen, "%s: bad overflow subfile", path); rc = GT16_EDATA; break;
    }
    {
      std::vector<std::vector<uint8_t> > tmp(hs->nRegions);
      size_t total = 0;
      bool ok = true, bad = false;
      for (uint32_t r = 0; r < hs->nRegions && ok; ++r) {
        uint32_t sz, lim;
        if (fread(&sz, 4, 1, f) != 1 || fread(&lim, 4, 1, f) != 1) { ok = false; break; }
        // the writer stores `size` records behind every (size, limit) pair (SnapshotHandler.cpp:214-219) and size <= limit
        // always holds there; anything else is not a snapshot the reference could have written (its reader would skip the
        // records of a limit-0 region and mis-parse the rest, :313)
        if (sz > lim) { bad = true; break; }
        if ((uint64_t) sz * 32 > fileSize) { ok = false; break; }
        hs->ovfSize.push_back(sz);
        if (lim == 0) continue;                       // SnapshotHandler.cpp:313
        tmp[r].resize((size_t) sz * 32);
        if (sz && fread(tmp[r].data(), 32, sz, f) != sz) { ok = false; break; }
        total += (size_t) sz * 32;
      }
      if (bad) { set_err(err, errLen, "%s: overflow region with more records than its limit", path); rc = GT16_EDATA; break; }
      if (!ok) { set_err(err, errLen, "%s: truncated overflow arrays", path); rc = GT16_EIO; break; }
      if (!hs->ovfAll.alloc(total + 32, pinned != 0)) { set_err(err, errLen, "Memory Allocation Failed!"); rc = GT16_EIO; break; }
      size_t off = 0;
      for (uint32_t r = 0; r < hs->nRegions; ++r) {
        hs->ovfPtr.push_back((char *) hs->ovfAll.p + off);
        if (!tmp[r].empty()) memcpy((char *) hs->ovfAll.p + off, tmp[r].data(), tmp[r].size());
        off += tmp[r].size();
      }
    }
    // subfile 3: invalid bitmap (absent in sparse snapshots)
    size_t bmWords = ((size_t) L + 31) >> 5;
    if (!hs->invalid.alloc(bmWords * 4 + 4, pinned != 0)) { set_err(err, errLen, "Memory Allocation Failed!"); rc = GT16_EIO; break; }
    if (len[3]) {
      if (!pread_parallel(path, start[3], hs->invalid.p, bmWords * 4)) { set_err(err, errLen, "%s: truncated bitmap", path); rc = GT16_EIO; break; }
      hs->hasInvalid = true;
    } else {
      memset(hs->invalid.p, 0, bmWords * 4);
    }
    // subfile 0: memory pool; only the first curPtr bytes carry data
    uint8_t ph[12];
    if (!read_at(f, start[0], ph, 12)) { set_err(err, errLen, "%s: no pool subfile", path); rc = GT16_EIO; break; }
    uint32_t curPtr; uint64_t capacity;
    memcpy(&curPtr, ph, 4); memcpy(&capacity, ph + 4, 8);
    if (curPtr > capacity) { set_err(err, errLen, "%s: pool pointer beyond capacity", path); rc = GT16_EDATA; break; }
    if (curPtr > fileSize) { set_err(err, errLen, "%s: truncated pool", path); rc = GT16_EIO; break; }
    hs->poolBytes = curPtr;
    if (!hs->pool.alloc((size_t) curPtr + 16, pinned != 0)) { set_err(err, errLen, "Memory Allocation Failed!"); rc = GT16_EIO; break; }
    if (curPtr && fread(hs->pool.p, 1, curPtr, f) != curPtr) { set_err(err, errLen, "%s: truncated pool", path); rc = GT16_EIO; break; }
  } while (0);
  fclose(f);
  if (rc != GT16_OK) { gt16_snapshot_free(hs); return rc; }
  *out = hs;
  return GT16_OK;
}

void gt16_snapshot_view_of(const gt16_host_snapshot *hs, gt16_snapshot_view *v) {
  memset(v, 0, sizeof(*v));
  // indexed by genome position whatever the window: with a window the pointer is biased and only positions
  // [pos_lo, pos_hi) may be dereferenced (calls whose range lies inside the window touch nothing else)
  v->counters = hs->counters.p ? (const uint64_t *) hs->counters.p - hs->posLo : nullptr;
  v->sparse_entries = hs->counters.p ? nullptr : hs->sparse.p;
  v->n_sparse_entries = hs->counters.p ? 0 : hs->nSparse;
  v->text_length = hs->textLength;
  v->invalid_bitmap = hs->hasInvalid ? (const uint32_t *) hs->invalid.p : nullptr;
  v->n_overflow_regions = hs->nRegions;
  v->overflow = hs->ovfPtr.data();
  v->overflow_size = hs->ovfSize.data();
  v->pool = (const uint8_t *) hs->pool.p;
  v->pool_bytes = hs->poolBytes;
}

void gt16_snapshot_free(gt16_host_snapshot *hs) {
  if (!hs) return;
  hs->counters.release(); hs->invalid.release(); hs->pool.release(); hs->ovfAll.release(); hs->sparse.release();
  delete hs;
}

// ---- index files (.pac/.ann/.tra) -----------------------------------------------------
int gt16_index_read(const char *prefix, gt16_index **out, char *err, size_t errLen) {
  if (!prefix || !out) return GT16_EINVAL;
  std::string pre(prefix);
  gt16_index *ix = new gt16_index();
  int rc = GT16_OK;
  do {
    // .ann: "%u %d %u" then per sequence "%u %s" (gi, name up to end of line) and "%u %u %d" (readIndex.cpp:96-115)
    FILE *f = fopen((pre + ".ann").c_str(), "r");
    if (!f) { set_err(err, errLen, "Cannot open %s.ann", prefix); rc = GT16_EIO; break; }
    unsigned len = 0, seed = 0; int nSeq = 0;
    if (fscanf(f, "%u %d %u", &len, &nSeq, &seed) != 3 || nSeq <= 0) { fclose(f); set_err(err, errLen, "%s.ann: bad header", prefix); rc = GT16_EDATA; break; }
    ix->length = len;
    char line[4096];
    if (!fgets(line, sizeof(line), f)) { fclose(f); rc = GT16_EDATA; break; }   // rest of header line
    bool ok = true;
    for (int i = 0; i < nSeq; ++i) {
      if (!fgets(line, sizeof(line), f)) { ok = false; break; }
      char *sp = strchr(line, ' ');
      std::string name = sp ? std::string(sp + 1) : std::string(line);
      while (!name.empty() && (name.back() == '\n' || name.back() == '\r')) name.pop_back();
      ix->names.push_back(name);
      if (!fgets(line, sizeof(line), f)) { ok = false; break; }
    }
    fclose(f);
    if (!ok) { set_err(err, errLen, "%s.ann: truncated", prefix); rc = GT16_EDATA; break; }
    // .tra: "%u %d %d %u" header, gridEntries lines, then translate entries (readIndex.cpp:24-60)
    f = fopen((pre + ".tra").c_str(), "r");
    if (!f) { set_err(err, errLen, "Cannot open %s.tra", prefix); rc = GT16_EIO; break; }
    unsigned tl = 0, grid = 0; int ns = 0, removed = 0;
    if (fscanf(f, "%u %d %d %u", &tl, &ns, &removed, &grid) != 4) { fclose(f); set_err(err, errLen, "%s.tra: bad header", prefix); rc = GT16_EDATA; break; }
    for (unsigned g = 0; g < grid; ++g) { unsigned v; if (fscanf(f, "%u", &v) != 1) { ok = false; break; } }
    int nTr = ns + removed;
    for (int i = 0; i < nTr && ok; ++i) {
      unsigned sp, corr; unsigned short chr;
      if (fscanf(f, "%u %hu %u", &sp, &chr, &corr) != 3) { ok = false; break; }
      gt16_index::Tr t; t.startPos = sp; t.chrID = chr; t.correction = corr;
      ix->translate.push_back(t);
    }
    fclose(f);
    if (!ok || ix->translate.empty()) { set_err(err, errLen, "%s.tra: truncated", prefix); rc = GT16_EDATA; break; }
    {
      // bisection gives the scan's answer when (a) a chromosome's entries are contiguous, (b) they begin at ascending
      // chromosome positions and (c) the chromosome positions between two entries' beginnings are at least as many as
      // the first one's bases (removed stretches only ever add to the gap): then no earlier entry can match a position
      // at or behind a later entry's beginning
      const size_t n = ix->translate.size();
      unsigned maxId = 0;
      for (const auto &t : ix->translate) maxId = std::max<unsigned>(maxId, t.chrID);
      ix->chrLo.assign(maxId + 1, 0); ix->chrHi.assign(maxId + 1, 0); ix->chrBegin.resize(n);
      std::vector<char> seen(maxId + 1, 0);
      bool good = true;
      for (size_t k = 0; k < n && good; ++k) {
        const gt16_index::Tr &t = ix->translate[k];
        ix->chrBegin[k] = t.startPos - t.correction;
        if (k && ix->translate[k - 1].chrID == t.chrID) {
          const gt16_index::Tr &q = ix->translate[k - 1];
          good = t.startPos >= q.startPos && ix->chrBegin[k] >= ix->chrBegin[k - 1] &&
                 ix->chrBegin[k] - ix->chrBegin[k - 1] >= t.startPos - q.startPos;
        } else {
          if (seen[t.chrID]) good = false;
          seen[t.chrID] = 1;
          ix->chrLo[t.chrID] = (uint32_t) k;
        }
        ix->chrHi[t.chrID] = (uint32_t) k + 1;
      }
      ix->bisect = good;
    }
    // .pac: 4 bases / byte, first base in bits 7..6; trailing byte = length % 4
    // read straight into the word array (parallel pread), then swap each word's bytes in place: base 0 of a word
    // ends up in bits 31..30
    size_t nBytes = ((size_t) len + 3) / 4;
    size_t nWords = ((size_t) len + 15) / 16 + 2;
    ix->packed.assign(nWords, 0);
    {
      struct stat st;
      if (stat((pre + ".pac").c_str(), &st) != 0) { set_err(err, errLen, "Cannot open %s.pac", prefix); rc = GT16_EIO; break; }
      if ((uint64_t) st.st_size < nBytes || !pread_parallel((pre + ".pac").c_str(), 0, ix->packed.data(), nBytes)) {
        set_err(err, errLen, "%s.pac: truncated", prefix); rc = GT16_EIO; break;
      }
    }
    {
      const size_t nw = (nBytes + 3) / 4;
      const unsigned nt = nw < (1u << 20) ? 1u : io_threads();
      const size_t per = (nw + nt - 1) / nt;
      uint32_t *w = ix->packed.data();
      auto swap = [&](unsigned t) { for (size_t k = t * per, e = std::min(nw, k + per); k < e; ++k) w[k] = __builtin_bswap32(w[k]); };
      if (nt == 1) swap(0);
      else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(swap, t);
        for (auto &x : th) x.join();
      }
    }
    // clear bits beyond the last base so reads past the end see 'A' like a zeroed trailer
    if (len % 16) ix->packed[len / 16] &= ~0u << (32 - 2 * (len % 16));
  } while (0);
  if (rc != GT16_OK) { delete ix; return rc; }
  *out = ix;
  return GT16_OK;
}

uint32_t gt16_index_length(const gt16_index *ix) { return ix->length; }
uint32_t gt16_index_sequences(const gt16_index *ix) { return ix ? (uint32_t) ix->names.size() : 0; }
const uint32_t *gt16_index_packed(const gt16_index *ix, uint32_t *nWords) {
  if (nWords) *nWords = (uint32_t) ix->packed.size();
  return ix->packed.data();
}

// Page-lock the packed reference where it lies, so that host-fed calls copy it with a plain asynchronous DMA
// (from pageable memory the driver stages it synchronously on the calling thread, ahead of the counter chunks).
int gt16_index_pin(gt16_index *ix) {
  if (!ix) return GT16_EINVAL;
  if (ix->pinned || ix->packed.empty()) return GT16_OK;
  if (cudaHostRegister(ix->packed.data(), ix->packed.size() * 4, cudaHostRegisterDefault) != cudaSuccess) {
    (void) cudaGetLastError();
    return GT16_ECUDA;
  }
  ix->pinned = true;
  return GT16_OK;
}

// getChrAndPos: last translate entry starting at or before amb; position = amb - correction
const char *gt16_index_translate(const gt16_index *ix, uint32_t amb, uint32_t *chrPos) {
  size_t lo = 0, hi = ix->translate.size();
  while (hi - lo > 1) { size_t mid = (lo + hi) / 2; if (ix->translate[mid].startPos <= amb) lo = mid; else hi = mid; }
  const gt16_index::Tr &t = ix->translate[lo];
  if (chrPos) *chrPos = amb - t.correction;      // low 32 bits, printed with %u (interpreter.cpp:116)
  size_t c = (size_t) t.chrID - 1;
  return c < ix->names.size() ? ix->names[c].c_str() : "?";
}

int gt16_index_sequence_id(const gt16_index *ix, const char *chrName) {
  if (!ix || !chrName) return -1;
  // getChrIDFromName: exact name, else the name without a leading "chr" (indexFunction.cpp:42-66; its third
  // attempt compares the unprefixed name again and so never matches anything new)
  size_t id = ix->names.size();
  for (size_t i = 0; i < ix->names.size(); ++i) if (ix->names[i] == chrName) { id = i; break; }
  if (id == ix->names.size() && strncmp(chrName, "chr", 3) == 0)
    for (size_t i = 0; i < ix->names.size(); ++i) if (ix->names[i] == chrName + 3) { id = i; break; }
  return id == ix->names.size() ? -1 : (int) id + 1;
}

int gt16_index_locate_id(const gt16_index *ix, int chrID, uint32_t chrPos, uint32_t *amb) {
  if (!ix || !amb || chrID < 1) return -1;
  // getAmbPos: the concatenated position whose translation is (chrID, chrPos)
  if (ix->bisect && (size_t) chrID < ix->chrLo.size() && ix->chrHi[chrID] > ix->chrLo[chrID]) {
    const uint32_t *b = ix->chrBegin.data();
    const size_t k = (size_t) (std::upper_bound(b + ix->chrLo[chrID], b + ix->chrHi[chrID], chrPos) - b);
    if (k == ix->chrLo[chrID]) return -1;                      // in front of the chromosome's first entry
    const gt16_index::Tr &t = ix->translate[k - 1];
    const uint32_t a = chrPos + t.correction;
    const uint64_t next = k < ix->translate.size() ? ix->translate[k].startPos : (uint64_t) ix->length;
    if (a >= t.startPos && (uint64_t) a < next && a < ix->length) { *amb = a; return 0; }
    return -1;
  }
  for (size_t k = 0; k < ix->translate.size(); ++k) {
    const gt16_index::Tr &t = ix->translate[k];
    if ((int) t.chrID != chrID) continue;
    const uint32_t a = chrPos + t.correction;
    const uint64_t next = k + 1 < ix->translate.size() ? ix->translate[k + 1].startPos : (uint64_t) ix->length;
    if (a >= t.startPos && (uint64_t) a < next && a < ix->length) { *amb = a; return 0; }
  }
  return -1;
}

int gt16_index_locate(const gt16_index *ix, const char *chrName, uint32_t chrPos, uint32_t *amb) {
  return gt16_index_locate_id(ix, gt16_index_sequence_id(ix, chrName), chrPos, amb);
}

void gt16_index_free(gt16_index *ix) {
  if (ix && ix->pinned) cudaHostUnregister(ix->packed.data());
  delete ix;
}

// ---- .txt emitter (interpreter.cpp:5-156) ---------------------------------------------
static const char kGenotypeString[][4] = {"AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT",
                                          "AX", "CX", "GX", "TX", "XX", "XY", "NN"};
static const char kSym[] = {'A', 'C', 'G', 'T', 'X', 'Y', 'N'};

// one row of the .txt (printPossibleSNP, interpreter.cpp:107-156) appended to `out`; lq / rq are the values the
// reference would print for this row (its ReadQuality variable survives rows that do not assign it)
namespace {
// ---- number formatting -------------------------------------------------------------------------------------
// A row holds ~60 "%.4e" fields; through snprintf they cost more than everything else the drop-in does on the
// host after the GPU is done.  fmt_e4 / fmt_f4 produce the very bytes glibc's printf produces (exact decimal value
// of the double, round-half-even): the scaled value is a 128-bit integer product (relative error < 2^-63), and
// whenever its fraction lies within 2.4e-7 of one half -- where that error could matter, exact ties included --
// the field is handed to snprintf.  tests/test_capi_cpu.py compares both on millions of values.
// 10^k = m * 2^e with a 64-bit mantissa (top bit set), correctly rounded: taken from strtold, whose x87 format has just
// that mantissa.  A double's 53 bits times the mantissa is a 128-bit integer whose error against the exact scaled value
// is below 2^-63 relative: an integer part of five digits and a fraction good to ~2^-46.
struct Pow10 {
  uint64_t m[701];
  int e[701];
  Pow10() {
    char b[16];
    for (int k = 0; k <= 700; ++k) {
      snprintf(b, sizeof(b), "1e%d", k - 350);
      int ex;
      const long double f = frexpl(strtold(b, nullptr), &ex);        // in [0.5, 1)
      m[k] = (uint64_t) ldexpl(f, 64);
      e[k] = ex - 64;
    }
  }
};
const Pow10 kPow10;

inline char *put_u32(char *o, uint32_t v) {
  char t[10]; int n = 0;
  do { t[n++] = (char) ('0' + v % 10); v /= 10; } while (v);
  while (n) *o++ = t[--n];
  return o;
}

// |v| = m * 2^e2 with the top bit of m set (v finite, not zero)
inline void split_double(double v, uint64_t *m, int *e2) {
  uint64_t bits;
  memcpy(&bits, &v, 8);
  const uint64_t frac = bits & ((1ull << 52) - 1);
  const int be = (int) ((bits >> 52) & 0x7FF);
  if (be) { *m = (frac | (1ull << 52)) << 11; *e2 = be - 1075 - 11; }
  else { const int lz = __builtin_clzll(frac); *m = frac << lz; *e2 = -1074 - lz; }
}
const uint64_t kHalf = 1ull << 63, kTieBand = 1ull << 41;      // fractions within 2.4e-7 of one half go to snprintf

// "%.4e" of v followed by a tab; returns the end of the output
inline char *fmt_e4(char *o, double v) {
  if (v == 0.0 && !std::signbit(v)) { memcpy(o, "0.0000e+00\t", 11); return o + 11; }
  if (v == 1.0) { memcpy(o, "1.0000e+00\t", 11); return o + 11; }
  const double a = fabs(v);
  if (!(a > 0.0) || !(a <= DBL_MAX)) return o + snprintf(o, 40, "%.4e\t", v);      // -0, inf, nan
  uint64_t m;
  int e2;
  split_double(a, &m, &e2);
  // a lies in [2^(e2+63), 2^(e2+64)): floor(log10 a) is this estimate or one more
  int e = (int) (((int64_t) (e2 + 63) * 1292913986ll) >> 32);                     // log10(2) * 2^32
  uint64_t fl = 0, fr = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    const int k = 4 - e + 350;
    if (k < 0 || k > 700) return o + snprintf(o, 40, "%.4e\t", v);
    const unsigned __int128 P = (unsigned __int128) m * kPow10.m[k];
    const int s = -(e2 + kPow10.e[k]);                                            // a * 10^(4-e) = P / 2^s
    if (s <= 64 || s >= 128) return o + snprintf(o, 40, "%.4e\t", v);
    fl = (uint64_t) (P >> s);
    fr = (uint64_t) ((P << (128 - s)) >> 64);
    if (fl < 100000u) break;
    ++e;
  }
  if (fl < 10000u || fl >= 100000u) return o + snprintf(o, 40, "%.4e\t", v);
  if ((fr > kHalf ? fr - kHalf : kHalf - fr) < kTieBand) return o + snprintf(o, 40, "%.4e\t", v);
  uint32_t d = (uint32_t) fl + (fr > kHalf ? 1u : 0u);
  if (d >= 100000u) { d = 10000u; ++e; }                         // 9.99996 -> 1.0000e+01
  if (std::signbit(v)) *o++ = '-';
  o[0] = (char) ('0' + d / 10000u); o[1] = '.';
  o[2] = (char) ('0' + d / 1000u % 10u); o[3] = (char) ('0' + d / 100u % 10u);
  o[4] = (char) ('0' + d / 10u % 10u); o[5] = (char) ('0' + d % 10u);
  o[6] = 'e'; o[7] = e < 0 ? '-' : '+';
  const uint32_t ae = (uint32_t) (e < 0 ? -e : e);
  o += 8;
  if (ae < 10) { *o++ = '0'; *o++ = (char) ('0' + ae); } else o = put_u32(o, ae);
  *o++ = '\t';
  return o;
}

// "%.4f" of v (a float widened to double) followed by a tab
inline char *fmt_f4(char *o, double v) {
  const double a = fabs(v);
  if (!(a < 4.0e5)) return o + snprintf(o, 64, "%.4f\t", v);     // large, inf, nan
  uint64_t d = 0;
  if (a > 0.0) {
    uint64_t m;
    int e2;
    split_double(a, &m, &e2);
    const unsigned __int128 P = (unsigned __int128) m * (10000ull << 50);          // 10^4, exactly
    const int s = -(e2 - 50);                                                      // a * 10^4 = P / 2^s, s >= 96 here
    if (s < 65) return o + snprintf(o, 64, "%.4f\t", v);
    if (s <= 128) {
      const uint64_t fl = s == 128 ? 0 : (uint64_t) (P >> s);
      const uint64_t fr = s == 128 ? (uint64_t) (P >> 64) : (uint64_t) ((P << (128 - s)) >> 64);
      if ((fr > kHalf ? fr - kHalf : kHalf - fr) < kTieBand) return o + snprintf(o, 64, "%.4f\t", v);
      d = fl + (fr > kHalf ? 1u : 0u);
    }                                                                              // else below 0.00005 by far: 0.0000
  }
  if (std::signbit(v)) *o++ = '-';
  o = put_u32(o, (uint32_t) (d / 10000u));
  const uint32_t f = (uint32_t) (d % 10000u);
  o[0] = '.'; o[1] = (char) ('0' + f / 1000u); o[2] = (char) ('0' + f / 100u % 10u);
  o[3] = (char) ('0' + f / 10u % 10u); o[4] = (char) ('0' + f % 10u); o[5] = '\t';
  return o + 6;
}

struct RowBuf {
  std::string &s;
  char tmp[96];
  explicit RowBuf(std::string &str) : s(str) {}
  void e4(double v) { s.append(tmp, (size_t) (fmt_e4(tmp, v) - tmp)); }
  void f4(double v) { s.append(tmp, (size_t) (fmt_f4(tmp, v) - tmp)); }
  void u(unsigned v, char sep = '\t') { char *e = put_u32(tmp, v); *e++ = sep; s.append(tmp, (size_t) (e - tmp)); }
  void hu(unsigned short v) { u(v); }
  void lit(const char *t) { s.append(t); }
  void ch(char c) { s.push_back(c); s.push_back('\t'); }
};

void format_indel(RowBuf &o, const gt16_indel &d) {
  if (d.type == '*') { o.lit("*\t*\t*\t*\t"); return; }
  o.s.append(o.tmp, (size_t) snprintf(o.tmp, sizeof(o.tmp), "%u%c\t", (unsigned) d.len, (char) d.type));
  for (unsigned i = 0; i < d.len; ++i) o.s.push_back(kSym[(d.seq[i >> 2] >> ((i & 3) << 1)) & 3]);
  o.s.push_back('\t');
  o.u(d.hq); o.u(d.lq);
}

void format_row(std::string &out, const gt16_index *ix, const gt16_called_site &s, float lq, float rq) {
  RowBuf o(out);
  uint32_t chrPos = 0;
  const char *chr = gt16_index_translate(ix, s.amb, &chrPos);
  out.append(chr); out.push_back('\t');
  o.u(chrPos);
  o.ch(kSym[s.ref]);
  if (s.is_indel) o.ch(kSym[s.ref]); else o.lit("*\t");
  for (int g = 0; g < 16; ++g) o.e4(s.flh[g]);
  out.append(kGenotypeString[s.genotype <= 16 ? s.genotype : 16]); out.push_back('\t');
  o.e4(s.bestD); o.e4(s.secondD); o.e4(s.pD);
  o.e4(0.0); o.lit("0\t");                                            // rfProb, rfMode (SNP_Caller.cpp:583-584)
  format_indel(o, s.indel[0]);
  format_indel(o, s.indel[1]);
  o.u(s.W[0] + s.W[1] + s.W[2] + s.W[3]);
  for (int b = 0; b < 4; ++b) o.hu((unsigned short) s.W[b]);
  for (int b = 0; b < 4; ++b) { o.hu((unsigned short) s.F[b]); o.hu((unsigned short) s.R[b]); }
  for (int g = 0; g < 16; ++g) o.e4(s.lh[g]);
  if (s.is_indel) {
    for (int k3 = 0; k3 < 3; ++k3) for (int b = 0; b < 4; ++b) o.e4(s.sb[k3][b]);
    for (int k3 = 0; k3 < 3; ++k3) for (int b = 0; b < 4; ++b) o.e4(s.bqb[k3][b]);
  } else {
    for (int b = 0; b < 4; ++b) o.e4(s.sb[0][b]);
    o.lit("*\t*\t*\t*\t*\t*\t*\t*\t");
    for (int b = 0; b < 4; ++b) o.e4(s.bqb[0][b]);
    o.lit("*\t*\t*\t*\t*\t*\t*\t*\t");
  }
  o.f4(s.rpb); o.f4(s.lgc); o.f4(s.rgc);
  o.u(s.lg); o.u(s.rg);
  o.f4(s.asc); o.f4(lq); o.f4(rq);
  o.hu(s.polyrun);
  o.u(s.total_hq); o.u(s.total_lq, '\n');
}
}  // namespace

// Rows are formatted by worker threads into per-range strings and written in order: the only
// row-to-row dependency, the surviving LQ/RQ values, is resolved by a sequential pre-pass.
int gt16_write_txt(const char *path, const gt16_index *ix, const gt16_called_site *sites, uint64_t n, int append,
                   float *carry) {
  FILE *f = fopen(path, append ? "a" : "w");
  if (!f) return GT16_EIO;
  if (!append) {
    fputs("Chr\tPos\tRef\tPreBase\t"
          "pAA\tpAC\tpAG\tpAT\tpCC\tpCG\tpCT\tpGG\tpGT\tpTT\tpAX\tpCX\tpGX\tpTX\tpXX\tpXY\t"
          "GT\tOptb\tSuboptb\tpD\t"
          "rfProb\trfMode\t"
          "optIn/Del\tPatt\tHQ\tLQ\t"
          "suboptIn/Del\tPatt\tHQ\tLQ\t"
          "Depth\tA\tC\tG\tT\tA+\tA-\tC+\tC-\tG+\tG-\tT+\tT-\t"
          "AA\tAC\tAG\tAT\tCC\tCG\tCT\tGG\tGT\tTT\tAX\tCX\tGX\tTX\tXX\tXY\t"
          "SB-A\tSB-C\tSB-G\tSB-T\t"
          "SB1-A\tSB1-C\tSB1-G\tSB1-T\t"
          "SB2-A\tSB2-C\tSB2-G\tSB2-T\t"
          "BQB-A\tBQB-C\tBQB-G\tBQB-T\t"
          "BQB1-A\tBQB1-C\tBQB1-G\tBQB1-T\t"
          "BQB2-A\tBQB2-C\tBQB2-G\tBQB2-T\t"
          "RPB\t"
          "LGC\tRGC\t"
          "LG\tRG\t"
          "ASC\t"
          "LQ\tRQ\t"
          "polyrun\t"
          "totalHQ\ttotalLQ\n", f);
  }
  float lq = carry ? carry[0] : 0.0f, rq = carry ? carry[1] : 0.0f;
  // Blocks of 1,024 rows (~0.7 MB of text) are formatted by worker threads into buffers out of a small pool and written
  // by this thread in order while later blocks are being formatted: the text stays in cache-sized buffers that are used
  // over and over (a string per worker holding its whole share page-faulted 0.7 GB per million rows, with all threads
  // queueing for the address space).
  const uint64_t kBlockRows = 1024;
  const uint64_t nBlocks = (n + kBlockRows - 1) / kBlockRows;
  const unsigned nt = (unsigned) std::max<uint64_t>(1, std::min<uint64_t>(io_threads(), nBlocks));
  // values in force at the start of each block
  std::vector<float> lq0(nBlocks, 0.f), rq0(nBlocks, 0.f);
  for (uint64_t blk = 0; blk < nBlocks; ++blk) {
    lq0[blk] = lq; rq0[blk] = rq;
    const uint64_t a = blk * kBlockRows, b = std::min<uint64_t>(n, a + kBlockRows);
    for (uint64_t k = a; k < b; ++k) if (sites[k].rq_assigned) { lq = sites[k].lq; rq = sites[k].rq; }
  }
  bool bad = false;
  auto format_block = [&](uint64_t blk, std::string &out) {
    const uint64_t a = blk * kBlockRows, b = std::min<uint64_t>(n, a + kBlockRows);
    float l = lq0[blk], r = rq0[blk];
    for (uint64_t k = a; k < b; ++k) {
      if (sites[k].rq_assigned) { l = sites[k].lq; r = sites[k].rq; }   // else: previous row's values survive
      format_row(out, ix, sites[k], l, r);
    }
  };
  if (nt == 1) {
    std::string out;
    out.reserve((size_t) std::min<uint64_t>(n, kBlockRows) * 760);
    for (uint64_t blk = 0; blk < nBlocks; ++blk) {
      out.clear();
      format_block(blk, out);
      if (!out.empty() && fwrite(out.data(), 1, out.size(), f) != out.size()) bad = true;
    }
  } else {
    const uint64_t window = 4ull * nt;                  // blocks formatted ahead of the one being written
    std::mutex mu;
    std::condition_variable cv;
    std::vector<std::string> done(nBlocks);
    std::vector<char> ready(nBlocks, 0);
    std::vector<std::string> pool;
    uint64_t next = 0, written = 0;
    auto work = [&]() {
      std::string out;
      for (;;) {
        uint64_t blk;
        {
          std::unique_lock<std::mutex> lk(mu);
          blk = next++;
          if (blk >= nBlocks) return;
          cv.wait(lk, [&] { return blk < written + window; });
          if (!pool.empty()) { out.swap(pool.back()); pool.pop_back(); }
        }
        out.clear();
        if (out.capacity() < kBlockRows * 760) out.reserve(kBlockRows * 760);
        format_block(blk, out);
        {
          std::lock_guard<std::mutex> lk(mu);
          done[blk].swap(out);
          ready[blk] = 1;
        }
        cv.notify_all();
      }
    };
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(work);
    std::string cur;
    for (uint64_t blk = 0; blk < nBlocks; ++blk) {
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return ready[blk] != 0; });
        cur.swap(done[blk]);
      }
      if (!cur.empty() && fwrite(cur.data(), 1, cur.size(), f) != cur.size()) bad = true;
      {
        std::lock_guard<std::mutex> lk(mu);
        pool.emplace_back();
        pool.back().swap(cur);
        ++written;
      }
      cv.notify_all();
    }
    for (auto &x : th) x.join();
  }
  if (carry) { carry[0] = lq; carry[1] = rq; }
  if (ferror(f)) bad = true;
  if (fclose(f) != 0 || bad) return GT16_EIO;
  return GT16_OK;
}

// test hook: the two fast formatters (tests compare them with printf on millions of values)
int gt16_format_number(double v, int kind, char *out) {
  char *e = kind == 0 ? fmt_e4(out, v) : fmt_f4(out, v);
  e[-1] = 0;                                   // drop the tab
  return (int) (e - 1 - out);
}

}  // extern "C"
