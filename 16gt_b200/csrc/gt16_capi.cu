// gt16_capi.cu -- extern "C" layer of libgt16.so (include/gt16.h): context, table prefill,
// the reference's operator seam, and the fused snapshot -> called-sites path with a copy
// stream that feeds HBM while the previous block is being classified.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <functional>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>
#include "gt16_kernels.cuh"

using namespace gt16;

static thread_local std::string g_createError;      // per thread: the drop-in creates one context per shard thread
#ifndef GT16_BIAS_CACHE_LOG2
#define GT16_BIAS_CACHE_LOG2 20
#endif
static const uint32_t kBiasCacheEntries = 1u << GT16_BIAS_CACHE_LOG2;     // 8 MB, stays in L2
// room in the list of Fisher tests the memo cache did not hold: four per record of the bound.  A context's first call finds
// the cache empty and lists every test (~8 per SNP row); the record bound is ~7x the called fraction of a 30x genome, so
// that fits; whatever does not is walked in place by the record kernel.
static const uint32_t kMissedPerRecord = 4;

namespace {

// grow-only device buffer: the fused path allocates nothing in steady state
struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// page-locked result buffers are recycled between calls (cudaHostAlloc of tens of MB costs milliseconds);
// gt16_calls objects hold a reference, so they may outlive the context
struct PinPool {
  std::mutex mu;
  std::vector<std::pair<void *, size_t> > free_;
  void *take(size_t bytes, size_t *cap) {
    {
      std::lock_guard<std::mutex> g(mu);
      for (size_t k = 0; k < free_.size(); ++k)
        if (free_[k].second >= bytes) { void *p = free_[k].first; *cap = free_[k].second; free_.erase(free_.begin() + k); return p; }
    }
    size_t want = bytes + bytes / 4 + 4096;
    void *p = nullptr;
    if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) { (void) cudaGetLastError(); return nullptr; }
    *cap = want;
    return p;
  }
  void give(void *p, size_t cap) {
    std::lock_guard<std::mutex> g(mu);
    if (free_.size() >= 4) { cudaFreeHost(p); return; }
    free_.push_back(std::make_pair(p, cap));
  }
  ~PinPool() { for (auto &f : free_) cudaFreeHost(f.first); }
};

}  // namespace

struct gt16_ctx {
  int device = 0;
  cudaStream_t compute = nullptr, aux = nullptr, copy = nullptr, d2h = nullptr;   // kernels / Fisher-test kernels / host->device / device->host
  std::string err;
  // tables
  double *tabOwned[10] = {nullptr};
  int nOwned = 0;
  double *tab[10] = {nullptr};
  double *ycsq = nullptr, *prior = nullptr;
  uint32_t *g2i = nullptr;
  float *lnfact = nullptr;
  uint32_t *s1 = nullptr, *s2p = nullptr, *s2q = nullptr, *g4 = nullptr;
  uint32_t *errWord = nullptr, *dTotal = nullptr, *dFlagged = nullptr, *dMissed = nullptr;   // one allocation: error bits, compaction total, sites evaluated exactly, Fisher tests left to the walk kernel, ...
  uint32_t *dCounts = nullptr;          // ... and per segment a pair: [2 seg] side-buffer entries the evaluator reserved (its device-side counter, read by the
                                        // kernels behind it; includes entries reserved and left unused), [2 seg + 1] the number of candidates (the scan's total)
  static const int kMaxSegs = 512;
  uint32_t *hPinnedRing = nullptr;      // kCallRing slots of 8 words
  uint32_t *hPinned = nullptr;          // the current call's slot: [0] entries reserved / [1] records of a device-resident call, [2] error word, [3] -, [4] sites evaluated exactly
  uint32_t *hCounts = nullptr;          // page-locked copy of dCounts (host-result calls)
  uint32_t recCap = 0;                  // records the candidate buffers (part, rankOf, list, recs, recs2) hold per segment
  gt16_called_site *extDst = nullptr;   // gt16_call_sites_host_to: the caller's result buffer for the call in progress
  uint64_t extCap = 0;
  uint32_t recCapWanted = 0;            // set when a call found more candidates than recCap: the retry allocates this many
  bool havePrefill = false, haveDecisions = false, useDecisions = true;
  double balance = 0.01, unbalance = 0.005;
  float prefillMs = 0.f, decisionMs = 0.f;
  uint32_t launches = 0;
  uint32_t shortcutAlt = 0;               // stream_shortcut_bound() of the current decision bitmaps
  uint2 *biasCache = nullptr;             // memoised Fisher tests (k_bias), kBiasCacheEntries entries
  int numSMs = 148;
  // persistent scratch of the fused path
  DevBuf mask, counts, part, rankOf, list, missed, recs, recs2, planStart, planFirst, planTile;
  std::shared_ptr<const void> planCache; // host side of what planStart / planFirst / planTile hold (PlanData, make_plan)
  DevBuf sCounters, sInvalid, sOvf, sOvfBase, sPool, sRef;      // streamed (host-fed) snapshot
  DevBuf sSparse;                                               // records of a sparse view on their way into the counter array
  cudaEvent_t sparseCopied = nullptr;                           // ... and the event that orders a chunk's scatter behind its copy
  // events and the small pinned read-back slot of a call come from ring entry evCur: a call that returns before its tail
  // has run (gt16_call_sites_device_nowait) keeps its entry until kCallRing further calls have been made
  static const int kCallRing = 128;
  std::vector<cudaEvent_t> evPool[kCallRing];
  int evCur = 0;
  std::shared_ptr<PinPool> pins = std::make_shared<PinPool>();
  DevBuf opIn, opOut;                     // operator seam: device scratch, ...
  void *opPin = nullptr;                  // ... its two page-locked slots ...
  size_t opPinCap = 0;
  cudaEvent_t opEv[2] = {nullptr, nullptr};   // ... and the events that say a slot's chunk has been copied
  void *stagePin = nullptr;               // page-locked ring for callers whose counters live in pageable memory
  cudaEvent_t stageEv[32] = {nullptr};    // slot k's last copy has left the ring
};

struct gt16_dev_snapshot {
  SnapView v;
  // `counters` is indexed by genome position but only [winLo, winHi) (+ padding) is backed by memory: the allocation
  // starts at countersAlloc = counters + winLo (a whole snapshot has the window [0, L))
  void *counters = nullptr, *countersAlloc = nullptr, *invalid = nullptr, *ovf = nullptr, *ovfBase = nullptr, *pool = nullptr, *ref = nullptr;
  uint64_t winLo = 0, winHi = 0;
  bool borrowed = false;      // buffers belong to the context's streaming arena
};

struct gt16_dev_calls {          // records left in the context's record buffer
  const gt16_called_site *recs = nullptr;
  uint64_t n = 0;                 // known once the call has completed (gt16_dev_calls_wait)
  gt16_ctx *ctx = nullptr;
  // what the call was, in case it has to be repeated with larger record buffers
  const gt16_dev_snapshot *snap = nullptr;
  std::vector<gt16_region> regions;
  uint64_t begin = 0, end = 0;
  uint32_t cap = 0;
  // completion of a call that returned before its tail had run (gt16_dev_calls_wait)
  bool pending = false, useTables = false;
  cudaEvent_t evBegin = nullptr, evEnd = nullptr;
  std::vector<cudaEvent_t> timing;
  const uint32_t *slot = nullptr;         // the call's pinned read-back slot
  gt16_stats st;
};

struct gt16_calls {
  std::shared_ptr<PinPool> pins;      // null: `sites` is the caller's own buffer (gt16_call_sites_host_to)
  gt16_called_site *sites = nullptr;
  size_t cap = 0;                     // bytes
  uint64_t n = 0;
  bool tooSmall = false;              // caller's buffer could not take every record
  ~gt16_calls() { if (sites && pins) pins->give(sites, cap); }
};

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char b_[512];                                                                                \
      snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      ctx->err = b_;                                                                               \
      return GT16_ECUDA;                                                                           \
    }                                                                                              \
  } while (0)

static int fail(gt16_ctx *ctx, int code, const char *msg) { if (ctx) ctx->err = msg; return code; }
// for gt16_pileup.cu (the producer lives in its own translation unit)
int gt16_pileup_set_error(gt16_ctx *ctx, int code, const char *msg) { return fail(ctx, code, msg); }
int gt16_pileup_device(gt16_ctx *ctx) {
  if (cudaSetDevice(ctx->device) != cudaSuccess) { (void) cudaGetLastError(); return fail(ctx, GT16_ECUDA, "cudaSetDevice failed"); }
  return GT16_OK;
}

static EvalParams eval_params(const gt16_ctx *ctx, double balance, double unbalance) {
  EvalParams P;
  for (int k = 0; k < 10; ++k) P.lh.t[k] = ctx->tab[k];
  P.lh.ycsq = ctx->ycsq;
  P.lh.p_err_s = balance / 4;             // FisherExactTest.cpp:298
  P.lh.p_err_i = unbalance / 2;           // :299
  P.prior = ctx->prior;
  P.g2i = ctx->g2i;
  P.lnfact.tab = ctx->lnfact;
  P.lnfact.size = kLnFactTab;
  P.lnfact.far = nullptr;
  P.lnfact.farSize = 0;
  return P;
}
static EvalParams eval_params(const gt16_ctx *ctx) { return eval_params(ctx, ctx->balance, ctx->unbalance); }

// scratch of the operator seam: freed on every return path
namespace {
struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  cudaError_t create() { cudaError_t e = cudaEventCreate(&a); return e != cudaSuccess ? e : cudaEventCreate(&b); }
};
}  // namespace

// defaults of selectPossibleSNPs (VariantCaller.cpp:506-513)
static cudaError_t install_default_tables(gt16_ctx *ctx) {
  const unsigned g2i[16] = {0, 1, 2, 3, 9, 10, 11, 18, 19, 27, 4, 12, 20, 28, 36, 37};
  double prior[256];
  make_prior(prior);
  cudaError_t e = cudaMemcpy(ctx->g2i, g2i, sizeof(g2i), cudaMemcpyHostToDevice);
  if (e != cudaSuccess) return e;
  return cudaMemcpy(ctx->prior, prior, sizeof(prior), cudaMemcpyHostToDevice);
}

extern "C" {

const char *gt16_last_error(const gt16_ctx *ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

int gt16_create_snp_prior(double out256[256]) {
  if (!out256) return GT16_EINVAL;
  make_prior(out256);
  return GT16_OK;
}

int gt16_device_count(int *n) {
  if (!n) return GT16_EINVAL;
  *n = 0;
  if (cudaGetDeviceCount(n) != cudaSuccess) { (void) cudaGetLastError(); *n = 0; return GT16_ENODEVICE; }
  return GT16_OK;
}

int gt16_ctx_create(int device, gt16_ctx **out) {
  if (!out) return GT16_EINVAL;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    g_createError = std::string("no CUDA device: ") + cudaGetErrorString(e) + " (libgt16 has no CPU path)";
    (void) cudaGetLastError();
    return GT16_ENODEVICE;
  }
  if (device < 0 || device >= n) { g_createError = "device index out of range"; return GT16_EINVAL; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
    g_createError = "device is not Blackwell-class (kernels are built for sm_100a only)";
    return GT16_ENODEVICE;
  }
  gt16_ctx *ctx = new gt16_ctx();
  ctx->device = device;
  ctx->numSMs = prop.multiProcessorCount;
  auto bail = [&](const char *what, cudaError_t ce) {
    g_createError = std::string(what) + ": " + cudaGetErrorString(ce);
    (void) cudaGetLastError();
    gt16_ctx_destroy(ctx);                 // streams, device and page-locked allocations made so far
    return GT16_ECUDA;
  };
  if ((e = cudaSetDevice(device)) != cudaSuccess) return bail("cudaSetDevice", e);
  if ((e = configure_kernels()) != cudaSuccess) return bail("cudaFuncSetAttribute", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking)) != cudaSuccess) return bail("stream", e);
  if ((e = cudaEventCreateWithFlags(&ctx->sparseCopied, cudaEventDisableTiming)) != cudaSuccess) return bail("event", e);
  if ((e = cudaMalloc(&ctx->prior, 256 * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&ctx->g2i, 16 * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&ctx->ycsq, kYcsqN * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&ctx->lnfact, (size_t) kLnFactTab * sizeof(float))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMalloc(&ctx->errWord, (4 + 2 * gt16_ctx::kMaxSegs) * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMalloc", e);
  ctx->dTotal = ctx->errWord + 1;
  ctx->dFlagged = ctx->errWord + 2;
  ctx->dMissed = ctx->errWord + 3;
  ctx->dCounts = ctx->errWord + 4;       // 8-byte aligned: a segment's pair is read back with one copy
  if ((e = cudaHostAlloc(&ctx->hCounts, 2 * gt16_ctx::kMaxSegs * sizeof(uint32_t), cudaHostAllocDefault)) != cudaSuccess) return bail("cudaHostAlloc", e);
  if ((e = cudaHostAlloc(&ctx->hPinnedRing, gt16_ctx::kCallRing * 8 * sizeof(uint32_t), cudaHostAllocDefault)) != cudaSuccess) return bail("cudaHostAlloc", e);
  ctx->hPinned = ctx->hPinnedRing;
  if ((e = cudaMemset(ctx->errWord, 0, (4 + 2 * gt16_ctx::kMaxSegs) * sizeof(uint32_t))) != cudaSuccess) return bail("cudaMemset", e);
  if ((e = cudaMalloc(&ctx->biasCache, (size_t) kBiasCacheEntries * sizeof(uint2))) != cudaSuccess) return bail("cudaMalloc", e);
  if ((e = cudaMemset(ctx->biasCache, 0xFF, (size_t) kBiasCacheEntries * sizeof(uint2))) != cudaSuccess) return bail("cudaMemset", e);
  if ((e = install_default_tables(ctx)) != cudaSuccess) return bail("cudaMemcpy", e);
  // float log-factorial table: lnfactC(n) = lgammaf((float)(n+1)) from the host libm, i.e. the
  // very function the reference calls (SNPFunctions.cpp:82-84) -- a constant table, like the prior
  {
    std::vector<float> tab(kLnFactTab);
    unsigned nt = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t)
      th.emplace_back([&, t]() { for (uint32_t k = t; k < kLnFactTab; k += nt) tab[k] = lgammaf((float) (k + 1)); });
    for (auto &x : th) x.join();
    if ((e = cudaMemcpy(ctx->lnfact, tab.data(), tab.size() * sizeof(float), cudaMemcpyHostToDevice)) != cudaSuccess)
      return bail("cudaMemcpy", e);
  }
  *out = ctx;
  return GT16_OK;
}

void gt16_ctx_destroy(gt16_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int k = 0; k < ctx->nOwned; ++k) cudaFree(ctx->tabOwned[k]);
  cudaFree(ctx->ycsq); cudaFree(ctx->prior); cudaFree(ctx->g2i); cudaFree(ctx->lnfact);
  cudaFree(ctx->s1); cudaFree(ctx->s2p); cudaFree(ctx->s2q); cudaFree(ctx->g4);
  cudaFree(ctx->errWord);
  cudaFree(ctx->biasCache);
  if (ctx->hPinnedRing) cudaFreeHost(ctx->hPinnedRing);
  if (ctx->hCounts) cudaFreeHost(ctx->hCounts);
  for (DevBuf *b : {&ctx->mask, &ctx->counts, &ctx->part, &ctx->rankOf, &ctx->list, &ctx->missed, &ctx->recs, &ctx->recs2, &ctx->planStart, &ctx->planFirst, &ctx->planTile,
                    &ctx->sCounters, &ctx->sInvalid, &ctx->sOvf, &ctx->sOvfBase, &ctx->sPool, &ctx->sRef, &ctx->sSparse}) b->release();
  for (auto &pool : ctx->evPool) for (auto e : pool) cudaEventDestroy(e);
  if (ctx->stagePin) cudaFreeHost(ctx->stagePin);
  ctx->opIn.release(); ctx->opOut.release();
  if (ctx->opPin) cudaFreeHost(ctx->opPin);
  for (auto e : ctx->opEv) if (e) cudaEventDestroy(e);
  for (auto e : ctx->stageEv) if (e) cudaEventDestroy(e);
  for (cudaStream_t st : {ctx->compute, ctx->copy, ctx->d2h, ctx->aux}) if (st) cudaStreamDestroy(st);
  if (ctx->sparseCopied) cudaEventDestroy(ctx->sparseCopied);
  (void) cudaGetLastError();
  delete ctx;
}

int gt16_copy_genotype_table(gt16_ctx *ctx, const unsigned int g2i[16]) {
  if (!ctx || !g2i) return GT16_EINVAL;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpy(ctx->g2i, g2i, 16 * sizeof(unsigned), cudaMemcpyHostToDevice));
  ctx->haveDecisions = false;             // decision bitmaps were built for the previous table
  return GT16_OK;
}

int gt16_copy_snp_prior(gt16_ctx *ctx, const double prior[256]) {
  if (!ctx || !prior) return GT16_EINVAL;
  CU(cudaSetDevice(ctx->device));
  CU(cudaMemcpy(ctx->prior, prior, 256 * sizeof(double), cudaMemcpyHostToDevice));
  ctx->haveDecisions = false;
  return GT16_OK;
}

int gt16_prefill_likelihood_cache(gt16_ctx *ctx, float b_err, float ub_err, void *lc_host) {
  if (!ctx) return GT16_EINVAL;
  CU(cudaSetDevice(ctx->device));
  ErrFractions ef = err_fractions(b_err, ub_err);
  for (int k = 0; k < ctx->nOwned; ++k) { cudaFree(ctx->tabOwned[k]); ctx->tabOwned[k] = nullptr; }
  ctx->nOwned = 0;
  ctx->havePrefill = false; ctx->haveDecisions = false;
  EventPair ev;
  CU(ev.create());
  const cudaEvent_t e0 = ev.a, e1 = ev.b;
  CU(cudaEventRecord(e0, ctx->compute));
  const size_t bytes = (size_t) kFisherN * kFisherN * sizeof(double);
  for (int k = 0; k < 10; ++k) {
    int same = -1;
    for (int j = 0; j < k; ++j) if (memcmp(&ef.e[j], &ef.e[k], sizeof(float)) == 0) { same = j; break; }
    if (same >= 0) { ctx->tab[k] = ctx->tab[same]; continue; }   // identical error fraction -> identical table
    double *d = nullptr;
    CU(cudaMalloc(&d, bytes));
    ctx->tabOwned[ctx->nOwned++] = d;
    ctx->tab[k] = d;
    launch_build_table(d, ef.e[k], ctx->compute);
    ctx->launches++;
  }
  launch_build_ycsq(ctx->ycsq, ctx->compute);
  ctx->launches++;
  CU(cudaEventRecord(e1, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  CU(cudaGetLastError());
  CU(cudaEventElapsedTime(&ctx->prefillMs, e0, e1));
  ctx->havePrefill = true;
  if (lc_host) {                                  // LikelihoodCache layout, likelihood_cache.h:13-31
    char *dst = (char *) lc_host;
    for (int k = 0; k < 10; ++k) CU(cudaMemcpy(dst + (size_t) k * bytes, ctx->tab[k], bytes, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(dst + 10 * bytes, ctx->ycsq, kYcsqN * sizeof(double), cudaMemcpyDeviceToHost));
  }
  return GT16_OK;
}

int gt16_configure(gt16_ctx *ctx, double balance_sub_error, double unbalance_sub_error, int use_decision_tables) {
  if (!ctx) return GT16_EINVAL;
  ctx->balance = balance_sub_error;
  ctx->unbalance = unbalance_sub_error;
  ctx->useDecisions = use_decision_tables != 0;
  // start-up state of selectPossibleSNPs (VariantCaller.cpp:506-513): the default genotype2int and prior, whatever an
  // earlier gt16_copy_* installed.  The decision bitmaps below are built from them, and key_needs_exact's labelling
  // argument (one representative reference base) needs the role symmetry of the default prior; a caller that installs
  // its own tables AFTER configure gets the exact-inline path (gt16_copy_* clear haveDecisions).
  CU(cudaSetDevice(ctx->device));
  CU(install_default_tables(ctx));
  // snapshotSnpcaller.cpp:230-231 passes the ini doubles through float parameters
  int rc = gt16_prefill_likelihood_cache(ctx, (float) balance_sub_error, (float) unbalance_sub_error, nullptr);
  if (rc != GT16_OK) return rc;
  if (!ctx->useDecisions) return GT16_OK;
  if (!ctx->s1) {
    CU(cudaMalloc(&ctx->s1, kS1Words * 4)); CU(cudaMalloc(&ctx->s2p, kS2Words * 4));
    CU(cudaMalloc(&ctx->s2q, kS2Words * 4)); CU(cudaMalloc(&ctx->g4, kG4Words * 4));
  }
  EventPair ev;
  CU(ev.create());
  const cudaEvent_t e0 = ev.a, e1 = ev.b;
  CU(cudaEventRecord(e0, ctx->compute));
  launch_build_decisions(eval_params(ctx), ctx->s1, ctx->s2p, ctx->s2q, ctx->g4, ctx->compute);
  ctx->launches++;
  CU(cudaEventRecord(e1, ctx->compute));
  CU(cudaStreamSynchronize(ctx->compute));
  CU(cudaGetLastError());
  CU(cudaEventElapsedTime(&ctx->decisionMs, e0, e1));
  {
    std::vector<uint32_t> h1(kS1Words), hp(kS2Words), hq(kS2Words);
    CU(cudaMemcpy(h1.data(), ctx->s1, kS1Words * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hp.data(), ctx->s2p, kS2Words * 4, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(hq.data(), ctx->s2q, kS2Words * 4, cudaMemcpyDeviceToHost));
    ctx->shortcutAlt = stream_shortcut_bound(h1.data(), hp.data(), hq.data());
    if (getenv("GT16_NO_SHORTCUT")) ctx->shortcutAlt = 0;
  }
  ctx->haveDecisions = true;
  return GT16_OK;
}

int gt16_setup_times(const gt16_ctx *ctx, float *prefill_ms, float *decision_ms) {
  if (!ctx) return GT16_EINVAL;
  if (prefill_ms) *prefill_ms = ctx->prefillMs;
  if (decision_ms) *decision_ms = ctx->decisionMs;
  return GT16_OK;
}

// ---- operator seam ------------------------------------------------------------------------
// The reference calls these once per batch of up to 8,388,608 sites with its array-of-structures buffers
// (VariantCaller.cpp:620-707).  The sites go through in chunks: worker threads gather a chunk's fields out of the
// caller's strides into one of two page-locked slots while the previous chunk is copied and computed; device and
// page-locked scratch belong to the context and only grow, so a loop over batches allocates once.
}  // extern "C"
static uint32_t op_chunk_sites() {                        // GT16_OP_CHUNK: test hook (a small chunk exercises the ring)
  uint32_t n = 1u << 18;
  if (const char *e = getenv("GT16_OP_CHUNK")) { long v = atol(e); if (v >= 1 && v <= (1l << 22)) n = (uint32_t) v; }
  return n;
}
static cudaError_t op_prepare(gt16_ctx *ctx, size_t slotBytes, size_t devIn, size_t devOut) {
  if (2 * slotBytes > ctx->opPinCap) {
    if (ctx->opPin) cudaFreeHost(ctx->opPin);
    ctx->opPin = nullptr; ctx->opPinCap = 0;
    cudaError_t e = cudaHostAlloc(&ctx->opPin, 2 * slotBytes, cudaHostAllocDefault);
    if (e != cudaSuccess) return e;
    ctx->opPinCap = 2 * slotBytes;
  }
  for (int k = 0; k < 2; ++k)
    if (!ctx->opEv[k]) { cudaError_t e = cudaEventCreateWithFlags(&ctx->opEv[k], cudaEventDisableTiming); if (e != cudaSuccess) return e; }
  cudaError_t e = ctx->opIn.ensure(devIn);
  return e != cudaSuccess ? e : ctx->opOut.ensure(devOut);
}
template <class F>
static void op_parallel(uint32_t n, F f) {                // f(lo, hi) over [0, n) on a few threads
  unsigned nt = std::thread::hardware_concurrency();
  nt = std::max(1u, std::min(8u, nt ? nt : 4u));
  if (n < (1u << 15)) nt = 1;
  if (nt == 1) { f(0u, n); return; }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) {
    const uint32_t lo = (uint32_t) ((uint64_t) n * t / nt), hi = (uint32_t) ((uint64_t) n * (t + 1) / nt);
    th.emplace_back([=]() { f(lo, hi); });
  }
  for (auto &x : th) x.join();
}
extern "C" {

int gt16_compute_genotype_likelihood(gt16_ctx *ctx, const void *w, size_t in_stride, const void *ref_char,
                                     size_t ref_stride, double *lh, uint32_t n, double balance, double unbalance) {
  if (!ctx || !w || !ref_char || !lh || in_stride < 24 || ref_stride < 1) return GT16_EINVAL;
  if (!ctx->havePrefill) return fail(ctx, GT16_ESTATE, "likelihood cache not prefilled");
  if (!n) return GT16_OK;
  CU(cudaSetDevice(ctx->device));
  // the call's own error rates go straight into the kernel parameters: the context's (which the fused path and its
  // decision bitmaps were configured with) are not touched
  const uint32_t chunk = std::min(op_chunk_sites(), n);
  const size_t slotBytes = (size_t) chunk * 25;           // only W[0..5] and refChar are read by the operator
  CU(op_prepare(ctx, slotBytes, slotBytes, (size_t) chunk * 128));
  uint8_t *dIn = (uint8_t *) ctx->opIn.p;
  double *dL = (double *) ctx->opOut.p;
  const EvalParams P = eval_params(ctx, balance, unbalance);
  // chunk k + 1 is gathered while chunk k is on the device; a copy into pageable `lh` returns when it is done
  auto gather = [&](uint32_t k) -> cudaError_t {
    const uint32_t a = k * chunk, m = std::min(chunk, n - a);
    uint8_t *pin = (uint8_t *) ctx->opPin + (size_t) (k & 1) * slotBytes;
    if (k >= 2) { cudaError_t e = cudaEventSynchronize(ctx->opEv[k & 1]); if (e != cudaSuccess) return e; }   // the slot's previous chunk has left it
    op_parallel(m, [=](uint32_t lo, uint32_t hi) {
      for (uint32_t i = lo; i < hi; ++i) {
        memcpy(pin + (size_t) i * 24, (const char *) w + (size_t) (a + i) * in_stride, 24);
        pin[(size_t) chunk * 24 + i] = *((const uint8_t *) ref_char + (size_t) (a + i) * ref_stride);
      }
    });
    return cudaSuccess;
  };
  const uint32_t nChunks = (n + chunk - 1) / chunk;
  CU(gather(0));
  for (uint32_t k = 0; k < nChunks; ++k) {
    const uint32_t a = k * chunk, m = std::min(chunk, n - a);
    const uint8_t *pin = (const uint8_t *) ctx->opPin + (size_t) (k & 1) * slotBytes;
    CU(cudaMemcpyAsync(dIn, pin, (size_t) chunk * 24 + m, cudaMemcpyHostToDevice, ctx->compute));
    CU(cudaEventRecord(ctx->opEv[k & 1], ctx->compute));
    launch_op_likelihood(P, dIn, 24, dIn + (size_t) chunk * 24, 1, dL, m, ctx->compute);
    ctx->launches++;
    if (k + 1 < nChunks) CU(gather(k + 1));
    CU(cudaMemcpyAsync(lh + (size_t) a * 16, dL, (size_t) m * 128, cudaMemcpyDeviceToHost, ctx->compute));
  }
  CU(cudaStreamSynchronize(ctx->compute));
  CU(cudaGetLastError());
  return GT16_OK;
}

int gt16_compute_genotype(gt16_ctx *ctx, const unsigned char *ref_char, const double *lh, void *info,
                          size_t info_stride, uint32_t n) {
  if (!ctx || !ref_char || !lh || !info || info_stride < 160) return GT16_EINVAL;
  if (!n) return GT16_OK;
  CU(cudaSetDevice(ctx->device));
  const size_t dStride = 160;   // pD, genotype(+pad), flh[16], bestD, secondD
  const uint32_t chunk = std::min(op_chunk_sites(), n);
  const size_t slotBytes = (size_t) chunk * dStride;
  CU(op_prepare(ctx, slotBytes, (size_t) chunk * 129, slotBytes));
  uint8_t *dR = (uint8_t *) ctx->opIn.p + (size_t) chunk * 128, *dI = (uint8_t *) ctx->opOut.p;
  double *dL = (double *) ctx->opIn.p;
  const EvalParams P = eval_params(ctx);
  // scatter into SnpCallingInfo (SNP_Meta.h:115-125); chunk k - 1 is scattered while chunk k is on the device
  auto scatter = [&](uint32_t a, uint32_t m, const uint8_t *pin) {
    op_parallel(m, [=](uint32_t lo, uint32_t hi) {
      for (uint32_t i = lo; i < hi; ++i) {
        char *d = (char *) info + (size_t) (a + i) * info_stride;
        const uint8_t *s = pin + (size_t) i * dStride;
        memcpy(d + 0, s + 0, 8);                  // pD
        memcpy(d + 8, s + 8, 4);                  // genotype (secondGenotype @12 untouched)
        memcpy(d + 16, s + 16, 128);              // flh
        memcpy(d + 144, s + 144, 16);             // bestD, secondD (rfProb/rfMode untouched)
      }
    });
  };
  const uint32_t nChunks = (n + chunk - 1) / chunk;
  uint32_t prevA = 0, prevM = 0;
  for (uint32_t k = 0; k < nChunks; ++k) {
    const uint32_t a = k * chunk, m = std::min(chunk, n - a);
    uint8_t *pin = (uint8_t *) ctx->opPin + (size_t) (k & 1) * slotBytes;
    CU(cudaMemcpyAsync(dR, ref_char + a, m, cudaMemcpyHostToDevice, ctx->compute));
    CU(cudaMemcpyAsync(dL, lh + (size_t) a * 16, (size_t) m * 128, cudaMemcpyHostToDevice, ctx->compute));
    CU(cudaMemsetAsync(dI, 0, (size_t) m * dStride, ctx->compute));
    launch_op_genotype(P, dR, dL, dI, dStride, m, ctx->compute);
    ctx->launches++;
    // slot k & 1 was scattered (chunk k - 2) before chunk k - 1 was issued: it is free
    CU(cudaMemcpyAsync(pin, dI, (size_t) m * dStride, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaEventRecord(ctx->opEv[k & 1], ctx->compute));
    if (k >= 1) {
      CU(cudaEventSynchronize(ctx->opEv[(k - 1) & 1]));
      scatter(prevA, prevM, (const uint8_t *) ctx->opPin + (size_t) ((k - 1) & 1) * slotBytes);
    }
    prevA = a; prevM = m;
  }
  CU(cudaStreamSynchronize(ctx->compute));
  CU(cudaGetLastError());
  scatter(prevA, prevM, (const uint8_t *) ctx->opPin + (size_t) ((nChunks - 1) & 1) * slotBytes);
  return GT16_OK;
}

int gt16_compute_bias(gt16_ctx *ctx, const void *w, const void *f, const void *r, size_t stride, uint32_t n,
                      float *sb, float *bqb) {
  if (!ctx || !w || !f || !r || !sb || !bqb || stride < 16) return GT16_EINVAL;
  if (!n) return GT16_OK;
  CU(cudaSetDevice(ctx->device));
  const uint32_t chunk = std::min(op_chunk_sites(), n);
  const size_t slotBytes = (size_t) chunk * 48;           // W, F, R of a site side by side
  CU(op_prepare(ctx, slotBytes, slotBytes, (size_t) chunk * 32));
  uint8_t *d = (uint8_t *) ctx->opIn.p;
  float *dS = (float *) ctx->opOut.p, *dB = dS + (size_t) chunk * 4;
  const EvalParams P = eval_params(ctx);
  auto gather = [&](uint32_t k) -> cudaError_t {
    const uint32_t a = k * chunk, m = std::min(chunk, n - a);
    uint8_t *pin = (uint8_t *) ctx->opPin + (size_t) (k & 1) * slotBytes;
    if (k >= 2) { cudaError_t e = cudaEventSynchronize(ctx->opEv[k & 1]); if (e != cudaSuccess) return e; }
    op_parallel(m, [=](uint32_t lo, uint32_t hi) {
      for (uint32_t i = lo; i < hi; ++i) {
        memcpy(pin + (size_t) i * 48 + 0, (const char *) w + (size_t) (a + i) * stride, 16);
        memcpy(pin + (size_t) i * 48 + 16, (const char *) f + (size_t) (a + i) * stride, 16);
        memcpy(pin + (size_t) i * 48 + 32, (const char *) r + (size_t) (a + i) * stride, 16);
      }
    });
    return cudaSuccess;
  };
  const uint32_t nChunks = (n + chunk - 1) / chunk;
  CU(gather(0));
  for (uint32_t k = 0; k < nChunks; ++k) {
    const uint32_t a = k * chunk, m = std::min(chunk, n - a);
    const uint8_t *pin = (const uint8_t *) ctx->opPin + (size_t) (k & 1) * slotBytes;
    CU(cudaMemcpyAsync(d, pin, (size_t) m * 48, cudaMemcpyHostToDevice, ctx->compute));
    CU(cudaEventRecord(ctx->opEv[k & 1], ctx->compute));
    launch_op_bias(P, d, d + 16, d + 32, 48, m, dS, dB, ctx->compute);
    ctx->launches++;
    if (k + 1 < nChunks) CU(gather(k + 1));
    CU(cudaMemcpyAsync(sb + (size_t) a * 4, dS, (size_t) m * 16, cudaMemcpyDeviceToHost, ctx->compute));
    CU(cudaMemcpyAsync(bqb + (size_t) a * 4, dB, (size_t) m * 16, cudaMemcpyDeviceToHost, ctx->compute));
  }
  CU(cudaStreamSynchronize(ctx->compute));
  CU(cudaGetLastError());
  return GT16_OK;
}

}  // extern "C"

// ---- fused path -----------------------------------------------------------------------------
// A context is driven by one host thread at a time; all scratch is owned by the context and only grows.
static const uint32_t kCopyChunkSites = 1u << 22;     // 32 MiB of counters per cudaMemcpyAsync (page-locked source)
static const uint32_t kStageChunkSites = 1u << 20;    // 8 MiB per chunk when the source is pageable and goes through the ring
static const int kStageThreadsMax = 16, kStageSlots = 2 * kStageThreadsMax;  // stager threads, ring slots (two per thread)
static int stage_threads() {                              // GT16_STAGE_THREADS, default 4
  int n = 4;
  if (const char *e = getenv("GT16_STAGE_THREADS")) { int v = atoi(e); if (v >= 1 && v <= kStageThreadsMax) n = v; }
  return n;
}
// Record buffers are sized by a bound, not by the count (which only the device knows until the call completes): one
// candidate per 128 sites plus a floor -- about seven times the called fraction of a 30x genome, 2.5 times that of the
// 60x stress configs.  A call that finds more raises kErrRecordOverflow, and is repeated once with buffers of the size
// it asked for (the buffers only grow, so this happens at most once per context and workload).
// (The floor also covers the side-buffer entries the evaluator's warps reserve in chunks of 8 and leave unused: at most 7
// per warp of its grid, 33,152 in all.)
static uint32_t record_bound(uint64_t segSites) {
  uint64_t per = 128, floor = 65536;
  if (const char *e = getenv("GT16_RECORD_BOUND")) {      // test hook "sites per record:floor": a tiny bound exercises the repeat
    unsigned long a = 0, b = 0;
    if (sscanf(e, "%lu:%lu", &a, &b) == 2 && a >= 1) { per = a; floor = b; }
  }
  return (uint32_t) std::min<uint64_t>(segSites / per + floor, 0x7FFFFFFFull / 12);
}
static const int kRetryLargerBuffers = 1;             // internal return code of call_common
static const uint32_t kSegBatchesDefault = 2;         // batches per segment of a host-fed range (GT16_SEG_BATCHES overrides)

#define CUF(call)                                                                                  \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      char b_[512];                                                                                \
      snprintf(b_, sizeof(b_), "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
      ctx->err = b_;                                                                               \
      (void) cudaGetLastError();                                                                   \
      return GT16_ECUDA;                                                                           \
    }                                                                                              \
  } while (0)

static int get_event(gt16_ctx *ctx, size_t k, cudaEvent_t *out) {
  std::vector<cudaEvent_t> &pool = ctx->evPool[ctx->evCur];
  while (pool.size() <= k) {
    cudaEvent_t e;
    CUF(cudaEventCreate(&e));
    pool.push_back(e);
  }
  *out = pool[k];
  return GT16_OK;
}

// Device arrays of one snapshot.  Every array is over-allocated by kStreamPadSites positions so that the
// TMA streaming kernel may copy whole tiles; counters[L] duplicates counters[L-1] (the reference clamps the
// indel look-ahead of the last position, see decode_indel_refs).
static int alloc_snapshot(gt16_ctx *ctx, const gt16_snapshot_view *view, gt16_dev_snapshot *ds,
                          std::vector<uint32_t> &ovfBase, bool arena, uint64_t winLo, uint64_t winHi) {
  const uint32_t L = view->text_length;
  size_t nOvf = 0;
  ovfBase.assign(view->n_overflow_regions + 1, 0);
  for (uint32_t r = 0; r < view->n_overflow_regions; ++r) { ovfBase[r] = (uint32_t) nOvf; nOvf += view->overflow_size[r]; }
  ovfBase[view->n_overflow_regions] = (uint32_t) nOvf;
  ds->winLo = winLo; ds->winHi = winHi;
  const size_t bC = ((size_t) (winHi - winLo) + kStreamPadSites) * 8, bO = nOvf * 32 + 32, bB = ovfBase.size() * 4,
               bP = (size_t) view->pool_bytes + 16,
               bR = ((size_t) view->packed_ref_words + kStreamPadSites / 16 + 4) * 4,
               bI = ((((size_t) L + 31) >> 5) + kStreamPadSites / 32 + 4) * 4;
  ds->borrowed = arena;
  if (arena) {
    CUF(ctx->sCounters.ensure(bC)); CUF(ctx->sOvf.ensure(bO)); CUF(ctx->sOvfBase.ensure(bB)); CUF(ctx->sPool.ensure(bP));
    CUF(ctx->sRef.ensure(bR));
    if (view->invalid_bitmap) CUF(ctx->sInvalid.ensure(bI));
    ds->countersAlloc = ctx->sCounters.p; ds->ovf = ctx->sOvf.p; ds->ovfBase = ctx->sOvfBase.p; ds->pool = ctx->sPool.p;
    ds->ref = ctx->sRef.p; ds->invalid = view->invalid_bitmap ? ctx->sInvalid.p : nullptr;
  } else {
    CUF(cudaMalloc(&ds->countersAlloc, bC)); CUF(cudaMalloc(&ds->ovf, bO)); CUF(cudaMalloc(&ds->ovfBase, bB));
    CUF(cudaMalloc(&ds->pool, bP)); CUF(cudaMalloc(&ds->ref, bR));
    if (view->invalid_bitmap) CUF(cudaMalloc(&ds->invalid, bI));
  }
  ds->counters = (char *) ds->countersAlloc - (size_t) winLo * 8;     // winLo is a multiple of 4096: alignment is kept
  SnapView &v = ds->v;
  v.counters = (const uint64_t *) ds->counters;
  v.invalid = (const uint32_t *) ds->invalid;
  v.ovf = (const uint8_t *) ds->ovf;
  v.ovfBase = (const uint32_t *) ds->ovfBase;
  v.pool = (const uint8_t *) ds->pool;
  v.packedRef = (const uint32_t *) ds->ref;
  v.L = L;
  v.nOvfRegions = view->n_overflow_regions;
  v.poolBytes = view->pool_bytes;
  v.refWords = view->packed_ref_words;
  set_ovf_geometry(v);
  return GT16_OK;
}

// everything except the counters (0.375 B/site + overflow records + pool), plus the padding.  [loPos, hiPos] is the
// range of positions the call evaluates: only what those positions can reach is copied --
//   packed reference: +-70,000 positions (a homopolymer run is followed until its count saturates at 65,535,
//                     VariantCaller.cpp:659-685; deleted bases and windows reach < 1,000 positions),
//   invalid bitmap  : +-1,024 positions (windows, neighbour tests, look-ahead),
//   overflow arrays : the regions (findRegionByPosition) of positions within +-1,024,
// so that a shard of a sharded genome moves its own part of the side arrays over the host link, not all of them.
static const uint32_t kRefReach = 70000, kSideReach = 1024;
static uint32_t host_ovf_region(const gt16_snapshot_view *view, uint32_t p) {
  const uint32_t n = view->n_overflow_regions, len = view->text_length / n;
  if (len == 0) return n - 1;
  const uint32_t r = p / len;
  return r < n - 1 ? r : n - 1;
}
static int upload_aux(gt16_ctx *ctx, const gt16_snapshot_view *view, gt16_dev_snapshot *ds,
                      const std::vector<uint32_t> &ovfBase, cudaStream_t s, uint32_t loPos, uint32_t hiPos) {
  const uint32_t L = view->text_length;
  if (hiPos >= L) hiPos = L - 1;
  if (loPos > hiPos) loPos = hiPos;
  const uint32_t sLo = loPos > kSideReach ? loPos - kSideReach : 0, sHi = (uint32_t) std::min<uint64_t>((uint64_t) hiPos + kSideReach, L - 1);
  if (view->n_overflow_regions) {
    const uint32_t rLo = host_ovf_region(view, sLo), rHi = host_ovf_region(view, sHi);
    for (uint32_t r = rLo; r <= rHi; ++r)
      if (view->overflow_size[r])
        CUF(cudaMemcpyAsync((char *) ds->ovf + (size_t) ovfBase[r] * 32, view->overflow[r], (size_t) view->overflow_size[r] * 32,
                            cudaMemcpyHostToDevice, s));
  }
  CUF(cudaMemcpyAsync(ds->ovfBase, ovfBase.data(), ovfBase.size() * 4, cudaMemcpyHostToDevice, s));
  if (view->pool_bytes) CUF(cudaMemcpyAsync(ds->pool, view->pool, view->pool_bytes, cudaMemcpyHostToDevice, s));
  {
    const size_t wLo = (loPos > kRefReach ? loPos - kRefReach : 0) >> 4;
    const size_t wHi = std::min<size_t>(view->packed_ref_words, (((size_t) hiPos + kRefReach) >> 4) + 1);
    if (wHi > wLo)
      CUF(cudaMemcpyAsync((uint32_t *) ds->ref + wLo, view->packed_ref + wLo, (wHi - wLo) * 4, cudaMemcpyHostToDevice, s));
    CUF(cudaMemsetAsync((char *) ds->ref + (size_t) view->packed_ref_words * 4, 0, (kStreamPadSites / 16 + 4) * 4, s));
  }
  if (view->invalid_bitmap) {
    const size_t w = ((size_t) L + 31) >> 5;
    const size_t wLo = sLo >> 5, wHi = std::min<size_t>(w, ((size_t) sHi >> 5) + 1);
    if (wHi > wLo)
      CUF(cudaMemcpyAsync((uint32_t *) ds->invalid + wLo, view->invalid_bitmap + wLo, (wHi - wLo) * 4, cudaMemcpyHostToDevice, s));
    CUF(cudaMemsetAsync((char *) ds->invalid + w * 4, 0, (kStreamPadSites / 32 + 4) * 4, s));
  }
  CUF(cudaMemsetAsync((char *) ds->counters + (size_t) ds->winHi * 8, 0, kStreamPadSites * 8, s));
  return GT16_OK;
}

// Counter positions [lo, hi) that calls on the processed range [begin, end) may touch: their own sites; the sites after
// `end` that the windows / neighbour tests of the last sites walk in processed order when the range does not end on a
// batch boundary (~360 sites, which in a multi-region list may lie in later regions -- overlapping regions: the highest
// end so far); 512 positions before the first site (left window margin) and 1,024 after the last (indel look-ahead,
// right margin).  lo is a multiple of 4096 (TMA source alignment of the biased counter pointer).
// (touch_range(), below)

// after the counter element L-1 has been copied: counters[L] = counters[L-1]
static int finish_counters(gt16_ctx *ctx, gt16_dev_snapshot *ds, cudaStream_t s) {
  const uint32_t L = ds->v.L;
  if (ds->winHi != L) return GT16_OK;
  CUF(cudaMemcpyAsync((char *) ds->counters + (size_t) L * 8, (char *) ds->counters + (size_t) (L - 1) * 8, 8,
                      cudaMemcpyDeviceToDevice, s));
  return GT16_OK;
}

// ---- sparse views: {u32 position; u64 counter} records, 12 bytes each, ascending positions -----------------------
static const uint64_t kSparseChunk = 1u << 21;          // records per cudaMemcpyAsync (24 MiB)
static inline uint32_t sparse_pos(const gt16_snapshot_view *v, uint64_t k) {
  uint32_t p;
  memcpy(&p, (const char *) v->sparse_entries + k * 12, 4);
  return p;
}
// first record whose position is >= pos
static uint64_t sparse_lower_bound(const gt16_snapshot_view *v, uint64_t pos) {
  uint64_t lo = 0, hi = v->n_sparse_entries;
  while (lo < hi) { const uint64_t mid = (lo + hi) / 2; if ((uint64_t) sparse_pos(v, mid) < pos) lo = mid + 1; else hi = mid; }
  return lo;
}
// copy records [eLo, eHi) to the context's staging buffer and scatter them into ds->counters (positions in the window
// only) on stream s, in sparse_chunks(eLo, eHi) chunks; chunkDone (optional) holds one event per chunk, recorded behind
// its scatter (the last one also behind counters[L] = counters[L-1])
static size_t sparse_chunks(uint64_t eLo, uint64_t eHi) { return eHi > eLo ? (size_t) ((eHi - eLo + kSparseChunk - 1) / kSparseChunk) : 1; }
static int finish_counters(gt16_ctx *ctx, gt16_dev_snapshot *ds, cudaStream_t s);
static int upload_sparse(gt16_ctx *ctx, const gt16_snapshot_view *view, gt16_dev_snapshot *ds, uint64_t eLo, uint64_t eHi,
                         cudaStream_t s, const std::vector<cudaEvent_t> *chunkDone) {
  if (eHi < eLo) eHi = eLo;
  CUF(ctx->sSparse.ensure((size_t) (eHi - eLo) * 12 + 16));
  const size_t nChunks = sparse_chunks(eLo, eHi);
  for (size_t c = 0; c < nChunks; ++c) {
    const uint64_t a = eLo + c * kSparseChunk, b = std::min<uint64_t>(a + kSparseChunk, eHi);
    // the copies queue up back to back on s; the scatter of a chunk runs on the aux stream behind its copy
    if (b > a) {
      char *dst = (char *) ctx->sSparse.p + (size_t) (a - eLo) * 12;
      CUF(cudaMemcpyAsync(dst, (const char *) view->sparse_entries + a * 12, (size_t) (b - a) * 12, cudaMemcpyHostToDevice, s));
      CUF(cudaEventRecord(ctx->sparseCopied, s));
      CUF(cudaStreamWaitEvent(ctx->aux, ctx->sparseCopied, 0));
      launch_scatter_sparse(dst, b - a, (uint64_t *) ds->counters, view->text_length, (uint32_t) ds->winLo,
                            (uint32_t) std::min<uint64_t>(ds->winHi, view->text_length), a ? sparse_pos(view, a - 1) : 0xFFFFFFFFu,
                            ctx->errWord, ctx->aux);
      ctx->launches++;
    } else if (c == 0) {                       // no records at all: the aux stream still has to stand behind the zeroing
      CUF(cudaEventRecord(ctx->sparseCopied, s));
      CUF(cudaStreamWaitEvent(ctx->aux, ctx->sparseCopied, 0));
    }
    if (c + 1 == nChunks) { int rc = finish_counters(ctx, ds, ctx->aux); if (rc) return rc; }
    if (chunkDone) CUF(cudaEventRecord((*chunkDone)[c], ctx->aux));
  }
  CUF(cudaEventRecord(ctx->sparseCopied, ctx->aux));       // whoever synchronises s next also waits for the scatters
  CUF(cudaStreamWaitEvent(s, ctx->sparseCopied, 0));
  return GT16_OK;
}
// a whole window from a sparse view: zeros, then the records that fall into it
static int upload_sparse_window(gt16_ctx *ctx, const gt16_snapshot_view *view, gt16_dev_snapshot *ds, cudaStream_t s) {
  CUF(cudaMemsetAsync((char *) ds->counters + (size_t) ds->winLo * 8, 0, (size_t) (ds->winHi - ds->winLo) * 8, s));
  return upload_sparse(ctx, view, ds, sparse_lower_bound(view, ds->winLo), sparse_lower_bound(view, ds->winHi), s, nullptr);
}

// after a synchronised sparse upload: did a record lie out of order or beyond the genome?
static int sparse_upload_errors(gt16_ctx *ctx) {
  uint32_t errw = 0;
  CUF(cudaMemcpy(&errw, ctx->errWord, 4, cudaMemcpyDeviceToHost));
  if (errw & kErrBadSnapshot) {
    CUF(cudaMemset(ctx->errWord, 0, 4));
    return fail(ctx, GT16_EDATA, "sparse snapshot records are not in ascending position order or lie beyond the genome");
  }
  return GT16_OK;
}

static int check_view(gt16_ctx *ctx, const gt16_snapshot_view *v) {
  if (!v || (!v->counters && !v->sparse_entries) || !v->text_length || !v->packed_ref || !v->n_overflow_regions ||
      !v->overflow_size || !v->overflow)
    return fail(ctx, GT16_EINVAL, "incomplete snapshot view");
  if ((size_t) v->packed_ref_words * 16 < v->text_length) return fail(ctx, GT16_EINVAL, "packed reference shorter than the snapshot");
  return GT16_OK;
}

extern "C" void gt16_snapshot_release(gt16_ctx *ctx, gt16_dev_snapshot *ds) {
  if (!ds) return;
  if (ctx) cudaSetDevice(ctx->device);
  if (!ds->borrowed) {
    cudaFree(ds->countersAlloc); cudaFree(ds->invalid); cudaFree(ds->ovf); cudaFree(ds->ovfBase); cudaFree(ds->pool); cudaFree(ds->ref);
  }
  delete ds;
}

extern "C" int gt16_snapshot_upload(gt16_ctx *ctx, const gt16_snapshot_view *view, gt16_dev_snapshot **out) {
  if (!ctx || !out) return GT16_EINVAL;
  int rc = check_view(ctx, view);
  if (rc) return rc;
  CU(cudaSetDevice(ctx->device));
  gt16_dev_snapshot *ds = new gt16_dev_snapshot();
  std::vector<uint32_t> ovfBase;
  rc = alloc_snapshot(ctx, view, ds, ovfBase, false, 0, view->text_length);
  if (rc == GT16_OK) rc = upload_aux(ctx, view, ds, ovfBase, ctx->copy, 0, view->text_length - 1);
  if (rc == GT16_OK && !view->counters) rc = upload_sparse_window(ctx, view, ds, ctx->copy);
  else if (rc == GT16_OK) {
    const uint32_t L = view->text_length;
    for (uint64_t a = 0; a < L && rc == GT16_OK; a += kCopyChunkSites) {
      uint64_t b = std::min<uint64_t>(a + kCopyChunkSites, L);
      if (cudaMemcpyAsync((char *) ds->counters + a * 8, view->counters + a, (b - a) * 8, cudaMemcpyHostToDevice, ctx->copy) != cudaSuccess)
        rc = fail(ctx, GT16_ECUDA, "counter upload failed");
    }
  }
  if (rc == GT16_OK) rc = finish_counters(ctx, ds, ctx->copy);
  if (rc == GT16_OK && cudaStreamSynchronize(ctx->copy) != cudaSuccess) rc = fail(ctx, GT16_ECUDA, "upload synchronize failed");
  if (rc == GT16_OK && !view->counters) rc = sparse_upload_errors(ctx);
  if (rc != GT16_OK) { gt16_snapshot_release(ctx, ds); return rc; }
  *out = ds;
  return GT16_OK;
}

namespace {

// The region list of a call in processed order.  The host arrays are shared with the context's cache: a call with the
// same list as the previous one (every call of a bench loop, every piece of a shard; 200,000 targets for an exome) builds
// nothing and copies nothing.
struct PlanData {
  std::vector<gt16_region> source;  // the caller's list as it was passed (cache key)
  uint32_t L = 0;
  std::vector<uint32_t> start;
  std::vector<uint64_t> first;      // nRegions + 1
  std::vector<uint64_t> maxEnd;     // running maximum of the region ends (regions may overlap)
  bool ascending = true;
};
struct RegionPlan {
  std::shared_ptr<const PlanData> d;
  DevSiteMap map;
  const std::vector<uint32_t> &start() const { return d->start; }
  const std::vector<uint64_t> &first() const { return d->first; }
  const std::vector<uint64_t> &maxEnd() const { return d->maxEnd; }
  bool ascending() const { return d->ascending; }
  uint32_t region_of(uint64_t i) const {           // index (in start / first) of the region that holds processed site i
    size_t lo = 0, hi = d->start.size();
    while (hi - lo > 1) { size_t mid = (lo + hi) / 2; if (d->first[mid] <= i) lo = mid; else hi = mid; }
    return (uint32_t) lo;
  }
  uint32_t pos_of(uint64_t i) const { const uint32_t r = region_of(i); return d->start[r] + (uint32_t) (i - d->first[r]); }
  uint64_t max_end_upto(uint64_t i) const { return d->maxEnd[region_of(i)]; }   // highest region end among the regions up to the one holding site i
  uint64_t hi_pos(uint64_t i) const {              // highest position any site up to processed index i can have
    const uint64_t p = pos_of(i);
    return d->start.size() > 1 ? std::max<uint64_t>(p, max_end_upto(i) - 1) : p;
  }
};

int make_plan(gt16_ctx *ctx, uint32_t L, const gt16_region *regions, uint32_t nRegions, RegionPlan &pl) {
  gt16_region whole = {0, L};
  if (!regions || !nRegions) { regions = &whole; nRegions = 1; }
  const PlanData *cached = static_cast<const PlanData *>(ctx->planCache.get());
  if (!(cached && ctx->planStart.p && cached->L == L && cached->source.size() == nRegions &&
        memcmp(cached->source.data(), regions, (size_t) nRegions * sizeof(gt16_region)) == 0)) {
    std::shared_ptr<PlanData> fresh = std::make_shared<PlanData>();
    PlanData &d = *fresh;
    d.L = L;
    d.source.assign(regions, regions + nRegions);
    uint64_t acc = 0;
    for (uint32_t r = 0; r < nRegions; ++r) {
      if (regions[r].start >= L || regions[r].end > L)       // VariantCaller.cpp:623-627
        return fail(ctx, GT16_EINVAL, "[Error] Positions in exome region are incompatible with the index, please generate the exome list with the correct index.");
      if (regions[r].end <= regions[r].start) continue;      // empty loop body in the reference
      if (!d.start.empty() && regions[r].start < d.start.back()) d.ascending = false;
      d.start.push_back(regions[r].start);
      d.first.push_back(acc);
      d.maxEnd.push_back(std::max<uint64_t>(d.maxEnd.empty() ? 0 : d.maxEnd.back(), regions[r].end));
      acc += regions[r].end - regions[r].start;
    }
    if (d.start.empty()) { d.start.push_back(0); d.first.push_back(0); d.maxEnd.push_back(0); }
    d.first.push_back(acc);
    // the device copy of the plan
    ctx->planCache.reset();
    CUF(ctx->planStart.ensure(d.start.size() * 4));
    CUF(ctx->planFirst.ensure(d.first.size() * 8));
    CUF(cudaMemcpyAsync(ctx->planStart.p, d.start.data(), d.start.size() * 4, cudaMemcpyHostToDevice, ctx->compute));
    CUF(cudaMemcpyAsync(ctx->planFirst.p, d.first.data(), d.first.size() * 8, cudaMemcpyHostToDevice, ctx->compute));
    std::vector<uint32_t> tile;
    if (d.start.size() > 1) {
      // cursor of the streaming kernel: the region that holds processed site 256 t
      const size_t nTiles = (size_t) ((acc + kCursorTile - 1) / kCursorTile);
      tile.resize(nTiles + 1);
      size_t r = 0;
      for (size_t t = 0; t < nTiles; ++t) {
        const uint64_t i = (uint64_t) t * kCursorTile;
        while (r + 1 < d.start.size() && d.first[r + 1] <= i) ++r;
        tile[t] = (uint32_t) r;
      }
      tile[nTiles] = (uint32_t) (d.start.size() - 1);
      CUF(ctx->planTile.ensure(tile.size() * 4));
      CUF(cudaMemcpyAsync(ctx->planTile.p, tile.data(), tile.size() * 4, cudaMemcpyHostToDevice, ctx->compute));
    }
    CUF(cudaStreamSynchronize(ctx->compute));     // the host vectors may be pageable
    ctx->planCache = fresh;
  }
  pl.d = std::static_pointer_cast<const PlanData>(ctx->planCache);
  pl.map.m.regStart = (const uint32_t *) ctx->planStart.p;
  pl.map.m.regFirst = (const uint64_t *) ctx->planFirst.p;
  pl.map.m.nRegions = (uint32_t) pl.d->start.size();
  pl.map.m.tileReg = pl.d->start.size() > 1 ? (const uint32_t *) ctx->planTile.p : nullptr;
  pl.map.totalSites = pl.d->first.back();
  return GT16_OK;
}

void touch_range(const RegionPlan &pl, uint32_t L, uint64_t begin, uint64_t end, uint32_t *loPos, uint32_t *hiPos,
                 uint64_t *lo, uint64_t *hi) {
  const uint64_t reachEnd = std::min<uint64_t>(pl.map.totalSites, end + 512);
  *loPos = pl.ascending() ? pl.pos_of(begin) : 0;
  *hiPos = pl.ascending() ? (uint32_t) std::min<uint64_t>(L - 1, pl.hi_pos(reachEnd - 1)) : L - 1;
  *lo = *loPos > 512 ? (uint64_t) ((*loPos - 512) & ~4095u) : 0;
  *hi = std::min<uint64_t>((uint64_t) *hiPos + 1024, L);
}

float ms_between(cudaEvent_t a, cudaEvent_t b) { float m = 0.f; cudaEventElapsedTime(&m, a, b); return m; }

// can [pos(begin), pos(end)) go through the TMA-staged kernel?
bool tma_eligible(const gt16_ctx *ctx, const RegionPlan &pl, uint64_t begin, bool useTables) {
  return useTables && pl.start().size() == 1 && ((pl.start()[0] + (begin - pl.first()[0])) % 128) == 0 && ctx->s1 != nullptr;
}

void stream_range(gt16_ctx *ctx, const SnapView &V, const RegionPlan &pl, uint64_t a, uint64_t b, bool useTables,
                  const DecisionTables &D, const EvalParams &P, uint32_t *mask) {
  if (tma_eligible(ctx, pl, a, useTables)) {
    uint32_t pBeg = pl.start()[0] + (uint32_t) (a - pl.first()[0]);
    launch_stream_tma(V, pBeg, pBeg + (uint32_t) (b - a), D, ctx->shortcutAlt, mask, ctx->numSMs, ctx->compute);
  } else {
    launch_stream(V, pl.map, a, (uint32_t) (b - a), useTables, D, ctx->shortcutAlt, P, mask, ctx->errWord, ctx->compute);
  }
  ctx->launches++;
}

// make room for `need` more records in the pinned result buffer of `calls` (contents are preserved)
int reserve_results(gt16_ctx *ctx, gt16_calls *calls, uint64_t need, uint64_t hint) {
  const uint64_t want = calls->n + need;
  if (want * sizeof(gt16_called_site) <= calls->cap) return GT16_OK;
  if (!calls->pins) { calls->tooSmall = true; return GT16_OK; }
  const uint64_t target = std::max<uint64_t>(want + want / 2, hint);
  size_t cap = 0;
  gt16_called_site *fresh = (gt16_called_site *) ctx->pins->take((size_t) target * sizeof(gt16_called_site), &cap);
  if (!fresh) return fail(ctx, GT16_ECUDA, "cudaHostAlloc of the result buffer failed");
  if (calls->sites) {
    CUF(cudaStreamSynchronize(ctx->d2h));                  // copies into the old buffer must have landed
    memcpy(fresh, calls->sites, (size_t) calls->n * sizeof(gt16_called_site));
    ctx->pins->give(calls->sites, calls->cap);
  }
  calls->sites = fresh;
  calls->cap = cap;
  return GT16_OK;
}

// One segment [a, b) of the processed range after its streaming pass, all of it enqueued without a host round trip:
//   k_eval_mask   exact evaluation of the flagged sites straight from the mask; clears the bits of non-candidates and
//                 leaves likelihoods / posteriors of the candidates in part[] (device-side counter dCounts[seg])
//   count + scan  block counts of what is left in the mask
//   k_rank        place of every candidate in the ordered output
//   k_candidates  records (persistent grid that reads the count on the device), with every strand-bias / base-quality
//                 Fisher test the memo cache holds; the others go to a list
//   k_bias        the walks of the listed tests (persistent grid, device-side count; next to nothing once the cache is warm)
// `mask` is the flag bitmap of the whole call (bit 0 = site `begin`).
int finish_segment(gt16_ctx *ctx, const SnapView &V, RegionPlan &pl, uint64_t begin, uint64_t a, uint64_t b,
                   bool usedTables, const DecisionTables &D, size_t &evNext, uint32_t seg, cudaEvent_t evStream0,
                   cudaEvent_t evStream1, std::vector<cudaEvent_t> &timing, const std::vector<cudaEvent_t> &segCopied) {
  EvalParams P = eval_params(ctx);
  const uint32_t n = (uint32_t) (b - a), off = (uint32_t) (a - begin), cap = ctx->recCap;
  uint32_t *maskAll = (uint32_t *) ctx->mask.p, *counts = (uint32_t *) ctx->counts.p;
  uint32_t *dCount = ctx->dCounts + 2 * seg;
  EvalOut *part = (EvalOut *) ctx->part.p;
  uint32_t *rankOf = (uint32_t *) ctx->rankOf.p, *list = (uint32_t *) ctx->list.p;
  cudaEvent_t ev[4];
  int rc;
  for (int k = 0; k < 4; ++k) if ((rc = get_event(ctx, evNext++, &ev[k]))) return rc;
  CUF(cudaEventRecord(ev[0], ctx->compute));
  launch_eval_mask(V, pl.map, begin, off, n, usedTables, D, P, maskAll, ctx->errWord, ctx->dFlagged, part, dCount, cap, ctx->compute);
  CUF(cudaEventRecord(ev[1], ctx->compute));
  uint32_t *chunkOffsets = counts + scan_blocks(n) + 1;        // behind the block counts (counts.ensure)
  uint32_t *subCounts = chunkOffsets + 1024;                   // ... and the counts of the blocks' runs of 32 words behind those
  launch_mask_count(maskAll + off / 32, n, counts, subCounts, ctx->compute);
  launch_scan_counts(counts, scan_blocks(n), chunkOffsets, ctx->dCounts + 2 * seg + 1, ctx->compute);
  launch_rank(maskAll, off, counts, chunkOffsets, subCounts, part, dCount, cap, rankOf, list, ctx->compute);
  CUF(cudaEventRecord(ev[2], ctx->compute));
  // record buffers alternate between segments: a segment's copy home may still be in flight while the next one is written
  gt16_called_site *recs = (gt16_called_site *) ((seg & 1) ? ctx->recs2.p : ctx->recs.p);
  if (seg >= 2) CUF(cudaStreamWaitEvent(ctx->compute, segCopied[seg - 2], 0));   // this buffer's previous contents are on the host
  CUF(cudaMemsetAsync(ctx->dMissed, 0, 4, ctx->compute));
  const uint32_t missedCap = (uint32_t) std::min<uint64_t>((uint64_t) cap * kMissedPerRecord, 0x7FFFFFF0ull);
  launch_candidates(V, pl.map, begin, part, list, dCount, cap, P, ctx->biasCache, kBiasCacheEntries - 1,
                    (BiasTask *) ctx->missed.p, ctx->dMissed, missedCap, recs, ctx->errWord, ctx->numSMs, ctx->compute);
  launch_bias((BiasTask *) ctx->missed.p, ctx->dMissed, missedCap, P, ctx->biasCache, kBiasCacheEntries - 1, recs, ctx->numSMs, ctx->compute);
  ctx->launches += 7;
  CUF(cudaEventRecord(ev[3], ctx->compute));
  timing.push_back(evStream0); timing.push_back(evStream1);
  for (int k = 0; k < 4; ++k) timing.push_back(ev[k]);
  return GT16_OK;
}

}  // namespace

static void fill_timings(gt16_stats &st, cudaEvent_t evBegin, cudaEvent_t evEnd, const std::vector<cudaEvent_t> &timing) {
  float lastCand = 0.f;
  for (size_t k = 0; k + 6 <= timing.size(); k += 6) {
    st.ms_stream += ms_between(timing[k], timing[k + 1]);
    st.ms_scan += ms_between(timing[k + 1], timing[k + 2]) + ms_between(timing[k + 3], timing[k + 4]);
    st.ms_eval += ms_between(timing[k + 2], timing[k + 3]);
    st.ms_candidates += ms_between(timing[k + 4], timing[k + 5]);
    lastCand = ms_between(evBegin, timing[k + 5]);
  }
  st.ms_total_device = ms_between(evBegin, evEnd);
  st.ms_d2h = st.ms_total_device - lastCand;            // tail after the last candidate slice was written
}

static int call_common(gt16_ctx *ctx, const gt16_snapshot_view *hostView, const gt16_dev_snapshot *resident,
                       const gt16_region *regions, uint32_t nRegions, uint64_t begin, uint64_t end,
                       gt16_calls **out, gt16_stats *stats, gt16_dev_calls **devOutHandle = nullptr, bool nowait = false);

static int report_device_errors(gt16_ctx *ctx, uint32_t errw) {
  if (errw) CUF(cudaMemset(ctx->errWord, 0, 4));
  if (errw & kErrBadSnapshot) return fail(ctx, GT16_EDATA, "snapshot refers to an overflow record or pool range that does not exist");
  if (errw & kErrGenotype) return fail(ctx, GT16_EGENOTYPE, "Error in selecting Genotype");
  return GT16_OK;
}

// completion of a device-resident call: wait for its tail, read the record count and the error word, fill in flagged and
// the timings.  A call that found more candidates than the record buffers hold is repeated here, blocking, with larger ones.
static int finish_dev_call(gt16_ctx *ctx, gt16_dev_calls *c) {
  if (!c->pending) return GT16_OK;
  CUF(cudaEventSynchronize(c->evEnd));
  CUF(cudaGetLastError());
  c->pending = false;
  const uint32_t reserved = c->slot[0], found = c->slot[1], errw = c->slot[2];
  if (reserved > c->cap) {
    CUF(cudaDeviceSynchronize());                                 // later calls on the context may be in flight
    CUF(cudaMemset(ctx->errWord, 0, 4));
    ctx->recCapWanted = reserved + reserved / 8 + 1024;
    gt16_dev_calls *again = nullptr;
    gt16_stats st2;
    int rc = call_common(ctx, nullptr, c->snap, c->regions.empty() ? nullptr : c->regions.data(), (uint32_t) c->regions.size(),
                         c->begin, c->end, nullptr, &st2, &again, false);
    if (rc != GT16_OK) return rc;
    c->recs = again->recs; c->n = again->n; c->cap = again->cap; c->st = again->st;
    delete again;
    return GT16_OK;
  }
  c->n = found;
  c->st.called = found;
  c->st.flagged = c->slot[4];
  int rc = report_device_errors(ctx, errw);
  if (rc != GT16_OK) return rc;
  fill_timings(c->st, c->evBegin, c->evEnd, c->timing);
  return GT16_OK;
}

static int call_common(gt16_ctx *ctx, const gt16_snapshot_view *hostView, const gt16_dev_snapshot *resident,
                       const gt16_region *regions, uint32_t nRegions, uint64_t begin, uint64_t end,
                       gt16_calls **out, gt16_stats *stats, gt16_dev_calls **devOutHandle, bool nowait) {
  if (!ctx || (!out && !devOutHandle)) return GT16_EINVAL;
  ctx->evCur = (ctx->evCur + 1) % gt16_ctx::kCallRing;
  ctx->hPinned = ctx->hPinnedRing + 8 * ctx->evCur;
  if (!ctx->havePrefill) return fail(ctx, GT16_ESTATE, "gt16_configure / gt16_prefill_likelihood_cache must run first");
  std::unique_ptr<gt16_dev_calls> devOut(devOutHandle ? new gt16_dev_calls() : nullptr);
  CU(cudaSetDevice(ctx->device));
  const uint32_t L = resident ? resident->v.L : hostView->text_length;
  RegionPlan pl;
  int rc = make_plan(ctx, L, regions, nRegions, pl);
  if (rc) return rc;
  if (end > pl.map.totalSites) end = pl.map.totalSites;
  if (begin > end) begin = end;
  std::unique_ptr<gt16_calls> calls(new gt16_calls());
  if (ctx->extDst) { calls->sites = ctx->extDst; calls->cap = (size_t) ctx->extCap * sizeof(gt16_called_site); }
  else calls->pins = ctx->pins;
  // an error return must not hand the page-locked result buffer back to the pool while copies into it are queued
  struct DrainD2H {
    gt16_ctx *c; bool armed = true;
    ~DrainD2H() { if (armed) { cudaStreamSynchronize(c->d2h); (void) cudaGetLastError(); } }
  } drainD2H{ctx};
  gt16_stats st;
  memset(&st, 0, sizeof(st));
  if (begin == end) {                        // an empty shard (more ranks than batches) is not an error, aligned or not
    drainD2H.armed = false;
    if (devOutHandle) *devOutHandle = devOut.release(); else *out = calls.release();
    if (stats) *stats = st;
    return GT16_OK;
  }
  if (begin % kBatch) return fail(ctx, GT16_EINVAL, "site_begin must be a multiple of GT16_BATCH_SITES");
  if (end - begin >= 0xFFFFFFFFull) return fail(ctx, GT16_EINVAL, "range too long");
  const uint32_t n = (uint32_t) (end - begin);
  const uint32_t launches0 = ctx->launches;
  const bool useTables = ctx->useDecisions && ctx->haveDecisions;
  DecisionTables D = {ctx->s1, ctx->s2p, ctx->s2q, ctx->g4};
  EvalParams P = eval_params(ctx);
  st.sites = n;

  size_t evNext = 0;
  cudaEvent_t evBegin, evEnd;
  if ((rc = get_event(ctx, evNext++, &evBegin)) || (rc = get_event(ctx, evNext++, &evEnd))) return rc;
  CUF(ctx->mask.ensure((((size_t) n + 31) / 32 + 64) * 4));
  // A call whose records stay in HBM is one segment: one streaming launch and six more kernels, nothing else.  A call
  // that returns its records to the host is cut into segments of kSegBatches batches, so that a segment's records
  // travel home while the next one computes -- and, when the counters come from the host too, everything after a
  // segment's streaming pass runs while later counter chunks are still on the bus.
  uint32_t segBatches = kSegBatchesDefault;
  if (const char *e = getenv("GT16_SEG_BATCHES")) { int v = atoi(e); if (v >= 1 && v <= 64) segBatches = (uint32_t) v; }
  const uint64_t segSites = devOutHandle ? (uint64_t) n : (uint64_t) segBatches * kBatch;
  const uint64_t segLongest = std::min<uint64_t>(segSites, n);
  CUF(ctx->counts.ensure(((size_t) scan_blocks((uint32_t) segLongest) * (1 + kScanBlock / 32) + 1 + 1024) * 4));    // block counts + chunk offsets + run counts
  {
    const uint32_t capWant = std::max(std::max(ctx->recCap, ctx->recCapWanted), record_bound(segLongest));
    ctx->recCap = capWant;
    CUF(ctx->part.ensure((size_t) capWant * sizeof(EvalOut)));
    CUF(ctx->rankOf.ensure((size_t) capWant * 4));
    CUF(ctx->list.ensure((size_t) capWant * 4));
    CUF(ctx->missed.ensure((size_t) capWant * kMissedPerRecord * sizeof(BiasTask)));
    CUF(ctx->recs.ensure((size_t) capWant * sizeof(gt16_called_site)));
    if (!devOutHandle && n > kBatch) CUF(ctx->recs2.ensure((size_t) capWant * sizeof(gt16_called_site)));   // more than one segment
  }
  uint32_t *mask = (uint32_t *) ctx->mask.p;
  const uint64_t resultHint = std::max<uint64_t>(4096, (uint64_t) n / 512);
  std::vector<cudaEvent_t> timing, segCopied, segDone;
  const uint32_t nSegsBound = (uint32_t) ((n + segSites - 1) / segSites) + 1;
  if (nSegsBound > (uint32_t) gt16_ctx::kMaxSegs) return fail(ctx, GT16_EINVAL, "range too long for GT16_SEG_BATCHES");
  CUF(cudaMemsetAsync(ctx->dFlagged, 0, 4, ctx->compute));
  CUF(cudaMemsetAsync(ctx->dCounts, 0, nSegsBound * 8, ctx->compute));
  CUF(cudaEventRecord(evBegin, ctx->compute));

  gt16_dev_snapshot streamed;
  const SnapView *V = nullptr;
  uint64_t copyLo = 0, copyHi = 0;
  size_t nChunks = 0, chunkEv0 = 0;
  uint32_t chunkSites = kCopyChunkSites;
  bool staged = false;
  const bool sparse = !resident && hostView && !hostView->counters;
  std::vector<uint32_t> sparseLastPos;          // position of the last record of each chunk of a sparse view
  std::unique_ptr<std::atomic<int>[]> enqueued;
  std::atomic<int> stageError{0};
  std::vector<std::thread> stagers;
  std::vector<cudaEvent_t> chunkEvents;
  std::function<bool(uint64_t, uint64_t)> chunk_needed;                     // function scope: the stager threads use them
  std::function<cudaError_t(size_t, const void *, bool)> enqueue_chunk;
  if (resident) {
    V = &resident->v;
    uint32_t loPos, hiPos; uint64_t lo, hi;
    touch_range(pl, L, begin, end, &loPos, &hiPos, &lo, &hi);
    if (lo < resident->winLo || hi > resident->winHi)
      return fail(ctx, GT16_EINVAL, "site range reaches outside the part of the snapshot that was uploaded (gt16_snapshot_upload_range)");
  } else {
    // counters stream in on the copy stream in genome order (32 MiB chunks, all queued up front)
    std::vector<uint32_t> ovfBase;
    uint32_t loPos, hiPos;
    touch_range(pl, L, begin, end, &loPos, &hiPos, &copyLo, &copyHi);
    rc = alloc_snapshot(ctx, hostView, &streamed, ovfBase, true, copyLo, copyHi);
    if (rc == GT16_OK) rc = upload_aux(ctx, hostView, &streamed, ovfBase, ctx->copy, loPos, hiPos);
    if (rc != GT16_OK) return rc;
    V = &streamed.v;
    // page-locked source: cudaMemcpyAsync straight from the caller's buffer.  Pageable source (what a caller gets
    // from malloc/fread; pinning half a gigabyte costs more than copying it): stager threads move 8 MiB chunks
    // through a small page-locked ring.  (The counters pointer may be biased -- a windowed view -- so it is looked
    // at where the call starts reading.)
    uint64_t eLo = 0, eHi = 0;
    if (sparse) {
      // a sparse view: its records (12 B per covered position) travel as they are and are expanded on the device
      eLo = sparse_lower_bound(hostView, copyLo); eHi = sparse_lower_bound(hostView, copyHi);
      nChunks = sparse_chunks(eLo, eHi);
      for (size_t c = 0; c < nChunks; ++c) {
        const uint64_t b = std::min<uint64_t>(eLo + (c + 1) * kSparseChunk, eHi);
        sparseLastPos.push_back(b > eLo + c * kSparseChunk ? sparse_pos(hostView, b - 1) : 0u);
      }
    } else {
      cudaPointerAttributes pa;
      staged = !(cudaPointerGetAttributes(&pa, hostView->counters + copyLo) == cudaSuccess && pa.type == cudaMemoryTypeHost);
      (void) cudaGetLastError();
      chunkSites = staged ? kStageChunkSites : kCopyChunkSites;
      nChunks = (size_t) ((copyHi - copyLo + chunkSites - 1) / chunkSites);
    }
    chunkEv0 = evNext;
    evNext += nChunks;
    chunkEvents.resize(nChunks);      // own copy: stager threads must not look into evPool while it grows
    for (size_t c = 0; c < nChunks; ++c) { rc = get_event(ctx, chunkEv0 + c, &chunkEvents[c]); if (rc) return rc; }
    if (sparse) {
      // positions without a record hold zero counters: the targets' spans (and what windows, look-ahead and batch margins
      // reach around them) are cleared, or the whole window when the call is not a list of ascending targets
      if (pl.ascending() && pl.start().size() > 1) {
        const uint32_t rA = pl.region_of(begin), rB = pl.region_of(std::min<uint64_t>(pl.map.totalSites, end + 512) - 1);
        launch_zero_spans(pl.map.m.regStart, pl.map.m.regFirst, rA, rB, (uint64_t *) streamed.counters, (uint32_t) copyLo,
                          (uint32_t) copyHi, kSideReach, ctx->copy);
        ctx->launches++;
      } else {
        CUF(cudaMemsetAsync((char *) streamed.counters + copyLo * 8, 0, (size_t) (copyHi - copyLo) * 8, ctx->copy));
      }
      if ((rc = upload_sparse(ctx, hostView, &streamed, eLo, eHi, ctx->copy, &chunkEvents))) return rc;
    }
    chunk_needed = [&](uint64_t a, uint64_t b) -> bool {
      if (!(pl.ascending() && pl.start().size() > 1)) return true;
      size_t lo = 0, hi = pl.start().size();                 // sparse targets: skip chunks no region touches
      while (hi - lo > 1) { size_t mid = (lo + hi) / 2; if ((uint64_t) pl.start()[mid] <= a + 0) lo = mid; else hi = mid; }
      // regions may nest or overlap (starts ascend, ends need not: the reference's merge can emit such lists,
      // buildRegionList.cpp:199-216): an earlier region reaches into [a, b) iff the running maximum of the ends does
      if (lo > 0 && pl.maxEnd()[lo - 1] + 512 > a) return true;
      for (size_t r = lo; r < pl.start().size(); ++r) {
        uint64_t rs = pl.start()[r], re = rs + (pl.first()[r + 1] - pl.first()[r]);
        if (rs > b + 512) break;
        if (re + 512 > a && rs < b + 512) return true;
      }
      return false;
    };
    // enqueue chunk c on the copy stream (from `src`, the caller's buffer or a ring slot) and record its event
    enqueue_chunk = [&](size_t c, const void *src, bool needed) -> cudaError_t {
      const uint64_t a = copyLo + c * (uint64_t) chunkSites, b = std::min<uint64_t>(a + chunkSites, copyHi);
      cudaError_t e = needed ? cudaMemcpyAsync((char *) streamed.counters + a * 8, src, (b - a) * 8, cudaMemcpyHostToDevice, ctx->copy)
                             : cudaMemsetAsync((char *) streamed.counters + a * 8, 0, (b - a) * 8, ctx->copy);
      if (e == cudaSuccess && c + 1 == nChunks && copyHi == L)
        e = cudaMemcpyAsync((char *) streamed.counters + (size_t) L * 8, (char *) streamed.counters + (size_t) (L - 1) * 8, 8,
                            cudaMemcpyDeviceToDevice, ctx->copy);          // counters[L] = counters[L-1]
      if (e == cudaSuccess) e = cudaEventRecord(chunkEvents[c], ctx->copy);
      return e;
    };
    if (sparse) {
    } else if (!staged) {
      for (size_t c = 0; c < nChunks; ++c) {
        const uint64_t a = copyLo + c * (uint64_t) chunkSites, b = std::min<uint64_t>(a + chunkSites, copyHi);
        CUF(enqueue_chunk(c, hostView->counters + a, chunk_needed(a, b)));
      }
    } else {
      if (!ctx->stagePin) {
        CUF(cudaHostAlloc(&ctx->stagePin, (size_t) 2 * stage_threads() * kStageChunkSites * 8, cudaHostAllocDefault));
        for (int k = 0; k < kStageSlots; ++k) CUF(cudaEventCreateWithFlags(&ctx->stageEv[k], cudaEventDisableTiming));
      }
      enqueued.reset(new std::atomic<int>[nChunks]);
      for (size_t c = 0; c < nChunks; ++c) enqueued[c].store(0);
      const int dev = ctx->device;
      const int kStageThreads = stage_threads();
      for (int t = 0; t < kStageThreads; ++t)
        stagers.emplace_back([&, t, dev, kStageThreads]() {
          cudaSetDevice(dev);
          int use = 0;                                       // this thread alternates between its two ring slots
          for (size_t c = (size_t) t; c < nChunks; c += kStageThreads, use ^= 1) {
            const int slot = t * 2 + use;
            const uint64_t a = copyLo + c * (uint64_t) chunkSites, b = std::min<uint64_t>(a + chunkSites, copyHi);
            const bool needed = chunk_needed(a, b);
            char *ring = (char *) ctx->stagePin + (size_t) slot * kStageChunkSites * 8;
            cudaError_t e = cudaSuccess;
            if (needed) {
              if (c >= (size_t) 2 * kStageThreads) e = cudaEventSynchronize(ctx->stageEv[slot]);   // the slot's previous copy has left
              if (e == cudaSuccess) memcpy(ring, hostView->counters + a, (size_t) (b - a) * 8);
            }
            if (e == cudaSuccess) e = enqueue_chunk(c, ring, needed);
            if (e == cudaSuccess && needed) e = cudaEventRecord(ctx->stageEv[slot], ctx->copy);
            if (e != cudaSuccess) stageError.store((int) e);
            enqueued[c].store(1, std::memory_order_release);
          }
        });
    }
  }
  // a stream must not be told to wait for an event that has not been recorded yet (that wait would be a no-op):
  // with stager threads the host first waits until chunk c has been enqueued
  auto host_wait_enqueued = [&](size_t c) {
    if (!staged) return;
    while (!enqueued[c].load(std::memory_order_acquire)) std::this_thread::yield();
  };
  struct JoinStagers {
    std::vector<std::thread> &v;
    ~JoinStagers() { for (auto &t : v) if (t.joinable()) t.join(); }
  } joinStagers{stagers};
  // event of the chunk that holds position `need - 1` (host-fed ranges only)
  // make the compute stream wait until every chunk up to the one holding position `need - 1` has arrived.
  // Chunks are enqueued out of order by the stager threads, so each chunk's own event is waited for (once).
  size_t chunksWaited = 0;
  auto wait_chunks = [&](uint64_t need) -> int {
    if (resident) return GT16_OK;
    size_t c = nChunks - 1;
    if (pl.ascending()) {
      need = std::min<uint64_t>(std::max<uint64_t>(need, copyLo + 1), copyHi);
      if (sparse) c = std::min((size_t) (std::lower_bound(sparseLastPos.begin(), sparseLastPos.end(), (uint32_t) (need - 1)) - sparseLastPos.begin()), nChunks - 1);
      else c = std::min((size_t) ((need - 1 - copyLo) / chunkSites), nChunks - 1);
    }
    for (; chunksWaited <= c; ++chunksWaited) {
      host_wait_enqueued(chunksWaited);
      CUF(cudaStreamWaitEvent(ctx->compute, chunkEvents[chunksWaited], 0));
    }
    return GT16_OK;
  };

  // segment boundaries; a host-fed range ends with a segment of its own for the last batch: everything after the
  // streaming pass of that segment is exposed (the bus is idle by then), so it should be short
  std::vector<uint64_t> cuts;
  for (uint64_t a = begin; a < end; a += segSites) cuts.push_back(a);
  if (!resident) {
    const uint64_t lastBatch = begin + ((uint64_t) (n - 1) / kBatch) * kBatch;
    if (lastBatch > cuts.back()) cuts.push_back(lastBatch);
  }
  cuts.push_back(end);
  // records of segment `sg` -> the caller's page-locked buffer, once the segment has finished on the device: the one
  // place where the host learns a count, and it gets there after the NEXT segment has been enqueued
  bool overflow = false;
  auto collect = [&](uint32_t sg) -> int {
    CUF(cudaEventSynchronize(segDone[sg]));
    const uint32_t reserved = ctx->hCounts[2 * sg], cnt = ctx->hCounts[2 * sg + 1];
    cudaEvent_t done;
    int rc2 = get_event(ctx, evNext++, &done);
    if (rc2) return rc2;
    if (reserved > ctx->recCap) {
      overflow = true;
      ctx->recCapWanted = std::max(ctx->recCapWanted, reserved + reserved / 8 + 1024);
    } else if (cnt && !overflow) {
      if ((rc2 = reserve_results(ctx, calls.get(), cnt, resultHint))) return rc2;
      if (!calls->tooSmall) {
        const gt16_called_site *recs = (const gt16_called_site *) ((sg & 1) ? ctx->recs2.p : ctx->recs.p);
        CUF(cudaStreamWaitEvent(ctx->d2h, segDone[sg], 0));
        CUF(cudaMemcpyAsync(calls->sites + calls->n, recs, (size_t) cnt * sizeof(gt16_called_site), cudaMemcpyDeviceToHost, ctx->d2h));
        calls->n += cnt;
      }
      st.called += cnt;
    }
    CUF(cudaEventRecord(done, ctx->d2h));
    segCopied.push_back(done);
    return GT16_OK;
  };
  uint32_t seg = 0;
  for (size_t ci = 0; ci + 1 < cuts.size(); ++ci, ++seg) {
    const uint64_t a = cuts[ci], b = cuts[ci + 1];
    cudaEvent_t es0, es1;
    if ((rc = get_event(ctx, evNext++, &es0)) || (rc = get_event(ctx, evNext++, &es1))) return rc;
    CUF(cudaEventRecord(es0, ctx->compute));
    for (uint64_t c = a; c < b; c += kBatch) {           // classification of a batch waits for the chunk with its last position + 1
      const uint64_t d = std::min<uint64_t>(c + kBatch, b);
      if ((rc = wait_chunks(pl.ascending() ? pl.hi_pos(d - 1) + 2 : copyHi))) return rc;
      if (resident) { stream_range(ctx, *V, pl, a, b, useTables, D, P, mask + (a - begin) / 32); break; }
      stream_range(ctx, *V, pl, c, d, useTables, D, P, mask + (c - begin) / 32);
    }
    CUF(cudaEventRecord(es1, ctx->compute));
    // windows, neighbours' Fisher tests and the indel look-ahead reach up to ~360 positions past the segment
    if ((rc = wait_chunks(pl.ascending() ? pl.hi_pos(std::min<uint64_t>(pl.map.totalSites, b + 512) - 1) + 600 : copyHi))) return rc;
    rc = finish_segment(ctx, *V, pl, begin, a, b, useTables, D, evNext, seg, es0, es1, timing, segCopied);
    if (rc != GT16_OK) return rc;
    if (!devOutHandle) {
      cudaEvent_t sd;
      if ((rc = get_event(ctx, evNext++, &sd))) return rc;
      CUF(cudaMemcpyAsync(ctx->hCounts + 2 * seg, ctx->dCounts + 2 * seg, 8, cudaMemcpyDeviceToHost, ctx->compute));
      CUF(cudaEventRecord(sd, ctx->compute));
      segDone.push_back(sd);
      if (seg >= 1 && (rc = collect(seg - 1))) return rc;
    }
  }
  if (!devOutHandle && (rc = collect(seg - 1))) return rc;
  // drain: results on the host, error word
  cudaEvent_t evD2H;
  if ((rc = get_event(ctx, evNext++, &evD2H))) return rc;
  CUF(cudaEventRecord(evD2H, ctx->d2h));
  CUF(cudaStreamWaitEvent(ctx->compute, evD2H, 0));
  if ((rc = wait_chunks(copyHi))) return rc;
  if (staged) {
    for (auto &t : stagers) t.join();
    if (stageError.load()) { ctx->err = std::string("staged upload failed: ") + cudaGetErrorString((cudaError_t) stageError.load()); return GT16_ECUDA; }
  }
  if (devOutHandle) CUF(cudaMemcpyAsync(ctx->hPinned, ctx->dCounts, 8, cudaMemcpyDeviceToHost, ctx->compute));   // entries reserved / records of the (single) segment
  CUF(cudaMemcpyAsync(ctx->hPinned + 2, ctx->errWord, 12, cudaMemcpyDeviceToHost, ctx->compute));   // error bits, -, sites evaluated
  CUF(cudaEventRecord(evEnd, ctx->compute));
  st.kernel_launches = ctx->launches - launches0;
  if (devOutHandle) {
    // the rest -- waiting for the tail, the record count, the error word, the timings -- is gt16_dev_calls_wait's
    devOut->pending = true; devOut->useTables = useTables; devOut->ctx = ctx;
    devOut->recs = (const gt16_called_site *) ctx->recs.p; devOut->cap = ctx->recCap;
    devOut->snap = resident; devOut->begin = begin; devOut->end = end;
    if (regions && nRegions) devOut->regions.assign(regions, regions + nRegions);
    devOut->evBegin = evBegin; devOut->evEnd = evEnd; devOut->timing = timing; devOut->slot = ctx->hPinned; devOut->st = st;
    if (!nowait) {
      rc = finish_dev_call(ctx, devOut.get());
      if (rc != GT16_OK) return rc;
    }
    if (stats) *stats = devOut->st;
    drainD2H.armed = false;
    *devOutHandle = devOut.release();
    return GT16_OK;
  }
  CUF(cudaStreamSynchronize(ctx->compute));
  CUF(cudaGetLastError());
  const uint32_t errw = ctx->hPinned[2];
  st.flagged = ctx->hPinned[4];
  if (overflow) {                                   // more candidates than the record buffers hold: again, with larger ones
    CUF(cudaDeviceSynchronize());
    CUF(cudaMemset(ctx->errWord, 0, 4));
    return kRetryLargerBuffers;
  }
  if ((rc = report_device_errors(ctx, errw))) return rc;
  fill_timings(st, evBegin, evEnd, timing);
  drainD2H.armed = false;
  *out = calls.release();
  if (stats) *stats = st;
  return GT16_OK;
}

// host-result calls: repeated once when the first attempt found more candidates than the record buffers hold
static int call_with_retry(gt16_ctx *ctx, const gt16_snapshot_view *hostView, const gt16_dev_snapshot *resident,
                           const gt16_region *regions, uint32_t nRegions, uint64_t begin, uint64_t end,
                           gt16_calls **out, gt16_stats *stats) {
  int rc = kRetryLargerBuffers;
  for (int attempt = 0; attempt < 3 && rc == kRetryLargerBuffers; ++attempt)
    rc = call_common(ctx, hostView, resident, regions, nRegions, begin, end, out, stats, nullptr, false);
  if (rc == kRetryLargerBuffers) return fail(ctx, GT16_ECUDA, "record buffers could not be sized");
  return rc;
}

extern "C" {

int gt16_call_sites(gt16_ctx *ctx, const gt16_dev_snapshot *snap, const gt16_region *regions, uint32_t n_regions,
                    uint64_t site_begin, uint64_t site_end, gt16_calls **out, gt16_stats *stats) {
  if (!snap) return GT16_EINVAL;
  return call_with_retry(ctx, nullptr, snap, regions, n_regions, site_begin, site_end, out, stats);
}

int gt16_call_sites_device(gt16_ctx *ctx, const gt16_dev_snapshot *snap, const gt16_region *regions, uint32_t n_regions,
                           uint64_t site_begin, uint64_t site_end, gt16_dev_calls **out, gt16_stats *stats) {
  if (!snap || !out) return GT16_EINVAL;
  return call_common(ctx, nullptr, snap, regions, n_regions, site_begin, site_end, nullptr, stats, out, false);
}
int gt16_call_sites_device_nowait(gt16_ctx *ctx, const gt16_dev_snapshot *snap, const gt16_region *regions, uint32_t n_regions,
                                  uint64_t site_begin, uint64_t site_end, gt16_dev_calls **out, gt16_stats *stats) {
  if (!snap || !out) return GT16_EINVAL;
  return call_common(ctx, nullptr, snap, regions, n_regions, site_begin, site_end, nullptr, stats, out, true);
}
int gt16_dev_calls_wait(gt16_ctx *ctx, gt16_dev_calls *c, gt16_stats *stats) {
  if (!ctx || !c) return GT16_EINVAL;
  CU(cudaSetDevice(ctx->device));
  int rc = finish_dev_call(ctx, c);
  if (stats) *stats = c->st;
  return rc;
}
// the count (and with it the extent of the record array) is known when the call has completed: both complete it if need be
static void complete_quietly(const gt16_dev_calls *c) {
  if (c && c->pending && c->ctx) { cudaSetDevice(c->ctx->device); (void) finish_dev_call(c->ctx, const_cast<gt16_dev_calls *>(c)); }
}
uint64_t gt16_dev_calls_count(const gt16_dev_calls *c) { complete_quietly(c); return c ? c->n : 0; }
const void *gt16_dev_calls_device_ptr(const gt16_dev_calls *c) { complete_quietly(c); return c && c->n ? c->recs : nullptr; }
void gt16_dev_calls_free(gt16_dev_calls *c) { delete c; }
int gt16_dev_calls_fetch(gt16_ctx *ctx, const gt16_dev_calls *c, gt16_calls **out) {
  if (!ctx || !c || !out) return GT16_EINVAL;
  CU(cudaSetDevice(ctx->device));
  if (c->pending) { int rc = finish_dev_call(ctx, const_cast<gt16_dev_calls *>(c)); if (rc != GT16_OK) return rc; }
  std::unique_ptr<gt16_calls> calls(new gt16_calls());
  calls->pins = ctx->pins;
  if (c->n) {
    calls->sites = (gt16_called_site *) ctx->pins->take((size_t) c->n * sizeof(gt16_called_site), &calls->cap);
    if (!calls->sites) return fail(ctx, GT16_ECUDA, "cudaHostAlloc of the result buffer failed");
    CU(cudaMemcpyAsync(calls->sites, c->recs, (size_t) c->n * sizeof(gt16_called_site), cudaMemcpyDeviceToHost, ctx->d2h));
    CU(cudaStreamSynchronize(ctx->d2h));
    calls->n = c->n;
  }
  *out = calls.release();
  return GT16_OK;
}

int gt16_call_sites_host(gt16_ctx *ctx, const gt16_snapshot_view *view, const gt16_region *regions,
                         uint32_t n_regions, uint64_t site_begin, uint64_t site_end, gt16_calls **out,
                         gt16_stats *stats) {
  if (!ctx) return GT16_EINVAL;
  int rc = check_view(ctx, view);
  if (rc) return rc;
  return call_with_retry(ctx, view, nullptr, regions, n_regions, site_begin, site_end, out, stats);
}

int gt16_call_sites_host_to(gt16_ctx *ctx, const gt16_snapshot_view *view, const gt16_region *regions, uint32_t n_regions,
                            uint64_t site_begin, uint64_t site_end, gt16_called_site *dst, uint64_t dst_capacity,
                            uint64_t *n_out, gt16_stats *stats) {
  if (!ctx || !n_out || (!dst && dst_capacity)) return GT16_EINVAL;
  int rc = check_view(ctx, view);
  if (rc) return rc;
  static gt16_called_site none;
  ctx->extDst = dst ? dst : &none; ctx->extCap = dst_capacity;
  gt16_calls *calls = nullptr;
  gt16_stats st;
  memset(&st, 0, sizeof(st));
  rc = call_with_retry(ctx, view, nullptr, regions, n_regions, site_begin, site_end, &calls, &st);
  ctx->extDst = nullptr; ctx->extCap = 0;
  if (stats) *stats = st;
  *n_out = st.called;
  const bool tooSmall = calls && calls->tooSmall;
  delete calls;
  if (rc == GT16_OK && tooSmall) return fail(ctx, GT16_EINVAL, "result buffer too small (n_out = records found)");
  return rc;
}

int gt16_snapshot_upload_range(gt16_ctx *ctx, const gt16_snapshot_view *view, const gt16_region *regions, uint32_t n_regions,
                               uint64_t site_begin, uint64_t site_end, gt16_dev_snapshot **out) {
  if (!ctx || !out) return GT16_EINVAL;
  int rc = check_view(ctx, view);
  if (rc) return rc;
  CU(cudaSetDevice(ctx->device));
  const uint32_t L = view->text_length;
  RegionPlan pl;
  if ((rc = make_plan(ctx, L, regions, n_regions, pl))) return rc;
  if (site_end > pl.map.totalSites) site_end = pl.map.totalSites;
  if (site_begin >= site_end) return fail(ctx, GT16_EINVAL, "empty site range");
  uint32_t loPos, hiPos; uint64_t lo, hi;
  touch_range(pl, L, site_begin, site_end, &loPos, &hiPos, &lo, &hi);
  gt16_dev_snapshot *ds = new gt16_dev_snapshot();
  std::vector<uint32_t> ovfBase;
  rc = alloc_snapshot(ctx, view, ds, ovfBase, false, lo, hi);
  if (rc == GT16_OK) rc = upload_aux(ctx, view, ds, ovfBase, ctx->copy, loPos, hiPos);
  if (rc == GT16_OK && !view->counters) rc = upload_sparse_window(ctx, view, ds, ctx->copy);
  for (uint64_t a = lo; view->counters && a < hi && rc == GT16_OK; a += kCopyChunkSites) {
    const uint64_t b = std::min<uint64_t>(a + kCopyChunkSites, hi);
    if (cudaMemcpyAsync((char *) ds->counters + a * 8, view->counters + a, (b - a) * 8, cudaMemcpyHostToDevice, ctx->copy) != cudaSuccess)
      rc = fail(ctx, GT16_ECUDA, "counter upload failed");
  }
  if (rc == GT16_OK) rc = finish_counters(ctx, ds, ctx->copy);
  if (rc == GT16_OK && cudaStreamSynchronize(ctx->copy) != cudaSuccess) rc = fail(ctx, GT16_ECUDA, "upload synchronize failed");
  if (rc == GT16_OK && !view->counters) rc = sparse_upload_errors(ctx);
  if (rc != GT16_OK) { (void) cudaGetLastError(); gt16_snapshot_release(ctx, ds); return rc; }
  *out = ds;
  return GT16_OK;
}

uint64_t gt16_calls_count(const gt16_calls *c) { return c ? c->n : 0; }
const gt16_called_site *gt16_calls_data(const gt16_calls *c) { return c && c->n ? c->sites : nullptr; }
void gt16_calls_free(gt16_calls *c) { delete c; }

}  // extern "C"
