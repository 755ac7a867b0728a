// TEST INFRASTRUCTURE: compiles the product's host/device per-site routines
// (16gt_b200/csrc/gt16_core.cuh) with g++ so they can be unit-tested against the reference
// objects on a machine without a GPU.  Nothing here is linked into libgt16.so.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <string>
// Table reads go through emu_lazy_entry(): a null table pointer means "evaluate the entry on first use
// with the product's own table builder (table_entry) and memoise it", so the emulation needs no 320 MB
// prefill; a non-null pointer is read directly (LikelihoodCache layout).
static double emu_lazy_entry(const double *tab, int k, int w_total, int w_target);
#define GT16_TABLE_READ(T, k, w_total, w_target) emu_lazy_entry((T).t[k], k, w_total, w_target)
#include "../../16gt_b200/csrc/gt16_core.cuh"

using namespace gt16;

static ErrFractions g_lazyErr;
static std::vector<std::vector<double> > g_lazyRows[10];
static double emu_lazy_entry(const double *tab, int k, int w_total, int w_target) {
  if (tab) return tab[(size_t) w_total * kFisherN + w_target];
  if (g_lazyRows[k].empty()) g_lazyRows[k].resize(kFisherN);
  std::vector<double> &row = g_lazyRows[k][w_total];
  if (row.empty()) row.assign(kFisherN, -1.0);
  if (row[w_target] < 0) row[w_target] = table_entry(w_total, w_target, g_lazyErr.e[k]);
  return row[w_target];
}

static std::vector<float> g_lnfact;
static const LnFact &lnfact_table() {
  static LnFact L;
  if (g_lnfact.empty()) {
    g_lnfact.resize(1u << 16);
    for (uint32_t k = 0; k < g_lnfact.size(); ++k) g_lnfact[k] = lgammaf((float) (k + 1));
    L.tab = g_lnfact.data(); L.size = (uint32_t) g_lnfact.size(); L.far = nullptr; L.farSize = 0;
  }
  return L;
}

static EvalParams make_params(const double *lc, double balance, double unbalance, double *prior, uint32_t *g2i) {
  static const uint32_t kG2I[16] = {0, 1, 2, 3, 9, 10, 11, 18, 19, 27, 4, 12, 20, 28, 36, 37};
  EvalParams P;
  static std::vector<double> ycsq;
  if (!lc) {                                   // lazy tables built by the product's own routines
    ErrFractions e = err_fractions((float) balance, (float) unbalance);
    if (memcmp(&e, &g_lazyErr, sizeof(e)) != 0) { g_lazyErr = e; for (auto &t : g_lazyRows) t.clear(); }
    if (ycsq.empty()) { ycsq.resize(kYcsqN); for (int i = 0; i < kYcsqN; ++i) ycsq[i] = chisqr1(i / 10.0); }
  }
  for (int k = 0; k < 10; ++k) P.lh.t[k] = lc ? lc + (size_t) k * kFisherN * kFisherN : nullptr;
  P.lh.ycsq = lc ? lc + (size_t) 10 * kFisherN * kFisherN : ycsq.data();
  P.lh.p_err_s = balance / 4; P.lh.p_err_i = unbalance / 2;
  make_prior(prior);
  memcpy(g2i, kG2I, sizeof(kG2I));
  P.prior = prior; P.g2i = g2i;
  P.lnfact = lnfact_table();
  return P;
}

extern "C" {

double emu_table_entry(int i, int j, float err) { return table_entry(i, j, err); }
void emu_err_fractions(float b, float ub, float *out10) { ErrFractions e = err_fractions(b, ub); memcpy(out10, e.e, 40); }
double emu_chisqr1(double cv) { return chisqr1(cv); }
float emu_expf(float x) { return expf_glibc(x); }
void emu_prior(double *out) { make_prior(out); }

// lc: LikelihoodCache-layout doubles (10 tables + ycsq)
void emu_call(const double *lc, double balance, double unbalance, const uint32_t *W, const uint8_t *ref, uint32_t n,
              double *lh, double *flh, double *pD, double *bestD, double *secondD, int *genotype) {
  double prior[256]; uint32_t g2i[16];
  EvalParams P = make_params(lc, balance, unbalance, prior, g2i);
  for (uint32_t i = 0; i < n; ++i) {
    genotype_likelihood16(P.lh, W + (size_t) i * 6, ref[i], lh + (size_t) i * 16);
    Posterior po;
    select_genotype(ref[i], lh + (size_t) i * 16, P.prior, P.g2i, po);
    memcpy(flh + (size_t) i * 16, po.flh, 128);
    pD[i] = po.pD; bestD[i] = po.bestD; secondD[i] = po.secondD; genotype[i] = po.genotype;
  }
}

void emu_bias(const uint32_t *W, const uint32_t *F, const uint32_t *R, uint32_t n, float *sb, float *bqb) {
  const LnFact &L = lnfact_table();
  for (uint32_t i = 0; i < n; ++i)
    for (int b = 0; b < 4; ++b) {
      sb[i * 4 + b] = strand_bias(L, F[i * 4 + b], R[i * 4 + b]);
      bqb[i * 4 + b] = base_quality_bias(L, W[i * 4 + b], F[i * 4 + b], R[i * 4 + b]);
    }
}

// Serial emulation of the fused path: every processed site through site_is_candidate, the
// candidates through build_called_site, rows through gt16_write_txt.
int emu_run(const char *indexPrefix, const char *snapshotPath, const char *regionBin, const double *lc,
            double balance, double unbalance, const char *txtOut) {
  char err[512];
  gt16_host_snapshot *hs = nullptr; gt16_index *ix = nullptr;
  if (gt16_snapshot_read(snapshotPath, 0, &hs, err, sizeof(err))) { fprintf(stderr, "%s\n", err); return 1; }
  if (gt16_index_read(indexPrefix, &ix, err, sizeof(err))) { fprintf(stderr, "%s\n", err); return 1; }
  gt16_snapshot_view v;
  gt16_snapshot_view_of(hs, &v);
  uint32_t nWords; const uint32_t *packed = gt16_index_packed(ix, &nWords);
  std::vector<uint8_t> ovf; std::vector<uint32_t> ovfBase(v.n_overflow_regions + 1, 0);
  for (uint32_t r = 0; r < v.n_overflow_regions; ++r) {
    ovfBase[r] = (uint32_t) (ovf.size() / 32);
    ovf.insert(ovf.end(), (const uint8_t *) v.overflow[r], (const uint8_t *) v.overflow[r] + (size_t) v.overflow_size[r] * 32);
  }
  ovfBase[v.n_overflow_regions] = (uint32_t) (ovf.size() / 32);
  ovf.resize(ovf.size() + 32);
  SnapView V;
  V.counters = v.counters; V.invalid = v.invalid_bitmap; V.ovf = ovf.data(); V.ovfBase = ovfBase.data();
  V.pool = v.pool; V.packedRef = packed; V.L = v.text_length; V.nOvfRegions = v.n_overflow_regions;
  V.poolBytes = v.pool_bytes; V.refWords = nWords;
  set_ovf_geometry(V);
  std::vector<uint32_t> rs; std::vector<uint64_t> rf;
  uint64_t acc = 0;
  if (regionBin && *regionBin) {
    FILE *f = fopen(regionBin, "rb"); if (!f) return 1;
    uint32_t n; if (fread(&n, 4, 1, f) != 1) return 1;
    for (uint32_t k = 0; k < n; ++k) { uint32_t se[2]; if (fread(se, 4, 2, f) != 2) return 1; if (se[1] > se[0]) { rs.push_back(se[0]); rf.push_back(acc); acc += se[1] - se[0]; } }
    fclose(f);
  } else { rs.push_back(0); rf.push_back(0); acc = V.L; }
  rf.push_back(acc);
  DevSiteMap M; M.m.regStart = rs.data(); M.m.regFirst = rf.data(); M.m.nRegions = (uint32_t) rs.size(); M.m.tileReg = nullptr; M.totalSites = acc;
  double prior[256]; uint32_t g2i[16];
  EvalParams P = make_params(lc, balance, unbalance, prior, g2i);
  std::vector<gt16_called_site> calls;
  uint32_t e = 0;
  for (uint64_t i = 0; i < acc; ++i) {
    if (site_is_candidate(V, P, site_pos(M, i), e)) {
      gt16_called_site s; memset(&s, 0, sizeof(s));
      build_called_site(V, M, i, P, s, e);
      calls.push_back(s);
    }
  }
  if (e) { fprintf(stderr, "emu_run: error bits %u\n", e); return 2; }
  float carry[2] = {0, 0};
  int rc = gt16_write_txt(txtOut, ix, calls.data(), calls.size(), 0, carry);
  printf("Called Entries : %zu\n", calls.size());
  gt16_snapshot_free(hs); gt16_index_free(ix);
  return rc;
}

}  // extern "C"
