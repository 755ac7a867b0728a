// TEST INFRASTRUCTURE ONLY.  extern "C" handles onto the reference's operator-level
// functions (FisherExactTest.cpp, likelihood_cache.cpp, fisher.cpp, ycsq.cpp,
// SNPFunctions.cpp compiled -fPIC from /root/reference) so tests can call them via
// ctypes.  No arithmetic lives here.
#include <stdlib.h>
#include <string.h>
#include "SNP_Meta.h"
#include "dependencies.h"
#include "FisherExactTest.h"
#include "likelihood_cache.h"
#include "SNPFunctions.h"
#include "fisher.h"
#include "ycsq.h"

extern "C" {

void *refops_prefill(float b_err, float ub_err) {
  LikelihoodCache *lc = (LikelihoodCache *) malloc(sizeof(LikelihoodCache));
  prefill_likelihood_cache_with_p_err(lc, b_err, ub_err);
  unsigned int g2i[16] = {0, 1, 2, 3, 9, 10, 11, 18, 19, 27, 4, 12, 20, 28, 36, 37};  // VariantCaller.cpp:506-509
  copyGenotypeTableToGPU(g2i);
  double *prior = createSnpPriorArray();
  copySnpPriorToGPU(prior);
  freeSnpPriorArray(prior);
  return lc;
}

void refops_free(void *lc) { free(lc); }

// table k in 1..10 -> pointer to 2000*2000 doubles, k==0 -> ycsq_cache (15000)
const double *refops_table(void *lcv, int k) {
  LikelihoodCache *lc = (LikelihoodCache *) lcv;
  switch (k) {
    case 0: return lc->ycsq_cache;
    case 1: return lc->fc.cache1;  case 2: return lc->fc.cache2;
    case 3: return lc->fc.cache3;  case 4: return lc->fc.cache4;
    case 5: return lc->fc.cache5;  case 6: return lc->fc.cache6;
    case 7: return lc->fc.cache7;  case 8: return lc->fc.cache8;
    case 9: return lc->fc.cache9;  case 10: return lc->fc.cache10;
  }
  return NULL;
}

void refops_prior(double *out256) {
  double *p = createSnpPriorArray();
  memcpy(out256, p, 256 * sizeof(double));
  freeSnpPriorArray(p);
}

// W: n x 6 u32, ref: n bytes; lh: n x 16 doubles; out: flh n x 16, pD/best/second n, genotype n
void refops_call(void *lcv, const unsigned int *W, const unsigned char *ref, unsigned int n, int threads,
                 double b_err, double ub_err,
                 double *lh, double *flh, double *pD, double *bestD, double *secondD, int *genotype) {
  LikelihoodCache *lc = (LikelihoodCache *) lcv;
  InputStat *in = (InputStat *) calloc(n, sizeof(InputStat));
  MetaReference *mr = (MetaReference *) calloc(n, sizeof(MetaReference));
  GenotypeLikelihood *gl = (GenotypeLikelihood *) calloc(n, sizeof(GenotypeLikelihood));
  SnpCallingInfo *info = (SnpCallingInfo *) calloc(n, sizeof(SnpCallingInfo));
  for (unsigned int i = 0; i < n; ++i) {
    for (int k = 0; k < 6; ++k) in[i].W[k] = W[i * 6 + k];
    mr[i].refChar = ref[i];
  }
  IniParams ini;
  memset(&ini, 0, sizeof(ini));
  ini.Ini_NumOfCpuThreads = threads;
  ini.Ini_BalanceSubError = b_err;
  ini.Ini_UnbalanceSubError = ub_err;
  computeGenotypeLikelihoodWrapper(in, mr, gl, ini, n, lc);
  computeGenotypeWrapper((unsigned char *) ref, gl, info, ini, n);
  for (unsigned int i = 0; i < n; ++i) {
    memcpy(lh + i * 16, gl[i].likelihood, 16 * sizeof(double));
    memcpy(flh + i * 16, info[i].flh, 16 * sizeof(double));
    pD[i] = info[i].pD; bestD[i] = info[i].bestD; secondD[i] = info[i].secondD; genotype[i] = info[i].genotype;
  }
  free(in); free(mr); free(gl); free(info);
}

// W,F,R: n x 4 u32 -> sb, bqb: n x 4 floats
void refops_bias(const unsigned int *W, const unsigned int *F, const unsigned int *R, unsigned int n,
                 float *sb, float *bqb) {
  for (unsigned int i = 0; i < n; ++i) {
    InputStat in;
    memset(&in, 0, sizeof(in));
    for (int k = 0; k < 4; ++k) { in.W[k] = W[i * 4 + k]; in.F[k] = F[i * 4 + k]; in.R[k] = R[i * 4 + k]; }
    StrandBias s; BaseQualityBias b;
    computeStrandBias(&in, &s);
    computeBaseQualityBias(&in, &b);
    memcpy(sb + i * 4, s.bias, 16); memcpy(bqb + i * 4, b.bias, 16);
  }
}

double refops_fisher2(int a, int b, int c, int d) { return fisher2(a, b, c, d); }
double refops_hyper(int a, int b, int c, int d) { return hyper_323(a, b, c, d); }
double refops_chisqr(int dof, double cv) { return chisqr(dof, cv); }
double refops_chi2dist(double a, double b, double c, double d) { return chi2UniformDistance(a, b, c, d); }

}  // extern "C"
