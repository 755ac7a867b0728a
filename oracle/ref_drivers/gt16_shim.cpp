// TEST / INTEGRATION INFRASTRUCTURE.  The shim a reference maintainer would compile INSTEAD of
// FisherExactTest.cpp: the five functions of the reference's vestigial GPU seam
// (FisherExactTest.h:10-23), implemented by calls into libgt16.so.  oracle/build_ref.sh links it with the
// otherwise unmodified reference objects (VariantCaller.o, SNP_Caller.o, SNPFunctions.o, ...) into
// oracle/_ref/ref_caller_gt16: the reference's own selectPossibleSNPs loop, consumer thread and .txt writer,
// with the likelihood / posterior operators running on the GPU.  tests/test_gpu_parity.py checks that its
// output is the reference's.  No arithmetic lives here.
#include <stdio.h>
#include <stdlib.h>
#include "dependencies.h"
#include "FisherExactTest.h"      // the reference's own declarations
#include "gt16.h"

static gt16_ctx *g_ctx = NULL;
static gt16_ctx *ctx() {
  if (!g_ctx && gt16_ctx_create(0, &g_ctx) != GT16_OK) {
    fprintf(stderr, "gt16: %s\n", gt16_last_error(NULL));
    exit(1);
  }
  return g_ctx;
}
static void must(int rc) {
  if (rc != GT16_OK) { fprintf(stderr, "gt16: %s\n", gt16_last_error(g_ctx)); exit(1); }
}

void copyGenotypeTableToGPU(unsigned int *genotype2int) { must(gt16_copy_genotype_table(ctx(), genotype2int)); }   // FisherExactTest.cpp:19-21
void copySnpPriorToGPU(double *snpPrior) { must(gt16_copy_snp_prior(ctx(), snpPrior)); }                              // :23-25

int prefill_likelihood_cache_with_p_err(LikelihoodCache *lc, float b_err, float ub_err) {                              // :61-92
  (void) lc;     // the tables stay in HBM; pass lc instead of NULL to receive a copy in the reference layout
  must(gt16_prefill_likelihood_cache(ctx(), b_err, ub_err, NULL));
  return 1;
}

void computeGenotypeLikelihoodWrapper(InputStat *in, MetaReference *reference, GenotypeLikelihood *lh, IniParams ini_params,
                                      unsigned int input_size, LikelihoodCache *lc) {                                 // :291-301
  (void) lc;
  must(gt16_compute_genotype_likelihood(ctx(), in[0].W, sizeof(InputStat), &reference[0].refChar, sizeof(MetaReference),
                                        lh[0].likelihood, input_size, ini_params.Ini_BalanceSubError,
                                        ini_params.Ini_UnbalanceSubError));
}

void computeGenotypeWrapper(unsigned char *refChar, GenotypeLikelihood *lh, SnpCallingInfo *info, IniParams ini_params,
                            unsigned int input_size) {                                                                 // :511-521
  (void) ini_params;
  must(gt16_compute_genotype(ctx(), refChar, lh[0].likelihood, info, sizeof(SnpCallingInfo), input_size));
}

// declared in FisherExactTest.h:21 but never called by the caller (SNP_Caller.cpp uses SNPFunctions.cpp's overload)
void computeStrandBias(InputStat *in, StrandBias *sb, IniParams ini_params, unsigned int input_size) {
  (void) in; (void) sb; (void) ini_params; (void) input_size;
  fprintf(stderr, "gt16 shim: computeStrandBias(InputStat*, ...) is not on the calling path\n");
  exit(1);
}
