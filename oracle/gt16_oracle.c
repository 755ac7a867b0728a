/* gt16_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, loaded by or called from the product
 * (16gt_b200/lib/libgt16.so, 16gt_b200/bin/snapshotSnpcaller).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline leg may use it, and there only as the checker.
 *
 * A plain-C restatement of the reference's algorithm for the calling hot path of 16GT
 * (aquaskyline/16GT, sources under /root/reference), written to be read next to the reference:
 * it keeps the reference's batch structure (8,388,608-site batches, the mwInfo array with its
 * +100 offset, the per-batch IndelInfo byte stream) so that the reference's indexing quirks fall
 * out of the same loops instead of being re-derived.  Every function cites the file:line it follows.
 *
 * Parity pin: this oracle reproduces, byte for byte, the .txt files that the UNMODIFIED reference
 * objects (oracle/_ref/ref_caller, built by oracle/build_ref.sh from /root/reference) write for the
 * configs of tests/golden_configs.py (tests/golden/<config>.txt.gz), and bit for bit the operator
 * outputs of oracle/_ref/libref16gt_ops.so on random inputs (tests/test_oracle.py).  Identity caveat
 * (SURVEY 8c-7): "the reference" is what g++ 13 / glibc 2.39 make of it in this image -- double fabs
 * in ycsq.cpp:15-18, lgammaf/expf for SNPFunctions.cpp:82-92.
 *
 * The one deliberate difference is cost, not values: the reference fills ten dense 2000x2000 tables
 * up front (~80 s); the oracle evaluates a table entry the first time it is read, with the very same
 * expression the prefill uses (likelihood_cache.cpp:21-30), and memoises it.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define ALPHABET_SIZE 4
#define COUNTER_NUM 6              /* SNP_Meta.h:22 */
#define FISHER_CACHE_SIZE 2000     /* likelihood_cache.h:10 */
#define YCSQ_CACHE_SIZE 15000      /* likelihood_cache.h:11 */
#define WINDOW 100                 /* SNP_META_WINDOW_SIZE, SNP_Meta.h:10 */
#define BATCH_SIZE 8388608u        /* 256 * 32768, VariantCaller.cpp:526-528 */
#define OVERFLOW_HANDLER 3.7537584144023501e+255   /* FisherExactTest.cpp:99 */

/* ------------------------------------------------------------------ fisher.cpp:6-112 (double) */
static double lngamm(int z) {                       /* :6-23, Lanczos series */
  double x = 0;
  x += 0.1659470187408462e-06 / (z + 7);
  x += 0.9934937113930748e-05 / (z + 6);
  x -= 0.1385710331296526 / (z + 5);
  x += 12.50734324009056 / (z + 4);
  x -= 176.6150291498386 / (z + 3);
  x += 771.3234287757674 / (z + 2);
  x -= 1259.139216722289 / (z + 1);
  x += 676.5203681218835 / (z);
  x += 0.9999999999995183;
  return (log(x) - 5.58106146679532777 - z + (z - 0.5) * log(z + 6.5));
}
static double lnfact(int n) { return n <= 1 ? 0 : lngamm(n + 1); }                        /* :27-30 */
static double lnbico(int n, int k) { return lnfact(n) - lnfact(k) - lnfact(n - k); }       /* :33-35 */
static double hyper_323(int n11, int n1_, int n_1, int n) {                                /* :38-40 */
  return exp(lnbico(n1_, n11) + lnbico(n - n1_, n_1 - n11) - lnbico(n, n_1));
}
double oracle_hyper(int a, int b, int c, int d) { return hyper_323(a, b, c, d); }

/* two-tailed p, fisher.cpp:42-112 (the x == s-1 / x == s+1 branches are always taken) */
double oracle_fisher2(int a, int b, int c, int d) {
  if (a < 0) a = -a;
  if (b < 0) b = -b;
  if (c < 0) c = -c;
  if (d < 0) d = -d;
  int n11 = a, n1_ = a + b, n_1 = a + c, n = a + b + c + d;
  int min = (0 > n1_ + n_1 - n) ? 0 : n1_ + n_1 - n;
  int max = (n1_ < n_1) ? n1_ : n_1;
  double poa = hyper_323(n11, n1_, n_1, n);
  double two = 0, pox, pos = poa;
  int s = n11, x;
  for (x = n11 - 1; x >= min; x--) {
    pox = pos * (double) (s * (n + s - n1_ - n_1)) / ((n1_ - x) * (n_1 - x));
    s = x;
    pos = pox;
    if (pox < poa * 1.000000001) two += pox;
  }
  s = n11;
  pos = poa;
  for (x = n11 + 1; x <= max; x++) {
    pox = pos * (double) (n1_ - s) * (n_1 - s) / ((n + x - n1_ - n_1) * x);
    s = x;
    pos = pox;
    if (pox < poa * 1.000000001) two += pox;
  }
  two += poa;
  return two;
}

/* ------------------------------------------------------------------ ycsq.cpp:3-33 */
double oracle_chi2dist(double a, double b, double c, double d) {
  const double ab = a + b, cd = c + d, ac = a + c, bd = b + d, t = a + b + c + d;
  const double ea = ab * ac / t, eb = ab * bd / t, ec = ac * cd / t, ed = cd * bd / t;
  double dist = 0.;
  if (ea > 0) dist += (pow((fabs(a - ea) - 0.5), 2) / ea);      /* abs(double): fabs as built by g++ 13 */
  if (eb > 0) dist += (pow((fabs(b - eb) - 0.5), 2) / eb);
  if (ec > 0) dist += (pow((fabs(c - ec) - 0.5), 2) / ec);
  if (ed > 0) dist += (pow((fabs(d - ed) - 0.5), 2) / ed);
  return dist;
}
double oracle_chisqr(int Dof, double Cv) {
  if (Cv < 0 || Dof < 1) return 0.0;
  double K = ((double) Dof) * 0.5;
  double PValue = pow(0.5, K) / tgamma(K) * pow(Cv, K - 1) * exp((-0.5) * Cv);
  return PValue > 1.0 ? 1.0 : PValue;
}

/* ------------------------------------------------------------------ likelihood_cache.cpp + FisherExactTest.cpp:61-92 */
typedef struct {
  float b_err, ub_err;
  float err[10];                        /* the ten error fractions, in cache1..cache10 order */
  double *row[10][FISHER_CACHE_SIZE];   /* lazily allocated rows; entry < 0 = not yet evaluated */
  double ycsq[YCSQ_CACHE_SIZE];
} Tables;
static Tables *g_tables = NULL;

static Tables *tables_for(float b_err, float ub_err) {
  if (g_tables && g_tables->b_err == b_err && g_tables->ub_err == ub_err) return g_tables;
  if (g_tables) {
    for (int k = 0; k < 10; ++k)
      for (int i = 0; i < FISHER_CACHE_SIZE; ++i) free(g_tables->row[k][i]);
    free(g_tables);
  }
  Tables *T = (Tables *) calloc(1, sizeof(Tables));
  T->b_err = b_err;
  T->ub_err = ub_err;
  /* prefill_likelihood_cache_with_p_err, FisherExactTest.cpp:61-91 -- float arithmetic throughout */
  float P_err_s = b_err / ALPHABET_SIZE;
  float P_err_i = ub_err / 2;
  float P_err_base = 3 * P_err_s + 2 * P_err_i;
  T->err[0] = 1.0f - P_err_base;        T->err[1] = P_err_base;          /* hom base       cache1, cache2 */
  float P_err_indel = 4 * P_err_s + P_err_i;
  T->err[2] = 1.0f - P_err_indel;       T->err[3] = P_err_indel;         /* hom indel      cache3, cache4 */
  float P_err_base_base = 2 * P_err_s + 2 * P_err_i;
  T->err[4] = 0.5f - P_err_base_base;   T->err[5] = P_err_base_base;     /* het base       cache5, cache6 */
  float P_err_base_indel = 3 * P_err_s + P_err_i;
  T->err[6] = 0.5f - P_err_base_indel;  T->err[7] = P_err_base_indel;    /* het base/indel cache7, cache8 */
  float P_err_indel_indel = 4 * P_err_s;
  T->err[8] = 0.5f - P_err_indel_indel; T->err[9] = P_err_indel_indel;   /* het indel/indel cache9, cache10 */
  for (int i = 0; i < YCSQ_CACHE_SIZE; i++) T->ycsq[i] = oracle_chisqr(1, i / 10.0);   /* likelihood_cache.cpp:50-53 */
  g_tables = T;
  return T;
}

/* cache_k[i * 2000 + j] = fisher2(j, i, i * err_k + 0.5f, i): float product, float sum, truncation
 * (likelihood_cache.cpp:21-30).  volatile keeps the two float roundings apart. */
static double table_entry(Tables *T, int k, int i, int j) {
  double *row = T->row[k][i];
  if (!row) {
    row = (double *) malloc(FISHER_CACHE_SIZE * sizeof(double));
    for (int c = 0; c < FISHER_CACHE_SIZE; ++c) row[c] = -1.0;
    T->row[k][i] = row;
  }
  if (row[j] < 0) {
    volatile float prod = (float) i * T->err[k];
    volatile float sum = prod + 0.5f;
    row[j] = oracle_fisher2(j, i, (int) sum, i);
  }
  return row[j];
}
double oracle_table_entry(float b_err, float ub_err, int k /* 1..10 */, int i, int j) {
  return table_entry(tables_for(b_err, ub_err), k - 1, i, j);
}

static double ycsq_p(Tables *T, int a, int b, int c, int d) {                 /* likelihood_cache.cpp:57-66 */
  double dist = oracle_chi2dist(a, b, c, d);
  int dist_int = (int) (dist * 10 + 0.5);
  if (dist_int >= YCSQ_CACHE_SIZE) dist_int = YCSQ_CACHE_SIZE - 1;
  return T->ycsq[dist_int];
}
/* fisher_homo_base ... fisher_het_indel_e_2 (likelihood_cache.cpp:68-234): after a full prefill the
 * "not cached" branch is dead, so an accessor is the chi-square fallback or a table read */
static double cache_read(Tables *T, int k, int w_target, int err, int w_total) {
  if (w_target >= FISHER_CACHE_SIZE || w_total >= FISHER_CACHE_SIZE) return ycsq_p(T, w_target, w_total, err, w_total);
  return table_entry(T, k, w_total, w_target);
}

/* ------------------------------------------------------------------ FisherExactTest.cpp:101-289 */
static double checkOverflowAndReturn(double ret) {                             /* :101-114 */
  int ret_exp;
  frexp(ret, &ret_exp);
  if (ret == .0 || ret_exp - 849 < -1500) return 0;
  double downRet = ret / OVERFLOW_HANDLER;
  if (downRet != .0) return downRet;
  return DBL_MIN;
}
static int flank_centre(int num) { return num < 0 ? 0 : num; }   /* GenerateIntegerArrayFlanking5 :116-133: only ary[5] is used */

static double Likelihood_homo(Tables *T, unsigned int W[], int target, int W_total, double P_err, int k, int ke) {
  /* Likelihood_homo_base :135-152 (k = cache1, ke = cache2), Likelihood_homo_indel :154-171 (cache3, cache4) */
  int error0 = W_total * (1.0f - P_err) + 0.5f;
  int error1 = W_total * P_err + 0.5f;
  double hom = cache_read(T, k, W[target], error0, W_total);
  double hom_e = cache_read(T, ke, flank_centre(W_total - W[target]), error1, W_total);
  double ret = hom * OVERFLOW_HANDLER * hom_e;
  return checkOverflowAndReturn(ret);
}
static double Likelihood_het(Tables *T, unsigned int W[], int t1, int t2, int W_total, double P_err, int k1, int k2, int ke) {
  /* Likelihood_het_base :173-195 (cache5, cache5, cache6); _het_indel_1 :197-221 (cache7 for the base,
   * cache9 for the indel, cache8); _het_indel_2 :223-247 (cache9, cache9, cache10) */
  double P_err_half = P_err * 0.5f;
  int error0 = W_total * (0.5f - P_err_half) + 0.5f;
  int error1 = W_total * P_err_half + 0.5f;
  int W_total_half = W_total * 0.5f + 0.5f;
  double h1 = cache_read(T, k1, W[t1], error0, W_total_half);
  double h2 = cache_read(T, k2, W[t2], error0, W_total_half);
  double he = cache_read(T, ke, flank_centre(W_total - W[t1] - W[t2]), error1, W_total);
  double ret = pow(h1, 0.5f) * pow(h2, 0.5f) * OVERFLOW_HANDLER * he;
  return checkOverflowAndReturn(ret);
}

/* computeGenotypeLikelihood :249-289; W is restored like the reference does */
static void computeGenotypeLikelihood(Tables *T, unsigned int W[], char refChar, double lh[16], double p_err_s, double p_err_i) {
  unsigned int indelCount = W[4] + W[5];
  int idx = (int) refChar;
  unsigned int refWeightBackup = 0;
  if (indelCount) {
    refWeightBackup = W[idx];
    double tmp = (double) indelCount * (1.375f) + 0.5f;
    tmp = W[idx] - tmp;
    if (tmp <= 0) tmp = 0;
    W[idx] = tmp;
  }
  int W_total = 0, i, j, index = 0;
  for (i = 0; i < COUNTER_NUM; ++i) W_total += W[i];
  for (i = 0; i < ALPHABET_SIZE; i++)
    for (j = i; j < ALPHABET_SIZE; j++) {
      if (i == j) lh[index] = fmin(1.0, Likelihood_homo(T, W, i, W_total, 3 * p_err_s + 2 * p_err_i, 0, 1));
      else lh[index] = fmin(1.0, Likelihood_het(T, W, i, j, W_total, 2 * p_err_s + 2 * p_err_i, 4, 4, 5));
      index++;
    }
  for (i = 0; i < 4; ++i) lh[index++] = fmin(1.0, Likelihood_het(T, W, i, 4, W_total, 3 * p_err_s + p_err_i, 6, 8, 7));
  lh[index++] = fmin(1.0, Likelihood_homo(T, W, 4, W_total, 4 * p_err_s + p_err_i, 2, 3));
  lh[index++] = fmin(1.0, Likelihood_het(T, W, 5, 4, W_total, 4 * p_err_s, 8, 8, 9));
  if (indelCount) W[idx] = refWeightBackup;
}

/* ------------------------------------------------------------------ SNPFunctions.cpp:13-70, VariantCaller.cpp:506-509 */
static const unsigned int genotype2int[16] = {0, 1, 2, 3, 9, 10, 11, 18, 19, 27, 4, 12, 20, 28, 36, 37};
static double g_prior[256];
static int g_prior_ready = 0;
void oracle_prior(double *ary) {
  memset(ary, 0, sizeof(double) * 256);
  for (int t = 0; t < 4; ++t)
    for (int a1 = 0; a1 < 6; ++a1)
      for (int a2 = 0; a2 < 6; ++a2) {
        double *e = &ary[t << 6 | a1 << 3 | a2];
        if (a1 == t && a2 == t) *e = 1;
        else if (a1 == t || a2 == t) *e = (a1 > 3 || a2 > 3) ? 0.0001 : 0.001;
        else if (a1 == a2) *e = (a1 > 3) ? 0.00005 : 0.0005;
        else if (a1 > 3 && a2 > 3) *e = 0.00005 * 0.0001;
        else if (a1 > 3 || a2 > 3) *e = 0.0005 * 0.0001;
        else *e = 0.0005 * 0.001;
        if (abs(a1 - t) == 2) *e *= 2.1;
        if (abs(a2 - t) == 2) *e *= 2.1;
      }
}

typedef struct { double pD; int genotype; double flh[16]; double bestD, secondD; } CallInfo;

/* selectGenotype, FisherExactTest.cpp:317-509 */
static void selectGenotype(unsigned char refChar, const double lh[16], CallInfo *info) {
  if (!g_prior_ready) { oracle_prior(g_prior); g_prior_ready = 1; }
  info->pD = lh[0] + lh[1] + lh[2] + lh[3] + lh[4] + lh[5] + lh[6] + lh[7] + lh[8] + lh[9] + lh[10] + lh[11] +
             lh[12] + lh[13] + lh[14] + lh[15];
  if (info->pD == 0.0) {
    static const int genotypeMap[4][4] = {{0, 1, 2, 3}, {1, 4, 5, 6}, {2, 5, 7, 8}, {3, 6, 8, 9}};
    info->genotype = genotypeMap[refChar][refChar];
    for (int i = 0; i < 16; ++i) info->flh[i] = 0;
    info->bestD = 0;
    info->secondD = 0;
    return;
  }
  for (int g = 0; g < 16; ++g) info->flh[g] = lh[g] / info->pD * g_prior[((refChar) << 6) | (genotype2int[g])];
  int bestI;
  double flh0, flh1;
  if (info->flh[0] >= info->flh[1]) { bestI = 0; flh0 = info->flh[0]; flh1 = info->flh[1]; }
  else { bestI = 1; flh0 = info->flh[1]; flh1 = info->flh[0]; }
  for (int g = 2; g < 16; ++g) {
    if (info->flh[g] > flh0) { flh1 = flh0; flh0 = info->flh[g]; bestI = g; }
    else if (info->flh[g] > flh1) flh1 = info->flh[g];
  }
  info->bestD = flh0;
  info->secondD = flh1;
  info->genotype = bestI;
}

/* operator-level entry for tests: computeGenotypeLikelihoodWrapper + computeGenotypeWrapper (:291-301, :511-521) */
void oracle_call(double balance, double unbalance, const unsigned int *W, const unsigned char *ref, unsigned int n,
                 double *lh, double *flh, double *pD, double *bestD, double *secondD, int *genotype) {
  Tables *T = tables_for((float) balance, (float) unbalance);      /* snapshotSnpcaller.cpp:230-231: float parameters */
  for (unsigned int i = 0; i < n; ++i) {
    unsigned int w[6];
    memcpy(w, W + (size_t) i * 6, sizeof(w));
    computeGenotypeLikelihood(T, w, (char) ref[i], lh + (size_t) i * 16, balance / ALPHABET_SIZE, unbalance / 2);
    CallInfo info;
    selectGenotype(ref[i], lh + (size_t) i * 16, &info);
    memcpy(flh + (size_t) i * 16, info.flh, sizeof(info.flh));
    pD[i] = info.pD; bestD[i] = info.bestD; secondD[i] = info.secondD; genotype[i] = info.genotype;
  }
}

/* ------------------------------------------------------------------ SNPFunctions.cpp:82-159 (float) */
static float lnfactC(int n) { return lgammaf((float) (n + 1)); }      /* lgamma(float) -> lgammaf as built */
static float lnbicoC(int n, int k) { return (lnfactC(n) - lnfactC(k) - lnfactC(n - k)); }
static float hyper_323C(int n11, int n1_, int n_1, int n) {
  return (expf(lnbicoC(n1_, n11) + lnbicoC(n - n1_, n_1 - n11) - lnbicoC(n, n_1)));      /* exp(float) -> expf */
}
static float fisher2tailC(int n11, int n1_, int n_1, int n) {          /* :94-135 */
  n += n11 + n1_ + n_1;
  n1_ += n11;
  n_1 += n11;
  float poa = hyper_323C(n11, n1_, n_1, n);
  int x;
  float pox = poa;
  float p_twotail = 0;
  p_twotail += poa;
  float poa_e = poa * 1.000000001f;
  int min = 0;
  if (n1_ + n_1 - n > 0) min = n1_ + n_1 - n;
  for (x = n11; x > min; --x) {
    pox = pox * (float) (int) ((unsigned) x * (unsigned) (n + x - n1_ - n_1)) / (int) ((unsigned) (n1_ - x + 1) * (unsigned) (n_1 - x + 1));
    p_twotail += pox * (pox < poa_e);
  }
  int max = n_1;
  if (n1_ < n_1) max = n1_;
  pox = poa;
  for (x = n11 + 1; x <= max; ++x) {
    pox = pox * (float) (n1_ - (x - 1)) * (n_1 - (x - 1)) / (int) ((unsigned) (n + x - n1_ - n_1) * (unsigned) x);
    p_twotail += pox * (pox < poa_e);
  }
  return p_twotail;
}
static void computeStrandBias(const unsigned int F[], const unsigned int R[], float bias[4]) {          /* :145-151 */
  for (int i = 0; i < ALPHABET_SIZE; ++i) {
    int S = F[i] + R[i];
    bias[i] = fmin(1.0f, fisher2tailC(F[i], R[i], S * 0.5, S * 0.5));
  }
}
static void computeBaseQualityBias(const unsigned int W[], const unsigned int F[], const unsigned int R[], float bias[4]) {   /* :153-159 */
  for (int i = 0; i < ALPHABET_SIZE; ++i) {
    int S = (F[i] + R[i]) * 4;
    bias[i] = fmin(1.0f, fisher2tailC(W[i], S, S, S));
  }
}
void oracle_bias(const unsigned int *W, const unsigned int *F, const unsigned int *R, unsigned int n, float *sb, float *bqb) {
  for (unsigned int i = 0; i < n; ++i) {
    computeStrandBias(F + i * 4, R + i * 4, sb + i * 4);
    computeBaseQualityBias(W + i * 4, F + i * 4, R + i * 4, bqb + i * 4);
  }
}

/* ------------------------------------------------------------------ inputs: coreStruct.h, SnapshotHandler.cpp:238-347, readIndex.cpp */
#pragma pack(push, 1)
typedef struct {                 /* SnpCounter, coreStruct.h:10-30 */
  unsigned short weightedCount1, weightedCount2;
  union {
    struct { unsigned char posStrandCount1, negStrandCount1, posStrandCount2, negStrandCount2; } sc;
    unsigned int arrayIndex;
  } data;
} SnpCounter;
#pragma pack(pop)
typedef struct {                 /* SnpOverflowCounter, coreStruct.h:34-67 */
  unsigned short weightedCount[4], posStrandCount[4], negStrandCount[4];
  union { struct { unsigned char insertSeq[3]; unsigned char lqCount; } counters; unsigned int ptr; } insertion;
  union { struct { unsigned char deletionLength, hqCount, lqCount, pad; } counters; unsigned int ptr; } deletion;
} SnpOverflowCounter;

typedef struct {
  unsigned int textLength, numOfCPUThreads;
  SnpCounter *snpCounter;
  SnpOverflowCounter **overflow;     /* per region */
  unsigned int *invalidPos;
  unsigned char *pool;
  unsigned int *packedSeq;
  unsigned int dnaLength;
  /* annotation / translate */
  int nSeq; char (*seqName)[256];
  int nTr; unsigned int *trStart, *trCorr; unsigned short *trChr;
} Genome;

static int read_snapshot(Genome *G, const char *path) {          /* readSnpInfoSnapshot :259-347 */
  FILE *f = fopen(path, "rb");
  if (!f) { fprintf(stderr, "Error opening %s\n", path); return 1; }
  int pageSize; size_t start[5], len[5];
  if (fread(&pageSize, sizeof(int), 1, f) != 1) return 1;
  for (int i = 0; i < 5; ++i) { if (fread(&start[i], sizeof(size_t), 1, f) != 1 || fread(&len[i], sizeof(size_t), 1, f) != 1) return 1; }
  fseek(f, start[1], SEEK_SET);
  char mode = 0;
  if (fread(&mode, 1, 1, f) != 1 || fread(&G->textLength, 4, 1, f) != 1) return 1;
  unsigned int L = G->textLength;
  G->snpCounter = (SnpCounter *) calloc((size_t) L + 1, sizeof(SnpCounter));
  if (mode & 1) { if (fread(G->snpCounter, sizeof(SnpCounter), L, f) != L) return 1; }
  else if (mode & 2) {
    size_t n = len[1] / (sizeof(unsigned int) + sizeof(SnpCounter));
    while (n--) { unsigned int pos; if (fread(&pos, 4, 1, f) != 1) break; if (fread(&G->snpCounter[pos], sizeof(SnpCounter), 1, f) != 1) break; }
  }
  fseek(f, start[2], SEEK_SET);
  if (fread(&G->numOfCPUThreads, 4, 1, f) != 1) return 1;
  G->overflow = (SnpOverflowCounter **) calloc(G->numOfCPUThreads, sizeof(*G->overflow));
  for (unsigned int i = 0; i < G->numOfCPUThreads; ++i) {
    unsigned int size, limit;
    if (fread(&size, 4, 1, f) != 1 || fread(&limit, 4, 1, f) != 1) return 1;
    G->overflow[i] = (SnpOverflowCounter *) calloc((size_t) limit + 1, sizeof(SnpOverflowCounter));
    if (limit == 0) continue;
    if (size && fread(G->overflow[i], sizeof(SnpOverflowCounter), size, f) != size) return 1;
  }
  fseek(f, start[3], SEEK_SET);
  G->invalidPos = (unsigned int *) calloc(((size_t) L + 31) / 32 + 1, 4);
  if (len[3]) { if (fread(G->invalidPos, 4, (L + 31) >> 5, f) != ((L + 31) >> 5)) return 1; }
  fseek(f, start[0], SEEK_SET);
  unsigned int curPtr; size_t capacity;
  if (fread(&curPtr, 4, 1, f) != 1 || fread(&capacity, sizeof(size_t), 1, f) != 1) return 1;
  G->pool = (unsigned char *) calloc((size_t) curPtr + 16, 1);   /* only [0, curPtr) carries data */
  if (curPtr && fread(G->pool, 1, curPtr, f) != curPtr) return 1;
  fclose(f);
  return 0;
}

static int read_index(Genome *G, const char *prefix) {           /* .ann readIndex.cpp:96-115, .tra :24-60, .pac */
  char path[4096], line[4096];
  snprintf(path, sizeof(path), "%s.ann", prefix);
  FILE *f = fopen(path, "r");
  if (!f) return 1;
  unsigned int len, seed;
  if (fscanf(f, "%u %d %u", &len, &G->nSeq, &seed) != 3) return 1;
  G->dnaLength = len;
  if (!fgets(line, sizeof(line), f)) return 1;
  G->seqName = calloc(G->nSeq, 256);
  for (int i = 0; i < G->nSeq; ++i) {
    if (!fgets(line, sizeof(line), f)) return 1;
    char *sp = strchr(line, ' ');
    strncpy(G->seqName[i], sp ? sp + 1 : line, 255);
    G->seqName[i][strcspn(G->seqName[i], "\r\n")] = 0;
    if (!fgets(line, sizeof(line), f)) return 1;
  }
  fclose(f);
  snprintf(path, sizeof(path), "%s.tra", prefix);
  f = fopen(path, "r");
  if (!f) return 1;
  unsigned int tl, grid; int ns, removed;
  if (fscanf(f, "%u %d %d %u", &tl, &ns, &removed, &grid) != 4) return 1;
  for (unsigned int g = 0; g < grid; ++g) { unsigned int v; if (fscanf(f, "%u", &v) != 1) return 1; }
  G->nTr = ns + removed;
  G->trStart = calloc(G->nTr, 4); G->trCorr = calloc(G->nTr, 4); G->trChr = calloc(G->nTr, 2);
  for (int i = 0; i < G->nTr; ++i) if (fscanf(f, "%u %hu %u", &G->trStart[i], &G->trChr[i], &G->trCorr[i]) != 3) return 1;
  fclose(f);
  snprintf(path, sizeof(path), "%s.pac", prefix);
  f = fopen(path, "rb");
  if (!f) return 1;
  size_t nBytes = ((size_t) len + 3) / 4;
  unsigned char *bytes = calloc(nBytes + 8, 1);
  if (fread(bytes, 1, nBytes, f) != nBytes) return 1;
  fclose(f);
  size_t nWords = ((size_t) len + 15) / 16 + 2;
  G->packedSeq = calloc(nWords, 4);                               /* HSP.packedDNA: 16 bases / word, first base in bits 31..30 */
  for (size_t w = 0; w * 4 < nBytes; ++w)
    G->packedSeq[w] = ((unsigned) bytes[w * 4] << 24) | ((unsigned) bytes[w * 4 + 1] << 16) | ((unsigned) bytes[w * 4 + 2] << 8) | bytes[w * 4 + 3];
  if (len % 16) G->packedSeq[len / 16] &= ~0u << (32 - 2 * (len % 16));
  free(bytes);
  return 0;
}
static void free_genome(Genome *G) {
  free(G->snpCounter);
  if (G->overflow) { for (unsigned int i = 0; i < G->numOfCPUThreads; ++i) free(G->overflow[i]); free(G->overflow); }
  free(G->invalidPos); free(G->pool); free(G->packedSeq); free(G->seqName); free(G->trStart); free(G->trCorr); free(G->trChr);
}
/* getChrAndPos (lib.a; HSP.h:273-275): last translate entry starting at or before amb */
static const char *chr_and_pos(const Genome *G, unsigned int amb, unsigned int *chrPos) {
  int lo = 0, hi = G->nTr;
  while (hi - lo > 1) { int mid = (lo + hi) / 2; if (G->trStart[mid] <= amb) lo = mid; else hi = mid; }
  *chrPos = amb - G->trCorr[lo];
  return G->seqName[G->trChr[lo] - 1];
}
static unsigned int findRegionByPosition(unsigned int pos, unsigned int len, unsigned int n) {   /* coreStruct.h:141-154 */
  unsigned int regionLength = len / n;
  if (regionLength == 0) return n - 1;
  unsigned int r = pos / regionLength;
  return r >= n ? n - 1 : r;
}
static unsigned char base_at(const unsigned int *packedSeq, unsigned int p) { return (packedSeq[p >> 4] >> (30 - ((p & 15) << 1))) & 3; }

/* ------------------------------------------------------------------ getSnpMetrics, VariantCaller.cpp:8-491 */
typedef struct { unsigned int W[COUNTER_NUM], F[COUNTER_NUM], R[COUNTER_NUM]; } InputStat;
typedef struct { unsigned char *array; unsigned int size, limit; } IndelInfo;
static void indel_reserve(IndelInfo *ii, unsigned int extra) {
  if (ii->size + extra + 16 > ii->limit) { ii->limit = (ii->limit + extra + 16) * 2; ii->array = realloc(ii->array, ii->limit); }
}
/* one pattern record appended to the byte stream: type, length, ceil(len/4) packed bases, hq u16, lq u16 */
static unsigned int indel_append(IndelInfo *ii, char type, unsigned char length, const unsigned char *seq, unsigned short hq, unsigned short lq) {
  unsigned int at = ii->size, nb = (length + 3) / 4;
  indel_reserve(ii, 6 + nb);
  ii->array[ii->size++] = type;
  ii->array[ii->size++] = length;
  for (unsigned int i = 0; i < nb; ++i) ii->array[ii->size++] = seq[i];
  ii->array[ii->size++] = hq & 0xff; ii->array[ii->size++] = (hq >> 8) & 0xff;
  ii->array[ii->size++] = lq & 0xff; ii->array[ii->size++] = (lq >> 8) & 0xff;
  return at;
}

static char getSnpMetrics(const Genome *G, unsigned int position, InputStat *iStat, unsigned int *weightedCount,
                          unsigned int *posStrandCount, unsigned int *negStrandCount, unsigned char *gcCount,
                          unsigned char *gCount, unsigned int *indelHqCount, unsigned int *indelLqCount, IndelInfo *indelInfo) {
  const unsigned int dnaLength = G->dnaLength;
  memset(iStat, 0, sizeof(InputStat));
  *weightedCount = *posStrandCount = *negStrandCount = 0;
  *gcCount = *gCount = 0;
  *indelHqCount = *indelLqCount = 0;
  if (position >= dnaLength || (G->invalidPos[position >> 5] & (1u << (31 - (position & 31))))) return 0;     /* :39 */
  const SnpCounter *sc = G->snpCounter;
  unsigned int region = findRegionByPosition(position, dnaLength, G->numOfCPUThreads);
  if (((sc[position].weightedCount1 & sc[position].weightedCount2) & 0xC000) != 0xC000) {                     /* compact :49-86 */
    for (int slot = 0; slot < 2; ++slot) {
      unsigned short w = slot ? sc[position].weightedCount2 : sc[position].weightedCount1;
      unsigned char pos = slot ? sc[position].data.sc.posStrandCount2 : sc[position].data.sc.posStrandCount1;
      unsigned char neg = slot ? sc[position].data.sc.negStrandCount2 : sc[position].data.sc.negStrandCount1;
      if (pos | neg) {
        unsigned char base = (w >> 14) & 3;
        iStat->W[base] = w & 0x3FFF; iStat->F[base] = pos; iStat->R[base] = neg;
        *weightedCount += iStat->W[base]; *posStrandCount += iStat->F[base]; *negStrandCount += iStat->R[base];
        if (base == 1 || base == 2) { if (base == 2) *gCount = 1; *gcCount = 1; }
      }
    }
  } else {                                                                                                    /* overflow :100-123 */
    const SnpOverflowCounter *oc = &G->overflow[region][sc[position].data.arrayIndex];
    for (unsigned int i = 0; i < ALPHABET_SIZE; ++i)
      if (oc->posStrandCount[i] | oc->negStrandCount[i]) {
        iStat->W[i] = oc->weightedCount[i]; iStat->F[i] = oc->posStrandCount[i]; iStat->R[i] = oc->negStrandCount[i];
        *weightedCount += iStat->W[i]; *posStrandCount += iStat->F[i]; *negStrandCount += iStat->R[i];
        if (i == 1 || i == 2) { if (i == 2) *gCount = 1; *gcCount = 1; }
      }
  }
  /* insertion / deletion lists recorded at position + 1 (:125-488) */
  unsigned int pn = position + 1;
  pn = pn >= dnaLength ? dnaLength - 1 : pn;
  if (((sc[pn].weightedCount1 & sc[pn].weightedCount2) & 0xC000) != 0xC000) return 1;
  /* quirk kept: `region` is that of `position`, not of position + 1 (:47,150,268) */
  const SnpOverflowCounter *oc = &G->overflow[region][sc[pn].data.arrayIndex];
  unsigned int deepest[3] = {0, 0, 0}, second[3] = {0, 0, 0};
  IndelInfo scratch = {NULL, 0, 0};
  IndelInfo *out = indelInfo ? &scratch : NULL;       /* patterns are staged, then the two survivors are appended deepest first */
  unsigned int at0 = 0, len0 = 0, at1 = 0, len1 = 0;
  unsigned char delSeq[64];
#define CONSIDER(TYPE, LENGTH, SEQ, HQ, LQ)                                                                      \
  do {                                                                                                           \
    *indelHqCount += (HQ); *indelLqCount += (LQ);                                                                \
    if ((unsigned) (HQ) + (LQ) > deepest[1] + deepest[2]) {                                                      \
      if (out) { at1 = at0; len1 = len0; at0 = indel_append(out, TYPE, LENGTH, SEQ, HQ, LQ); len0 = out->size - at0; } \
      second[0] = deepest[0]; second[1] = deepest[1]; second[2] = deepest[2];                                    \
      deepest[0] = (HQ) * 4; deepest[1] = (HQ); deepest[2] = (LQ);                                               \
    } else if ((unsigned) (HQ) + (LQ) > second[1] + second[2]) {                                                 \
      if (out) { at1 = indel_append(out, TYPE, LENGTH, SEQ, HQ, LQ); len1 = out->size - at1; }                   \
      second[0] = (HQ) * 4; second[1] = (HQ); second[2] = (LQ);                                                  \
    }                                                                                                            \
  } while (0)
  if (sc[pn].weightedCount1 & 0x2000) {
    if (sc[pn].weightedCount1 & 0x1000) {                                      /* pooled insertion patterns :147-210 */
      unsigned short arraySize = sc[pn].weightedCount1 & 0xFFF;
      const unsigned char *info = G->pool + oc->insertion.ptr;
      unsigned int k = 0;
      while (k < arraySize) {
        unsigned char length = info[k++];
        const unsigned char *seq = info + k;
        k += (length + 3) / 4;
        unsigned short hq = info[k] | (info[k + 1] << 8), lq = info[k + 2] | (info[k + 3] << 8);
        k += 4;
        CONSIDER('I', length, seq, hq, lq);
      }
    } else {                                                                   /* single inline insertion :211-261 */
      unsigned char length = (sc[pn].weightedCount1 >> 8) & 0xF, hq = sc[pn].weightedCount1 & 0xFF, lq = oc->insertion.counters.lqCount;
      CONSIDER('I', length, oc->insertion.counters.insertSeq, hq, lq);
    }
  }
  if (sc[pn].weightedCount2 & 0x2000) {
    if (sc[pn].weightedCount2 & 0x1000) {                                      /* pooled deletion patterns :265-337 */
      unsigned short arraySize = sc[pn].weightedCount2 & 0xFFF;
      const unsigned char *info = G->pool + oc->deletion.ptr;
      unsigned int k = 0;
      while (k < arraySize) {
        unsigned char length = info[k];
        unsigned short hq = info[k + 1] | (info[k + 2] << 8), lq = info[k + 3] | (info[k + 4] << 8);
        k += 5;
        memset(delSeq, 0, sizeof(delSeq));                                     /* deleted bases come from the reference :290-296 */
        for (unsigned int i = 0; i < length; ++i) delSeq[i >> 2] |= base_at(G->packedSeq, pn + i) << ((i & 3) << 1);
        CONSIDER('D', length, delSeq, hq, lq);
      }
    } else {                                                                   /* single deletion :338-400 */
      unsigned char length = oc->deletion.counters.deletionLength, hq = oc->deletion.counters.hqCount, lq = oc->deletion.counters.lqCount;
      memset(delSeq, 0, sizeof(delSeq));
      for (unsigned int i = 0; i < length; ++i) delSeq[i >> 2] |= base_at(G->packedSeq, pn + i) << ((i & 3) << 1);
      CONSIDER('D', length, delSeq, hq, lq);
    }
  }
#undef CONSIDER
  iStat->W[4] = deepest[0]; iStat->F[4] = deepest[1]; iStat->R[4] = deepest[2];        /* :403-408 */
  iStat->W[5] = second[0]; iStat->F[5] = second[1]; iStat->R[5] = second[2];
  char ret = 1;
  if (second[1] | second[2]) ret = 3;                                                  /* :436-478: deepest first, then second */
  else if (deepest[1] | deepest[2]) ret = 2;                                           /* :481-487 */
  if (indelInfo && ret >= 2) {
    indel_reserve(indelInfo, len0 + len1);
    memcpy(indelInfo->array + indelInfo->size, scratch.array + at0, len0); indelInfo->size += len0;
    if (ret == 3) { memcpy(indelInfo->array + indelInfo->size, scratch.array + at1, len1); indelInfo->size += len1; }
  }
  free(scratch.array);
  return ret;
}

/* ------------------------------------------------------------------ selectPossibleSNPs + filterSNP + computeMeta + printPossibleSNP */
typedef struct {                    /* MetaWindowInfo, SNP_Meta.h */
  unsigned int weightedCount, posStrandCount, negStrandCount, indelHqCount, indelLqCount;
  unsigned short polyrun;
  unsigned char gcStrandCount, gStrandCount, preBase;
  char isValid;
} MetaWindowInfo;

static const char *genotypeString[] = {"AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT", "AX", "CX", "GX", "TX", "XX", "XY", "NN"};
static const char sym[] = {'A', 'C', 'G', 'T', 'X', 'Y', 'N'};

static void jumpIndelInfoIndex(const IndelInfo *ii, unsigned int *idx) {               /* SNP_Caller.cpp jumpIndelInfoIndex */
  unsigned char length = ii->array[*idx + 1];
  *idx += 2 + (length + 3) / 4 + 4;
}
static void printIndel(FILE *f, const IndelInfo *ii, unsigned int *idx, int present) { /* interpreter.cpp:76-91 + SNP_Caller.cpp:547-575 */
  if (!present) { fprintf(f, "*\t*\t*\t*\t"); return; }
  const unsigned char *r = ii->array + *idx;
  char type = r[0]; unsigned char length = r[1]; unsigned int nb = (length + 3) / 4;
  unsigned int hq = r[2 + nb] | (r[3 + nb] << 8), lq = r[4 + nb] | (r[5 + nb] << 8);
  fprintf(f, "%u%c\t", length, type);
  for (unsigned int i = 0; i < length; ++i) fprintf(f, "%c", sym[(r[2 + (i >> 2)] >> ((i & 3) << 1)) & 3]);
  fprintf(f, "\t%u\t%u\t", hq, lq);
  jumpIndelInfoIndex(ii, idx);
}
static void printHeader(FILE *f) {                                                     /* interpreter.cpp:8-31 */
  fputs("Chr\tPos\tRef\tPreBase\t"
        "pAA\tpAC\tpAG\tpAT\tpCC\tpCG\tpCT\tpGG\tpGT\tpTT\tpAX\tpCX\tpGX\tpTX\tpXX\tpXY\t"
        "GT\tOptb\tSuboptb\tpD\t" "rfProb\trfMode\t" "optIn/Del\tPatt\tHQ\tLQ\t" "suboptIn/Del\tPatt\tHQ\tLQ\t"
        "Depth\tA\tC\tG\tT\tA+\tA-\tC+\tC-\tG+\tG-\tT+\tT-\t"
        "AA\tAC\tAG\tAT\tCC\tCG\tCT\tGG\tGT\tTT\tAX\tCX\tGX\tTX\tXX\tXY\t"
        "SB-A\tSB-C\tSB-G\tSB-T\t" "SB1-A\tSB1-C\tSB1-G\tSB1-T\t" "SB2-A\tSB2-C\tSB2-G\tSB2-T\t"
        "BQB-A\tBQB-C\tBQB-G\tBQB-T\t" "BQB1-A\tBQB1-C\tBQB1-G\tBQB1-T\t" "BQB2-A\tBQB2-C\tBQB2-G\tBQB2-T\t"
        "RPB\t" "LGC\tRGC\t" "LG\tRG\t" "ASC\t" "LQ\tRQ\t" "polyrun\t" "totalHQ\ttotalLQ\n", f);
}

typedef struct { unsigned int startPos, endPos; } ExomeRegion;                        /* struct.h:5-13 */

/* the whole path: snapshotSnpcaller.cpp:133-260 minus the index load, which read_index() replaces */
int oracle_run(const char *indexPrefix, const char *snapshotPath, const char *regionBin, double balance, double unbalance,
               const char *txtOut, unsigned long long *calledOut) {
  Genome G;
  memset(&G, 0, sizeof(G));
  if (read_index(&G, indexPrefix)) { fprintf(stderr, "oracle: cannot read index %s\n", indexPrefix); return 1; }
  if (read_snapshot(&G, snapshotPath)) { fprintf(stderr, "oracle: cannot read snapshot %s\n", snapshotPath); return 1; }
  const unsigned int dnaLength = G.dnaLength;
  Tables *T = tables_for((float) balance, (float) unbalance);
  const double p_err_s = balance / ALPHABET_SIZE, p_err_i = unbalance / 2;            /* FisherExactTest.cpp:298-299 */
  ExomeRegion whole = {0, dnaLength}, *regions = &whole;
  unsigned int nRegions = 1;
  if (regionBin && *regionBin) {                                                       /* snapshotSnpcaller.cpp:172-187 */
    FILE *rf = fopen(regionBin, "rb");
    if (!rf || fread(&nRegions, 4, 1, rf) != 1) return 1;
    regions = malloc((size_t) nRegions * sizeof(ExomeRegion));
    if (fread(regions, sizeof(ExomeRegion), nRegions, rf) != nRegions) return 1;
    fclose(rf);
  }
  FILE *out = fopen(txtOut, "w");
  if (!out) return 1;
  printHeader(out);

  /* per-batch buffers (VariantCaller.cpp:531-572), trimmed to what the later stages read */
  const unsigned int batchSize = BATCH_SIZE;
  unsigned int *amb = malloc((size_t) batchSize * 4);
  unsigned char *genotype = malloc(batchSize);
  MetaWindowInfo *mwInfo = calloc((size_t) batchSize + 2 * WINDOW + 512, sizeof(MetaWindowInfo));   /* +512: a deletion near the batch end reads past batch+200 in the reference (UB there, zeros here) */
  IndelInfo indelInfo = {malloc(1 << 20), 0, 1 << 20};
  unsigned long long called = 0;
  float rqLeft = 0.0f, rqRight = 0.0f;       /* ReadQuality is assigned only in one branch (SNP_Caller.cpp:338-353): values persist */
  unsigned int batchIndex = 0;
  InputStat st;
  unsigned int wc, psc, nsc, hqc, lqc; unsigned char gcc, gc;

  for (unsigned int group = 0; group <= nRegions; ++group) {
    unsigned int start = 0, end = 0;
    if (group < nRegions) {
      start = regions[group].startPos; end = regions[group].endPos;
      if (start >= dnaLength || end > dnaLength) {                                    /* VariantCaller.cpp:623-627 */
        fprintf(stderr, "[Error] Positions in exome region are incompatible with the index, please generate the exome list with the correct index.\n");
        return 1;
      }
    }
    unsigned int position = start;
    for (;;) {
      int flush = 0;
      if (group < nRegions && position < end) {
        /* one position of the producer loop (:628-687) */
        char isValid = getSnpMetrics(&G, position, &st, &wc, &psc, &nsc, &gcc, &gc, &hqc, &lqc, &indelInfo);
        unsigned char refBase = base_at(G.packedSeq, position);
        amb[batchIndex] = position;
        MetaWindowInfo *mw = &mwInfo[WINDOW + batchIndex];
        mw->weightedCount = wc; mw->posStrandCount = psc; mw->negStrandCount = nsc; mw->gcStrandCount = gcc; mw->gStrandCount = gc;
        mw->indelHqCount = hqc; mw->indelLqCount = lqc; mw->isValid = isValid;
        unsigned short polyrun = 1;                                                   /* :659-685 */
        unsigned int testPos = position;
        while (testPos > 0 && base_at(G.packedSeq, testPos - 1) == refBase) { if (polyrun < 0xFFFF) polyrun++; testPos--; }
        testPos = position + 1;
        while (testPos < dnaLength && base_at(G.packedSeq, testPos) == refBase) { if (polyrun < 0xFFFF) polyrun++; testPos++; }
        mw->polyrun = polyrun;
        /* operators, fused per site: computeGenotypeLikelihoodWrapper + computeGenotypeWrapper (:700-702) */
        double lh[16]; CallInfo info;
        computeGenotypeLikelihood(T, st.W, (char) refBase, lh, p_err_s, p_err_i);
        selectGenotype(refBase, lh, &info);
        genotype[batchIndex] = (unsigned char) info.genotype;
        batchIndex++;
        position++;
        if (batchIndex == batchSize) flush = 1;
      } else if (group == nRegions && batchIndex > 0) {
        flush = 1;                                                                    /* final partial batch :770-839 */
      } else break;
      if (!flush) continue;

      /* window margins (:712-747) */
      if (amb[0] > WINDOW) {
        unsigned int mwIdx = 0;
        for (unsigned int i = amb[0] - WINDOW; i < amb[0]; ++i, ++mwIdx) {
          char v = getSnpMetrics(&G, i, &st, &wc, &psc, &nsc, &gcc, &gc, &hqc, &lqc, NULL);
          MetaWindowInfo *mw = &mwInfo[mwIdx];
          mw->weightedCount = wc; mw->posStrandCount = psc; mw->negStrandCount = nsc; mw->gcStrandCount = gcc; mw->gStrandCount = gc; mw->isValid = v;
        }
      }
      {
        unsigned int mwIdx = WINDOW + batchIndex;
        for (unsigned int i = amb[batchIndex - 1] + 1; i < dnaLength && i < amb[batchIndex - 1] + WINDOW; ++i, ++mwIdx) {
          char v = getSnpMetrics(&G, i, &st, &wc, &psc, &nsc, &gcc, &gc, &hqc, &lqc, NULL);
          MetaWindowInfo *mw = &mwInfo[mwIdx];
          mw->weightedCount = wc; mw->posStrandCount = psc; mw->negStrandCount = nsc; mw->gcStrandCount = gcc; mw->gStrandCount = gc; mw->isValid = v;
        }
      }
      /* filterSNP (SNP_Caller.cpp:395-438) + computeMeta (:211-393) + printPossibleSNP (interpreter.cpp:107-156) */
      static const unsigned int filterTable[4] = {0, 4, 7, 9};
      unsigned int indelInfoIndex = 0;
      for (unsigned int index = 0; index < batchIndex; ++index) {
        const MetaWindowInfo *mw = &mwInfo[index + WINDOW];
        unsigned char refChar = base_at(G.packedSeq, amb[index]);
        if (mw->isValid && !(filterTable[refChar] == genotype[index] || genotype[index] >= 16)) {
          called++;
          char isIndel = 0; unsigned char nextBp = 0;
          if (genotype[index] > 9 && mw->isValid >= 2) {
            if (indelInfo.array[indelInfoIndex] == 'D') nextBp = indelInfo.array[indelInfoIndex + 1];
            isIndel = mw->isValid;
          } else if (genotype[index] > 9) {
            fprintf(stderr, "Error in selecting Genotype : %u\n", amb[index]);
            return 2;
          }
          /* the site's own counters and operator outputs, recomputed (the reference keeps them in batch arrays) */
          InputStat is; double lh[16]; CallInfo info;
          getSnpMetrics(&G, amb[index], &is, &wc, &psc, &nsc, &gcc, &gc, &hqc, &lqc, NULL);
          computeGenotypeLikelihood(T, is.W, (char) refChar, lh, p_err_s, p_err_i);
          selectGenotype(refChar, lh, &info);
          float sb[3][4], bqb[3][4];
          memset(sb, 0, sizeof(sb)); memset(bqb, 0, sizeof(bqb));
          if (isIndel) {                                                               /* :248-270 */
            InputStat nb;
            if (index != 0) { getSnpMetrics(&G, amb[index - 1], &nb, &wc, &psc, &nsc, &gcc, &gc, &hqc, &lqc, NULL); computeStrandBias(nb.F, nb.R, sb[0]); computeBaseQualityBias(nb.W, nb.F, nb.R, bqb[0]); }
            if (index + nextBp < batchIndex) { getSnpMetrics(&G, amb[index + nextBp], &nb, &wc, &psc, &nsc, &gcc, &gc, &hqc, &lqc, NULL); computeStrandBias(nb.F, nb.R, sb[1]); computeBaseQualityBias(nb.W, nb.F, nb.R, bqb[1]); }
            computeStrandBias(is.F, is.R, sb[2]); computeBaseQualityBias(is.W, is.F, is.R, bqb[2]);
          } else {
            computeStrandBias(is.F, is.R, sb[0]); computeBaseQualityBias(is.W, is.F, is.R, bqb[0]);
          }
          /* window sums over mwInfo with the reference's un-offset indices (:276-367) */
          int strandSum[2] = {0, 0}, deltaSum[2] = {0, 0}, depthSum[2] = {0, 0};
          unsigned char gcSum[2] = {0, 0}, gSum[2] = {0, 0}, validSum[2] = {0, 0};
          float avgSC[2] = {0.0f, 0.0f}, avgD[2] = {0.0f, 0.0f};
          unsigned int i;
          for (i = index - WINDOW; i < index; ++i) {
            if (mwInfo[i].isValid) {
              unsigned int s0 = mwInfo[i].posStrandCount + mwInfo[i].negStrandCount;
              strandSum[0] += s0; depthSum[0] += mwInfo[i].weightedCount; gcSum[0] += mwInfo[i].gcStrandCount; validSum[0] += 1;
              if (mwInfo[i + 1].isValid) { unsigned int s1 = mwInfo[i + 1].posStrandCount + mwInfo[i + 1].negStrandCount; deltaSum[0] += (s0 > s1 ? s0 - s1 : s1 - s0); }
            }
          }
          for (i = index + nextBp; i < index + nextBp + WINDOW; ++i) {
            if (mwInfo[i].isValid) {
              unsigned int s0 = mwInfo[i].posStrandCount + mwInfo[i].negStrandCount;
              strandSum[1] += s0; depthSum[1] += mwInfo[i].weightedCount; gcSum[1] += mwInfo[i].gcStrandCount; validSum[1] += 1;
              if (mwInfo[i + 1].isValid) { unsigned int s1 = mwInfo[i + 1].posStrandCount + mwInfo[i + 1].negStrandCount; deltaSum[1] += (s0 > s1 ? s0 - s1 : s1 - s0); }
            }
          }
          float rpb, lgc, rgc, asc;
          if (deltaSum[0] | deltaSum[1]) rpb = (deltaSum[0] - deltaSum[1] + 0.0f) / (deltaSum[0] + deltaSum[1]); else rpb = 0.0f;
          if (validSum[0]) { lgc = (gcSum[0] + 0.0f) / validSum[0]; avgSC[0] = (strandSum[0] + 0.0f) / validSum[0]; avgD[0] = (depthSum[0] + 0.0f) / validSum[0]; } else lgc = 0.0f;
          if (validSum[1]) { rgc = (gcSum[1] + 0.0f) / validSum[1]; avgSC[1] = (strandSum[1] + 0.0f) / validSum[1]; avgD[1] = (depthSum[1] + 0.0f) / validSum[1]; } else rgc = 0.0f;
          if (fabs(avgSC[0] + avgSC[1]) > FLT_EPSILON) {
            asc = (avgSC[0] - avgSC[1]) / (avgSC[0] + avgSC[1]);
            rqLeft = fabs(avgSC[0]) > FLT_EPSILON ? avgD[0] / avgSC[0] : 0.0f;
            rqRight = fabs(avgSC[1]) > FLT_EPSILON ? avgD[1] / avgSC[1] : 0.0f;
          } else asc = 0.0f;
          for (i = index - 3; i < index; ++i) if (mwInfo[i].isValid) gSum[0] += mwInfo[i].gStrandCount;
          for (i = index + nextBp; i < index + nextBp + 3; ++i) if (mwInfo[i].isValid) gSum[1] += mwInfo[i].gStrandCount;

          unsigned int chrPos; const char *chrName = chr_and_pos(&G, amb[index], &chrPos);
          fprintf(out, "%s\t%u\t%c\t", chrName, chrPos, sym[refChar]);
          if (isIndel) fprintf(out, "%c\t", sym[refChar]); else fprintf(out, "*\t");
          for (int g = 0; g < 16; ++g) fprintf(out, "%.4e\t", info.flh[g]);
          fprintf(out, "%s\t", genotypeString[info.genotype]);
          fprintf(out, "%.4e\t%.4e\t%.4e\t", info.bestD, info.secondD, info.pD);
          fprintf(out, "%.4e\t%d\t", 0.0, 0);                                          /* rfProb, rfMode forced 0 (SNP_Caller.cpp:583-584) */
          printIndel(out, &indelInfo, &indelInfoIndex, isIndel >= 2);
          printIndel(out, &indelInfo, &indelInfoIndex, isIndel == 3);
          fprintf(out, "%u\t", is.W[0] + is.W[1] + is.W[2] + is.W[3]);
          for (int b = 0; b < 4; ++b) fprintf(out, "%hu\t", (unsigned short) is.W[b]);
          for (int b = 0; b < 4; ++b) fprintf(out, "%hu\t%hu\t", (unsigned short) is.F[b], (unsigned short) is.R[b]);
          for (int g = 0; g < 16; ++g) fprintf(out, "%.4e\t", lh[g]);
          if (isIndel) {
            for (int k = 0; k < 3; ++k) for (int b = 0; b < 4; ++b) fprintf(out, "%.4e\t", sb[k][b]);
            for (int k = 0; k < 3; ++k) for (int b = 0; b < 4; ++b) fprintf(out, "%.4e\t", bqb[k][b]);
          } else {
            for (int b = 0; b < 4; ++b) fprintf(out, "%.4e\t", sb[0][b]);
            fprintf(out, "*\t*\t*\t*\t*\t*\t*\t*\t");
            for (int b = 0; b < 4; ++b) fprintf(out, "%.4e\t", bqb[0][b]);
            fprintf(out, "*\t*\t*\t*\t*\t*\t*\t*\t");
          }
          fprintf(out, "%.4f\t%.4f\t%.4f\t%u\t%u\t%.4f\t%.4f\t%.4f\t%hu\t%u\t%u\n", rpb, lgc, rgc, (unsigned) gSum[0], (unsigned) gSum[1],
                  asc, rqLeft, rqRight, mw->polyrun, mw->indelHqCount, mw->indelLqCount);
          if (!isIndel) {                                                              /* non-indel call: skip this site's patterns (:240-245) */
            if (mw->isValid == 2) jumpIndelInfoIndex(&indelInfo, &indelInfoIndex);
            else if (mw->isValid == 3) { jumpIndelInfoIndex(&indelInfo, &indelInfoIndex); jumpIndelInfoIndex(&indelInfo, &indelInfoIndex); }
          } else if (isIndel == 2) {
            /* one pattern printed, none left */
          }
        } else if (mw->isValid == 2) {
          jumpIndelInfoIndex(&indelInfo, &indelInfoIndex);
        } else if (mw->isValid == 3) {
          jumpIndelInfoIndex(&indelInfo, &indelInfoIndex);
          jumpIndelInfoIndex(&indelInfo, &indelInfoIndex);
        }
      }
      /* next batch (:749-767) */
      memset(mwInfo, 0, ((size_t) batchSize + 2 * WINDOW + 512) * sizeof(MetaWindowInfo));
      indelInfo.size = 0;
      batchIndex = 0;
      if (group == nRegions) break;
    }
  }
  fclose(out);
  if (calledOut) *calledOut = called;
  free(amb); free(genotype); free(mwInfo); free(indelInfo.array);
  if (regions != &whole) free(regions);
  free_genome(&G);
  return 0;
}
