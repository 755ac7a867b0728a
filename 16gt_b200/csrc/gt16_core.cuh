// gt16_core.cuh -- per-site arithmetic of the 16GT calling hot path, written once as
// host/device inline functions.  The CUDA kernels (gt16_kernels.cu) are the only product
// callers; tests/host_emu compiles the same functions with g++ to unit-test them on a
// machine without a GPU (test infrastructure -- libgt16.so exports no CPU path).
//
// Every function cites the reference lines whose behaviour it reproduces.  Arithmetic
// order, operand widths (float vs double) and truncations follow the reference exactly;
// this translation unit must be compiled with -fmad=false (no FMA contraction), because the
// reference binary is built for baseline x86-64 (Makefile:3, no -mfma).
#pragma once
#include <stdint.h>
#include <math.h>
#include <float.h>
#include <string.h>

#if defined(__CUDACC__)
#define GT16_HD __host__ __device__ __forceinline__
#define GT16_HDN inline __host__ __device__
#else
#define GT16_HD inline
#define GT16_HDN inline
#endif

namespace gt16 {

constexpr int kFisherN = 2000;            // FISHER_CACHE_SIZE, likelihood_cache.h:10
constexpr int kYcsqN = 15000;             // YCSQ_CACHE_SIZE,  likelihood_cache.h:11
constexpr uint32_t kBatch = 8388608u;     // 256 x 32768, VariantCaller.cpp:526-528
constexpr int kWindow = 100;              // SNP_META_WINDOW_SIZE, SNP_Meta.h:10
constexpr double kOverflowHandler = 3.7537584144023501e+255;  // FisherExactTest.cpp:99
constexpr uint32_t kLnFactTab = 1u << 21; // entries of the float log-factorial table

// ------------------------------------------------------------------------------------
// Double-precision Fisher exact test used to BUILD the error tables (fisher.cpp:6-112).
// ------------------------------------------------------------------------------------
GT16_HD double lngamm(int z) {  // fisher.cpp:6-23 (Lanczos series, not lgamma)
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
GT16_HD double lnfact(int n) { return n <= 1 ? 0.0 : lngamm(n + 1); }                 // fisher.cpp:27-30
GT16_HD double lnbico(int n, int k) { return lnfact(n) - lnfact(k) - lnfact(n - k); } // fisher.cpp:33-35
GT16_HD double hyper323(int n11, int n1_, int n_1, int n) {                            // fisher.cpp:38-40
  return exp(lnbico(n1_, n11) + lnbico(n - n1_, n_1 - n11) - lnbico(n, n_1));
}

// two-tailed p of [[a,b],[c,d]] (fisher.cpp:42-112); a..d >= 0 here
GT16_HDN double fisher2(int a, int b, int c, int d) {
  const int n11 = a, n1_ = a + b, n_1 = a + c, n = a + b + c + d;
  const int lo = (0 > n1_ + n_1 - n) ? 0 : n1_ + n_1 - n;
  const int hi = (n1_ < n_1) ? n1_ : n_1;
  const double poa = hyper323(n11, n1_, n_1, n);
  const double thr = poa * 1.000000001;
  double two = 0.0;
  double pos = poa;
  int s = n11;
  for (int x = n11 - 1; x >= lo; --x) {         // left tail, multiplicative recurrence :74-86
    double pox = pos * (double) (s * (n + s - n1_ - n_1)) / ((n1_ - x) * (n_1 - x));
    s = x;
    pos = pox;
    if (pox < thr) two += pox;
    if (pox == 0.0) break;                      // every later term is exactly 0
  }
  s = n11;
  pos = poa;
  for (int x = n11 + 1; x <= hi; ++x) {         // right tail :88-99
    double pox = pos * (double) (n1_ - s) * (n_1 - s) / ((n + x - n1_ - n_1) * x);
    s = x;
    pos = pox;
    if (pox < thr) two += pox;
    if (pox == 0.0) break;
  }
  two += poa;
  return two;
}

// column index of the "expected" cell baked into table row i (likelihood_cache.cpp:21-30):
// FLOAT arithmetic, un-fused.
GT16_HD int table_expected(int i, float err) {
#if defined(__CUDA_ARCH__)
  return __float2int_rz(__fadd_rn(__fmul_rn((float) i, err), 0.5f));
#else
  volatile float m = (float) i * err;
  volatile float s = m + 0.5f;
  return (int) s;
#endif
}
GT16_HD double table_entry(int i, int j, float err) { return fisher2(j, i, table_expected(i, err), i); }

// the ten error fractions (FisherExactTest.cpp:61-91), float arithmetic
struct ErrFractions { float e[10]; };
GT16_HD ErrFractions err_fractions(float b_err, float ub_err) {
  ErrFractions r;
  float ps = b_err / 4, pi = ub_err / 2;
  float pb = 3 * ps + 2 * pi;   r.e[0] = 1.0f - pb;   r.e[1] = pb;   // hom base      -> cache1, cache2
  float pn = 4 * ps + pi;       r.e[2] = 1.0f - pn;   r.e[3] = pn;   // hom indel     -> cache3, cache4
  float pbb = 2 * ps + 2 * pi;  r.e[4] = 0.5f - pbb;  r.e[5] = pbb;  // het base/base -> cache5, cache6
  float pbi = 3 * ps + pi;      r.e[6] = 0.5f - pbi;  r.e[7] = pbi;  // het base/indel-> cache7, cache8
  float pii = 4 * ps;           r.e[8] = 0.5f - pii;  r.e[9] = pii;  // het indel/ind -> cache9, cache10
  return r;
}

// chi-square fallback (ycsq.cpp:3-33, likelihood_cache.cpp:57-66)
GT16_HD double chi2_uniform_distance(double a, double b, double c, double d) {
  const double ab = a + b, cd = c + d, ac = a + c, bd = b + d, t = a + b + c + d;
  const double ea = ab * ac / t, eb = ab * bd / t, ec = ac * cd / t, ed = cd * bd / t;
  double dist = 0.;
  double u;
  if (ea > 0) { u = fabs(a - ea) - 0.5; dist += (u * u) / ea; }   // pow(.,2) == exact square
  if (eb > 0) { u = fabs(b - eb) - 0.5; dist += (u * u) / eb; }
  if (ec > 0) { u = fabs(c - ec) - 0.5; dist += (u * u) / ec; }
  if (ed > 0) { u = fabs(d - ed) - 0.5; dist += (u * u) / ed; }
  return dist;
}
GT16_HD double chisqr1(double cv) {  // chisqr(1, cv), ycsq.cpp:22-33
  if (cv < 0) return 0.0;
  const double K = 0.5;
  double p = pow(0.5, K) / tgamma(K) * pow(cv, K - 1) * exp((-0.5) * cv);
  return p > 1.0 ? 1.0 : p;
}

// ------------------------------------------------------------------------------------
// Likelihood of the 16 genotypes for one site (FisherExactTest.cpp:101-289)
// ------------------------------------------------------------------------------------
struct LhTables {
  const double *t[10];   // cache1..cache10 (entries may alias when error fractions coincide)
  const double *ycsq;    // 15000 entries
  double p_err_s;        // Ini_BalanceSubError / 4    (FisherExactTest.cpp:298)
  double p_err_i;        // Ini_UnbalanceSubError / 2  (:299)
};

// test builds (tests/host_emu) may route table reads through a lazily filled table
#ifndef GT16_TABLE_READ
#define GT16_TABLE_READ(T, k, w_total, w_target) ((T).t[k][(w_total) * kFisherN + (w_target)])
#endif

// chi-square fallback of a table read (likelihood_cache.cpp:57-72): deep sites only, kept out of line
GT16_HDN double table_fallback(const LhTables &T, int w_target, int err, int w_total) {
  double dist = chi2_uniform_distance(w_target, w_total, err, w_total);
  int di = (int) (dist * 10 + 0.5);
  if (di >= kYcsqN) di = kYcsqN - 1;
  return T.ycsq[di];
}
GT16_HD double table_lookup(const LhTables &T, int k, int w_target, int err, int w_total) {
  if (w_target >= kFisherN || w_total >= kFisherN) return table_fallback(T, w_target, err, w_total);   // :70-72
  return GT16_TABLE_READ(T, k, w_total, w_target);
}

GT16_HD double scale_back(double ret) {  // checkOverflowAndReturn, FisherExactTest.cpp:101-114
  int e;
  frexp(ret, &e);
  if (ret == .0 || e - 849 < -1500) return 0;
  double down = ret / kOverflowHandler;
  if (down != .0) return down;
  return DBL_MIN;
}

GT16_HD int nonneg(int v) { return v < 0 ? 0 : v; }  // centre element of GenerateIntegerArrayFlanking5

// One genotype class of FisherExactTest.cpp:135-247 in a single code path (the five reference functions differ
// only in their tables, their error rate and hom/het):
//   hom(x)  : T[ka][wt][W[x]] * OH * T[ke][wt][max(0, wt - W[x])]                                        (:135-171)
//   het(x,y): sqrt(T[k1][wh][W[x]]) * sqrt(T[k2][wh][W[y]]) * OH * T[ke][wt][max(0, wt - W[x] - W[y])]    (:173-247)
// with wh = (int)(wt * 0.5f + 0.5f) in float arithmetic (:179).  The expected counts error0/error1 (used by the
// chi-square fallback only) are computed in double from the class error rate: p_err for hom, p_err * 0.5f for het.
struct LhClass { int x, y; bool het; double p_err; int k1, k2, ke; };

GT16_HD double lh_class(const LhTables &T, const uint32_t *W, int wt, const LhClass &c) {
  const double q = c.het ? c.p_err * 0.5f : c.p_err;
  const double full = c.het ? 0.5f : 1.0f;
  const int error0 = (int) (wt * (full - q) + 0.5f);
  const int error1 = (int) (wt * q + 0.5f);
  const int wx = (int) W[c.x], wy = c.het ? (int) W[c.y] : 0;
  const int wa = c.het ? (int) (wt * 0.5f + 0.5f) : wt;
  double first = table_lookup(T, c.k1, wx, error0, wa);
  if (c.het) first = sqrt(first) * sqrt(table_lookup(T, c.k2, wy, error0, wa));        // pow(.,0.5f) (:192)
  const double e = table_lookup(T, c.ke, nonneg(wt - wx - wy), error1, wt);
  return scale_back(first * kOverflowHandler * e);
}

// ref-weight penalty when indel evidence is present (:251-260) and W_total (:262-266)
GT16_HD int penalised_weights(const uint32_t *Win, int ref, uint32_t *W) {
  for (int i = 0; i < 6; ++i) W[i] = Win[i];
  uint32_t indelCount = W[4] + W[5];
  if (indelCount) {
    double tmp = (double) indelCount * (1.375f) + 0.5f;
    tmp = W[ref] - tmp;
    if (tmp <= 0) tmp = 0;
    W[ref] = (uint32_t) tmp;
  }
  int wt = 0;
  for (int i = 0; i < 6; ++i) wt += (int) W[i];
  return wt;
}

// class of genotype g (0..9 = AA,AC,AG,AT,CC,CG,CT,GG,GT,TT; 10..13 = AX,CX,GX,TX; 14 = XX; 15 = XY)  (:268-284)
GT16_HD LhClass genotype_class(const LhTables &T, int g) {
  const double ps = T.p_err_s, pi = T.p_err_i;
  LhClass c;
  if (g < 10) {
    int i = 0, rem = g;                       // (i, j) of the g-th pair in the i <= j enumeration
    while (rem >= 4 - i) { rem -= 4 - i; ++i; }
    c.x = i; c.y = i + rem; c.het = rem != 0;
    c.p_err = c.het ? 2 * ps + 2 * pi : 3 * ps + 2 * pi;
    c.k1 = c.het ? 4 : 0; c.k2 = 4; c.ke = c.het ? 5 : 1;
  } else if (g < 14) {
    c.x = g - 10; c.y = 4; c.het = true; c.p_err = 3 * ps + pi; c.k1 = 6; c.k2 = 8; c.ke = 7;
  } else if (g == 14) {
    c.x = 4; c.y = 4; c.het = false; c.p_err = 4 * ps + pi; c.k1 = 2; c.k2 = 2; c.ke = 3;
  } else {
    c.x = 5; c.y = 4; c.het = true; c.p_err = 4 * ps; c.k1 = 8; c.k2 = 8; c.ke = 9;
  }
  return c;
}

// likelihood of genotype g from the penalised weights
GT16_HDN double genotype_lh_one(const LhTables &T, const uint32_t *W, int wt, int g) {
  return fmin(1.0, lh_class(T, W, wt, genotype_class(T, g)));
}

// W[6] = A,C,G,T,deepest indel,second indel weighted counts; ref in 0..3; lh[16] out (:249-289)
GT16_HDN void genotype_likelihood16(const LhTables &T, const uint32_t *Win, int ref, double *lh) {
  uint32_t W[6];
  int wt = penalised_weights(Win, ref, W);
  for (int g = 0; g < 16; ++g) lh[g] = genotype_lh_one(T, W, wt, g);
}

// ------------------------------------------------------------------------------------
// Posterior + argmax (selectGenotype, FisherExactTest.cpp:317-509)
// ------------------------------------------------------------------------------------
struct Posterior {
  double pD, flh[16], bestD, secondD;
  int genotype;
};
GT16_HD int hom_ref_genotype(int ref) { return ref == 0 ? 0 : ref == 1 ? 4 : ref == 2 ? 7 : 9; }  // SNP_Caller.cpp:400

// best / second-best scan over the 16 posteriors with the reference's tie rules (:383-504)
GT16_HD void argmax_posteriors(Posterior &o) {
  int best; double f0, f1;
  if (o.flh[0] >= o.flh[1]) { best = 0; f0 = o.flh[0]; f1 = o.flh[1]; }
  else { best = 1; f0 = o.flh[1]; f1 = o.flh[0]; }
  for (int g = 2; g < 16; ++g) {
    if (o.flh[g] > f0) { f1 = f0; f0 = o.flh[g]; best = g; }
    else if (o.flh[g] > f1) { f1 = o.flh[g]; }
  }
  o.bestD = f0; o.secondD = f1; o.genotype = best;
}

GT16_HDN void select_genotype(int ref, const double *lh, const double *prior256, const uint32_t *g2i,
                              Posterior &o) {
  double pD = lh[0];
  for (int g = 1; g < 16; ++g) pD = pD + lh[g];               // left-to-right sum :333-348
  o.pD = pD;
  if (pD == 0.0) {                                            // :351-361
    o.genotype = hom_ref_genotype(ref);
    for (int g = 0; g < 16; ++g) o.flh[g] = 0;
    o.bestD = 0; o.secondD = 0;
    return;
  }
  for (int g = 0; g < 16; ++g) o.flh[g] = lh[g] / pD * prior256[(ref << 6) | g2i[g]];
  argmax_posteriors(o);
}

// prior table (createSnpPriorArray, SNPFunctions.cpp:13-70); index ref<<6 | a1<<3 | a2
GT16_HDN void make_prior(double *ary) {
  const double altHom = 0.0005, refHet = 0.001, indelAltHom = 0.00005, indelRefHet = 0.0001, titv = 2.1;
  for (int i = 0; i < 256; ++i) ary[i] = 0;
  for (int t = 0; t < 4; ++t)
    for (int a1 = 0; a1 < 6; ++a1)
      for (int a2 = 0; a2 < 6; ++a2) {
        double v;
        if (a1 == t && a2 == t) v = 1;
        else if (a1 == t || a2 == t) v = (a1 > 3 || a2 > 3) ? indelRefHet : refHet;
        else if (a1 == a2) v = (a1 > 3) ? indelAltHom : altHom;
        else if (a1 > 3 && a2 > 3) v = indelAltHom * indelRefHet;
        else if (a1 > 3 || a2 > 3) v = altHom * indelRefHet;
        else v = altHom * refHet;
        int d1 = a1 - t, d2 = a2 - t;
        if (d1 == 2 || d1 == -2) v *= titv;
        if (d2 == 2 || d2 == -2) v *= titv;
        ary[t << 6 | a1 << 3 | a2] = v;
      }
}

// ------------------------------------------------------------------------------------
// Float Fisher test for strand bias / base-quality bias (SNPFunctions.cpp:82-159).
// lnfactC(n) = lgammaf((float)(n+1)) comes from a table filled by the host's lgammaf (bit-
// identical to the reference's libm call); expf follows glibc's algorithm (sysdeps/ieee754/
// flt-32/e_expf.c, Szabolcs Nagy's exp2f_data: N=32, degree-3 polynomial evaluated in
// double) so that poa -- and with it every "pox < poa" tie decision -- matches bit for bit.
// ------------------------------------------------------------------------------------
GT16_HD uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
GT16_HD double u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
GT16_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

GT16_HD uint64_t exp2f_tab(int i) {
  // asuint64(2^(i/32)) - (i << 47), correctly rounded (generated with 60-digit decimals)
  const uint64_t T[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};
  return T[i];
}

GT16_HDN float expf_glibc(float x) {
  const double InvLn2N = 0x1.71547652b82fep+0 * 32, Shift = 0x1.8p+52;
  const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32,
               C2 = 0x1.62e42ff0c52d6p-1 / 32;
  uint32_t abstop = (f2u(x) >> 20) & 0x7ff;
  if (abstop >= (0x42b00000u >> 20)) {                 // |x| >= 88 or NaN
    if (f2u(x) == 0xff800000u) return 0.0f;
    if (abstop >= (0x7f800000u >> 20)) return x + x;
    if (x > 0x1.62e42ep6f) return INFINITY;
    if (x < -0x1.9fe368p6f) return 0.0f;
    if (x < -0x1.9d1d9ep6f) return 0x1p-149f;           // __math_may_uflowf
  }
  double xd = (double) x;
  double z = InvLn2N * xd;
  double kd = z + Shift;
  uint64_t ki = d2u(kd);
  kd -= Shift;
  double r = z - kd;
  uint64_t t = exp2f_tab((int) (ki % 32));
  t += ki << (52 - 5);
  double s = u2d(t);
  z = C0 * r + C1;
  double r2 = r * r;
  double y = C2 * r + 1;
  y = z * r2 + y;
  y = y * s;
  return (float) y;
}

struct LnFact {
  const float *tab;   // tab[n] = lgammaf((float)(n+1)), n < size  (the Fisher kernel points this at shared memory)
  uint32_t size;
  const float *far;   // continuation of the same table in global memory, n < farSize (may be null)
  uint32_t farSize;
};
GT16_HD float lnfactC(const LnFact &L, int n) {
  if ((uint32_t) n < L.size) return L.tab[n];
  if ((uint32_t) n < L.farSize) return L.far[n];
  return lgammaf((float) (n + 1));  // beyond 2^21: documented tolerance path (depth > 130k)
}
GT16_HD float lnbicoC(const LnFact &L, int n, int k) { return (lnfactC(L, n) - lnfactC(L, k) - lnfactC(L, n - k)); }

GT16_HDN float fisher2tailC(const LnFact &L, int n11, int n1_, int n_1, int n) {  // SNPFunctions.cpp:94-135
  n += n11 + n1_ + n_1;
  n1_ += n11;
  n_1 += n11;
  float poa = expf_glibc(lnbicoC(L, n1_, n11) + lnbicoC(L, n - n1_, n_1 - n11) - lnbicoC(L, n, n_1));
  float pox = poa;
  float two = 0;
  two += poa;
  const float poa_e = poa * 1.000000001f;
  // Once a tail term has underflowed to zero every later term of that tail is zero too and adds
  // nothing, so the walk may stop there -- provided the integer products below cannot wrap (then
  // every factor is a positive finite float); for deeper tables the full walk is kept.
  const bool stopAtZero = n <= 46340;
  int lo = 0;
  if (n1_ + n_1 - n > 0) lo = n1_ + n_1 - n;
  for (int x = n11; x > lo; --x) {
    int num = (int) ((uint32_t) x * (uint32_t) (n + x - n1_ - n_1));
    int den = (int) ((uint32_t) (n1_ - x + 1) * (uint32_t) (n_1 - x + 1));
    pox = pox * (float) num / den;
    two += pox * (pox < poa_e);
    if (stopAtZero && pox == 0.0f) break;
  }
  int hi = n_1;
  if (n1_ < n_1) hi = n1_;
  pox = poa;
  for (int x = n11 + 1; x <= hi; ++x) {
    int den = (int) ((uint32_t) (n + x - n1_ - n_1) * (uint32_t) x);
    pox = pox * (float) (n1_ - (x - 1)) * (n_1 - (x - 1)) / den;
    two += pox * (pox < poa_e);
    if (stopAtZero && pox == 0.0f) break;
  }
  return two;
}
GT16_HD float strand_bias(const LnFact &L, uint32_t F, uint32_t R) {      // :137-139,145-151
  int S = (int) (F + R);
  return fminf(1.0f, fisher2tailC(L, (int) F, (int) R, (int) (S * 0.5), (int) (S * 0.5)));
}
GT16_HD float base_quality_bias(const LnFact &L, uint32_t W, uint32_t F, uint32_t R) {  // :141-143,153-159
  int S = (int) ((F + R) * 4);
  return fminf(1.0f, fisher2tailC(L, (int) W, S, S, S));
}

// ------------------------------------------------------------------------------------
// Snapshot view + site decoding (coreStruct.h:10-67,141-154; VariantCaller.cpp:8-491)
// ------------------------------------------------------------------------------------
struct SnapView {
  const uint64_t *counters;   // SnpCounter[L]: w1 | w2<<16 | data<<32 (little endian)
  const uint32_t *invalid;    // MSB-first bitmap, may be null (sparse snapshots carry none)
  const uint8_t *ovf;         // all regions' SnpOverflowCounter records (32 B each), concatenated
  const uint32_t *ovfBase;    // [nOvfRegions + 1] first record of each region in `ovf`
  const uint8_t *pool;        // memory pool bytes [0, poolBytes)
  const uint32_t *packedRef;  // 16 bases per word, first base in bits 31..30
  uint32_t L;                 // dnaLength == textLength
  uint32_t nOvfRegions;       // numOfCPUThreads of the snapshot (SnapshotHandler.cpp:297)
  uint32_t poolBytes;
  uint32_t refWords;
  uint32_t ovfLen;            // L / nOvfRegions, positions per overflow region (set_ovf_geometry)
  double ovfInv;              // 1.0 / ovfLen
};
inline void set_ovf_geometry(SnapView &V) {
  V.ovfLen = V.nOvfRegions ? V.L / V.nOvfRegions : 0;
  V.ovfInv = V.ovfLen ? 1.0 / (double) V.ovfLen : 0.0;
}

GT16_HD uint32_t ref_base(const SnapView &V, uint32_t p) {            // VariantCaller.cpp:629
  uint32_t w = p >> 4;
  if (w >= V.refWords) return 0;
  return (V.packedRef[w] >> (30 - ((p & 15) << 1))) & 3;
}
GT16_HD bool is_invalid(const SnapView &V, uint32_t p) {              // :39
  if (p >= V.L) return true;
  if (!V.invalid) return false;
  return (V.invalid[p >> 5] >> (31 - (p & 31))) & 1;
}
GT16_HD uint32_t ovf_region(const SnapView &V, uint32_t p) {          // findRegionByPosition: min(p / (L / n), n - 1)
  const uint32_t n = V.nOvfRegions, len = V.ovfLen;
  if (len == 0) return n - 1;
  // p / len without an integer division (the region length is fixed per snapshot): the product with the reciprocal is
  // within one of the quotient, one comparison on either side makes it exact
  uint32_t r = (uint32_t) ((double) p * V.ovfInv);
  if ((uint64_t) r * len > p) --r;
  else if ((uint64_t) (r + 1) * len <= p) ++r;
  return r < n - 1 ? r : n - 1;
}
GT16_HD bool is_overflow_type(uint64_t c) { return (((uint32_t) c & (uint32_t) (c >> 16)) & 0xC000u) == 0xC000u; }

// pointer to the overflow record `idx` of `region`, or null if out of range (reference: UB)
GT16_HD const uint8_t *ovf_record(const SnapView &V, uint32_t region, uint32_t idx) {
  uint32_t base = V.ovfBase[region], end = V.ovfBase[region + 1];
  if (idx >= end - base) return nullptr;
  return V.ovf + (size_t) (base + idx) * 32;
}
GT16_HD uint32_t rd16(const uint8_t *p) { return (uint32_t) p[0] | ((uint32_t) p[1] << 8); }
GT16_HD uint32_t rd32(const uint8_t *p) { return rd16(p) | (rd16(p + 2) << 16); }

struct BaseCounts {
  uint32_t W[4], F[4], R[4];
  uint32_t wsum, fsum, rsum;     // weightedCount / posStrandCount / negStrandCount
  uint8_t gc, g;                 // C-or-G seen, G seen
  uint8_t valid;                 // 0 when invalid bit set or p >= L
  uint8_t bad;                   // overflow index out of range (corrupt snapshot)
};

// base counters of one position (VariantCaller.cpp:39-123)
GT16_HDN void decode_bases(const SnapView &V, uint32_t p, BaseCounts &o) {
  for (int i = 0; i < 4; ++i) { o.W[i] = 0; o.F[i] = 0; o.R[i] = 0; }
  o.wsum = o.fsum = o.rsum = 0; o.gc = o.g = 0; o.valid = 0; o.bad = 0;
  if (is_invalid(V, p)) return;
  o.valid = 1;
  uint64_t c = V.counters[p];
  if (!is_overflow_type(c)) {
    uint32_t w1 = (uint32_t) c & 0xFFFF, w2 = (uint32_t) (c >> 16) & 0xFFFF;
    uint32_t p1 = (uint32_t) (c >> 32) & 0xFF, n1 = (uint32_t) (c >> 40) & 0xFF;
    uint32_t p2 = (uint32_t) (c >> 48) & 0xFF, n2 = (uint32_t) (c >> 56) & 0xFF;
    if (p1 | n1) { uint32_t b = w1 >> 14; o.W[b] = w1 & 0x3FFF; o.F[b] = p1; o.R[b] = n1; }
    if (p2 | n2) { uint32_t b = w2 >> 14; o.W[b] = w2 & 0x3FFF; o.F[b] = p2; o.R[b] = n2; }
    // note: a second slot naming the same base overwrites the first, as in the reference,
    // but the running sums below add both (VariantCaller.cpp:57-59,76-78)
    if (p1 | n1) { o.wsum += w1 & 0x3FFF; o.fsum += p1; o.rsum += n1; uint32_t b = w1 >> 14; if (b == 1 || b == 2) { o.gc = 1; if (b == 2) o.g = 1; } }
    if (p2 | n2) { o.wsum += w2 & 0x3FFF; o.fsum += p2; o.rsum += n2; uint32_t b = w2 >> 14; if (b == 1 || b == 2) { o.gc = 1; if (b == 2) o.g = 1; } }
  } else {
    const uint8_t *r = ovf_record(V, ovf_region(V, p), (uint32_t) (c >> 32));
    if (!r) { o.bad = 1; return; }
    for (int i = 0; i < 4; ++i) {
      uint32_t f = rd16(r + 8 + 2 * i), rv = rd16(r + 16 + 2 * i);
      if (f | rv) {
        o.W[i] = rd16(r + 2 * i); o.F[i] = f; o.R[i] = rv;
        o.wsum += o.W[i]; o.fsum += f; o.rsum += rv;
        if (i == 1 || i == 2) { o.gc = 1; if (i == 2) o.g = 1; }
      }
    }
  }
}

struct IndelPat {
  uint8_t type;     // 'I' / 'D' / 0 = none
  uint8_t len;
  uint16_t hq, lq;
  uint8_t seq[64];  // 2-bit bases, base i in bits (i&3)*2 of byte i>>2 (interpreter.cpp:85-87)
};
struct IndelInfo {
  uint32_t W4, F4, R4, W5, F5, R5;   // iStat.W/F/R[4], [5] (VariantCaller.cpp:403-408)
  uint32_t hqTotal, lqTotal;         // indelHqCount / indelLqCount
  uint8_t isValidPlus;               // 0: no pattern, 1: one pattern (isValid 2), 2: two (isValid 3)
  uint8_t bad;
};

// running "deepest / second deepest" selection of VariantCaller.cpp:166-209 etc.
struct PatRef { uint8_t type, len; uint16_t hq, lq; const uint8_t *seq; uint32_t delPos; };

GT16_HD void consider_pattern(PatRef &d0, PatRef &d1, const PatRef &p) {
  uint32_t depth = (uint32_t) p.hq + p.lq;
  if (depth > (uint32_t) d0.hq + d0.lq) { d1 = d0; d0 = p; }
  else if (depth > (uint32_t) d1.hq + d1.lq) { d1 = p; }
}

GT16_HDN void materialise(const SnapView &V, const PatRef &p, IndelPat *o) {
  o->type = p.type; o->len = p.len; o->hq = p.hq; o->lq = p.lq;
  uint32_t nb = ((uint32_t) p.len + 3) / 4;
  for (uint32_t i = 0; i < 64; ++i) o->seq[i] = 0;
  if (p.type == 'I') {
    for (uint32_t i = 0; i < nb; ++i) o->seq[i] = p.seq[i];
  } else {                                     // deleted bases come from the reference (:290-296)
    for (uint32_t i = 0; i < p.len; ++i) o->seq[i >> 2] |= (uint8_t) (ref_base(V, p.delPos + i) << ((i & 3) << 1));
  }
}

// insertion/deletion lists of position+1 (VariantCaller.cpp:125-488); d0 / d1 receive references to
// the deepest / second-deepest pattern (type 0 = none)
GT16_HDN void decode_indel_refs(const SnapView &V, uint32_t p, IndelInfo &o, PatRef &d0, PatRef &d1) {
  o.W4 = o.F4 = o.R4 = o.W5 = o.F5 = o.R5 = 0; o.hqTotal = o.lqTotal = 0; o.isValidPlus = 0; o.bad = 0;
  d0.type = d1.type = 0; d0.len = d1.len = 0; d0.hq = d0.lq = d1.hq = d1.lq = 0; d0.seq = d1.seq = nullptr; d0.delPos = d1.delPos = 0;
  uint32_t q = p + 1;
  if (q >= V.L) q = V.L - 1;
  uint64_t c = V.counters[q];
  if (!is_overflow_type(c)) return;
  uint32_t w1 = (uint32_t) c & 0xFFFF, w2 = (uint32_t) (c >> 16) & 0xFFFF;
  if (!((w1 | w2) & 0x2000)) return;
  // quirk kept: the region is that of `p`, not of p+1 (:47,150,268)
  const uint8_t *r = ovf_record(V, ovf_region(V, p), (uint32_t) (c >> 32));
  if (!r) { o.bad = 1; return; }
  PatRef cur;
  if (w1 & 0x2000) {
    if (w1 & 0x1000) {                                       // pooled pattern list :147-210
      uint32_t size = w1 & 0xFFF, ptr = rd32(r + 24), i = 0;
      if ((uint64_t) ptr + size > V.poolBytes) { o.bad = 1; return; }
      const uint8_t *info = V.pool + ptr;
      while (i < size) {
        // an entry is 1 + ceil(len/4) + 4 bytes; one that runs past the list is a corrupt snapshot (the reference
        // would read whatever follows in the pool)
        if (i + 1 + ((uint32_t) info[i] + 3) / 4 + 4 > size) { o.bad = 1; return; }
        cur.type = 'I'; cur.len = info[i++]; cur.seq = info + i; cur.delPos = 0;
        i += ((uint32_t) cur.len + 3) / 4;
        cur.hq = (uint16_t) rd16(info + i); cur.lq = (uint16_t) rd16(info + i + 2); i += 4;
        o.hqTotal += cur.hq; o.lqTotal += cur.lq;
        consider_pattern(d0, d1, cur);
      }
    } else {                                                 // single inline pattern :211-261
      cur.type = 'I'; cur.len = (uint8_t) ((w1 >> 8) & 0xF); cur.hq = (uint16_t) (w1 & 0xFF); cur.lq = r[27];
      cur.seq = r + 24; cur.delPos = 0;
      o.hqTotal += cur.hq; o.lqTotal += cur.lq;
      consider_pattern(d0, d1, cur);
    }
  }
  if (w2 & 0x2000) {
    if (w2 & 0x1000) {                                       // pooled deletion list :265-337
      uint32_t size = w2 & 0xFFF, ptr = rd32(r + 28), i = 0;
      if ((uint64_t) ptr + size > V.poolBytes) { o.bad = 1; return; }
      const uint8_t *info = V.pool + ptr;
      while (i < size) {
        if (i + 5 > size) { o.bad = 1; return; }
        cur.type = 'D'; cur.len = info[i]; cur.hq = (uint16_t) rd16(info + i + 1); cur.lq = (uint16_t) rd16(info + i + 3);
        i += 5; cur.seq = nullptr; cur.delPos = q;
        o.hqTotal += cur.hq; o.lqTotal += cur.lq;
        consider_pattern(d0, d1, cur);
      }
    } else {                                                 // single deletion :338-400
      cur.type = 'D'; cur.len = r[28]; cur.hq = r[29]; cur.lq = r[30]; cur.seq = nullptr; cur.delPos = q;
      o.hqTotal += cur.hq; o.lqTotal += cur.lq;
      consider_pattern(d0, d1, cur);
    }
  }
  o.W4 = (uint32_t) d0.hq * 4; o.F4 = d0.hq; o.R4 = d0.lq;
  o.W5 = (uint32_t) d1.hq * 4; o.F5 = d1.hq; o.R5 = d1.lq;
  if (d1.hq | d1.lq) o.isValidPlus = 2;                       // return 3 (:436-478)
  else if (d0.hq | d0.lq) o.isValidPlus = 1;                  // return 2 (:481-487)
}

// same, materialising the two pattern records when pats != null
GT16_HDN void decode_indels(const SnapView &V, uint32_t p, IndelInfo &o, IndelPat *pats /*[2] or null*/) {
  PatRef d0, d1;
  decode_indel_refs(V, p, o, d0, d1);
  if (pats) {
    pats[0].type = pats[1].type = 0;
    if (o.isValidPlus >= 1) materialise(V, d0, &pats[0]);
    if (o.isValidPlus >= 2) materialise(V, d1, &pats[1]);
  }
}

// homopolymer run through p, capped at 65535 (VariantCaller.cpp:659-685)
GT16_HDN uint32_t polyrun(const SnapView &V, uint32_t p) {
  uint32_t rb = ref_base(V, p), run = 1, t = p;
  while (t > 0 && ref_base(V, t - 1) == rb) { if (run < 0xFFFF) ++run; --t; }
  t = p + 1;
  while (t < V.L && ref_base(V, t) == rb) { if (run < 0xFFFF) ++run; ++t; }
  return run;
}

// ------------------------------------------------------------------------------------
// Processed-order <-> genome-position map (region list of snapshotSnpcaller -e, or [0,L))
// ------------------------------------------------------------------------------------
struct SiteMap {
  const uint32_t *regStart;     // [nRegions] start position of each region
  const uint64_t *regFirst;     // [nRegions + 1] processed index of each region's first site
  uint32_t nRegions;
  const uint32_t *tileReg;      // cursor for the streaming kernel, may be null: [ceil(total / 256) + 1] region that holds
                                // processed site 256 t (the last entry repeats the last region)
};
constexpr uint32_t kCursorTile = 256;
GT16_HD uint32_t pos_of(const SiteMap &M, uint64_t i) {
  uint32_t lo = 0, hi = M.nRegions;              // last region with regFirst <= i
  while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (M.regFirst[mid] <= i) lo = mid; else hi = mid; }
  return M.regStart[lo] + (uint32_t) (i - M.regFirst[lo]);
}

// the same through the per-tile cursor: the search is confined to the regions that overlap site i's tile of 256 processed
// sites (one or two for exome targets), not the whole list
GT16_HD uint32_t pos_of_cursor(const SiteMap &M, uint64_t i) {
  const uint64_t t = i / kCursorTile;
  uint32_t lo = M.tileReg[t], hi = M.tileReg[t + 1] + 1;      // last region with regFirst <= i lies in [lo, hi)
  while (hi - lo > 1) { uint32_t mid = (lo + hi) >> 1; if (M.regFirst[mid] <= i) lo = mid; else hi = mid; }
  return M.regStart[lo] + (uint32_t) (i - M.regFirst[lo]);
}

// ------------------------------------------------------------------------------------
// +-100 window features with the reference's indexing (SNP_Caller.cpp:276-367)
// ------------------------------------------------------------------------------------
struct WinEntry { uint32_t sc, w; uint8_t valid, gc, g; };

struct BatchGeom {
  uint64_t first;    // processed index of the batch's first site
  uint32_t n;        // sites in this batch
  uint32_t amb0;     // genome position of its first site
  uint32_t ambLast;  // ... and of its last site
};

// entry k of the reference's mwInfo array for this batch (k in [0, n + 200))
GT16_HDN WinEntry window_entry(const SnapView &V, const SiteMap &M, const BatchGeom &B, uint64_t k) {
  WinEntry e; e.sc = e.w = 0; e.valid = e.gc = e.g = 0;
  uint32_t p;
  if (k < (uint64_t) kWindow) {                                   // left margin, VariantCaller.cpp:713-729
    if (!(B.amb0 > (uint32_t) kWindow)) return e;
    p = B.amb0 - kWindow + (uint32_t) k;
  } else if (k < (uint64_t) kWindow + B.n) {
    p = pos_of(M, B.first + (k - kWindow));
  } else {                                                        // right margin, :731-747
    uint64_t j = k - kWindow - B.n;
    if (j >= (uint64_t) kWindow - 1) return e;                    // only 99 entries are ever filled
    uint64_t pp = (uint64_t) B.ambLast + 1 + j;
    if (pp >= V.L) return e;
    p = (uint32_t) pp;
  }
  BaseCounts bc;
  decode_bases(V, p, bc);
  e.valid = bc.valid; e.sc = bc.fsum + bc.rsum; e.w = bc.wsum; e.gc = bc.gc; e.g = bc.g;
  return e;
}

// error bits raised by the per-site routines
enum : uint32_t { kErrBadSnapshot = 1u, kErrGenotype = 2u, kErrRecordOverflow = 4u };

struct EvalParams {
  LhTables lh;
  const double *prior;      // 256 entries, createSnpPriorArray layout
  const uint32_t *g2i;      // genotype2int[16]
  LnFact lnfact;
};

struct DevSiteMap {          // SiteMap + total number of processed sites
  SiteMap m;
  uint64_t totalSites;
};

GT16_HD uint32_t site_pos(const DevSiteMap &M, uint64_t i) {
  if (M.m.nRegions == 1) return M.m.regStart[0] + (uint32_t) (i - M.m.regFirst[0]);
  return M.m.tileReg ? pos_of_cursor(M.m, i) : pos_of(M.m, i);
}

// exact candidate predicate of one site: getSnpMetrics + computeGenotypeLikelihood +
// selectGenotype + the filter of SNP_Caller.cpp:416-417.  err receives kErr* bits.
GT16_HDN bool site_is_candidate(const SnapView &V, const EvalParams &P, uint32_t p, uint32_t &err) {
  BaseCounts bc;
  decode_bases(V, p, bc);
  if (bc.bad) { err |= kErrBadSnapshot; return false; }
  if (!bc.valid) return false;
  IndelInfo ii;
  decode_indels(V, p, ii, nullptr);
  if (ii.bad) { err |= kErrBadSnapshot; return false; }
  uint32_t W[6] = {bc.W[0], bc.W[1], bc.W[2], bc.W[3], ii.W4, ii.W5};
  int ref = (int) ref_base(V, p);
  double lh[16];
  genotype_likelihood16(P.lh, W, ref, lh);
  Posterior po;
  select_genotype(ref, lh, P.prior, P.g2i, po);
  if (po.genotype > 9 && ii.isValidPlus == 0) err |= kErrGenotype;   // SNP_Caller.cpp:235-239
  return po.genotype != hom_ref_genotype(ref);
}

GT16_HDN void bias_of(const LnFact &L, const BaseCounts &bc, float *sb, float *bqb) {
  for (int b = 0; b < 4; ++b) {
    sb[b] = strand_bias(L, bc.F[b], bc.R[b]);
    bqb[b] = base_quality_bias(L, bc.W[b], bc.F[b], bc.R[b]);
  }
}

GT16_HDN void window_side(const SnapView &V, const SiteMap &M, const BatchGeom &B, uint64_t k0, int &strandSum,
                          int &deltaSum, int &depthSum, uint8_t &gcSum, uint8_t &validSum) {
  WinEntry cur = window_entry(V, M, B, k0);
  for (int t = 0; t < kWindow; ++t) {                                // SNP_Caller.cpp:286-314
    WinEntry nxt = window_entry(V, M, B, k0 + t + 1);
    if (cur.valid) {
      strandSum += (int) cur.sc;
      depthSum += (int) cur.w;
      gcSum = (uint8_t) (gcSum + cur.gc);
      validSum = (uint8_t) (validSum + 1);
      if (nxt.valid) deltaSum += (int) (cur.sc > nxt.sc ? cur.sc - nxt.sc : nxt.sc - cur.sc);
    }
    cur = nxt;
  }
}

}  // namespace gt16

#include "../../include/gt16.h"

namespace gt16 {

// Everything one .txt row needs for the candidate at processed index i (computeMeta +
// outputPossibleSNPs, SNP_Caller.cpp:101-393).
GT16_HDN void build_called_site(const SnapView &V, const DevSiteMap &M, uint64_t i, const EvalParams &P,
                                gt16_called_site &o, uint32_t &err) {
  const uint32_t p = site_pos(M, i);
  BaseCounts bc;
  decode_bases(V, p, bc);
  IndelInfo ii;
  IndelPat pats[2];
  decode_indels(V, p, ii, pats);
  if (bc.bad || ii.bad) err |= kErrBadSnapshot;
  const int ref = (int) ref_base(V, p);
  uint32_t W[6] = {bc.W[0], bc.W[1], bc.W[2], bc.W[3], ii.W4, ii.W5};
  double lh[16];
  genotype_likelihood16(P.lh, W, ref, lh);
  Posterior po;
  select_genotype(ref, lh, P.prior, P.g2i, po);

  o.amb = p; o.ref = (uint8_t) ref; o.genotype = (uint8_t) po.genotype;
  const uint8_t isValid = (uint8_t) (1 + ii.isValidPlus);
  uint8_t isIndel = 0, nextBp = 0;
  if (po.genotype > 9) {                                              // SNP_Caller.cpp:228-239
    if (isValid >= 2) { isIndel = isValid; if (pats[0].type == 'D') nextBp = pats[0].len; }
    else err |= kErrGenotype;
  }
  o.is_indel = isIndel;
  for (int b = 0; b < 4; ++b) { o.W[b] = bc.W[b]; o.F[b] = bc.F[b]; o.R[b] = bc.R[b]; }
  o.W[4] = ii.W4; o.F[4] = ii.F4; o.R[4] = ii.R4; o.W[5] = ii.W5; o.F[5] = ii.F5; o.R[5] = ii.R5;
  for (int g = 0; g < 16; ++g) { o.lh[g] = lh[g]; o.flh[g] = po.flh[g]; }
  o.pD = po.pD; o.bestD = po.bestD; o.secondD = po.secondD;

  BatchGeom B;
  B.first = (i / kBatch) * (uint64_t) kBatch;
  uint64_t left = M.totalSites - B.first;
  B.n = left < kBatch ? (uint32_t) left : kBatch;
  B.amb0 = site_pos(M, B.first);
  B.ambLast = site_pos(M, B.first + B.n - 1);
  const uint32_t idx = (uint32_t) (i - B.first);

  // strand bias / base-quality bias: the site (SNP) or {index-1, index+nextBp, index} (indel) :248-274
  for (int k = 0; k < 3; ++k) for (int b = 0; b < 4; ++b) { o.sb[k][b] = 0.f; o.bqb[k][b] = 0.f; }
  if (isIndel) {
    BaseCounts nb;
    if (idx != 0) { decode_bases(V, site_pos(M, i - 1), nb); bias_of(P.lnfact, nb, o.sb[0], o.bqb[0]); }
    if (idx + nextBp < B.n) { decode_bases(V, site_pos(M, i + nextBp), nb); bias_of(P.lnfact, nb, o.sb[1], o.bqb[1]); }
    bias_of(P.lnfact, bc, o.sb[2], o.bqb[2]);
  } else {
    bias_of(P.lnfact, bc, o.sb[0], o.bqb[0]);
  }

  // window features with the reference's (un-offset) mwInfo indexing :276-367
  int strandSum[2] = {0, 0}, deltaSum[2] = {0, 0}, depthSum[2] = {0, 0};
  uint8_t gcSum[2] = {0, 0}, gSum[2] = {0, 0}, validSum[2] = {0, 0};
  float avgSC[2] = {0.f, 0.f}, avgD[2] = {0.f, 0.f};
  if (idx >= (uint32_t) kWindow)
    window_side(V, M.m, B, (uint64_t) idx - kWindow, strandSum[0], deltaSum[0], depthSum[0], gcSum[0], validSum[0]);
  window_side(V, M.m, B, (uint64_t) idx + nextBp, strandSum[1], deltaSum[1], depthSum[1], gcSum[1], validSum[1]);
  if (deltaSum[0] | deltaSum[1]) o.rpb = (deltaSum[0] - deltaSum[1] + 0.0f) / (deltaSum[0] + deltaSum[1]);
  else o.rpb = 0.0f;
  if (validSum[0]) { o.lgc = (gcSum[0] + 0.0f) / validSum[0]; avgSC[0] = (strandSum[0] + 0.0f) / validSum[0]; avgD[0] = (depthSum[0] + 0.0f) / validSum[0]; }
  else o.lgc = 0.0f;
  if (validSum[1]) { o.rgc = (gcSum[1] + 0.0f) / validSum[1]; avgSC[1] = (strandSum[1] + 0.0f) / validSum[1]; avgD[1] = (depthSum[1] + 0.0f) / validSum[1]; }
  else o.rgc = 0.0f;
  o.lq = 0.f; o.rq = 0.f; o.rq_assigned = 0;
  if (fabsf(avgSC[0] + avgSC[1]) > FLT_EPSILON) {
    o.asc = (avgSC[0] - avgSC[1]) / (avgSC[0] + avgSC[1]);
    o.lq = fabsf(avgSC[0]) > FLT_EPSILON ? avgD[0] / avgSC[0] : 0.0f;
    o.rq = fabsf(avgSC[1]) > FLT_EPSILON ? avgD[1] / avgSC[1] : 0.0f;
    o.rq_assigned = 1;
  } else {
    o.asc = 0.0f;
  }
  if (idx >= 3)
    for (uint32_t k = idx - 3; k < idx; ++k) { WinEntry e = window_entry(V, M.m, B, k); if (e.valid) gSum[0] = (uint8_t) (gSum[0] + e.g); }
  for (uint64_t k = (uint64_t) idx + nextBp; k < (uint64_t) idx + nextBp + 3; ++k) { WinEntry e = window_entry(V, M.m, B, k); if (e.valid) gSum[1] = (uint8_t) (gSum[1] + e.g); }
  o.lg = gSum[0]; o.rg = gSum[1];

  o.polyrun = (uint16_t) polyrun(V, p);
  o.total_hq = ii.hqTotal; o.total_lq = ii.lqTotal;
  for (int k = 0; k < 2; ++k) {
    gt16_indel &d = o.indel[k];
    bool have = isIndel && (k == 0 || isIndel == 3);
    d.type = have ? pats[k].type : (uint8_t) '*';
    d.len = have ? pats[k].len : 0; d.hq = have ? pats[k].hq : 0; d.lq = have ? pats[k].lq : 0;
    for (int b = 0; b < 64; ++b) d.seq[b] = have ? pats[k].seq[b] : 0;
  }
}

}  // namespace gt16
