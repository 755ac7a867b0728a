// txt2vcf -- native replacement for the reference's txt2vcf.pl:
//   txt2vcf <txt Input> <Sample Name> <Reference FASTA>  >  sample.vcf
// Reads the .txt written by snapshotSnpcaller (98 tab-separated columns, interpreter.cpp:107-156) and writes the VCF
// txt2vcf.pl writes, byte for byte: same header, same row order, same skipped rows, same arithmetic (the Perl script's
// Inline-C helpers qual / phred / phred2 / clog are restated below; everything else is double arithmetic on the PRINTED
// values of the .txt, as in Perl), same number formatting (Perl stringifies numbers with "%.15g", "%d" truncates).
// The Perl script stays the parity reference: tests/test_txt2vcf.py runs both on the golden .txt files.
// It takes minutes per million rows; this takes seconds.
#include <float.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <complex>
#include <map>
#include <algorithm>
#include <string>
#include <thread>
#include <vector>

// ---- the script's Inline-C helpers (txt2vcf.pl:546-600) ----------------------------------------------------------------
static int phred(double num) {
  if (num < 0 || num > 1.01) { fprintf(stderr, "Input for phred < 0 or > 1: %f\n", num); exit(EXIT_FAILURE); }
  if (num > 1.) num = 1.;
  return (int) ((num == 0.) ? (10 * DBL_MAX_10_EXP) : (-10 * log10(num)));
}
static int phred2(double num) {
  if (num < 0) { fprintf(stderr, "Input for phred2 < 0: %f \n", num); exit(EXIT_FAILURE); }
  return (int) ((num == 0.) ? (0) : (10 * log10(num)));
}
static double qual(double subOpt, double opt, double qd_nom, double qd_den) {
  // C semantics: log2(5.f) is the double function applied to 5.0 (C++ would pick the float overload)
  double rt = (-4.343 * log(subOpt / opt)) * (1 - (log2(5.f - qd_nom / qd_den) / log2((double) 5.f)));
  if (rt != rt) fprintf(stderr, "%e, %e, %e, %f, %f, %f\n", rt, subOpt, opt, qd_nom, qd_den, qd_nom / qd_den);
  return rt;
}
static double cmax(double m, double n) { return m > n ? m : n; }
static double cmin(double m, double n) { return m < n ? m : n; }

// ---- Perl semantics ---------------------------------------------------------------------------------------------------
static double num(const std::string &s) { return strtod(s.c_str(), NULL); }      // numification ("*" -> 0)
static std::string pstr(double v) {                                             // stringification of a number
  char b[64];
  if (v == floor(v) && fabs(v) < 1e15) snprintf(b, sizeof(b), "%.0f", v);
  else snprintf(b, sizeof(b), "%.15g", v);
  return b;
}
static long long ipart(double v) { return (long long) v; }                       // "%d" / int(): truncation
static std::string fmt(const char *f, ...) __attribute__((format(printf, 1, 2)));
#include <stdarg.h>
static std::string fmt(const char *f, ...) {
  char b[4096];
  va_list ap; va_start(ap, f); vsnprintf(b, sizeof(b), f, ap); va_end(ap);
  return b;
}
static bool ends_with(const std::string &s, char c) { return !s.empty() && s[s.size() - 1] == c; }

static int rci(char c) { switch (c) { case 'A': return 0; case 'C': return 1; case 'G': return 2; case 'T': return 3; case 'X': return 4; case 'Y': return 5; } return -1; }
static int gtci(const std::string &g) {                                          // %GTci
  static const char *names[] = {"AA", "AC", "AG", "AT", "CC", "CG", "CT", "GG", "GT", "TT", "CA", "GC", "GA", "TG", "TC", "TA",
                                "AX", "XA", "CX", "XC", "GX", "XG", "TX", "XT", "XX", "XY", "YX"};
  static const int idx[] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 1, 5, 2, 8, 6, 3, 10, 10, 11, 11, 12, 12, 13, 13, 14, 15, 15};
  for (size_t k = 0; k < sizeof(idx) / sizeof(idx[0]); ++k) if (g == names[k]) return idx[k];
  return -1;
}
static const int kGTiii[16][2] = {{0, 0}, {0, 1}, {0, 2}, {0, 3}, {1, 1}, {1, 2}, {1, 3}, {2, 2}, {2, 3}, {3, 3}, {0, 4}, {1, 4}, {2, 4}, {3, 4}, {4, 4}, {4, 5}};

struct Field { const char *p; size_t n; };
struct Row {
  Field f[98];
  std::string a(int i) const { return std::string(f[i].p, f[i].n); }            // the text of column i
  double n(int i) const { return f[i].n ? strtod(f[i].p, NULL) : 0.0; }          // ... and its numification (ends at the tab)
  std::string chr, refBase, preRefBase, altGT, indelMode, indelPattern, indelHQ, indelLQ, indelMode2, indelPattern2, indelHQ2, indelLQ2;
  long long pos;
  double probGT[16], optProb, subOptProb;
  double depth[6], strandCount[8], agg[6], strandBias[6], bqb[6];
  double totalHQ, totalLQ, readPosBias, leftGC, rightGC, depthBalance, leftQ, rightQ;
};

static const char *kHeader[] = {
  "##fileformat=VCFv4.1",
  "##FILTER=<ID=LowQual,Description=\"Low quality\">",
  "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">",
  "##FORMAT=<ID=AD,Number=.,Type=Integer,Description=\"Allelic depths for the ref and alt alleles in the order listed\">",
  "##FORMAT=<ID=DP,Number=1,Type=Integer,Description=\"Read depth\">",
  "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Genotype Quality\">",
  "##FORMAT=<ID=PL,Number=G,Type=Integer,Description=\"Normalized, Phred-scaled likelihoods for genotypes as defined in the VCF specification\">",
  "##INFO=<ID=AC,Number=A,Type=Integer,Description=\"Allele count in genotypes, for each ALT allele, in the same order as listed\">",
  "##INFO=<ID=AF,Number=A,Type=Float,Description=\"Allele Frequency, for each ALT allele, in the same order as listed\">",
  "##INFO=<ID=AN,Number=1,Type=Integer,Description=\"Total number of alleles in called genotypes\">",
  "##INFO=<ID=FRAC,Number=1,Type=Float,Description=\"Alternative allele fraction\">",
  "##INFO=<ID=BQFisherPhredRef,Number=1,Type=Float,Description=\"Phred-scaled p-value using Fisher's exact test of ref vs. perfect base qualities\">",
  "##INFO=<ID=BQFisherPhredAlt,Number=1,Type=Float,Description=\"Phred-scaled p-value using Fisher's exact test of alt vs. perfect base qualities\">",
  "##INFO=<ID=BQFisherPhredAlt2,Number=1,Type=Float,Description=\"Phred-scaled p-value using Fisher's exact test of alt2 vs. perfect base qualities\">",
  "##INFO=<ID=BestGT,Number=1,Type=String,Description=\"Best Genotype among the 16 candidate\">",
  "##INFO=<ID=DP,Number=1,Type=Integer,Description=\"Total read depth\">",
  "##INFO=<ID=DP4,Number=4,Type=Integer,Description=\"Read depth of A,C,G,T\">",
  "##INFO=<ID=DP8,Number=8,Type=Integer,Description=\"Read depth of A+,A-,C+,C-,G+,G-,T+,T-\">",
  "##INFO=<ID=INDEL,Number=0,Type=Flag,Description=\"Genotype with Indel at the genome position\">",
  "##INFO=<ID=IndelMeta,Number=8,Type=String,Description=\"Details of the deepest and second deepest Indel of the position. Format: OptMode/OptPattern/OptHQ/OptLQ/SubOptMode/SubOptPattern/SubOptHQ/SubOptLQ\">",
  "##INFO=<ID=ILDP,Number=1,Type=Integer,Description=\"Reads supporting the Indel patterns with average base quality<=13\">",
  "##INFO=<ID=IHDP,Number=1,Type=Integer,Description=\"Reads supporting the Indel patterns with average base quality>13\">",
  "##INFO=<ID=LeftG,Number=1,Type=Integer,Description=\"# of consecutive base G on the left of the Indel\">",
  "##INFO=<ID=RightG,Number=1,Type=Integer,Description=\"# of consecutive base G on the right of the Indel\">",
  "##INFO=<ID=HRun,Number=1,Type=Integer,Description=\"Largest Contiguous Homopolymer Run of Variant Allele In Either Direction\">",
  "##INFO=<ID=QD,Number=1,Type=Float,Description=\"Variant (Confidence/Quality) by Depth\">",
  "##INFO=<ID=SBFisherPhredRef,Number=1,Type=Float,Description=\"Phred-scaled p-value using Fisher's exact test to detect strand bias of ref allele\">",
  "##INFO=<ID=SBFisherPhredAlt,Number=1,Type=Float,Description=\"Phred-scaled p-value using Fisher's exact test to detect strand bias of alt allele\">",
  "##INFO=<ID=SBFisherPhredAlt2,Number=1,Type=Float,Description=\"Phred-scaled p-value using Fisher's exact test to detect strand bias of alt2 allele\">",
  "##INFO=<ID=ReadPosBias,Number=1,Type=Float,Description=\"Read position bias, [-1,1], 0:no bias, positive: bias into left 100bp, negative: bias into right 100bp\">",
  "##INFO=<ID=DepthBias,Number=1,Type=Float,Description=\"Read depth balance between the left and right 100bp, [-1,1], 0: perfect balance, positive: deeper in left 100bp, negative: deeper in right 100bp\">",
  "##INFO=<ID=SOR,Number=1,Type=Float,Description=\"Symmetric Odds Ratio of 2x2 contingency table to detect strand bias\">",
  "##INFO=<ID=LeftGCPercent,Number=1,Type=Integer,Description=\"The percentage of GC content of the 100bp bases on the left\">",
  "##INFO=<ID=RightGCPercent,Number=1,Type=Integer,Description=\"The percentage of GC content of the 100bp bases on the right\">",
  "##INFO=<ID=LeftQuality,Number=1,Type=Integer,Description=\"The approximate average base quality of the 100bp bases on the left\">",
  "##INFO=<ID=RightQuality,Number=1,Type=Integer,Description=\"The approximate average base quality of the 100bp bases on the right\">",
  "##INFO=<ID=OptSuboptRatio,Number=1,Type=Float,Description=\"10*log10((Highest posterior probability of the 16 genotypes)/(Second highest posterior probability of the 16 genotypes))\">",
};

// the part of every INFO string from BestGT on (the leading AC/AF/../BQFisher fields and the SB fields differ by case)
struct Common {
  std::string bestGT, dp4, dp8, indelMeta;
  long long DP;
};
static std::string info_tail(const Row &r, bool withSOR, double QD, double SOR, int optSubopt) {
  std::string s = fmt("QD=%.3f;ReadPosBias=%.5f;DepthBias=%.5f;", QD, r.readPosBias, r.depthBalance);
  if (withSOR) s += fmt("SOR=%.5f;", SOR);
  s += fmt("LeftGCPercent=%lld;RightGCPercent=%lld;LeftQuality=%lld;RightQuality=%lld;OptSuboptRatio=%d;LeftG=%lld;RightG=%lld;HRun=%lld;IHDP=%lld;ILDP=%lld",
           ipart(r.leftGC), ipart(r.rightGC), ipart(r.leftQ), ipart(r.rightQ), optSubopt, ipart(r.n(90)), ipart(r.n(91)),
           ipart(r.n(95)), ipart(r.totalHQ), ipart(r.totalLQ));
  return s;
}

// symmetric odds ratio of strands i (first) and j (second) (txt2vcf.pl:160-166)
static double sor(const Row &r, int i, int j) {
  const double *c = r.strandCount;
  double S = 0.0;
  S += ((c[i * 2] + 1) / (c[i * 2 + 1] + 1)) * ((c[j * 2 + 1] + 1) / (c[j * 2] + 1));
  S += ((c[i * 2 + 1] + 1) / (c[i * 2] + 1)) * ((c[j * 2] + 1) / (c[j * 2 + 1] + 1));
  const double refRatio = cmin(c[i * 2] + 1, c[i * 2 + 1] + 1) / cmax(c[i * 2] + 1, c[i * 2 + 1] + 1);
  const double altRatio = cmin(c[j * 2] + 1, c[j * 2 + 1] + 1) / cmax(c[j * 2] + 1, c[j * 2 + 1] + 1);
  S = S * refRatio / altRatio;
  // The script ends with `$SOR = clog($SOR)`, clog being its Inline-C wrapper of log().  In the reference as it builds
  // here (perl 5.38 + gcc 13, the oracle of this repository) that name binds to C99's COMPLEX logarithm:
  // the call passes SOR in xmm0 and inherits the imaginary part from xmm1, which still holds the divisor of the
  // statement before -- altRatio -- so what gets printed is Re clog(SOR + i altRatio) = log(hypot(SOR, altRatio)).
  // Reproduced through the same libm entry; GT16_SOR_PLAIN=1 prints the log(SOR) the script's author meant.
  static const bool plain = getenv("GT16_SOR_PLAIN") != NULL;
  if (plain) return log(S);
  return std::log(std::complex<double>(S, altRatio)).real();
}

static double safe_div(double a, double b) {
  if (b == 0) { fprintf(stderr, "Illegal division by zero\n"); exit(255); }     // Perl dies here
  return a / b;
}

struct Converter {
  std::map<std::string, std::string> ref;
  double qual_max;
  // one .txt row -> zero, one or two VCF rows appended to `out`; false = fatal (message on stderr), as where Perl dies
  bool row(const char *line, size_t n, long long lineNo, std::string &out) const;
};

#define OUTF(...) do { char ob_[8192]; int on_ = snprintf(ob_, sizeof(ob_), __VA_ARGS__); out.append(ob_, (size_t) (on_ < (int) sizeof(ob_) ? on_ : (int) sizeof(ob_) - 1)); } while (0)

bool Converter::row(const char *line, size_t n, long long lineNo, std::string &out) const {
  Row r;
  {                                                                              // split /\s+/
    int k = 0;
    const char *p = line, *e = line + n;
    while (p < e && k < 98) {
      const char *q = p;
      while (q < e && !(*q == ' ' || *q == '\t' || *q == '\n' || *q == '\r' || *q == '\f' || *q == '\v')) ++q;
      r.f[k].p = p; r.f[k].n = (size_t) (q - p); ++k;
      p = q;
      while (p < e && (*p == ' ' || *p == '\t' || *p == '\n' || *p == '\r' || *p == '\f' || *p == '\v')) ++p;
    }
    for (; k < 98; ++k) { r.f[k].p = ""; r.f[k].n = 0; }
  }
  r.optProb = r.n(21); r.subOptProb = r.n(22);
  if (r.optProb == 0) return true;
  if (r.subOptProb == 0) return true;
  r.chr = r.a(0); r.pos = ipart(r.n(1)); r.refBase = r.a(2); r.preRefBase = r.a(3);
  for (int g = 0; g < 16; ++g) r.probGT[g] = r.n(4 + g);
  r.altGT = r.a(20);
  r.indelMode = r.a(26); r.indelPattern = r.a(27); r.indelHQ = r.a(28); r.indelLQ = r.a(29);
  r.indelMode2 = r.a(30); r.indelPattern2 = r.a(31); r.indelHQ2 = r.a(32); r.indelLQ2 = r.a(33);
  for (int b = 0; b < 4; ++b) r.depth[b] = r.n(35 + b);
  r.depth[4] = r.depth[5] = 0;
  for (int b = 0; b < 8; ++b) r.strandCount[b] = r.n(39 + b);
  for (int b = 0; b < 4; ++b) { r.strandBias[b] = r.n(63 + b); r.bqb[b] = r.n(75 + b); }
  r.readPosBias = r.n(87);
  r.leftGC = r.n(88) * 100; r.rightGC = r.n(89) * 100;
  r.depthBalance = r.n(92);
  r.leftQ = r.n(93) * 10; r.rightQ = r.n(94) * 10;
  r.totalHQ = r.n(96); r.totalLQ = r.n(97);

  const bool isSnp = !r.altGT.empty() && r.altGT.find_first_not_of("ACGT") == std::string::npos;
  const bool isIndel = r.altGT.find('X') != std::string::npos;
  if (!isSnp && !isIndel) { fprintf(stderr, "%lld: Not SNP and Indel for entry: %.*s", lineNo, (int) n, line); return false; }
  for (int b = 0; b < 4; ++b) r.agg[b] = r.strandCount[2 * b] + r.strandCount[2 * b + 1];
  if (isIndel) {
    if (r.indelHQ2 == "*") r.indelHQ2 = "0";
    if (r.indelLQ2 == "*") r.indelLQ2 = "0";
    r.agg[4] = num(r.indelHQ) + num(r.indelLQ);
    r.agg[5] = num(r.indelHQ2) + num(r.indelLQ2);
    r.depth[4] += num(r.indelHQ) * 4; r.depth[5] += num(r.indelHQ2) * 4;
    // strand bias / base-quality bias of the indel alleles: the minimum over the left and right neighbours' values
    r.strandBias[4] = r.strandBias[5] = 1;
    for (int k = 67; k <= 74; ++k) if (r.n(k) < r.strandBias[4]) r.strandBias[4] = r.strandBias[5] = r.n(k);
    r.bqb[4] = r.bqb[5] = 1;
    for (int k = 79; k <= 86; ++k) if (r.n(k) < r.bqb[4]) r.bqb[4] = r.bqb[5] = r.n(k);
  }
  double DP = r.agg[0] + r.agg[1] + r.agg[2] + r.agg[3] + r.totalHQ + r.totalLQ;
  for (int g = 0; g < 16; ++g) r.probGT[g] = r.probGT[g] / r.optProb;

  const int gt = gtci(r.altGT);
  if (gt < 0) { fprintf(stderr, "%lld: unknown genotype %s\n", lineNo, r.altGT.c_str()); return false; }
  // $altBase =~ s/$refBase//; $altBase =~ s/(.*)\1/$1/g;
  std::string alt = r.altGT;
  {
    size_t k = alt.find(r.refBase);
    if (!r.refBase.empty() && k != std::string::npos) alt.erase(k, r.refBase.size());
    if (alt.size() == 2 && alt[0] == alt[1]) alt.erase(1);
  }
  const char *filter = ".";
  const int optSubopt = phred2(r.optProb / r.subOptProb);
  double qualV = qual_max;
  const char refC = r.refBase.empty() ? '?' : r.refBase[0];
  const int rRef = rci(refC);
  const std::string dp4 = pstr(r.agg[0]) + "," + pstr(r.agg[1]) + "," + pstr(r.agg[2]) + "," + pstr(r.agg[3]);
  std::string dp8 = r.a(39);
  for (int b = 1; b < 8; ++b) dp8 += "," + r.a(39 + b);
  const std::string meta = r.indelMode + "," + r.indelPattern + "," + r.indelHQ + "," + r.indelLQ + "," + r.indelMode2 + "," +
                           r.indelPattern2 + "," + r.indelHQ2 + "," + r.indelLQ2;
  auto PLof = [&](const std::string &g) -> int {
    int k = gtci(g);
    if (k < 0) { fprintf(stderr, "%lld: no genotype %s\n", lineNo, g.c_str()); exit(255); }
    return phred(r.probGT[k]);
  };
  auto s1 = [](char c) { return std::string(1, c); };
  auto refMatches = [&](const std::string &pat) -> bool {
    std::map<std::string, std::string>::const_iterator it = ref.find(r.chr);
    if (it == ref.end()) return pat.empty();
    if (r.pos - 1 < 0 || (size_t) (r.pos - 1) > it->second.size()) return pat.empty();
    return it->second.compare((size_t) (r.pos - 1), pat.size(), pat) == 0 && (size_t) (r.pos - 1) + pat.size() <= it->second.size();
  };
  auto emit = [&](const std::string &refPat, const std::string &altPat, const std::string &info, const std::string &format) {
    OUTF("%s\t%lld\t.\t%s\t%s\t%.1f\t%s\t", r.chr.c_str(), r.pos, refPat.c_str(), altPat.c_str(), qualV, filter);
    out += info; out += "\tGT:AD:DP:GQ:PL\t"; out += format; out += '\n';
  };
  auto format_of = [&](const char *GT, const std::string &AD, double DPF, double QD, const std::string &PL) {
    return fmt("%s:%s:%lld:%lld:%s", GT, AD.c_str(), ipart(DPF), ipart(QD + 0.499), PL.c_str());
  };
  auto indel_patterns = [&](std::string &refPat, std::string &altPat) {        // single indel allele
    if (ends_with(r.indelMode, 'I')) { refPat = r.preRefBase; altPat = r.preRefBase + r.indelPattern; }
    else if (ends_with(r.indelMode, 'D')) { refPat = r.preRefBase + r.indelPattern; altPat = r.preRefBase; }
    else { fprintf(stderr, "%lld: Unknown Indel Mode: %s\n", lineNo, r.indelMode.c_str()); exit(255); }
  };

  if (isSnp) {
    if (kGTiii[gt][0] == kGTiii[gt][1]) {                                       // 1/1
      if (alt.size() != 1) { fprintf(stderr, "%lld: scalar (@altBase) != 1", lineNo); return false; }
      const int rAlt = rci(alt[0]);
      const int FSAlt = phred(r.strandBias[rAlt]), BQAlt = phred(r.bqb[rAlt]);
      const double QD = safe_div(r.depth[rAlt], r.agg[rAlt]) * 10;
      qualV = qual(r.subOptProb, r.optProb, r.depth[rAlt], r.agg[rAlt]);
      const double SOR = sor(r, rRef, rAlt);
      const std::string AD = pstr(r.agg[rRef]) + "," + pstr(r.agg[rAlt]);
      const std::string PL = fmt("%d,%d,%d", PLof(s1(refC) + refC), PLof(s1(refC) + alt[0]), PLof(s1(alt[0]) + alt[0]));
      const std::string info = fmt("AC=2;AF=1;AN=2;FRAC=1.000;BQFisherPhredAlt=%d;BestGT=%s;DP=%lld;DP4=%s;DP8=%s;IndelMeta=%s;SBFisherPhredAlt=%d;",
                                   BQAlt, r.altGT.c_str(), ipart(DP), dp4.c_str(), dp8.c_str(), meta.c_str(), FSAlt) +
                               info_tail(r, true, QD, SOR, optSubopt);
      if (!refMatches(r.refBase)) return true;
      emit(r.refBase, alt, info, format_of("1/1", AD, DP, QD, PL));
    } else if (alt.size() == 1) {                                              // 0/1
      const int rAlt = rci(alt[0]);
      const double FRAC = safe_div(r.agg[rAlt], r.agg[rAlt] + r.agg[rRef]);
      const int BQRef = phred(r.bqb[rRef]), BQAlt = phred(r.bqb[rAlt]), FSRef = phred(r.strandBias[rRef]), FSAlt = phred(r.strandBias[rAlt]);
      const double QD = safe_div(r.depth[rAlt] + r.depth[rRef], r.agg[rAlt] + r.agg[rRef]) * 10;
      qualV = qual(r.subOptProb, r.optProb, r.depth[rAlt] + r.depth[rRef], r.agg[rAlt] + r.agg[rRef]);
      const double SOR = sor(r, rRef, rAlt);
      const std::string AD = pstr(r.agg[rRef]) + "," + pstr(r.agg[rAlt]);
      const std::string PL = fmt("%d,%d,%d", PLof(s1(refC) + refC), PLof(s1(refC) + alt[0]), PLof(s1(alt[0]) + alt[0]));
      const std::string info = fmt("AC=1;AF=0.5;AN=2;FRAC=%.3f;BQFisherPhredRef=%d;BQFisherPhredAlt=%d;BestGT=%s;DP=%lld;DP4=%s;DP8=%s;IndelMeta=%s;SBFisherPhredRef=%d;SBFisherPhredAlt=%d;",
                                   FRAC, BQRef, BQAlt, r.altGT.c_str(), ipart(DP), dp4.c_str(), dp8.c_str(), meta.c_str(), FSRef, FSAlt) +
                               info_tail(r, true, QD, SOR, optSubopt);
      if (!refMatches(r.refBase)) return true;
      emit(r.refBase, alt, info, format_of("0/1", AD, DP, QD, PL));
    } else if (alt.size() == 2) {                                              // 1/2
      const int r0 = rci(alt[0]), r1 = rci(alt[1]);
      const int BQAlt = phred(r.bqb[r0]), BQAlt2 = phred(r.bqb[r1]), FSAlt = phred(r.strandBias[r0]), FSAlt2 = phred(r.strandBias[r1]);
      const double SOR = sor(r, r0, r1);
      const double QD = safe_div(r.depth[r0] + r.depth[r1], r.agg[r0] + r.agg[r1]) * 10;
      qualV = qual(r.subOptProb, r.optProb, r.depth[r0] + r.depth[r1], r.agg[r0] + r.agg[r1]);
      const std::string AD = pstr(r.agg[rRef]) + "," + pstr(r.agg[r0]) + "," + pstr(r.agg[r1]);
      const std::string PL = fmt("%d,%d,%d,%d,%d,%d", PLof(s1(refC) + refC), PLof(s1(refC) + alt[0]), PLof(s1(refC) + alt[1]),
                                 PLof(s1(alt[0]) + alt[0]), PLof(s1(alt[0]) + alt[1]), PLof(s1(alt[1]) + alt[1]));
      const std::string info = fmt("AC=1,1;AF=0.500,0.500;AN=2;FRAC=1.000;BQFisherPhredAlt=%d;BQFisherPhredAlt2=%d;BestGT=%s;DP=%lld;DP4=%s;DP8=%s;IndelMeta=%s;SBFisherPhredAlt=%d;SBFisherPhredAlt2=%d;",
                                   BQAlt, BQAlt2, r.altGT.c_str(), ipart(DP), dp4.c_str(), dp8.c_str(), meta.c_str(), FSAlt, FSAlt2) +
                               info_tail(r, true, QD, SOR, optSubopt);
      if (!refMatches(r.refBase)) return true;
      emit(r.refBase, s1(alt[0]) + "," + alt[1], info, format_of("1/2", AD, DP, QD, PL));
    } else { fprintf(stderr, "%lld: @altBase != 1 and != 2 in Het.\n", lineNo); return false; }
    return true;
  }

  // ---- indel rows --------------------------------------------------------------------------------------------------
  std::string refPat, altPat;
  if (kGTiii[gt][0] == kGTiii[gt][1]) {                                         // XX
    if (alt.size() != 1) { fprintf(stderr, "%lld: scalar (@altBase) != 1", lineNo); return false; }
    const int rAlt = rci(alt[0]);
    DP -= r.agg[rRef];
    const int FSAlt = phred(r.strandBias[rAlt]), BQAlt = phred(r.bqb[rAlt]);
    const double QD = safe_div(r.depth[rAlt], r.agg[rAlt]) * 10;
    qualV = qual(r.subOptProb, r.optProb, r.depth[rAlt], r.agg[rAlt]);
    const std::string AD = pstr(r.agg[rRef]) + "," + pstr(r.agg[rAlt]);
    const std::string PL = fmt("%d,%d,%d", PLof(s1(refC) + refC), PLof(s1(refC) + alt[0]), PLof(s1(alt[0]) + alt[0]));
    const std::string info = fmt("INDEL;AC=2;AF=1;AN=2;FRAC=1.000;BQFisherPhredAlt=%d;BestGT=%s;DP=%lld;DP4=%s;DP8=%s;IndelMeta=%s;SBFisherPhredAlt=%d;",
                                 BQAlt, r.altGT.c_str(), ipart(DP), dp4.c_str(), dp8.c_str(), meta.c_str(), FSAlt) +
                             info_tail(r, false, QD, 0, optSubopt);
    indel_patterns(refPat, altPat);
    if (!refMatches(refPat)) return true;
    emit(refPat, altPat, info, format_of("1/1", AD, DP, QD, PL));
  } else if (alt.size() == 1) {                                                // ref + X
    const int rAlt = rci(alt[0]);
    DP -= (double) ipart(r.agg[rRef] / 2);
    const double FRAC = safe_div(r.agg[rAlt], r.agg[rAlt] + r.agg[rRef]);
    const int FSRef = phred(r.strandBias[rRef]), BQRef = phred(r.bqb[rRef]), FSAlt = phred(r.strandBias[rAlt]), BQAlt = phred(r.bqb[rAlt]);
    const double QD = safe_div(r.depth[rAlt] + r.depth[rRef], r.agg[rAlt] + r.agg[rRef]) * 10;
    qualV = qual(r.subOptProb, r.optProb, r.depth[rAlt] + r.depth[rRef], r.agg[rAlt] + r.agg[rRef]);
    const std::string AD = pstr(r.agg[rRef]) + "," + pstr(r.agg[rAlt]);
    const std::string PL = fmt("%d,%d,%d", PLof(s1(refC) + refC), PLof(s1(refC) + alt[0]), PLof(s1(alt[0]) + alt[0]));
    const std::string info = fmt("INDEL;AC=1;AF=0.5;AN=2;FRAC=%.3f;BQFisherPhredRef=%d;BQFisherPhredAlt=%d;BestGT=%s;DP=%lld;DP4=%s;DP8=%s;IndelMeta=%s;SBFisherPhredRef=%d;SBFisherPhredAlt=%d;",
                                 FRAC, BQRef, BQAlt, r.altGT.c_str(), ipart(DP), dp4.c_str(), dp8.c_str(), meta.c_str(), FSRef, FSAlt) +
                             info_tail(r, false, QD, 0, optSubopt);
    indel_patterns(refPat, altPat);
    if (!refMatches(refPat)) return true;
    emit(refPat, altPat, info, format_of("0/1", AD, DP, QD, PL));
  } else if (alt.size() == 2) {
    if (r.altGT == "XY") {
      const int r0 = rci(alt[0]), r1 = rci(alt[1]);
      DP -= r.agg[rRef];
      const int BQAlt = phred(r.bqb[r0]), BQAlt2 = phred(r.bqb[r1]), FSAlt = phred(r.strandBias[r0]), FSAlt2 = phred(r.strandBias[r1]);
      const double QD = safe_div(r.depth[r0] + r.depth[r1], r.agg[r0] + r.agg[r1]) * 10;
      qualV = qual(r.subOptProb, r.optProb, r.depth[r0] + r.depth[r1], r.agg[r0] + r.agg[r1]);
      const std::string AD = pstr(r.agg[rRef]) + "," + pstr(r.agg[r0]) + "," + pstr(r.agg[r1]);
      const std::string PL = fmt("%d,%d,%d,%d,%d,%d", PLof(s1(refC) + refC), PLof(s1(refC) + alt[0]), phred(0), PLof(s1(alt[0]) + alt[0]),
                                 PLof(s1(alt[0]) + alt[1]), phred(0));
      const std::string AF = "0.500,0.500";
      const std::string &p1 = r.indelPattern, &p2 = r.indelPattern2;
      const bool D1 = ends_with(r.indelMode, 'D'), I1 = ends_with(r.indelMode, 'I'), D2 = ends_with(r.indelMode2, 'D'), I2 = ends_with(r.indelMode2, 'I');
      if (D1 && D2) {
        const std::string &lng = p1.size() > p2.size() ? p1 : p2, &sht = p1.size() > p2.size() ? p2 : p1;
        if (p1.size() == p2.size()) { fprintf(stderr, "%lld: Equal length of deletion: %s %s %s %s\n", lineNo, r.indelMode.c_str(), p1.c_str(), r.indelMode2.c_str(), p2.c_str()); return false; }
        std::string rest = lng;
        size_t k = rest.find(sht);                                             // s/$short//: the first occurrence is removed
        if (k == std::string::npos) { fprintf(stderr, "%lld: Unable to substract %s from %s.\n", lineNo, sht.c_str(), lng.c_str()); return false; }
        rest.erase(k, sht.size());
        refPat = r.preRefBase + lng;
        altPat = r.preRefBase + "," + r.preRefBase + rest;
      } else if (I1 && I2) {
        refPat = r.preRefBase; altPat = r.preRefBase + p1 + "," + r.preRefBase + p2;
      } else if (I1 && D2) {
        refPat = r.preRefBase + p2; altPat = r.preRefBase + "," + r.preRefBase + p2 + p1;
      } else if (D1 && I2) {
        refPat = r.preRefBase + p1; altPat = r.preRefBase + "," + r.preRefBase + p1 + p2;
      } else { fprintf(stderr, "%lld: Unknown Indel Mode: %s\n", lineNo, r.indelMode.c_str()); return false; }
      const std::string info = fmt("INDEL;AC=1,1;AF=%s;AN=2;FRAC=1.000;BQFisherPhredAlt=%d;BQFisherPhredAlt2=%d;BestGT=%s;DP=%lld;DP4=%s;DP8=%s;IndelMeta=%s;SBFisherPhredAlt=%d;SBFisherPhredAlt2=%d;",
                                   AF.c_str(), BQAlt, BQAlt2, r.altGT.c_str(), ipart(DP), dp4.c_str(), dp8.c_str(), meta.c_str(), FSAlt, FSAlt2) +
                               info_tail(r, false, QD, 0, optSubopt);
      if (!refMatches(refPat)) return true;
      emit(refPat, altPat, info, format_of("1/2", AD, DP, QD, PL));
    } else if (r.altGT.size() == 2 && strchr("ACGT", r.altGT[0]) && r.altGT[1] == 'X') {   // non-reference base + X: two rows
      {
        const int rAlt = rci(alt[1]);
        DP -= r.agg[rRef];
        const int FSAlt = phred(r.strandBias[rAlt]), BQAlt = phred(r.bqb[rAlt]);
        const double QD = safe_div(r.depth[rAlt], r.agg[rAlt]) * 10;
        qualV = qual(r.subOptProb, r.optProb, r.depth[rAlt], r.agg[rAlt]);
        const std::string AD = pstr(r.agg[rRef]) + "," + pstr(r.agg[rAlt]);
        const std::string PL = fmt("%d,%d,%d", PLof(s1(refC) + refC), PLof(s1(refC) + alt[1]), PLof(s1(alt[1]) + alt[1]));
        const std::string info = fmt("INDEL;AC=1;AF=0.5;AN=2;FRAC=1.000;BQFisherPhredAlt=%d;BestGT=%s;DP=%lld;DP4=%s;DP8=%s;IndelMeta=%s;SBFisherPhredAlt=%d;",
                                     BQAlt, r.altGT.c_str(), ipart(DP), dp4.c_str(), dp8.c_str(), meta.c_str(), FSAlt) +
                                 info_tail(r, false, QD, 0, optSubopt);
        indel_patterns(refPat, altPat);
        if (refMatches(refPat)) emit(refPat, altPat, info, format_of("0/1", AD, DP, QD, PL));
      }
      {                                                                        // NEXTALLELE: the base allele
        const int rAlt = rci(alt[0]);
        const int BQAlt = phred(r.bqb[rAlt]), FSAlt = phred(r.strandBias[rAlt]);
        const double QD = safe_div(r.depth[rAlt], r.agg[rAlt]) * 10;
        qualV = qual(r.subOptProb, r.optProb, r.depth[rAlt], r.agg[rAlt]);
        const std::string AD = pstr(r.agg[rRef]) + "," + pstr(r.agg[rAlt]);
        const std::string PL = fmt("%d,%d,%d", PLof(s1(refC) + refC), PLof(s1(refC) + alt[0]), PLof(s1(alt[0]) + alt[0]));
        const std::string info = fmt("AC=1;AF=0.5;AN=2;FRAC=1.000;BQFisherPhredAlt=%d;BestGT=%s;DP=%lld;DP4=%s;DP8=%s;IndelMeta=%s;SBFisherPhredAlt=%d;",
                                     BQAlt, r.altGT.c_str(), ipart(DP), dp4.c_str(), dp8.c_str(), meta.c_str(), FSAlt) +
                                 info_tail(r, false, QD, 0, optSubopt);
        if (!refMatches(r.refBase)) return true;
        emit(r.refBase, s1(alt[0]), info, format_of("0/1", AD, DP, QD, PL));
      }
    } else { fprintf(stderr, "%lld: Problematic alternative GT: %s\n", lineNo, r.altGT.c_str()); return false; }
  } else { fprintf(stderr, "%lld: @altBase != 1 and != 2 in Het.\n", lineNo); return false; }
  return true;
}

int main(int argc, char **argv) {
  if (argc < 4) { fprintf(stderr, "perl %s <txt Input> <Sample Name> <Referece FASTA>\n", argv[0]); return 255; }
  Converter C;
  C.qual_max = 4.343 * log(DBL_MAX);
  // reference FASTA: id = header up to the first blank, sequence upper-cased (acgtn only, as the script does)
  {
    FILE *fa = fopen(argv[3], "r");
    if (!fa) { perror(argv[3]); return 2; }
    std::string *seq = NULL;
    char *line = NULL; size_t cap = 0; ssize_t n;
    while ((n = getline(&line, &cap, fa)) > 0) {
      if (line[0] == '>') {
        char *e = line + 1;
        while (*e && *e != ' ' && *e != '\t' && *e != '\n' && *e != '\r' && *e != '\f' && *e != '\v') ++e;
        seq = &C.ref[std::string(line + 1, e)];
        seq->clear();
      } else if (seq) {
        for (ssize_t k = 0; k < n; ++k) {
          char c = line[k];
          if (c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v') continue;
          switch (c) { case 'a': c = 'A'; break; case 'c': c = 'C'; break; case 'g': c = 'G'; break; case 't': c = 'T'; break; case 'n': c = 'N'; break; }
          seq->push_back(c);
        }
      }
    }
    free(line);
    fclose(fa);
  }
  std::string cmd = std::string("gzip -dcf '") + argv[1] + "'";
  FILE *in = popen(cmd.c_str(), "r");
  if (!in) { perror(argv[1]); return 2; }
  for (size_t k = 0; k < sizeof(kHeader) / sizeof(kHeader[0]); ++k) { fputs(kHeader[k], stdout); fputc('\n', stdout); }
  printf("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t%s\n", argv[2]);

  // blocks of whole lines; the lines of a block are converted by worker threads (rows are independent), output in order
  unsigned nt = std::thread::hardware_concurrency();
  nt = nt ? (nt > 16 ? 16 : nt) : 4;
  if (const char *e = getenv("GT16_VCF_THREADS")) { int v = atoi(e); if (v >= 1 && v <= 64) nt = (unsigned) v; }
  const size_t kBlock = (size_t) 32 << 20;
  std::vector<char> buf(kBlock + 1);
  size_t have = 0;
  long long lineNo = 0;
  bool first = true, failed = false;
  for (;;) {
    const size_t got = fread(buf.data() + have, 1, buf.size() - 1 - have, in);
    have += got;
    if (have == 0) break;
    size_t end = have;
    if (got != 0) {                                                              // keep the last, possibly partial, line for the next round
      while (end > 0 && buf[end - 1] != '\n') --end;
      if (end == 0) {
        if (have == buf.size() - 1) { buf.resize(buf.size() * 2); continue; }     // a line longer than the block
        continue;
      }
    }
    buf[have] = 0;
    // line starts of [0, end)
    std::vector<size_t> starts;
    for (size_t p = 0; p < end;) {
      starts.push_back(p);
      const char *nl = (const char *) memchr(buf.data() + p, '\n', end - p);
      p = nl ? (size_t) (nl - buf.data()) + 1 : end;
    }
    starts.push_back(end);
    size_t firstLine = 0;
    if (first && starts.size() > 1) { firstLine = 1; first = false; }              // the column header
    const size_t nLines = starts.size() - 1;
    const unsigned T = nLines - firstLine < 256 ? 1u : nt;
    std::vector<std::string> outs(T);
    std::vector<char> ok(T, 1);
    auto work = [&](unsigned t) {
      const size_t per = (nLines - firstLine + T - 1) / T, a = firstLine + t * per, b = std::min(nLines, a + per);
      outs[t].reserve((b > a ? b - a : 0) * 640);
      for (size_t k = a; k < b; ++k)
        if (!C.row(buf.data() + starts[k], starts[k + 1] - starts[k], lineNo + (long long) k + 1, outs[t])) { ok[t] = 0; return; }
    };
    if (T == 1) work(0);
    else {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < T; ++t) th.emplace_back(work, t);
      for (auto &x : th) x.join();
    }
    for (unsigned t = 0; t < T; ++t) {
      fwrite(outs[t].data(), 1, outs[t].size(), stdout);
      if (!ok[t]) { failed = true; break; }
    }
    if (failed) break;
    lineNo += (long long) nLines;
    memmove(buf.data(), buf.data() + end, have - end);
    have -= end;
    if (got == 0) break;
  }
  pclose(in);
  return failed ? 255 : 0;
}
