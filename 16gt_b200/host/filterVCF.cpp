// filterVCF -- native replacement for the reference's filterVCF.pl, the last step of the 16GT pipeline:
//   filterVCF <16GT VCF Input> <dbSNP VCF Input> [Min QUAL] [Min FRAC SNP] [Min FRAC INDEL] [Max SOR SNP] [Max FS INDEL]
//             [Min DP] [Max DP Multiplier] [Min Strand Count] [Min QD SNP] [Min QD INDEL]   >  filtered.vcf
// Same arguments and defaults, same stderr protocol, same output bytes: known dbSNP sites get their rs ID, the rest
// is tagged QUALlow / DPhigh / DPlow / SORhigh / FShigh / FRAClow / SClow / QDlow (filterVCF.pl:126-225).  The input
// may be a .vcf, a .vcf.gz or a .list of such files (then every file gets its <name>.filtered.vcf).  One difference, and
// it is the script's: a site matching several rs IDs prints them in Perl's randomised hash order there, in order of
// first appearance here.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <string>
#include <unordered_map>
#include <vector>

static bool exists(const char *p) { struct stat st; return stat(p, &st) == 0; }
static bool ends_with(const std::string &s, const char *suf) { size_t n = strlen(suf); return s.size() >= n && s.compare(s.size() - n, n, suf) == 0; }
static double num(const std::string &s) { return strtod(s.c_str(), NULL); }
static std::string pstr(double v) {                                             // Perl's stringification of a number
  char b[64];
  if (v == floor(v) && fabs(v) < 1e15) snprintf(b, sizeof(b), "%.0f", v);
  else snprintf(b, sizeof(b), "%.15g", v);
  return b;
}
static bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

// /.vcf.gz/ resp. /\.vcf.gz$/ : the script's dots are unescaped (any character) except the leading one of the anchored form
static bool has_vcf_gz(const std::string &s, bool atEnd) {
  if (atEnd) return s.size() >= 7 && s[s.size() - 7] == '.' && s.compare(s.size() - 6, 3, "vcf") == 0 && s.compare(s.size() - 2, 2, "gz") == 0;
  for (size_t p = 1; p + 6 <= s.size(); ++p) if (s.compare(p, 3, "vcf") == 0 && s.compare(p + 4, 2, "gz") == 0) return true;
  return false;
}

static bool read_lines(const std::string &fn, std::vector<std::string> &lines) {  // open "gzip -dcf $fn |"
  std::string cmd = "gzip -dcf " + fn;
  FILE *f = popen(cmd.c_str(), "r");
  if (!f) return false;
  char *line = NULL; size_t cap = 0; ssize_t n;
  while ((n = getline(&line, &cap, f)) > 0) lines.push_back(std::string(line, (size_t) n));
  free(line);
  pclose(f);
  return true;
}

// /KEY=(\d+?);/ : first "KEY=" followed by one or more digits and a semicolon
static bool find_uint(const std::string &s, const char *key, std::string &out) {
  const size_t kl = strlen(key);
  for (size_t p = s.find(key); p != std::string::npos; p = s.find(key, p + 1)) {
    size_t q = p + kl;
    while (q < s.size() && s[q] >= '0' && s[q] <= '9') ++q;
    if (q > p + kl && q < s.size() && s[q] == ';') { out = s.substr(p + kl, q - p - kl); return true; }
  }
  return false;
}
// /KEY=([-+]?[0-9]*\.?[0-9]+);/
static bool find_float(const std::string &s, const char *key, std::string &out) {
  const size_t kl = strlen(key);
  for (size_t p = s.find(key); p != std::string::npos; p = s.find(key, p + 1)) {
    size_t q = p + kl, b = q;
    if (q < s.size() && (s[q] == '-' || s[q] == '+')) ++q;
    size_t d0 = q;
    while (q < s.size() && s[q] >= '0' && s[q] <= '9') ++q;
    size_t intDigits = q - d0, fracDigits = 0;
    if (q < s.size() && s[q] == '.') {
      size_t r = q + 1;
      while (r < s.size() && s[r] >= '0' && s[r] <= '9') ++r;
      fracDigits = r - q - 1;
      if (fracDigits) q = r;                                   // "12." alone does not match \.?[0-9]+ unless digits follow
    }
    if ((fracDigits || intDigits) && q < s.size() && s[q] == ';') { out = s.substr(b, q - b); return true; }
  }
  return false;
}

int main(int argc, char **argv) {
  const double defMinQual = 10, defMinFracSNP = 0.3, defMinFracINDEL = 0.15, defMaxSORSNP = 3, defMaxFSINDEL = 30, defMinDP = 3,
               defMaxDPMultiplier = 5, defMinStrandCount = 1, defMinQDSNP = 24, defMinQDINDEL = 20;
  if (argc < 3) {
    fprintf(stderr, "perl %s <16GT VCF Input> <dbSNP VCF Input> [Min QUAL, %s] [Min FRAC SNP, %s] [Min FRAC INDEL, %s] [Max SOR SNP, %s] "
                    "[Max FS INDEL, %s] [Min DP, %s] [Max DP Multiplier, meanDP+FLOAT*sqrt(meanDP), %s] [Min Strand Count, %s] [Min QD SNP, %s] "
                    "[Min QD INDEL, %s]\n", argv[0], pstr(defMinQual).c_str(), pstr(defMinFracSNP).c_str(), pstr(defMinFracINDEL).c_str(),
            pstr(defMaxSORSNP).c_str(), pstr(defMaxFSINDEL).c_str(), pstr(defMinDP).c_str(), pstr(defMaxDPMultiplier).c_str(),
            pstr(defMinStrandCount).c_str(), pstr(defMinQDSNP).c_str(), pstr(defMinQDINDEL).c_str());
    return 0;
  }
  if (!exists(argv[1])) { fprintf(stderr, "Failed to open %s, exiting...\n", argv[1]); return 2; }
  if (!exists(argv[2])) { fprintf(stderr, "Failed to open %s, exiting...\n", argv[2]); return 2; }

  std::unordered_map<std::string, std::string> dbSNP;
  {
    fprintf(stderr, "Loading dbSNP %s...\n", argv[2]);
    std::vector<std::string> lines;
    read_lines(argv[2], lines);
    for (const std::string &l : lines) {
      if (!l.empty() && l[0] == '#') continue;
      std::vector<std::string> F;                                // split /\t/ (the newline stays on the last field)
      for (size_t p = 0;;) { size_t q = l.find('\t', p); F.push_back(l.substr(p, q == std::string::npos ? q : q - p)); if (q == std::string::npos) break; p = q + 1; }
      if (F.size() < 5) continue;
      for (size_t p = 0;;) {
        size_t q = F[4].find(',', p);
        const std::string alt = F[4].substr(p, q == std::string::npos ? q : q - p);
        dbSNP[F[0] + "-" + F[1] + "-" + F[3] + "-" + alt] = F[2];
        if (q == std::string::npos) break;
        p = q + 1;
      }
    }
    fprintf(stderr, "Loaded %zu dbSNP entries from %s\n", dbSNP.size(), argv[2]);
  }

  const std::string in0 = argv[1];
  bool isList = false;
  std::vector<std::string> files;
  if (ends_with(in0, ".list")) {
    isList = true;
    FILE *f = fopen(argv[1], "r");
    if (!f) { perror(argv[1]); return 2; }
    char *line = NULL; size_t cap = 0; ssize_t n;
    while ((n = getline(&line, &cap, f)) > 0) { std::string s(line, (size_t) n); if (!s.empty() && s.back() == '\n') s.pop_back(); files.push_back(s); }
    free(line);
    fclose(f);
  } else if (ends_with(in0, ".vcf") || has_vcf_gz(in0, false)) {
    files.push_back(in0);
  } else {
    fprintf(stderr, "Input is neither a VCF file ended with .vcf or .vcf.gz, nor a list file ended with .list\n");
  }

  auto arg = [&](int k, double def) { return argc > k ? std::string(argv[k]) : pstr(def); };
  for (const std::string &fn : files) {
    FILE *out = stdout;
    if (isList) {
      std::string ofn = fn;
      if (ends_with(ofn, ".vcf")) ofn.erase(ofn.size() - 4);
      else if (has_vcf_gz(ofn, true)) ofn.erase(ofn.size() - 7);
      ofn += ".filtered.vcf";
      out = fopen(ofn.c_str(), "w");
      if (!out) { perror(ofn.c_str()); return 2; }
    }
    fprintf(stderr, "Loading VCF %s...\n", fn.c_str());
    std::vector<std::string> lines;
    read_lines(fn, lines);
    fprintf(stderr, "Loaded %zu lines from %s\n", lines.size(), fn.c_str());

    long long meanDP = 0;
    {
      fprintf(stderr, "First pass (determine mean depth) ...\n");
      std::vector<long long> dpCount;
      std::string m;
      for (const std::string &l : lines) {
        if (!l.empty() && l[0] == '#') continue;
        if (find_uint(l, "DP=", m)) {
          const size_t d = (size_t) strtoull(m.c_str(), NULL, 10);
          if (dpCount.size() <= d) dpCount.resize(d + 1, 0);
          ++dpCount[d];
        }
      }
      long long mx = 0;
      for (size_t i = 0; i < dpCount.size(); ++i)
        if (dpCount[i] > 0 && dpCount[i] >= mx) { meanDP = (long long) i; mx = dpCount[i]; }   // the most frequent depth (largest on ties)
      fprintf(stderr, "Mean DP: %lld\n", meanDP);
    }
    const std::string minQual = arg(3, defMinQual), minFracSNP = arg(4, defMinFracSNP), minFracINDEL = arg(5, defMinFracINDEL),
                      maxSORSNP = arg(6, defMaxSORSNP), maxFSINDEL = arg(7, defMaxFSINDEL), minDP = arg(8, defMinDP),
                      maxDPMultiplier = arg(9, defMaxDPMultiplier), minStrandCount = arg(10, defMinStrandCount),
                      minQDSNP = arg(11, defMinQDSNP), minQDINDEL = arg(12, defMinQDINDEL);
    double maxDP = (double) meanDP + num(maxDPMultiplier) * sqrt((double) meanDP);
    maxDP = (double) (long long) (maxDP + 0.51);
    fprintf(stderr, "Parameters:\nMin Qual: %s\nMin FRAC SNP: %s\nMin FRAC INDEL: %s\nMax SOR SNP: %s\nMax FS INDEL: %s\nMin DP: %s\n"
                    "Max DP Multiplier: %s\nMax DP: %s\nMin Strand Count: %s\nMin QD SNP: %s\nMin QD INDEL: %s\n",
            minQual.c_str(), minFracSNP.c_str(), minFracINDEL.c_str(), maxSORSNP.c_str(), maxFSINDEL.c_str(), minDP.c_str(),
            maxDPMultiplier.c_str(), pstr(maxDP).c_str(), minStrandCount.c_str(), minQDSNP.c_str(), minQDINDEL.c_str());
    fprintf(stderr, "Second pass (add filter tags) ...\n");

    std::string m;
    for (const std::string &l : lines) {
      if (!l.empty() && l[0] == '#') {
        if (l.compare(0, 6, "#CHROM") == 0) {
          fprintf(out, "##FILTER=<ID=QUALlow,Description=\"Low Quality: QUAL<%s\">\n", minQual.c_str());
          fprintf(out, "##FILTER=<ID=DPhigh,Description=\"High read depth: DP>%s\">\n", pstr(maxDP).c_str());
          fprintf(out, "##FILTER=<ID=DPlow,Description=\"Low read depth: DP<%s\">\n", minDP.c_str());
          fprintf(out, "##FILTER=<ID=SORhigh,Description=\"Large SOR at SNPs: SOR>%s\">\n", maxSORSNP.c_str());
          fprintf(out, "##FILTER=<ID=FShigh,Description=\"Large Fisher-Strand bias at INDELs: SBFisherPhredAlt>%s\">\n", maxFSINDEL.c_str());
          fprintf(out, "##FILTER=<ID=FRAClow,Description=\"Low fraction of non-reference reads: FRAC<%s at SNPs or FRAC<%s at INDELs\">\n", minFracSNP.c_str(), minFracINDEL.c_str());
          fprintf(out, "##FILTER=<ID=SClow,Description=\"Low double-strand support at SNPs: Lowest Strand Count Support of Alt<%s\">\n", minStrandCount.c_str());
          fprintf(out, "##FILTER=<ID=QDlow,Description=\"Low QD: QD<%s at SNPs or QD<%s at INDELs\">\n", minQDSNP.c_str(), minQDINDEL.c_str());
        }
        fputs(l.c_str(), out);
        continue;
      }
      std::vector<std::string> a;                                // split /\s+/
      for (size_t p = 0; p < l.size();) {
        size_t q = p;
        while (q < l.size() && !is_space(l[q])) ++q;
        a.push_back(l.substr(p, q - p));
        p = q;
        while (p < l.size() && is_space(l[p])) ++p;
      }
      if (a.size() < 8) a.resize(8);
      std::vector<const char *> tags;
      const std::string &info = a[7];
      if (num(a[5]) < num(minQual)) tags.push_back("QUALlow");
      if (find_uint(info, "DP=", m)) {
        if (num(m) > maxDP) tags.push_back("DPhigh");
        if (num(m) < num(minDP)) tags.push_back("DPlow");
      } else { fprintf(stderr, "no DP tag: %s", info.c_str()); return 255; }
      if (info.find("INDEL;") != std::string::npos) {
        if (find_uint(info, "SBFisherPhredAlt=", m)) { if (num(m) > num(maxFSINDEL)) tags.push_back("FShigh"); }
        if (find_float(info, "FRAC=", m)) { if (num(m) < num(minFracINDEL)) tags.push_back("FRAClow"); }
        else { fprintf(stderr, "no FRAC tag: %s", info.c_str()); return 255; }
        if (find_float(info, "QD=", m)) { if (num(m) < num(minQDINDEL)) tags.push_back("QDlow"); }
        else { fprintf(stderr, "no QD tag: %s", info.c_str()); return 255; }
      } else {
        if (find_float(info, "SOR=", m)) { if (num(m) > num(maxSORSNP)) tags.push_back("SORhigh"); }
        else if (find_uint(info, "SBFisherPhredAlt=", m)) { if (num(m) > num(maxFSINDEL)) tags.push_back("FShigh"); }
        if (find_float(info, "FRAC=", m)) { if (num(m) < num(minFracSNP)) tags.push_back("FRAClow"); }
        else { fprintf(stderr, "no FRAC tag: %s", info.c_str()); return 255; }
        if (find_float(info, "QD=", m)) { if (num(m) < num(minQDSNP)) tags.push_back("QDlow"); }
        else { fprintf(stderr, "no QD tag: %s", info.c_str()); return 255; }
        // DP8=(\S+?); : the eight per-strand depths; both strands of every alternative base must be supported
        size_t p = info.find("DP8=");
        size_t q = p == std::string::npos ? p : info.find(';', p + 4);
        if (p == std::string::npos || q == std::string::npos || q == p + 4) { fprintf(stderr, "no DP8 tag: %s", info.c_str()); return 255; }
        std::vector<double> sc;
        {
          const std::string v = info.substr(p + 4, q - p - 4);
          for (size_t b = 0;;) { size_t e = v.find(',', b); sc.push_back(num(v.substr(b, e == std::string::npos ? e : e - b))); if (e == std::string::npos) break; b = e + 1; }
        }
        bool flag = false;
        for (size_t b = 0;;) {
          size_t e = a[4].find(',', b);
          std::string alt = a[4].substr(b, e == std::string::npos ? e : e - b);
          for (char &c : alt) { if (c == 'A') c = '0'; else if (c == 'C') c = '2'; else if (c == 'G') c = '4'; else if (c == 'T') c = '6'; }   // tr/ACGT/0246/
          const long long i = (long long) num(alt);
          const double s0 = i >= 0 && (size_t) i < sc.size() ? sc[(size_t) i] : 0.0, s1 = i + 1 >= 0 && (size_t) (i + 1) < sc.size() ? sc[(size_t) (i + 1)] : 0.0;
          if (s0 < num(minStrandCount) || s1 < num(minStrandCount)) flag = true;
          if (e == std::string::npos) break;
          b = e + 1;
        }
        if (flag) tags.push_back("SClow");
      }
      std::vector<std::string> rs;
      for (size_t b = 0;;) {
        size_t e = a[4].find(',', b);
        const std::string alt = a[4].substr(b, e == std::string::npos ? e : e - b);
        auto it = dbSNP.find(a[0] + "-" + a[1] + "-" + a[3] + "-" + alt);
        if (it != dbSNP.end()) {
          bool seen = false;
          for (const std::string &x : rs) if (x == it->second) seen = true;
          if (!seen) rs.push_back(it->second);
        }
        if (e == std::string::npos) break;
        b = e + 1;
      }
      if (!rs.empty()) { a[2] = rs[0]; for (size_t k = 1; k < rs.size(); ++k) a[2] += "," + rs[k]; }
      if (!tags.empty() && a[2] == ".") {
        a[6] = tags[0];
        for (size_t k = 1; k < tags.size(); ++k) { a[6] += ";"; a[6] += tags[k]; }
      }
      for (size_t k = 0; k < a.size(); ++k) { if (k) fputc('\t', out); fputs(a[k].c_str(), out); }
      fputc('\n', out);
    }
    fprintf(stderr, "Filtering %s done.\n", fn.c_str());
    if (isList) fclose(out);
  }
  return 0;
}
