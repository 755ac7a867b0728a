// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Driver for the *unmodified reference objects* (compiled from /root/reference
// where they lie, see oracle/build_ref.sh).  It performs the same sequence of
// calls as the reference's main (snapshotSnpcaller.cpp:133-260) with ONE
// substitution: loadIndex() (which needs a genuine SOAP3-dp BWT/SA that cannot
// be built offline) is replaced by HSPLoad() from lib/lib.a, which loads the
// .pac/.ann/.amb/.tra files -- the only index members the caller ever reads
// (the BWT is freed unused at snapshotSnpcaller.cpp:196-197).
//
// usage: ref_caller <indexPrefix> <outPrefix> [region.bin] [-v] [-t threads] [-c tables.bin]
// reads  <outPrefix>.snapshot, writes <outPrefix>.txt
// -c: keep the reference's own prefilled LikelihoodCache in a file between runs (the prefill
//     takes ~26-90 s and is identical for identical ini error rates).  Used when generating
//     golden fixtures and by bench.py's reference arm: the prefill is outside the timed quantity
//     (REF_SELECT_SECONDS = selectPossibleSNPs alone) on both arms, so reusing the reference's own
//     table bytes between its steps changes nothing that is measured.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <sys/time.h>

#include "dependencies.h"
#include "CPUfunctions.h"
#include "VariantCaller.h"
#include "FisherExactTest.h"
#include "SnapshotHandler.h"

// present in lib/lib.a (HSP.o) but not declared by the shipped headers
HSP *HSPLoad(const char *pac, const char *ann, const char *amb, const char *tra,
             unsigned int trailerBufferInWord, char shareIndex);

// own wall clock for the REF_* lines (lib.a's getElapsedTime is only used where the reference uses it)
static double wall() { struct timeval tv; gettimeofday(&tv, NULL); return tv.tv_sec + tv.tv_usec * 1e-6; }

int main(int argc, char **argv) {
  if (argc < 3) {
    fprintf(stderr, "usage: %s <indexPrefix> <outPrefix> [region.bin] [-v] [-t n]\n", argv[0]);
    return 2;
  }
  std::string idx = argv[1], out = argv[2];
  const char *regionFile = NULL;
  bool verbose = false;
  int threads = 0;
  const char *lcFile = NULL;
  for (int i = 3; i < argc; ++i) {
    if (!strcmp(argv[i], "-v")) verbose = true;
    else if (!strcmp(argv[i], "-c") && i + 1 < argc) lcFile = argv[++i];
    else if (!strcmp(argv[i], "-t") && i + 1 < argc) threads = atoi(argv[++i]);
    else regionFile = argv[i];
  }

  IniParams ini;
  std::string iniName = std::string(argv[0]) + ".ini";
  if (ParseIniFile((char *) iniName.c_str(), ini) != 0) {
    fprintf(stderr, "Failed to parse configuration file ... %s\n", iniName.c_str());
    return 1;
  }
  if (threads > 0) ini.Ini_NumOfCpuThreads = threads;

  double startTime = setStartTime();
  HSP *hsp = HSPLoad((idx + ".pac").c_str(), (idx + ".ann").c_str(), (idx + ".amb").c_str(),
                     (idx + ".tra").c_str(), 1, 0);
  double lastEventTime = getElapsedTime(startTime);
  printf("Reference sequence length: %u\n", hsp->dnaLength);

  hsp->snpBundle.numOfCPUThreads = ini.Ini_NumOfCpuThreads;
  hsp->snpFlag = SNP_STAT_FLAG;

  InputOptions opt;
  memset(&opt, 0, sizeof(opt));
  opt.resultPrefix = (char *) out.c_str();
  opt.verbose = verbose;
  ExomeRegion *regions = NULL;
  unsigned int nRegions = 0;
  if (regionFile) {
    FILE *f = fopen(regionFile, "rb");
    if (!f) { fprintf(stderr, "Exome Region Index %s doesn't exist\n", regionFile); return 1; }
    if (fread(&nRegions, sizeof(unsigned int), 1, f) != 1) return 1;
    regions = (ExomeRegion *) malloc(nRegions * sizeof(ExomeRegion));
    if (fread(regions, sizeof(ExomeRegion), nRegions, f) != nRegions) return 1;
    fclose(f);
    opt.isExome = 1;
    opt.exomeRegionFileName = (char *) regionFile;
    printf("%u Exome Regions loaded\n", nRegions);
  }

  readSnpInfoSnapshot(&(hsp->snpBundle), hsp->dnaLength, (out + ".snapshot").c_str());
  double t = getElapsedTime(startTime);
  printf("Read snapshot in %9.4f seconds\n", t - lastEventTime);
  lastEventTime = t;

  double w0 = wall();
  LikelihoodCache *lc = (LikelihoodCache *) malloc(sizeof(LikelihoodCache));
  bool haveLc = false;
  if (lcFile) {
    FILE *f = fopen(lcFile, "rb");
    float key[2] = {0, 0};
    if (f && fread(key, sizeof(float), 2, f) == 2 && key[0] == (float) ini.Ini_BalanceSubError &&
        key[1] == (float) ini.Ini_UnbalanceSubError && fread(lc, sizeof(LikelihoodCache), 1, f) == 1)
      haveLc = true;
    if (f) fclose(f);
  }
  if (!haveLc) {
    prefill_likelihood_cache_with_p_err(lc, ini.Ini_BalanceSubError, ini.Ini_UnbalanceSubError);
    if (lcFile) {
      FILE *f = fopen(lcFile, "wb");
      float key[2] = {(float) ini.Ini_BalanceSubError, (float) ini.Ini_UnbalanceSubError};
      if (f) { fwrite(key, sizeof(float), 2, f); fwrite(lc, sizeof(LikelihoodCache), 1, f); fclose(f); }
    }
  }
  t = getElapsedTime(startTime);
  printf("REF_PREFILL_SECONDS %9.4f\n", wall() - w0);
  lastEventTime = t;
  w0 = wall();

  unsigned int attriSize = 0;
  std::string tmp = out + ".tmpresult", txt = out + ".txt";
  selectPossibleSNPs(hsp->snpBundle, hsp->packedDNA, hsp->dnaLength, hsp->annotation, &attriSize,
                     opt.isExome, regions, nRegions, lc, opt, ini, hsp->ambiguityMap, hsp->translate,
                     tmp.c_str(), startTime, lastEventTime);
  t = getElapsedTime(startTime);
  printf("REF_SELECT_SECONDS %9.4f\n", wall() - w0);
  lastEventTime = t;
  w0 = wall();

  calRFpedictProb(tmp.c_str(), txt.c_str(), attriSize, verbose);
  printf("REF_EMIT_SECONDS %9.4f\n", wall() - w0);
  remove(tmp.c_str());
  printf("Total Running Time: %9.4f\n", getElapsedTime(startTime));
  return 0;
}
