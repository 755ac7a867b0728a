// TEST INFRASTRUCTURE ONLY.  FASTA -> .pac/.ann/.amb/.tra through the reference's
// own packer (HSPParseFASTAToPacked in lib/lib.a, HSP.o); used to validate the
// index files written by tools/gt16_synth.
// usage: ref_mkindex <in.fa> <indexPrefix>
#include <stdio.h>
#include <string>
unsigned int HSPParseFASTAToPacked(const char *fasta, const char *ann, const char *pac, const char *amb,
                                   const char *tra, unsigned int seed, int maskLowerCase);
int main(int argc, char **argv) {
  if (argc != 3) { fprintf(stderr, "usage: %s in.fa indexPrefix\n", argv[0]); return 2; }
  std::string p = argv[2];
  HSPParseFASTAToPacked(argv[1], (p + ".ann").c_str(), (p + ".pac").c_str(), (p + ".amb").c_str(),
                        (p + ".tra").c_str(), 0, 1);
  return 0;
}
