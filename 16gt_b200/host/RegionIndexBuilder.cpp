// RegionIndexBuilder -- drop-in for the reference tool of the same name (buildRegionList.cpp:234-285):
//   RegionIndexBuilder [Index] [Input Region List] [Output List Name] -bed|-gff
// BED / GFF intervals -> region.bin = u32 n; {u32 start, end}[n] in concatenated coordinates, sorted by start and
// merged the way the reference merges them (including its quirk, see merge()), the input of snapshotSnpcaller -e.
// Needs only <Index>.ann and <Index>.tra, read through libgt16's index reader.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <string>
#include <vector>
#include "gt16.h"

struct IndexRegion { unsigned int startPos, endPos; };

// isOverlap (buildRegionList.cpp:15-43): a and b overlap when the later start lies before the earlier region's end
static bool overlap(const IndexRegion &a, const IndexRegion &b, IndexRegion &c) {
  const IndexRegion &l = a.startPos < b.startPos ? a : b, &r = a.startPos < b.startPos ? b : a;
  if (!(r.startPos < l.endPos)) return false;
  c.startPos = l.startPos;
  c.endPos = r.endPos < l.endPos ? l.endPos : r.endPos;
  return true;
}

// The reference's merge pass (buildRegionList.cpp:188-212).  It compares every following region with region[i],
// the FIRST region of the current group as it stands in the array, and stores the union of just those two at
// the output cursor; the union only accumulates while the output cursor is still on region[i] itself (i.e. until
// the first non-overlap after a merge).  Reproduced as is: region.bin must be byte-identical.
static unsigned int merge(std::vector<IndexRegion> &reg) {
  const unsigned int n = (unsigned int) reg.size();
  if (n == 0) return 0;
  unsigned int newSize = 0, i = 0, nextI = 1;
  for (;;) {
    if (i == n - 1 || nextI == n) { newSize++; break; }
    IndexRegion c;
    if (overlap(reg[i], reg[nextI], c)) {
      reg[newSize] = c;
      nextI++;
    } else {
      newSize++;
      i = nextI++;
      reg[newSize] = reg[i];
    }
  }
  return newSize;
}

int main(int argc, char **argv) {
  if (argc != 5) {
    fprintf(stderr, "Usage:\n\t%s [Index] [Input Region List] [Output List Name] [format]\n", argv[0]);
    fprintf(stderr, "\tformat: -bed OR -gff\n");
    return 1;
  }
  printf("Loading Index\n");
  char err[512] = "";
  gt16_index *ix = NULL;
  if (gt16_index_read(argv[1], &ix, err, sizeof(err)) != GT16_OK) { fprintf(stderr, "%s\n", err); return 1; }
  const unsigned int textLength = gt16_index_length(ix);
  printf("Loaded Translate : %u\n", textLength);
  int fileType = strcmp(argv[4], "-gff") == 0 ? 1 : strcmp(argv[4], "-bed") == 0 ? 2 : 0;
  if (!fileType) { printf("Invalid input format\n"); return 1; }
  FILE *in = fopen(argv[2], "r");
  if (!in) { fprintf(stderr, "Cannot open file : %s\n", argv[2]); return 1; }
  FILE *out = fopen(argv[3], "wb");
  if (!out) { fprintf(stderr, "Cannot open file : %s\n", argv[3]); return 1; }

  std::vector<IndexRegion> reg;
  std::vector<char> line(1 << 16);
  while (fgets(line.data(), (int) line.size(), in)) {
    char *s = line.data();
    if (*s == '#' || *s == '\n' || *s == 0) continue;                 // comment lines (buildRegionList.cpp:92-96)
    // name, [GFF: two more fields], start, end -- separated by blanks or tabs; the rest of the line is ignored
    char *save = NULL;
    char *name = strtok_r(s, " \t\n", &save);
    if (!name) continue;
    if (fileType == 1) { strtok_r(NULL, " \t\n", &save); strtok_r(NULL, " \t\n", &save); }
    char *a = strtok_r(NULL, " \t\n", &save), *b = strtok_r(NULL, " \t\n", &save);
    if (!a || !b) { printf("File Format Error\n"); break; }
    for (const char *p = a; *p; ++p) if (*p < '0' || *p > '9') { fprintf(stderr, "Error in parsing gff file : position is not a number\n"); return 1; }
    for (const char *p = b; *p; ++p) if (*p < '0' || *p > '9') { fprintf(stderr, "Error in parsing gff file : position is not a number\n"); return 1; }
    if (reg.size() % 10000 == 0) printf("Handled %u entries\n", (unsigned) reg.size());
    const unsigned int chrPos0 = (unsigned int) strtoul(a, NULL, 10) + 1, chrPos1 = (unsigned int) strtoul(b, NULL, 10) + 1;   // :163-164
    unsigned int amb0, amb1;
    if (gt16_index_locate(ix, name, chrPos0, &amb0) == 0 && gt16_index_locate(ix, name, chrPos1, &amb1) == 0 &&
        amb0 < textLength && amb1 < textLength) {
      IndexRegion r = {amb0, amb1};
      reg.push_back(r);
    }
  }
  // two stable 16-bit radix passes on startPos == a stable sort by startPos (:182-186)
  std::stable_sort(reg.begin(), reg.end(), [](const IndexRegion &x, const IndexRegion &y) { return x.startPos < y.startPos; });
  const unsigned int newSize = merge(reg);
  fwrite(&newSize, sizeof(unsigned int), 1, out);
  if (newSize) fwrite(reg.data(), sizeof(IndexRegion), newSize, out);
  printf("\nRemoved Overlapped entries : %u remained\n", newSize);
  fclose(in);
  fclose(out);
  gt16_index_free(ix);
  return 0;
}
