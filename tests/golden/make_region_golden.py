#!/usr/bin/env python
"""Generate tests/golden/regions/: seeded BED and GFF interval lists over the stress_small index and the region.bin the
UNMODIFIED reference RegionIndexBuilder (oracle/_ref/ref_RegionIndexBuilder, buildRegionList.cpp) writes for them.

Run in the build container:  python tests/golden/make_region_golden.py
The lists cover: blank/tab separators, comment lines, names with and without the "chr" prefix (getChrIDFromName's
leniency), unknown sequences, positions beyond a sequence's end, empty and reversed intervals, heavy overlap (the
merge pass and its first-of-group quirk).
"""
import gzip, os, random, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gt16_testlib as T  # noqa: E402

LENS = {"chr1": 222222, "chr2": 177778, "1": 222222, "2": 177778}


def interval_list(fmt, seed, n=1200):
    rnd = random.Random(seed)
    lines = ["# seeded interval list, %s, seed %d" % (fmt, seed), ""]
    for i in range(n):
        name = rnd.choice(["chr1", "chr2", "chr1", "chr2", "1", "2", "chrX", "chrchr1"])
        L = LENS.get(name, 100000)
        a = rnd.randrange(0, L + 50)
        b = max(0, a + rnd.choice([0, 1, 5, 100, 300, 2000, -3]))
        sep = rnd.choice(["\t", " "])
        if fmt == "bed":
            lines.append(sep.join([name, str(a), str(b), "t%d" % i, "0", "+"]))
        else:
            lines.append(sep.join([name, "src", "exon", str(a), str(b), ".", "+", ".", "id=%d" % i]))
    return "\n".join(lines) + "\n"


def main():
    outdir = os.path.join(HERE, "regions")
    os.makedirs(outdir, exist_ok=True)
    wd = tempfile.mkdtemp(prefix="gt16_region_golden_")
    idx, _, _ = T.regenerate("stress_small", wd)
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_RegionIndexBuilder")
    for fmt, seed in (("bed", 3), ("gff", 4)):
        src = os.path.join(outdir, "targets." + fmt)
        open(src, "w").write(interval_list(fmt, seed))
        out = os.path.join(outdir, "targets." + fmt + ".region.bin")
        subprocess.run([ref, idx, src, out, "-" + fmt], check=True, stdout=subprocess.DEVNULL)
        print(fmt, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
