#!/usr/bin/env python
"""Generate tests/golden/*.vcf.gz: the VCF the UNMODIFIED reference script txt2vcf.pl (perl + Inline::C, as it builds in
this container) writes for the golden .txt files of wgs_small and stress_small, with the FASTA tools/gt16_synth writes
for the same config (--fasta).  16gt_b200/bin/txt2vcf must reproduce them byte for byte (tests/test_txt2vcf.py).

Run in the build container:  python tests/golden/make_vcf_golden.py
"""
import gzip, os, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gt16_testlib as T  # noqa: E402
from golden_configs import synth_command  # noqa: E402

SAMPLE = "SAMPLE1"
FILTER_ARGS = ["30", "0.25", "0.2", "2.5", "20", "5", "4", "2", "26", "22"]


def fake_dbsnp(vcf_text, seed):
    """a seeded stand-in for a dbSNP VCF: ~30 % of the called sites get an rs id (first alt only, or all alts under one
    id), a few positions get an entry whose allele does not match"""
    import random
    rnd = random.Random(seed)
    out = ["##fileformat=VCFv4.1", "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO"]
    k = 0
    for l in vcf_text.split("\n"):
        if not l or l.startswith("#"):
            continue
        r = l.split("\t")
        if rnd.random() < 0.3:
            k += 1
            alt = r[4].split(",")[0] if rnd.random() < 0.7 else r[4]
            out.append("\t".join([r[0], r[1], "rs%d" % k, r[3], alt, ".", ".", "."]))
        elif rnd.random() < 0.05:
            out.append("\t".join([r[0], r[1], "rsX%d" % k, r[3], "N", ".", ".", "."]))
    return "\n".join(out) + "\n"


def main():
    for name in ("wgs_small", "stress_small"):
        wd = tempfile.mkdtemp(prefix="gt16_vcf_golden_")
        argv, idx, out, region = synth_command(T.SYNTH, name, wd)
        fa = os.path.join(wd, name + ".fa")
        subprocess.run(argv + ["--fasta", fa], check=True, stderr=subprocess.DEVNULL)
        txt = os.path.join(wd, name + ".txt")
        open(txt, "wb").write(T.golden_txt(name))
        r = subprocess.run(["perl", "/root/reference/txt2vcf.pl", txt, SAMPLE, fa], check=True, stdout=subprocess.PIPE, cwd=wd)   # perl Inline caches its build in the cwd
        with open(os.path.join(HERE, name + ".vcf.gz"), "wb") as f:
            f.write(gzip.compress(r.stdout, 9, mtime=0))
        print(name, r.stdout.count(b"\n"), "lines")
        # filterVCF.pl on that VCF with a seeded stand-in for dbSNP (tests/test_txt2vcf.py: fake_dbsnp), defaults and
        # non-default thresholds
        vcf = os.path.join(wd, name + ".vcf")
        open(vcf, "wb").write(r.stdout)
        db = os.path.join(wd, name + ".db.vcf")
        open(db, "w").write(fake_dbsnp(r.stdout.decode(), 5))
        for tag, extra in (("filtered", []), ("filtered2", FILTER_ARGS)):
            f = subprocess.run(["perl", "/root/reference/filterVCF.pl", vcf, db] + extra, check=True, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL)
            with open(os.path.join(HERE, "%s.%s.vcf.gz" % (name, tag)), "wb") as g:
                g.write(gzip.compress(f.stdout, 9, mtime=0))


if __name__ == "__main__":
    main()
