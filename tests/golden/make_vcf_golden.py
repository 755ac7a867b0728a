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


def main():
    for name in ("wgs_small", "stress_small"):
        wd = tempfile.mkdtemp(prefix="gt16_vcf_golden_")
        argv, idx, out, region = synth_command(T.SYNTH, name, wd)
        fa = os.path.join(wd, name + ".fa")
        subprocess.run(argv + ["--fasta", fa], check=True, stderr=subprocess.DEVNULL)
        txt = os.path.join(wd, name + ".txt")
        open(txt, "wb").write(T.golden_txt(name))
        r = subprocess.run(["perl", "/root/reference/txt2vcf.pl", txt, SAMPLE, fa], check=True, stdout=subprocess.PIPE)
        with open(os.path.join(HERE, name + ".vcf.gz"), "wb") as f:
            f.write(gzip.compress(r.stdout, 9, mtime=0))
        print(name, r.stdout.count(b"\n"), "lines")


if __name__ == "__main__":
    main()
