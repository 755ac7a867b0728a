#!/usr/bin/env python
"""snapshot file -> .txt -> VCF with the two drop-in binaries on the bench workload (64,444,167-bp contig, 30x): wall
times on the GPU box, and -- where the reference script and perl exist -- the VCF compared with txt2vcf.pl's.

    python tools/cli_vcf_check.py [workdir]
"""
import hashlib, os, shutil, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    wd = sys.argv[1] if len(sys.argv) > 1 else "/tmp/gt16_vcf"
    os.makedirs(wd, exist_ok=True)
    g, s = os.path.join(wd, "g"), os.path.join(wd, "s")
    subprocess.run([os.path.join(ROOT, "tools", "gt16_synth"), "--index", g, "--snapshot", s, "--fasta", g + ".fa",
                    "--contigs", "chr20:64444167", "--seed", "1", "--threads", str(min(32, os.cpu_count() or 8))],
                   check=True, stderr=subprocess.DEVNULL)
    t0 = time.time()
    r = subprocess.run([os.path.join(ROOT, "16gt_b200", "bin", "snapshotSnpcaller"), "-i", g, "-o", s], stdout=subprocess.PIPE, text=True)
    t1 = time.time()
    assert r.returncode == 0, r.stdout[-2000:]
    print("snapshotSnpcaller: %.2f s wall;" % (t1 - t0), [l for l in r.stdout.splitlines() if "Called Entries" in l][0])
    with open(s + ".vcf", "wb") as f:
        t0 = time.time()
        subprocess.run([os.path.join(ROOT, "16gt_b200", "bin", "txt2vcf"), s + ".txt", "SAMPLE", g + ".fa"], stdout=f, check=True)
        t1 = time.time()
    vcf = open(s + ".vcf", "rb").read()
    print("txt2vcf: %.2f s wall, %d VCF rows, sha256 %s" % (t1 - t0, sum(1 for l in vcf.split(b"\n") if l and not l.startswith(b"#")),
                                                             hashlib.sha256(vcf).hexdigest()[:16]))
    if os.path.exists("/root/reference/txt2vcf.pl") and shutil.which("perl"):
        t0 = time.time()
        p = subprocess.run(["perl", "/root/reference/txt2vcf.pl", s + ".txt", "SAMPLE", g + ".fa"], stdout=subprocess.PIPE)
        print("txt2vcf.pl: %.2f s wall; output %s" % (time.time() - t0, "IDENTICAL" if p.stdout == vcf else "DIFFERENT"))
    shutil.rmtree(wd, ignore_errors=True)


if __name__ == "__main__":
    main()
