"""16gt_b200/bin/txt2vcf -- the native replacement of the reference's txt2vcf.pl -- against the reference
script itself: byte-identical VCF for the golden .txt files (tests/golden/*.vcf.gz, written by the unmodified Perl script,
tests/golden/make_vcf_golden.py), whatever the number of worker threads; and, where perl and /root/reference exist (the
build container), live on a larger seeded call set that exercises every branch of the script (hom / het / two-alt SNPs,
XX, ref+X, base+X -- two rows --, XY with every insertion / deletion combination, rows skipped by the reference check)."""
import ctypes
import gzip
import os
import shutil
import subprocess

import pytest

import gt16_testlib as T
from golden_configs import synth_command

BIN = os.path.join(T.PKG, "bin", "txt2vcf")
FILTER = os.path.join(T.PKG, "bin", "filterVCF")
REF_FILTER = "/root/reference/filterVCF.pl"
REF_SCRIPT = "/root/reference/txt2vcf.pl"


def _inputs(name, wd):
    argv, idx, out, region = synth_command(T.SYNTH, name, str(wd))
    fa = os.path.join(str(wd), name + ".fa")
    subprocess.run(argv + ["--fasta", fa], check=True, stderr=subprocess.DEVNULL)
    txt = os.path.join(str(wd), name + ".txt")
    open(txt, "wb").write(T.golden_txt(name))
    return txt, fa


@pytest.mark.parametrize("name", ["wgs_small", "stress_small"])
def test_vcf_matches_reference_script(tmp_path, name):
    txt, fa = _inputs(name, tmp_path)
    want = gzip.decompress(open(os.path.join(T.GOLDEN, name + ".vcf.gz"), "rb").read())
    for threads in ("1", "5"):
        r = subprocess.run([BIN, txt, "SAMPLE1", fa], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                           env=dict(os.environ, GT16_VCF_THREADS=threads))
        assert r.returncode == 0, r.stderr
        assert r.stdout == want
    # gzip'ed input, like the script (it reads through `gzip -dcf`)
    with open(txt, "rb") as f, gzip.open(txt + ".gz", "wb") as g:
        shutil.copyfileobj(f, g)
    r = subprocess.run([BIN, txt + ".gz", "SAMPLE1", fa], stdout=subprocess.PIPE)
    assert r.stdout == want
    # usage
    r = subprocess.run([BIN, txt], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode != 0 and "<txt Input> <Sample Name>" in r.stderr


@pytest.mark.skipif(not (os.path.exists(REF_SCRIPT) and shutil.which("perl")), reason="needs perl and /root/reference (build container)")
def test_vcf_matches_reference_script_live(tmp_path):
    wd = str(tmp_path)
    subprocess.run([T.SYNTH, "--index", wd + "/v.idx", "--snapshot", wd + "/v", "--fasta", wd + "/v.fa", "--genome-len", "1500000",
                    "--n-contigs", "3", "--depth", "40", "--snp-rate", "5e-3", "--indel-rate", "6e-3", "--multi-pattern-frac", "0.7",
                    "--long-indel-frac", "0.2", "--hotspot-rate", "2e-4", "--hotspot-depth", "1500", "--seed", "31"],
                   check=True, stderr=subprocess.DEVNULL)
    emu = ctypes.CDLL(T.EMU_LIB)      # the product's per-site source compiled for the host (tests/host_emu): writes the .txt
    emu.emu_run.argtypes = [ctypes.c_char_p] * 3 + [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_char_p]
    assert emu.emu_run((wd + "/v.idx").encode(), (wd + "/v.snapshot").encode(), b"", None, 0.01, 0.005, (wd + "/v.txt").encode()) == 0
    gts = [l.split("\t")[20] for l in open(wd + "/v.txt").read().split("\n")[1:] if l]
    for g in ("XX", "XY", "AX", "CX", "AC", "GG"):
        assert g in gts
    a = subprocess.run(["perl", REF_SCRIPT, wd + "/v.txt", "S", wd + "/v.fa"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=wd)   # perl Inline caches its build in the cwd
    b = subprocess.run([BIN, wd + "/v.txt", "S", wd + "/v.fa"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert a.returncode == 0 and b.returncode == 0
    assert a.stdout.count(b"\n") > 10000
    assert a.stdout == b.stdout
    # a FASTA that does not belong to the call set: the rows failing the reference check are dropped, the same ones
    _, other = _inputs("wgs_small", tmp_path)
    subprocess.run(["sed", "-i", "s/^>chrA/>chr1/", other], check=True)
    a = subprocess.run(["perl", REF_SCRIPT, wd + "/v.txt", "S", other], stdout=subprocess.PIPE, stderr=subprocess.PIPE, cwd=wd)
    b = subprocess.run([BIN, wd + "/v.txt", "S", other], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert a.stdout == b.stdout and 40 < a.stdout.count(b"\n") < 5000


def _golden_maker():
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_vcf_golden", os.path.join(T.GOLDEN, "make_vcf_golden.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("name", ["wgs_small", "stress_small"])
def test_filtervcf_matches_reference_script(tmp_path, name):
    """16gt_b200/bin/filterVCF against filterVCF.pl: same bytes for the golden VCF and a seeded stand-in for dbSNP, with
    the default and with non-default thresholds, plain, gzip'ed and .list input"""
    G = _golden_maker()
    vcf_bytes = gzip.decompress(open(os.path.join(T.GOLDEN, name + ".vcf.gz"), "rb").read())
    vcf, db = str(tmp_path / (name + ".vcf")), str(tmp_path / (name + ".db.vcf"))
    open(vcf, "wb").write(vcf_bytes)
    open(db, "w").write(G.fake_dbsnp(vcf_bytes.decode(), 5))
    for tag, extra in (("filtered", []), ("filtered2", G.FILTER_ARGS)):
        want = gzip.decompress(open(os.path.join(T.GOLDEN, "%s.%s.vcf.gz" % (name, tag)), "rb").read())
        r = subprocess.run([FILTER, vcf, db] + extra, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert r.returncode == 0, r.stderr
        assert r.stdout == want
        assert b"Mean DP: " in r.stderr and b"Second pass (add filter tags) ..." in r.stderr
    assert b"\trs" in want and b"QUALlow" in want
    with open(vcf, "rb") as f, gzip.open(vcf + ".gz", "wb") as g:
        shutil.copyfileobj(f, g)
    open(str(tmp_path / "in.list"), "w").write(vcf + ".gz\n")
    r = subprocess.run([FILTER, str(tmp_path / "in.list"), db] + G.FILTER_ARGS, stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0 and r.stdout == b""
    assert open(str(tmp_path / (name + ".filtered.vcf")), "rb").read() == want
    r = subprocess.run([FILTER], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 0 and b"<16GT VCF Input> <dbSNP VCF Input> [Min QUAL, 10]" in r.stderr
    if os.path.exists(REF_FILTER) and shutil.which("perl"):      # live, stderr protocol included
        a = subprocess.run(["perl", REF_FILTER, vcf, db], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        b = subprocess.run([FILTER, vcf, db], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
        assert a.stdout == b.stdout and a.stderr == b.stderr
