"""Helpers shared by the tests: paths, regeneration of the golden configs, .txt comparison and
ctypes bindings of the C ABI (include/gt16.h) and of the test-only checkers."""
import ctypes
import gzip
import hashlib
import json
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "16gt_b200")
LIB = os.path.join(PKG, "lib", "libgt16.so")
CLI = os.path.join(PKG, "bin", "snapshotSnpcaller")
SYNTH = os.path.join(ROOT, "tools", "gt16_synth")
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")
REF_OPS = os.path.join(REF_DIR, "libref16gt_ops.so")
REF_CALLER = os.path.join(REF_DIR, "ref_caller")
ORACLE_LIB = os.path.join(ROOT, "oracle", "libgt16_oracle.so")
EMU_LIB = os.path.join(ROOT, "tests", "host_emu", "libgt16_emu.so")

from golden_configs import CONFIGS, CONFIGS_ALL, synth_command  # noqa: E402
import importlib, sys
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
capi_mod = importlib.import_module("16gt_b200.capi")
capi, Caller, LoadedInput, call_to_txt, Stats, View = (capi_mod.capi, capi_mod.Caller, capi_mod.LoadedInput,
                                                      capi_mod.call_to_txt, capi_mod.Stats, capi_mod.View)

G2I = np.array([0, 1, 2, 3, 9, 10, 11, 18, 19, 27, 4, 12, 20, 28, 36, 37], dtype=np.uint32)


def sha256(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def manifest():
    return json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))


def config_ini(name):
    """(BalanceSubError, UnbalanceSubError) the config is called with"""
    ini = CONFIGS_ALL[name].get("ini") or {}
    return float(ini.get("BalanceSubError", 0.01)), float(ini.get("UnbalanceSubError", 0.005))


def golden_txt(name):
    return gzip.decompress(open(os.path.join(GOLDEN, name + ".txt.gz"), "rb").read())


def regenerate(name, workdir, check=True):
    """Re-create config `name` with gt16_synth; verify the inputs are the ones the golden was made from."""
    argv, idx, out, region = synth_command(SYNTH, name, str(workdir))
    subprocess.run(argv, check=True, stderr=subprocess.DEVNULL)
    if check:
        want = manifest()[name]["inputs"]
        for p in [out + ".snapshot", idx + ".pac", idx + ".ann", idx + ".tra"] + ([region] if region else []):
            assert sha256(p) == want[os.path.basename(p)], "regenerated input differs from golden input: " + p
    return idx, out, region


def callset_digest(txt: bytes):
    """sha256 over the columns that must be byte-identical (everything except the floating-point
    columns): position, alleles, genotype, indel patterns, counts, integer window features."""
    h = hashlib.sha256()
    rows = 0
    for line in txt.split(b"\n")[1:]:
        if not line:
            continue
        f = line.split(b"\t")
        h.update(b"\t".join(x for c, x in enumerate(f, start=1) if c not in FLOAT_COLS))
        h.update(b"\n")
        rows += 1
    return h.hexdigest(), rows


# ---- .txt comparison ------------------------------------------------------------------------
# 1-based column numbers of SURVEY Appendix C
FLOAT_COLS = set(range(5, 21)) | {22, 23, 24, 25} | set(range(48, 64)) | set(range(64, 88)) | {88, 89, 90, 93, 94, 95}


def compare_txt(got: bytes, want: bytes, rel=1e-6):
    """Call-set columns (position, alleles, genotype, indel patterns, counts, integer features) must
    be byte-identical; floating-point columns are printed with 5 significant digits ('%.4e'/'%.4f'),
    so two values that agree to `rel` may still print one unit apart in the last digit: accept
    |a-b| <= 1.001 units of the last printed digit and report how many fields were not identical."""
    gl, wl = got.split(b"\n"), want.split(b"\n")
    assert gl[0] == wl[0], "header differs"
    assert len(gl) == len(wl), "row count differs: %d vs %d" % (len(gl) - 2, len(wl) - 2)
    nonident = 0
    for r, (g, w) in enumerate(zip(gl[1:], wl[1:]), start=1):
        if g == w:
            continue
        gf, wf = g.split(b"\t"), w.split(b"\t")
        assert len(gf) == len(wf), "row %d: column count differs" % r
        for c, (a, b) in enumerate(zip(gf, wf), start=1):
            if a == b:
                continue
            assert c in FLOAT_COLS, "row %d col %d: %r != %r" % (r, c, a, b)
            assert a != b"*" and b != b"*", "row %d col %d: %r != %r" % (r, c, a, b)
            x, y = float(a), float(b)
            if b"e" in b:
                mant_exp = int(b.split(b"e")[1])
                ulp = 10.0 ** (mant_exp - 4)
            else:
                ulp = 1e-4
            assert abs(x - y) <= 1.001 * ulp + rel * abs(y), "row %d col %d: %r != %r" % (r, c, a, b)
            nonident += 1
    return nonident


