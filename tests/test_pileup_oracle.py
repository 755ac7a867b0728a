"""The producer's CPU restatement (oracle/gt16_pileup_oracle.py) pinned against the reference itself: on reads made by
tools/gt16_readsim it must say exactly what the decoded snapshot of oracle/_ref/ref_bam2snapshot (the unmodified
bam2snapshot) says -- invalid positions, overflow-type positions, per-base totals, pattern lists in order.  The GPU producer
is then compared with both (tests/test_gpu_pileup.py)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

import gt16_testlib as T

sys.path.insert(0, os.path.join(T.ROOT, "oracle"))
import gt16_pileup_oracle as PO  # noqa: E402

# small enough for a Python loop over the reads; dense enough to meet every rule: SNPs, 1-30 bp indels (pooled patterns),
# compound sites, soft clips, duplicates, a 1,200x hotspot (overflow promotion by strand count) with indels under it
READSIM = ["--contigs", "chrA:50000,chrB:20000", "--depth", "25", "--snp-rate", "2e-3", "--indel-rate", "2e-3", "--long-indel-frac", "0.2",
           "--hotspots", "1:150:1200", "--seed", "11"]


def make_inputs(wd):
    for exe in (T.REF_MKINDEX, T.REF_BAM2SNAPSHOT, T.READSIM):
        if not os.path.exists(exe):
            pytest.skip("%s not built" % os.path.relpath(exe, T.ROOT))
    pre = os.path.join(str(wd), "p")
    subprocess.run([T.READSIM, "--out", pre] + READSIM, check=True, stderr=subprocess.DEVNULL)
    subprocess.run([T.REF_MKINDEX, pre + ".fa", pre + ".index"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([T.REF_BAM2SNAPSHOT, "-i", pre + ".index", "-s", pre + ".sam", "-o", pre + ".ref"], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    return pre


def oracle_of(pre):
    L = T.capi()
    err = ctypes.create_string_buffer(256)
    ix = ctypes.c_void_p()
    assert L.gt16_index_read((pre + ".index").encode(), ctypes.byref(ix), err, 256) == 0

    def locate(name, pos1):
        amb = ctypes.c_uint32()
        return amb.value if L.gt16_index_locate(ix, name.encode(), pos1, ctypes.byref(amb)) == 0 else None
    try:
        return PO.pileup(pre + ".sam", locate, L.gt16_index_length(ix))
    finally:
        L.gt16_index_free(ix)


def assert_same(a, b):
    ok = ~(a["invalid"] | b["invalid"])
    assert np.array_equal(a["invalid"], b["invalid"]), np.nonzero(a["invalid"] != b["invalid"])[0][:10]
    assert np.array_equal(a["overflow"][ok], b["overflow"][ok]), np.nonzero((a["overflow"] != b["overflow"]) & ok)[0][:10]
    for k in ("W", "F", "R"):
        bad = np.nonzero((a[k] != b[k]).any(axis=1) & ok)[0]
        assert len(bad) == 0, (k, bad[:5], a[k][bad[:5]], b[k][bad[:5]])
    for k in ("ins", "dele"):
        x = {p: v for p, v in a[k].items() if ok[p]}
        y = {p: v for p, v in b[k].items() if ok[p]}
        assert sorted(x) == sorted(y), (k, sorted(set(x) ^ set(y))[:10])
        diff = [p for p in x if x[p] != y[p]]
        assert not diff, (k, diff[:3], [x[p] for p in diff[:3]], [y[p] for p in diff[:3]])


def test_oracle_equals_the_reference_producer(tmp_path):
    pre = make_inputs(tmp_path)
    ref = T.decode_snapshot(pre + ".ref.snapshot")
    ora = oracle_of(pre)
    assert_same(ora, ref)
    assert ref["overflow"].sum() > 200 and len(ref["ins"]) > 20 and len(ref["dele"]) > 20
    assert any(len(v) > 1 for v in ref["ins"].values()) or any(len(v) > 1 for v in ref["dele"].values())     # multi-pattern lists are met


@pytest.mark.gpu
def test_gpu_producer_equals_the_oracle(tmp_path):
    pre = make_inputs(tmp_path)
    r = subprocess.run([os.path.join(T.PKG, "bin", "bam2snapshot"), "-i", pre + ".index", "-s", pre + ".sam", "-o", pre + ".mine"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    assert_same(T.decode_snapshot(pre + ".mine.snapshot"), oracle_of(pre))
