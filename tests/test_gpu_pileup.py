"""SURVEY 8(f)3, the producer in front of the path: 16gt_b200/bin/bam2snapshot (SAM -> snapshot, counters built on the
GPU by gt16_pileup.cu) against the reference's own bam2snapshot (oracle/_ref/ref_bam2snapshot, unmodified) on the same
reads.  The bar: the DECODED snapshots are equal -- invalid bits, overflow-type positions, per-base weight /
forward / reverse totals, insertion and deletion pattern lists in order -- and the caller turns ours into the very .txt
the reference pipeline produced (tests/golden/prod_*.txt.gz)."""
import os
import subprocess

import numpy as np
import pytest

import gt16_testlib as T

pytestmark = pytest.mark.gpu
B2S = os.path.join(T.PKG, "bin", "bam2snapshot")


@pytest.fixture(scope="module")
def built(tmp_path_factory):
    made = {}

    def make(name):
        if name not in made:
            for exe in (T.REF_MKINDEX, T.REF_BAM2SNAPSHOT, T.READSIM):
                if not os.path.exists(exe):
                    pytest.skip("%s not built" % os.path.relpath(exe, T.ROOT))
            wd = str(tmp_path_factory.mktemp(name))
            idx, out = T.producer_pipeline(name, wd, keep_sam=True)          # the reference's tools
            sam = os.path.join(wd, name + ".sam")
            mine = os.path.join(wd, "mine")
            r = subprocess.run([B2S, "-i", idx, "-s", sam, "-o", mine], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1200)
            assert r.returncode == 0, r.stderr + r.stdout
            for marker in ("Loading reference...", "Updating snapshot...", "Total number of alignments:", "Dealing with overflows...",
                           "Writing snapshot..."):
                assert marker in r.stdout, marker
            os.remove(sam)
            made[name] = (idx, out, mine, r.stdout)
        return made[name]
    return make


@pytest.mark.parametrize("name", list(T.PRODUCER_CONFIGS))
def test_decoded_snapshot_equals_the_reference_producers(built, name):
    idx, out, mine, log = built(name)
    ref = T.decode_snapshot(out + ".snapshot")
    got = T.decode_snapshot(mine + ".snapshot")
    assert got["L"] == ref["L"] and got["nreg"] == ref["nreg"] and got["mode"] == ref["mode"]
    assert np.array_equal(got["invalid"], ref["invalid"]), "invalid bitmaps differ at %s" % np.nonzero(got["invalid"] != ref["invalid"])[0][:10]
    ok = ~ref["invalid"]                                # what an invalid position holds is whatever was there when it saturated
    assert np.array_equal(got["overflow"][ok], ref["overflow"][ok]), "overflow-type positions differ at %s" % np.nonzero((got["overflow"] != ref["overflow"]) & ok)[0][:10]
    for k in ("W", "F", "R"):
        bad = np.nonzero((got[k] != ref[k]).any(axis=1) & ok)[0]
        assert len(bad) == 0, "%s differs at %s: %s vs %s" % (k, bad[:5], got[k][bad[:5]], ref[k][bad[:5]])
    for k in ("ins", "dele"):
        a = {p: v for p, v in got[k].items() if ok[p]}
        b = {p: v for p, v in ref[k].items() if ok[p]}
        assert sorted(a) == sorted(b), "%s positions differ: %s" % (k, sorted(set(a) ^ set(b))[:10])
        diff = [p for p in a if a[p] != b[p]]
        assert not diff, "%s lists differ at %s: %s vs %s" % (k, diff[:3], [a[p] for p in diff[:3]], [b[p] for p in diff[:3]])
    assert len(got["ins"]) + len(got["dele"]) > 100 and got["overflow"].sum() > 1000      # the config does exercise the rules


@pytest.mark.parametrize("name", list(T.PRODUCER_CONFIGS))
def test_caller_output_from_our_snapshot_equals_the_reference_pipelines(built, name):
    """reads -> OUR bam2snapshot -> OUR snapshotSnpcaller  ==  reads -> reference bam2snapshot -> reference caller"""
    idx, out, mine, log = built(name)
    r = subprocess.run([T.CLI, "-i", idx, "-o", mine], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert "Called Entries : %d" % T.manifest()[name]["rows"] in r.stdout
    assert T.compare_txt(open(mine + ".txt", "rb").read(), T.golden_txt(name)) == 0
    # and the reference's own reader (readSnpInfoSnapshot), producer loop and .txt writer accept the container we wrote:
    # its harness with the libgt16 operator shim in place of FisherExactTest.o (no 26 s CPU prefill)
    exe = os.path.join(T.REF_DIR, "ref_caller_gt16")
    if name == "prod_5mb" and os.path.exists(exe):
        os.remove(mine + ".txt")
        r = subprocess.run([exe, idx, mine], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
        assert r.returncode == 0, r.stderr[-2000:]
        assert T.compare_txt(open(mine + ".txt", "rb").read(), T.golden_txt(name)) == 0


def _decoded_equal(a, b):
    ok = ~a["invalid"]
    assert np.array_equal(a["invalid"], b["invalid"]) and np.array_equal(a["overflow"][ok], b["overflow"][ok])
    for k in ("W", "F", "R"):
        assert np.array_equal(a[k][ok], b[k][ok]), k
    for k in ("ins", "dele"):
        assert {p: v for p, v in a[k].items() if ok[p]} == {p: v for p, v in b[k].items() if ok[p]}, k


def test_bam_input_and_several_files(tmp_path):
    """-b: the same alignments as BAM (BGZF inflated with zlib, records as bam_read1 hands them over); and two input
    files (-s first half, -b second half) processed in order, like the reference's loop over its inputs
    (bam2snapshot.cpp:263-293).  The reference reads the very same BAM (its samtools layer) on the way."""
    for exe in (T.REF_MKINDEX, T.REF_BAM2SNAPSHOT, T.READSIM):
        if not os.path.exists(exe):
            pytest.skip("%s not built" % os.path.relpath(exe, T.ROOT))
    pre = str(tmp_path / "p")
    subprocess.run([T.READSIM, "--out", pre, "--contigs", "chrA:300000,chrB:100000", "--depth", "20", "--seed", "5"], check=True,
                   stderr=subprocess.DEVNULL)
    subprocess.run([T.REF_MKINDEX, pre + ".fa", pre + ".index"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    T.sam_to_bam(pre + ".sam", pre + ".bam")
    lines = open(pre + ".sam").read().split("\n")
    head = [l for l in lines if l.startswith("@")]
    body = [l for l in lines if l and not l.startswith("@")]
    open(pre + ".1.sam", "w").write("\n".join(head + body[:len(body) // 2]) + "\n")
    open(pre + ".2.sam", "w").write("\n".join(head + body[len(body) // 2:]) + "\n")
    T.sam_to_bam(pre + ".2.sam", pre + ".2.bam")
    subprocess.run([T.REF_BAM2SNAPSHOT, "-i", pre + ".index", "-b", pre + ".bam", "-o", pre + ".ref"], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    ref = T.decode_snapshot(pre + ".ref.snapshot")
    for tag, args in (("bam", ["-b", pre + ".bam"]), ("two", ["-s", pre + ".1.sam", "-b", pre + ".2.bam"])):
        r = subprocess.run([B2S, "-i", pre + ".index"] + args + ["-o", pre + "." + tag], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr
        assert r.stdout.count("Processing ") == len(args) // 2
        _decoded_equal(ref, T.decode_snapshot(pre + "." + tag + ".snapshot"))


def test_deep_tile_with_more_indel_events_than_the_shared_memory_list(tmp_path):
    """an 8,000x hotspot over planted indels: the tile reports far more than 2,048 indel events, is left out by the first
    launch and done by the second with its event list in global memory -- same decoded snapshot as the reference's"""
    for exe in (T.REF_MKINDEX, T.REF_BAM2SNAPSHOT, T.READSIM):
        if not os.path.exists(exe):
            pytest.skip("%s not built" % os.path.relpath(exe, T.ROOT))
    pre = str(tmp_path / "p")
    subprocess.run([T.READSIM, "--out", pre, "--contigs", "chrA:60000", "--depth", "10", "--indel-rate", "1e-2", "--hotspots", "1:300:8000",
                    "--seed", "3"], check=True, stderr=subprocess.DEVNULL)
    subprocess.run([T.REF_MKINDEX, pre + ".fa", pre + ".index"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    subprocess.run([T.REF_BAM2SNAPSHOT, "-i", pre + ".index", "-s", pre + ".sam", "-o", pre + ".ref"], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    r = subprocess.run([B2S, "-i", pre + ".index", "-s", pre + ".sam", "-o", pre + ".mine"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    import re
    m = re.search(r"\((\d+) deep tiles in a second launch\)", r.stdout)
    assert m and int(m.group(1)) >= 1, r.stdout
    ref, got = T.decode_snapshot(pre + ".ref.snapshot"), T.decode_snapshot(pre + ".mine.snapshot")
    assert max(sum(h + l for _, _, h, l in v) for v in got["ins"].values()) + max(sum(h + l for _, h, l in v) for v in got["dele"].values()) > 2048
    _decoded_equal(ref, got)


def test_read_records_are_checked_before_the_kernel_indexes_by_them():
    """gt16_build_snapshot straight through the C ABI: hand-made read records give the expected totals, and records that
    point outside the arrays (or whose CIGAR is longer than the read) are refused with GT16_EINVAL, not run."""
    import ctypes
    L = T.capi()
    M = T.capi_mod
    glen = 6000
    # read 0: forward, 20M at 100; read 1: reverse, 8M2I10M at 3000 (trim 5 at both ends of each)
    start = np.array([100, 3000], np.uint32); rev = np.array([0, 1], np.uint8); length = np.array([20, 20], np.uint32)
    seq_off = np.array([0, 20], np.uint64); cig_off = np.array([0, 1], np.uint64); n_cig = np.array([1, 3], np.uint32)
    bases = np.tile(np.array([0, 1, 2, 3], np.uint8), 10); quals = np.full(40, 30, np.uint8)
    cigar = np.array([20 << 4, 8 << 4, (2 << 4) | 1, 10 << 4], np.uint32)

    def reads_of(**over):
        a = dict(start=start, rev=rev, length=length, seq_off=seq_off, cig_off=cig_off, n_cig=n_cig, bases=bases, quals=quals, cigar=cigar)
        a.update(over)
        r = M.Reads()
        r.n_reads = len(a["start"]); r.n_bases = len(a["bases"]); r.n_cigar_ops = len(a["cigar"])
        for f, k in (("start", "start"), ("reverse", "rev"), ("length", "length"), ("seq_off", "seq_off"), ("cigar_off", "cig_off"),
                     ("n_cigar", "n_cig"), ("bases", "bases"), ("quals", "quals"), ("cigar", "cigar")):
            setattr(r, f, a[k].ctypes.data)
        return r, a                       # the arrays stay alive with the struct

    p = M.PileupParams()
    L.gt16_pileup_default_params(ctypes.byref(p))
    with T.Caller(0) as c:
        r, keep = reads_of()
        b = M.BuiltSnapshot()
        c.check(L.gt16_build_snapshot(c.ctx, ctypes.byref(r), glen, ctypes.byref(p), ctypes.byref(b)))
        cnt = np.ctypeslib.as_array(ctypes.cast(b.view.counters, ctypes.POINTER(ctypes.c_uint64)), (glen,)).copy()
        covered = np.nonzero(cnt)[0]
        # 10 applied bases of read 0 (positions 105..114); read 1: bases 5..7 before the insertion, 10..14 after it; its
        # insertion makes position 3008 overflow-type, which is a non-zero counter as well
        assert list(covered) == list(range(105, 115)) + list(range(3005, 3008)) + list(range(3008, 3013)), covered
        assert b.view.overflow_size and b.aligned_bases == 40
        L.gt16_built_snapshot_free(ctypes.byref(b))
        bad = [dict(start=np.array([100, glen], np.uint32)),                              # a read starting behind the genome
               dict(seq_off=np.array([0, 21], np.uint64)),                                # bases beyond the array
               dict(seq_off=np.array([0, 2 ** 63], np.uint64)),
               dict(cig_off=np.array([0, 2], np.uint64)),                                 # CIGAR operations beyond the array
               dict(n_cig=np.array([1, 0], np.uint32)),                                   # no CIGAR
               dict(length=np.array([20, 0], np.uint32)),                                 # no bases
               dict(cigar=np.array([20 << 4, 9 << 4, (2 << 4) | 1, 10 << 4], np.uint32)),   # CIGAR longer than the read
               dict(cigar=np.array([20 << 4, 8 << 4, (2 << 4) | 9, 10 << 4], np.uint32))]   # unknown operation
        for over in bad:
            r, keep = reads_of(**over)
            rc = L.gt16_build_snapshot(c.ctx, ctypes.byref(r), glen, ctypes.byref(p), ctypes.byref(b))
            assert rc == -1, (over, rc)
            assert b"read 1" in L.gt16_last_error(c.ctx)
        # and the context is still good for a correct call
        r, keep = reads_of()
        c.check(L.gt16_build_snapshot(c.ctx, ctypes.byref(r), glen, ctypes.byref(p), ctypes.byref(b)))
        L.gt16_built_snapshot_free(ctypes.byref(b))
