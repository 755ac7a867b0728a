"""Edge cases of the fused path on the GPU, checked against the C oracle (which is pinned on the reference's
golden vectors; the reference binary itself cannot run most of these: tiny genomes trip its fixed-size buffers
and its exact-multiple-of-a-batch hang).  Inputs come from tools/gt16_synth with odd sizes and from hand-built
region lists: genomes shorter than a window, lengths around tile / word / batch boundaries, regions that touch
the first and last position, single-site and empty regions, unsorted regions, a region list whose total is an
exact multiple of nothing in particular."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import gt16_testlib as T

pytestmark = pytest.mark.gpu


def oracle_txt(idx, snap, region, out):
    O = ctypes.CDLL(T.ORACLE_LIB)
    O.oracle_run.argtypes = [ctypes.c_char_p] * 3 + [ctypes.c_double, ctypes.c_double, ctypes.c_char_p,
                                                     ctypes.POINTER(ctypes.c_ulonglong)]
    n = ctypes.c_ulonglong()
    rc = O.oracle_run(idx.encode(), snap.encode(), (region or "").encode(), 0.01, 0.005, out.encode(), ctypes.byref(n))
    assert rc == 0
    return n.value


def synth(tmp_path, tag, args):
    idx, out = str(tmp_path / (tag + ".idx")), str(tmp_path / tag)
    subprocess.run([T.SYNTH, "--index", idx, "--snapshot", out] + args, check=True, stderr=subprocess.DEVNULL)
    return idx, out


@pytest.mark.parametrize("length", [1, 2, 31, 33, 99, 100, 101, 127, 129, 255, 4095, 4096, 4097, 8191, 12289, 70001])
def test_odd_genome_lengths(gpu_caller, tmp_path, length):
    # dense variants so that even tiny genomes call something; one contig
    idx, out = synth(tmp_path, "g%d" % length, ["--contigs", "c:%d" % length, "--snp-rate", "0.05", "--indel-rate", "0.01",
                                                 "--seed", str(1000 + length)])
    want_n = oracle_txt(idx, out + ".snapshot", None, out + ".oracle.txt")
    for mode in ("host", "resident"):
        inp = T.LoadedInput(idx, out + ".snapshot", None, pinned=0)
        st = T.call_to_txt(gpu_caller, inp, out + ".txt", resident=(mode == "resident"))
        inp.close()
        assert sum(s.called for s in st) == want_n
        T.compare_txt(open(out + ".txt", "rb").read(), open(out + ".oracle.txt", "rb").read())


def write_regions(path, regions):
    with open(path, "wb") as f:
        f.write(np.uint32(len(regions)).tobytes())
        f.write(np.array(regions, dtype=np.uint32).tobytes())


@pytest.mark.parametrize("case", ["ends", "single_sites", "empty_and_adjacent", "unsorted", "one_aligned_region", "many_small"])
def test_region_lists(gpu_caller, tmp_path, case):
    L = 300000
    idx, out = synth(tmp_path, "r", ["--contigs", "a:200000,b:100000", "--snp-rate", "0.01", "--indel-rate", "0.003",
                                     "--seed", "77"])
    regions = {
        "ends": [(0, 150), (L - 130, L)],
        "single_sites": [(5, 6), (1000, 1001), (1001, 1002), (199999, 200001), (L - 1, L)],
        "empty_and_adjacent": [(100, 100), (200, 400), (400, 700), (900, 900), (5000, 9000)],
        "unsorted": [(50000, 52000), (100, 3000), (250000, 251000), (100, 3000)],
        "one_aligned_region": [(4096, 200000)],           # single region whose start is tile-aligned: TMA kernel with pBeg != 0
        "many_small": [(k * 997, k * 997 + 1 + (k * 37) % 300) for k in range(1, 290)],
    }[case]
    region_bin = str(tmp_path / (case + ".bin"))
    write_regions(region_bin, regions)
    want_n = oracle_txt(idx, out + ".snapshot", region_bin, out + ".oracle.txt")
    inp = T.LoadedInput(idx, out + ".snapshot", region_bin)
    for resident in (False, True):
        st = T.call_to_txt(gpu_caller, inp, out + ".txt", resident=resident)
        assert sum(s.called for s in st) == want_n
        T.compare_txt(open(out + ".txt", "rb").read(), open(out + ".oracle.txt", "rb").read())
    inp.close()
    # the drop-in binary takes the same file
    r = subprocess.run([T.CLI, "-i", idx, "-o", out, "-e", region_bin], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    T.compare_txt(open(out + ".txt", "rb").read(), open(out + ".oracle.txt", "rb").read())


def test_batch_boundary_neighbourhood(gpu_caller, tmp_path):
    """a genome one site longer / shorter than one batch and exactly one batch (the reference hangs on the latter)"""
    B = 8388608
    for length in (B - 1, B, B + 1):
        idx, out = synth(tmp_path, "b%d" % length, ["--contigs", "c:%d" % length, "--snp-rate", "2e-5", "--indel-rate", "1e-5",
                                                     "--seed", "5"])
        want_n = oracle_txt(idx, out + ".snapshot", None, out + ".oracle.txt")
        inp = T.LoadedInput(idx, out + ".snapshot", None)
        st = T.call_to_txt(gpu_caller, inp, out + ".txt", shards=2)
        inp.close()
        assert sum(s.called for s in st) == want_n
        T.compare_txt(open(out + ".txt", "rb").read(), open(out + ".oracle.txt", "rb").read())
        os.remove(out + ".snapshot")


@pytest.mark.parametrize("name,mode", [("stress_small", "host"), ("stress_small", "resident"), ("stress_small", "device_nowait"),
                                       ("two_batch", "host"), ("two_batch", "resident")])       # two_batch: several segments
def test_more_candidates_than_the_record_bound(tmp_path, monkeypatch, name, mode):
    """Record buffers are sized by a bound; a call that finds more candidates is repeated once with buffers of the size it
    asked for (GT16_RECORD_BOUND shrinks the bound so that the small stress config overflows it): same .txt."""
    import ctypes
    monkeypatch.setenv("GT16_RECORD_BOUND", "1000000:24")
    monkeypatch.setenv("GT16_SEG_BATCHES", "1")
    idx, out, region = T.regenerate(name, tmp_path)
    rows = T.manifest()[name]["rows"]
    assert rows > 100
    with T.Caller(0) as c:
        inp = T.LoadedInput(idx, out + ".snapshot", region)
        try:
            if mode == "device_nowait":
                L = c.L
                dev, calls, st, host = ctypes.c_void_p(), ctypes.c_void_p(), T.Stats(), ctypes.c_void_p()
                c.check(L.gt16_snapshot_upload(c.ctx, ctypes.byref(inp.view), ctypes.byref(dev)))
                c.check(L.gt16_call_sites_device_nowait(c.ctx, dev, None, 0, 0, inp.total_sites, ctypes.byref(calls), ctypes.byref(st)))
                c.check(L.gt16_dev_calls_wait(c.ctx, calls, ctypes.byref(st)))
                assert st.called == rows == L.gt16_dev_calls_count(calls)
                c.check(L.gt16_dev_calls_fetch(c.ctx, calls, ctypes.byref(host)))
                assert L.gt16_write_txt((out + ".txt").encode(), inp.ix, L.gt16_calls_data(host), L.gt16_calls_count(host), 0, None) == 0
                L.gt16_calls_free(host); L.gt16_dev_calls_free(calls); L.gt16_snapshot_release(c.ctx, dev)
            else:
                st = T.call_to_txt(c, inp, out + ".txt", resident=(mode == "resident"))
                assert sum(x.called for x in st) == rows
        finally:
            inp.close()
    assert T.compare_txt(open(out + ".txt", "rb").read(), T.golden_txt(name)) <= 2
