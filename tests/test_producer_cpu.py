"""Parity pinned on inputs written by the reference's OWN producer: reads from
tools/gt16_readsim go through oracle/_ref/ref_mkindex and oracle/_ref/ref_bam2snapshot (unmodified reference code:
bam2snapshot.cpp:591-716, SnapshotHandler.cpp:178-236, SNP.cpp:745-891), and the golden .txt is the reference
caller's output on that snapshot.  CPU side: the product's snapshot reader takes the writeSnpInfoSnapshot container
as it is (page-aligned subfiles, full 1 GiB pool, soft-clip subfile) and sees what an independent parse sees; the C
oracle and the product's per-site source (host build) reproduce the reference's .txt byte for byte."""
import ctypes

import numpy as np
import pytest

import gt16_testlib as T
from producer_fixture import producer_inputs

NAMES = list(T.PRODUCER_CONFIGS)


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mapped", [True, False])
def test_reader_takes_the_producers_container(name, mapped, monkeypatch):
    idx, out = producer_inputs(name)
    if not mapped:
        monkeypatch.setenv("GT16_NO_MMAP", "1")
    want = T.parse_snapshot_container(out + ".snapshot")
    assert want["pool_capacity"] == 1 << 30 and want["mode"] == 1            # full pool, dense counters
    assert all(int(s) % want["page"] == 0 for s in want["subfiles"][:, 0])   # page-aligned subfiles
    assert int(want["subfiles"][4, 0]) > 0                                   # the soft-clip subfile is there (unused by the caller)
    inp = T.LoadedInput(idx, out + ".snapshot", None, pinned=0)
    try:
        v = inp.view
        L = v.text_length
        assert L == want["L"] and v.pool_bytes == len(want["pool"]) and v.n_overflow_regions == len(want["ovf"])
        got = np.ctypeslib.as_array(ctypes.cast(v.counters, ctypes.POINTER(ctypes.c_uint64)), (L,))
        assert np.array_equal(got, want["counters"])
        bm = np.ctypeslib.as_array(ctypes.cast(v.invalid_bitmap, ctypes.POINTER(ctypes.c_uint32)), ((L + 31) >> 5,))
        assert np.array_equal(bm, want["bitmap"])
        assert ctypes.string_at(v.pool, v.pool_bytes) == want["pool"]
        sizes = np.ctypeslib.as_array(ctypes.cast(v.overflow_size, ctypes.POINTER(ctypes.c_uint32)), (v.n_overflow_regions,))
        ptrs = ctypes.cast(v.overflow, ctypes.POINTER(ctypes.c_void_p))
        for r, blob in enumerate(want["ovf"]):
            assert int(sizes[r]) * 32 == len(blob)
            assert ctypes.string_at(ptrs[r], len(blob)) == blob
    finally:
        inp.close()


def test_producer_snapshots_carry_the_hard_cases():
    """what the synthesiser was written to imitate really occurs in producer-written input: overflow promotion,
    pooled insertion AND deletion lists, invalid bits, compact counters with two live slots"""
    idx, out = producer_inputs("prod_5mb")
    c = T.parse_snapshot_container(out + ".snapshot")
    lo = (c["counters"] & 0xFFFFFFFF).astype(np.uint32)
    ovf = ((lo & (lo >> 16)) & 0xC000) == 0xC000
    assert ovf.sum() > 10000
    assert (ovf & ((lo & 0x3000) == 0x3000)).sum() > 10            # pooled insertion lists
    assert (ovf & (((lo >> 16) & 0x3000) == 0x3000)).sum() > 3     # pooled deletion lists
    assert np.unpackbits(c["bitmap"].view(np.uint8)).sum() > 0     # invalid positions (20000x hotspot)
    hi = (c["counters"] >> 32).astype(np.uint32)
    assert ((~ovf) & ((hi & 0xFFFF) != 0) & ((hi >> 16) != 0)).sum() > 100000


@pytest.mark.parametrize("name", NAMES)
def test_oracle_on_producer_snapshot(name, tmp_path):
    idx, out = producer_inputs(name)
    O = ctypes.CDLL(T.ORACLE_LIB)
    O.oracle_run.argtypes = [ctypes.c_char_p] * 3 + [ctypes.c_double, ctypes.c_double, ctypes.c_char_p,
                                                     ctypes.POINTER(ctypes.c_ulonglong)]
    n = ctypes.c_ulonglong()
    txt = str(tmp_path / "oracle.txt")
    assert O.oracle_run(idx.encode(), (out + ".snapshot").encode(), b"", 0.01, 0.005, txt.encode(), ctypes.byref(n)) == 0
    assert n.value == T.manifest()[name]["rows"]
    assert open(txt, "rb").read() == T.golden_txt(name)


@pytest.mark.parametrize("name", NAMES)
def test_product_source_on_producer_snapshot(name, tmp_path):
    """gt16_core.cuh compiled for the host (tests/host_emu): the functions the kernels are made of"""
    idx, out = producer_inputs(name)
    E = ctypes.CDLL(T.EMU_LIB)
    E.emu_run.argtypes = [ctypes.c_char_p] * 3 + [ctypes.c_void_p, ctypes.c_double, ctypes.c_double, ctypes.c_char_p]
    txt = str(tmp_path / "emu.txt")
    assert E.emu_run(idx.encode(), (out + ".snapshot").encode(), b"", None, 0.01, 0.005, txt.encode()) == 0
    assert open(txt, "rb").read() == T.golden_txt(name)
