"""CPU tests of the drop-in boundary: libgt16.so loads and exports every symbol include/gt16.h declares,
refuses to compute without a GPU (there is no CPU path), and its host-side format code (index reader,
coordinate translation, snapshot reader, prior table) behaves like the reference's."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

import gt16_testlib as T


def have_gpu():
    n = ctypes.c_int()
    return T.capi().gt16_device_count(ctypes.byref(n)) == 0 and n.value > 0


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(T.ROOT, "include", "gt16.h")).read()
    declared = sorted(set(re.findall(r"\b(gt16_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 28
    lib = ctypes.CDLL(T.LIB)
    for name in declared:
        assert hasattr(lib, name), "libgt16.so does not export " + name
    bound = set(T.capi()._signatures)
    assert bound == set(declared), "ctypes binding and header disagree: %s" % (bound ^ set(declared))


def test_no_cpu_fallback():
    if have_gpu():
        pytest.skip("a GPU is present")
    L = T.capi()
    ctx = ctypes.c_void_p()
    assert L.gt16_ctx_create(0, ctypes.byref(ctx)) == -2          # GT16_ENODEVICE
    assert b"no CPU path" in L.gt16_last_error(None)
    r = subprocess.run([T.CLI, "-i", "/nonexistent/x", "-o", "/tmp/y"], stdout=subprocess.PIPE, stderr=subprocess.PIPE)
    assert r.returncode == 1


def test_prior_table_matches_reference():
    g = np.load(os.path.join(T.GOLDEN, "operator_vectors.npz"))
    prior = np.zeros(256)
    assert T.capi().gt16_create_snp_prior(prior.ctypes.data) == 0
    assert np.array_equal(prior, g["prior"])


def test_index_and_snapshot_readers(tmp_path):
    idx, out, region = T.regenerate("stress_small", tmp_path)
    inp = T.LoadedInput(idx, out + ".snapshot", None, pinned=0)
    try:
        L = T.capi()
        n = L.gt16_index_length(inp.ix)
        assert n == inp.view.text_length == 400000
        assert inp.view.n_overflow_regions == 7
        # coordinate translation: contig boundaries, 1-based positions (getChrAndPos)
        pos = ctypes.c_uint32()
        assert L.gt16_index_translate(inp.ix, 0, ctypes.byref(pos)) == b"chr1" and pos.value == 1
        ann = open(idx + ".ann").read().split("\n")
        start2 = int(ann[4].split()[0])
        assert L.gt16_index_translate(inp.ix, start2 - 1, ctypes.byref(pos)) == b"chr1" and pos.value == start2
        assert L.gt16_index_translate(inp.ix, start2, ctypes.byref(pos)) == b"chr2" and pos.value == 1
        # packed reference equals the .pac bytes re-packed 16 bases per word
        nw = ctypes.c_uint32()
        ptr = L.gt16_index_packed(inp.ix, ctypes.byref(nw))
        words = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_uint32)), (nw.value,))
        pac = np.fromfile(idx + ".pac", dtype=np.uint8)[:(n + 3) // 4]
        pad = np.zeros((-len(pac)) % 4, dtype=np.uint8)
        want = np.concatenate([pac, pad]).reshape(-1, 4).astype(np.uint32)
        want = (want[:, 0] << 24) | (want[:, 1] << 16) | (want[:, 2] << 8) | want[:, 3]
        assert np.array_equal(words[:len(want)], want)
    finally:
        inp.close()
    err = ctypes.create_string_buffer(256)
    hs = ctypes.c_void_p()
    assert T.capi().gt16_snapshot_read(b"/nonexistent.snapshot", 0, ctypes.byref(hs), err, 256) == -4
    assert b"Error opening" in err.value


@pytest.mark.parametrize("mapped", [True, False])
def test_sparse_snapshot_kept_as_records(tmp_path, mapped, monkeypatch):
    """GT16_READ_KEEP_SPARSE: the counter subfile of a mode-2 snapshot stays the {u32 position; u64 counter} records the
    producer wrote (mapped or read); scattered into an array of zeros they give exactly what the expanding reader gives"""
    if not mapped:
        monkeypatch.setenv("GT16_NO_MMAP", "1")
    idx, out, region = T.regenerate("exome_small", tmp_path)
    dense = T.LoadedInput(idx, out + ".snapshot", region, pinned=0)
    recs = T.LoadedInput(idx, out + ".snapshot", region, pinned=0, keep_sparse=True)
    try:
        n = dense.view.text_length
        assert recs.view.text_length == n and not recs.view.counters and dense.view.counters and not dense.view.n_sparse_entries
        k = recs.view.n_sparse_entries
        raw = np.ctypeslib.as_array(ctypes.cast(recs.view.sparse_entries, ctypes.POINTER(ctypes.c_uint8)), (k * 12,)).reshape(k, 12)
        pos = raw[:, :4].copy().view(np.uint32)[:, 0]
        val = raw[:, 4:].copy().view(np.uint64)[:, 0]
        assert np.all(pos[1:] > pos[:-1]) and pos[-1] < n
        want = np.ctypeslib.as_array(ctypes.cast(dense.view.counters, ctypes.POINTER(ctypes.c_uint64)), (n,))
        got = np.zeros(n, dtype=np.uint64)
        got[pos] = val
        assert np.array_equal(got, want)
        # a dense snapshot is read as before whatever the flag
        idx2, out2, _ = T.regenerate("wgs_small", tmp_path)
        d2 = T.LoadedInput(idx2, out2 + ".snapshot", None, pinned=0, keep_sparse=True)
        assert d2.view.counters and not d2.view.sparse_entries
        d2.close()
    finally:
        dense.close(); recs.close()


def test_region_index_builder_matches_reference(tmp_path):
    """RegionIndexBuilder (buildRegionList.cpp): byte-identical region.bin for the seeded BED and GFF lists of
    tests/golden/regions (made by the unmodified reference tool, tests/golden/make_region_golden.py); plus the
    coordinate lookup it rests on (getChrIDFromName / getAmbPos, indexFunction.cpp:6-66)."""
    idx, out, region = T.regenerate("stress_small", tmp_path)
    rib = os.path.join(T.PKG, "bin", "RegionIndexBuilder")
    for fmt in ("bed", "gff"):
        src = os.path.join(T.GOLDEN, "regions", "targets." + fmt)
        got = str(tmp_path / ("region." + fmt + ".bin"))
        r = subprocess.run([rib, idx, src, got, "-" + fmt], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
        assert r.returncode == 0, r.stderr
        want = open(src + ".region.bin", "rb").read()
        assert open(got, "rb").read() == want
        n = int(np.frombuffer(want[:4], dtype=np.uint32)[0])
        assert "Removed Overlapped entries : %u remained" % n in r.stdout
        reg = np.frombuffer(want[4:], dtype=np.uint32).reshape(n, 2)
        assert np.all(reg[1:, 0] >= reg[:-1, 0])
    r = subprocess.run([rib, idx, src, got], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Usage:" in r.stderr
    r = subprocess.run([rib, idx, src, got, "-vcf"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Invalid input format" in r.stdout
    L = T.capi()
    err = ctypes.create_string_buffer(256)
    ix = ctypes.c_void_p()
    assert L.gt16_index_read(idx.encode(), ctypes.byref(ix), err, 256) == 0
    try:
        amb, pos = ctypes.c_uint32(), ctypes.c_uint32()
        assert L.gt16_index_locate(ix, b"chr2", 1, ctypes.byref(amb)) == 0 and amb.value == 222222
        assert L.gt16_index_locate(ix, b"chrchr2", 5, ctypes.byref(amb)) == 0 and amb.value == 222226   # one "chr" is dropped
        assert L.gt16_index_locate(ix, b"2", 5, ctypes.byref(amb)) == -1           # ... but never added (:54-63)
        assert L.gt16_index_locate(ix, b"chrX", 5, ctypes.byref(amb)) == -1
        assert L.gt16_index_locate(ix, b"chr1", 222223, ctypes.byref(amb)) == -1       # beyond the sequence
        if not have_gpu():                                                            # page-locking needs a device: reported, not faked
            assert L.gt16_index_pin(ix) == -3                                         # GT16_ECUDA
        for a in (0, 1, 222221, 222222, 399999):                                      # locate inverts translate
            name = L.gt16_index_translate(ix, a, ctypes.byref(pos))
            assert L.gt16_index_locate(ix, name, pos.value, ctypes.byref(amb)) == 0 and amb.value == a
    finally:
        L.gt16_index_free(ix)


def test_fast_number_formatting_equals_printf():
    """The .txt emitter formats "%.4e" / "%.4f" fields itself (128-bit integer scaling, snprintf whenever the value lies close
    to a rounding boundary): every byte must be what libc's printf writes -- random bit patterns over all exponents,
    subnormals, values next to d.dddd5 boundaries and next to powers of ten, exact ties of the fixed format."""
    import math, random, struct
    L = T.capi()
    libc = ctypes.CDLL("libc.so.6")
    libc.snprintf.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_double]
    got, want = ctypes.create_string_buffer(128), ctypes.create_string_buffer(128)
    rnd = random.Random(11)

    def check(v, kind):
        n = L.gt16_format_number(v, kind, got)
        m = libc.snprintf(want, 128, b"%.4e" if kind == 0 else b"%.4f", v)
        assert got.raw[:n] == want.raw[:m], (v, kind, got.raw[:n], want.raw[:m])

    for _ in range(60000):
        v = struct.unpack("<d", struct.pack("<Q", rnd.getrandbits(64)))[0]
        if math.isfinite(v):
            check(v, 0)
        check(10 ** rnd.uniform(-320, 0.5), 0)
        k = rnd.randint(-300, 10)
        v = (rnd.randint(10000, 99999) + 0.5) * 10.0 ** (k - 4)
        for w in (v, math.nextafter(v, 0), math.nextafter(v, math.inf)):
            check(w, 0)
        f = struct.unpack("<f", struct.pack("<f", rnd.uniform(-3, 3) * 10 ** rnd.uniform(-6, 5)))[0]
        check(f, 1)
    for k in range(-320, 300):
        for m in (9.99995, 9.999949999, 9.99995000001, 1.0, 0.99999999999, 1.00005):
            check(m * 10.0 ** k, 0)
    for v in (0.0, -0.0, 0.5, 0.00005, 0.125, 1.00005, 2.5e-5, -2.5e-5, 1e-9, -1e-9, 399999.99996, 4e5, 1e10, 0.99995,
              5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, float("inf"), float("nan")):
        check(v, 0)
        if not abs(v) > 3.5e38:                                                # "%.4f" fields are floats
            check(v, 1)
    check(3.4028234663852886e38, 1)
    for i in range(1, 40000):
        check(i / 32768.0, 1)                                                  # exact ties
        check(struct.unpack("<f", struct.pack("<f", i * 0.00005))[0], 1)


def test_cli_usage_errors():
    r = subprocess.run([T.CLI], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Usage:" in r.stderr and "-e <Exome Region Index>" in r.stderr
    r = subprocess.run([T.CLI, "-i", "a", "-i", "b"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Multiple indexes" in r.stderr
    r = subprocess.run([T.CLI, "-o", "x"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Missing index" in r.stderr
    r = subprocess.run([T.CLI, "-q"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Invalid arguments" in r.stderr


@pytest.mark.parametrize("name", ["stress_small", "exome_small"])
def test_snapshot_writer_round_trip(tmp_path, name):
    """gt16_snapshot_write (writeSnpInfoSnapshot, SnapshotHandler.cpp:83-236): a snapshot read by the library and written
    again means the same (dense stays dense, a sparsely populated one is written as mode-2 records)"""
    idx, out, region = T.regenerate(name, tmp_path)
    inp = T.LoadedInput(idx, out + ".snapshot", region, pinned=0)
    err = ctypes.create_string_buffer(256)
    again = str(tmp_path / "again.snapshot")
    try:
        assert T.capi().gt16_snapshot_write(again.encode(), ctypes.byref(inp.view), err, 256) == 0, err.value
    finally:
        inp.close()
    a, b = T.decode_snapshot(out + ".snapshot"), T.decode_snapshot(again)
    assert a["mode"] == b["mode"] and a["L"] == b["L"] and a["nreg"] == b["nreg"]
    for k in ("invalid", "overflow", "W", "F", "R"):
        assert np.array_equal(a[k], b[k]), k
    assert a["ins"] == b["ins"] and a["dele"] == b["dele"] and len(a["ins"]) + len(a["dele"]) > 0
    assert T.snapshot_content_digest(out + ".snapshot") == T.snapshot_content_digest(again) or name == "exome_small"


def test_sam_reader_filters(tmp_path):
    """gt16_sam_read applies bamToAlignment / updateSnapshotForFile's filters (bam2snapshot.cpp:79-186)"""
    idx, out, region = T.regenerate("wgs_small", tmp_path)
    L = T.capi()
    sam = tmp_path / "r.sam"
    seq, q = "ACGTACGTAC", "IIIIIIIII#"
    rows = [
        ("ok_fwd", 0, "chrA", 11, 60, "10M", seq, q),
        ("ok_rev", 16, "chrA", 21, 60, "2S3M1I2M1D2M", seq, q),
        ("dup", 0x400, "chrA", 31, 60, "10M", seq, q),            # duplicate: skipped
        ("lowq", 0, "chrA", 41, 3, "10M", seq, q),                # MAPQ below the threshold: skipped
        ("unmapped", 4, "*", 0, 0, "*", seq, q),                  # skipped
        ("lenmiss", 0, "chrA", 51, 60, "9M", seq, q),             # l_qseq != M + I + S: skipped
        ("chrB", 0, "chrB", 1, 60, "10M", "NNNNNNNNNN", "*"),     # second contig, N bases, no qualities
        ("offend", 0, "chrA", 999999999, 60, "10M", seq, q),      # beyond the contig: skipped
    ]
    with open(sam, "w") as f:
        f.write("@HD\tVN:1.0\n@SQ\tSN:chrA\tLN:180000\n@SQ\tSN:chrB\tLN:70000\n")
        for n, fl, rn, pos, mq, cg, s, ql in rows:
            f.write("%s\t%d\t%s\t%d\t%d\t%s\t*\t0\t0\t%s\t%s\n" % (n, fl, rn, pos, mq, cg, s, ql))
    err = ctypes.create_string_buffer(256)
    ix = ctypes.c_void_p(); h = ctypes.c_void_p()
    assert L.gt16_index_read(idx.encode(), ctypes.byref(ix), err, 256) == 0
    assert L.gt16_sam_read(str(sam).encode(), ix, 10, ctypes.byref(h), err, 256) == 0, err.value
    r = T.capi_mod.Reads()
    L.gt16_sam_reads(h, ctypes.byref(r))
    assert L.gt16_sam_total_alignments(h) == 8 and r.n_reads == 3
    start = np.ctypeslib.as_array(ctypes.cast(r.start, ctypes.POINTER(ctypes.c_uint32)), (3,))
    rev = np.ctypeslib.as_array(ctypes.cast(r.reverse, ctypes.POINTER(ctypes.c_uint8)), (3,))
    amb = ctypes.c_uint32()
    assert L.gt16_index_locate(ix, b"chrB", 1, ctypes.byref(amb)) == 0
    assert list(start) == [10, 20, amb.value] and list(rev) == [0, 1, 0]
    bases = np.ctypeslib.as_array(ctypes.cast(r.bases, ctypes.POINTER(ctypes.c_uint8)), (30,))
    quals = np.ctypeslib.as_array(ctypes.cast(r.quals, ctypes.POINTER(ctypes.c_uint8)), (30,))
    assert list(bases[:10]) == [0, 1, 2, 3, 0, 1, 2, 3, 0, 1] and list(bases[20:]) == [2] * 10      # N -> G (bamcode2snpcode[15])
    assert list(quals[:10]) == [40] * 9 + [2] and list(quals[20:]) == [255] * 10
    cig = np.ctypeslib.as_array(ctypes.cast(r.cigar, ctypes.POINTER(ctypes.c_uint32)), (int(r.n_cigar_ops),))
    assert list(cig) == [10 << 4, (2 << 4) | 4, 3 << 4, (1 << 4) | 1, 2 << 4, (1 << 4) | 2, 2 << 4, 10 << 4]
    # the same alignments as BAM (BGZF + binary records): identical reads
    bam = tmp_path / "r.bam"
    T.sam_to_bam(str(sam), str(bam))
    hb = ctypes.c_void_p()
    assert L.gt16_bam_read(str(bam).encode(), ix, 10, ctypes.byref(hb), err, 256) == 0, err.value
    rb = T.capi_mod.Reads()
    L.gt16_sam_reads(hb, ctypes.byref(rb))
    assert L.gt16_sam_total_alignments(hb) == 8 and rb.n_reads == 3 and rb.n_bases == r.n_bases and rb.n_cigar_ops == r.n_cigar_ops
    for field, ct, n in (("start", ctypes.c_uint32, 3), ("reverse", ctypes.c_uint8, 3), ("length", ctypes.c_uint32, 3),
                         ("seq_off", ctypes.c_uint64, 3), ("cigar_off", ctypes.c_uint64, 3), ("n_cigar", ctypes.c_uint32, 3),
                         ("bases", ctypes.c_uint8, 30), ("quals", ctypes.c_uint8, 30), ("cigar", ctypes.c_uint32, int(r.n_cigar_ops))):
        a = np.ctypeslib.as_array(ctypes.cast(getattr(r, field), ctypes.POINTER(ct)), (n,))
        b = np.ctypeslib.as_array(ctypes.cast(getattr(rb, field), ctypes.POINTER(ct)), (n,))
        assert np.array_equal(a, b), field
    L.gt16_sam_free(hb)
    assert L.gt16_bam_read(str(sam).encode(), ix, 10, ctypes.byref(hb), err, 256) == -5        # a SAM file is not a BAM file
    L.gt16_sam_free(h)
    L.gt16_index_free(ix)
    p = T.capi_mod.PileupParams()
    L.gt16_pileup_default_params(ctypes.byref(p))
    assert list(p.weight_map)[9:22] == [0, 0, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 4] and (p.trim_size, p.indel_weight_threshold, p.n_overflow_regions) == (5, 1, 16)


def test_sam_and_bam_readers_survive_damaged_input(tmp_path):
    """Mutated SAM lines and cut / corrupted BAM files: the readers return (reads or an error code), they never crash and
    every read they hand on is self-consistent (lengths, offsets, positions inside the genome)."""
    import random
    idx, out, region = T.regenerate("wgs_small", tmp_path)
    L = T.capi()
    err = ctypes.create_string_buffer(256)
    ix = ctypes.c_void_p()
    assert L.gt16_index_read(idx.encode(), ctypes.byref(ix), err, 256) == 0
    glen = L.gt16_index_length(ix)
    rng = random.Random(5)
    good = ["r%d\t%d\tchrA\t%d\t60\t%s\t*\t0\t0\t%s\t%s" % (i, rng.choice([0, 16]), rng.randrange(1, 170000), cg, "ACGTN"[i % 5] * 12, "I" * 12)
            for i, cg in enumerate(["12M", "3S9M", "4M2I6M", "5M3D7M", "6M100N6M", "2H12M", "1M1I1M1I1M1I1M1I1M1I1M1I"] * 6)]

    def check(h):
        r = T.capi_mod.Reads()
        L.gt16_sam_reads(h, ctypes.byref(r))
        n = int(r.n_reads)
        if n:
            start = np.ctypeslib.as_array(ctypes.cast(r.start, ctypes.POINTER(ctypes.c_uint32)), (n,))
            length = np.ctypeslib.as_array(ctypes.cast(r.length, ctypes.POINTER(ctypes.c_uint32)), (n,)).astype(np.int64)
            so = np.ctypeslib.as_array(ctypes.cast(r.seq_off, ctypes.POINTER(ctypes.c_uint64)), (n,)).astype(np.int64)
            co = np.ctypeslib.as_array(ctypes.cast(r.cigar_off, ctypes.POINTER(ctypes.c_uint64)), (n,)).astype(np.int64)
            nc = np.ctypeslib.as_array(ctypes.cast(r.n_cigar, ctypes.POINTER(ctypes.c_uint32)), (n,)).astype(np.int64)
            assert (start < glen).all() and (length > 0).all() and (nc > 0).all()
            assert so[0] == 0 and (so[1:] == so[:-1] + length[:-1]).all() and so[-1] + length[-1] == r.n_bases
            assert co[0] == 0 and (co[1:] == co[:-1] + nc[:-1]).all() and co[-1] + nc[-1] == r.n_cigar_ops
            bases = np.ctypeslib.as_array(ctypes.cast(r.bases, ctypes.POINTER(ctypes.c_uint8)), (int(r.n_bases),))
            assert bases.max() <= 3
        return n

    junk = ["\t", "*", "999999999999999999999", "-1", "M", "0M", "12", "chrZ", "", "\t\t\t", "12X", "268435456M", "\xff"]
    sam = tmp_path / "m.sam"
    kept = 0
    for trial in range(60):
        lines = []
        for ln in good:
            f = ln.split("\t")
            for _ in range(rng.randrange(0, 3)):
                f[rng.randrange(len(f))] = rng.choice(junk)
            if trial == 59 and len(lines) == 7:
                f[5] = "\x00"                     # a NUL cuts the line short: fewer than 11 fields
            lines.append("\t".join(f))
        with open(sam, "w", encoding="latin-1") as fh:
            fh.write("@SQ\tSN:chrA\tLN:180000\n" + "\n".join(lines) + ("\n" if trial & 1 else ""))
        h = ctypes.c_void_p()
        rc = L.gt16_sam_read(str(sam).encode(), ix, 10, ctypes.byref(h), err, 256)
        assert rc in (0, -5), rc                 # -5: a line with fewer than 11 fields
        if rc == 0:
            kept += check(h)
            L.gt16_sam_free(h)
        assert (rc == -5) == (trial == 59)
    assert kept > 500
    # BAM: cut at every few bytes, and single bytes overwritten
    with open(sam, "w") as fh:
        fh.write("@SQ\tSN:chrA\tLN:180000\n@SQ\tSN:chrB\tLN:70000\n" + "\n".join(good) + "\n")
    bam = tmp_path / "m.bam"
    T.sam_to_bam(str(sam), str(bam))
    blob = open(bam, "rb").read()
    hurt = tmp_path / "hurt.bam"
    outcomes = set()
    for trial in range(120):
        b = bytearray(blob)
        if trial < 60:
            b = b[:rng.randrange(0, len(b))]
        else:
            for _ in range(rng.randrange(1, 4)):
                b[rng.randrange(len(b))] = rng.randrange(256)
        open(hurt, "wb").write(bytes(b))
        h = ctypes.c_void_p()
        rc = L.gt16_bam_read(str(hurt).encode(), ix, 10, ctypes.byref(h), err, 256)
        outcomes.add(rc)
        assert rc in (0, -5), rc
        if rc == 0:
            check(h)
            L.gt16_sam_free(h)
    assert -5 in outcomes
    L.gt16_index_free(ix)


@pytest.mark.parametrize("name", ["stress_small", "exome_small"])
def test_snapshot_reader_survives_damaged_containers(tmp_path, name, monkeypatch):
    """Cut and overwritten snapshot containers (dense and sparse), mapped and read: an error code or a view, no crash, and
    nothing allocated by a size only the damaged bytes vouch for."""
    import random
    idx, out, region = T.regenerate(name, tmp_path)
    L = T.capi()
    blob = open(out + ".snapshot", "rb").read()
    rng = random.Random(11)
    hurt = str(tmp_path / "hurt.snapshot")
    err = ctypes.create_string_buffer(256)
    outcomes = set()
    for trial in range(80):
        b = bytearray(blob)
        if trial % 2:
            b = b[:rng.randrange(0, len(b))]
        else:
            for _ in range(rng.randrange(1, 4)):
                at = rng.randrange(0, 4096) if rng.random() < 0.7 else rng.randrange(len(b))     # mostly the header page
                b[at] = rng.randrange(256)
        open(hurt, "wb").write(bytes(b))
        for flags, nommap in ((0, False), (0, True), (2, False)):
            if nommap:
                monkeypatch.setenv("GT16_NO_MMAP", "1")
            else:
                monkeypatch.delenv("GT16_NO_MMAP", raising=False)
            hs = ctypes.c_void_p()
            rc = L.gt16_snapshot_read_ex(hurt.encode(), flags, 0, 0xFFFFFFFF, ctypes.byref(hs), err, 256)
            outcomes.add(rc)
            assert rc in (0, -4, -5), (trial, rc, err.value)          # GT16_EIO, GT16_EDATA
            if rc == 0:
                v = T.capi_mod.View()
                L.gt16_snapshot_view_of(hs, ctypes.byref(v))
                assert v.text_length > 0
                L.gt16_snapshot_free(hs)
    assert 0 in outcomes and (-4 in outcomes or -5 in outcomes)


@pytest.mark.parametrize("bgzf_batch", ["", "100000", "1"])          # "1": one member per batch, the empty end-of-file member alone in the last
def test_sam_slices_joined_in_file_order(tmp_path, monkeypatch, bgzf_batch):
    """A SAM file of more than 4 MB is parsed in slices by several threads and the lists are joined in parallel; the BAM
    reader inflates BGZF members batch by batch and parses each batch's records on several threads (GT16_BGZF_BATCH
    shrinks the batches to two members, so that most of them end inside a record): the same reads, in the same order."""
    if bgzf_batch:
        monkeypatch.setenv("GT16_BGZF_BATCH", bgzf_batch)
    import random
    idx, out, region = T.regenerate("wgs_small", tmp_path)
    L = T.capi()
    err = ctypes.create_string_buffer(256)
    ix = ctypes.c_void_p()
    assert L.gt16_index_read(idx.encode(), ctypes.byref(ix), err, 256) == 0
    rng = random.Random(3)
    sam = tmp_path / "big.sam"
    cigars = ["100M", "10S90M", "40M3I57M", "50M2D50M", "30M1I20M4D49M", "95M5S"]
    with open(sam, "w") as f:
        f.write("@SQ\tSN:chrA\tLN:180000\n@SQ\tSN:chrB\tLN:70000\n")
        for i in range(22000):
            seq = "".join(rng.choice("ACGT") for _ in range(100))
            qual = "".join(chr(33 + rng.randrange(2, 41)) for _ in range(100))
            contig, span = ("chrA", 179000) if i % 3 else ("chrB", 69000)
            f.write("r%d\t%d\t%s\t%d\t%d\t%s\t*\t0\t0\t%s\t%s\n" % (i, rng.choice([0, 16, 0x400]), contig, rng.randrange(1, span),
                                                                    rng.choice([0, 20, 60]), cigars[i % len(cigars)], seq, qual))
    assert os.path.getsize(sam) > (4 << 20)
    bam = tmp_path / "big.bam"
    T.sam_to_bam(str(sam), str(bam))
    hs, hb = ctypes.c_void_p(), ctypes.c_void_p()
    assert L.gt16_sam_read(str(sam).encode(), ix, 10, ctypes.byref(hs), err, 256) == 0, err.value
    assert L.gt16_bam_read(str(bam).encode(), ix, 10, ctypes.byref(hb), err, 256) == 0, err.value
    rs, rb = T.capi_mod.Reads(), T.capi_mod.Reads()
    L.gt16_sam_reads(hs, ctypes.byref(rs)); L.gt16_sam_reads(hb, ctypes.byref(rb))
    n = int(rs.n_reads)
    assert n > 5000 and n == rb.n_reads and rs.n_bases == rb.n_bases and rs.n_cigar_ops == rb.n_cigar_ops
    assert L.gt16_sam_total_alignments(hs) == L.gt16_sam_total_alignments(hb) == 22000
    for field, ct, k in (("start", ctypes.c_uint32, n), ("reverse", ctypes.c_uint8, n), ("length", ctypes.c_uint32, n),
                         ("seq_off", ctypes.c_uint64, n), ("cigar_off", ctypes.c_uint64, n), ("n_cigar", ctypes.c_uint32, n),
                         ("bases", ctypes.c_uint8, int(rs.n_bases)), ("quals", ctypes.c_uint8, int(rs.n_bases)),
                         ("cigar", ctypes.c_uint32, int(rs.n_cigar_ops))):
        a = np.ctypeslib.as_array(ctypes.cast(getattr(rs, field), ctypes.POINTER(ct)), (k,))
        b = np.ctypeslib.as_array(ctypes.cast(getattr(rb, field), ctypes.POINTER(ct)), (k,))
        assert np.array_equal(a, b), field
    L.gt16_sam_free(hs); L.gt16_sam_free(hb); L.gt16_index_free(ix)


def test_coordinate_translation_with_removed_stretches(tmp_path):
    """gt16_index_locate (getAmbPos, indexFunction.cpp) finds a read's translate entry by bisection: same answers as the
    reference's scan over the whole table, on a table with removed stretches inside the chromosomes -- and on a table
    whose entries are out of order, where the scan itself is used."""
    import random
    import shutil
    idx, out, region = T.regenerate("wgs_small", tmp_path)
    L = T.capi()
    head = open(idx + ".tra").read().split("\n")
    tl, ns, removed, grid = (int(x) for x in head[0].split())
    assert (tl, ns, removed) == (250000, 2, 0)
    # (startPos, chrID, correction): chrA loses 1,000 positions before amb 50,000 and 500 before amb 120,000; chrB (from amb
    # 180,000) loses 250 before amb 200,000
    entries = [(0, 1, 0xFFFFFFFF), (50000, 1, 0xFFFFFFFF - 1000), (120000, 1, 0xFFFFFFFF - 1500),
               (180000, 2, 179999), (200000, 2, 179999 - 250)]

    def scan(table, chr_id, pos):                      # the reference's rule, entry by entry
        for k, (sp, c, corr) in enumerate(table):
            if c != chr_id:
                continue
            a = (pos + corr) & 0xFFFFFFFF
            nxt = table[k + 1][0] if k + 1 < len(table) else tl
            if sp <= a < nxt and a < tl:
                return a
        return None

    rng = random.Random(2)
    for label, table in (("ordered", entries), ("shuffled", [entries[3], entries[0], entries[1], entries[4], entries[2]])):
        pre = str(tmp_path / ("idx_" + label))
        for ext in (".pac", ".ann"):
            shutil.copy(idx + ext, pre + ext)
        with open(pre + ".tra", "w") as f:
            f.write("%d %d %d %d\n" % (tl, ns, len(table) - ns, grid))
            f.write("\n".join(head[1:1 + grid]) + "\n")
            for sp, c, corr in table:
                f.write("%d %d %d\n" % (sp, c, corr))
        err = ctypes.create_string_buffer(256)
        ix = ctypes.c_void_p()
        assert L.gt16_index_read(pre.encode(), ctypes.byref(ix), err, 256) == 0, err.value
        amb = ctypes.c_uint32()
        probes = [(c, p) for c in (1, 2) for p in (0, 1, 2, 49999, 50000, 50001, 51000, 51001, 119999, 121499, 121500, 121501, 181500, 181501,
                                                    19999, 20000, 20001, 20250, 20251, 70000, 70250, 70251, 0xFFFFFFFF)]
        probes += [(rng.choice([1, 2, 3]), rng.randrange(0, 200000)) for _ in range(20000)]
        hits = 0
        for c, p in probes:
            rc = L.gt16_index_locate_id(ix, c, p, ctypes.byref(amb))
            want = scan(table, c, p)
            assert (rc == 0) == (want is not None) and (want is None or amb.value == want), (label, c, p, rc, amb.value, want)
            hits += rc == 0
        assert hits > 5000
        L.gt16_index_free(ix)


def test_txt_emitter_blocks_equal_row_by_row_calls(tmp_path):
    """gt16_write_txt formats blocks of rows on worker threads and writes them in order; the LQ/RQ values that survive from
    row to row are carried across blocks and calls.  One call for 50,000 rows must write the very bytes that appending
    them in pieces of a few hundred rows (the single-threaded path) writes."""
    idx, out, region = T.regenerate("wgs_small", tmp_path)
    L = T.capi()
    M = T.capi_mod
    err = ctypes.create_string_buffer(256)
    ix = ctypes.c_void_p()
    assert L.gt16_index_read(idx.encode(), ctypes.byref(ix), err, 256) == 0
    n = 50000
    rng = np.random.default_rng(7)
    r = np.zeros(n, dtype=M.CALLED_SITE_DTYPE)
    r["amb"] = np.sort(rng.integers(0, 250000, n)).astype(np.uint32)
    r["ref"] = rng.integers(0, 4, n); r["genotype"] = rng.integers(0, 16, n); r["rq_assigned"] = rng.integers(0, 2, n)
    r["is_indel"] = rng.choice([0, 0, 0, 2, 3], n)
    for f in ("W", "F", "R"):
        r[f] = rng.integers(0, 70000, (n, 6))
    r["lh"] = 10.0 ** rng.uniform(-320, 0, (n, 16)); r["flh"] = 10.0 ** rng.uniform(-320, 0, (n, 16))
    r["lh"][rng.random((n, 16)) < 0.1] = 0.0
    r["pD"] = 10.0 ** rng.uniform(-300, 0, n); r["bestD"] = rng.uniform(0, 1, n); r["secondD"] = rng.uniform(0, 1, n)
    r["sb"] = rng.uniform(0, 1, (n, 3, 4)); r["bqb"] = rng.uniform(0, 1, (n, 3, 4))
    for f in ("rpb", "lgc", "rgc", "asc", "lq", "rq"):
        r[f] = rng.uniform(0, 3, n)
    r["polyrun"] = rng.integers(1, 100, n); r["total_hq"] = rng.integers(0, 100, n); r["total_lq"] = rng.integers(0, 100, n)
    for k in (0, 1):
        r["indel"]["type"][:, k] = rng.choice([ord("*"), ord("I"), ord("D")], n)
        r["indel"]["len"][:, k] = rng.integers(1, 40, n)
        r["indel"]["hq"][:, k] = rng.integers(0, 500, n); r["indel"]["lq"][:, k] = rng.integers(0, 500, n)
        r["indel"]["seq"][:, k, :] = rng.integers(0, 256, (n, 64))
    whole, pieces = str(tmp_path / "whole.txt").encode(), str(tmp_path / "pieces.txt").encode()
    c1 = (ctypes.c_float * 2)(0.25, 0.75)
    assert L.gt16_write_txt(whole, ix, r.ctypes.data, n, 0, c1) == 0
    c2 = (ctypes.c_float * 2)(0.25, 0.75)
    at = 0
    while at < n:
        m = min(int(rng.integers(1, 900)), n - at)
        assert L.gt16_write_txt(pieces, ix, r[at:].ctypes.data, m, 1 if at else 0, c2) == 0
        at += m
    a, b = open(whole, "rb").read(), open(pieces, "rb").read()
    assert a == b and a.count(b"\n") == n + 1 and (c1[0], c1[1]) == (c2[0], c2[1])
    row = a.split(b"\n")[1].split(b"\t")
    assert len(row) == 98
    L.gt16_index_free(ix)
