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
    argv += ["--threads", str(max(1, min(32, os.cpu_count() or 8)))]        # the output does not depend on the thread count
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


BLOCK_ROWS = 1000


def block_digests(txt: bytes):
    """8-hex-digit digests of consecutive blocks of BLOCK_ROWS rows (header excluded): lets a test on a box without the
    reference's multi-megabyte .txt say HOW MUCH of a full-size output is not byte-identical, not only whether"""
    lines = txt.split(b"\n")[1:]
    if lines and not lines[-1]:
        lines.pop()
    return [hashlib.sha256(b"\n".join(lines[k:k + BLOCK_ROWS])).hexdigest()[:8] for k in range(0, len(lines), BLOCK_ROWS)]


def assert_full_size_txt(name, txt: bytes):
    """whole-file identity with the reference's .txt, or -- should a floating-point column print differently somewhere --
    identical call-set columns and at most 1 % of the 1000-row blocks touched"""
    m = manifest()[name]
    digest, rows = callset_digest(txt)
    assert rows == m["rows"], "row count %d != %d" % (rows, m["rows"])
    assert digest == m["callset_sha256"], "call-set columns differ from the reference's"
    if hashlib.sha256(txt).hexdigest() == m["txt_sha256"]:
        return 0
    want = m.get("block_sha8")
    assert want, "%s: .txt differs from the reference's in a floating-point column and the manifest has no block digests" % name
    got = block_digests(txt)
    assert len(got) == len(want)
    bad = sum(1 for a, b in zip(got, want) if a != b)
    assert bad <= max(1, len(want) // 100), "%d of %d blocks differ from the reference's .txt" % (bad, len(want))
    return bad


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




# ---- snapshots written by the reference's OWN producer ------------------------------------------
# tools/gt16_readsim (FASTA + SAM, a pure function of its arguments) -> oracle/_ref/ref_mkindex (the reference's
# packer) -> oracle/_ref/ref_bam2snapshot (the unmodified bam2snapshot: SNP.cpp promotion rules, the
# writeSnpInfoSnapshot container with its page-aligned subfiles, the full 1 GiB pool and the soft-clip subfile).
READSIM = os.path.join(ROOT, "tools", "gt16_readsim")
REF_MKINDEX = os.path.join(REF_DIR, "ref_mkindex")
REF_BAM2SNAPSHOT = os.path.join(REF_DIR, "ref_bam2snapshot")

from golden_configs import PRODUCER_CONFIGS  # noqa: E402


def parse_snapshot_container(path):
    """Independent (numpy) reading of the container writeSnpInfoSnapshot produces (SnapshotHandler.cpp:83-236): returns
    a dict of its meaningful content.  The file itself is not byte-reproducible (the pool is written at its full
    capacity and only its first curPtr bytes are initialised), so identity of two snapshots is the identity of this
    content."""
    with open(path, "rb") as f:
        hdr = f.read(4 + 5 * 16 + 1)
        page = int(np.frombuffer(hdr, "<i4", 1, 0)[0])
        tab = np.frombuffer(hdr, "<u8", 10, 4).reshape(5, 2)
        f.seek(int(tab[0, 0]))
        cur = int(np.frombuffer(f.read(4), "<u4")[0]); cap = int(np.frombuffer(f.read(8), "<u8")[0])
        pool = f.read(cur)
        f.seek(int(tab[1, 0]))
        mode = f.read(1)[0]
        L = int(np.frombuffer(f.read(4), "<u4")[0])
        if mode & 1:
            counters = np.fromfile(f, "<u8", L)
        else:
            rec = np.fromfile(f, np.dtype([("pos", "<u4"), ("c", "<u8")]), int(tab[1, 1]) // 12)
            counters = np.zeros(L, "<u8"); counters[rec["pos"]] = rec["c"]
        f.seek(int(tab[2, 0]))
        nreg = int(np.frombuffer(f.read(4), "<u4")[0])
        ovf = []
        for _ in range(nreg):
            sz, lim = (int(x) for x in np.frombuffer(f.read(8), "<u4"))
            ovf.append(f.read(sz * 32) if lim else b"")
        bitmap = None
        if int(tab[3, 1]):
            f.seek(int(tab[3, 0]))
            bitmap = np.fromfile(f, "<u4", (L + 31) >> 5)
    return dict(page=page, subfiles=tab, pool=pool, pool_capacity=cap, mode=mode, L=L, counters=counters, ovf=ovf,
                bitmap=bitmap, amplicon=hdr[-1])


def snapshot_content_digest(path):
    """sha256 of what a snapshot MEANS.  bam2snapshot resolves its overflow buffers with one thread per region, and
    those threads allocate pattern lists from the shared pool in whatever order they get there: pool offsets (and so
    the pointer bytes of the overflow records) differ from run to run, the lists themselves do not.  The digest
    therefore replaces every pool pointer by the bytes it points to."""
    c = parse_snapshot_container(path)
    L, nreg = c["L"], len(c["ovf"])
    cnt = c["counters"]
    lo = (cnt & 0xFFFFFFFF).astype(np.uint32)
    recs = [np.frombuffer(o, np.uint8).reshape(-1, 32).copy() for o in c["ovf"]]
    lists = []
    pos = np.nonzero(((lo & (lo >> 16)) & 0xC000) == 0xC000)[0]
    seg = max(1, L // nreg)
    for p in pos[((lo[pos] & 0x3000) == 0x3000) | (((lo[pos] >> 16) & 0x3000) == 0x3000)]:
        r = min(int(p) // seg, nreg - 1) if L // nreg else nreg - 1      # findRegionByPosition (coreStruct.h:141-154)
        rec = recs[r][int(cnt[p] >> 32)]
        for w, off in ((int(lo[p]) & 0xFFFF, 24), (int(lo[p]) >> 16, 28)):
            if (w & 0x3000) == 0x3000:
                ptr = int(rec[off:off + 4].view("<u4")[0])
                lists.append(np.array([p, off, w & 0xFFF], "<u8").tobytes() + c["pool"][ptr:ptr + (w & 0xFFF)])
                rec[off:off + 4] = 0
    h = hashlib.sha256()
    h.update(np.array([L, c["mode"], nreg, len(c["pool"])], "<u8").tobytes())
    h.update(cnt.tobytes())
    for o in recs:
        h.update(np.array([len(o)], "<u8").tobytes()); h.update(o.tobytes())
    if c["bitmap"] is not None:
        h.update(c["bitmap"].tobytes())
    for x in lists:
        h.update(x)
    return h.hexdigest()


def decode_snapshot(path):
    """What a snapshot MEANS, position by position (the comparison bar of the producer): the invalid bitmap,
    which positions are overflow-type, the weight / forward / reverse totals of the four bases, and the insertion and
    deletion pattern lists (pattern, hq, lq) in list order.  Slot order, array indices and pool offsets are gone."""
    c = parse_snapshot_container(path)
    L, nreg = c["L"], len(c["ovf"])
    cnt = c["counters"]
    lo = (cnt & 0xFFFFFFFF).astype(np.uint32)
    hi = (cnt >> 32).astype(np.uint32)
    w1, w2 = lo & 0xFFFF, lo >> 16
    ovf = ((w1 & w2) & 0xC000) == 0xC000
    W = np.zeros((L, 4), np.uint32); F = np.zeros((L, 4), np.uint32); R = np.zeros((L, 4), np.uint32)
    idx = np.arange(L)
    for w, f, r in ((w1, hi & 0xFF, (hi >> 8) & 0xFF), (w2, (hi >> 16) & 0xFF, hi >> 24)):
        live = (~ovf) & ((f | r) != 0)
        b = (w >> 14)[live]
        W[idx[live], b] = (w & 0x3FFF)[live]; F[idx[live], b] = f[live]; R[idx[live], b] = r[live]
    recs = [np.frombuffer(o, np.uint8).reshape(-1, 32) for o in c["ovf"]]
    seg = L // nreg
    ins, dele = {}, {}
    pool = c["pool"]
    for p in np.nonzero(ovf)[0]:
        p = int(p)
        r = min(p // seg, nreg - 1) if seg else nreg - 1
        rec = recs[r][int(hi[p])]
        u16 = rec[:24].view("<u2")
        f, rv = u16[4:8].astype(np.uint32), u16[8:12].astype(np.uint32)
        seen = (f | rv) != 0
        W[p] = np.where(seen, u16[0:4], 0); F[p] = f; R[p] = rv
        a, b = int(w1[p]), int(w2[p])
        if a & 0x2000:
            lst = []
            if a & 0x1000:
                ptr, size, i = int(rec[24:28].view("<u4")[0]), a & 0xFFF, 0
                info = pool[ptr:ptr + size]
                while i < size:
                    n = info[i]; nb = (n + 3) // 4
                    lst.append((n, bytes(info[i + 1:i + 1 + nb]), info[i + 1 + nb] | info[i + 2 + nb] << 8, info[i + 3 + nb] | info[i + 4 + nb] << 8))
                    i += 1 + nb + 4
            else:
                n = (a >> 8) & 0xF
                lst.append((n, bytes(rec[24:24 + (n + 3) // 4]), a & 0xFF, int(rec[27])))
            ins[p] = lst
        if b & 0x2000:
            lst = []
            if b & 0x1000:
                ptr, size, i = int(rec[28:32].view("<u4")[0]), b & 0xFFF, 0
                info = pool[ptr:ptr + size]
                while i < size:
                    lst.append((info[i], info[i + 1] | info[i + 2] << 8, info[i + 3] | info[i + 4] << 8))
                    i += 5
            else:
                lst.append((int(rec[28]), int(rec[29]), int(rec[30])))
            dele[p] = lst
    bitmap = c["bitmap"] if c["bitmap"] is not None else np.zeros((L + 31) >> 5, np.uint32)
    bits = np.unpackbits(bitmap.astype(">u4").view(np.uint8))[:L].astype(bool)      # bit 31 - (p & 31) of word p >> 5
    return dict(L=L, invalid=bits, overflow=ovf, W=W, F=F, R=R, ins=ins, dele=dele, mode=c["mode"], nreg=nreg)


def producer_pipeline(name, workdir, keep_sam=False):
    """reads -> index -> snapshot with the reference's own tools; returns (index_prefix, out_prefix)"""
    cfg = PRODUCER_CONFIGS[name]
    wd = str(workdir)
    pre = os.path.join(wd, name)
    subprocess.run([READSIM, "--out", pre] + cfg["readsim"], check=True, stderr=subprocess.DEVNULL)
    idx = pre + ".index"
    subprocess.run([REF_MKINDEX, pre + ".fa", idx], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    os.makedirs(os.path.join(wd, "out"), exist_ok=True)
    out = os.path.join(wd, "out", name)
    subprocess.run([REF_BAM2SNAPSHOT, "-i", idx, "-s", pre + ".sam", "-o", out], check=True, stdout=subprocess.DEVNULL,
                   stderr=subprocess.DEVNULL)
    if not keep_sam:
        os.remove(pre + ".sam")
    return idx, out


def regenerate_producer(name, workdir, check=True):
    """producer_pipeline + proof that the inputs are the ones the committed golden was made from"""
    idx, out = producer_pipeline(name, workdir)
    if check:
        want = manifest()[name]["inputs"]
        for p in (idx + ".pac", idx + ".ann", idx + ".tra"):
            assert sha256(p) == want[os.path.basename(p)], "regenerated input differs from golden input: " + p
        assert snapshot_content_digest(out + ".snapshot") == want["snapshot_content"], "producer-written snapshot differs"
    return idx, out


# ---- SAM -> BAM (test fixture: the producer's -b input) ---------------------------------------------
def sam_to_bam(sam_path, bam_path):
    """Write the alignments of a SAM file as BAM (SAM specification 4.2: BGZF blocks of a binary header and records; no
    optional fields).  The reference's bam2snapshot -b (samtools' reader) and gt16_bam_read both take the result."""
    import struct
    import zlib
    nib = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    ops = {c: i for i, c in enumerate("MIDNSHP=X")}

    def reg2bin(beg, end):
        end -= 1
        for shift, base in ((14, 4681), (17, 585), (20, 73), (23, 9), (26, 1)):
            if beg >> shift == end >> shift:
                return base + (beg >> shift)
        return 0

    header, refs, body = [], [], bytearray()
    ref_id = {}
    with open(sam_path) as f:
        for line in f:
            if line.startswith("@"):
                header.append(line)
                if line.startswith("@SQ"):
                    d = dict(x.split(":", 1) for x in line.rstrip("\n").split("\t")[1:])
                    ref_id[d["SN"]] = len(refs)
                    refs.append((d["SN"], int(d["LN"])))
                continue
            t = line.rstrip("\n").split("\t")
            if len(t) < 11:
                continue
            qname, flag, rname, pos, mapq, cigar, rnext, pnext, tlen, seq, qual = t[:11]
            flag, pos, mapq, pnext, tlen = int(flag), int(pos), int(mapq), int(pnext), int(tlen)
            rid = ref_id.get(rname, -1)
            nrid = rid if rnext == "=" else ref_id.get(rnext, -1)
            cig = []
            if cigar != "*":
                num = ""
                for ch in cigar:
                    if ch.isdigit():
                        num += ch
                    else:
                        cig.append((int(num) << 4) | ops[ch]); num = ""
            lseq = 0 if seq == "*" else len(seq)
            span = sum(c >> 4 for c in cig if (c & 15) in (0, 2, 3, 7, 8)) or 1
            name = qname.encode() + b"\0"
            rec = struct.pack("<iiBBHHHiiii", rid, pos - 1, len(name), mapq, reg2bin(max(pos - 1, 0), max(pos - 1, 0) + span) & 0xFFFF, len(cig), flag,
                              lseq, nrid, pnext - 1, tlen) + name + struct.pack("<%dI" % len(cig), *cig)
            packed = bytearray((lseq + 1) // 2)
            for k in range(lseq):
                packed[k >> 1] |= nib.get(seq[k].upper(), 15) << (4 if k % 2 == 0 else 0)
            rec += bytes(packed) + (b"\xff" * lseq if qual == "*" else bytes(ord(c) - 33 for c in qual))
            body += struct.pack("<i", len(rec)) + rec
    text = "".join(header).encode()
    head = b"BAM\1" + struct.pack("<i", len(text)) + text + struct.pack("<i", len(refs))
    for name, ln in refs:
        head += struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", ln)
    data = head + bytes(body)
    with open(bam_path, "wb") as out:
        for a in range(0, len(data), 0xFF00):
            chunk = data[a:a + 0xFF00]
            c = zlib.compressobj(6, zlib.DEFLATED, -15)
            comp = c.compress(chunk) + c.flush()
            out.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", len(comp) + 25) + comp +
                      struct.pack("<II", zlib.crc32(chunk) & 0xFFFFFFFF, len(chunk)))
        out.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))      # the BGZF end-of-file block
