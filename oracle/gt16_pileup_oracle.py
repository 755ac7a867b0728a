"""TEST INFRASTRUCTURE ONLY -- CPU restatement of what the reference producer (bam2snapshot) leaves in a snapshot, in
decoded form: per position the weight / forward / reverse totals of the four bases, whether the position is overflow-type,
whether it is invalid, and the insertion / deletion pattern lists in order of first appearance.  Plain Python + numpy,
read by read in input order, following the reference's own structure:

  filters                 bamToAlignment / updateSnapshotForFile      bam2snapshot.cpp:79-206
  per-read walk           updateSnpCounterForReads                    SNP.cpp:438-741   (trims: getTrimSize(NULL, ...) = the ini's trim size)
  compact counters        updateSnpCounter                            SNP.cpp:745-891
  overflow replay         readFilesForOverflow, updateSnpCounterFor{Overflow,Insert,Delete}   SNP.cpp:893-1688
  quality -> weight       loadBam2snapshotParas                       bam2snapshot.cpp:336-369

Pinned against the reference itself: tests/test_producer_cpu.py compares it with the decoded snapshot that
oracle/_ref/ref_bam2snapshot (the unmodified tool) writes for the same reads.  Only tests may import this module; the
product (16gt_b200/) never does.
"""
import numpy as np

CODE = {"A": 0, "C": 1, "G": 2, "T": 3, "N": 2}            # bam_nt16_table -> bamcode2snpcode (bam2snapshot.h:52-58): N -> 15 -> 2
OPS = "MIDNSHP=X"


def weight_map(bounds=(10, 13, 17, 20)):                    # bam2snapshot.cpp:352-363
    return [0 if q <= bounds[0] else 1 if q <= bounds[1] else 2 if q <= bounds[2] else 3 if q <= bounds[3] else 4 for q in range(41)]


def pileup(sam_path, locate, L, trim=5, mapq_threshold=0, indel_weight_threshold=1, wmap=None):
    """locate(name, pos1) -> concatenated position or None.  Returns a dict shaped like tests/gt16_testlib.decode_snapshot."""
    wmap = wmap or weight_map()
    W = np.zeros((L, 4), np.int64); F = np.zeros((L, 4), np.int64); R = np.zeros((L, 4), np.int64)
    ins, dele = {}, {}                                         # pos -> {pattern: [hq, lq]} in order of first appearance

    def record(table, pos, pattern, hq):
        if pos >= L:
            return
        d = table.setdefault(pos, {})
        c = d.setdefault(pattern, [0, 0])
        c[0 if hq else 1] += 1

    for line in open(sam_path):
        if line.startswith("@"):
            continue
        t = line.rstrip("\n").split("\t")
        if len(t) < 11:
            continue
        flag, rname, pos1, mapq, cigar, seq, qual = int(t[1]), t[2], int(t[3]), int(t[4]), t[5], t[9], t[10]
        if mapq < mapq_threshold or rname == "*" or pos1 <= 0 or (flag & 0x400) or seq == "*" or cigar == "*":
            continue
        ops, num = [], ""
        for ch in cigar:
            if ch.isdigit():
                num += ch
            else:
                ops.append((int(num), ch)); num = ""
        n = len(seq)
        if sum(k for k, c in ops if c in "MIS") != n:         # getLenFromCigar
            continue
        start = locate(rname, pos1)
        if start is None or start >= L:
            continue
        base = [CODE.get(c.upper(), 0 if c in "MRSVWYHKDB=mrsvwyhkdb" else 2) for c in seq]
        if qual == "*" or len(qual) != n:
            wt = [0] * n                                       # 0xFF = negative quality, corrected to 0 (SNP.cpp:478-481)
        else:
            wt = [wmap[min(ord(c) - 33, 40)] for c in qual]
        strand_fwd = not (flag & 16)                           # updateSnpCounterForReads gets 1 - align.strand: forward -> posStrandCount
        S = F if strand_fwd else R
        gpos, rpos, trimmed = start, 0, 0
        tail_from = (n - trim) & 0xFFFFFFFF                    # unsigned arithmetic of the reference (:564)
        ins_len, ins_start, del_len, del_pos, del_active = 0, 0, 0, 0, False
        held = []                                              # indels that still need a match base behind them (lastMPos, :735-741)
        alive = True
        for length, op in ops:
            if not alive:
                break
            if op in "M=X":
                rng = length
                if rpos < trim:
                    tt = min(rng, trim - trimmed)
                    gpos += tt; rpos += tt; rng -= tt; trimmed += tt
                if rng == 0:
                    continue
                if rpos >= tail_from:
                    alive = False; break
                k = min(rng, tail_from - rpos)
                # the first base of the run closes what is pending (readFilesForOverflow case 1, :957-981)
                if ins_len:
                    avg = int(sum(wt[ins_start:ins_start + ins_len]) / np.float32(ins_len & 0xFF)) & 0xFF
                    held.append((ins, gpos, (ins_len & 0xFF, tuple(base[ins_start:ins_start + ins_len])), avg > indel_weight_threshold))
                    ins_len = 0
                if del_active:
                    held.append((dele, del_pos, (del_len & 0xFF,), wt[rpos] > indel_weight_threshold))
                    del_active, del_len = False, 0
                for table, p, pat, hq in held:
                    record(table, p, pat, hq)
                held = []
                for j in range(k):
                    p = gpos + j
                    if p >= L:
                        break
                    b = base[rpos + j]
                    W[p, b] += wt[rpos + j]
                    S[p, b] += 1
                if k < rng:
                    alive = False; break
                gpos += k; rpos += k
            elif op in "IS":
                if rpos < trim:
                    rpos += length; trimmed += length
                    continue
                if rpos + length >= n:
                    alive = False; break
                if op == "I":
                    if del_active:                             # "deletion before insertion" (:983-989)
                        held.append((dele, del_pos, (del_len & 0xFF,), wt[rpos] > indel_weight_threshold))
                        del_active, del_len = False, 0
                    if not ins_len:
                        ins_start = rpos
                    ins_len += length
                rpos += length
            elif op == "D":
                if rpos < trim:
                    gpos += length
                    continue
                if rpos + 1 >= n:
                    alive = False; break
                if ins_len:                                    # :1007-1016
                    avg = int(sum(wt[ins_start:ins_start + ins_len]) / np.float32(ins_len & 0xFF)) & 0xFF
                    held.append((ins, gpos, (ins_len & 0xFF, tuple(base[ins_start:ins_start + ins_len])), avg > indel_weight_threshold))
                    ins_len = 0
                if not del_active:
                    del_active, del_pos, del_len = True, gpos, 0
                del_len += length
                gpos += length
            # N, H, P: not handled by the reference's walk

    # ---- what the counters end up as
    invalid = ((F >= 0xFFFC) | (R >= 0xFFFC) | (W >= 0xFFFF)).any(axis=1)          # SNP.cpp:783-813, :1123-1135
    alleles = ((F | R) != 0).sum(axis=1)
    deep = ((F >= 255) | (R >= 255)).any(axis=1)                                  # :852-861: a compact strand count stops at 254
    has_indel = np.zeros(L, bool)
    ins_out, del_out = {}, {}
    for table, out, ins_kind in ((ins, ins_out, True), (dele, del_out, False)):
        for p, d in table.items():
            has_indel[p] = True
            lst, size = [], 0
            for pat, (hq, lq) in d.items():
                if hq >= 0xFFFF or lq >= 0xFFFF:
                    invalid[p] = True                                             # :1343-1372
                if ins_kind:
                    ln, sq = pat
                    packed = bytearray((ln + 3) // 4)
                    for i, b in enumerate(sq):
                        packed[i >> 2] |= b << ((i & 3) << 1)
                    lst.append((ln, bytes(packed), hq, lq)); size += 1 + (ln + 3) // 4 + 4
                else:
                    lst.append((pat[0], hq, lq)); size += 5
            if size >= 4096:
                invalid[p] = True                                                 # :1408-1419, :1445-1456
            out[p] = lst
    overflow = (alleles >= 3) | deep | has_indel
    return dict(L=L, invalid=invalid, overflow=overflow, W=W.astype(np.uint32), F=F.astype(np.uint32), R=R.astype(np.uint32),
                ins=ins_out, dele=del_out)
