#!/usr/bin/env python
"""Time the producer on the GPU box: reads -> snapshot with the reference's bam2snapshot (oracle/_ref, CPU) and with
16gt_b200/bin/bam2snapshot (GPU counter construction) on a tools/gt16_readsim workload; prints both wall times, the
library's own breakdown, and whether the decoded snapshots are equal.
    python tools/producer_check.py [config of tests/golden_configs.PRODUCER_CONFIGS]"""
import os, subprocess, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import gt16_testlib as T

name = sys.argv[1] if len(sys.argv) > 1 else "prod_5mb"
wd = tempfile.mkdtemp(prefix="gt16_prodcheck_")
cfg = T.PRODUCER_CONFIGS[name]
pre = os.path.join(wd, name)
subprocess.run([T.READSIM, "--out", pre] + cfg["readsim"], check=True, stderr=subprocess.DEVNULL)
idx = pre + ".index"
subprocess.run([T.REF_MKINDEX, pre + ".fa", idx], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
t0 = time.time()
subprocess.run([T.REF_BAM2SNAPSHOT, "-i", idx, "-s", pre + ".sam", "-o", pre + ".ref"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
t_ref = time.time() - t0
t0 = time.time()
r = subprocess.run([os.path.join(T.PKG, "bin", "bam2snapshot"), "-i", idx, "-s", pre + ".sam", "-o", pre + ".mine"], check=True, stdout=subprocess.PIPE, text=True)
t_mine = time.time() - t0
a, b = T.decode_snapshot(pre + ".ref.snapshot"), T.decode_snapshot(pre + ".mine.snapshot")
ok = ~a["invalid"]
same = (np.array_equal(a["invalid"], b["invalid"]) and np.array_equal(a["overflow"][ok], b["overflow"][ok]) and
        all(np.array_equal(a[k][ok], b[k][ok]) for k in ("W", "F", "R")) and
        {p: v for p, v in a["ins"].items() if ok[p]} == {p: v for p, v in b["ins"].items() if ok[p]} and
        {p: v for p, v in a["dele"].items() if ok[p]} == {p: v for p, v in b["dele"].items() if ok[p]})
print("%s: SAM %.0f MB, %d positions; reference bam2snapshot %.1f s wall; 16gt_b200/bin/bam2snapshot %.1f s wall; decoded snapshots %s"
      % (name, os.path.getsize(pre + ".sam") / 1e6, a["L"], t_ref, t_mine, "EQUAL" if same else "DIFFER"))
print([l for l in r.stdout.split("\n") if l.startswith("[GT16]")][0])
