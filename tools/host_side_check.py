#!/usr/bin/env python3
"""Host-side stages of the drop-in tools timed on this machine's cores (no GPU needed): the SAM and BAM readers of
16gt_b200/bin/bam2snapshot and the .txt emitter of 16gt_b200/bin/snapshotSnpcaller.  Writes profiles/r2_host_side.log.

    python tools/host_side_check.py [workdir]

Inputs: prod_5mb reads from tools/gt16_readsim (the config of tests/golden_configs.py), the first 300,000 alignments
of it as BAM (written by the test library's sam_to_bam), one million random call records."""
import ctypes
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gt16_testlib as T                                    # noqa: E402
from golden_configs import PRODUCER_CONFIGS                 # noqa: E402


def main():
    wd = sys.argv[1] if len(sys.argv) > 1 else "/tmp/gt16_host_side"
    os.makedirs(wd, exist_ok=True)
    pre = os.path.join(wd, "prod_5mb")
    if not os.path.exists(pre + ".sam"):
        subprocess.run([T.READSIM, "--out", pre] + PRODUCER_CONFIGS["prod_5mb"]["readsim"], check=True, stderr=subprocess.DEVNULL)
        subprocess.run([T.REF_MKINDEX, pre + ".fa", pre + ".index"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    part = os.path.join(wd, "part")
    if not os.path.exists(part + ".bam"):
        with open(pre + ".sam") as f, open(part + ".sam", "w") as g:
            for k, line in enumerate(f):
                if k >= 300003:
                    break
                g.write(line)
        T.sam_to_bam(part + ".sam", part + ".bam")
    L, M = T.capi(), T.capi_mod
    err = ctypes.create_string_buffer(256)
    ix = ctypes.c_void_p()
    assert L.gt16_index_read((pre + ".index").encode(), ctypes.byref(ix), err, 256) == 0
    lines = ["host-side stages on %d cores (%s)" % (os.cpu_count(), time.strftime("%Y-%m-%d"))]

    def best(fn, reps=3):
        ts = []
        for _ in range(reps):
            t = time.time(); fn(); ts.append(time.time() - t)
        return min(ts)

    def read(fn, path):
        def go():
            h = ctypes.c_void_p()
            assert fn(path.encode(), ix, 0, ctypes.byref(h), err, 256) == 0, err.value
            go.n = L.gt16_sam_total_alignments(h)
            L.gt16_sam_free(h)
        return go

    g = read(L.gt16_sam_read, pre + ".sam")
    t = best(g)
    lines.append("gt16_sam_read  %4.0f MB SAM, %d alignments: %.3f s (%.2f GB/s)" % (os.path.getsize(pre + ".sam") / 1e6, g.n, t, os.path.getsize(pre + ".sam") / t / 1e9))
    g = read(L.gt16_bam_read, part + ".bam")
    t = best(g)
    lines.append("gt16_bam_read  %4.0f MB BAM (%.0f MB of SAM), %d alignments: %.3f s" % (os.path.getsize(part + ".bam") / 1e6, os.path.getsize(part + ".sam") / 1e6, g.n, t))
    os.environ["GT16_BGZF_BATCH"] = "1"            # one member per batch and nothing in parallel: what a streaming reader does
    t1 = best(g, 1)
    del os.environ["GT16_BGZF_BATCH"]
    lines.append("               the same file member by member on one thread: %.3f s" % t1)
    n = 1_000_000
    rng = np.random.default_rng(1)
    r = np.zeros(n, dtype=M.CALLED_SITE_DTYPE)
    r["amb"] = np.sort(rng.integers(0, 5_000_000, n)).astype(np.uint32)
    r["ref"] = rng.integers(0, 4, n); r["genotype"] = rng.integers(0, 10, n); r["rq_assigned"] = 1
    for f in ("W", "F", "R"):
        r[f] = rng.integers(0, 200, (n, 6))
    r["lh"] = 10.0 ** rng.uniform(-300, 0, (n, 16)); r["flh"] = 10.0 ** rng.uniform(-300, 0, (n, 16))
    r["pD"] = 10.0 ** rng.uniform(-300, 0, n); r["bestD"] = rng.uniform(0, 1, n); r["secondD"] = rng.uniform(0, 1, n)
    r["sb"] = rng.uniform(0, 1, (n, 3, 4)); r["bqb"] = rng.uniform(0, 1, (n, 3, 4))
    for f in ("rpb", "lgc", "rgc", "asc", "lq", "rq"):
        r[f] = rng.uniform(0, 1, n)
    r["polyrun"] = rng.integers(1, 10, n); r["total_hq"] = rng.integers(0, 100, n)
    r["indel"]["type"] = ord("*")
    txt = os.path.join(wd, "bench.txt")
    t = best(lambda: L.gt16_write_txt(txt.encode(), ix, r.ctypes.data, n, 0, None))
    lines.append("gt16_write_txt %d rows, %.0f MB of text: %.3f s (%.2f M rows/s)" % (n, os.path.getsize(txt) / 1e6, t, n / t / 1e6))
    out = os.path.join(ROOT, "profiles", "r2_host_side.log")
    with open(out, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
