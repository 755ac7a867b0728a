#!/usr/bin/env python
"""Turn what tools/profile_round.sh brought back in gpurun_out/ into the tracked evidence under profiles/ (run in the build
container):  python tools/collect_profiles.py r2f r2     (gpurun_out tag, profiles/ prefix)"""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, pre = sys.argv[1], sys.argv[2]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")

# launch list + per-kernel summary
src = os.path.join(G, tag + "_launches.csv")
if os.path.exists(src):
    shutil.copy(src, os.path.join(P, pre + "_launches.csv"))
    rows = list(csv.reader(open(src)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[hdr + 2:]:
        if len(r) > vi:
            name = re.sub(r"\(.*", "", r[ki]).replace("gt16::", "").replace("void ", "")
            agg.setdefault(name, []).append(float(r[vi].replace(",", "")) / 1000.0)
    with open(os.path.join(P, pre + "_launch_summary.txt"), "w") as f:
        f.write("ncu --metrics gpu__time_duration.sum --clock-control none, bench.py --workload chr20 --steps 2 --warmup 1 (first 400 launches;\n"
                "cold, serialised: shares, not absolute times)\n%-28s %8s %12s %12s\n" % ("kernel", "launches", "avg us", "total us"))
        for k, v in agg.items():
            f.write("%-28s %8d %12.2f %12.1f\n" % (k, len(v), sum(v) / len(v), sum(v)))
        # one device-resident step: a k_stream_tma launch over the whole contig (> 80 us) and the seven launches behind it
        seq = [(re.sub(r"\(.*", "", r[ki]).replace("gt16::", "").replace("void ", ""), float(r[vi].replace(",", "")) / 1000.0)
               for r in rows[hdr + 2:] if len(r) > vi]
        found = [i for i, (k, d) in enumerate(seq) if k.startswith("k_stream_tma") and d > 80.0 and i + 8 <= len(seq) and seq[i + 7][0] == "k_bias"]
        for what, i in (("first (memo cache of the Fisher tests still empty)", found[0]), ("last (steady state)", found[-1])) if found else ():
            step = seq[i:i + 8]
            tot = sum(d for _, d in step)
            f.write("\none device-resident step over the 64,444,167-site contig, the %s, launches %d..%d of the list: %.1f us\n" % (what, i, i + 7, tot))
            for k2, d2 in step:
                f.write("  %-26s %6.1f us  %5.1f %%\n" % (k2, d2, 100.0 * d2 / tot))

# full captures: summary + raw page
for part in ("kernels", "exome", "pileup"):
    rep = os.path.join(G, "%s_%s.ncu-rep" % (tag, part))
    if not os.path.exists(rep):
        continue
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], stdout=subprocess.PIPE, text=True).stdout
    open(os.path.join(P, "%s_%s_summary.txt" % (pre, part)), "w").write(out.replace(G + "/", "gpurun_out/"))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    open(os.path.join(P, "%s_%s_raw.csv" % (pre, part)), "w").write(raw)
    if part == "kernels":
        rows = list(csv.reader(raw.splitlines()))
        H = rows[0]
        for r in rows[2:]:
            if "k_stream_tma" in r[H.index("Kernel Name")]:
                rd, wr = float(r[H.index("dram__bytes_read.sum")]), float(r[H.index("dram__bytes_write.sum")])
                unit_r, unit_w = rows[1][H.index("dram__bytes_read.sum")], rows[1][H.index("dram__bytes_write.sum")]
                mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tot = rd * mul[unit_r] + wr * mul[unit_w]
                json.dump({"kernel": "k_stream_tma", "sites_per_launch": 64444167, "dram_bytes_per_launch": tot,
                           "algorithmic_bytes_per_launch": 64444167 * 8.375, "ratio": tot / (64444167 * 8.375),
                           "source": "profiles/%s_kernels_raw.csv (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)" % pre},
                          open(os.path.join(P, "stream_kernel_traffic.json"), "w"), indent=1)
        for k in ("k_stream_tma", "k_candidates", "k_eval_mask"):
            lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), rep, k, "1.0"], stdout=subprocess.PIPE, text=True).stdout
            open(os.path.join(P, "%s_%s_lines.txt" % (pre, k)), "w").write(lines)

# SASS of the roofline kernel (both variants): proof of the TMA / mbarrier path
lib = os.path.join(ROOT, "16gt_b200", "build", "gt16_kernels.o")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True).stdout
keep, on = [], False
for line in sass.splitlines():
    if "Function :" in line:
        on = "k_stream_tma" in line
    if on:
        keep.append(re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", line))
open(os.path.join(P, pre + "_k_stream_tma_sass.txt"), "w").write("\n".join(keep) + "\n")
n = lambda pat: sum(1 for l in keep if pat in l)
print("SASS: %d lines; UBLKCP %d, SYNCS.PHASECHK.TRANS64.TRYWAIT %d, SYNCS.ARRIVE.TRANS64 %d" % (len(keep), n("UBLKCP"), n("TRYWAIT"), n("SYNCS.ARRIVE")))
