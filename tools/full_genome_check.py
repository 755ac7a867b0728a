#!/usr/bin/env python
"""Full-size parity check of BASELINE.json configs[2] (3.1 Gb genome, 24 contigs, 30x, dense snapshot: 24.8 GB of counters,
370 batches, positions beyond 2^31).  Regenerates the snapshot with tools/gt16_synth (a pure function of its arguments),
runs the drop-in binary 16gt_b200/bin/snapshotSnpcaller on it with every visible GPU, and compares the .txt with the digest
of the UNMODIFIED reference's .txt for the same snapshot (tests/golden/MANIFEST.json: wgs_full / stress_huge, made in the build
container by tests/golden/make_golden.py wgs_full -- 25 min of reference CPU time on 8 cores).

    python tools/full_genome_check.py [workdir] [--config wgs_full|stress_huge] [ENV=v1,v2 ...]   (needs ~40 GB of disk, 70 GB of RAM)
stress_huge is configs[4] at full size (3.1 Gb at 60x, indels 1/500 with pooled multi-pattern lists, depth hotspots, invalid
bits, gaps; reference: 9,235,032 rows in 38 min).  ENV=v1,v2 runs the binary once per value of an environment knob first
(timing only).
"""
import hashlib, json, os, shutil, subprocess, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gt16_testlib as T  # noqa: E402
from golden_configs import synth_command  # noqa: E402

def main():
    args = sys.argv[1:]
    NAME = "wgs_full"                    # or --config stress_huge (configs[4] at full size: 60x, 9,235,032 rows)
    if "--config" in args:
        k = args.index("--config")
        NAME = args[k + 1]
        del args[k:k + 2]
    wd = args[0] if args and "=" not in args[0] else "/tmp/gt16_full"
    os.makedirs(wd, exist_ok=True)
    free_disk = shutil.disk_usage(wd).free / 2**30
    avail_ram = 0
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable"):
            avail_ram = int(line.split()[1]) / 2**20
    print("workdir %s: %.0f GiB disk free, %.0f GiB RAM available, %d cores" % (wd, free_disk, avail_ram, os.cpu_count()), flush=True)
    if free_disk < 40 or avail_ram < 70:
        print("FULL GENOME: skipped (needs 40 GiB of disk and 70 GiB of RAM)")
        return 0
    m = T.manifest()[NAME]
    argv, idx, out, region = synth_command(T.SYNTH, NAME, wd)
    t0 = time.time()
    subprocess.run(argv + ["--threads", str(min(32, os.cpu_count() or 8))], check=True, stderr=subprocess.DEVNULL)
    print("synth: %.1f s, snapshot %.1f GB" % (time.time() - t0, os.path.getsize(out + ".snapshot") / 1e9), flush=True)
    for p in (idx + ".pac", idx + ".ann", idx + ".tra"):
        assert T.sha256(p) == m["inputs"][os.path.basename(p)], "regenerated input differs from the golden input: " + p
    for setting in [a for a in args if "=" in a]:         # optional sweeps: NAME=v1,v2 (timed only)
        key, vals = setting.split("=", 1)
        for v in vals.split(","):
            t0 = time.time()
            subprocess.run([T.CLI, "-i", idx, "-o", out], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, env=dict(os.environ, **{key: v}))
            print("sweep %s=%s: %.2f s wall" % (key, v, time.time() - t0), flush=True)
    t0 = time.time()
    r = subprocess.run([T.CLI, "-i", idx, "-o", out, "-v"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    wall = time.time() - t0
    keep = [l for l in r.stdout.splitlines() if any(k in l for k in ("GT16", "Called", "Total Running", "Read snapshot", "Elapsed"))]
    print("\n".join(keep[-20:]))
    if r.returncode != 0:
        print("FULL GENOME: drop-in binary failed:", r.stderr[-2000:])
        return 1
    print("drop-in snapshotSnpcaller: %.1f s wall (snapshot file -> .txt)" % wall, flush=True)
    os.remove(out + ".snapshot")
    h = hashlib.sha256()
    rows = -1
    with open(out + ".txt", "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
            rows += blk.count(b"\n")
    print("rows %d (reference %d)" % (rows, m["rows"]))
    if h.hexdigest() == m["txt_sha256"]:
        print("FULL GENOME: .txt is byte-identical to the reference's (sha256 %s)" % h.hexdigest())
        return 0
    digest, nrows = T.callset_digest(open(out + ".txt", "rb").read())
    ok = digest == m["callset_sha256"] and nrows == m["rows"]
    print("FULL GENOME: call-set columns %s the reference's (floating-point columns differ in print somewhere)"
          % ("IDENTICAL to" if ok else "DIFFER from"))
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
