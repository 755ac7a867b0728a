#!/usr/bin/env python
"""Generate tests/golden/*.txt.gz by running the UNMODIFIED reference objects (oracle/_ref/ref_caller,
compiled from /root/reference by oracle/build_ref.sh) on the configs of tests/golden_configs.py.

Run in the build container (needs /root/reference for the one-off build of oracle/_ref):
    python tests/golden/make_golden.py [config ...]
Writes <config>.txt.gz (the reference's .txt, byte for byte) and MANIFEST.json (sha256 of every
generated input file, so a test can prove it regenerated the very same snapshot).
The reference's ~80 s table prefill is cached in /tmp between configs (ref_caller -c): the cache
holds the reference's own LikelihoodCache bytes and is keyed by the ini error rates.
"""
import gzip, hashlib, json, os, shutil, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from golden_configs import CONFIGS, CONFIGS_ALL, FULL_CONFIGS, synth_command  # noqa: E402
import gt16_testlib as T  # noqa: E402


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 22), b""):
            h.update(blk)
    return h.hexdigest()


def main():
    if "--operators" in sys.argv:
        return
    names = sys.argv[1:] or list(CONFIGS)
    ref = os.path.join(ROOT, "oracle", "_ref", "ref_caller")
    synth = os.path.join(ROOT, "tools", "gt16_synth")
    mpath = os.path.join(HERE, "MANIFEST.json")
    manifest = json.load(open(mpath)) if os.path.exists(mpath) else {}
    for name in names:
        wd = tempfile.mkdtemp(prefix="gt16_golden_")
        try:
            argv, idx, out, region = synth_command(synth, name, wd)
            subprocess.run(argv, check=True, stderr=subprocess.DEVNULL)
            exe = ref
            ini = CONFIGS_ALL[name].get("ini")
            cache = "/tmp/gt16_ref_lc.bin"
            if ini:      # the reference reads <argv0>.ini: run it through a symlink with its own ini beside it
                exe = os.path.join(wd, "ref_caller")
                os.symlink(ref, exe)
                text = open(ref + ".ini").read()
                for k, v in ini.items():
                    text = "\n".join(("%s=%g" % (k, v)) if l.split("=")[0].strip() == k else l for l in text.split("\n"))
                open(exe + ".ini", "w").write(text)
                cache = "/tmp/gt16_ref_lc_%s.bin" % name
            cmd = [exe, idx, out] + ([region] if region else []) + ["-c", cache]
            r = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, text=True)
            called = [l for l in r.stdout.splitlines() if l.startswith("Called Entries")]
            txt = open(out + ".txt", "rb").read()
            if name not in FULL_CONFIGS:
                with open(os.path.join(HERE, name + ".txt.gz"), "wb") as f:
                    f.write(gzip.compress(txt, 9, mtime=0))
            files = {os.path.basename(p): sha(p) for p in
                     [out + ".snapshot", idx + ".pac", idx + ".ann", idx + ".tra"] + ([region] if region else [])}
            manifest[name] = dict(inputs=files, ini=CONFIGS_ALL[name].get("ini"), rows=txt.count(b"\n") - 1, called=called[0] if called else "",
                                  txt_sha256=hashlib.sha256(txt).hexdigest(), callset_sha256=T.callset_digest(txt)[0])
            if name in FULL_CONFIGS:
                manifest[name]["block_sha8"] = T.block_digests(txt)
            print(name, manifest[name]["rows"], "rows")
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    json.dump(manifest, open(mpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()


def make_operator_vectors():
    """Operator-level golden vectors from the reference objects (oracle/_ref/libref16gt_ops.so):
    lc_samples.npz       entries of the reference's own prefilled LikelihoodCache (the ten tables at
                         random (i, j), the complete rows 60 / 120 / 121 / 1999, the chi-square table)
    operator_vectors.npz computeGenotypeLikelihoodWrapper + computeGenotypeWrapper outputs for seeded
                         W[6] vectors (typical 30x sites, hets, indel evidence, deep sites that take
                         the chi-square fallback, zero coverage) and computeStrandBias /
                         computeBaseQualityBias outputs for seeded (W, F, R)."""
    import ctypes
    import numpy as np
    ops = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref16gt_ops.so"))
    cache = "/tmp/gt16_ref_lc.bin"
    if not os.path.exists(cache):
        raise SystemExit("run a config first: it leaves the reference's prefilled tables in " + cache)
    raw = np.fromfile(cache, dtype=np.float64, offset=8)            # LikelihoodCache bytes: 10 tables + ycsq
    N = 2000
    rng = np.random.default_rng(20261019)
    ks, iis, js = [], [], []
    for k in range(10):
        i = np.concatenate([rng.integers(0, 400, 2200), rng.integers(0, N, 800)])
        j = np.minimum(rng.integers(0, N, 3000), np.where(rng.random(3000) < 0.8, i, N - 1))
        for row in (60, 120, 121, 1999):
            i = np.concatenate([i, np.full(N, row)]); j = np.concatenate([j, np.arange(N)])
        ks.append(np.full(i.size, k + 1)); iis.append(i); js.append(j)
    k = np.concatenate(ks).astype(np.int32); i = np.concatenate(iis).astype(np.int32); j = np.concatenate(js).astype(np.int32)
    val = raw[(k.astype(np.int64) - 1) * N * N + i.astype(np.int64) * N + j]
    np.savez_compressed(os.path.join(HERE, "lc_samples.npz"), k=k, i=i, j=j, value=val, ycsq=raw[10 * N * N:],
                        balance=0.01, unbalance=0.005)

    n = 6000
    W = np.zeros((n, 6), dtype=np.uint32)
    ref = rng.integers(0, 4, n).astype(np.uint8)
    for s in range(n):
        kind = s % 8
        depth = int(rng.poisson(30)) if kind < 6 else int(rng.integers(400, 900))
        w = lambda d: int(sum(rng.choice([1, 2, 3, 4, 4, 4, 4, 4], size=d))) if d else 0
        if kind == 0:   W[s, ref[s]] = w(depth)
        elif kind == 1: W[s, ref[s]] = w(depth); W[s, (ref[s] + 1 + rng.integers(0, 3)) % 4] = w(int(rng.integers(1, 3)))
        elif kind == 2: a = (ref[s] + 1 + rng.integers(0, 3)) % 4; h = depth // 2; W[s, ref[s]] = w(h); W[s, a] = w(depth - h)
        elif kind == 3: a = (ref[s] + 1 + rng.integers(0, 3)) % 4; W[s, a] = w(depth); W[s, ref[s]] = w(int(rng.integers(0, 2)))
        elif kind == 4: W[s, ref[s]] = w(depth); W[s, 4] = 4 * int(rng.integers(1, 20)); W[s, 5] = 4 * int(rng.integers(0, 6))
        elif kind == 5: W[s, :4] = [w(int(rng.integers(0, 12))) for _ in range(4)]
        elif kind == 6: W[s, ref[s]] = w(depth); W[s, (ref[s] + 2) % 4] = w(int(rng.integers(0, 300)))      # W >= 2000: chi-square path
        else:           W[s, ref[s]] = w(depth); W[s, 4] = 4 * int(rng.integers(50, 400))
    W[:8] = 0                                                         # zero coverage
    lh = np.zeros((n, 16)); flh = np.zeros((n, 16)); pD = np.zeros(n); best = np.zeros(n); second = np.zeros(n)
    gt = np.zeros(n, dtype=np.int32)
    lc = np.ascontiguousarray(raw)
    vp = ctypes.c_void_p
    ops.refops_prefill_globals_only = None
    # genotype2int / prior live in globals of FisherExactTest.o: set them the way selectPossibleSNPs does
    g2i = (ctypes.c_uint * 16)(0, 1, 2, 3, 9, 10, 11, 18, 19, 27, 4, 12, 20, 28, 36, 37)
    ops.copyGenotypeTableToGPU = getattr(ops, "_Z22copyGenotypeTableToGPUPj")
    ops.copySnpPriorToGPU = getattr(ops, "_Z17copySnpPriorToGPUPd")
    prior = np.zeros(256)
    ops.refops_prior.argtypes = [vp]
    ops.refops_prior(prior.ctypes.data)
    ops.copyGenotypeTableToGPU(g2i)
    ops.copySnpPriorToGPU.argtypes = [vp]
    ops.copySnpPriorToGPU(prior.ctypes.data)
    ops.refops_call.argtypes = [vp, vp, vp, ctypes.c_uint, ctypes.c_int, ctypes.c_double, ctypes.c_double] + [vp] * 6
    ops.refops_call(lc.ctypes.data, W.ctypes.data, ref.ctypes.data, n, 4, 0.01, 0.005, lh.ctypes.data, flh.ctypes.data,
                    pD.ctypes.data, best.ctypes.data, second.ctypes.data, gt.ctypes.data)
    m = 4000
    F = rng.integers(0, 60, (m, 4)).astype(np.uint32); R = rng.integers(0, 60, (m, 4)).astype(np.uint32)
    BW = (rng.integers(0, 5, (m, 4)) * (F + R)).astype(np.uint32)
    z = rng.random((m, 4)) < 0.4
    F[z] = 0; R[z] = 0; BW[z] = 0
    deep = rng.integers(0, m, 40)
    F[deep] *= 30; R[deep] *= 27; BW[deep] = (F[deep] + R[deep]) * 3
    sb = np.zeros((m, 4), dtype=np.float32); bqb = np.zeros((m, 4), dtype=np.float32)
    ops.refops_bias.argtypes = [vp, vp, vp, ctypes.c_uint, vp, vp]
    ops.refops_bias(BW.ctypes.data, F.ctypes.data, R.ctypes.data, m, sb.ctypes.data, bqb.ctypes.data)
    np.savez_compressed(os.path.join(HERE, "operator_vectors.npz"), W=W, ref=ref, lh=lh, flh=flh, pD=pD, bestD=best,
                        secondD=second, genotype=gt, prior=prior, bias_W=BW, bias_F=F, bias_R=R, sb=sb, bqb=bqb,
                        balance=0.01, unbalance=0.005)
    print("operator vectors:", n, "sites,", m, "bias sets; genotypes", np.bincount(gt, minlength=16).tolist())


if __name__ == "__main__" and "--operators" in sys.argv:
    make_operator_vectors()
