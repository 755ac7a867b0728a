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
            cmd = [ref, idx, out] + ([region] if region else []) + ["-c", "/tmp/gt16_ref_lc.bin"]
            r = subprocess.run(cmd, check=True, stdout=subprocess.PIPE, text=True)
            called = [l for l in r.stdout.splitlines() if l.startswith("Called Entries")]
            txt = open(out + ".txt", "rb").read()
            if name not in FULL_CONFIGS:
                with open(os.path.join(HERE, name + ".txt.gz"), "wb") as f:
                    f.write(gzip.compress(txt, 9, mtime=0))
            files = {os.path.basename(p): sha(p) for p in
                     [out + ".snapshot", idx + ".pac", idx + ".ann", idx + ".tra"] + ([region] if region else [])}
            manifest[name] = dict(inputs=files, rows=txt.count(b"\n") - 1, called=called[0] if called else "",
                                  txt_sha256=hashlib.sha256(txt).hexdigest(), callset_sha256=T.callset_digest(txt)[0])
            print(name, manifest[name]["rows"], "rows")
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    json.dump(manifest, open(mpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
