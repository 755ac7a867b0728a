#!/usr/bin/env python
"""Golden .txt of the PRODUCER-WRITTEN configs (tests/golden_configs.py: PRODUCER_CONFIGS): every step is the
reference's own code -- ref_mkindex (its FASTA packer), ref_bam2snapshot (its unmodified producer, SAM input) and
ref_caller (its caller) -- on reads made by tools/gt16_readsim.  Run in the build container:
    python tests/golden/make_producer_golden.py [config ...]
Writes <config>.txt.gz and the MANIFEST.json entry (index-file hashes, the content digest of the snapshot -- the file
itself is not byte-reproducible, see gt16_testlib.parse_snapshot_container --, row count, .txt digests)."""
import gzip, hashlib, json, os, shutil, subprocess, sys, tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import gt16_testlib as T  # noqa: E402


def main():
    names = sys.argv[1:] or list(T.PRODUCER_CONFIGS)
    mpath = os.path.join(HERE, "MANIFEST.json")
    manifest = json.load(open(mpath))
    for name in names:
        wd = tempfile.mkdtemp(prefix="gt16_prod_")
        try:
            idx, out = T.producer_pipeline(name, wd)
            r = subprocess.run([T.REF_CALLER, idx, out, "-c", "/tmp/gt16_ref_lc.bin"], check=True, stdout=subprocess.PIPE, text=True)
            called = [l for l in r.stdout.splitlines() if l.startswith("Called Entries")]
            txt = open(out + ".txt", "rb").read()
            with open(os.path.join(HERE, name + ".txt.gz"), "wb") as f:
                f.write(gzip.compress(txt, 9, mtime=0))
            files = {os.path.basename(p): T.sha256(p) for p in (idx + ".pac", idx + ".ann", idx + ".tra")}
            files["snapshot_content"] = T.snapshot_content_digest(out + ".snapshot")
            manifest[name] = dict(inputs=files, ini=None, rows=txt.count(b"\n") - 1, called=called[0] if called else "",
                                  txt_sha256=hashlib.sha256(txt).hexdigest(), callset_sha256=T.callset_digest(txt)[0],
                                  producer="oracle/_ref/ref_bam2snapshot (unmodified reference)")
            print(name, manifest[name]["rows"], "rows")
        finally:
            shutil.rmtree(wd, ignore_errors=True)
    json.dump(manifest, open(mpath, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
