"""GPU parity tests proper: the CUDA path, called through the C ABI (and through the drop-in
snapshotSnpcaller binary), against the UNMODIFIED reference's output for the same input
(tests/golden/*.txt.gz, made by tests/golden/make_golden.py from oracle/_ref/ref_caller).

Bar (BASELINE.json north_star): identical called positions, genotypes, alleles, indel patterns,
counts and integer features (byte-identical columns); likelihood / posterior / bias columns equal
to 1e-6 relative (they are printed with 5 significant digits, so at most the last printed digit
may differ by one unit; the number of such fields is reported and bounded)."""
import os
import subprocess

import pytest

import gt16_testlib as T

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(T.CONFIGS))
@pytest.mark.parametrize("mode", ["host_stream", "host_pageable", "resident"])
def test_fused_path_matches_reference(gpu_caller, tmp_path, name, mode):
    """host_stream: page-locked host buffers copied chunk by chunk; host_pageable: ordinary malloc'd buffers staged
    through the library's page-locked ring by its stager threads; resident: snapshot uploaded once, then called"""
    idx, out, region = T.regenerate(name, tmp_path)
    inp = T.LoadedInput(idx, out + ".snapshot", region, pinned=0 if mode == "host_pageable" else 1)
    bal, unbal = T.config_ini(name)
    caller = gpu_caller if (bal, unbal) == (0.01, 0.005) else T.Caller(0, balance=bal, unbalance=unbal)
    try:
        st = T.call_to_txt(caller, inp, out + ".txt", resident=(mode == "resident"))
    finally:
        inp.close()
        if caller is not gpu_caller:
            caller.close()
    got = open(out + ".txt", "rb").read()
    want = T.golden_txt(name)
    nonident = T.compare_txt(got, want)
    rows = T.manifest()[name]["rows"]
    assert sum(s.called for s in st) == rows
    assert sum(s.kernel_launches for s in st) > 0
    assert nonident <= max(2, rows * 98 // 1000), "too many fields differ in the last digit: %d" % nonident


@pytest.mark.parametrize("mode", ["host_stream", "host_pageable", "resident", "resident_range"])
def test_sparse_snapshot_as_records(gpu_caller, tmp_path, mode):
    """A sparse (mode 2, exome) snapshot handed over as the {position, counter} records it consists of
    (GT16_READ_KEEP_SPARSE; view.counters == NULL): copied to HBM as they are and expanded there -- from page-locked
    records, from pageable (mapped) ones, into a resident snapshot (whole genome / one shard's range)."""
    import ctypes
    name = "exome_small"
    idx, out, region = T.regenerate(name, tmp_path)
    inp = T.LoadedInput(idx, out + ".snapshot", region, pinned=0 if mode == "host_pageable" else 1, keep_sparse=True)
    assert not inp.view.counters and inp.view.n_sparse_entries > 0
    try:
        if mode == "resident_range":
            L = gpu_caller.L
            dev, calls, st = ctypes.c_void_p(), ctypes.c_void_p(), T.Stats()
            gpu_caller.check(L.gt16_snapshot_upload_range(gpu_caller.ctx, ctypes.byref(inp.view), inp.regions_ptr(), inp.n_regions,
                                                          0, inp.total_sites, ctypes.byref(dev)))
            gpu_caller.check(L.gt16_call_sites(gpu_caller.ctx, dev, inp.regions_ptr(), inp.n_regions, 0, inp.total_sites,
                                               ctypes.byref(calls), ctypes.byref(st)))
            assert L.gt16_write_txt((out + ".txt").encode(), inp.ix, L.gt16_calls_data(calls), L.gt16_calls_count(calls), 0, None) == 0
            L.gt16_calls_free(calls)
            L.gt16_snapshot_release(gpu_caller.ctx, dev)
            st = [st]
        else:
            st = T.call_to_txt(gpu_caller, inp, out + ".txt", resident=(mode == "resident"))
    finally:
        inp.close()
    assert sum(s.called for s in st) == T.manifest()[name]["rows"]
    assert T.compare_txt(open(out + ".txt", "rb").read(), T.golden_txt(name)) <= 2


@pytest.mark.parametrize("name", ["wgs_small", "two_batch"])
def test_exact_inline_mode_matches_reference(tmp_path, name):
    """DecisionTables=0: every site goes through the exact evaluator in the streaming kernel."""
    idx, out, region = T.regenerate(name, tmp_path)
    with T.Caller(0, decision_tables=False) as c:
        inp = T.LoadedInput(idx, out + ".snapshot", region)
        st = T.call_to_txt(c, inp, out + ".txt")
        inp.close()
    T.compare_txt(open(out + ".txt", "rb").read(), T.golden_txt(name))
    assert st[0].flagged == st[0].called       # inline mode: the mask IS the candidate set


def test_shards_are_invisible(gpu_caller, tmp_path):
    """N batch-aligned shards (what N GPUs would each do) concatenate to the same bytes."""
    idx, out, region = T.regenerate("two_batch", tmp_path)
    inp = T.LoadedInput(idx, out + ".snapshot", region)
    T.call_to_txt(gpu_caller, inp, out + ".1.txt", shards=1)
    T.call_to_txt(gpu_caller, inp, out + ".2.txt", shards=2)
    inp.close()
    a, b = open(out + ".1.txt", "rb").read(), open(out + ".2.txt", "rb").read()
    assert a == b
    T.compare_txt(a, T.golden_txt("two_batch"))


@pytest.mark.parametrize("name", ["wgs_small", "exome_small"])
def test_dropin_binary(tmp_path, name):
    """The drop-in CLI: same flags, same <prefix>.snapshot -> <prefix>.txt, same stdout protocol."""
    idx, out, region = T.regenerate(name, tmp_path)
    cmd = [T.CLI, "-i", idx, "-o", out, "-v"] + (["-e", region] if region else [])
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    for marker in ("Loading reference index...", "Reference sequence length:", "Reading snapshot...",
                   "Handling SNP Counter Result", "Writing variants to", "Called Entries : %d" % T.manifest()[name]["rows"],
                   "Total Running Time:"):
        assert marker in r.stdout, marker
    T.compare_txt(open(out + ".txt", "rb").read(), T.golden_txt(name))


def test_dropin_binary_multi_shard(tmp_path):
    """the drop-in's multi-GPU path (one thread + context per shard, batch-aligned cuts, ordered gather) with two
    shards on whatever GPUs are visible (GT16_SHARDS): byte-identical to the single-shard output"""
    idx, out, region = T.regenerate("two_batch", tmp_path)
    env = dict(os.environ, GT16_SHARDS="2")
    r = subprocess.run([T.CLI, "-i", idx, "-o", out, "-v"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                       timeout=600, env=env)
    assert r.returncode == 0, r.stderr
    assert r.stdout.count("[GT16] GPU") == 2
    T.compare_txt(open(out + ".txt", "rb").read(), T.golden_txt("two_batch"))


def test_dropin_binary_pieces(tmp_path):
    """a shard is called in pieces whose rows are appended while the next piece computes (GT16_PIECE_BATCHES=1: every
    batch its own call, with its own shard-local upload of the side arrays); the read-instead-of-map path of the
    snapshot reader (GT16_NO_MMAP) gives the same bytes"""
    idx, out, region = T.regenerate("two_batch", tmp_path)
    want = T.golden_txt("two_batch")
    for extra in (dict(GT16_PIECE_BATCHES="1"), dict(GT16_PIECE_BATCHES="1", GT16_NO_MMAP="1", GT16_STAGE_THREADS="7")):
        r = subprocess.run([T.CLI, "-i", idx, "-o", out, "-v"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True,
                           timeout=600, env=dict(os.environ, **extra))
        assert r.returncode == 0, r.stderr
        assert "2 pieces" in r.stdout
        T.compare_txt(open(out + ".txt", "rb").read(), want)
        os.remove(out + ".txt")


def test_dropin_binary_errors(tmp_path):
    r = subprocess.run([T.CLI], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1 and "Usage:" in r.stderr
    r = subprocess.run([T.CLI, "-i", str(tmp_path / "nope"), "-o", str(tmp_path / "x")], stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True)
    assert r.returncode == 1


@pytest.mark.parametrize("name,shards", [("chr20_full", 3), ("stress_full", 2), ("exome_mid", 1), ("exome_full", 2)])
def test_full_size_callset_digest(gpu_caller, tmp_path, name, shards):
    """Full-size configs (chr20_full is what bench.py times; stress_full is the config-5 flavour with 189,080
    rows; exome_mid the config-4 flavour with a sparse snapshot and 20,000 regions; exome_full is config 4 at full
    size -- 200,000 targets on a 3.1 Gb genome, positions beyond 2^31): the call-set columns of
    every row hash to the digest of the reference's own .txt for the same snapshot, and the number of rows
    whose floating-point columns print differently is bounded."""
    idx, out, region = T.regenerate(name, tmp_path)
    # one of them through the staging ring; the full-size exome as the sparse records it is written as (expanded in HBM)
    inp = T.LoadedInput(idx, out + ".snapshot", region, pinned=0 if name == "stress_full" else 1, keep_sparse=(name == "exome_full"))
    try:
        st = T.call_to_txt(gpu_caller, inp, out + ".txt", shards=shards)
    finally:
        inp.close()
    txt = open(out + ".txt", "rb").read()
    assert T.manifest()[name]["rows"] == sum(s.called for s in st)
    bad = T.assert_full_size_txt(name, txt)
    print("%s: whole .txt %s" % (name, "IDENTICAL to the reference's" if not bad else
                                 "differs from the reference's in %d blocks of %d rows (floating-point print only)" % (bad, T.BLOCK_ROWS)))


@pytest.mark.parametrize("name", ["wgs_full", "stress_huge"])
def test_whole_genome_configs_through_dropin_binary(tmp_path, name):
    """BASELINE.json configs[2] and configs[4] at FULL size (3.1 Gb, 24 contigs, 370 batches, positions beyond 2^31;
    25.5 / 26.5 GB snapshots) through the drop-in binary on every visible GPU: .txt identical to the reference's
    (3,480,483 / 9,235,032 rows; the reference needs 25 / 38 minutes on 8 cores for them)."""
    import shutil
    wd = os.environ.get("GT16_HUGE_DIR") or ("/dev/shm" if shutil.disk_usage("/dev/shm").free > (60 << 30) else str(tmp_path))
    wd = os.path.join(wd, "gt16_huge_" + name)
    avail_ram = 0
    for line in open("/proc/meminfo"):
        if line.startswith("MemAvailable"):
            avail_ram = int(line.split()[1]) >> 20
    if shutil.disk_usage(os.path.dirname(wd)).free < (40 << 30) or avail_ram < 70:
        pytest.skip("needs 40 GiB of scratch space and 70 GiB of RAM")
    os.makedirs(wd, exist_ok=True)
    try:
        idx, out, region = T.regenerate(name, wd, check=False)
        m = T.manifest()[name]
        for p in (idx + ".pac", idx + ".ann", idx + ".tra"):
            assert T.sha256(p) == m["inputs"][os.path.basename(p)], "regenerated input differs from the golden input: " + p
        r = subprocess.run([T.CLI, "-i", idx, "-o", out, "-v"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=1200)
        assert r.returncode == 0, r.stderr[-2000:]
        os.remove(out + ".snapshot")
        assert "Called Entries : %d" % m["rows"] in r.stdout
        bad = T.assert_full_size_txt(name, open(out + ".txt", "rb").read())
        print("%s: whole .txt %s" % (name, "IDENTICAL to the reference's" if not bad else "differs in %d blocks" % bad))
    finally:
        shutil.rmtree(wd, ignore_errors=True)


@pytest.mark.parametrize("name", ["wgs_small", "stress_small"])
def test_reference_harness_with_gt16_operator_shim(tmp_path, name):
    """Drop-in at operator level: the UNMODIFIED reference objects -- its producer loop,
    consumer thread, window code and .txt writer -- linked with oracle/ref_drivers/gt16_shim.cpp in place of
    FisherExactTest.o, so that prefill / likelihood / posterior run in libgt16.so on the GPU."""
    exe = os.path.join(T.REF_DIR, "ref_caller_gt16")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/ref_caller_gt16 not built (needs /root/reference at build time)")
    idx, out, region = T.regenerate(name, tmp_path)
    r = subprocess.run([exe, idx, out] + ([region] if region else []), stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "Called Entries : %d" % T.manifest()[name]["rows"] in r.stdout
    T.compare_txt(open(out + ".txt", "rb").read(), T.golden_txt(name))


def test_error_paths(tmp_path):
    """failures are reported through status codes + gt16_last_error, never by computing something else"""
    import ctypes
    import numpy as np
    L = T.capi()
    idx, out, region = T.regenerate("wgs_small", tmp_path)
    inp = T.LoadedInput(idx, out + ".snapshot", None)
    ctx = ctypes.c_void_p()
    assert L.gt16_ctx_create(0, ctypes.byref(ctx)) == 0
    try:
        calls = ctypes.c_void_p(); st = T.Stats()
        # no tables yet
        assert L.gt16_call_sites_host(ctx, ctypes.byref(inp.view), None, 0, 0, inp.total_sites, ctypes.byref(calls), ctypes.byref(st)) == -6
        assert b"gt16_configure" in L.gt16_last_error(ctx)
        assert L.gt16_configure(ctx, 0.01, 0.005, 1) == 0
        # site_begin must be batch aligned
        assert L.gt16_call_sites_host(ctx, ctypes.byref(inp.view), None, 0, 5, inp.total_sites, ctypes.byref(calls), ctypes.byref(st)) == -1
        # region list that does not fit the index (VariantCaller.cpp:623-627)
        bad = np.array([0, inp.view.text_length + 10], dtype=np.uint32)
        assert L.gt16_call_sites_host(ctx, ctypes.byref(inp.view), bad.ctypes.data, 1, 0, 10, ctypes.byref(calls), ctypes.byref(st)) == -1
        assert b"incompatible with the index" in L.gt16_last_error(ctx)
        # a counter that points at an overflow record which does not exist -> GT16_EDATA, not a wild read
        n = inp.view.text_length
        counters = np.ctypeslib.as_array(ctypes.cast(inp.view.counters, ctypes.POINTER(ctypes.c_uint64)), (n,))
        saved = int(counters[1000])
        counters[1000] = 0xC000 | (0xC000 << 16) | (0x7FFFFFF0 << 32)
        try:
            rc = L.gt16_call_sites_host(ctx, ctypes.byref(inp.view), None, 0, 0, inp.total_sites, ctypes.byref(calls), ctypes.byref(st))
            assert rc == -5 and b"overflow record" in L.gt16_last_error(ctx)
        finally:
            counters[1000] = saved
        # and the context is still usable afterwards
        assert L.gt16_call_sites_host(ctx, ctypes.byref(inp.view), None, 0, 0, inp.total_sites, ctypes.byref(calls), ctypes.byref(st)) == 0
        assert L.gt16_calls_count(calls) == T.manifest()["wgs_small"]["rows"]
        L.gt16_calls_free(calls)
        # empty range
        assert L.gt16_call_sites_host(ctx, ctypes.byref(inp.view), None, 0, 0, 0, ctypes.byref(calls), ctypes.byref(st)) == 0
        assert L.gt16_calls_count(calls) == 0
        L.gt16_calls_free(calls)
    finally:
        L.gt16_ctx_destroy(ctx)
        inp.close()


@pytest.mark.parametrize("name", ["wgs_small", "stress_small"])
def test_table_driven_stream_variant(tmp_path, name, monkeypatch):
    """the streaming kernel without the bitmap-derived shortcut (what a non-default ini may get): same output"""
    monkeypatch.setenv("GT16_NO_SHORTCUT", "1")
    idx, out, region = T.regenerate(name, tmp_path)
    with T.Caller(0) as c:
        inp = T.LoadedInput(idx, out + ".snapshot", region)
        T.call_to_txt(c, inp, out + ".txt", resident=True)
        inp.close()
    T.compare_txt(open(out + ".txt", "rb").read(), T.golden_txt(name))


@pytest.mark.parametrize("balance,unbalance", [(0.02, 0.01), (0.004, 0.001)])
def test_non_default_ini_against_oracle(tmp_path, balance, unbalance):
    """BalanceSubError / UnbalanceSubError other than the shipped ones: ten distinct tables, other decision bitmaps,
    another (or no) shortcut bound.  The checker is the C oracle (pinned on the reference at the default ini; its
    table and likelihood code is parametric in the two rates exactly like the reference's)."""
    import ctypes
    O = ctypes.CDLL(T.ORACLE_LIB)
    O.oracle_run.argtypes = [ctypes.c_char_p] * 3 + [ctypes.c_double, ctypes.c_double, ctypes.c_char_p,
                                                     ctypes.POINTER(ctypes.c_ulonglong)]
    idx, out, region = T.regenerate("stress_small", tmp_path)
    n = ctypes.c_ulonglong()
    assert O.oracle_run(idx.encode(), (out + ".snapshot").encode(), b"", balance, unbalance, (out + ".oracle.txt").encode(),
                        ctypes.byref(n)) == 0
    with T.Caller(0, balance=balance, unbalance=unbalance) as c:
        inp = T.LoadedInput(idx, out + ".snapshot", None)
        st = T.call_to_txt(c, inp, out + ".txt")
        inp.close()
    assert sum(s.called for s in st) == n.value
    T.compare_txt(open(out + ".txt", "rb").read(), open(out + ".oracle.txt", "rb").read())


def test_device_resident_results(gpu_caller, tmp_path):
    """gt16_call_sites_device leaves the records in HBM; fetching them gives exactly the records gt16_call_sites returns"""
    import ctypes
    import numpy as np
    L = gpu_caller.L
    idx, out, region = T.regenerate("stress_small", tmp_path)
    inp = T.LoadedInput(idx, out + ".snapshot", None)
    dev = ctypes.c_void_p()
    gpu_caller.check(L.gt16_snapshot_upload(gpu_caller.ctx, ctypes.byref(inp.view), ctypes.byref(dev)))
    try:
        a = ctypes.c_void_p(); st = T.Stats()
        gpu_caller.check(L.gt16_call_sites(gpu_caller.ctx, dev, None, 0, 0, inp.total_sites, ctypes.byref(a), ctypes.byref(st)))
        ra = T.capi_mod.calls_array(L, a)
        L.gt16_calls_free(a)
        d = ctypes.c_void_p()
        gpu_caller.check(L.gt16_call_sites_device(gpu_caller.ctx, dev, None, 0, 0, inp.total_sites, ctypes.byref(d), ctypes.byref(st)))
        assert L.gt16_dev_calls_count(d) == len(ra) == T.manifest()["stress_small"]["rows"]
        assert L.gt16_dev_calls_device_ptr(d)
        b = ctypes.c_void_p()
        gpu_caller.check(L.gt16_dev_calls_fetch(gpu_caller.ctx, d, ctypes.byref(b)))
        rb = T.capi_mod.calls_array(L, b)
        L.gt16_calls_free(b); L.gt16_dev_calls_free(d)
        assert ra.tobytes() == rb.tobytes()
        assert st.flagged > 0 and st.ms_total_device > 0
        # the call that returns before its tail has run: three in a row without waiting (stream-ordered on the context),
        # then completed out of order; the last one's records are the ones in the context's buffer
        hs, sts = [], []
        for _ in range(3):
            h, s0 = ctypes.c_void_p(), T.Stats()
            gpu_caller.check(L.gt16_call_sites_device_nowait(gpu_caller.ctx, dev, None, 0, 0, inp.total_sites, ctypes.byref(h), ctypes.byref(s0)))
            assert s0.sites == inp.total_sites and s0.kernel_launches > 0      # the count is not known yet: nothing has been waited for
            hs.append(h); sts.append(s0)
        for h in (hs[2], hs[0], hs[1]):
            s1 = T.Stats()
            gpu_caller.check(L.gt16_dev_calls_wait(gpu_caller.ctx, h, ctypes.byref(s1)))
            assert s1.flagged == st.flagged and s1.called == len(ra) and s1.ms_stream > 0 and s1.ms_total_device > 0
            assert L.gt16_dev_calls_count(h) == len(ra)
            gpu_caller.check(L.gt16_dev_calls_wait(gpu_caller.ctx, h, ctypes.byref(s1)))      # idempotent
        c = ctypes.c_void_p()
        gpu_caller.check(L.gt16_dev_calls_fetch(gpu_caller.ctx, hs[2], ctypes.byref(c)))
        assert T.capi_mod.calls_array(L, c).tobytes() == ra.tobytes()
        L.gt16_calls_free(c)
        h, s0 = ctypes.c_void_p(), T.Stats()                                          # fetch completes a pending call by itself
        gpu_caller.check(L.gt16_call_sites_device_nowait(gpu_caller.ctx, dev, None, 0, 0, inp.total_sites, ctypes.byref(h), ctypes.byref(s0)))
        gpu_caller.check(L.gt16_dev_calls_fetch(gpu_caller.ctx, h, ctypes.byref(c)))
        assert T.capi_mod.calls_array(L, c).tobytes() == ra.tobytes()
        L.gt16_calls_free(c)
        for x in hs + [h]:
            L.gt16_dev_calls_free(x)
    finally:
        L.gt16_snapshot_release(gpu_caller.ctx, dev)
        inp.close()
