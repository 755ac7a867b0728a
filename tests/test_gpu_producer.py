"""GPU parity on snapshots written by the reference's OWN producer (see tests/test_producer_cpu.py for the
pipeline): through the C ABI (host-streamed from mapped pageable counters, HBM-resident, two batch-aligned shards)
and through the drop-in binary, against the reference caller's .txt for the same snapshot."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import gt16_testlib as T
from producer_fixture import producer_inputs

pytestmark = pytest.mark.gpu
NAMES = list(T.PRODUCER_CONFIGS)


def check_txt(name, got):
    want = T.golden_txt(name)
    nonident = T.compare_txt(got, want)
    m = T.manifest()[name]
    digest, rows = T.callset_digest(got)
    assert rows == m["rows"] and digest == m["callset_sha256"]
    assert nonident <= max(2, rows // 1000), "%d float fields differ in the last printed digit" % nonident
    return nonident


@pytest.mark.parametrize("name", NAMES)
@pytest.mark.parametrize("mode", ["host_pageable", "host_pinned", "resident", "shards2"])
def test_c_abi_on_producer_snapshot(gpu_caller, tmp_path, name, mode):
    idx, out = producer_inputs(name)
    inp = T.LoadedInput(idx, out + ".snapshot", None, pinned=1 if mode in ("host_pinned", "shards2") else 0)
    txt = str(tmp_path / "got.txt")
    try:
        st = T.call_to_txt(gpu_caller, inp, txt, shards=2 if mode == "shards2" else 1, resident=(mode == "resident"))
    finally:
        inp.close()
    assert sum(s.called for s in st) == T.manifest()[name]["rows"]
    assert sum(s.kernel_launches for s in st) > 0
    check_txt(name, open(txt, "rb").read())


@pytest.mark.parametrize("name", NAMES)
def test_dropin_binary_on_producer_snapshot(tmp_path, name):
    """the CLI exactly as a user of the reference would run it, on the files the reference's tools wrote"""
    idx, out = producer_inputs(name)
    r = subprocess.run([T.CLI, "-i", idx, "-o", out, "-v"], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr
    assert "Called Entries : %d" % T.manifest()[name]["rows"] in r.stdout
    got = open(out + ".txt", "rb").read()
    os.remove(out + ".txt")
    check_txt(name, got)


def test_producer_records_against_oracle(gpu_caller):
    """raw records of the 5 Mb producer-written config: doubles within 1e-6 in log space of the C oracle's operator
    outputs for the same W vectors"""
    from test_gpu_operators import log_close
    O = ctypes.CDLL(T.ORACLE_LIB)
    vp = ctypes.c_void_p
    O.oracle_call.argtypes = [ctypes.c_double, ctypes.c_double, vp, vp, ctypes.c_uint] + [vp] * 6
    idx, out = producer_inputs("prod_5mb")
    inp = T.LoadedInput(idx, out + ".snapshot", None, pinned=0)
    L = gpu_caller.L
    calls = ctypes.c_void_p(); st = T.Stats()
    try:
        gpu_caller.check(L.gt16_call_sites_host(gpu_caller.ctx, ctypes.byref(inp.view), None, 0, 0, inp.total_sites,
                                                ctypes.byref(calls), ctypes.byref(st)))
        recs = T.capi_mod.calls_array(L, calls)
        L.gt16_calls_free(calls)
    finally:
        inp.close()
    n = len(recs)
    assert n == T.manifest()["prod_5mb"]["rows"]
    W = np.ascontiguousarray(recs["W"]); ref = np.ascontiguousarray(recs["ref"])
    lh = np.zeros((n, 16)); flh = np.zeros((n, 16)); pD = np.zeros(n); b = np.zeros(n); s = np.zeros(n)
    gt = np.zeros(n, dtype=np.int32)
    O.oracle_call(0.01, 0.005, W.ctypes.data, ref.ctypes.data, n, lh.ctypes.data, flh.ctypes.data, pD.ctypes.data,
                  b.ctypes.data, s.ctypes.data, gt.ctypes.data)
    assert np.array_equal(recs["genotype"], gt)
    log_close(recs["lh"], lh); log_close(recs["flh"], flh); log_close(recs["pD"], pD)
    log_close(recs["bestD"], b); log_close(recs["secondD"], s)
