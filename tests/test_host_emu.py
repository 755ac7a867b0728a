"""CPU tests of the PRODUCT's per-site source (16gt_b200/csrc/gt16_core.cuh -- the functions the CUDA
kernels are made of) compiled for the host by tests/host_emu: table builder, likelihood, posterior, float
Fisher tests with the re-implemented expf, counter decoding, window features, .txt writer.  This is a check of
the product's arithmetic on a machine without a GPU; libgt16.so itself has no CPU path."""
import ctypes
import os

import numpy as np
import pytest

import gt16_testlib as T

vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def E():
    lib = ctypes.CDLL(T.EMU_LIB)
    lib.emu_run.argtypes = [ctypes.c_char_p] * 3 + [vp, ctypes.c_double, ctypes.c_double, ctypes.c_char_p]
    lib.emu_table_entry.restype = ctypes.c_double
    lib.emu_table_entry.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_float]
    lib.emu_err_fractions.argtypes = [ctypes.c_float, ctypes.c_float, vp]
    lib.emu_call.argtypes = [vp, ctypes.c_double, ctypes.c_double, vp, vp, ctypes.c_uint] + [vp] * 6
    lib.emu_bias.argtypes = [vp, vp, vp, ctypes.c_uint, vp, vp]
    lib.emu_expf.restype = ctypes.c_float
    lib.emu_expf.argtypes = [ctypes.c_float]
    lib.emu_chisqr1.restype = ctypes.c_double
    lib.emu_chisqr1.argtypes = [ctypes.c_double]
    return lib


def test_table_builder_matches_reference_prefill(E):
    g = np.load(os.path.join(T.GOLDEN, "lc_samples.npz"))
    err = np.zeros(10, dtype=np.float32)
    E.emu_err_fractions(0.01, 0.005, err.ctypes.data)
    sel = np.random.default_rng(1).choice(len(g["k"]), 20000, replace=False)
    got = np.array([E.emu_table_entry(int(g["i"][s]), int(g["j"][s]), float(err[g["k"][s] - 1])) for s in sel])
    assert np.array_equal(got.view(np.uint64), g["value"][sel].view(np.uint64))
    ycsq = np.array([E.emu_chisqr1(t / 10.0) for t in range(15000)])
    assert np.array_equal(ycsq.view(np.uint64), g["ycsq"].view(np.uint64))


def test_operators_match_reference(E):
    g = np.load(os.path.join(T.GOLDEN, "operator_vectors.npz"))
    n = len(g["ref"])
    W = np.ascontiguousarray(g["W"]); ref = np.ascontiguousarray(g["ref"])
    lh = np.zeros((n, 16)); flh = np.zeros((n, 16)); pD = np.zeros(n); b = np.zeros(n); s = np.zeros(n)
    gt = np.zeros(n, dtype=np.int32)
    E.emu_call(None, 0.01, 0.005, W.ctypes.data, ref.ctypes.data, n, lh.ctypes.data, flh.ctypes.data, pD.ctypes.data,
               b.ctypes.data, s.ctypes.data, gt.ctypes.data)
    assert np.array_equal(gt, g["genotype"])
    # The product takes sqrt() where the reference calls pow(x, 0.5f) (FisherExactTest.cpp:192): glibc's pow is
    # not always correctly rounded, so het likelihoods may differ in the last ulp (42 of 96,000 values here);
    # everything else is bit-identical.  The bar is 1e-6 relative in log space; hold the host build to 1e-14.
    for got, name in ((lh, "lh"), (flh, "flh"), (pD, "pD"), (b, "bestD"), (s, "secondD")):
        assert np.array_equal(got == 0, g[name] == 0), name
        np.testing.assert_allclose(got, g[name], rtol=1e-14, atol=0, err_msg=name)
    assert (lh.view(np.uint64) != g["lh"].view(np.uint64)).mean() < 1e-3
    m = len(g["bias_W"])
    sb = np.zeros((m, 4), dtype=np.float32); bqb = np.zeros((m, 4), dtype=np.float32)
    BW, F, R = (np.ascontiguousarray(g[x]) for x in ("bias_W", "bias_F", "bias_R"))
    E.emu_bias(BW.ctypes.data, F.ctypes.data, R.ctypes.data, m, sb.ctypes.data, bqb.ctypes.data)
    assert np.array_equal(sb.view(np.uint32), g["sb"].view(np.uint32))      # includes the early-exit walk
    assert np.array_equal(bqb.view(np.uint32), g["bqb"].view(np.uint32))


def test_expf_reimplementation_is_glibc_expf(E):
    """expf_glibc() must equal the host libm's expf bit for bit (tie decisions of the float Fisher test)"""
    libm = ctypes.CDLL("libm.so.6")
    libm.expf.restype = ctypes.c_float
    libm.expf.argtypes = [ctypes.c_float]
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(-104, 2, 200000), rng.uniform(-1e-3, 1e-3, 2000), [0.0, -0.0, -87.5, -103.9, -104.5, 88.8]])
    for x in xs.astype(np.float32):
        a, b = E.emu_expf(float(x)), libm.expf(float(x))
        assert np.float32(a).view(np.uint32) == np.float32(b).view(np.uint32), x


@pytest.mark.parametrize("name", list(T.CONFIGS))
def test_whole_path_matches_reference_txt(E, tmp_path, name):
    idx, out, region = T.regenerate(name, tmp_path)
    bal, unbal = T.config_ini(name)
    rc = E.emu_run(idx.encode(), (out + ".snapshot").encode(), (region or "").encode(), None, bal, unbal,
                   (out + ".txt").encode())
    assert rc == 0
    assert open(out + ".txt", "rb").read() == T.golden_txt(name)
