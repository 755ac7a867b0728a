"""CPU tests of the test oracle itself (oracle/gt16_oracle.c, the plain-C restatement of the reference
algorithm): it is pinned against golden vectors produced by the UNMODIFIED reference objects
(tests/golden/, made by tests/golden/make_golden.py) before anything else is allowed to trust it."""
import ctypes
import os

import numpy as np
import pytest

import gt16_testlib as T

vp = ctypes.c_void_p


@pytest.fixture(scope="module")
def O():
    lib = ctypes.CDLL(T.ORACLE_LIB)
    lib.oracle_run.argtypes = [ctypes.c_char_p] * 3 + [ctypes.c_double, ctypes.c_double, ctypes.c_char_p,
                                                       ctypes.POINTER(ctypes.c_ulonglong)]
    lib.oracle_table_entry.restype = ctypes.c_double
    lib.oracle_table_entry.argtypes = [ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    lib.oracle_fisher2.restype = ctypes.c_double
    lib.oracle_hyper.restype = ctypes.c_double
    lib.oracle_chisqr.restype = ctypes.c_double
    lib.oracle_chisqr.argtypes = [ctypes.c_int, ctypes.c_double]
    lib.oracle_chi2dist.restype = ctypes.c_double
    lib.oracle_chi2dist.argtypes = [ctypes.c_double] * 4
    lib.oracle_call.argtypes = [ctypes.c_double, ctypes.c_double, vp, vp, ctypes.c_uint] + [vp] * 6
    lib.oracle_bias.argtypes = [vp, vp, vp, ctypes.c_uint, vp, vp]
    lib.oracle_prior.argtypes = [vp]
    return lib


def oracle_call(O, W, ref, balance=0.01, unbalance=0.005):
    n = len(ref)
    W = np.ascontiguousarray(W, dtype=np.uint32); ref = np.ascontiguousarray(ref, dtype=np.uint8)
    out = dict(lh=np.zeros((n, 16)), flh=np.zeros((n, 16)), pD=np.zeros(n), bestD=np.zeros(n), secondD=np.zeros(n),
               genotype=np.zeros(n, dtype=np.int32))
    O.oracle_call(balance, unbalance, W.ctypes.data, ref.ctypes.data, n, out["lh"].ctypes.data, out["flh"].ctypes.data,
                  out["pD"].ctypes.data, out["bestD"].ctypes.data, out["secondD"].ctypes.data, out["genotype"].ctypes.data)
    return out


def test_known_answers_of_the_reference_objects(O):
    """values printed by throw-away drivers linked against the reference's own .o files"""
    assert O.oracle_fisher2(3, 10, 9, 10) == 0.26725312458977984
    assert O.oracle_fisher2(120, 120, 118, 120) == 0.92742489228285319
    assert O.oracle_fisher2(0, 120, 2, 120) == 0.4979596035801177
    assert O.oracle_hyper(3, 13, 12, 43) == 0.26676645989789538
    assert O.oracle_chisqr(1, 3.7) == 0.032611022214010063
    assert O.oracle_chi2dist(2500, 2600, 2567, 2600) == 0.42266452720581277
    e = lambda k, i, j: O.oracle_table_entry(0.01, 0.005, k, i, j)
    assert e(1, 120, 120) == 0.99999999999931721 and e(2, 120, 0) == 0.99999999999998801
    assert e(5, 60, 58) == 0.022650099892931784 and e(6, 120, 3) == 0.62194559704920249
    assert e(9, 60, 0) == 5.5813637451038962e-08
    r = oracle_call(O, [[120, 0, 0, 0, 0, 0], [60, 0, 58, 0, 0, 0], [3, 0, 117, 0, 0, 0], [0] * 6, [70, 0, 0, 2, 48, 0],
                        [40, 1, 0, 0, 36, 28]], [0, 0, 0, 2, 0, 0])
    assert r["genotype"].tolist() == [0, 2, 7, 7, 14, 15]
    np.testing.assert_allclose(r["pD"], [1.0000003582e+00, 1.8644055190e-02, 5.7671370776e-01, 16, 1.7495204709e-01,
                                         8.4012088789e-02], rtol=2e-10)
    np.testing.assert_allclose(r["bestD"], [9.9999964181e-01, 2.1e-03, 2.2049912172e-03, 6.25e-02, 4.6458831455e-05,
                                            4.9999979884e-09], rtol=2e-10)


def test_tables_match_the_reference_prefill(O):
    """entries of the reference's own LikelihoodCache (incl. the float index arithmetic that differs from
    double at rows 120, 200, ...) and its chi-square table: bit for bit"""
    g = np.load(os.path.join(T.GOLDEN, "lc_samples.npz"))
    k, i, j, want = g["k"], g["i"], g["j"], g["value"]
    got = np.array([O.oracle_table_entry(0.01, 0.005, int(a), int(b), int(c)) for a, b, c in zip(k, i, j)])
    assert np.array_equal(got.view(np.uint64), want.view(np.uint64))
    ycsq = np.array([O.oracle_chisqr(1, t / 10.0) for t in range(15000)])
    assert np.array_equal(ycsq.view(np.uint64), g["ycsq"].view(np.uint64))


def test_operators_match_the_reference(O):
    """computeGenotypeLikelihoodWrapper + computeGenotypeWrapper + computeStrandBias/BaseQualityBias outputs
    of the reference objects on 6000 seeded sites covering all 16 genotypes: bit for bit"""
    g = np.load(os.path.join(T.GOLDEN, "operator_vectors.npz"))
    r = oracle_call(O, g["W"], g["ref"])
    assert np.array_equal(r["genotype"], g["genotype"])
    for f in ("lh", "flh", "pD", "bestD", "secondD"):
        assert np.array_equal(r[f].view(np.uint64), g[f].view(np.uint64)), f
    prior = np.zeros(256)
    O.oracle_prior(prior.ctypes.data)
    assert np.array_equal(prior, g["prior"])
    m = len(g["bias_W"])
    sb = np.zeros((m, 4), dtype=np.float32); bqb = np.zeros((m, 4), dtype=np.float32)
    W, F, R = (np.ascontiguousarray(g[x]) for x in ("bias_W", "bias_F", "bias_R"))
    O.oracle_bias(W.ctypes.data, F.ctypes.data, R.ctypes.data, m, sb.ctypes.data, bqb.ctypes.data)
    assert np.array_equal(sb.view(np.uint32), g["sb"].view(np.uint32))
    assert np.array_equal(bqb.view(np.uint32), g["bqb"].view(np.uint32))


@pytest.mark.parametrize("name", list(T.CONFIGS))
def test_whole_path_matches_the_reference_txt(O, tmp_path, name):
    """snapshot in, .txt out: byte-identical to what oracle/_ref/ref_caller wrote for the same input"""
    idx, out, region = T.regenerate(name, tmp_path)
    n = ctypes.c_ulonglong()
    bal, unbal = T.config_ini(name)
    rc = O.oracle_run(idx.encode(), (out + ".snapshot").encode(), (region or "").encode(), bal, unbal,
                      (out + ".txt").encode(), ctypes.byref(n))
    assert rc == 0 and n.value == T.manifest()[name]["rows"]
    assert open(out + ".txt", "rb").read() == T.golden_txt(name)


@pytest.mark.skipif(not os.path.exists(T.REF_OPS), reason="oracle/_ref not built (needs /root/reference)")
def test_against_live_reference_objects(O):
    """when the reference objects are present: random 2x2 tables through the reference's own fisher2 / chisqr"""
    R = ctypes.CDLL(T.REF_OPS)
    R.refops_fisher2.restype = ctypes.c_double
    R.refops_chisqr.restype = ctypes.c_double
    R.refops_chisqr.argtypes = [ctypes.c_int, ctypes.c_double]
    R.refops_chi2dist.restype = ctypes.c_double
    R.refops_chi2dist.argtypes = [ctypes.c_double] * 4
    rng = np.random.default_rng(3)
    for a, b, c, d in rng.integers(0, 400, (3000, 4)):
        assert O.oracle_fisher2(int(a), int(b), int(c), int(d)) == R.refops_fisher2(int(a), int(b), int(c), int(d))
    for a, b, c, d in rng.integers(1, 5000, (2000, 4)).astype(float):
        assert O.oracle_chi2dist(a, b, c, d) == R.refops_chi2dist(a, b, c, d)
    for cv in rng.random(500) * 1500:
        assert O.oracle_chisqr(1, cv) == R.refops_chisqr(1, cv)
