"""ctypes binding of libgt16.so (include/gt16.h) -- the Python-side mirror of the reference's
operator seam (FisherExactTest.h:10-23) and of the fused snapshot -> called-sites path that the
snapshotSnpcaller drop-in uses.  Everything computes on the GPU inside libgt16.so; if the
library is not built, or no sm_100 device is usable, these calls raise -- there is no CPU path.
"""
import ctypes
import os

import numpy as np

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
LIB = os.environ.get("GT16_LIB") or os.path.join(PKG, "lib", "libgt16.so")   # GT16_LIB: kernel-variant builds (tools/probes)
CLI = os.path.join(PKG, "bin", "snapshotSnpcaller")
BATCH_SITES = 8388608   # VariantCaller.cpp:526-528

# ---- C ABI bindings --------------------------------------------------------------------------
class Stats(ctypes.Structure):
    _fields_ = [("sites", ctypes.c_uint64), ("flagged", ctypes.c_uint64), ("called", ctypes.c_uint64),
                ("kernel_launches", ctypes.c_uint32), ("ms_h2d", ctypes.c_float), ("ms_stream", ctypes.c_float),
                ("ms_scan", ctypes.c_float), ("ms_eval", ctypes.c_float), ("ms_candidates", ctypes.c_float),
                ("ms_d2h", ctypes.c_float), ("ms_total_device", ctypes.c_float)]


class View(ctypes.Structure):
    _fields_ = [("counters", ctypes.c_void_p), ("text_length", ctypes.c_uint32), ("invalid_bitmap", ctypes.c_void_p),
                ("n_overflow_regions", ctypes.c_uint32), ("overflow", ctypes.c_void_p),
                ("overflow_size", ctypes.c_void_p), ("pool", ctypes.c_void_p), ("pool_bytes", ctypes.c_uint32),
                ("packed_ref", ctypes.c_void_p), ("packed_ref_words", ctypes.c_uint32),
                ("sparse_entries", ctypes.c_void_p), ("n_sparse_entries", ctypes.c_uint64)]


class Reads(ctypes.Structure):            # gt16_reads
    _fields_ = [("n_reads", ctypes.c_uint64), ("start", ctypes.c_void_p), ("reverse", ctypes.c_void_p), ("length", ctypes.c_void_p),
                ("seq_off", ctypes.c_void_p), ("cigar_off", ctypes.c_void_p), ("n_cigar", ctypes.c_void_p),
                ("n_bases", ctypes.c_uint64), ("n_cigar_ops", ctypes.c_uint64), ("bases", ctypes.c_void_p),
                ("quals", ctypes.c_void_p), ("cigar", ctypes.c_void_p)]


class PileupParams(ctypes.Structure):     # gt16_pileup_params
    _fields_ = [("weight_map", ctypes.c_uint8 * 41), ("trim_size", ctypes.c_uint32),
                ("indel_weight_threshold", ctypes.c_uint32), ("n_overflow_regions", ctypes.c_uint32)]


class BuiltSnapshot(ctypes.Structure):    # gt16_built_snapshot
    _fields_ = [("view", View), ("overflow_storage", ctypes.c_void_p), ("ms_pack", ctypes.c_float), ("ms_h2d", ctypes.c_float),
                ("ms_kernel", ctypes.c_float), ("ms_d2h", ctypes.c_float), ("aligned_bases", ctypes.c_uint64),
                ("heavy_tiles", ctypes.c_uint32)]


class Indel(ctypes.Structure):
    _fields_ = [("type", ctypes.c_uint8), ("len", ctypes.c_uint8), ("hq", ctypes.c_uint16), ("lq", ctypes.c_uint16),
                ("seq", ctypes.c_uint8 * 64)]


class CalledSite(ctypes.Structure):      # gt16_called_site, include/gt16.h
    _fields_ = [("amb", ctypes.c_uint32), ("ref", ctypes.c_uint8), ("is_indel", ctypes.c_uint8),
                ("genotype", ctypes.c_uint8), ("rq_assigned", ctypes.c_uint8),
                ("W", ctypes.c_uint32 * 6), ("F", ctypes.c_uint32 * 6), ("R", ctypes.c_uint32 * 6),
                ("lh", ctypes.c_double * 16), ("flh", ctypes.c_double * 16),
                ("pD", ctypes.c_double), ("bestD", ctypes.c_double), ("secondD", ctypes.c_double),
                ("sb", (ctypes.c_float * 4) * 3), ("bqb", (ctypes.c_float * 4) * 3),
                ("rpb", ctypes.c_float), ("lgc", ctypes.c_float), ("rgc", ctypes.c_float), ("asc", ctypes.c_float),
                ("lq", ctypes.c_float), ("rq", ctypes.c_float),
                ("lg", ctypes.c_uint8), ("rg", ctypes.c_uint8), ("polyrun", ctypes.c_uint16),
                ("total_hq", ctypes.c_uint32), ("total_lq", ctypes.c_uint32), ("indel", Indel * 2)]


CALLED_SITE_DTYPE = np.dtype([
    ("amb", "<u4"), ("ref", "u1"), ("is_indel", "u1"), ("genotype", "u1"), ("rq_assigned", "u1"),
    ("W", "<u4", 6), ("F", "<u4", 6), ("R", "<u4", 6), ("lh", "<f8", 16), ("flh", "<f8", 16),
    ("pD", "<f8"), ("bestD", "<f8"), ("secondD", "<f8"), ("sb", "<f4", (3, 4)), ("bqb", "<f4", (3, 4)),
    ("rpb", "<f4"), ("lgc", "<f4"), ("rgc", "<f4"), ("asc", "<f4"), ("lq", "<f4"), ("rq", "<f4"),
    ("lg", "u1"), ("rg", "u1"), ("polyrun", "<u2"), ("total_hq", "<u4"), ("total_lq", "<u4"),
    ("indel", [("type", "u1"), ("len", "u1"), ("hq", "<u2"), ("lq", "<u2"), ("seq", "u1", 64)], 2)], align=True)
assert CALLED_SITE_DTYPE.itemsize == ctypes.sizeof(CalledSite)


def calls_array(L, calls):
    """copy of the called-site records of a gt16_calls handle as a numpy structured array"""
    n = L.gt16_calls_count(calls)
    if n == 0:
        return np.zeros(0, dtype=CALLED_SITE_DTYPE)
    buf = ctypes.string_at(L.gt16_calls_data(calls), n * CALLED_SITE_DTYPE.itemsize)
    return np.frombuffer(buf, dtype=CALLED_SITE_DTYPE).copy()


_lib = None


def capi():
    """libgt16.so through ctypes; raises if the library is missing (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB):
        raise RuntimeError("libgt16.so is not built: run python -c 'import __graft_entry__ as g; g.build()'")
    L = ctypes.CDLL(LIB)
    vp, u32, u64, cp, dbl = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_char_p, ctypes.c_double
    sz, i32 = ctypes.c_size_t, ctypes.c_int
    P = ctypes.POINTER
    sig = {
        "gt16_device_count": (i32, [P(i32)]),
        "gt16_ctx_create": (i32, [i32, P(vp)]),
        "gt16_ctx_destroy": (None, [vp]),
        "gt16_last_error": (cp, [vp]),
        "gt16_copy_genotype_table": (i32, [vp, vp]),
        "gt16_copy_snp_prior": (i32, [vp, vp]),
        "gt16_create_snp_prior": (i32, [vp]),
        "gt16_prefill_likelihood_cache": (i32, [vp, ctypes.c_float, ctypes.c_float, vp]),
        "gt16_configure": (i32, [vp, dbl, dbl, i32]),
        "gt16_setup_times": (i32, [vp, P(ctypes.c_float), P(ctypes.c_float)]),
        "gt16_compute_genotype_likelihood": (i32, [vp, vp, sz, vp, sz, vp, u32, dbl, dbl]),
        "gt16_compute_genotype": (i32, [vp, vp, vp, vp, sz, u32]),
        "gt16_compute_bias": (i32, [vp, vp, vp, vp, sz, u32, vp, vp]),
        "gt16_snapshot_upload": (i32, [vp, P(View), P(vp)]),
        "gt16_snapshot_release": (None, [vp, vp]),
        "gt16_snapshot_upload_range": (i32, [vp, P(View), vp, u32, u64, u64, P(vp)]),
        "gt16_call_sites_host_to": (i32, [vp, P(View), vp, u32, u64, u64, vp, u64, P(u64), P(Stats)]),
        "gt16_snapshot_read_range": (i32, [cp, i32, u32, u32, P(vp), cp, sz]),
        "gt16_call_sites": (i32, [vp, vp, vp, u32, u64, u64, P(vp), P(Stats)]),
        "gt16_call_sites_host": (i32, [vp, P(View), vp, u32, u64, u64, P(vp), P(Stats)]),
        "gt16_call_sites_device": (i32, [vp, vp, vp, u32, u64, u64, P(vp), P(Stats)]),
        "gt16_call_sites_device_nowait": (i32, [vp, vp, vp, u32, u64, u64, P(vp), P(Stats)]),
        "gt16_dev_calls_wait": (i32, [vp, vp, P(Stats)]),
        "gt16_dev_calls_count": (u64, [vp]),
        "gt16_dev_calls_device_ptr": (vp, [vp]),
        "gt16_dev_calls_fetch": (i32, [vp, vp, P(vp)]),
        "gt16_dev_calls_free": (None, [vp]),
        "gt16_calls_count": (u64, [vp]),
        "gt16_calls_data": (vp, [vp]),
        "gt16_calls_free": (None, [vp]),
        "gt16_snapshot_read": (i32, [cp, i32, P(vp), cp, sz]),
        "gt16_snapshot_read_ex": (i32, [cp, i32, u32, u32, P(vp), cp, sz]),
        "gt16_snapshot_view_of": (None, [vp, P(View)]),
        "gt16_snapshot_free": (None, [vp]),
        "gt16_index_read": (i32, [cp, P(vp), cp, sz]),
        "gt16_index_length": (u32, [vp]),
        "gt16_index_packed": (vp, [vp, P(u32)]),
        "gt16_index_translate": (cp, [vp, u32, P(u32)]),
        "gt16_index_locate": (i32, [vp, cp, u32, P(u32)]),
        "gt16_index_pin": (i32, [vp]),
        "gt16_format_number": (i32, [ctypes.c_double, i32, cp]),
        "gt16_index_free": (None, [vp]),
        "gt16_write_txt": (i32, [cp, vp, vp, u64, i32, vp]),
        "gt16_index_sequences": (u32, [vp]),
        "gt16_index_sequence_id": (i32, [vp, cp]),
        "gt16_index_locate_id": (i32, [vp, i32, u32, P(u32)]),
        "gt16_pileup_default_params": (None, [vp]),
        "gt16_build_snapshot": (i32, [vp, vp, u32, vp, vp]),
        "gt16_built_snapshot_free": (None, [vp]),
        "gt16_snapshot_write": (i32, [cp, P(View), cp, sz]),
        "gt16_sam_read": (i32, [cp, vp, u32, P(vp), cp, sz]),
        "gt16_bam_read": (i32, [cp, vp, u32, P(vp), cp, sz]),
        "gt16_sam_reads": (None, [vp, vp]),
        "gt16_sam_total_alignments": (u64, [vp]),
        "gt16_sam_free": (None, [vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    L._signatures = sig
    _lib = L
    return L


class Caller:
    """Thin RAII-ish wrapper used by tests / bench / smoke: one context on one device."""

    def __init__(self, device=0, balance=0.01, unbalance=0.005, decision_tables=True):
        self.L = capi()
        self.ctx = ctypes.c_void_p()
        rc = self.L.gt16_ctx_create(device, ctypes.byref(self.ctx))
        if rc != 0:
            raise RuntimeError("gt16_ctx_create: %d %s" % (rc, self.L.gt16_last_error(None).decode()))
        self.check(self.L.gt16_configure(self.ctx, balance, unbalance, 1 if decision_tables else 0))

    def check(self, rc):
        if rc != 0:
            raise RuntimeError("libgt16 error %d: %s" % (rc, self.L.gt16_last_error(self.ctx).decode()))

    def close(self):
        if self.ctx:
            self.L.gt16_ctx_destroy(self.ctx)
            self.ctx = ctypes.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


class LoadedInput:
    """index + snapshot read through the library's own readers.

    snapshot_path=None: index (and region list) only.  pos_range=(lo, hi): keep the counters of genome positions
    [lo, hi) only (gt16_snapshot_read_range: one shard of a sharded genome).  index=<LoadedInput>: share that input's
    index instead of reading it again."""

    def __init__(self, idx_prefix, snapshot_path, region_bin=None, pinned=1, pos_range=None, index=None, keep_sparse=False):
        L = capi()
        self.L = L
        err = ctypes.create_string_buffer(512)
        self.ix = ctypes.c_void_p()
        self.hs = ctypes.c_void_p()
        self.owns_index = index is None
        if index is None:
            if L.gt16_index_read(idx_prefix.encode(), ctypes.byref(self.ix), err, 512) != 0:
                raise RuntimeError(err.value.decode())
            if pinned:
                L.gt16_index_pin(self.ix)        # best effort: without a device the reference stays pageable
        else:
            self.ix = index.ix
        self.length = L.gt16_index_length(self.ix)
        self.view = View()
        if snapshot_path is not None:
            if keep_sparse:      # a sparse snapshot stays the list of {position, counter} records it is (GT16_READ_KEEP_SPARSE)
                lo, hi = (0, 0xFFFFFFFF) if pos_range is None else (int(pos_range[0]), int(pos_range[1]))
                rc = L.gt16_snapshot_read_ex(snapshot_path.encode(), (1 if pinned else 0) | 2, lo, hi, ctypes.byref(self.hs), err, 512)
            elif pos_range is None:
                rc = L.gt16_snapshot_read(snapshot_path.encode(), pinned, ctypes.byref(self.hs), err, 512)
            else:
                rc = L.gt16_snapshot_read_range(snapshot_path.encode(), pinned, int(pos_range[0]), int(pos_range[1]),
                                                ctypes.byref(self.hs), err, 512)
            if rc != 0:
                raise RuntimeError(err.value.decode())
            L.gt16_snapshot_view_of(self.hs, ctypes.byref(self.view))
        n = ctypes.c_uint32()
        self.view.packed_ref = L.gt16_index_packed(self.ix, ctypes.byref(n))
        self.view.packed_ref_words = n.value
        self.regions = None
        self.n_regions = 0
        self.total_sites = self.view.text_length if snapshot_path is not None else self.length
        if region_bin:
            raw = np.fromfile(region_bin, dtype=np.uint32)
            self.n_regions = int(raw[0])
            self.regions = np.ascontiguousarray(raw[1:1 + 2 * self.n_regions])
            r = self.regions.reshape(-1, 2).astype(np.int64)
            self.total_sites = int(np.maximum(r[:, 1] - r[:, 0], 0).sum())

    def regions_ptr(self):
        return self.regions.ctypes.data if self.regions is not None else None

    def close(self):
        if self.hs:
            self.L.gt16_snapshot_free(self.hs)
            self.hs = ctypes.c_void_p()
        if self.ix and self.owns_index:
            self.L.gt16_index_free(self.ix)
        self.ix = ctypes.c_void_p()


def call_to_txt(caller, inp, txt_path, shards=1, resident=False):
    """Run the fused path over all processed sites in `shards` batch-aligned shards and write .txt."""
    L = caller.L
    from .sharding import shard_cuts
    cuts = shard_cuts(inp.total_sites, shards)
    shards = len(cuts) - 1
    carry = (ctypes.c_float * 2)(0.0, 0.0)
    stats_all = []
    dev = ctypes.c_void_p()
    if resident:
        caller.check(L.gt16_snapshot_upload(caller.ctx, ctypes.byref(inp.view), ctypes.byref(dev)))
    try:
        for d in range(shards):
            calls = ctypes.c_void_p()
            st = Stats()
            if resident:
                rc = L.gt16_call_sites(caller.ctx, dev, inp.regions_ptr(), inp.n_regions, cuts[d], cuts[d + 1],
                                       ctypes.byref(calls), ctypes.byref(st))
            else:
                rc = L.gt16_call_sites_host(caller.ctx, ctypes.byref(inp.view), inp.regions_ptr(), inp.n_regions,
                                            cuts[d], cuts[d + 1], ctypes.byref(calls), ctypes.byref(st))
            caller.check(rc)
            rc = L.gt16_write_txt(txt_path.encode(), inp.ix, L.gt16_calls_data(calls), L.gt16_calls_count(calls),
                                  1 if d else 0, carry)
            L.gt16_calls_free(calls)
            assert rc == 0
            stats_all.append(st)
    finally:
        if resident:
            L.gt16_snapshot_release(caller.ctx, dev)
    return stats_all
