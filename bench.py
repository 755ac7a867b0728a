#!/usr/bin/env python
"""bench.py -- headline benchmark of the 16GT calling hot path on B200.

Metric (BASELINE.json): called sites/sec at 30x WGS = genome positions evaluated per second (SURVEY
8d: the unit of work is one position; "Called Entries" is reported beside it), plus the HBM
roofline fraction of the likelihood (streaming) kernel.

Workload at N=1 = BASELINE.json configs[1]: one 64,444,167-bp chr20-sized synthetic contig, 30x WGS
snapshot (tools/gt16_synth, seed 1 + rank).  At N>1 every rank calls its own contig of that size
(independent genome regions, no collective on the data path): weak scaling.

A step = one pass of the hot path over the whole snapshot:
  value : snapshot resident in HBM when the timed region starts (gt16_call_sites)
  e2e   : through the C ABI with HOST buffers (gt16_call_sites_host): pinned counters stream to HBM on
          the copy stream inside the timed region, called-site records come back to the host
  --impl reference : the UNMODIFIED reference objects (oracle/_ref/ref_caller = selectPossibleSNPs
          etc. compiled from /root/reference) on the host cores, on a bounded sample of the workload

One JSON line on stdout (rank 0).
"""
import argparse
import ctypes
import importlib
import json
import os
import re
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sites_per_sec_30x_wgs"
UNIT = "sites/s"
WORKLOAD_SITES = 64444167            # chr20-sized, deliberately not a multiple of 8,388,608
# Bounded CPU sample of the same workload: half a chr20 (not a multiple of 8,388,608).  It has to be this large for
# the reference's fixed per-run costs (18 GB of batch buffers, sleep(1) hand-offs) to amortise: on a
# 4 M-site sample the reference reads 1.3-1.9 M sites/s, on this one and on the full workload 5-6 M sites/s.
SAMPLE_SITES = 32222083
BYTES_PER_SITE = 8.375               # 8 B SnpCounter + 2-bit reference + 1 invalid bit
SYNTH = os.path.join(ROOT, "tools", "gt16_synth")
REF_CALLER = os.path.join(ROOT, "oracle", "_ref", "ref_caller")


def synth(workdir, sites, seed, threads):
    idx, out = os.path.join(workdir, "g"), os.path.join(workdir, "s")
    subprocess.run([SYNTH, "--index", idx, "--snapshot", out, "--contigs", "chr20:%d" % sites, "--seed", str(seed),
                    "--threads", str(threads)], check=True, stderr=subprocess.DEVNULL)
    return idx, out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.th.join(timeout=2)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_reference_once(idx, out, threads, lc_cache):
    """one pass of the reference's own CPU path; returns (select_seconds, prefill_seconds, called); the operator-only
    time the reference itself reports with -v is left in run_reference_once.operators"""
    cmd = [REF_CALLER, idx, out, "-v", "-t", str(threads)]
    if lc_cache:
        cmd += ["-c", lc_cache]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, check=True)
    sel = float(re.search(r"REF_SELECT_SECONDS\s+([0-9.]+)", r.stdout).group(1))
    pre = float(re.search(r"REF_PREFILL_SECONDS\s+([0-9.]+)", r.stdout).group(1))
    called = int(re.search(r"Called Entries : (\d+)", r.stdout).group(1))
    m = re.search(r"Total Likelihood and Bias computing time :\s+([0-9.]+)", r.stdout)
    run_reference_once.operators = float(m.group(1)) if m else None
    return sel, pre, called


def ref_threads():
    return max(1, min(os.cpu_count() or 1, 99))      # the reference's ini parser rejects >= 100


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if not os.path.exists(REF_CALLER):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_caller not built"}), flush=True)
        return 0
    wd = tempfile.mkdtemp(prefix="gt16_refarm_")
    try:
        threads = ref_threads()
        idx, out = synth(wd, SAMPLE_SITES, 1, min(threads, 32))
        cache = os.path.join(wd, "lc.bin")        # the reference's own tables, built by its own prefill in run 1
        times, prefill, called = [], None, 0
        for it in range(args.warmup + args.steps):
            sel, pre, called = run_reference_once(idx, out, threads, cache)
            if it == 0:
                prefill = pre
            if it >= args.warmup:
                times.append(sel)
        total = sum(times)
        value = SAMPLE_SITES * len(times) / total
        sample = ("first-contig sample of %d sites of the same 30x synthetic workload per step; selectPossibleSNPs "
                  "wall time (serial decode loop + OpenMP operators + consumer thread); the one-off table prefill "
                  "(%.1f s on %d threads) is outside the timed steps, like the GPU arm's table build" %
                  (SAMPLE_SITES, prefill, threads))
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1] (64,444,167-bp chr20-sized 30x WGS snapshot), bounded sample: %d sites "
                                   "per step on the host CPU" % SAMPLE_SITES, "sites_per_step": SAMPLE_SITES,
                       "called_entries": called},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    return 0


def bind_to_gpu_numa_node(device):
    """Run this rank (and allocate its page-locked buffers) on the CPUs next to its GPU: with one process per GPU on
    a two-socket box, host<->device copies from the far socket cross the inter-socket link and collapse when all
    ranks copy at once.  Uses NVML's CPU affinity for the device; silently does nothing if that is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(device)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if cpus:
            os.sched_setaffinity(0, cpus)
        return sorted(cpus)
    except Exception:
        return None


def measure_host_link(torch, nbytes):
    """What the host link of this box delivers: one pinned->device and one device->pinned copy of nbytes, CUDA
    events, best of 5.  Reported beside e2e (which moves 545 MB in and 46 MB out per step) as its floor."""
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    best = {"h2d": 1e9, "d2h": 1e9}
    for _ in range(6):
        for kind in ("h2d", "d2h"):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            if kind == "h2d":
                dev.copy_(host, non_blocking=True)
            else:
                host.copy_(dev, non_blocking=True)
            b.record()
            b.synchronize()
            best[kind] = min(best[kind], a.elapsed_time(b))
    return {"h2d_gbs": nbytes / best["h2d"] / 1e6, "d2h_gbs": nbytes / best["d2h"] / 1e6, "bytes": nbytes}


def main():
    # exactly ONE line on stdout: libraries (NCCL's version banner, ...) that print to fd 1 are sent to stderr
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sites", type=int, default=WORKLOAD_SITES, help="override the workload size (debug only)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        sys.stderr.write("bench.py: no CUDA device -- the hot path has no CPU implementation\n")
        return 2
    torch.cuda.set_device(local)
    bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    capi = importlib.import_module("16gt_b200.capi")
    sharding = importlib.import_module("16gt_b200.sharding")
    L = capi.capi()
    wd = tempfile.mkdtemp(prefix="gt16_bench_%d_" % rank)
    try:
        threads = max(1, min(32, (os.cpu_count() or 8) // max(1, world)))
        idx, out = synth(wd, args.sites, 1 + rank, threads)
        caller = capi.Caller(local)
        pre_ms, dec_ms = ctypes.c_float(), ctypes.c_float()
        L.gt16_setup_times(caller.ctx, ctypes.byref(pre_ms), ctypes.byref(dec_ms))
        inp = capi.LoadedInput(idx, out + ".snapshot", None, pinned=1)
        sites = inp.total_sites
        dev = ctypes.c_void_p()
        caller.check(L.gt16_snapshot_upload(caller.ctx, ctypes.byref(inp.view), ctypes.byref(dev)))

        pending, completed = [], []          # calls in flight (handle, n); final stats of the completed ones, in call order

        def retire(count):
            """gt16_dev_calls_wait + free for the `count` oldest calls in flight (record count, error word, CUDA-event timings)"""
            for calls in pending[:count]:
                st = capi.Stats()
                caller.check(L.gt16_dev_calls_wait(caller.ctx, calls, ctypes.byref(st)))
                L.gt16_dev_calls_free(calls)
                completed.append((st, st.called))
            del pending[:count]

        def step_resident():
            # inputs resident in HBM, records left in HBM: the device-side pipeline.  gt16_call_sites_device_nowait returns
            # when the number of called sites is known and the records + Fisher tests are enqueued, so the host prepares
            # step k+1 while the GPU finishes step k; a step is completed (gt16_dev_calls_wait) after the timed region's
            # closing synchronize -- or, in a run of more than 96 steps, once it is 48 steps old and long finished (the
            # library keeps the events of the last 128 calls)
            if len(pending) >= 96:
                retire(48)
            calls, st = ctypes.c_void_p(), capi.Stats()
            caller.check(L.gt16_call_sites_device_nowait(caller.ctx, dev, None, 0, 0, sites, ctypes.byref(calls), ctypes.byref(st)))
            pending.append(calls)
            return st, None

        def complete_resident(stats):
            """complete every step still in flight (after the timed region) and return the final stats of the last len(stats)"""
            retire(len(pending))
            done = completed[-len(stats):] if stats else []
            del completed[:]
            return done

        def step_resident_host_results():
            # same, but the records are also copied to page-locked host memory inside the step (gt16_call_sites)
            calls, st = ctypes.c_void_p(), capi.Stats()
            caller.check(L.gt16_call_sites(caller.ctx, dev, None, 0, 0, sites, ctypes.byref(calls), ctypes.byref(st)))
            n = L.gt16_calls_count(calls)
            L.gt16_calls_free(calls)
            return st, n

        def step_host():
            calls, st = ctypes.c_void_p(), capi.Stats()
            caller.check(L.gt16_call_sites_host(caller.ctx, ctypes.byref(inp.view), None, 0, 0, sites,
                                                ctypes.byref(calls), ctypes.byref(st)))
            n = L.gt16_calls_count(calls)
            L.gt16_calls_free(calls)
            return st, n

        def timed(fn, warmup, steps):
            for _ in range(warmup):
                fn()
            barrier()
            t0 = time.perf_counter()
            stats = [fn() for _ in range(steps)]
            barrier()
            dt = time.perf_counter() - t0
            dt = sharding.max_over_ranks(dt, dist if world > 1 else None, "cuda")
            return dt, stats

        # clocks / throttle reasons are sampled over the whole measurement phase (warm-ups and the three timed loops
        # are all under load; one timed loop alone lasts only milliseconds, less than nvidia-smi's sampling period)
        sampler = ClockSampler(local) if rank == 0 else None
        if sampler:
            sampler.start()
            for _ in range(40):            # keep the GPU busy until the sampler has certainly started
                step_resident()
            complete_resident([(None, 0)] * 40)
        dt, stats = timed(step_resident, args.warmup, args.steps)
        stats = complete_resident(stats)
        dt_hr, stats_hr = timed(step_resident_host_results, 1, max(2, min(args.steps, 5)))
        dt_e2e, stats_e2e = timed(step_host, 2, max(2, args.steps))
        clocks = sampler.stop() if sampler else None
        n_e2e = len(stats_e2e)
        link = measure_host_link(torch, min(sites * 8, 1 << 29))

        launches = sum(s.kernel_launches for s, _ in stats)
        called = stats[0][1]
        flagged = stats[0][0].flagged
        ms_stream = sum(s.ms_stream for s, _ in stats) / len(stats)
        ms_dev = sum(s.ms_total_device for s, _ in stats) / len(stats)
        parts = {k: sum(getattr(s, k) for s, _ in stats) / len(stats)
                 for k in ("ms_stream", "ms_scan", "ms_eval", "ms_candidates", "ms_d2h")}
        v = inp.view
        ovf_records = 0
        if v.n_overflow_regions:
            arr = (ctypes.c_uint32 * v.n_overflow_regions).from_address(v.overflow_size)
            ovf_records = sum(arr)
        h2d = sites * 8 + ((sites + 31) // 32) * 4 + v.packed_ref_words * 4 + ovf_records * 32 + v.pool_bytes
        d2h = called * ctypes.sizeof(capi.CalledSite) + 8

        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = sites * BYTES_PER_SITE / (ms_stream * 1e-3) / 1e9

        line = {
            "metric": METRIC, "value": world * sites * args.steps / dt, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[1]: 64,444,167-bp chr20-sized synthetic contig, 30x WGS snapshot, one "
                                   "contig of that size per GPU (seed 1+rank)",
                       "sites_per_step_per_gpu": sites, "called_entries_per_gpu": called,
                       "exactly_evaluated_sites_per_gpu": flagged,
                       "l2": "inputs (%.0f MB of counters per pass) exceed the 126 MB L2; no flush needed" % (sites * 8 / 1e6),
                       "parallelism": "region-sharded x%d, no collective on the data path" % world,
                       "setup_ms": {"table_prefill": pre_ms.value, "decision_bitmaps": dec_ms.value}},
            "e2e": {"value": world * sites * n_e2e / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": 1000.0 * dt_e2e / n_e2e,
                    "host_link": dict(link, h2d_floor_ms=h2d / (link["h2d_gbs"] * 1e6),
                                      note="plain cudaMemcpyAsync from page-locked memory on this box, best of 5: the "
                                           "time the step's input bytes need on the bus whatever the kernels do")},
            "value_with_host_results": {"value": world * sites * len(stats_hr) / dt_hr, "unit": UNIT,
                                        "ms_per_step": 1000.0 * dt_hr / len(stats_hr), "d2h_bytes_per_step": d2h,
                                        "note": "value = inputs AND called-site records resident in HBM (gt16_call_sites_device_nowait); "
                                                "this line adds the copy of the records to page-locked host memory"},
            "gpu_launches": launches,
            "device_ms_per_step": ms_dev, "device_ms_breakdown": parts,
            "roofline": {"bound": "hbm", "kernel": "k_stream (decode + 16-genotype classification of every site)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": sites * BYTES_PER_SITE, "avg_launch_ms": ms_stream},
            "clocks": clocks,
        }
        tr = os.path.join(ROOT, "profiles", "stream_kernel_traffic.json")
        if os.path.exists(tr):
            try:
                line["roofline"]["traffic"] = json.load(open(tr)).get("dram_bytes_per_launch")
            except (ValueError, OSError):
                pass

        if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(REF_CALLER):
            swd = tempfile.mkdtemp(prefix="gt16_cpu_")
            try:
                th = ref_threads()
                sidx, sout = synth(swd, SAMPLE_SITES, 1, min(th, 32))
                sel, pre, ccalled = run_reference_once(sidx, sout, th, None)
                line["cpu_baseline"] = {
                    "value": SAMPLE_SITES / sel, "unit": UNIT, "cores": th, "kind": "reference",
                    "sample": "oracle/_ref/ref_caller (unmodified reference objects) on the first %d sites of the same "
                              "synthetic 30x workload: selectPossibleSNPs %.2f s (serial decode loop + OpenMP operators "
                              "on %d threads + consumer thread), %d called; of that, the operators alone (the reference's "
                              "own 'Total Likelihood and Bias computing time'): %s s; one-off table prefill %.1f s not included"
                              % (SAMPLE_SITES, sel, th, ccalled,
                                 "%.2f" % run_reference_once.operators if run_reference_once.operators is not None else "n/a", pre),
                    "select_seconds": sel, "operator_seconds": run_reference_once.operators, "prefill_seconds": pre}
            finally:
                shutil.rmtree(swd, ignore_errors=True)
        elif rank == 0 and world == 1:
            line["cpu_baseline"] = None

        L.gt16_snapshot_release(caller.ctx, dev)
        inp.close()
        caller.close()
        if rank == 0:
            print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(wd, ignore_errors=True)
        if world > 1:
            dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
