#!/usr/bin/env python
"""bench.py -- headline benchmark of the 16GT calling hot path on B200.

Metric (BASELINE.json): called sites/sec at 30x WGS = genome positions evaluated per second (the unit of
work is one position; "Called Entries" is reported beside it), plus the HBM roofline fraction of the likelihood
(streaming) kernel.

Workload = BASELINE.json configs[2], the configuration the metric is quoted on: ONE 3.1 Gb human-sized synthetic genome
(24 contigs, 30x, dense snapshot: 24.8 GB of counters, 370 batches of 8,388,608 sites, positions beyond 2^31), made
once per box by tools/gt16_synth.  With N GPUs the genome is PARTITIONED: rank r calls the contiguous, batch-aligned
range sharding.my_shard(total, r, N) -- strong scaling, no collective on the data path; the only communication is the
final ordered gather of the calls (each rank's records land in its slab of one shared, rank-ordered host buffer;
the counts are all-gathered) and the max-over-ranks of the measured time.

A step = one pass of the hot path over the whole genome (every rank over its shard):
  value    : the shard's counters resident in HBM when the timed region starts, records left in HBM
             (gt16_call_sites_device_nowait: 7 kernel launches per step, no host round trip)
  e2e      : through the C ABI with HOST buffers (gt16_call_sites_host_to): page-locked counters stream to HBM on the copy
             stream inside the timed region, called-site records come back into the rank-ordered host buffer
  e2e_file : the drop-in binary (16gt_b200/bin/snapshotSnpcaller) on the snapshot FILE, all N GPUs: file -> .txt, wall clock
  --impl reference : the UNMODIFIED reference objects (oracle/_ref/ref_caller = selectPossibleSNPs etc. compiled from
             /root/reference) on the host cores; a step = the full 64,444,167-site chr20-sized workload (configs[1], same
             synthesiser parameters) -- the whole 3.1 Gb genome takes the reference 25 minutes per pass

One JSON line on stdout (rank 0).  --workload chr20 | exome select the other configurations (debug / extra lines).
"""
import argparse
import ctypes
import hashlib
import importlib
import json
import mmap
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
BYTES_PER_SITE = 8.375               # 8 B SnpCounter + 2-bit reference + 1 invalid bit
SYNTH = os.path.join(ROOT, "tools", "gt16_synth")
REF_CALLER = os.path.join(ROOT, "oracle", "_ref", "ref_caller")
CLI = os.path.join(ROOT, "16gt_b200", "bin", "snapshotSnpcaller")
CHR20_SITES = 64444167               # configs[1]; deliberately not a multiple of 8,388,608

WORKLOADS = {
    # BASELINE.json configs[2] (tests/golden_configs.py: wgs_full -- the GPU suite checks its .txt against the reference's)
    "wgs": dict(synth=["--genome-len", "3100000000", "--n-contigs", "24", "--gaps", "8:100000", "--seed", "3"], exome=False,
                name="configs[2]: 3.1 Gb human-sized synthetic genome, 24 contigs, 30x WGS, dense snapshot (24.8 GB of counters, "
                     "370 batches), region-sharded"),
    # configs[1]
    "chr20": dict(synth=["--contigs", "chr20:%d" % CHR20_SITES, "--seed", "1"], exome=False,
                  name="configs[1]: 64,444,167-bp chr20-sized synthetic contig, 30x WGS snapshot"),
    # configs[3] (exome_full): 200,000 targets of ~300 bp at 150x on the 3.1 Gb genome, sparse snapshot, -e region.bin
    "exome": dict(synth=["--genome-len", "3100000000", "--n-contigs", "24", "--exome", "200000:300", "--depth", "150", "--seed", "6"],
                  exome=True, name="configs[3]: exome mode, 200,000 targets x 300 bp at 150x on the 3.1 Gb genome, sparse snapshot, "
                                   "-e region.bin"),
}


def workload_dir(key, synth_args):
    """where the workload's files live: a per-box cache in /dev/shm (RAM: the snapshot is 25.5 GB and every rank reads
    its shard from it) when there is room, else a directory under /tmp"""
    tag = hashlib.sha256(" ".join(synth_args).encode()).hexdigest()[:12]
    for base in (os.environ.get("GT16_BENCH_DIR"), "/dev/shm", tempfile.gettempdir()):
        if not base or not os.path.isdir(base):
            continue
        d = os.path.join(base, "gt16_bench_%s_%s" % (key, tag))
        if os.path.exists(os.path.join(d, ".done")) or shutil.disk_usage(base).free > (64 << 30):
            return d
    return os.path.join(tempfile.gettempdir(), "gt16_bench_%s_%s" % (key, tag))


def ensure_workload(key, rank):
    """rank 0 synthesises (once per box: the driver's N = 1, 2, 4, 8 runs share it); the others wait for the marker"""
    cfg = WORKLOADS[key]
    d = workload_dir(key, cfg["synth"])
    idx, out = os.path.join(d, "g"), os.path.join(d, "s")
    region = os.path.join(d, "region.bin") if cfg["exome"] else None
    done = os.path.join(d, ".done")
    t0 = time.time()
    if rank == 0 and not os.path.exists(done):
        shutil.rmtree(d, ignore_errors=True)
        os.makedirs(d, exist_ok=True)
        argv = [SYNTH, "--index", idx, "--snapshot", out] + cfg["synth"] + ["--threads", str(max(1, min(32, os.cpu_count() or 8)))]
        if region:
            argv += ["--region-out", region]
        subprocess.run(argv, check=True, stderr=subprocess.DEVNULL)
        open(done, "w").write("ok\n")
    while not os.path.exists(done):
        if time.time() - t0 > 1200:
            raise RuntimeError("workload was not synthesised within 20 minutes")
        time.sleep(0.5)
    return idx, out, region, time.time() - t0


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


REF_NOTE = ("the reference needs 25 minutes (8 cores) for one pass over configs[2], so its step is the full configs[1] workload "
            "-- one 64,444,167-site chr20-sized contig made by the same synthesiser with the same parameters (30x, same variant "
            "rates); its per-site cost does not depend on which contig it walks")


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if not os.path.exists(REF_CALLER):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_caller not built"}), flush=True)
        return 0
    threads = ref_threads()
    idx, out, _, _ = ensure_workload("chr20", 0)
    wd = tempfile.mkdtemp(prefix="gt16_refarm_")
    try:
        cache = os.path.join(wd, "lc.bin")        # the reference's own tables, built by its own prefill in run 1
        times, prefill, called = [], None, 0
        for it in range(args.warmup + args.steps):
            sel, pre, called = run_reference_once(idx, out, threads, cache)
            if it == 0:
                prefill = pre
            if it >= args.warmup:
                times.append(sel)
        total = sum(times)
        value = CHR20_SITES * len(times) / total
        sample = ("%d sites per step (the full configs[1] contig): selectPossibleSNPs wall time (serial decode loop + OpenMP "
                  "operators on %d threads + consumer thread); the one-off table prefill (%.1f s) is outside the timed steps, "
                  "like the GPU arm's table build; %s" % (CHR20_SITES, threads, prefill, REF_NOTE))
        line = {
            "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * total / len(times),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOADS["wgs"]["name"] + " -- reference arm: bounded sample, " + REF_NOTE,
                       "sites_per_step": CHR20_SITES, "called_entries": called},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "reference", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line), flush=True)
    finally:
        shutil.rmtree(wd, ignore_errors=True)
    return 0


def bind_to_gpu_cpus(torch, device):
    """Run this rank (and allocate its page-locked buffers) on the CPUs next to its GPU: with one process per GPU on a
    multi-socket box, host<->device copies from the far socket cross the inter-socket link.  The CPU list comes from
    sysfs (/sys/bus/pci/devices/<bdf>/local_cpulist); nothing is changed when that is unavailable or names every CPU."""
    try:
        p = torch.cuda.get_device_properties(device)
        bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        text = open("/sys/bus/pci/devices/%s/local_cpulist" % bdf).read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        allowed = set(os.sched_getaffinity(0))
        cpus &= allowed
        node = open("/sys/bus/pci/devices/%s/numa_node" % bdf).read().strip()
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
        return {"pci": bdf, "numa_node": int(node), "cpus": len(cpus), "bound": bool(cpus and cpus != allowed)}
    except Exception as e:  # noqa: BLE001
        return {"numa_node": None, "error": str(e)[:80]}


def measure_host_link(torch, dist, world, nbytes):
    """What the host link of this box delivers: pinned->device and device->pinned copies of nbytes (CUDA events, best of
    5), first every rank alone, then ALL ranks copying at once (barrier-aligned): the aggregate is the floor the
    end-to-end curve has to be judged against when N ranks share one host."""
    host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")

    def one(kind):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if kind == "h2d":
            dev.copy_(host, non_blocking=True)
        else:
            host.copy_(dev, non_blocking=True)
        b.record()
        b.synchronize()
        return a.elapsed_time(b)

    out = {"bytes": nbytes}
    rank = dist.get_rank() if world > 1 else 0
    alone = {"h2d": 1e9, "d2h": 1e9}
    for r in range(world):                       # one rank at a time
        if world > 1:
            dist.barrier()
        if r == rank:
            for _ in range(5):
                for kind in alone:
                    alone[kind] = min(alone[kind], one(kind))
    out["h2d_gbs"] = nbytes / alone["h2d"] / 1e6
    out["d2h_gbs"] = nbytes / alone["d2h"] / 1e6
    if world > 1:
        worst = {"h2d": 1e9, "d2h": 1e9}
        for kind in worst:
            for _ in range(4):
                dist.barrier()
                torch.cuda.synchronize()
                t = torch.tensor([one(kind)], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                worst[kind] = min(worst[kind], float(t.item()))
        out["aggregate_h2d_gbs"] = world * nbytes / worst["h2d"] / 1e6
        out["aggregate_d2h_gbs"] = world * nbytes / worst["d2h"] / 1e6
    else:
        out["aggregate_h2d_gbs"], out["aggregate_d2h_gbs"] = out["h2d_gbs"], out["d2h_gbs"]
    return out


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
    ap.add_argument("--workload", default="wgs", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-file", action="store_true", help="skip the e2e_file (drop-in binary) measurement")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import numpy as np
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        sys.stderr.write("bench.py: no CUDA device -- the hot path has no CPU implementation\n")
        return 2
    torch.cuda.set_device(local)
    binding = bind_to_gpu_cpus(torch, local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    capi = importlib.import_module("16gt_b200.capi")
    sharding = importlib.import_module("16gt_b200.sharding")
    L = capi.capi()
    cfg = WORKLOADS[args.workload]
    idx, out, region, synth_s = ensure_workload(args.workload, rank)
    caller = capi.Caller(local)
    pre_ms, dec_ms = ctypes.c_float(), ctypes.c_float()
    L.gt16_setup_times(caller.ctx, ctypes.byref(pre_ms), ctypes.byref(dec_ms))

    # ---- this rank's shard of the ONE genome: contiguous, batch-aligned range of the processed-site sequence
    t_load = time.time()
    meta = capi.LoadedInput(idx, None, region, pinned=1)                 # index (+ region list) only
    total = meta.total_sites
    begin, end = sharding.my_shard(total, rank, world)
    shard = end - begin
    regs = meta.regions.reshape(-1, 2).astype(np.int64) if meta.regions is not None else None
    if regs is None:
        pos_lo, pos_hi = max(0, begin - 8192), min(meta.length, end + 8192)
    else:                                                                # exome: the genome span of the shard's regions
        first = np.concatenate([[0], np.cumsum(np.maximum(regs[:, 1] - regs[:, 0], 0))])
        r0 = int(np.searchsorted(first, begin, side="right") - 1)
        r1 = int(np.searchsorted(first, max(begin, min(total, end + 512) - 1), side="right") - 1)
        r0, r1 = min(r0, len(regs) - 1), min(r1, len(regs) - 1)
        pos_lo, pos_hi = max(0, int(regs[r0, 0]) - 8192), min(meta.length, int(regs[:r1 + 1, 1].max()) + 8192)
    # (a sparse -- exome -- snapshot stays the list of records it is: 12 B per covered position travel, not 8 B per genome position)
    inp = capi.LoadedInput(idx, out + ".snapshot", region, pinned=1, pos_range=(pos_lo, pos_hi), index=meta, keep_sparse=True)
    dev = ctypes.c_void_p()
    if shard:
        caller.check(L.gt16_snapshot_upload_range(caller.ctx, ctypes.byref(inp.view), inp.regions_ptr(), inp.n_regions, begin, end,
                                                  ctypes.byref(dev)))
    load_s = time.time() - t_load

    # ---- the rank-ordered host buffer the calls are gathered into (e2e): one slab per rank
    rec_bytes = ctypes.sizeof(capi.CalledSite)
    slab_recs = shard // 400 + 65536
    gather_note = "single process: the records are written in order into one page-locked buffer"
    shm_path, shm_map = None, None
    if world == 1:
        slab = torch.empty(slab_recs * rec_bytes, dtype=torch.uint8, pin_memory=True)
        slab_ptr = slab.data_ptr()
    else:
        # every rank maps the same /dev/shm file and page-locks ITS slab: the library's device->host copies land in
        # rank order in memory that rank 0 reads -- the final ordered gather needs no further copy, only the counts
        slab_all = torch.tensor([slab_recs], dtype=torch.int64, device="cuda")
        dist.all_reduce(slab_all, op=dist.ReduceOp.MAX)
        slab_recs = int(slab_all.item())
        slab_bytes = (slab_recs * rec_bytes + 4095) & ~4095
        shm_path = "/dev/shm/gt16_gather_%s" % os.environ.get("MASTER_PORT", "0")
        if rank == 0:
            with open(shm_path, "wb") as f:
                f.truncate(world * slab_bytes)
        dist.barrier()
        fd = os.open(shm_path, os.O_RDWR)
        shm_map = mmap.mmap(fd, world * slab_bytes)
        os.close(fd)
        base = ctypes.addressof(ctypes.c_char.from_buffer(shm_map))
        slab_ptr = base + rank * slab_bytes
        rc = torch.cuda.cudart().cudaHostRegister(slab_ptr, slab_bytes, 0)
        ok = int(rc[0] if isinstance(rc, tuple) else rc) == 0
        gather_note = ("shared /dev/shm buffer, one slab per rank in rank order, %s; counts all-gathered over NCCL"
                       % ("page-locked in place (cudaHostRegister): device->host copies land in it directly" if ok else
                          "NOT page-locked (cudaHostRegister failed): copies are staged by the driver"))

    pending, completed = [], []          # calls in flight; final stats of the completed ones, in call order

    def retire(count):
        """gt16_dev_calls_wait + free for the `count` oldest calls in flight (record count, error word, CUDA-event timings)"""
        for calls in pending[:count]:
            st = capi.Stats()
            caller.check(L.gt16_dev_calls_wait(caller.ctx, calls, ctypes.byref(st)))
            L.gt16_dev_calls_free(calls)
            completed.append(st)
        del pending[:count]

    def step_resident():
        # inputs resident in HBM, records left in HBM: the device-side pipeline.  gt16_call_sites_device_nowait only
        # enqueues (no host round trip anywhere in the call); a step is completed (gt16_dev_calls_wait) after the timed
        # region's closing synchronize -- or, in a long run, once it is 48 steps old (the library keeps the events of
        # the last 128 calls)
        if not shard:
            return
        if len(pending) >= 96:
            retire(48)
        calls, st = ctypes.c_void_p(), capi.Stats()
        caller.check(L.gt16_call_sites_device_nowait(caller.ctx, dev, inp.regions_ptr(), inp.n_regions, begin, end,
                                                     ctypes.byref(calls), ctypes.byref(st)))
        pending.append(calls)

    def complete_resident(k):
        retire(len(pending))
        done = completed[-k:] if k else []
        del completed[:]
        return done

    def step_resident_host_results():
        # same, but the records are also copied to page-locked host memory inside the step (gt16_call_sites)
        if not shard:
            return
        calls, st = ctypes.c_void_p(), capi.Stats()
        caller.check(L.gt16_call_sites(caller.ctx, dev, inp.regions_ptr(), inp.n_regions, begin, end, ctypes.byref(calls), ctypes.byref(st)))
        L.gt16_calls_free(calls)

    counts_dev = torch.zeros(world, dtype=torch.int64, device="cuda")
    mine = torch.zeros(1, dtype=torch.int64, device="cuda")
    e2e_called = [0]

    def step_host():
        n = ctypes.c_uint64(0)
        st = capi.Stats()
        if shard:
            caller.check(L.gt16_call_sites_host_to(caller.ctx, ctypes.byref(inp.view), inp.regions_ptr(), inp.n_regions, begin, end,
                                                   slab_ptr, slab_recs, ctypes.byref(n), ctypes.byref(st)))
        if world > 1:                                # the ordered gather: records are in the shared buffer, counts to everybody
            mine[0] = n.value
            dist.all_gather_into_tensor(counts_dev, mine)
            e2e_called[0] = int(counts_dev.sum().item())
        else:
            e2e_called[0] = n.value

    def timed(fn, warmup, steps):
        for _ in range(warmup):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            fn()
        barrier()
        dt = time.perf_counter() - t0
        return sharding.max_over_ranks(dt, dist if world > 1 else None, "cuda")

    # clocks / throttle reasons are sampled over the whole measurement phase
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    for _ in range(10):                # keep the GPU busy until the sampler has certainly started
        step_resident()
    complete_resident(0)
    dt = timed(step_resident, args.warmup, args.steps)
    stats = complete_resident(args.steps)
    n_hr = max(2, min(args.steps, 3))
    dt_hr = timed(step_resident_host_results, 1, n_hr)
    n_e2e = max(2, args.steps)
    dt_e2e = timed(step_host, 2, n_e2e)
    clocks = sampler.stop() if sampler else None
    link = measure_host_link(torch, dist, world, 1 << 29)

    # ---- e2e_file: the drop-in binary, snapshot FILE -> .txt, every GPU of the job (the other ranks idle at the barrier)
    e2e_file = None
    if not args.no_file and os.path.exists(CLI):
        barrier()
        if rank == 0:
            runs = []
            prefix = out                                  # <prefix>.snapshot is the cached workload file
            env = dict(os.environ, GT16_DEVICES=str(world))
            for _ in range(2):
                t0 = time.perf_counter()
                r = subprocess.run([CLI, "-i", idx, "-o", prefix, "-v"] + (["-e", region] if region else []),
                                   stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env)
                wall = time.perf_counter() - t0
                m = re.search(r"Called Entries : (\d+)", r.stdout)
                runs.append({"wall_s": wall, "rc": r.returncode, "called": int(m.group(1)) if m else None})
            txt_bytes = os.path.getsize(prefix + ".txt") if os.path.exists(prefix + ".txt") else 0
            if os.path.exists(prefix + ".txt"):
                os.remove(prefix + ".txt")
            good = [x for x in runs if x["rc"] == 0]
            if good:
                best = min(x["wall_s"] for x in good)
                e2e_file = {"value": total / best, "unit": UNIT, "wall_s": best, "runs": runs, "txt_bytes": txt_bytes,
                            "what": "16gt_b200/bin/snapshotSnpcaller -i <index> -o <prefix> on the %.1f GB snapshot file (page cache / "
                                    "tmpfs), %d GPU(s): process start to exit, .txt written (mapped counters -> page-locked ring "
                                    "-> HBM, rows formatted beside the GPU work)" % (os.path.getsize(out + ".snapshot") / 1e9, world)}
            else:
                e2e_file = {"error": (runs[-1] if runs else None)}
        barrier()

    called_rank = stats[0].called if stats else 0
    flagged_rank = stats[0].flagged if stats else 0
    tot = torch.tensor([called_rank, flagged_rank, sum(s.kernel_launches for s in stats)], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(tot)
    called, flagged, launches = (int(x) for x in tot.tolist())
    parts = {k: (sum(getattr(s, k) for s in stats) / len(stats) if stats else 0.0)
             for k in ("ms_stream", "ms_scan", "ms_eval", "ms_candidates", "ms_d2h", "ms_total_device")}
    ms_stream = parts["ms_stream"]
    v = inp.view
    ovf_records = 0
    if v.n_overflow_regions:
        arr = (ctypes.c_uint32 * v.n_overflow_regions).from_address(v.overflow_size)
        ovf_records = sum(arr)
    # bytes this rank moves per e2e step (counters of its shard + the side arrays its positions reach), summed over ranks
    if v.counters:
        h2d_rank = shard * 8 + shard // 8 + shard // 4 + (ovf_records * 32) // world + v.pool_bytes
    else:                                # sparse view: the records whose positions lie in this rank's window, 12 bytes each
        k = v.n_sparse_entries
        raw = np.ctypeslib.as_array(ctypes.cast(v.sparse_entries, ctypes.POINTER(ctypes.c_uint8)), (k * 12,)).reshape(k, 12)
        pos = raw[:, :4].copy().view(np.uint32)[:, 0]
        in_window = int(np.searchsorted(pos, pos_hi) - np.searchsorted(pos, pos_lo)) if shard else 0
        # ... and the packed reference of the genome span its targets lie in (2 bits per position of the span, not per site)
        h2d_rank = in_window * 12 + (pos_hi - pos_lo) // 4 + (ovf_records * 32) // world + v.pool_bytes
    h2d_t = torch.tensor([h2d_rank], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(h2d_t)
    h2d = int(h2d_t.item())
    d2h = e2e_called[0] * rec_bytes + 8 * world

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = shard * BYTES_PER_SITE / (ms_stream * 1e-3) / 1e9 if ms_stream else 0.0

    floor_ms = h2d / (link["aggregate_h2d_gbs"] * 1e6)
    line = {
        "metric": METRIC, "value": total * args.steps / dt, "unit": UNIT, "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * dt / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": cfg["name"],
                   "sites_per_step": total, "sites_per_step_rank0": shard, "called_entries": called,
                   "exactly_evaluated_sites": flagged,
                   "l2": "every rank's inputs (%.1f GB of counters per pass) exceed the 126 MB L2; no flush needed" % (shard * 8 / 1e9),
                   "parallelism": "ONE genome partitioned into %d contiguous batch-aligned shards (sharding.my_shard), no collective on "
                                  "the data path; final ordered gather of the calls: %s" % (world, gather_note),
                   "setup_s": {"synthesise_or_wait": synth_s, "read_pin_upload_shard": load_s,
                               "table_prefill_ms": pre_ms.value, "decision_bitmaps_ms": dec_ms.value},
                   "cpu_binding": binding},
        "e2e": {"value": total * n_e2e / dt_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1000.0 * dt_e2e / n_e2e, "called_entries": e2e_called[0],
                "host_link": dict(link, h2d_floor_ms=floor_ms, frac_of_floor=floor_ms / (1000.0 * dt_e2e / n_e2e),
                                  note="plain cudaMemcpyAsync from page-locked memory on this box (CUDA events, best of 5), one rank "
                                       "alone and all ranks at once: the time the step's input bytes need on the bus whatever the "
                                       "kernels do, at the aggregate rate the box delivers to N concurrent ranks")},
        "e2e_file": e2e_file,
        "value_with_host_results": {"value": total * n_hr / dt_hr, "unit": UNIT, "ms_per_step": 1000.0 * dt_hr / n_hr,
                                    "note": "value = inputs AND called-site records resident in HBM; this line adds the copy of the "
                                            "records to page-locked host memory (gt16_call_sites)"},
        "gpu_launches": launches,
        "device_ms_per_step_rank0": parts["ms_total_device"],
        "device_ms_breakdown_rank0": {k: parts[k] for k in ("ms_stream", "ms_eval", "ms_scan", "ms_candidates", "ms_d2h")},
        "roofline": {"bound": "hbm", "kernel": ("k_stream_runs (decode + 16-genotype classification of every target site; the targets' runs of positions staged through a TMA ring), rank 0's launches"
                                                if region else "k_stream_tma (decode + 16-genotype classification of every site), rank 0's launch"),
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": shard * BYTES_PER_SITE, "avg_launch_ms": ms_stream},
        "clocks": clocks,
    }
    tr = os.path.join(ROOT, "profiles", "stream_kernel_traffic.json")
    if os.path.exists(tr) and not region:
        try:
            t = json.load(open(tr))
            # measured by ncu on one launch of `sites_per_launch` sites: bytes per site carry over to a launch of any length
            line["roofline"]["traffic"] = t.get("dram_bytes_per_launch") / t.get("sites_per_launch", CHR20_SITES) * shard
        except (ValueError, OSError, TypeError):
            pass

    if rank == 0 and world == 1 and not args.no_cpu_baseline and os.path.exists(REF_CALLER):
        th = ref_threads()
        sidx, sout, _, _ = ensure_workload("chr20", 0)
        sel, pre, ccalled = run_reference_once(sidx, sout, th, None)
        line["cpu_baseline"] = {
            "value": CHR20_SITES / sel, "unit": UNIT, "cores": th, "kind": "reference",
            "sample": "oracle/_ref/ref_caller (unmodified reference objects) on one %d-site chr20-sized contig of the same "
                      "synthetic 30x workload (configs[1]; the whole genome takes it 25 minutes): selectPossibleSNPs %.2f s (serial "
                      "decode loop + OpenMP operators on %d threads + consumer thread), %d called; of that, the operators alone (the "
                      "reference's own 'Total Likelihood and Bias computing time'): %s s; one-off table prefill %.1f s not included"
                      % (CHR20_SITES, sel, th, ccalled,
                         "%.2f" % run_reference_once.operators if run_reference_once.operators is not None else "n/a", pre),
            "select_seconds": sel, "operator_seconds": run_reference_once.operators, "prefill_seconds": pre}
    elif rank == 0 and world == 1:
        line["cpu_baseline"] = None

    if dev:
        L.gt16_snapshot_release(caller.ctx, dev)
    inp.close()
    meta.close()
    caller.close()
    if shm_map is not None:
        try:
            torch.cuda.cudart().cudaHostUnregister(slab_ptr)
        except Exception:  # noqa: BLE001
            pass
    if world > 1:
        dist.barrier()
        if rank == 0 and shm_path:
            try:
                os.remove(shm_path)
            except OSError:
                pass
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
