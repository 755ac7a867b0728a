import sys, time, ctypes, tempfile, subprocess, importlib
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
capi = importlib.import_module("16gt_b200.capi")
L = capi.capi()
wd = tempfile.mkdtemp()
subprocess.run(["/root/repo/tools/gt16_synth", "--index", wd+"/g", "--snapshot", wd+"/s", "--contigs", "chr20:64444167", "--seed", "1"], check=True, stderr=subprocess.DEVNULL)
c = capi.Caller(0)
for pinned in (1, 0, 0, 1):
    t0 = time.perf_counter()
    inp = capi.LoadedInput(wd+"/g", wd+"/s.snapshot", None, pinned=pinned)
    t1 = time.perf_counter()
    for rep in range(3):
        calls = ctypes.c_void_p(); st = capi.Stats()
        ta = time.perf_counter()
        c.check(L.gt16_call_sites_host(c.ctx, ctypes.byref(inp.view), None, 0, 0, inp.total_sites, ctypes.byref(calls), ctypes.byref(st)))
        tb = time.perf_counter()
        n = L.gt16_calls_count(calls); L.gt16_calls_free(calls)
        print("pinned", pinned, "read %.3f s" % (t1 - t0), "call %.1f ms" % ((tb - ta) * 1e3), "called", n, "device ms %.1f" % st.ms_total_device, flush=True)
    inp.close()
