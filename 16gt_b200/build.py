"""Build everything in-tree for sm_100a: libgt16.so (CUDA kernels + C ABI), the
snapshotSnpcaller drop-in binary, the synthetic-data generator and the test-only oracle.

    python 16gt_b200/build.py [--force]

nvcc cross-compiles without a GPU.  Outputs (git-ignored, shipped to the GPU box by gpurun):
    16gt_b200/lib/libgt16.so          16gt_b200/bin/snapshotSnpcaller (+ .ini)
    tools/gt16_synth                  oracle/libgt16_oracle.so, oracle/_ref/* (when /root/reference exists)
"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "16gt_b200")
CSRC = os.path.join(PKG, "csrc")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
HOST_CXX = "/usr/bin/g++"   # the image's $CXX points at a g++ without libgomp

# -fmad=false: the reference is built for baseline x86-64 (no FMA); contraction would change
# the truncations in the likelihood arithmetic (see gt16_core.cuh).
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-ccbin", HOST_CXX, "-Xcompiler", "-fPIC,-O2,-Wall,-Wno-unused-function", "-Xptxas", "-v",
]


def _newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def _run(cmd, log=None):
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if log:
        with open(log, "w") as f:
            f.write(" ".join(cmd) + "\n" + r.stdout)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("build step failed: " + " ".join(cmd))
    return r.stdout


def build(force=False, verbose=True):
    os.makedirs(os.path.join(PKG, "lib"), exist_ok=True)
    os.makedirs(os.path.join(PKG, "bin"), exist_ok=True)
    os.makedirs(os.path.join(PKG, "build"), exist_ok=True)
    hdrs = [os.path.join(CSRC, h) for h in ("gt16_core.cuh", "gt16_kernels.cuh", "gt16_pileup.cuh")] + [os.path.join(ROOT, "include", "gt16.h")]
    lib = os.path.join(PKG, "lib", "libgt16.so")
    units = ["gt16_kernels.cu", "gt16_capi.cu", "gt16_pileup.cu", "gt16_host_io.cpp", "gt16_sam.cpp"]
    objs = []
    for u in units:
        src = os.path.join(CSRC, u)
        obj = os.path.join(PKG, "build", u.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        if force or not _newer(obj, [src] + hdrs):
            if verbose:
                print("nvcc", u, flush=True)
            _run([NVCC] + NVCC_FLAGS + ["-x", "cu", "-c", src, "-o", obj], log=obj + ".log")
    if force or not _newer(lib, objs):
        _run([NVCC, "-shared", "-ccbin", HOST_CXX, "-gencode", "arch=compute_100a,code=sm_100a"] + objs +
             ["-o", lib, "-cudart", "static", "-lpthread", "-ldl", "-lrt", "-lz"])
    # drop-in CLI
    cli_src = os.path.join(PKG, "host", "snapshotSnpcaller.cpp")
    cli = os.path.join(PKG, "bin", "snapshotSnpcaller")
    if os.path.exists(cli_src) and (force or not _newer(cli, [cli_src, lib] + hdrs)):
        _run([HOST_CXX, "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), cli_src, "-o", cli,
              "-L", os.path.join(PKG, "lib"), "-lgt16", "-Wl,-rpath,$ORIGIN/../lib",
              "-Wl,-rpath-link,/usr/local/cuda/lib64", "-lpthread"])
        shutil.copyfile(os.path.join(PKG, "host", "snapshotSnpcaller.ini"), cli + ".ini")
    rib_src = os.path.join(PKG, "host", "RegionIndexBuilder.cpp")
    rib = os.path.join(PKG, "bin", "RegionIndexBuilder")
    if os.path.exists(rib_src) and (force or not _newer(rib, [rib_src, lib] + hdrs)):
        _run([HOST_CXX, "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), rib_src, "-o", rib,
              "-L", os.path.join(PKG, "lib"), "-lgt16", "-Wl,-rpath,$ORIGIN/../lib",
              "-Wl,-rpath-link,/usr/local/cuda/lib64", "-lpthread"])
    b2s_src = os.path.join(PKG, "host", "bam2snapshot.cpp")
    b2s = os.path.join(PKG, "bin", "bam2snapshot")
    if os.path.exists(b2s_src) and (force or not _newer(b2s, [b2s_src, lib] + hdrs)):
        _run([HOST_CXX, "-O2", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), b2s_src, "-o", b2s,
              "-L", os.path.join(PKG, "lib"), "-lgt16", "-Wl,-rpath,$ORIGIN/../lib",
              "-Wl,-rpath-link,/usr/local/cuda/lib64", "-lpthread"])
        shutil.copyfile(os.path.join(PKG, "host", "bam2snapshot.ini"), b2s + ".ini")
    t2v_src = os.path.join(PKG, "host", "txt2vcf.cpp")
    t2v = os.path.join(PKG, "bin", "txt2vcf")
    if os.path.exists(t2v_src) and (force or not _newer(t2v, [t2v_src])):
        _run([HOST_CXX, "-O2", "-std=c++17", "-Wall", t2v_src, "-o", t2v, "-lpthread"])
    flt_src = os.path.join(PKG, "host", "filterVCF.cpp")
    flt = os.path.join(PKG, "bin", "filterVCF")
    if os.path.exists(flt_src) and (force or not _newer(flt, [flt_src])):
        _run([HOST_CXX, "-O2", "-std=c++17", "-Wall", flt_src, "-o", flt])
    # synthetic data generator
    synth_src = os.path.join(ROOT, "tools", "gt16_synth.cpp")
    synth = os.path.join(ROOT, "tools", "gt16_synth")
    if force or not _newer(synth, [synth_src]):
        _run([HOST_CXX, "-O2", "-std=c++17", "-Wall", "-pthread", synth_src, "-o", synth])
    rs_src = os.path.join(ROOT, "tools", "gt16_readsim.cpp")
    rs = os.path.join(ROOT, "tools", "gt16_readsim")
    if force or not _newer(rs, [rs_src]):
        _run([HOST_CXX, "-O2", "-std=c++17", "-Wall", rs_src, "-o", rs])
    # test-only oracle (C restatement) and the reference objects
    mk = os.path.join(ROOT, "oracle", "Makefile")
    if os.path.exists(mk):
        _run(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    if os.path.isdir(os.environ.get("REF", "/root/reference")):
        _run(["bash", os.path.join(ROOT, "oracle", "build_ref.sh")])
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
