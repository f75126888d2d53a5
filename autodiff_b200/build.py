"""Builds libadb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libadb200.so")
EW_NSHARDS = 4   # must match N_SHARDS in tools/gen_ewise_catalog.py
# (source, extra nvcc flags, object name)
SOURCES = [(f, [], f.replace(".cu", ".o")) for f in
           ["adb_core.cu", "gemm_f64.cu", "gemm_tf32x3.cu", "ewise.cu", "reduce.cu", "softmax_ce_ring.cu", "einsum.cu", "conv.cu"]]
SOURCES += [("ewise_catalog.cu", ["-DEW_SHARD=%d" % k], "ewise_catalog_%d.o" % k) for k in range(EW_NSHARDS)]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]


def _stale():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "adb200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src, extra, objname in SOURCES:
        obj = os.path.join(HERE, "build", objname)
        objs.append(obj)
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((objname, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, out))
        if verbose and out.strip():
            print(out)
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-cudart", "static", "-Xlinker", "--exclude-libs,ALL"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
