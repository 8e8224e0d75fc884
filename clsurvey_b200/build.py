"""In-tree build of libclb.so (hand-written CUDA for sm_100a behind the C ABI of include/clb.h).

    python -m clsurvey_b200.build [--force]

nvcc cross-compiles without a GPU; the resulting clsurvey_b200/_lib/libclb.so is git-ignored but travels to the
GPU box with the gpurun snapshot.  No torch headers are involved: the library is plain CUDA runtime + dlopen'd NCCL.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_lib")
LIB = os.path.join(OUT_DIR, "libclb.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)) + ["../../include/clb.h"]:
        p = os.path.join(CSRC, f)
        if os.path.isfile(p):
            h.update(f.encode())
            h.update(open(p, "rb").read())
    h.update(" ".join(ARCH + FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False, extra_flags=(), out_name=None):
    """extra_flags / out_name: experimental variants (e.g. -DCLB_TC2_GROUPS=4 -> libclb_g4.so, loaded via CLB_LIB_PATH)."""
    os.makedirs(OUT_DIR, exist_ok=True)
    if out_name:
        return _build_variant(list(extra_flags), out_name)
    stamp = os.path.join(OUT_DIR, "build.stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    nvcc = _nvcc()
    objs = []

    def compile_one(src):
        obj = os.path.join(OUT_DIR, os.path.basename(src)[:-3] + ".o")
        cmd = [nvcc] + ARCH + FLAGS + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
        if verbose and r.stderr.strip():
            print(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
    open(stamp, "w").write(dig)
    return LIB


def _build_variant(extra, out_name):
    nvcc = _nvcc()
    vdir = os.path.join(OUT_DIR, "variant_" + out_name)
    os.makedirs(vdir, exist_ok=True)
    objs = []
    for src in sources():
        obj = os.path.join(vdir, os.path.basename(src)[:-3] + ".o")
        r = subprocess.run([nvcc] + ARCH + FLAGS + extra + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(r.stderr)
        objs.append(obj)
    lib = os.path.join(OUT_DIR, out_name)
    r = subprocess.run([nvcc] + ARCH + ["-shared", "-o", lib] + objs + ["-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    return lib


def build_probe():
    """Bring-up probes (csrc/probe/clb_debug.cu: MN-major descriptor, bf16 operand conventions, TMA box behaviour) as their own
    library _lib/libclb_probe.so -- test tooling for tools/*_probe.py, not part of the product ABI (include/clb.h)."""
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = _nvcc()
    lib = os.path.join(OUT_DIR, "libclb_probe.so")
    srcs = [os.path.join(CSRC, "probe", "clb_debug.cu"), os.path.join(CSRC, "clb_tma.cu"), os.path.join(CSRC, "clb_core.cu")]
    r = subprocess.run([nvcc] + ARCH + FLAGS + ["-shared", "-o", lib] + srcs + ["-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(r.stderr)
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
