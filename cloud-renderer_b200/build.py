"""Builds libcloud_renderer_b200.so (the C-ABI library) in-tree with nvcc for sm_100a.

The same sources and flags are described by the top-level CMakeLists.txt for C++ users;
this script exists so that `__graft_entry__.build()` and the tests do not depend on a
configure step.  k_prep_sort.cu and k_voxelize.cu carry the exact-parity arithmetic and
are compiled with -fmad=false (one IEEE rounding per written operation).
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libcloud_renderer_b200.so")
OBJ = os.path.join(HERE, "build")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC,-ffp-contract=off", "--expt-relaxed-constexpr"]
SOURCES = {
    "crn_api.cu": [],
    "k_prep_sort.cu": ["-fmad=false"],
    "k_bin.cu": [],
    "k_voxelize.cu": ["-fmad=false"],
    "k_mips.cu": [],
    "k_trace.cu": ["--use_fast_math"],      # image parity is PSNR-based; div/sqrt/rcp as MUFU approximations
    "k_skipmask.cu": [],
    "k_microbench.cu": [],
}


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA toolkit is required to build cloud-renderer_b200")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "cloud_renderer_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    host = ["-ccbin", "/usr/bin/g++"] if os.path.exists("/usr/bin/g++") else []
    objs, procs = [], []
    for src, extra in SOURCES.items():
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc] + host + ARCH + COMMON + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- {src}\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed")
    cmd = [nvcc] + host + ARCH + ["-shared", "-o", OUT] + objs
    subprocess.check_call(cmd)
    build_example()
    return OUT


def build_example():
    """examples/frame.cpp: the reference's frame loop against include/cloud_renderer_b200.hpp"""
    root = os.path.dirname(HERE)
    exe = os.path.join(OBJ, "crn_frame")
    gxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    subprocess.check_call([gxx, "-O2", "-std=c++17", "-I", os.path.join(root, "include"), os.path.join(root, "examples", "frame.cpp"),
                           "-o", exe, OUT, "-Wl,-rpath," + HERE])
    return exe


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
