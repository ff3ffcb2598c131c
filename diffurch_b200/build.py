"""Build libdiffurch_b200.so in-tree with nvcc for sm_100a.

Every kernel translation unit is compiled twice — STRICT (-fmad=false, no FMA
contraction: the reference's rounding) and FAST — and the general integrator is
split in groups so the objects build in parallel.  nvcc cross-compiles without a
GPU, so this runs on the CPU-only build box too.
"""
import concurrent.futures as cf
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_obj")
LIB = os.path.join(HERE, "libdiffurch_b200.so")
NVCC = os.environ.get("NVCC", "nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden"]
N_GROUPS = 5


def _jobs():
    jobs = []
    for arith, tag, extra in ((0, "strict", ["-fmad=false"]), (1, "fast", [])):
        jobs.append(("kernels_ode_fixed.cu", f"ode_fixed_{tag}.o", [f"-DDFB_ARITH={arith}"] + extra))
        jobs.append(("kernels_ode_event.cu", f"ode_event_{tag}.o", [f"-DDFB_ARITH={arith}"] + extra))
        for g in range(1, N_GROUPS + 1):
            jobs.append(("kernels_general.cu", f"general_g{g}_{tag}.o",
                         [f"-DDFB_ARITH={arith}", f"-DDFB_GROUP={g}"] + extra))
    jobs.append(("kernels_dde_fixed.cu", "dde_fixed_fast.o", ["-DDFB_ARITH=1"]))
    jobs.append(("kernels_dde_event.cu", "dde_event_fast.o", ["-DDFB_ARITH=1"]))
    jobs.append(("kernels_ode_adaptive.cu", "ode_adaptive_fast.o", ["-DDFB_ARITH=1"]))
    jobs.append(("nvrtc_systems.cu", "nvrtc_systems.o", []))
    jobs.append(("api.cu", "api.o", []))
    return jobs


def _deps_mtime():
    m = 0.0
    for root, _, files in os.walk(CSRC):
        for f in files:
            m = max(m, os.path.getmtime(os.path.join(root, f)))
    m = max(m, os.path.getmtime(os.path.join(HERE, "..", "include", "diffurch_b200.h")))
    m = max(m, os.path.getmtime(os.path.abspath(__file__)))
    return m


def _compile(job):
    src, obj, flags = job
    out = os.path.join(OBJ, obj)
    cmd = [NVCC] + ARCH + COMMON + flags + ["-c", os.path.join(CSRC, src), "-o", out]
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    return obj, r.returncode, r.stdout + r.stderr, time.time() - t0, cmd


def build(force=False, verbose=True, workers=None):
    os.makedirs(OBJ, exist_ok=True)
    dep = _deps_mtime()
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= dep:
        if verbose:
            print(f"[diffurch_b200.build] up to date: {LIB}")
        return LIB
    jobs = _jobs()
    todo = [j for j in jobs
            if force or not os.path.exists(os.path.join(OBJ, j[1]))
            or os.path.getmtime(os.path.join(OBJ, j[1])) < dep]
    workers = workers or min(len(todo) or 1, os.cpu_count() or 4)
    t0 = time.time()
    with cf.ThreadPoolExecutor(max_workers=workers) as ex:
        for obj, rc, log, dt, cmd in ex.map(_compile, todo):
            if verbose:
                print(f"[diffurch_b200.build] {obj}: {dt:.1f}s rc={rc}")
            if rc != 0:
                sys.stderr.write(" ".join(cmd) + "\n" + log + "\n")
                raise RuntimeError(f"nvcc failed on {obj}")
    objs = [os.path.join(OBJ, j[1]) for j in jobs]
    cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    if verbose:
        print(f"[diffurch_b200.build] linked {LIB} in {time.time() - t0:.1f}s")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv)
