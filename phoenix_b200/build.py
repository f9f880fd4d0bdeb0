"""Build libphoenix_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m phoenix_b200.build [--force] [--verbose]

Objects are compiled in parallel into phoenix_b200/csrc/_obj/ (the resident solver kernels are one translation unit
per (forward|adjoint, NV) instantiation) and linked into phoenix_b200/libphoenix_b200.so.  The .so is git-ignored but
travels to the GPU box with the gpurun snapshot.
"""
import argparse
import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_obj")
LIB = os.path.join(HERE, "libphoenix_b200.so")
HEADERS = [os.path.join(CSRC, "phx_common.cuh"), os.path.join(CSRC, "phx_resident.cuh"), os.path.join(CSRC, "phx_tc.cuh"),
           os.path.join(CSRC, "phx_rows.cuh"),
           os.path.join(HERE, "..", "include", "phoenix_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-fmad=false",  # elementwise solver math rounds op-by-op like ATen; dot products use explicit fmaf
    "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall",
]

# (object name, source, extra defines)
UNITS = [("phx_api", "phx_api.cu", []), ("phx_plan", "phx_plan.cu", []), ("phx_rhs", "phx_rhs.cu", []),
         ("phx_stream", "phx_stream.cu", []), ("phx_tc", "phx_tc.cu", []), ("phx_peer", "phx_peer.cu", [])]
for kind in (0, 1):
    UNITS.append(("phx_rows_%s" % ("adj" if kind else "fwd"), "phx_rows_inst.cu", ["-DPHX_KIND_ADJ=%d" % kind]))
    for nv in (1, 2, 4):
        UNITS.append(("phx_res_%s_nv%d" % ("adj" if kind else "fwd", nv), "phx_resident_inst.cu",
                      ["-DPHX_KIND_ADJ=%d" % kind, "-DPHX_NV=%d" % nv]))


def _nvcc():
    return os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(unit, verbose):
    name, src, defs = unit
    if os.environ.get("PHX_TC_PROFILE_BUILD") and name == "phx_tc":   # in-kernel cycle counters of the tcgen05 kernels
        defs = defs + ["-DPHX_TC_PROFILE=1"]
    obj = os.path.join(OBJ, name + ".o")
    cmd = [_nvcc()] + NVCC_FLAGS + defs + (["-Xptxas", "-v"] if verbose else []) + \
          ["-c", os.path.join(CSRC, src), "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return name, res.returncode, res.stdout + res.stderr


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    me = os.path.abspath(__file__)
    todo = [u for u in UNITS
            if force or _stale(os.path.join(OBJ, u[0] + ".o"), [os.path.join(CSRC, u[1]), me] + HEADERS)]
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(len(todo), os.cpu_count() or 4)) as ex:
            for name, rc, out in ex.map(lambda u: _compile(u, verbose), todo):
                if verbose or rc != 0:
                    sys.stderr.write("==== %s ====\n%s" % (name, out))
                if rc != 0:
                    raise RuntimeError("nvcc failed on %s" % name)
    objs = [os.path.join(OBJ, u[0] + ".o") for u in UNITS]
    if todo or _stale(LIB, objs):
        cmd = [_nvcc(), "--shared", "-cudart", "static", "-o", LIB] + objs
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link of libphoenix_b200.so failed")
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
