"""Builds libfvgn_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python finite-volume-graph-network_b200/build.py [--force]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.  Objects are cached under
csrc/_build/ keyed by source mtime."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "_build")
LIB = os.path.join(HERE, "libfvgn_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "--expt-relaxed-constexpr", "-Xptxas", "-v"] + os.environ.get("FVGN_NVCC_EXTRA", "").split()  # e.g. -DFVGN_TC_TIMING_EXPERIMENTS=1
# per-file extras: the geometry / pre-post kernels must not contract a*b+c into FMA (bit-parity with numpy/torch)
SOURCES = {
    "api.cu": [],
    "mlp_simt.cu": [],
    "mlp_tc.cu": [],
    "mlp_tc3.cu": ["--fmad=false"],
    "mlp_tc3_bwd.cu": ["--fmad=false"],
    "mlp_bwd.cu": [],
    "connectivity.cu": ["--fmad=false"],
    "prepost.cu": ["--fmad=false"],
    "executor.cu": [],
    "halo.cu": [],
}


def _deps():
    return [os.path.join(CSRC, "common.cuh"), os.path.join(HERE, "..", "include", "fvgn_b200.h")] + \
        [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh")]


def _stale(src, obj):
    if not os.path.exists(obj):
        return True
    t = os.path.getmtime(obj)
    return any(os.path.getmtime(d) > t for d in [src] + _deps())


def _compile(name, extra, force, verbose):
    src = os.path.join(CSRC, name)
    obj = os.path.join(OBJ, name.replace(".cu", ".o"))
    if not force and not _stale(src, obj):
        return name, "cached", ""
    cmd = [NVCC] + ARCH + COMMON + extra + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {name}:\n{r.stdout}\n{r.stderr}")
    return name, "built", r.stderr if verbose else ""


def build_variant(name, defines, only=("mlp_tc3.cu",), verbose=False):
    """experimental build `libfvgn_b200_<name>.so` (git-ignored; selected at run time with FVGN_B200_LIB=<path>): the sources in `only`
    recompiled with extra -D flags into csrc/_build_<name>/, every other object shared with the product build"""
    build()
    vobj = os.path.join(CSRC, "_build_" + name)
    os.makedirs(vobj, exist_ok=True)
    objs, logs = [], []
    for n, extra in SOURCES.items():
        if n in only:
            o = os.path.join(vobj, n.replace(".cu", ".o"))
            cmd = [NVCC] + ARCH + COMMON + extra + list(defines) + ["-c", os.path.join(CSRC, n), "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed for {n} [{name}]:\n{r.stdout}\n{r.stderr}")
            logs.append(r.stderr)
            objs.append(o)
        else:
            objs.append(os.path.join(OBJ, n.replace(".cu", ".o")))
    lib = os.path.join(HERE, f"libfvgn_b200_{name}.so")
    r = subprocess.run([NVCC] + ARCH + ["-shared", "-o", lib] + objs + ["-Xcompiler", "-fPIC", "-lcudart"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return lib, "\n".join(logs)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        results = list(ex.map(lambda kv: _compile(kv[0], kv[1], force, verbose), SOURCES.items()))
    objs = [os.path.join(OBJ, n.replace(".cu", ".o")) for n in SOURCES]
    need_link = force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs)
    if need_link:
        cmd = [NVCC] + ARCH + ["-shared", "-o", LIB] + objs + ["-Xcompiler", "-fPIC", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB, results


if __name__ == "__main__" and "--variant" in sys.argv:
    i = sys.argv.index("--variant")
    lib, log = build_variant(sys.argv[i + 1], [a for a in sys.argv[i + 2:] if a.startswith("-D")])
    print(log if "-v" in sys.argv else "", lib)
    sys.exit(0)

if __name__ == "__main__":
    lib, res = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    for n, s, log in res:
        print(n, s)
        if log:
            print(log)
    print(lib)
