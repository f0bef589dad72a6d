"""Builds libgsd_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB = os.path.join(PKG, "libgsd_b200.so")
OBJ = os.path.join(HERE, "_obj")
# debug variants (e.g. GSD_BUILD_DEFINES="-DGSD_RACECHECK_ARRIVE_ALL" GSD_BUILD_TAG=rc): separate objects and library name,
# loaded with GSD_LIB_PATH=<that .so>
if os.environ.get("GSD_BUILD_TAG"):
    LIB = os.path.join(PKG, "libgsd_b200_%s.so" % os.environ["GSD_BUILD_TAG"])
    OBJ = os.path.join(HERE, "_obj_%s" % os.environ["GSD_BUILD_TAG"])

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]
# per-file extra flags
EXTRA = {"raster_pre.cu": ["-fmad=false"]}


def sources():
    return sorted(f for f in os.listdir(HERE) if f.endswith(".cu"))


def _stale(out, deps):
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build_lib(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(PKG, "..", "include", "gsd.h"))
    srcs = sources()
    jobs = []
    for s in srcs:
        src = os.path.join(HERE, s)
        obj = os.path.join(OBJ, s[:-3] + ".o")
        if force or _stale(obj, [src] + headers):
            jobs.append((s, [nvcc] + ARCH + COMMON + EXTRA.get(s, []) + os.environ.get("GSD_BUILD_DEFINES", "").split() + ["-c", src, "-o", obj]))
    logs = {}

    def run(job):
        name, cmd = job
        r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        logs[name] = r.stdout
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (name, r.stdout))

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        list(ex.map(run, jobs))
    with open(os.path.join(OBJ, "ptxas.log"), "a" if not force else "w") as f:
        for k, v in logs.items():
            f.write("==== %s\n%s\n" % (k, v))
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in srcs]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc] + ARCH + ["-shared", "-o", LIB] + objs + ["-cudart", "static"]
        subprocess.check_call(cmd)
    if verbose:
        for k, v in logs.items():
            print("====", k)
            print(v)
    return LIB


if __name__ == "__main__":
    print(build_lib(force="--force" in sys.argv, verbose=True))
