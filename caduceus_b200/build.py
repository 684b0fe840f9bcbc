"""Build the sm_100a shared library `caduceus_b200/csrc/libcaduceus_b200.so` in-tree with nvcc.

    python -m caduceus_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the repo snapshot.  nvcc cross-compiles without a GPU.
"""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(CSRC, "libcaduceus_b200.so")
STAMP = os.path.join(CSRC, ".build_stamp")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo", "--use_fast_math", "-Xcompiler", "-fPIC", "-Xcompiler", "-O3",
    "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _digest():
    h = hashlib.sha256()
    for f in _sources() + sorted(glob.glob(os.path.join(CSRC, "*.cuh"))) + [
            os.path.join(os.path.dirname(HERE), "include", "caduceus_b200.h"), os.path.abspath(__file__)]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu into one shared library (one nvcc process per file, then a link)."""
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    objs, procs = [], []
    for src in _sources():
        obj = src[:-3] + ".o"
        objs.append(obj)
        cmd = [nvcc, *NVCC_FLAGS, "-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"==== {os.path.basename(src)}\n{out}")
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{out}")
    link = [nvcc, "-shared", "-o", LIB, *objs]  # cudart is linked statically (nvcc default)
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(os.path.join(CSRC, "ptxas.log"), "w") as fh:
        fh.write("\n".join(log))
    with open(STAMP, "w") as fh:
        fh.write(digest)
    if verbose:
        print("\n".join(log))
    return LIB


PROBE_SRC = os.path.join(os.path.dirname(HERE), "scripts", "hw_probe.cu")
PROBE_BIN = os.path.join(os.path.dirname(HERE), "scripts", "_bin", "hw_probe")


def build_probe(force=False):
    """Compile scripts/hw_probe.cu (the Python-free hardware probe of the scan kernels) against the in-tree library."""
    build()
    deps = [PROBE_SRC, LIB, os.path.join(os.path.dirname(HERE), "include", "caduceus_b200.h")]
    if not force and os.path.exists(PROBE_BIN) and all(os.path.getmtime(d) <= os.path.getmtime(PROBE_BIN) for d in deps):
        return PROBE_BIN
    os.makedirs(os.path.dirname(PROBE_BIN), exist_ok=True)
    cmd = [_nvcc(), "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
           "-I", os.path.join(os.path.dirname(HERE), "include"), PROBE_SRC, "-o", PROBE_BIN,
           "-L", CSRC, "-lcaduceus_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../caduceus_b200/csrc"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {PROBE_SRC}:\n{r.stdout}")
    return PROBE_BIN


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
    if "--probe" in sys.argv:
        print(build_probe(force="--force" in sys.argv))
