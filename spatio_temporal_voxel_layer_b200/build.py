"""Build libstvl_b200.so (CUDA kernels + C-ABI) in-tree for sm_100a with nvcc.

    python -m spatio_temporal_voxel_layer_b200.build [--force]

The .so is git-ignored but travels to the GPU box with the repo snapshot.
"""
import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(PKG_DIR)
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libstvl_b200.so")
SOURCES = ["stvl_mark.cu", "stvl_sweep.cu", "stvl_kernels.cu", "stvl_voxel_filter.cu", "stvl_costs.cu", "stvl_api.cu", "stvl_geometry.cpp"]
HEADERS = ["stvl_device.cuh", "stvl_common.cuh", "stvl_kernels.h", "stvl_geometry.h", "stvl_voxel_filter.h"]

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",                       # device: never contract a*b+c (bit-exact parity with the CPU path)
    "-Xcompiler", "-fPIC,-ffp-contract=off,-fno-fast-math,-Wall",
    "-cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or put /usr/local/cuda/bin on PATH)")


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    if os.path.exists(STAMP_PATH):
        return not library_matches_sources()
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.join(REPO_ROOT, "include", "stvl_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


TRACE_LIB_PATH = os.path.join(PKG_DIR, "libstvl_b200_trace.so")
STAMP_PATH = LIB_PATH + ".srchash"


def sources_hash():
    import hashlib
    h = hashlib.sha256()
    for f in sorted(SOURCES + HEADERS):
        h.update(f.encode())
        h.update(open(os.path.join(CSRC, f), "rb").read())
    h.update(open(os.path.join(REPO_ROOT, "include", "stvl_b200.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def library_matches_sources():
    """The library on disk was built from exactly the sources on disk (mtimes lie after a checkout or a snapshot copy)."""
    try:
        return os.path.exists(LIB_PATH) and open(STAMP_PATH).read().strip() == sources_hash()
    except OSError:
        return False


def build_library(force=False, verbose=False, trace=False, csrc=None, out=None):
    """Compile every source to an object file (in parallel) and link libstvl_b200.so.
    trace=True builds the profiling-only variant (-DSTVL_TRACE: per-CTA %globaltimer phase stamps, read back by
    profiles/trace_phases.py through STVL_B200_LIB) into a separate file; the product library never contains it.
    csrc / out: build another source tree into another file (scratch builds during development)."""
    if csrc is None and out is None and not trace and not force and not needs_build():
        return LIB_PATH
    import concurrent.futures
    import tempfile
    src_dir = csrc or CSRC
    out = out or (TRACE_LIB_PATH if trace else LIB_PATH)
    base = [_nvcc()] + NVCC_FLAGS + ["-I", os.path.join(REPO_ROOT, "include"), "-I", src_dir]
    if trace:
        base += ["-DSTVL_TRACE", "-rdc=true"]   # the stamp buffer is one device symbol shared by the kernel files
    if verbose:
        base += ["-Xptxas", "-v"]
    with tempfile.TemporaryDirectory(prefix="stvl_build_") as tmp:
        def compile_one(f):
            obj = os.path.join(tmp, f + ".o")
            res = subprocess.run(base + ["-x", "cu", "-c", os.path.join(src_dir, f), "-o", obj], capture_output=True, text=True)
            return f, obj, res
        with concurrent.futures.ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
            results = list(ex.map(compile_one, SOURCES))
        log = ""
        for f, obj, res in results:
            if res.returncode != 0:
                raise RuntimeError("nvcc failed on %s:\n%s%s" % (f, res.stdout, res.stderr))
            log += res.stderr
        link = [_nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + [obj for _, obj, _ in results]
        res = subprocess.run(link, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError("link failed:\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(log)
    if out == LIB_PATH:
        with open(STAMP_PATH, "w") as f:
            f.write(sources_hash() + "\n")
    return out


if __name__ == "__main__":
    def _opt(name):
        return sys.argv[sys.argv.index(name) + 1] if name in sys.argv else None
    print(build_library(force="--force" in sys.argv, verbose="-v" in sys.argv, trace="--trace" in sys.argv,
                        csrc=_opt("--csrc"), out=_opt("--out")))
