"""Builds librbphd.so (the sm_100a engine + C ABI) in-tree with nvcc.

    python -m monorfs_b200.build [--force]

The library is compiled for sm_100a only (no PTX for other architectures, no fallback path).
-fmad=false / -ffp-contract=off keep the FP64 arithmetic unfused so results match the reference's
scalar C# operation order.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "librbphd.so")
SOURCES = ["rbphd_kernels.cu", "rbphd_api.cu", "rbphd_microbench.cu", "rbphd_analysis.cu"]
HEADERS = ["rbphd_analysis.cuh", "rbphd_math.cuh", "rbphd_block.cuh", "rbphd_kernels.cuh", "rbphd_weight.cuh", "rbphd_murty.cuh",
           os.path.join("..", "..", "include", "rbphd.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",
    "-Xcompiler", "-fPIC,-ffp-contract=off,-O2",
    "-Xptxas", "-v",
    "--shared",
]


DEVICE_SOURCES = ["rbphd_kernels.cu", "rbphd_math.cuh", "rbphd_block.cuh", "rbphd_kernels.cuh", "rbphd_weight.cuh",
                  "rbphd_murty.cuh"]


def _code_only(text):
    """Source text without comments and with white space collapsed (string literals in the device sources hold no
    comment markers)."""
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    text = re.sub(r"//[^\n]*", " ", text)
    return re.sub(r"\s+", " ", text).strip()


def source_hash():
    """Hash of the CODE k_particle_update is compiled from (the device files without comments and white space, not
    the host-side ABI): ties ncu-derived counters under profiles/ to the kernel they were captured on, and survives
    edits of comments."""
    import hashlib
    h = hashlib.sha256()
    for f in sorted(DEVICE_SOURCES):
        path = os.path.join(CSRC, f)
        if os.path.exists(path):
            with open(path, "r", errors="replace") as fh:
                h.update(_code_only(fh.read()).encode())
    return h.hexdigest()[:16]


def nvcc_path():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build_variant(name, defines, verbose=False):
    """Experiment builds: librbphd_<name>.so with -D overrides of the launch shape (rbphd_block.cuh).
    Select one at run time with RBPHD_LIB=<path> (capi.py)."""
    os.makedirs(OUT_DIR, exist_ok=True)
    lib = os.path.join(OUT_DIR, "librbphd_%s.so" % name)
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-D%s=%s" % kv for kv in defines.items()] + ["-o", lib] + \
        [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(os.path.join(OUT_DIR, "build_%s.log" % name), "w") as fh:
        fh.write(" ".join(cmd) + "\n" + proc.stdout)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed for variant %s" % name)
    return lib


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = [nvcc_path()] + NVCC_FLAGS + ["-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    log = os.path.join(OUT_DIR, "build.log")
    with open(log, "w") as fh:
        fh.write(" ".join(cmd) + "\n" + proc.stdout)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("nvcc failed (see %s)" % log)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
