"""Build libpaos_b200.so in-tree with nvcc for sm_100a (no torch, no JIT cache: the .so travels with the repo).

Usage: ``python paos_b200/build.py [--force]`` (run it by path: ``python -m paos_b200.build`` imports the package first,
which needs an up-to-date library).  ``__graft_entry__.build()`` calls :func:`build`.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
# PAOS_BUILD_TAG=<tag> builds an experiment variant (with PAOS_NVCC_EXTRA flags) next to the product library:
# libpaos_b200_<tag>.so, loaded by setting PAOS_LIB to its path (tools/ only; the product and the tests use the default)
_TAG = os.environ.get("PAOS_BUILD_TAG", "")
OBJ = os.path.join(HERE, "build" + ("_" + _TAG if _TAG else ""))
LIB = os.path.join(HERE, "libpaos_b200" + ("_" + _TAG if _TAG else "") + ".so")
SOURCES = ["runtime.cu", "aux_kernels.cu", "pass_c128.cu", "pass_c64.cu", "comm.cu", "sag_kernels.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--fmad=true", "-Xptxas", "-v",
] + (os.environ.get("PAOS_NVCC_EXTRA", "").split())


INFO = os.path.join(HERE, "build_info" + ("_" + _TAG if _TAG else "") + ".json")


def source_hash():
    """sha1 over every file of csrc/ and include/ plus the compiler flags: what the binary was made from."""
    import hashlib

    h = hashlib.sha1(" ".join(NVCC_FLAGS).encode())
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            h.update(f.encode())
            with open(os.path.join(root, f), "rb") as fh:
                h.update(fh.read())
    return h.hexdigest()


def built_hash():
    """Source hash recorded next to the shipped library (None when there is no record)."""
    import json

    try:
        with open(INFO) as fh:
            return json.load(fh).get("source_hash")
    except (OSError, ValueError):
        return None


def build(force=False, verbose=False):
    """Compile when forced, when the library is missing, or when the sources no longer hash to what the shipped binary
    records (modification times do not survive a copy to another box).  The hash is also compiled into the library
    (``paos_build_info()``), so a stale binary cannot pass for a fresh one."""
    import json
    import time

    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    want = source_hash()
    if not force and os.path.exists(LIB) and built_hash() == want:
        return LIB
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, f'-DPAOS_SOURCE_HASH="{want}"', "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        with open(obj + ".log", "w") as fh:
            fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB, *objs, "-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    # the roofline probe (tools/peaks.cu): a stand-alone binary, not part of the library
    probe = os.path.join(HERE, "..", "tools", "peaks.cu")
    if os.path.exists(probe):
        r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o",
                            os.path.join(HERE, "..", "tools", "peaks_b200"), probe], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for tools/peaks.cu:\n{r.stdout}\n{r.stderr}")
    ver = subprocess.run([nvcc, "--version"], capture_output=True, text=True).stdout.strip().splitlines()
    with open(INFO, "w") as fh:
        json.dump({"source_hash": want, "nvcc": ver[-1] if ver else "", "flags": NVCC_FLAGS, "sources": SOURCES,
                   "built_at": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()), "forced": bool(force)}, fh, indent=1)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
