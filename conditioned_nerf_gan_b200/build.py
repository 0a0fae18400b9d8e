"""Build libcng_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI).

    python -m conditioned_nerf_gan_b200.build [--force] [--verbose]
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
LIB = os.path.join(PKG, "libcng_b200.so")
STAMP = os.path.join(PKG, "csrc", ".build_stamp")
SOURCES = ["cng_api.cu", "raymarch_gather.cu", "film_siren_simt.cu", "film_siren_tc.cu", "composite.cu", "sample_pdf.cu", "backward.cu", "film_siren_bwd_tc.cu", "group_norm.cu", "render.cu"]
# CNG_BUILD_EXPERIMENTAL=1 adds the two measured-slower K2 organisations kept for A/B work: CTA pairs
# (cta_group::2) and the layer-pipelined single-tile kernel; selected at run time with CNG_TC_CG=2 / CNG_TC_V=3
EXPERIMENTAL = os.environ.get("CNG_BUILD_EXPERIMENTAL", "0") == "1"
if EXPERIMENTAL:
    SOURCES += ["film_siren_tc2.cu", "film_siren_tc3.cu"]
# CNG_TC_EPI_WARPS (4, 8 or 12; 12 needs CNG_TC_EPI_PIPELINE=0 to fit 72 registers): epilogue warps per tile slot of the one-CTA-per-SM tcgen05 kernel (film_siren_tc.cu)
EPI_WARPS = os.environ.get("CNG_TC_EPI_WARPS", "8")
EPI_PIPELINE = os.environ.get("CNG_TC_EPI_PIPELINE", "1")    # 1: double-buffer the epilogue's TMEM loads
NVCC_FLAGS = [
    f"-DCNG_TC_EPI_WARPS={EPI_WARPS}", f"-DCNG_TC_EPI_PIPELINE={EPI_PIPELINE}", *(["-DCNG_WITH_EXPERIMENTAL_K2"] if EXPERIMENTAL else []),
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def _digest() -> str:
    h = hashlib.sha256()
    paths = [os.path.join(CSRC, s) for s in SOURCES] + sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    paths.append(os.path.join(ROOT, "include", "cng_b200.h"))
    for path in paths:
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, out: str = None, extra_flags=()) -> str:
    """Compile every .cu for sm_100a and link the shared library. Returns its path.
    ``out`` / ``extra_flags`` build an experimental variant next to the default library (see _lib.CNG_LIB)."""
    if out is not None:
        return _build_variant(out, list(extra_flags), verbose)
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read().strip() == digest:
        return LIB
    nvcc = _nvcc()
    objs = []
    procs = []
    for s in SOURCES:
        obj = os.path.join(CSRC, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", os.path.join(CSRC, s), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
            print(" ".join(cmd))
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for s, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{out}")
        if verbose and out:
            print(out)
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB, *objs, "-ldl"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    with open(STAMP, "w") as f:
        f.write(digest)
    return LIB


def _build_variant(out: str, extra_flags, verbose: bool) -> str:
    nvcc = _nvcc()
    tmp = os.path.join(CSRC, "_variant")
    os.makedirs(tmp, exist_ok=True)
    procs, objs = [], []
    for s in SOURCES:
        obj = os.path.join(tmp, s.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *extra_flags, "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-c", os.path.join(CSRC, s), "-o", obj]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for s, pr in procs:
        o, _ = pr.communicate()
        if pr.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}:\n{o}")
        if verbose and o:
            print(o)
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out, *objs, "-ldl"], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
