"""Build libavc_b200.so (sm_100a) in-tree with nvcc.

    python -m adaptive_voice_conversion_b200.build [--force] [--verbose]

Every .cu under csrc/ is compiled to an object file (in parallel) with
``-gencode arch=compute_100a,code=sm_100a -lineinfo -O3`` and linked into
``adaptive_voice_conversion_b200/libavc_b200.so``.  nvcc cross-compiles, so this works in
the GPU-less authoring container; the .so travels to the GPU box with the tree.

Two variants are built from the same sources: ``libavc_b200.so`` and ``libavc_b200_pdl.so``
(``-DAVC_PDL=1``: programmatic dependent launch, see csrc/common.cuh); ``_lib.py`` picks one.
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "csrc", "_obj")
LIB = os.path.join(PKG, "libavc_b200.so")
VARIANTS = {  # name -> (library, object dir, extra nvcc flags)
    "default": (LIB, OBJ, []),
    "pdl": (os.path.join(PKG, "libavc_b200_pdl.so"), os.path.join(PKG, "csrc", "_obj_pdl"), ["-DAVC_PDL=1"]),
}
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC"]


def nvcc() -> str:
    exe = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(exe):
        raise RuntimeError("nvcc not found: libavc_b200.so cannot be built")
    return exe


def _digest(paths, extra=()) -> str:
    h = hashlib.sha256()
    for p in sorted(paths):
        with open(p, "rb") as f:
            h.update(os.path.basename(p).encode())  # location independent: the tree moves to the GPU box
            h.update(f.read())
    h.update(" ".join(ARCH + FLAGS + list(extra)).encode())
    return h.hexdigest()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def headers():
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hs.append(os.path.join(os.path.dirname(PKG), "include", "avc_b200.h"))
    return hs


def build_variant(variant: str, force: bool = False, verbose: bool = False, allow_build: bool = True) -> str:
    lib, objdir, extra = VARIANTS[variant]
    os.makedirs(objdir, exist_ok=True)
    stamp = os.path.join(objdir, "stamp.txt")
    dig = _digest(sources() + headers(), extra)
    if not force and os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == dig:
        return lib
    if not allow_build:
        raise RuntimeError(f"{lib} is missing or older than its sources and building was not allowed")
    cc = nvcc()
    hdr_dig = _digest(headers(), extra)

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        ostamp = obj + ".stamp"
        d = _digest([src], extra) + hdr_dig
        if not force and os.path.exists(obj) and os.path.exists(ostamp) and open(ostamp).read() == d:
            return obj
        cmd = [cc, *ARCH, *FLAGS, *extra, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        with open(ostamp, "w") as f:
            f.write(d)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, sources()))
    cmd = [cc, *ARCH, "-shared", "-Xcompiler", "-fPIC", "-o", lib, *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(dig)
    return lib


def build(force: bool = False, verbose: bool = False, allow_build: bool = True) -> str:
    """Build (or stamp-check) every variant; returns the default library's path."""
    for v in VARIANTS:
        build_variant(v, force, verbose, allow_build)
    return LIB


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(a.force, a.verbose))
