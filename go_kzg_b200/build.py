"""Builds libb200kzg.so (hand-written sm_100a CUDA + the C ABI of include/b200_kzg.h) in-tree.

nvcc cross-compiles without a GPU; the three translation units compile in parallel.  The built
library lands in go_kzg_b200/lib/ (git-ignored, but it travels to the GPU box with gpurun)."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libb200kzg.so")
UNITS = ["api.cu", "kernels_fr.cu", "kernels_g1.cu", "kernels_msm.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"]


def _deps():
    out = []
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in os.listdir(root):
            if f.endswith((".cu", ".cuh", ".h")):
                out.append(os.path.join(root, f))
    return out


def _check_ptxas(unit: str, log: str) -> None:
    """Report register spills (performance signal only; ptxas accounts the spill slots of ABI
    device functions in the calling kernel's cumulative stack size)."""
    lines = log.splitlines()
    for i, ln in enumerate(lines):
        if "Function properties for" in ln and i + 1 < len(lines) and _spill_bytes(lines[i + 1]) > 256:
            sys.stderr.write("[build] %s: %s: %s\n" % (unit, ln.split("Function properties for")[1].strip()[:60], lines[i + 1].strip()))


def _spill_bytes(line: str) -> int:
    import re
    m = re.search(r"(\d+) bytes spill stores", line)
    return int(m.group(1)) if m else 0


def _compile(unit: str, verbose: bool, defines=(), tag: str = "") -> str:
    obj = os.path.join(OBJ, unit.replace(".cu", tag + ".o"))
    cmd = [NVCC, *FLAGS, *["-D" + d for d in defines], "-Xptxas", "-v", "-c", os.path.join(CSRC, unit), "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if verbose or r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode:
        raise RuntimeError("nvcc failed on " + unit)
    _check_ptxas(unit, r.stdout + r.stderr)
    return obj


def _source_digest() -> str:
    """SHA-256 over the names and contents of every source the library is built from (mtimes do not survive a copy of the
    tree to another machine in any useful order, contents do)."""
    import hashlib
    h = hashlib.sha256()
    for path in sorted(_deps()):
        h.update(os.path.basename(path).encode() + b"\0")
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = LIB + ".srchash"
    digest = _source_digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(len(UNITS)) as ex:
        objs = list(ex.map(lambda u: _compile(u, verbose), UNITS))
    subprocess.check_call([NVCC, "-shared", "-o", LIB, *objs, "-lcudart"])
    with open(stamp, "w") as f:
        f.write(digest + "\n")
    return LIB


def build_variant(tag: str, defines) -> str:
    """Experiment builds (tools/sessions): the same library with extra -D flags as lib/libb200kzg_<tag>.so, selected at
    run time with the environment variable B200_KZG_LIB.  Not part of the product build."""
    out = os.path.join(HERE, "lib", "libb200kzg_%s.so" % tag)
    os.makedirs(OBJ, exist_ok=True)
    with ThreadPoolExecutor(len(UNITS)) as ex:
        objs = list(ex.map(lambda u: _compile(u, False, defines, "_" + tag), UNITS))
    subprocess.check_call([NVCC, "-shared", "-o", out, *objs, "-lcudart"])
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build_variant(sys.argv[i + 1], sys.argv[i + 2:]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
