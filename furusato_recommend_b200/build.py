"""In-tree build of liblgcn_b200.so (nvcc, sm_100a only).

The library has no torch / Python dependency: it is the C ABI declared in
include/lgcn_b200.h.  `python -m furusato_recommend_b200.build` rebuilds it;
`__graft_entry__.build()` calls `build()`.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "liblgcn_b200.so"
STAMP = PKG / ".liblgcn_b200.stamp"

SOURCES = ["spmm.cu", "bpr.cu", "sampler.cu", "score_topk.cu", "score_topk_tc.cu", "metrics.cu", "ingest.cu", "ssm.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.exists(cand) else "nvcc"


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted([CSRC / x for x in SOURCES] + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "lgcn_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build_variant(tag: str, defines: dict, sources=("spmm.cu",)) -> Path:
    """Tuning aid: liblgcn_b200_<tag>.so with -D overrides applied to `sources` (the other objects
    come from the standard build).  Select it at run time with LGCN_B200_LIB=<path>."""
    build()
    out = PKG / f"liblgcn_b200_{tag}.so"
    objs = []
    for src in SOURCES:
        if src in sources:
            obj = CSRC / f"{src[:-3]}_{tag}.o"
            cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{k}={v}" for k, v in defines.items()], "-c", str(CSRC / src),
                   "-o", str(obj)]
            subprocess.run(cmd, check=True)
        else:
            obj = CSRC / (src[:-3] + ".o")
        objs.append(str(obj))
    subprocess.run([_nvcc(), "-shared", "-o", str(out), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
                    "-Xcompiler", "-fPIC", "-lcudart_static", "-ldl", "-lrt", "-lpthread"], check=True)
    return out


def build(force: bool = False, verbose: bool = False) -> Path:
    digest = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB
    objs = []
    procs = []
    for src in SOURCES:
        obj = CSRC / (src[:-3] + ".o")
        cmd = [_nvcc(), *NVCC_FLAGS, "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas")
            cmd.insert(2, "-v")
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(f"--- nvcc {src} ---\n{out}\n")
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError("nvcc failed building liblgcn_b200.so")
    link = [_nvcc(), "-shared", "-o", str(LIB), *objs, "-gencode", "arch=compute_100a,code=sm_100a",
            "-Xcompiler", "-fPIC", "-lcudart_static", "-ldl", "-lrt", "-lpthread"]
    subprocess.run(link, check=True)
    STAMP.write_text(digest)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
