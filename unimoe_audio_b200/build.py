"""Build libdcmoe_b200.so (hand-written sm_100a CUDA behind the C ABI of include/dcmoe_b200.h).

In-tree build with plain nvcc so that the shared library travels with the repository snapshot:
    python -m unimoe_audio_b200.build [--force]
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libdcmoe_b200.so")
SOURCES = ["api.cu", "router.cu", "plan_permute_combine.cu", "ffn_simt.cu", "ffn_tcgen05.cu", "ffn_tcgen05_stream.cu", "rmsnorm.cu", "ep.cu"]
HEADERS = [os.path.join(CSRC, "common.cuh"), os.path.join(CSRC, "ptx.cuh"), os.path.join(CSRC, "exp_fast.cuh"), os.path.join(CSRC, "route_token.cuh"), os.path.join(HERE, "..", "include", "dcmoe_b200.h")]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))] + HEADERS
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(args):
    src, obj, verbose = args
    cmd = [NVCC, *FLAGS, "-c", src, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    return src, res.returncode, res.stdout + res.stderr


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile the stale objects (all of them with ``force``) in parallel and link the shared library."""
    if not force and not _stale():
        return OUT
    from concurrent.futures import ThreadPoolExecutor

    objs, jobs = [], []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    newest_header = max(os.path.getmtime(h) for h in HEADERS)
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        if not os.path.exists(src):
            continue
        obj = os.path.join(HERE, "build", s.replace(".cu", ".o"))
        objs.append(obj)
        if force or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), newest_header):
            jobs.append((src, obj, verbose))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        for src, rc, log in pool.map(_compile, jobs):
            if verbose or rc != 0:
                sys.stderr.write(log)
            if rc != 0:
                raise RuntimeError(f"nvcc failed on {os.path.basename(src)}")
    cmd = [NVCC, "-shared", "-o", OUT, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
