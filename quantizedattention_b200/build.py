"""Build the C-ABI libraries (sm_100a) in-tree with nvcc.  No torch dependency in the libraries.

  libqattn.so      product: include/qattn.h only, no debug hooks in the kernels
  libqattn_dev.so  development (`--dev`): the same sources with -DQA_DEV_TIMELINE plus csrc/probe.cu; additionally
                   exports include/qattn_dev.h (hardware layout probes, kernel timeline hooks)
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libqattn.so")
LIB_DEV = os.path.join(HERE, "libqattn_dev.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
SOURCES = ["host.cu", "quant.cu", "attn_int8_fwd.cu", "attn_bf16_fwd.cu", "attn_bf16_fwd2.cu", "attn_jvp.cu", "attn_int8_bwd.cu",
           "attn_bf16_bwd.cu", "attn_bf16_bwd2.cu", "prepass.cu", "quant_fp4.cu", "attn_fp4_fwd.cu"]
DEV_SOURCES = ["probe.cu"]


def sources(dev: bool = False):
    names = SOURCES + (DEV_SOURCES if dev else [])
    return [os.path.join(CSRC, s) for s in names if os.path.exists(os.path.join(CSRC, s))]


def stale(dev: bool = False) -> bool:
    lib = LIB_DEV if dev else LIB
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = sources(dev) + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False, dev: bool = False) -> str:
    lib = LIB_DEV if dev else LIB
    if not force and not stale(dev):
        return lib
    objs = []
    bdir = os.path.join(HERE, "build_dev" if dev else "build")
    os.makedirs(bdir, exist_ok=True)
    procs = []
    for src in sources(dev):
        obj = os.path.join(bdir, os.path.basename(src) + ".o")
        objs.append(obj)
        cmd = [NVCC, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-c",
               "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v", "-I", CSRC, src, "-o", obj]
        if dev:
            cmd.append("-DQA_DEV_TIMELINE")
        cmd += os.environ.get("QA_NVCC_EXTRA", "").split()     # development switches
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log = []
    for src, p in procs:
        out, _ = p.communicate()
        log.append(f"== {os.path.basename(src)}\n{out}")
        if p.returncode != 0:
            sys.stderr.write("\n".join(log))
            raise RuntimeError(f"nvcc failed on {src}")
    link = [NVCC, "--shared", "-o", lib] + objs + ["-lcudart"]
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError("link failed")
    with open(os.path.join(bdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, dev="--dev" in sys.argv))
