"""Builds the in-tree shared library onedc_b200/libonedc_b200.so for sm_100a with nvcc.

The library is the product: there is no CPU fallback, `onedc_b200.lib.load()` raises if it is missing.
nvcc cross-compiles without a GPU, so this also runs in the CPU-only build container.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libonedc_b200.so")
SOURCES = ["core.cu", "igemm.cu", "attention.cu", "norm.cu", "elementwise.cu", "rans_host.cpp"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--use_fast_math", "-Xptxas", "-v"]


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "onedc_b200.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    procs = []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        obj = os.path.join(HERE, "build", s.rsplit(".", 1)[0] + ".o")
        objs.append(obj)
        if (not force) and os.path.exists(obj) and os.path.getmtime(obj) > max(
                os.path.getmtime(os.path.join(CSRC, f)) for f in os.listdir(CSRC)) and \
                os.path.getmtime(obj) > os.path.getmtime(os.path.join(HERE, "..", "include", "onedc_b200.h")):
            continue
        cmd = [nvcc] + NVCC_FLAGS + ["-c", src, "-o", obj]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for s, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed for {s}")
        with open(os.path.join(HERE, "build", s + ".ptxas.log"), "w") as f:
            f.write(out)
    cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-lcudart_static", "-lpthread", "-ldl", "-lrt"]
    subprocess.check_call(cmd)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
