"""Build libicadv_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCES = ["icadv_core.cu", "icadv_conv_tc.cu", "icadv_conv_simt.cu", "icadv_perturb.cu", "icadv_entropy.cu", "icadv_msssim.cu",
           "icadv_train.cu", "icadv_split.cu", "icadv_probe.cu", "icadv_wgrad_tc.cu", "icadv_rans.cu", "icadv_graph.cu"]
OUT = os.path.join(os.path.dirname(HERE), "libicadv_b200.so")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC",
         "--use_fast_math" if False else "-DICADV_NO_FAST_MATH", "-Xptxas", "-v"]


def _newer(a, b):
    return not os.path.exists(b) or os.path.getmtime(a) > os.path.getmtime(b)


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    headers = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(os.path.dirname(HERE)), "include", "icadv.h"))
    os.makedirs(os.path.join(HERE, "build"), exist_ok=True)
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(HERE, "build", src.replace(".cu", ".o"))
        if force or _newer(s, o) or any(_newer(h, o) for h in headers):
            cmd = [nvcc] + FLAGS + ["-c", s, "-o", o]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}")
        objs.append(o)
    if force or any(_newer(o, OUT) for o in objs):
        cmd = [nvcc, "-shared", "-o", OUT] + objs + ["-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
