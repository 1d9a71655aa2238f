#!/usr/bin/env python3
"""Build libsvk.so (the sm_100a kernel library) in-tree with nvcc.  No torch headers are needed: the library is a
plain C-ABI shared object (include/svk.h) loaded with ctypes."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(os.path.dirname(HERE), "svk")
SOURCES = ["svk_api.cu", "elementwise.cu", "conv_simt.cu", "conv_tc.cu", "conv_tc3.cu", "conv_tc_wgrad9.cu", "conv_tc_wgradr.cu", "pool_gemm.cu", "gemm_tc.cu", "aam_score.cu", "aam_fused.cu", "backend.cu"]
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "--extended-lambda", "-std=c++17",
         "-Xcompiler", "-fPIC,-fvisibility=hidden", "-Xptxas", "-v"]


def newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def build(force=False, verbose=False, extra_flags=(), out_name="libsvk.so"):
    """extra_flags / out_name: A/B builds of kernel variants (`build.py --variant NAME -DX=1 ...` -> svk/libsvk_NAME.so,
    selected at run time with SVK_LIB_PATH)."""
    out = os.path.join(OUT_DIR, out_name)
    srcs = [os.path.join(HERE, s) for s in SOURCES]
    deps = srcs + [os.path.join(HERE, "svk_common.cuh"), os.path.join(HERE, "tc_common.cuh"),
                   os.path.join(HERE, "..", "..", "include", "svk.h")]
    if not force and os.path.exists(out) and os.path.getmtime(out) >= newest(deps):
        return out
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    bdir = os.path.join(HERE, "build", out_name.replace(".so", ""))
    os.makedirs(bdir, exist_ok=True)
    for s in srcs:
        o = os.path.join(bdir, os.path.basename(s) + ".o")
        objs.append(o)
        procs.append((s, subprocess.Popen([nvcc] + FLAGS + list(extra_flags) + ["-c", s, "-o", o], stdout=subprocess.PIPE,
                                          stderr=subprocess.STDOUT, text=True)))
    log = []
    for s, p in procs:
        text, _ = p.communicate()
        log.append(text)
        if p.returncode != 0:
            sys.stderr.write(text)
            raise RuntimeError("nvcc failed on %s" % s)
    with open(os.path.join(bdir, "ptxas.log"), "w") as f:
        f.write("\n".join(log))
    if verbose:
        print("\n".join(log))
    subprocess.check_call([nvcc, "-shared", "-o", out] + objs + ["-lcudart"])
    import ctypes
    ctypes.CDLL(out)          # fail here, not on the GPU box, if a symbol is unresolved
    return out


if __name__ == "__main__":
    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        print(build(force=True, extra_flags=[a for a in sys.argv[i + 2:] if a.startswith("-D")], out_name="libsvk_%s.so" % sys.argv[i + 1]))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
