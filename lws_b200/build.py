"""Build liblws_b200.so in-tree with nvcc for sm_100a (B200).

    python -m lws_b200.build [--force]

The shared library is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.environ.get("LWSB_LIB_PATH") or os.path.join(HERE, "liblws_b200.so")
SOURCES = ["api.cu", "kernels_generic.cu", "kernels_batch.cu", "kernels_online.cu", "kernels_fft.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC,-O3,-Wall,-Wno-unused-function", "--shared", "-cudart", "static",
    "-Xptxas", "-v",
]


def _newest_source():
    t = 0.0
    for d in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(d):
            t = max(t, os.path.getmtime(os.path.join(d, f)))
    return max(t, os.path.getmtime(os.path.abspath(__file__)))


def build(force=False, verbose=False):
    """One nvcc process per translation unit (in parallel), then a link step.  LWSB_NVCC_EXTRA adds flags
    (e.g. -DLWSB_PAIR_EXPERIMENTS for the kernel-variant experiments of tools/gpu_pair.py)."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source():
        return LIB
    nvcc = os.environ.get("NVCC", "nvcc")
    extra = os.environ.get("LWSB_NVCC_EXTRA", "").split()
    objdir = os.path.join(HERE, os.environ.get("LWSB_OBJDIR", "build"))
    os.makedirs(objdir, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]
    compile_flags = [f for f in NVCC_FLAGS if f not in ("--shared",)]
    procs = []
    for s in srcs:
        obj = os.path.join(objdir, s[:-3] + ".o")
        cmd = [nvcc] + compile_flags + extra + ["-c", "-o", obj, os.path.join(CSRC, s)]
        procs.append((cmd, obj, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    log, failed = [], False
    for cmd, obj, pr in procs:
        out = pr.communicate()[0]
        log.append(" ".join(cmd) + "\n" + out)
        failed = failed or pr.returncode != 0
    if not failed:
        cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "--shared", "-cudart", "static", "-Xcompiler", "-fPIC",
               "-o", LIB] + [obj for _, obj, _ in procs]
        res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log.append(" ".join(cmd) + "\n" + res.stdout)
        failed = res.returncode != 0
    text = "\n".join(log)
    if verbose or failed:
        sys.stderr.write(text)
    if failed:
        raise RuntimeError("nvcc failed building liblws_b200.so")
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write(text)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
