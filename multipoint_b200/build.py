"""Build libmultipoint_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m multipoint_b200.build [--force]

The library is plain CUDA C++ behind the C ABI of include/multipoint_b200.h: it links only
cudart (static) and resolves cuTensorMapEncodeTiled from the driver at run time, so it loads on a
machine without a GPU (symbol checks) and travels to the GPU box as a built artefact.
"""
import concurrent.futures
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "lib", "obj")
LIB = os.path.join(HERE, "lib", "libmultipoint_b200.so")

SOURCES = ["mp_common.cu", "detector_head.cu", "descriptor_normalize.cu", "box_nms.cu",
           "sample_descriptors.cu", "match.cu", "match_tc.cu", "homographic.cu", "valid_mask.cu", "eval_points.cu", "activation_fused.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC,-fvisibility=hidden", "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]


def _nvcc():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: cannot build libmultipoint_b200.so")
    return nvcc


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "multipoint_b200.h"))
    nvcc = _nvcc()
    jobs = []
    for src in SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src[:-3] + ".o")
        if force or _stale(o, [s] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + ["-c", s, "-o", o])
    if jobs:
        with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, res in zip(jobs, ex.map(lambda c: subprocess.run(c, capture_output=True, text=True), jobs)):
                if verbose or res.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + res.stdout + res.stderr)
                if res.returncode != 0:
                    raise RuntimeError("nvcc failed for " + cmd[-3])
    objs = [os.path.join(OBJ, s[:-3] + ".o") for s in SOURCES]
    if force or jobs or _stale(LIB, objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
