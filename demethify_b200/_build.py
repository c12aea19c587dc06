"""In-tree build of libdemethify_sm100.so (nvcc, sm_100a only).  Used by __graft_entry__.build()."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libdemethify_sm100.so")
INCLUDE = os.path.join(HERE, "..", "include", "demethify_b200.h")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",            # elementwise arithmetic must round like numpy; dot products use explicit fma()
    "-Xcompiler", "-fPIC",
]


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    return [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))] + [INCLUDE, os.path.abspath(__file__)]


def _compile(nvcc, src, verbose):
    obj = os.path.join(OBJ, src[:-3] + ".o")
    srcp = os.path.join(CSRC, src)
    newest = max(os.path.getmtime(p) for p in [srcp] + _deps())
    if os.path.exists(obj) and os.path.getmtime(obj) >= newest:
        return obj, ""
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", srcp, "-o", obj]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{res.stdout}{res.stderr}")
    return obj, res.stderr


def build_variant(tag, flags):
    """Kernel experiments: the same sources with extra -D flags into demethify_b200/variants/lib_<tag>.so (load with DMF_LIB=...)."""
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    obj = os.path.join(HERE, "build_" + tag)
    os.makedirs(obj, exist_ok=True)
    os.makedirs(os.path.join(HERE, "variants"), exist_ok=True)
    out = os.path.join(HERE, "variants", f"lib_{tag}.so")

    def one(src):
        o = os.path.join(obj, src[:-3] + ".o")
        res = subprocess.run([nvcc] + NVCC_FLAGS + list(flags) + ["-c", os.path.join(CSRC, src), "-o", o], capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(res.stdout + res.stderr)
        return o
    with ThreadPoolExecutor(max_workers=8) as ex:
        objs = list(ex.map(one, _sources()))
    res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", out], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(res.stdout + res.stderr)
    return out


def build(force=False, verbose=False):
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    os.makedirs(OBJ, exist_ok=True)
    if force:
        for f in os.listdir(OBJ):
            os.remove(os.path.join(OBJ, f))
    srcs = _sources()
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        results = list(ex.map(lambda s: _compile(nvcc, s, verbose), srcs))
    objs = [o for o, _ in results]
    if verbose:
        sys.stderr.write("".join(log for _, log in results))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a"] + objs + ["-o", LIB],
                             capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}{res.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force="-f" in sys.argv, verbose="-v" in sys.argv))
