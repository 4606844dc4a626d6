"""Build libewvit.so in-tree with nvcc for sm_100a (no torch build machinery involved).

    python efficient-wavelet-vit_b200/build.py [--force]

The shared library lands in ``efficient-wavelet-vit_b200/ewvit/libewvit.so`` (git-ignored, but
it travels with the tree to the GPU box).  Object files are cached under ``build/``.
"""
import concurrent.futures
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "ewvit", "libewvit.so")
OBJ_DIR = os.path.join(HERE, "build")
INCLUDE = os.path.normpath(os.path.join(HERE, "..", "include"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "--use_fast_math" if False else "-DEWVIT_NO_FAST_MATH",   # parity first: IEEE fp32 math
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-I", INCLUDE,
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libewvit.so cannot be built (there is no fallback path)")


def _sources():
    return sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))


def _fingerprint():
    h = hashlib.sha1()
    for f in sorted(os.listdir(CSRC)) + [os.path.join(INCLUDE, "ewvit.h")]:
        path = f if os.path.isabs(f) else os.path.join(CSRC, f)
        h.update(f.encode())
        with open(path, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    stamp = os.path.join(OBJ_DIR, "fingerprint")
    fp = _fingerprint()
    if not force and os.path.exists(OUT) and os.path.exists(stamp) and open(stamp).read() == fp:
        return OUT
    os.makedirs(OBJ_DIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []

    def compile_one(src):
        obj = os.path.join(OBJ_DIR, src[:-3] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    link = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT, *objs,
            "-Xcompiler", "-fPIC", "-cudart", "static"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(fp)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
