"""Builds libkmcb200.so (the C-ABI CUDA library) in-tree with nvcc for sm_100a.

    python -m kmcislandgrowth_b200.build [--force]

Every .cu file of csrc/ is one translation unit, compiled in parallel (the persistent relaxation kernel alone
takes ~2 minutes) and only when it or a header changed; the objects are linked into one shared object.  The
shared object is git-ignored but travels to the GPU box with the repository snapshot.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
OUT = os.path.join(HERE, "libkmcb200.so")
SOURCES = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
HEADERS = sorted(f for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))) + [os.path.join("..", "..", "include", "kmc_b200.h")]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def nvcc_path():
    for p in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if p and (os.path.isabs(p) and os.path.exists(p) or not os.path.isabs(p)):
            return p
    return "nvcc"


def _mtime(p):
    return os.path.getmtime(p) if os.path.exists(p) else 0.0


def _obj(src):
    return os.path.join(OBJ, src[:-3] + ".o")


def _stale(src):
    t = _mtime(_obj(src))
    return t == 0.0 or _mtime(os.path.join(CSRC, src)) > t or any(_mtime(os.path.join(CSRC, h)) > t for h in HEADERS)


def up_to_date():
    if not os.path.exists(OUT):
        return False
    t = os.path.getmtime(OUT)
    return all(_mtime(os.path.join(CSRC, d)) <= t for d in SOURCES + HEADERS)


def build(force=False, verbose=False):
    if not force and up_to_date():
        return OUT
    os.makedirs(OBJ, exist_ok=True)
    env = dict(os.environ)
    env.pop("CC", None)
    env.pop("CXX", None)
    nvcc = nvcc_path()
    extra = os.environ.get("KMC_NVCC_EXTRA", "").split()

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + extra + ["-c", "-o", _obj(src), src]
        r = subprocess.run(cmd, cwd=CSRC, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        return src, " ".join(cmd), r.returncode, r.stdout

    todo = [s for s in SOURCES if force or extra or _stale(s)]
    logs = []
    with ThreadPoolExecutor(max_workers=max(1, min(len(todo), os.cpu_count() or 1))) as ex:
        results = list(ex.map(compile_one, todo))
    for src, cmd, rc, out in results:
        logs.append(cmd + "\n" + out)
    link = [nvcc, "-shared", "-o", OUT] + [_obj(s) for s in SOURCES]
    failed = [r for r in results if r[2] != 0]
    if not failed:
        r = subprocess.run(link, cwd=CSRC, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        logs.append(" ".join(link) + "\n" + r.stdout)
        if r.returncode != 0:
            failed = [("link", " ".join(link), r.returncode, r.stdout)]
    with open(os.path.join(HERE, "build.log"), "w") as f:
        f.write("\n".join(logs))
    if verbose:
        print("\n".join(logs))
    if failed:
        raise RuntimeError("nvcc failed:\n" + "\n".join(x[3][-4000:] for x in failed))
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
