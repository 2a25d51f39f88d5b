"""Builds weaklysuperviseddl_b200/libwsdl_b200.so (the C-ABI library, include/wsdl_b200.h) with nvcc for
sm_100a.  In-tree, no torch dependency: the .so travels to the GPU box with the repo snapshot.

    python -m weaklysuperviseddl_b200.build [--force]
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libwsdl_b200.so")
ARCH_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC"]
STAMP = os.path.join(HERE, "build", "flags.txt")


def effective_flags():
    """Everything that changes the object code: part of the up-to-date check (WSDL_NVCC_EXTRA included)."""
    return [*ARCH_FLAGS, *NVCC_FLAGS, *os.environ.get("WSDL_NVCC_EXTRA", "").split()]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(os.path.dirname(HERE), "include", "wsdl_b200.h"))
    deps.append(os.path.abspath(__file__))
    return deps


def up_to_date() -> bool:
    if not os.path.exists(LIB):
        return False
    try:
        with open(STAMP) as fh:
            if fh.read() != " ".join(effective_flags()):
                return False
    except OSError:
        return False
    t = os.path.getmtime(LIB)
    return all(os.path.getmtime(d) <= t for d in _deps())


def nvcc_path():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    return None


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and up_to_date():
        return LIB
    nvcc = nvcc_path()
    if nvcc is None:
        raise RuntimeError("nvcc not found: cannot build libwsdl_b200.so (there is no CPU fallback)")
    objs = []
    build_dir = os.path.join(HERE, "build")
    os.makedirs(build_dir, exist_ok=True)
    procs = []
    for src in sources():
        obj = os.path.join(build_dir, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        cmd = [nvcc, *effective_flags(), "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    cmd = [nvcc, *ARCH_FLAGS, "-shared", "-o", LIB + ".tmp", *objs]
    subprocess.check_call(cmd)
    os.replace(LIB + ".tmp", LIB)
    with open(STAMP, "w") as fh:
        fh.write(" ".join(effective_flags()))
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
