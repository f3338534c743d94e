"""Builds csrc/*.cu into the in-tree C-ABI shared library with nvcc for sm_100a.

    python -m moleculardiffusion_mivit_b200.build

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.

Staleness: an object is rebuilt when its source, ANY header under csrc/ (*.cuh, *.h) or include/, or the nvcc flag set
(hashed into build/flags.txt) is newer / different.  Concurrency: N data-parallel ranks may import the package at once on a
stale tree, so the whole build runs under an exclusive file lock and the library is linked to a temporary name and moved into
place atomically (a rank that waited for the lock finds the tree fresh and returns)."""
import fcntl
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BUILD = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libmivit_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h"))
                  + glob.glob(os.path.join(HERE, "..", "include", "*.h")))


def _flags_tag():
    return hashlib.sha256((NVCC + "\0" + "\0".join(FLAGS)).encode()).hexdigest()


def _flags_current():
    try:
        with open(os.path.join(BUILD, "flags.txt")) as f:
            return f.read().strip() == _flags_tag()
    except OSError:
        return False


def needs_build():
    if not os.path.exists(LIB):
        return True
    if not _flags_current():
        # a library shipped without its build directory (the gpurun snapshot keeps both; a bare .so is trusted as built)
        if os.path.isdir(BUILD) and os.path.exists(os.path.join(BUILD, "flags.txt")):
            return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in sources() + headers())


def _build_locked(force, verbose):
    if not force and not needs_build():
        return LIB
    hdr_time = max([0.0] + [os.path.getmtime(h) for h in headers()])
    flags_ok = _flags_current()
    objs, procs = [], []
    for src in sources():
        obj = os.path.join(BUILD, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if not force and flags_ok and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), hdr_time):
            continue
        cmd = [NVCC] + FLAGS + ["-c", src, "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s" % src)
        with open(os.path.join(BUILD, os.path.basename(src)[:-3] + ".ptxas.log"), "w") as f:
            f.write(out)
    # no -lcuda: the one driver entry point (cuTensorMapEncodeTiled) is resolved at run time (csrc/tma.cuh), so the library
    # loads on a machine without a driver
    tmp = LIB + ".tmp.%d" % os.getpid()
    cmd = [NVCC, "-shared", "-o", tmp] + objs + ["-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        if os.path.exists(tmp):
            os.unlink(tmp)
        raise RuntimeError("link failed")
    os.replace(tmp, LIB)
    with open(os.path.join(BUILD, "flags.txt"), "w") as f:
        f.write(_flags_tag() + "\n")
    return LIB


def build(force=False, verbose=False):
    if not force and not needs_build():
        return LIB
    os.makedirs(BUILD, exist_ok=True)
    with open(os.path.join(BUILD, ".lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            return _build_locked(force, verbose)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
