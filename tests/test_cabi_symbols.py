"""CPU test (-m "not gpu"): the in-tree C-ABI library exports every entry point include/mivit.h declares, and the ctypes
binding (moleculardiffusion_mivit_b200/_lib.py) declares a signature for each of them.  The library links against libcuda,
which the CPU container does not have, so the export table is read with `nm` instead of dlopen; the GPU tests load it."""
import os
import re
import subprocess

from moleculardiffusion_mivit_b200 import build as _build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "mivit.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    # function declarations: `<type> mivit_xxx(`; the callback typedef `(*mivit_allreduce_fn)` is not a symbol
    return sorted(set(re.findall(r"\b(mivit_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported():
    if _build.needs_build():
        _build.build()
    out = subprocess.run(["nm", "-D", "--defined-only", _build.LIB], check=True, stdout=subprocess.PIPE, text=True).stdout
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    names = _declared()
    assert len(names) >= 25, names
    missing = [n for n in names if n not in exported]
    assert not missing, "declared in include/mivit.h but not exported by libmivit_b200.so: %s" % missing
    # plain C linkage: no mangled duplicate of an entry point
    assert not [e for e in exported if e.startswith("_Z") and any(("%d%s" % (len(n), n)) in e and e.startswith("_Z%d" % len(n))
                                                                   for n in names)]


def test_ctypes_binding_covers_the_header():
    src = open(os.path.join(ROOT, "moleculardiffusion_mivit_b200", "_lib.py")).read()
    bound = set(re.findall(r'"(mivit_[a-z0-9_]+)"\s*:', src))
    missing = [n for n in _declared() if n not in bound]
    assert not missing, "no ctypes signature for: %s" % missing


def test_abi_version_matches_header():
    src = open(os.path.join(ROOT, "include", "mivit.h")).read()
    ver = int(re.search(r"#define\s+MIVIT_ABI_VERSION\s+(\d+)", src).group(1))
    entry = open(os.path.join(ROOT, "__graft_entry__.py")).read()
    assert "mivit_abi_version() == %d" % ver in entry


def test_library_loads_without_a_driver_and_reports_its_abi_version():
    """The library links only the CUDA runtime (the one driver entry point it needs is resolved at run time), so it must
    dlopen on a machine without libcuda, answer mivit_abi_version() and expose every declared entry point -- no compute call."""
    import ctypes
    if _build.needs_build():
        _build.build()
    lib = ctypes.CDLL(_build.LIB)
    src = open(os.path.join(ROOT, "include", "mivit.h")).read()
    ver = int(re.search(r"#define\s+MIVIT_ABI_VERSION\s+(\d+)", src).group(1))
    lib.mivit_abi_version.restype = ctypes.c_int
    assert lib.mivit_abi_version() == ver
    for name in _declared():
        assert hasattr(lib, name), name
    lib.mivit_last_error.restype = ctypes.c_char_p
    assert isinstance(lib.mivit_last_error(), bytes)
