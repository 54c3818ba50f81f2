"""CPU tests of the C-ABI boundary: the library builds for sm_100a, loads, and exports every symbol that
include/fcvsr_b200.h declares (no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

from fcvsr_b200 import _capi, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "fcvsr_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fcvsr_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fcvsr_b200.h but not exported"


def test_python_binding_covers_the_header():
    names = set(_declared()) - {"fcvsr_version"}
    assert names == set(_capi.SIGNATURES), names ^ set(_capi.SIGNATURES)
    assert _capi.version().startswith("fcvsr_b200")


def test_no_product_import_of_the_oracle():
    """The product package must never route through the oracle / a CPU fallback."""
    pkg = os.path.join(ROOT, "fcvsr_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_sass_contains_tcgen05_and_tma():
    """The conv kernel really is a tcgen05/TMEM/TMA kernel (B200_PROFILING.md evidence table)."""
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        return
    obj = os.path.join(ROOT, "fcvsr_b200", "_lib", "conv_tc.o")
    sass = subprocess.run([cuobjdump, "-sass", obj], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass and "UTMALDG" in sass and "LDTM" in sass
