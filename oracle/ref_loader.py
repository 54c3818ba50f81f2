"""Import the UNMODIFIED reference model from /root/reference (build container only).

TEST INFRASTRUCTURE ONLY.  /root/reference does not exist on the GPU box, so nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls this; it is used by ``oracle/make_golden.py``
(which produced ``tests/golden``) and by the ``not gpu`` test that re-checks the oracle against
the live reference when the reference tree is present.

The reference imports matplotlib at CVSR_freq.py:16 and *calls* it inside the forward
(featuremap_visual -> plt.title, :64), so the stub swallows arbitrary attribute access / calls.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = "/root/reference/CVSR_train"


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Stub(name)

    def __call__(self, *a, **k):
        return None


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "arch", "CVSR_freq.py"))


def load():
    """Returns the reference module CVSR_train/arch/CVSR_freq.py."""
    if not available():
        raise RuntimeError("reference tree not present (expected only in the build container)")
    for m in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot"):
        if m not in sys.modules:
            sys.modules[m] = _Stub(m)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import importlib
    return importlib.import_module("arch.CVSR_freq")
