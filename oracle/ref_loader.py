"""Test-infrastructure helper: import the UNMODIFIED reference (MAPLEAF) when an install exists.

The reference is pure Python + 5 Cython modules. `oracle/build_ref.sh` builds it under
`baseline/_ref/` (git-ignored). This loader stubs the modules
the reference imports at module scope but that are absent from this image
(matplotlib / mpl_toolkits / serial), and chdir-independent path handling is left to callers.

Only tests/, tests/golden/make_golden.py, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this file. The product package
(mapleaf_b200/) never does.
"""
import os
import sys
from unittest.mock import MagicMock

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(REPO, "baseline", "_ref")


def reference_available() -> bool:
    motion = os.path.join(REF_DIR, "MAPLEAF", "Motion")
    if not os.path.isdir(motion):
        return False
    return any(f.startswith("CythonVector") and f.endswith(".so") for f in os.listdir(motion))


def load_reference():
    """Returns the imported MAPLEAF package (reference). Raises ImportError if not installed."""
    if not reference_available():
        raise ImportError("reference install not found under baseline/_ref (run oracle/build_ref.sh)")
    for m in ("matplotlib", "matplotlib.pyplot", "matplotlib.animation", "mpl_toolkits",
              "mpl_toolkits.mplot3d", "mpl_toolkits.mplot3d.axes3d", "serial"):
        if m not in sys.modules:
            sys.modules[m] = MagicMock()
    sys.modules["matplotlib.pyplot"].subplots.return_value = (MagicMock(), MagicMock())
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import MAPLEAF  # noqa
    return MAPLEAF
