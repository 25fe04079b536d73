#!/bin/bash
# Builds the UNMODIFIED reference (henrystoldt/MAPLEAF) under baseline/_ref/ (git-ignored).
# Recipe verified in SURVEY.md Appendix B: the 5 .pyx files need Cython's
# c_api_binop_methods=True directive with Cython >= 3 (the reference targets 0.29 semantics).
# /root/reference is read-only, so the build happens in a copy.
set -e
REPO="$(cd "$(dirname "$0")/.." && pwd)"
SRC="${1:-/root/reference}"
if [ ! -d "$SRC/MAPLEAF" ]; then echo "reference source not found at $SRC"; exit 0; fi
mkdir -p "$REPO/baseline"
if [ ! -d "$REPO/baseline/_ref/MAPLEAF" ]; then
  cp -r "$SRC" "$REPO/baseline/_ref"
  chmod -R u+w "$REPO/baseline/_ref"
fi
cd "$REPO/baseline/_ref"
if ls MAPLEAF/Motion/CythonVector*.so >/dev/null 2>&1; then echo "reference already built"; exit 0; fi
python - <<'PY'
from setuptools import Extension, setup
from Cython.Build import cythonize
f=["MAPLEAF/Motion/CythonVector.pyx","MAPLEAF/Motion/CythonQuaternion.pyx","MAPLEAF/Motion/CythonAngularVelocity.pyx",
   "MAPLEAF/Rocket/CythonFinFunctions.pyx","MAPLEAF/IO/CythonLog.pyx"]
setup(name="x", script_args=["build_ext","--inplace"],
      ext_modules=cythonize([Extension(p[:-4].replace("/","."),[p]) for p in f], language_level="3",
                            compiler_directives={"c_api_binop_methods": True}))
PY
