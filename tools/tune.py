"""Kernel tuning aid. `tune.py build [names]` compiles variants of the CUDA library into mapleaf_b200/lib/variants/ (here, no GPU needed, in
parallel); on the GPU box `python tools/sweep.py --libs mapleaf_b200/lib/variants/a.so,... --blocks ...` times them."""
import os, subprocess, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
VDIR = os.path.join(REPO, "mapleaf_b200", "lib", "variants")
# name -> extra nvcc flags; edit for the experiment at hand (switches: MLB_FLIGHT_SHAPES, MLB_COMP_NOINLINE, MLB_COLD_MASK, MLB_MATH_NOINLINE,
# MLB_FIN_UNIFORM, MLB_FIN_LINEAR, MLB_CHUTE_BLOCK, MLB_CHUTE_CTAS, MLB_DERIV_NOINLINE)
# a model-specialised variant also needs the block header: `#include "<path>/libmapleaf_b200_<hash>.h"` (backend.writeSpecializationHeader) as
# the LAST line of the generated header, e.g. "-DMLB_SPEC_HEADER=..." is not needed: add the flag "NVCC:-include", "NVCC:<header>" instead
VARIANTS = {
    "base": ["-DMLB_FLIGHT_SHAPES(X)=X(1024,1)"],
    "split_barrier": ["-DMLB_FLIGHT_SHAPES(X)=X(1024,1)", "-DMLB_SPLIT_BARRIER=1"],
}


def build(names=None):
    from mapleaf_b200 import backend
    os.makedirs(VDIR, exist_ok=True)
    procs = []
    for name, flags in VARIANTS.items():
        if names and name not in names: continue
        out = os.path.join(VDIR, name + ".so")
        # function-like macros cannot be given with -D: they go through a generated header
        hdr = os.path.join(VDIR, name + ".h")
        with open(hdr, "w") as f:
            for fl in flags:
                if fl.startswith("NVCC:"):
                    continue
                k, v = fl[2:].split("=", 1)
                f.write("#define %s %s\n" % (k, v))
        cmd = ["nvcc"] + backend.NVCC_FLAGS + [fl[5:] for fl in flags if fl.startswith("NVCC:")] + ["-include", hdr, "-Xptxas", "-v", "-o", out, os.path.join(backend.CSRC_DIR, "mlb_api.cu")]
        procs.append((name, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for name, p in procs:
        txt = p.communicate()[0]
        lines = txt.splitlines()
        for i, l in enumerate(lines):
            if ("flightLanes" in l or "chuteLanes" in l) and "Compiling" in l:
                print(name, "|", l.split("'")[1][:40], "|", lines[i + 2].strip(), "|", lines[i + 3].strip())
        if p.returncode: print(name, "FAILED\n", txt[-2000:])


if __name__ == "__main__":
    build(sys.argv[2:] if len(sys.argv) > 1 and sys.argv[1] == "build" else sys.argv[1:])
