"""Build the CPU oracle's C core (TEST INFRASTRUCTURE ONLY).

`python oracle/build.py` compiles oracle/slic_core.c -> oracle/_build/libslic_oracle.so
with gcc.  -ffp-contract=off keeps float32 rounding identical to a build
without FMA contraction (see the header of slic_core.c).

The reference (/root/reference) is pure Python (no C /
C++ / Cython sources), so there is nothing to compile into oracle/_ref/.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "slic_core.c")
OUT_DIR = os.path.join(HERE, "_build")
OUT = os.path.join(OUT_DIR, "libslic_oracle.so")


def build(force: bool = False) -> str:
    os.makedirs(OUT_DIR, exist_ok=True)
    if (not force and os.path.exists(OUT)
            and os.path.getmtime(OUT) >= os.path.getmtime(SRC)):
        return OUT
    flags = ["-O2", "-std=c11", "-fPIC", "-ffp-contract=off", "-fno-fast-math"]
    o32 = os.path.join(OUT_DIR, "slic_core_f32.o")
    o64 = os.path.join(OUT_DIR, "slic_core_f64.o")
    subprocess.run(["gcc", *flags, "-c", SRC, "-o", o32], check=True)
    # float64 twin of the two slic loops (scikit-image's own known-answer
    # tests use float64 images); connectivity is type-independent.
    subprocess.run(["gcc", *flags, "-DREAL=double", "-DSUFFIX(n)=n##_f64",
                    "-DOBIA_ORACLE_NO_CC", "-c", SRC, "-o", o64], check=True)
    # float32 twin with the colour accumulation contracted to FMA (see slic_core.c)
    ofma = os.path.join(OUT_DIR, "slic_core_f32_fma.o")
    subprocess.run(["gcc", *flags, "-DORACLE_FMA", "-DREAL=float", "-DSUFFIX(n)=n##_fma",
                    "-DOBIA_ORACLE_NO_CC", "-c", SRC, "-o", ofma], check=True)
    subprocess.run(["gcc", "-shared", "-o", OUT, o32, o64, ofma, "-lm"], check=True)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
