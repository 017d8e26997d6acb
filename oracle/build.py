"""Build the C oracle into ``oracle/_build/liboracle.so`` (test infrastructure only).

``python oracle/build.py`` or ``oracle.build.build()``.  The reference is pure Python (no C
sources to compile), so there is no ``oracle/_ref`` for this repo -- see DESIGN.md.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")
SOURCES = [os.path.join(HERE, "cptrack_oracle.c")]  # (the preprocessing and motion oracles are numpy: preprocess_oracle.py, motion_oracle.py)


def build(force=False):
    os.makedirs(OUT_DIR, exist_ok=True)
    srcs = [s for s in SOURCES if os.path.exists(s)]
    if not force and os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(s) for s in srcs):
        return LIB
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-o", LIB] + srcs + ["-lm"]
    subprocess.run(cmd, check=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
