"""Compile oracle/step_math.c (test infrastructure) with gcc into oracle/_build/ (git-ignored)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_build", "libstepmath.so")


def build(force: bool = False) -> str:
    src = os.path.join(HERE, "step_math.c")
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(src):
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-ffp-contract=off", src, "-o", OUT, "-lm"], check=True)
    return OUT


if __name__ == "__main__":
    print(build(True))
