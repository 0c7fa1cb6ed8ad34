"""Build the oracle's C restatement (gcc + OpenMP) into oracle/cport/liboracle.so."""
from __future__ import annotations

import subprocess
from pathlib import Path

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
SRC = HERE / "iact_oracle.c"


def build(force: bool = False) -> Path:
    if not force and LIB.exists() and LIB.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB
    cmd = ["gcc", "-O2", "-march=x86-64-v2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared",
           str(SRC), "-o", str(LIB), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    print(build(force=True))
