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


LIB_F64 = HERE / "liboracle_f64.so"


def build_f64(force: bool = False) -> Path:
    """The same source with the per-ray chain in float64 (-DORACLE_REAL=double): the reference's operations at a
    precision where shadow and pixel-edge decisions are exact -- full-size image checks (tools/parity_fullsize.py)."""
    if not force and LIB_F64.exists() and LIB_F64.stat().st_mtime >= SRC.stat().st_mtime:
        return LIB_F64
    cmd = ["gcc", "-O2", "-march=x86-64-v2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-fPIC", "-shared",
           "-DORACLE_REAL=double", str(SRC), "-o", str(LIB_F64), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{r.stderr}")
    return LIB_F64


def build_native() -> Path:
    """The same source at full optimisation FOR THE HOST IT RUNS ON (-march=native), FMA contraction allowed: not
    bit-exact with the reference's op-by-op float32 any more, but the honest speed of this port on these cores
    (bench.py reports it beside the bit-exact build).  Always rebuilt, into a temporary directory: a -march=native
    binary must not travel between machines."""
    import tempfile
    out = Path(tempfile.mkdtemp(prefix="iact_oracle_native_")) / "liboracle_native.so"
    cmd = ["gcc", "-O3", "-march=native", "-fno-math-errno", "-fopenmp", "-fPIC", "-shared", str(SRC), "-o", str(out), "-lm"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"gcc failed:\n{r.stderr}")
    return out


if __name__ == "__main__":
    print(build(force=True))
