"""The C-ABI library loads and exports every symbol include/iactrace_b200.h declares (no compute)."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "iactrace_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(iact_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built_lib):
    lib = ctypes.CDLL(str(built_lib))
    names = _declared()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_ctypes_binding_covers_the_header(built_lib):
    from iactrace_b200 import _native
    assert sorted(_native.EXPORTED_SYMBOLS) == _declared()
    handle = _native.lib()
    assert handle.iact_version() >= 100
    assert isinstance(_native.last_error(), str)
    assert handle.iact_launch_count() >= 0


def test_struct_layouts_match_the_header(built_lib):
    """sizeof() AND every field offset of the ctypes mirrors equal what the C compiler lays out."""
    import subprocess
    import tempfile
    from iactrace_b200 import _native as N
    structs = (N.IactSurface, N.IactMirrorStage, N.IactSensor, N.IactScene, N.IactFacets, N.IactGrads)
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "iactrace_b200.h"', 'int main(){']
    for c in structs:
        lines.append(f'printf("{c.__name__} %zu", sizeof({c.__name__}));')
        for name, *_ in c._fields_:
            lines.append(f'printf(" %zu", offsetof({c.__name__}, {name}));')
        lines.append('printf("\\n");')
    lines.append('return 0;}')
    with tempfile.TemporaryDirectory() as d:
        (Path(d) / "s.c").write_text("\n".join(lines))
        subprocess.run(["gcc", "-I", str(ROOT / "include"), str(Path(d) / "s.c"), "-o", str(Path(d) / "s")], check=True)
        out = subprocess.run([str(Path(d) / "s")], capture_output=True, text=True, check=True).stdout.strip().splitlines()
    assert len(out) == len(structs)
    for c, line in zip(structs, out):
        got = line.split()
        assert got[0] == c.__name__
        want = [ctypes.sizeof(c)] + [getattr(c, name).offset for name, *_ in c._fields_]
        assert [int(x) for x in got[1:]] == want, c.__name__
