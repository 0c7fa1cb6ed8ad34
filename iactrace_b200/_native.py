"""ctypes binding of ``libiactrace_b200.so`` (the C ABI declared in ``include/iactrace_b200.h``).

There is no CPU fallback: if the shared library is missing or no CUDA device is
visible, every compute entry point raises ``RuntimeError``.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import os

# IACTRACE_B200_LIB selects another build of the same library (kernel tuning experiments); there is
# still no fallback: a missing file raises.
_LIB_PATH = Path(os.environ.get("IACTRACE_B200_LIB") or Path(__file__).resolve().parent / "csrc" / "libiactrace_b200.so")
_lib = None

MAX_ASPH = 8
MAX_STAGES = 4
MAX_POLY = 16
MIRROR_REC = 24
RUN_BOUND_FLOATS = 8    # IACT_RUN_BOUND_FLOATS: floats per 32-row run of IactScene.chunk_bounds

RNG_PARTITIONABLE = 0
RNG_LEGACY = 1
SOURCE_POINT = 0
SOURCE_PARALLEL = 1
SENSOR_SQUARE, SENSOR_HEX, SENSOR_SOFT_SQUARE, SENSOR_SOFT_HEX = 0, 1, 2, 3

_fp = C.c_void_p  # device pointers travel as plain addresses


class IactSurface(C.Structure):
    _fields_ = [("curvature", C.c_double), ("conic", C.c_double), ("n_aspheric", C.c_int32),
                ("aspheric", C.c_float * MAX_ASPH)]


class IactMirrorStage(C.Structure):
    _fields_ = [("n_mirrors", C.c_int32), ("records", _fp), ("verts", _fp)]


class IactSensor(C.Structure):
    _fields_ = [("kind", C.c_int32), ("position", C.c_float * 3), ("euler", C.c_float * 3),
                ("width", C.c_int32), ("height", C.c_int32),
                ("x0", C.c_double), ("y0", C.c_double), ("dx", C.c_double), ("dy", C.c_double),
                ("edge_width", C.c_double),
                ("hex_size", C.c_double), ("hex_inradius", C.c_double), ("grid_rotation", C.c_double),
                ("grid_offset", C.c_double * 2),
                ("q_min", C.c_int32), ("r_min", C.c_int32), ("table_q", C.c_int32), ("table_r", C.c_int32),
                ("n_pixels", C.c_int32), ("lookup", _fp),
                ("sigma", C.c_double), ("kernel_size", C.c_int32), ("hex_outer_radius", C.c_double)]


class IactScene(C.Structure):
    _fields_ = [("n_facets", C.c_int32), ("n_samples", C.c_int32), ("world", _fp), ("bounds", _fp),
                ("chunk_bounds", _fp),
                ("n_cyl", C.c_int32), ("cyl_p1", _fp), ("cyl_p2", _fp), ("cyl_r", _fp),
                ("n_box", C.c_int32), ("box_p1", _fp), ("box_p2", _fp),
                ("n_sph", C.c_int32), ("sph_c", _fp), ("sph_r", _fp),
                ("n_obox", C.c_int32), ("obox_c", _fp), ("obox_h", _fp), ("obox_R", _fp),
                ("n_tri", C.c_int32), ("tri_v0", _fp), ("tri_v1", _fp), ("tri_v2", _fp),
                ("n_stages", C.c_int32), ("stages", IactMirrorStage * MAX_STAGES),
                ("sensor", IactSensor), ("cull", C.c_int32)]


class IactFacets(C.Structure):
    _fields_ = [("n_facets", C.c_int32), ("n_samples", C.c_int32), ("positions", _fp), ("rotations", _fp),
                ("scale", _fp), ("points", _fp), ("normals", _fp), ("delta", _fp), ("weights", _fp)]


class IactGrads(C.Structure):
    _fields_ = [("positions", _fp), ("rotations", _fp), ("scale", _fp), ("weights", _fp), ("values", _fp),
                ("sources", _fp), ("sensor_position", _fp), ("sensor_euler", _fp),
                ("stage_positions", _fp), ("stage_rotations", _fp), ("points", _fp), ("nq", _fp), ("stage_surface", _fp)]


_KEY = C.c_uint32 * 2

_SIGNATURES = {
    "iact_last_error": (C.c_char_p, []),
    "iact_version": (C.c_int, []),
    "iact_device_count": (C.c_int, []),
    "iact_launch_count": (C.c_longlong, []),
    "iact_sample_disk_group": (C.c_int, [_KEY, C.c_int, C.c_int, C.c_int, C.POINTER(IactSurface), _fp, _fp,
                                         _fp, _fp, _fp, _fp, _fp]),
    "iact_sample_polygon_group": (C.c_int, [_KEY, C.c_int, C.c_int, C.c_int, C.POINTER(IactSurface), C.c_int,
                                            _fp, _fp, _fp, _fp, _fp, _fp, _fp]),
    "iact_sample_disk_group_rows": (C.c_int, [_KEY, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(IactSurface), _fp,
                                              _fp, _fp, _fp, _fp, _fp, _fp]),
    "iact_sample_polygon_group_rows": (C.c_int, [_KEY, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(IactSurface),
                                                 C.c_int, _fp, _fp, _fp, _fp, _fp, _fp, _fp]),
    "iact_random_normal": (C.c_int, [_KEY, C.c_int, C.c_int, _fp, _fp]),
    "iact_random_uniform": (C.c_int, [_KEY, C.c_int, C.c_int, C.c_float, C.c_float, _fp, _fp]),
    "iact_transform_to_world": (C.c_int, [C.POINTER(IactFacets), C.c_int, _fp, _fp, _fp]),
    "iact_transform_to_world_binned": (C.c_int, [C.POINTER(IactFacets), C.c_int, C.c_int, _fp, _fp, _fp, _fp]),
    "iact_render": (C.c_int, [C.POINTER(IactScene), _fp, _fp, C.c_int, C.c_int, _fp, _fp]),
    "iact_response_matrix": (C.c_int, [C.POINTER(IactScene), _fp, _fp, C.c_int, C.c_int, _fp, _fp]),
    "iact_render_debug": (C.c_int, [C.POINTER(IactScene), _fp, _fp, C.c_int, C.c_int, _fp, _fp, _fp, _fp]),
    "iact_render_vjp": (C.c_int, [C.POINTER(IactScene), C.POINTER(IactFacets), _fp, _fp, C.c_int, C.c_int,
                                  _fp, C.POINTER(IactGrads), _fp]),
    "iact_accumulate": (C.c_int, [C.POINTER(IactSensor), _fp, _fp, _fp, C.c_longlong, _fp, _fp]),
    "iact_cull_stats": (C.c_int, [C.POINTER(IactScene), _fp, C.c_int, C.c_int, _fp, _fp]),
    "iact_probe_fp32": (C.c_int, [C.c_int, C.POINTER(C.c_double), _fp]),
    "iact_work_plan": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_longlong, C.POINTER(C.c_int),
                                 C.POINTER(C.c_longlong)]),
    "iact_probe_smem_atomics": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double), _fp]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def library_path() -> Path:
    return _LIB_PATH


def lib():
    """Load the shared library (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not _LIB_PATH.exists():
            raise RuntimeError(
                f"iactrace_b200: CUDA extension not built ({_LIB_PATH} missing); run "
                "`python -c 'import __graft_entry__ as g; g.build()'` or "
                "`python iactrace_b200/csrc/build.py`.  There is no CPU fallback.")
        handle = C.CDLL(str(_LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def last_error() -> str:
    return lib().iact_last_error().decode()


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    msg = last_error()
    if rc == 1:
        raise ValueError(f"iactrace_b200 {what}: {msg}")
    if rc == 3:
        raise NotImplementedError(f"iactrace_b200 {what}: {msg}")
    raise RuntimeError(f"iactrace_b200 {what}: {msg}")


def require_cuda():
    """Return the torch module after asserting a usable CUDA device; raise loudly otherwise."""
    import torch
    lib()
    if not torch.cuda.is_available():
        raise RuntimeError("iactrace_b200: no CUDA device visible; the ray-tracing path has no CPU fallback")
    return torch


def ptr(t) -> int | None:
    """Device address of a contiguous float32/int32 CUDA tensor (None passes NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("iactrace_b200: expected a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError("iactrace_b200: expected a contiguous tensor")
    return t.data_ptr()


def stream_ptr():
    import torch
    return torch.cuda.current_stream().cuda_stream


def key_arg(key) -> "C.Array":
    return _KEY(int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF)
