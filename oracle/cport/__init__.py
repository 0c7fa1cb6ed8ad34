"""ctypes front end of the oracle's C restatement (``iact_oracle.c``).

TEST INFRASTRUCTURE / CPU BASELINE ONLY.  ``render`` mirrors ``oracle.trace.render`` for
single-stage telescopes with cylinder / box / sphere obstructions and hard sensors, multi-threaded
over facets; it is validated against the NumPy oracle in ``tests/test_oracle_cport.py``.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from .build import build, build_f64, build_native, LIB
from .. import trace as otrace


class _Sensor(C.Structure):
    _fields_ = [("kind", C.c_int), ("pos", C.c_float * 3), ("R", C.c_float * 9), ("width", C.c_int),
                ("height", C.c_int), ("x0", C.c_float), ("y0", C.c_float), ("dx", C.c_float), ("dy", C.c_float),
                ("edge_width", C.c_float), ("goff", C.c_float * 2), ("cr", C.c_float), ("sr", C.c_float),
                ("size", C.c_float), ("size_sqrt3", C.c_float), ("size_1p5", C.c_float), ("inradius", C.c_float),
                ("edge_thr", C.c_float), ("q_min", C.c_int), ("r_min", C.c_int), ("tq", C.c_int), ("tr", C.c_int),
                ("n_pixels", C.c_int), ("lookup", C.c_void_p)]


_lib = None
_lib_native = None
_lib_f64 = None


def _load(variant="exact"):
    global _lib, _lib_native, _lib_f64
    if variant == "f64":
        if _lib_f64 is None:
            _lib_f64 = C.CDLL(str(build_f64()))
            _lib_f64.oracle_render.restype = C.c_int
        return _lib_f64
    if variant == "native":
        if _lib_native is None:
            _lib_native = C.CDLL(str(build_native()))
            _lib_native.oracle_render.restype = C.c_int
        return _lib_native
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.oracle_render.restype = C.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _sensor_struct(s, keep):
    f = np.float32
    st = _Sensor()
    st.kind = 0 if s["type"] == "square" else 1
    R = otrace.euler_to_matrix(s["rotation"], f)
    for i in range(3):
        st.pos[i] = float(s["position"][i])
    for i in range(9):
        st.R[i] = float(R.reshape(-1)[i])
    if st.kind == 0:
        st.width, st.height = s["width"], s["height"]
        st.x0, st.y0, st.dx, st.dy, st.edge_width = s["x0"], s["y0"], s["dx"], s["dy"], s["edge_width"]
    else:
        st.goff[0], st.goff[1] = s["grid_offset"]
        st.cr = float(np.cos(f(-s["grid_rotation"])))
        st.sr = float(np.sin(f(-s["grid_rotation"])))
        st.size = s["hex_size"]
        st.size_sqrt3 = s["hex_size"] * 1.7320508075688772
        st.size_1p5 = s["hex_size"] * 1.5
        st.inradius = s["hex_inradius"]
        st.edge_thr = 1.0 - s["edge_width"] / s["hex_inradius"]
        st.q_min, st.r_min = s["q_min"], s["r_min"]
        lut = np.ascontiguousarray(s["lookup_table"], np.int32)
        keep.append(lut)
        st.tq, st.tr = lut.shape
        st.n_pixels = s["n_pixels"]
        st.lookup = lut.ctypes.data
    return st


def prepare(scene, sensor_idx=0):
    """World tables + obstruction tables + sensor struct (reusable across render calls)."""
    stages = otrace._stages(scene["groups"])
    if list(stages.keys()) != [0]:
        raise NotImplementedError("the C port covers single-stage telescopes")
    tp, tn, tw = otrace._primary_tables(stages, np.float32)
    obs = {g["type"]: g for g in scene["obstructions"]}
    if set(obs) - {"cylinder", "box", "sphere"}:
        raise NotImplementedError("the C port covers cylinder/box/sphere obstructions")
    c = lambda a: np.ascontiguousarray(a, np.float32)
    keep = []
    prep = dict(tp=c(tp), tn=c(tn), tw=c(tw[..., 0]), keep=keep,
                cyl=[c(obs["cylinder"][k]) for k in ("p1", "p2", "r")] if "cylinder" in obs else None,
                box=[c(obs["box"][k]) for k in ("p1", "p2")] if "box" in obs else None,
                sph=[c(obs["sphere"][k]) for k in ("centers", "radii")] if "sphere" in obs else None,
                sensor=scene["sensors"][sensor_idx])
    prep["struct"] = _sensor_struct(prep["sensor"], keep)
    return prep


def render(prep, sources, values, source_type="point", debug=False, threads=0, variant="exact"):
    """-> (image, n_threads_used) or, with debug, (xy, vals) in render_debug order.  ``variant="native"`` uses the
    -O3 -march=native build (timing only; the parity tests use the bit-exact one), ``variant="f64"`` the build whose
    per-ray chain runs in float64 (inputs and outputs stay float32 arrays)."""
    lib = _load(variant)
    src = np.ascontiguousarray(sources, np.float32)
    val = np.ascontiguousarray(values, np.float32)
    F, M = prep["tp"].shape[:2]
    S = len(src)
    s = prep["sensor"]
    shape = otrace.accumulator_shape(s)
    img = np.zeros(int(np.prod(shape)), np.float32)
    xy = np.zeros((F * S * M, 2), np.float32) if debug else None
    dv = np.zeros(F * S * M, np.float32) if debug else None
    cyl, box, sph = prep["cyl"], prep["box"], prep["sph"]
    nt = lib.oracle_render(
        C.c_int(F), C.c_int(M), _fp(prep["tp"]), _fp(prep["tn"]), _fp(prep["tw"]), C.c_int(S), _fp(src), _fp(val),
        C.c_int(0 if source_type == "point" else 1),
        C.c_int(len(cyl[2]) if cyl else 0), *(map(_fp, cyl) if cyl else (None, None, None)),
        C.c_int(len(box[0]) if box else 0), *(map(_fp, box) if box else (None, None)),
        C.c_int(len(sph[1]) if sph else 0), *(map(_fp, sph) if sph else (None, None)),
        C.byref(prep["struct"]), _fp(img), _fp(xy), _fp(dv), C.c_int(threads))
    if debug:
        return xy, dv
    return img.reshape(shape), nt
