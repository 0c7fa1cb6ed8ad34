"""Host arithmetic of the work queues (iact_work_plan, no device needed): the units the warps pull must cover every
(source, facet, sample row) exactly once, sample parts must start on 32-row boundaries (level-3 runs are row-aligned),
and large jobs must offer enough units per resident warp for the queue to balance."""
import ctypes as C

import numpy as np
import pytest

from iactrace_b200 import _native as N

CASES = [  # F, M, S
    (876, 115, 4096), (380, 1000, 1), (380, 64, 4096), (6, 16667, 10000), (1, 1, 1), (3, 31, 2), (876, 4096, 64),
    (7, 33, 5), (380, 1000, 7), (2, 100000, 1),
]


def _plan(kind, F, M, S, cull=1, warps=148 * 4 * 8):
    out = (C.c_int * 6)()
    n = C.c_longlong(0)
    rc = N.lib().iact_work_plan(kind, F, M, S, cull, warps, out, C.byref(n))
    assert rc == 0
    return list(out), n.value


@pytest.mark.parametrize("F,M,S", CASES)
def test_render_units_cover_every_ray_once(F, M, S):
    (fpu, runs, parts, msize, _, _), n_units = _plan(0, F, M, S)
    assert fpu >= 1 and runs == -(-F // fpu) and parts >= 1 and msize % 32 == 0
    assert n_units == S * runs * parts
    assert (parts - 1) * msize < M <= parts * msize
    # facet x row coverage of one source (every source gets the same units)
    cover = np.zeros((F, M), np.int32)
    for u in range(runs * parts):
        run, part = u // parts, u % parts
        f0, f1 = run * fpu, min(F, (run + 1) * fpu)
        m0, m1 = part * msize, min(M, (part + 1) * msize)
        assert f0 < f1 and m0 < m1
        cover[f0:f1, m0:m1] += 1
    assert (cover == 1).all()


@pytest.mark.parametrize("F,M,S", CASES)
def test_vjp_units_cover_every_ray_once(F, M, S):
    (slen, sruns, parts, msize, _, _), n_units = _plan(1, F, M, S, warps=148 * 2 * 8)
    assert 1 <= slen <= 32 and sruns == -(-S // slen) and msize % 32 == 0
    assert n_units == F * sruns * parts
    cover = np.zeros((S, M), np.int32)
    for u in range(sruns * parts):                           # the units of one facet
        sr, part = u % sruns, u // sruns
        s0, s1 = sr * slen, min(S, (sr + 1) * slen)
        m0, m1 = part * msize, min(M, (part + 1) * msize)
        assert s0 < s1 and m0 < m1
        cover[s0:s1, m0:m1] += 1
    assert (cover == 1).all()


def test_large_jobs_offer_many_units_per_warp_and_small_jobs_are_split():
    warps = 148 * 4 * 8
    (fpu, runs, parts, msize, _, _), n_units = _plan(0, 876, 115, 4096, warps=warps)
    assert n_units >= 32 * warps and parts == 1 and fpu <= 8
    (fpu, runs, parts, msize, _, _), n_units = _plan(0, 380, 1000, 1, warps=warps)     # BASELINE config 1: one source
    assert fpu == 1 and parts > 1 and n_units >= warps                                # split along the samples
    (slen, sruns, parts, msize, _, _), n_units = _plan(1, 876, 115, 4096, warps=148 * 2 * 8)
    assert n_units >= 32 * 148 * 2 * 8


def test_bad_arguments_are_reported():
    out = (C.c_int * 6)()
    n = C.c_longlong(0)
    assert N.lib().iact_work_plan(2, 1, 1, 1, 0, 1, out, C.byref(n)) != 0
    assert N.lib().iact_work_plan(0, 0, 1, 1, 0, 1, out, C.byref(n)) != 0
    assert b"iact_work_plan" in N.lib().iact_last_error()
