"""Gradients w.r.t. the surface parameters (reference core/surfaces.py:25-65: curvature, conic, parent-surface
offsets), against float64 autodiff of the oracle:

* stage-0 groups: the render uses the pre-sampled tables, so the parameters act through them -- the kernel returns
  d/d(local point) and d/d(local normal + scale * delta) per sample, and ``MirrorGroup.with_surface`` rebuilds the
  tables as differentiable functions of the parameters (what jax.grad through MCIntegrator.sample_group yields);
* stage >= 1 mirrors: the parameters enter the ray/surface intersection; the VJP kernel differentiates the Newton
  root implicitly and the direct dependence of the hit point and normal."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import iactrace_b200 as I
from iactrace_b200._util import replace
from iactrace_b200.core import render
from iactrace_b200.io import build_telescope
from iactrace_b200.sensors import DifferentiableSquareSensor
from oracle import trace_torch as ott
from _bridge import to_oracle_scene
from golden.cases import cfg_cassegrain

T = lambda a: torch.tensor(np.asarray(a, np.float64), dtype=ott.DT, requires_grad=True)


def _scene():
    tel = build_telescope(cfg_cassegrain(), I.MCIntegrator(24), I.random.key(0)).apply_roughness(20)
    sq = tel.sensors[0]
    tel = tel.replace_sensor(DifferentiableSquareSensor(sq.position, sq.rotation, 32, 32, (-0.5, 0.5, -0.5, 0.5),
                                                        sigma=0.8, kernel_size=2), 0)
    d = np.array([[0.002, -0.001, -1.0], [-0.004, 0.003, -1.0], [0.0, 0.0, -1.0], [0.001, 0.004, -1.0]])
    src = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    val = np.array([1.0, 0.6, 1.4, 0.8], np.float32)
    G = np.random.default_rng(5).normal(size=(32, 32))
    return tel, src, val, G


def _oracle_leaves(sc):
    g = sc["groups"][0]
    return dict(positions=T(g["positions"]), rotations=T(g["rotations"]), scale=T(g["scale"]), weights=T(g["weights"]),
                sensor_position=T(sc["sensors"][0]["position"]), sensor_rotation=T(sc["sensors"][0]["rotation"]))


def _rel(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).max() / np.abs(b).max())


def test_per_sample_point_and_normal_gradients():
    tel, src, val, G = _scene()
    g = tel.mirror_groups[0]
    pts = g.points.detach().clone().requires_grad_(True)
    nrm = g.normals.detach().clone().requires_grad_(True)
    dlt = g.perturbation_delta.detach().clone().requires_grad_(True)
    t2 = replace(tel, mirror_groups=[replace(g, points=pts, normals=nrm, perturbation_delta=dlt)] + tel.mirror_groups[1:])
    img = render(t2, src, val, "parallel", 0)
    (img * torch.tensor(G, device="cuda", dtype=torch.float32)).sum().backward()
    sc = to_oracle_scene(tel)
    lv = _oracle_leaves(sc)
    og = sc["groups"][0]
    lv.update(points=T(og["points"]), normals=T(og["normals"]), delta=T(og["delta"]))
    oimg = ott.render(sc, lv, T(src), T(val), "parallel", 0)
    (oimg * torch.tensor(G, dtype=ott.DT)).sum().backward()
    assert _rel(pts.grad.cpu().numpy(), lv["points"].grad.numpy()) < 1e-2
    assert _rel(nrm.grad.cpu().numpy(), lv["normals"].grad.numpy()) < 1e-2
    assert _rel(dlt.grad.cpu().numpy(), lv["delta"].grad.numpy()) < 1e-2


def test_secondary_surface_parameter_gradients():
    """d image / d(curvature, conic, offsets) of the stage-1 mirror through the implicit Newton root."""
    tel, src, val, G = _scene()
    sec = tel.mirror_groups[1]
    c = torch.tensor(float(sec.curvature), device="cuda", requires_grad=True)
    k = torch.tensor(float(sec.conic), device="cuda", requires_grad=True)
    off = (sec.offsets.detach() + torch.tensor([[0.05, -0.03]], device="cuda")).requires_grad_(True)
    t2 = replace(tel, mirror_groups=[tel.mirror_groups[0], sec.with_surface(c, k, off)])
    img = render(t2, src, val, "parallel", 0)
    (img * torch.tensor(G, device="cuda", dtype=torch.float32)).sum().backward()
    sc = to_oracle_scene(t2)
    sc["groups"][1]["curvature"], sc["groups"][1]["conic"] = float(c), float(k)
    lv = _oracle_leaves(sc)
    st = dict(positions=T(sc["groups"][1]["positions"]), rotations=T(sc["groups"][1]["rotations"]),
              curvature=T(float(c)), conic=T(float(k)), offsets=T(sc["groups"][1]["offsets"]))
    lv["stage"] = [st]
    oimg = ott.render(sc, lv, T(src), T(val), "parallel", 0)
    (oimg * torch.tensor(G, dtype=ott.DT)).sum().backward()
    np.testing.assert_allclose(img.detach().cpu().numpy(), oimg.detach().numpy(), rtol=5e-3, atol=2e-4 * float(oimg.max()))
    assert abs(float(c.grad) - float(st["curvature"].grad)) < 1e-2 * abs(float(st["curvature"].grad))
    assert abs(float(k.grad) - float(st["conic"].grad)) < 1e-2 * abs(float(st["conic"].grad))
    assert _rel(off.grad.cpu().numpy(), st["offsets"].grad.numpy()) < 1e-2


def test_primary_curvature_fit_through_with_surface():
    """Stage 0: loss(curvature, conic, offsets) through with_surface + render, against the same chain in float64
    (oracle render with points / normals / delta / weights rebuilt from the parameters in torch float64)."""
    tel, src, val, G = _scene()
    g = tel.mirror_groups[0]
    c = torch.tensor(float(g.curvature), device="cuda", requires_grad=True)
    k = torch.tensor(float(g.conic), device="cuda", requires_grad=True)
    off = g.offsets.detach().clone().requires_grad_(True)
    g2 = g.with_surface(c, k, off)
    # the rebuilt tables reproduce the sampler's (same parameters): forward consistency
    torch.testing.assert_close(g2.points, g.points, rtol=0, atol=2e-6)
    torch.testing.assert_close(g2.normals, g.normals, rtol=0, atol=2e-6)
    torch.testing.assert_close(g2.perturbation_delta, g.perturbation_delta, rtol=1e-4, atol=1e-5)
    torch.testing.assert_close(g2.weights, g.weights, rtol=1e-5, atol=0)
    t2 = replace(tel, mirror_groups=[g2] + tel.mirror_groups[1:])
    img = render(t2, src, val, "parallel", 0)
    (img * torch.tensor(G, device="cuda", dtype=torch.float32)).sum().backward()

    # float64 restatement of with_surface on the oracle side
    sc = to_oracle_scene(tel)
    og = sc["groups"][0]
    cc, kk, oo = T(float(c)), T(float(k)), T(og["offsets"])
    x, y = torch.tensor(og["points"][..., 0], dtype=ott.DT), torch.tensor(og["points"][..., 1], dtype=ott.DT)
    x0, y0 = oo[:, None, 0], oo[:, None, 1]
    z = ott._sag_t(x + x0, y + y0, cc, kk, og["aspheric"]) - ott._sag_t(x0, y0, cc, kk, og["aspheric"])
    sx, sy = ott._dsag_t(x + x0, y + y0, cc, kk, og["aspheric"])
    m = torch.stack([-sx, -sy, torch.ones_like(sx)], -1)
    n = m / m.norm(dim=-1, keepdim=True)
    on, od = torch.tensor(og["normals"], dtype=ott.DT), torch.tensor(og["delta"], dtype=ott.DT)

    def tangents(nn):
        ref = torch.where(nn[..., 2:3].abs() > 0.9, torch.tensor([1.0, 0.0, 0.0], dtype=ott.DT), torch.tensor([0.0, 0.0, 1.0], dtype=ott.DT))
        t1 = torch.cross(nn, ref.expand_as(nn), dim=-1)
        t1 = t1 / t1.norm(dim=-1, keepdim=True)
        return t1, torch.cross(nn, t1, dim=-1)

    t1o, t2o = tangents(on)
    th1, th2 = (od * t1o).sum(-1, keepdim=True), (od * t2o).sum(-1, keepdim=True)
    t1, t2v = tangents(n)
    lv = _oracle_leaves(sc)
    lv.update(points=torch.stack([x, y, z], -1), normals=n, delta=th1 * t1 + th2 * t2v)
    lv["weights"] = torch.tensor(og["weights"], dtype=ott.DT) * (n[..., 2:3] / on[..., 2:3])
    oimg = ott.render(sc, lv, T(src), T(val), "parallel", 0)
    (oimg * torch.tensor(G, dtype=ott.DT)).sum().backward()
    assert abs(float(c.grad) - float(cc.grad)) < 1e-2 * abs(float(cc.grad)), (float(c.grad), float(cc.grad))
    assert abs(float(k.grad) - float(kk.grad)) < 1e-2 * abs(float(kk.grad)) + 1e-6, (float(k.grad), float(kk.grad))
    assert _rel(off.grad.cpu().numpy(), oo.grad.numpy()) < 1e-2
