"""K6: the CUDA vector-Jacobian product (through torch.autograd over the C ABI) against reverse-mode
autodiff of the float64 oracle -- i.e. against what jax.grad of the reference's render would give."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

import iactrace_b200 as I
from iactrace_b200.core import render
from iactrace_b200.io import build_telescope, load_packed_config
from iactrace_b200.sensors import DifferentiableHexagonalSensor, DifferentiableSquareSensor
from iactrace_b200.telescope import operations as ops
from oracle import trace_torch as ott
from _bridge import to_oracle_scene, subset_config, point_grid, parallel_grid

LEAVES = ("positions", "rotations", "scale", "weights", "sensor_position", "sensor_rotation", "sources", "values")


def _make(sensor_kind, n_samples=12, step=110):
    cfg = subset_config(load_packed_config("CT5"), mirror_step=step)
    tel = build_telescope(cfg, I.MCIntegrator(n_samples), I.random.key(0)).apply_roughness(30)
    hard = tel.sensors[0]
    if sensor_kind == "soft_hex":
        tel = tel.replace_sensor(DifferentiableHexagonalSensor(hard.position, hard.rotation, hard.hex_centers, 0.5, 1,
                                                               grid=hard.grid_constants()), 0)
    elif sensor_kind == "soft_square":
        tel = tel.replace_sensor(DifferentiableSquareSensor(hard.position, hard.rotation, 64, 64, (-0.9, 0.9, -0.9, 0.9),
                                                            sigma=0.8, kernel_size=2), 0)
    return tel


def _grads_cuda(tel, src, val, stype, G):
    g = tel.mirror_groups[0]
    leaf = dict(positions=g.positions.detach().clone().requires_grad_(True),
                rotations=g.rotations.detach().clone().requires_grad_(True),
                scale=g.perturbation_scale.detach().clone().requires_grad_(True),
                weights=g.weights.detach().clone().requires_grad_(True),
                sensor_position=tel.sensors[0].position.detach().clone().requires_grad_(True),
                sensor_rotation=tel.sensors[0].rotation.detach().clone().requires_grad_(True),
                sources=torch.tensor(src, device="cuda", requires_grad=True),
                values=torch.tensor(val, device="cuda", requires_grad=True))
    from iactrace_b200._util import replace
    g2 = replace(g, positions=leaf["positions"], rotations=leaf["rotations"], perturbation_scale=leaf["scale"],
                 weights=leaf["weights"])
    s2 = replace(tel.sensors[0], position=leaf["sensor_position"], rotation=leaf["sensor_rotation"])
    later = []
    for i, gl in enumerate(tel.mirror_groups[1:]):
        leaf[f"stage{i}_positions"] = gl.positions.detach().clone().requires_grad_(True)
        leaf[f"stage{i}_rotations"] = gl.rotations.detach().clone().requires_grad_(True)
        later.append(replace(gl, positions=leaf[f"stage{i}_positions"], rotations=leaf[f"stage{i}_rotations"]))
    tel2 = replace(tel, mirror_groups=[g2] + later, sensors=[s2] + tel.sensors[1:])
    img = render(tel2, leaf["sources"], leaf["values"], stype, 0)
    assert img.requires_grad
    (img * torch.tensor(G, device="cuda", dtype=torch.float32).reshape(img.shape)).sum().backward()
    return img.detach().cpu().numpy(), {k: v.grad.detach().cpu().numpy().astype(np.float64) for k, v in leaf.items()}


def _grads_oracle(tel, src, val, stype, G):
    sc = to_oracle_scene(tel)
    g = sc["groups"][0]
    T = lambda a: torch.tensor(np.asarray(a, np.float64), dtype=ott.DT, requires_grad=True)
    leaves = dict(positions=T(g["positions"]), rotations=T(g["rotations"]), scale=T(g["scale"]), weights=T(g["weights"]),
                  sensor_position=T(sc["sensors"][0]["position"]), sensor_rotation=T(sc["sensors"][0]["rotation"]))
    stage = [dict(positions=T(gl["positions"]), rotations=T(gl["rotations"])) for gl in sc["groups"][1:]]
    leaves["stage"] = stage
    s, v = T(src), T(val)
    img = ott.render(sc, leaves, s, v, stype, 0)
    (img * torch.tensor(G, dtype=ott.DT).reshape(img.shape)).sum().backward()
    gnp = lambda t: t.grad.numpy() if t.grad is not None else np.zeros(tuple(t.shape))
    out = {k: gnp(t) for k, t in leaves.items() if k != "stage"}
    out.update(sources=gnp(s), values=gnp(v))
    for i, st in enumerate(stage):
        out[f"stage{i}_positions"], out[f"stage{i}_rotations"] = gnp(st["positions"]), gnp(st["rotations"])
    return img.detach().numpy(), out


def _check(got, want, names, rtol):
    for k in names:
        a, b = got[k], want[k].reshape(got[k].shape)
        scale = np.abs(b).max()
        assert scale > 0, k
        err = np.abs(a - b).max() / scale
        assert err < rtol, f"{k}: max err / max |grad| = {err:.3e}"


@pytest.mark.parametrize("stype", ["point", "parallel"])
def test_hard_hex_sensor_value_path_gradients(stype):
    """Hard sensors: only d(value) flows (pixel index is piecewise constant)."""
    tel = _make("hard")
    src = point_grid(2, 1.0) if stype == "point" else parallel_grid(2, 2.0)
    val = np.array([1.0, 0.7, 1.3, 0.9], np.float32)
    G = np.random.default_rng(1).normal(size=tel.sensors[0].n_pixels)
    img, got = _grads_cuda(tel, src, val, stype, G)
    oimg, want = _grads_oracle(tel, src, val, stype, G)
    assert abs(img.sum() - oimg.sum()) < 1e-3 * oimg.sum()
    names = ["rotations", "scale", "weights", "values"] + (["sources"] if stype == "parallel" else [])
    _check(got, want, names, 2e-3)
    # a far point source / a fixed sensor pose receive (numerically) no gradient through the value path
    assert np.abs(got["sensor_position"]).max() == 0.0


@pytest.mark.parametrize("kind,stype", [("soft_hex", "point"), ("soft_hex", "parallel"), ("soft_square", "parallel")])
def test_soft_sensor_full_gradients(kind, stype):
    tel = _make(kind)
    src = point_grid(2, 1.0) if stype == "point" else parallel_grid(2, 2.0)
    val = np.array([1.0, 0.7, 1.3, 0.9], np.float32)
    shape = tel.sensors[0].get_accumulator_shape()
    G = np.random.default_rng(2).normal(size=shape)
    img, got = _grads_cuda(tel, src, val, stype, G)
    oimg, want = _grads_oracle(tel, src, val, stype, G)
    np.testing.assert_allclose(img, oimg, rtol=5e-3, atol=1e-4 * oimg.max())
    names = ["rotations", "positions", "scale", "weights", "values", "sensor_position", "sensor_rotation"]
    if stype == "parallel":
        names.append("sources")
    # f32 kernel vs f64 autodiff; a ray within rounding noise of a hex boundary changes its tap set: 1 %
    _check(got, want, names, 1e-2)


def test_alignment_fit_loss_decreases():
    """BASELINE config 5 in miniature: gradient of 1/2 |img(theta) - img(theta*)|^2 w.r.t. facet tip/tilt."""
    tel = _make("soft_hex", n_samples=24, step=60)
    src = point_grid(2, 0.8)
    val = np.ones(4, np.float32)
    target = render(ops.apply_misalignment_to_group(tel, 0, 15, 10, I.random.key(4242)), src, val, "point", 0)
    g = tel.mirror_groups[0]
    rot = g.rotations.detach().clone().requires_grad_(True)
    from iactrace_b200._util import replace

    def loss_of(r):
        t = replace(tel, mirror_groups=[replace(g, rotations=r)])
        return 0.5 * ((render(t, src, val, "point", 0) - target) ** 2).sum()

    l0 = loss_of(rot)
    l0.backward()
    grad = rot.grad.clone()
    assert torch.isfinite(grad).all() and float(grad[:, :2].abs().max()) > 0
    with torch.no_grad():
        step = 1e-3 * float(l0) / float((grad ** 2).sum())
        l1 = loss_of(rot - step * grad)
    assert float(l1) < float(l0)


@pytest.mark.parametrize("polygon_secondary", [False, True])
def test_two_stage_gradients_through_the_secondary(polygon_secondary):
    """Cassegrain (BASELINE config 3 geometry): the VJP differentiates through the secondary mirror's
    Newton intersection (implicitly) and the second reflection; compared with f64 autodiff."""
    from golden.cases import cfg_cassegrain, cfg_polygon_secondary
    cfg = cfg_polygon_secondary() if polygon_secondary else cfg_cassegrain()
    tel = build_telescope(cfg, I.MCIntegrator(16), I.random.key(0)).apply_roughness(20)
    sq = tel.sensors[0]
    tel = tel.replace_sensor(DifferentiableSquareSensor(sq.position, sq.rotation, 32, 32, (-0.5, 0.5, -0.5, 0.5),
                                                        sigma=0.8, kernel_size=2), 0)
    d = np.array([[0.002, -0.001, -1.0], [-0.004, 0.003, -1.0], [0.0, 0.0, -1.0], [0.001, 0.004, -1.0]])
    src = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    val = np.array([1.0, 0.6, 1.4, 0.8], np.float32)
    G = np.random.default_rng(5).normal(size=(32, 32))
    img, got = _grads_cuda(tel, src, val, "parallel", G)
    oimg, want = _grads_oracle(tel, src, val, "parallel", G)
    assert oimg.sum() > 1.0
    np.testing.assert_allclose(img, oimg, rtol=5e-3, atol=2e-4 * oimg.max())
    _check(got, want, ["rotations", "positions", "scale", "weights", "values", "sources", "sensor_position",
                       "sensor_rotation", "stage0_positions", "stage0_rotations"], 1e-2)


def test_in_place_optimizer_steps_reach_the_kernels():
    """The alignment-fit loop: render -> backward -> optimizer.step() edits the leaves IN PLACE.  The stage >= 1
    record table, the obstruction tables and the world table are cached on the Telescope keyed on (storage, version),
    so the second render must see the moved secondary and the gradient must match the oracle at the NEW pose."""
    from golden.cases import cfg_cassegrain
    tel = build_telescope(cfg_cassegrain(), I.MCIntegrator(16), I.random.key(0)).apply_roughness(20)
    sq = tel.sensors[0]
    tel = tel.replace_sensor(DifferentiableSquareSensor(sq.position, sq.rotation, 32, 32, (-0.5, 0.5, -0.5, 0.5),
                                                        sigma=0.8, kernel_size=2), 0)
    d = np.array([[0.002, -0.001, -1.0], [-0.004, 0.003, -1.0], [0.0, 0.0, -1.0]])
    src = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    val = np.array([1.0, 0.6, 1.4], np.float32)
    G = torch.tensor(np.random.default_rng(5).normal(size=(32, 32)), device="cuda", dtype=torch.float32)
    sec = tel.mirror_groups[1]
    sec.positions.requires_grad_(True)
    sec.rotations.requires_grad_(True)
    prim = tel.mirror_groups[0]
    prim.rotations.requires_grad_(True)
    opt = torch.optim.SGD([sec.positions, sec.rotations, prim.rotations], lr=1.0)
    imgs = []
    for it in range(3):
        opt.zero_grad()
        img = render(tel, src, val, "parallel", 0)
        (img * G).sum().backward()
        imgs.append(img.detach().clone())
        # oracle gradient at the pose this render used
        _, want = _grads_oracle(tel, src, val, "parallel", G.cpu().numpy())
        got = dict(stage0_positions=sec.positions.grad.cpu().numpy().astype(np.float64),
                   stage0_rotations=sec.rotations.grad.cpu().numpy().astype(np.float64),
                   rotations=prim.rotations.grad.cpu().numpy().astype(np.float64))
        _check(got, want, ["stage0_positions", "stage0_rotations", "rotations"], 1e-2)
        with torch.no_grad():                           # a visible in-place move: 2 mm of despace, 0.01 deg of tilt
            sec.positions.grad.copy_(torch.tensor([[0.0, 0.0, -2e-3]], device="cuda"))
            sec.rotations.grad.copy_(torch.tensor([[-1e-2, 0.0, 0.0]], device="cuda"))
            prim.rotations.grad.mul_(0.0).add_(1e-3)
        opt.step()
    assert float((imgs[1] - imgs[0]).abs().max()) > 1e-3 * float(imgs[0].max())
    assert float((imgs[2] - imgs[1]).abs().max()) > 1e-3 * float(imgs[1].max())
    # obstruction tables follow in-place edits too
    from iactrace_b200.core import Sphere, group_obstructions
    t2 = I.Telescope(tel.mirror_groups, group_obstructions([Sphere([50.0, 0.0, 3.0], 0.5)]), tel.sensors)
    with torch.no_grad():
        a = render(t2, src, val, "parallel", 0)
        t2.obstruction_groups[0].centers.copy_(torch.tensor([[2.0, 0.0, 3.0]], device="cuda"))   # now above a segment
        b = render(t2, src, val, "parallel", 0)
    assert float(b.sum()) < 0.99 * float(a.sum())


def test_functional_edits_stay_differentiable():
    """apply_misalignment / apply_displacement / focus are out-of-place like the reference's `.at[].add()`: a fit
    through an edited telescope still reaches the original leaves."""
    tel = _make("soft_hex", n_samples=8, step=150)
    g = tel.mirror_groups[0]
    g.rotations.requires_grad_(True)
    g.positions.requires_grad_(True)
    tel.sensors[0].position.requires_grad_(True)
    t2 = ops.apply_displacement_to_group(ops.apply_misalignment_to_group(tel, 0, 15, 10, I.random.key(1)), 0, 1e-3, I.random.key(2))
    t2 = ops.focus(t2, 0.01, 0)
    src, val = point_grid(2, 0.8), np.ones(4, np.float32)
    render(t2, src, val, "point", 0).square().sum().backward()
    for t in (g.rotations, g.positions, tel.sensors[0].position):
        assert t.grad is not None and float(t.grad.abs().max()) > 0
