"""The gradient oracle (torch f64 autograd restatement): its forward equals the NumPy oracle and its
gradients equal central finite differences of the NumPy oracle (soft sensor => smooth image)."""
import numpy as np
import pytest
import torch

from iactrace_b200.io import load_packed_config
from oracle import prng, scene as oscene, trace as otrace, trace_torch as ott
from _bridge import subset_config, point_grid


def _setup(soft):
    sc = oscene.build_scene(subset_config(load_packed_config("CT5"), mirror_step=110), 12, prng.key(0))
    sc = oscene.apply_roughness(sc, 30)
    if soft:
        sc = dict(sc, sensors=[oscene.make_soft_hex_sensor(sc["sensors"][0], 0.5, 1)] + sc["sensors"][1:])
    g = sc["groups"][0]
    leaves = dict(positions=torch.tensor(g["positions"], dtype=ott.DT), rotations=torch.tensor(g["rotations"], dtype=ott.DT),
                  scale=torch.tensor(g["scale"], dtype=ott.DT), weights=torch.tensor(g["weights"], dtype=ott.DT),
                  sensor_position=torch.tensor(sc["sensors"][0]["position"], dtype=ott.DT),
                  sensor_rotation=torch.tensor(sc["sensors"][0]["rotation"], dtype=ott.DT))
    src = point_grid(2, 1.0)
    val = np.array([1.0, 0.7, 1.3, 0.9])
    return sc, leaves, src, val


@pytest.mark.parametrize("soft", [False, True])
def test_torch_forward_equals_numpy_oracle(soft):
    sc, leaves, src, val = _setup(soft)
    img = ott.render(sc, leaves, torch.tensor(src, dtype=ott.DT), torch.tensor(val, dtype=ott.DT), "point", 0)
    ref = otrace.render(sc, src, val, "point", 0, np.float64)
    np.testing.assert_allclose(img.numpy(), ref, rtol=1e-9, atol=1e-12)


def test_autograd_matches_finite_differences():
    sc, leaves, src, val = _setup(True)
    rng = np.random.default_rng(0)
    G = rng.normal(size=sc["sensors"][0]["n_pixels"])
    for k in ("rotations", "positions", "scale"):
        leaves[k].requires_grad_(True)
    img = ott.render(sc, leaves, torch.tensor(src, dtype=ott.DT), torch.tensor(val, dtype=ott.DT), "point", 0)
    (img * torch.tensor(G)).sum().backward()

    def loss(rot=None, pos=None):
        g = dict(sc["groups"][0])
        if rot is not None:
            g["rotations"] = rot
        if pos is not None:
            g["positions"] = pos
        return float((otrace.render(dict(sc, groups=[g]), src, val, "point", 0, np.float64) * G).sum())

    base_rot = sc["groups"][0]["rotations"].astype(np.float64)
    base_pos = sc["groups"][0]["positions"].astype(np.float64)
    for (f, k) in ((0, 0), (3, 1), (5, 0)):
        h = 1e-5
        rp, rm = base_rot.copy(), base_rot.copy()
        rp[f, k] += h; rm[f, k] -= h
        fd = (loss(rot=rp) - loss(rot=rm)) / (2 * h)
        assert abs(fd - float(leaves["rotations"].grad[f, k])) <= 2e-4 * max(1.0, abs(fd)), (f, k, fd)
    for (f, k) in ((1, 2), (4, 0)):
        h = 1e-6
        pp, pm = base_pos.copy(), base_pos.copy()
        pp[f, k] += h; pm[f, k] -= h
        fd = (loss(pos=pp) - loss(pos=pm)) / (2 * h)
        assert abs(fd - float(leaves["positions"].grad[f, k])) <= 2e-4 * max(1.0, abs(fd)), (f, k, fd)


def _cassegrain_setup():
    from golden.cases import cfg_cassegrain
    sc = oscene.build_scene(cfg_cassegrain(), 10, prng.key(0))
    sq = sc["sensors"][0]
    soft = oscene.make_soft_square_sensor(sq["position"], sq["rotation"], 32, 32, (-0.5, 0.5, -0.5, 0.5), 0.8, 2)
    sc = dict(sc, sensors=[soft])
    g = sc["groups"][0]
    leaves = dict(positions=torch.tensor(g["positions"], dtype=ott.DT), rotations=torch.tensor(g["rotations"], dtype=ott.DT),
                  scale=torch.tensor(g["scale"], dtype=ott.DT), weights=torch.tensor(g["weights"], dtype=ott.DT),
                  sensor_position=torch.tensor(soft["position"], dtype=ott.DT),
                  sensor_rotation=torch.tensor(soft["rotation"], dtype=ott.DT))
    d = np.array([[0.002, -0.001, -1.0], [-0.004, 0.003, -1.0], [0.0, 0.0, -1.0]])
    src = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    return sc, leaves, src, np.array([1.0, 0.6, 1.4])


def test_two_stage_forward_and_gradients():
    """Cassegrain: the torch restatement through the secondary equals the NumPy oracle, and its
    gradients (implicit differentiation of the Newton root) equal finite differences."""
    sc, leaves, src, val = _cassegrain_setup()
    for k in ("rotations", "positions"):
        leaves[k].requires_grad_(True)
    img = ott.render(sc, leaves, torch.tensor(src, dtype=ott.DT), torch.tensor(val, dtype=ott.DT), "parallel", 0)
    ref = otrace.render(sc, src, val, "parallel", 0, np.float64)
    assert ref.sum() > 0.5
    np.testing.assert_allclose(img.detach().numpy(), ref, rtol=1e-8, atol=1e-11)
    G = np.random.default_rng(3).normal(size=ref.shape)
    (img * torch.tensor(G)).sum().backward()

    def loss(rot=None, pos=None):
        g = dict(sc["groups"][0])
        if rot is not None:
            g["rotations"] = rot
        if pos is not None:
            g["positions"] = pos
        return float((otrace.render(dict(sc, groups=[g] + sc["groups"][1:]), src, val, "parallel", 0, np.float64) * G).sum())

    base_rot = sc["groups"][0]["rotations"].astype(np.float64)
    base_pos = sc["groups"][0]["positions"].astype(np.float64)
    for (f, k) in ((0, 0), (2, 1), (4, 2)):
        h = 2e-6
        rp, rm = base_rot.copy(), base_rot.copy()
        rp[f, k] += h; rm[f, k] -= h
        fd = (loss(rot=rp) - loss(rot=rm)) / (2 * h)
        got = float(leaves["rotations"].grad[f, k])
        assert abs(fd - got) <= 1e-3 * max(1.0, abs(fd)), ("rot", f, k, fd, got)
    for (f, k) in ((1, 2), (3, 0)):
        h = 1e-6
        pp, pm = base_pos.copy(), base_pos.copy()
        pp[f, k] += h; pm[f, k] -= h
        fd = (loss(pos=pp) - loss(pos=pm)) / (2 * h)
        got = float(leaves["positions"].grad[f, k])
        assert abs(fd - got) <= 1e-3 * max(1.0, abs(fd)), ("pos", f, k, fd, got)
