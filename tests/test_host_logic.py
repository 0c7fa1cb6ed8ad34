"""Host-side mirror of the reference API: YAML loader, grouping order, functional edits, sensors,
scene packing, error behaviour -- everything that needs no GPU."""
import re
from pathlib import Path

import numpy as np
import pytest
import torch
import yaml

import iactrace_b200 as I
from iactrace_b200.core import (AsphericSurface, Box, Cylinder, DiskAperture, OrientedBox, PolygonAperture, Sphere,
                                Triangle, euler_to_matrix, group_obstructions)
from iactrace_b200.io import build_telescope, load_packed_config
from iactrace_b200.io.scene_pack import pack_config, unpack_config
from iactrace_b200.sensors import HexagonalSensor, SquareSensor
from iactrace_b200.telescope import Mirror, group_mirrors
from oracle import scene as oscene, trace as otrace
from _bridge import cassegrain_config, subset_config

REF_CFG = Path("/root/reference/configs/HESS")


def test_public_names_match_reference():
    for n in ("Telescope", "Mirror", "Integrator", "MCIntegrator", "SquareSensor", "HexagonalSensor", "hexshow",
              "squareshow", "load_telescope"):
        assert hasattr(I, n)
    from iactrace_b200 import core
    for n in ("render", "render_debug", "render_response_matrix", "euler_to_matrix", "reflect", "AsphericSurface",
              "CylinderGroup", "BoxGroup", "SphereGroup", "OrientedBoxGroup", "TriangleGroup", "group_obstructions"):
        assert hasattr(core, n)


def test_packed_scenes_equal_reference_yaml():
    for name in ("CT3", "CT5"):
        cfg = load_packed_config(name)
        assert len(cfg["mirrors"]) == {"CT3": 380, "CT5": 876}[name]
        if (REF_CFG / f"{name}.yaml").exists():          # only in the build container
            ref = yaml.load(open(REF_CFG / f"{name}.yaml"), Loader=yaml.CSafeLoader)
            assert cfg == ref


def test_pack_roundtrip_synthetic(tmp_path):
    cfg = cassegrain_config()
    cfg["obstructions"].append(dict(type="triangle", v0=[0, 0, 1], v1=[1, 0, 1], v2=[0, 1, 1]))
    cfg["obstructions"].append(dict(type="oriented_box", center=[0, 0, 2], half_extents=[1, 1, 1],
                                    rotation=[[1, 0, 0], [0, 1, 0], [0, 0, 1]]))
    cfg["mirrors"][0]["aperture"] = dict(type="polygon", vertices=[[0.0, 0.0], [1.0, 0.0], [1.0, 1.0], [0.0, 1.0]])
    pack_config(cfg, tmp_path / "c.npz")
    assert unpack_config(tmp_path / "c.npz") == cfg


def test_yaml_loader_schema_and_errors(tmp_path):
    cfg = cassegrain_config()
    p = tmp_path / "cass.yaml"
    p.write_text(yaml.safe_dump(cfg))
    tel = build_telescope(yaml.safe_load(p.read_text()), None, None)
    info = tel.get_info()
    assert info["name"] == "test_cassegrain" and info["n_mirrors"] == 7 and info["optical_stages"] == [0, 1]
    assert info["mirror_types"] == ["disk", "disk"] and info["n_obstructions"] == 6 and info["sensor_types"] == ["square"]
    assert [type(g).__name__ for g in tel.obstruction_groups] == ["CylinderGroup", "BoxGroup", "SphereGroup"]
    bad = dict(cfg, mirrors=[dict(cfg["mirrors"][0], aperture=dict(type="star"))])
    with pytest.raises(ValueError, match="Unknown aperture type"):
        build_telescope(bad, None, None)
    with pytest.raises(ValueError, match="Unknown obstruction type"):
        build_telescope(dict(cfg, obstructions=[dict(type="torus")]), None, None)
    with pytest.raises(ValueError, match="Unknown sensor type"):
        build_telescope(dict(cfg, sensors=[dict(type="round")]), None, None)
    with pytest.raises(KeyError):
        build_telescope(dict(cfg, mirrors=[dict(cfg["mirrors"][0], template="nope")]), None, None)


def test_group_mirrors_order_matches_oracle():
    s1, s2 = AsphericSurface(0.05, -1.0), AsphericSurface(0.03, 0.0)
    sq = [[0, 0], [1, 0], [1, 1], [0, 1]]
    tri = [[0, 0], [1, 0], [0, 1]]
    spec = [(s1, DiskAperture(1.0), 0), (s2, PolygonAperture(sq), 0), (s2, DiskAperture(0.5), 0),
            (s1, PolygonAperture(tri), 0), (s1, DiskAperture(2.0), 1), (s1, DiskAperture(0.7), 0),
            (s2, PolygonAperture(sq), 0), (s1, PolygonAperture(sq), 0)]
    mirrors = [Mirror([i, 0, 0], [0, 0, 0], s, a, optical_stage=st) for i, (s, a, st) in enumerate(spec)]
    groups = group_mirrors(mirrors)
    omirrors = [oscene.make_mirror([i, 0, 0], [0, 0, 0], s.curvature, s.conic, [],
                                   ("disk", a.radius) if isinstance(a, DiskAperture) else ("polygon", a.vertices), st)
                for i, (s, a, st) in enumerate(spec)]
    ogroups = oscene.group_mirrors(omirrors)
    assert len(groups) == len(ogroups) == 6
    for g, og in zip(groups, ogroups):
        assert g.kind == og["kind"] and g.optical_stage == og["stage"]
        assert g.positions[:, 0].tolist() == og["positions"][:, 0].tolist()
    # stage ascending; disk groups before polygon groups; insertion order inside
    assert [(g.optical_stage, g.kind) for g in groups] == [(0, "disk"), (0, "disk"), (0, "polygon"), (0, "polygon"),
                                                           (0, "polygon"), (1, "disk")]
    assert groups[0].positions[:, 0].tolist() == [0.0, 5.0]
    assert group_mirrors([]) == []


def test_group_obstructions_order():
    obs = [Triangle([0, 0, 0], [1, 0, 0], [0, 1, 0]), Sphere([0, 0, 0], 1), Cylinder([0, 0, 0], [0, 0, 1], 0.1),
           OrientedBox([0, 0, 0], [1, 1, 1], np.eye(3)), Box([0, 0, 0], [1, 1, 1]), Cylinder([1, 0, 0], [1, 0, 1], 0.2)]
    gs = group_obstructions(obs)
    assert [type(g).__name__ for g in gs] == ["CylinderGroup", "BoxGroup", "SphereGroup", "OrientedBoxGroup", "TriangleGroup"]
    assert len(gs[0]) == 2 and gs[0].r.tolist() == pytest.approx([0.1, 0.2])
    assert group_obstructions([]) == []


def test_sensors_match_oracle_statics():
    cfg = load_packed_config("CT5")
    tel = build_telescope(cfg, None, None)
    osens = oscene.parse_config(cfg)[3]
    for s, o in zip(tel.sensors, osens):
        if isinstance(s, HexagonalSensor):
            assert s.hex_size == o["hex_size"] and s.hex_inradius == o["hex_inradius"]
            assert s.grid_rotation == o["grid_rotation"] and s.grid_offset == o["grid_offset"]
            assert (s.q_min, s.r_min) == (o["q_min"], o["r_min"])
            assert np.array_equal(s.lookup_table.cpu().numpy(), o["lookup_table"])
            assert s.get_accumulator_shape() == (o["n_pixels"],)
        else:
            assert (s.x0, s.y0, s.dx, s.dy) == (o["x0"], o["y0"], o["dx"], o["dy"])
            assert s.get_accumulator_shape() == (o["height"], o["width"]) == (1431, 1501)
    # pinned grid constants are accepted back verbatim
    h = tel.sensors[0]
    h2 = HexagonalSensor(h.position, h.rotation, h.hex_centers, h.edge_width, grid=h.grid_constants())
    assert h2.grid_constants()["grid_rotation"] == h.grid_rotation


def test_euler_to_matrix_matches_oracle():
    rng = np.random.default_rng(0)
    for ang in rng.uniform(-180, 180, (8, 3)).astype(np.float32):
        R = euler_to_matrix(ang).cpu().numpy()
        np.testing.assert_allclose(R, otrace.euler_to_matrix(ang, np.float64), atol=3e-7)
        np.testing.assert_allclose(R @ R.T, np.eye(3), atol=1e-6)
    # tip rotates about x: z-axis tilts towards -y
    R = euler_to_matrix([90.0, 0.0, 0.0]).cpu().numpy()
    np.testing.assert_allclose(R @ np.array([0, 0, 1.0]), [0, -1, 0], atol=1e-6)


def test_functional_edits_leave_the_original_untouched():
    tel = build_telescope(subset_config(load_packed_config("CT3"), n_mirrors=5), None, None)
    g0 = tel.mirror_groups[0]
    t2 = tel.apply_roughness(24)
    assert float(g0.perturbation_scale.abs().sum()) == 0.0
    assert float(t2.mirror_groups[0].perturbation_scale[0]) == pytest.approx(24 * np.pi / 648000, rel=1e-6)
    t3 = tel.set_mirror_rotations(0, torch.zeros(5, 3)).set_mirror_positions(0, torch.ones(5, 3))
    assert float(t3.mirror_groups[0].rotations.abs().sum()) == 0.0 and float(g0.rotations.abs().sum()) > 0
    assert float(t3.mirror_groups[0].positions.sum()) == 15.0
    t4 = tel.focus(0.01, 0)
    assert float(t4.sensors[0].position[2]) == pytest.approx(15.038, abs=1e-5) and float(tel.sensors[0].position[2]) == pytest.approx(15.028)
    assert tel.clear_obstructions().get_obstruction_count() == 0 and tel.get_obstruction_count() == 33
    assert tel.remove_sensor(0).get_sensor_count() == 1 and tel.add_sensor(tel.sensors[0]).get_sensor_count() == 3
    with pytest.raises(IndexError):
        tel.replace_sensor(tel.sensors[0], 7)
    with pytest.raises(IndexError):
        tel.remove_obstruction(3)
    with pytest.raises(IndexError):
        tel.clear_obstructions().remove_obstruction(0)
    c = tel.clone()
    assert c.mirror_groups[0].positions.data_ptr() != g0.positions.data_ptr()
    assert torch.equal(c.mirror_groups[0].positions, g0.positions)
    assert tel.get_mirrors_by_stage(0) == [0] and tel.get_mirror_count() == 5


def test_no_cpu_fallback():
    """Without a CUDA device the compute path must fail loudly, never fall back."""
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    cfg = subset_config(load_packed_config("CT3"), n_mirrors=2)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA extension"):
        build_telescope(cfg, I.MCIntegrator(4), None)
    tel = build_telescope(cfg, None, None)
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA extension"):
        tel(np.zeros((1, 3), np.float32), np.ones(1, np.float32))
    with pytest.raises(RuntimeError):
        I.random.normal(I.random.key(0), 4)


def test_product_never_imports_oracle():
    root = Path(__file__).resolve().parent.parent / "iactrace_b200"
    for p in root.rglob("*.py"):
        assert not re.search(r"^\s*(from|import)\s+oracle\b", p.read_text(), re.M), p
    for p in list(root.rglob("*.cu")) + list(root.rglob("*.cuh")):
        assert "oracle" not in p.read_text().lower(), p


def test_public_intersection_functions_match_reference_goldens():
    """iactrace_b200.core.intersect_* (torch API mirrors) against vectors from the executed reference."""
    from iactrace_b200 import core as C
    G = np.load(Path(__file__).parent / "golden" / "reference_golden.npz")
    o, d = torch.tensor(G["unit/o"]), torch.tensor(G["unit/d"])
    f = lambda *a: torch.tensor(a, dtype=torch.float32)

    def cmp(t, key):
        t, g = t.cpu().numpy(), G[key]
        assert np.array_equal(np.isfinite(t), np.isfinite(g)), key
        m = np.isfinite(g)
        np.testing.assert_allclose(t[m], g[m], rtol=2e-5, err_msg=key)

    cmp(C.intersect_cylinder(o, d, f(-1, 0.5, 2), f(2, -0.5, 4), 0.8), "unit/cylinder")
    cmp(C.intersect_box(o, d, f(-1, -2, 1), f(1.5, 0.5, 3)), "unit/box")
    cmp(C.intersect_sphere(o, d, f(0.5, 0.5, 3), 1.7), "unit/sphere")
    th = np.deg2rad(30.0)
    Rz = torch.tensor([[np.cos(th), -np.sin(th), 0], [np.sin(th), np.cos(th), 0], [0, 0, 1]], dtype=torch.float32)
    cmp(C.intersect_oriented_box(o, d, f(0.5, 0, 3), f(1.5, 0.6, 1.0), Rz), "unit/obox")
    cmp(C.intersect_triangle(o, d, f(-4, -4, 3), f(4, -3, 3.5), f(0, 4, 2.5)), "unit/triangle")
    cmp(C.intersect_conic(torch.tensor(G["unit/surf_o"]), torch.tensor(G["unit/surf_d"]), 0.05, -1.0), "unit/conic_t")
    p = C.intersect_plane(o, d, f(0.1, -0.2, 5), C.euler_to_matrix([3.0, -2.0, 20.0])).cpu().numpy()
    g = G["unit/plane"]
    assert np.array_equal(p[:, 0] > 1e9, g[:, 0] > 1e9)
    np.testing.assert_allclose(p[g[:, 0] < 1e9], g[g[:, 0] < 1e9], rtol=1e-4, atol=1e-5)


def test_config_builder_and_soft_sensor_yaml(tmp_path, capsys):
    from iactrace_b200.io import TelescopeConfigBuilder, build_telescope
    from iactrace_b200.sensors import DifferentiableHexagonalSensor, DifferentiableSquareSensor
    from iactrace_b200.utils import show_structure, trainable
    b = (TelescopeConfigBuilder("demo").add_mirror_template("m", 0.05, -1.0, [])
         .add_mirror_circular("A", "m", [2, 0, 0], [0, 0, 0], 1.0, offset=[2, 0])
         .add_mirror_polygon("B", "m", [0, 2, 0], [0, 0, 0], [[-0.5, -0.5], [0.5, -0.5], [0.5, 0.5], [-0.5, 0.5]])
         .add_mirror_circular("S", "m", [0, 0, 6], [180, 0, 0], 1.0, stage=1)
         .add_obstruction_cylinder("c", [0, 0, 1], [0, 0, 2], 0.1).add_obstruction_box("b", [1, 1, 1], [2, 2, 2])
         .add_obstruction("s", "sphere", center=[0, 1, 3], r=0.2)
         .add_obstruction("t", "triangle", v0=[0, 0, 4], v1=[1, 0, 4], v2=[0, 1, 4])
         .add_square_sensor_array("sq", [0, 0, -1], [0, 0, 0], 8, 4, [-1, 1, -0.5, 0.5], edge_width=0.01)
         .add_square_sensor_array("soft", [0, 0, -1], [0, 0, 0], 8, 4, [-1, 1, -0.5, 0.5], soft=dict(sigma=0.3, kernel_size=1))
         .add_hexagon_sensor_array("hx", [0, 0, -1], [0, 0, 0], [0.0, 0.1, 0.05, -0.05], [0.0, 0.0, 0.0866, 0.0866],
                                   soft=dict(sigma=0.5, kernel_size=1)))
    with pytest.raises(KeyError):
        b.add_mirror_circular("X", "nope", [0, 0, 0], [0, 0, 0], 1.0)
    with pytest.raises(ValueError):
        b.add_obstruction("z", "torus", r=1)
    path = b.save(tmp_path / "demo.yaml", precision=6)
    tel = build_telescope(yaml.safe_load(path.read_text()), None, None)
    assert tel.name == "demo" and [g.optical_stage for g in tel.mirror_groups] == [0, 0, 1]
    assert [g.kind for g in tel.mirror_groups] == ["disk", "polygon", "disk"]
    assert isinstance(tel.sensors[1], DifferentiableSquareSensor) and tel.sensors[1].sigma == 0.3
    assert isinstance(tel.sensors[2], DifferentiableHexagonalSensor) and tel.sensors[2].n_pixels == 4
    assert tel.sensors[0].edge_width == 0.01 and tel.get_obstruction_count() == 4
    show_structure(tel)
    out = capsys.readouterr().out
    assert "mirror_groups.0.rotations: (1, 3) float32" in out and "sensors.0.position" in out
    leaves = trainable(tel, "mirror_groups.*.rotations")
    assert len(leaves) == 3 and all(t.requires_grad for _, t in leaves)
