"""BASELINE config 5 timing: CT5 + soft hex sensor, loss + gradient w.r.t. facet rotations (not a bench contract)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import iactrace_b200 as I
from iactrace_b200._util import replace
from iactrace_b200.core import render
from iactrace_b200.io import build_telescope, load_packed_config
from iactrace_b200.sensors import DifferentiableHexagonalSensor
from iactrace_b200.workloads import cassegrain_config, point_grid, star_field


def timeit(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


def main():
    ct5 = load_packed_config("CT5")
    tel = build_telescope(ct5, I.MCIntegrator(115), I.random.key(0))
    hard = tel.sensors[0]
    tel = tel.replace_sensor(DifferentiableHexagonalSensor(hard.position, hard.rotation, hard.hex_centers, 0.5, 1,
                                                           grid=hard.grid_constants()), 0)
    src = torch.from_numpy(point_grid(64, 1.5)).cuda(); val = torch.ones(len(src), device="cuda")
    target = render(tel.apply_misalignment_to_group(0, 15, 10, I.random.key(4242)), src, val, "point", 0)
    g = tel.mirror_groups[0]
    rot = g.rotations.detach().clone().requires_grad_(True)

    def fwd():
        with torch.no_grad():
            render(tel, src, val, "point", 0)

    def fwd_bwd():
        rot.grad = None
        t = replace(tel, mirror_groups=[replace(g, rotations=rot)])
        loss = 0.5 * ((render(t, src, val, "point", 0) - target) ** 2).sum()
        loss.backward()
    a, b = timeit(fwd), timeit(fwd_bwd)
    print(f"config5 CT5 soft-hex S=4096 M=115: forward {a:.2f} ms, loss+gradient {b:.2f} ms (VJP ~{b - a:.2f} ms), |grad|max {float(rot.grad.abs().max()):.4g}")
    # Cassegrain: gradient w.r.t. the secondary's pose (stage path of the VJP)
    tel = build_telescope(cassegrain_config(True), I.MCIntegrator(4096), I.random.key(0))
    d, flux = star_field(256)
    src = torch.from_numpy(d).cuda(); val = torch.from_numpy(flux).cuda()
    sec = next((gq for gq in tel.mirror_groups if gq.optical_stage == 1), None)
    if sec is not None:
        pos = sec.positions.detach().clone().requires_grad_(True)
        groups = [replace(gq, positions=pos) if gq is sec else gq for gq in tel.mirror_groups]

        def cass():
            pos.grad = None
            t = replace(tel, mirror_groups=groups)
            (render(t, src, val, "parallel", 0) ** 2).sum().backward()
        c = timeit(cass)
        print(f"cassegrain S=256 M=4096 (6.3e6 rays): forward+VJP {c:.2f} ms")


if __name__ == "__main__":
    main()
