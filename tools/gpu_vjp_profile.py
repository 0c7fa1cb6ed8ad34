"""One forward + backward of BASELINE config 5 (CT5, soft hex sensor, 4096 sources, M = 115) for ncu."""
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
from iactrace_b200.workloads import point_grid

tel = build_telescope(load_packed_config("CT5"), I.MCIntegrator(115), I.random.key(0))
hard = tel.sensors[0]
tel = tel.replace_sensor(DifferentiableHexagonalSensor(hard.position, hard.rotation, hard.hex_centers, 0.5, 1, grid=hard.grid_constants()), 0)
src = torch.from_numpy(point_grid(64, 1.5)).cuda()
val = torch.ones(len(src), device="cuda")
target = render(tel.apply_misalignment_to_group(0, 15, 10, I.random.key(4242)), src, val, "point", 0)
g = tel.mirror_groups[0]
rot = g.rotations.detach().clone().requires_grad_(True)
for _ in range(3):
    rot.grad = None
    loss = 0.5 * ((render(replace(tel, mirror_groups=[replace(g, rotations=rot)]), src, val, "point", 0) - target) ** 2).sum()
    loss.backward()
torch.cuda.synchronize()
print("loss", float(loss), "grad max", float(rot.grad.abs().max()))
