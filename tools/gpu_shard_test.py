"""Device time of one render as a function of the source count (strong-scaling shards), per library variant."""
import os, sys, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import iactrace_b200 as I
from iactrace_b200.core import render, render_response_matrix
from iactrace_b200.io import build_telescope, load_packed_config
from iactrace_b200.workloads import point_grid, parallel_grid
tel = build_telescope(load_packed_config("CT5"), I.MCIntegrator(115), I.random.key(0))
full = torch.from_numpy(point_grid(64, 1.5)).cuda()
def t(fn):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(100): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 10
out = []
for S in (64, 512, 4096):
    src = full[:: 4096 // S].contiguous(); val = torch.ones(S, device="cuda")
    out.append(f"S={S}: {t(lambda: render(tel, src, val, 'point', 0)):8.1f} us")
ct3 = build_telescope(load_packed_config("CT3"), I.MCIntegrator(1000), I.random.key(42)).apply_roughness(24)
d = torch.from_numpy(parallel_grid(64, 5.5)).cuda()
for S in (512, 4096):
    src = d[:: 4096 // S].contiguous(); val = torch.ones(S, device="cuda")
    out.append(f"matrix M1000 S={S}: {t(lambda: render_response_matrix(ct3, src, val, 'parallel', 0)):8.1f} us")
print(os.environ.get("IACTRACE_B200_LIB", "product")[-12:], " | ".join(out))
