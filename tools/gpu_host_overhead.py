#!/usr/bin/env python
"""Host (enqueue) time vs device time of one render call as a function of the source count, and the same step
replayed from a CUDA graph.  Explains the strong-scaling limiter of small per-GPU jobs."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import iactrace_b200 as I
from iactrace_b200.core import render
from iactrace_b200.io import build_telescope, load_packed_config
from iactrace_b200.workloads import point_grid

tel = build_telescope(load_packed_config("CT5"), I.MCIntegrator(115), I.random.key(0))
full = torch.from_numpy(point_grid(64, 1.5)).cuda()
for S in (64, 256, 512, 1024, 4096):
    src, val = full[:S].contiguous(), torch.ones(S, device="cuda")
    for _ in range(5):
        render(tel, src, val, "point", 0)
    torch.cuda.synchronize()
    n = 200
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        render(tel, src, val, "point", 0)
    e1.record()
    t_host = (time.perf_counter() - t0) / n
    torch.cuda.synchronize()
    t_dev = e0.elapsed_time(e1) / n
    line = f"S={S:5d}: host enqueue {1e6 * t_host:7.1f} us/call, back-to-back device {1e3 * t_dev:7.1f} us/call"
    try:
        g = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            render(tel, src, val, "point", 0)
            with torch.cuda.graph(g, stream=s):
                out = render(tel, src, val, "point", 0)
        torch.cuda.synchronize()
        ref = render(tel, src, val, "point", 0)
        g.replay()
        torch.cuda.synchronize()
        ok = torch.allclose(out, ref, rtol=1e-4, atol=1e-6 * float(ref.max()))
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        line += f", CUDA-graph replay {1e3 * e0.elapsed_time(e1) / n:7.1f} us/call (matches: {ok})"
    except Exception as e:  # noqa: BLE001
        line += f", CUDA graph: {type(e).__name__}: {str(e)[:120]}"
    print(line, flush=True)
