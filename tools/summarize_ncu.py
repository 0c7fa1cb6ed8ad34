"""Summarise an .ncu-rep (read here, without a GPU) into a small text file for profiles/.

usage: python tools/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/NAME.txt [rays_per_launch [bench_workload]]

With a bench workload name the first kernel's figures also go to profiles/ncu_figures.json, which bench.py reads for
roofline.traffic / executed_fp32_frac / issue_slot_util / atomic.
"""
import json
import os
from pathlib import Path
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed.avg.per_cycle_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
        "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed_op_shared_atom.sum", "smsp__inst_executed_op_generic_atom_dot_alu.sum",
        "smsp__inst_executed_op_global_red.sum",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep, out = sys.argv[1], sys.argv[2]
    rays = float(sys.argv[3]) if len(sys.argv) > 3 else None
    rows = list(csv.reader(ncu(rep, "--page", "raw", "--csv").splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# ncu summary of {rep}", ""]
    figures = {}
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        lines.append(f"## kernel: {name[:160]}")
        vals = dict(zip(hdr, r))
        for k in KEYS:
            if k in vals:
                lines.append(f"{k} [{units[hdr.index(k)]}] = {vals[k]}")
        if rays:
            try:
                wi = float(vals["smsp__inst_executed.sum"].replace(",", ""))
                tpi = float(vals["smsp__thread_inst_executed_per_inst_executed.ratio"])
                cyc = float(vals["sm__cycles_elapsed.avg"].replace(",", ""))
                fl = sum(float(vals[f"smsp__sass_thread_inst_executed_op_{o}_pred_on.sum.per_cycle_elapsed"].replace(",", "")) * m
                         for o, m in (("ffma", 2), ("fadd", 1), ("fmul", 1))) * cyc
                lines.append(f"derived: warp instructions per ray = {wi / rays:.2f}; thread instructions per ray = {wi * tpi / rays:.0f}; "
                             f"executed FP32 flops per ray (FFMA=2) = {fl / rays:.0f}")
                if len(sys.argv) > 4 and not figures:
                    num = lambda k: float(vals[k].replace(",", ""))
                    unit = units[hdr.index("dram__bytes_read.sum")].lower()
                    scale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[unit]
                    wunit = units[hdr.index("dram__bytes_write.sum")].lower()
                    wscale = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[wunit]
                    figures.update(dram_bytes=int(num("dram__bytes_read.sum") * scale + num("dram__bytes_write.sum") * wscale),
                                   executed_fp32_frac=round(fl / cyc / (148 * 128 * 2), 4),
                                   issue_slot_util=round(num("smsp__issue_active.avg.pct_of_peak_sustained_active") / 100, 4),
                                   thread_inst_per_ray=round(wi * tpi / rays, 1),
                                   shared_atom_inst=int(num("smsp__inst_executed_op_shared_atom.sum")),
                                   kernel_ms_under_ncu=round(num("gpu__time_duration.sum"), 4), source=out)
            except Exception as e:  # pragma: no cover
                lines.append(f"derived: n/a ({e})")
        lines.append("")
    src = list(csv.reader(ncu(rep, "--page", "source", "--csv", "--print-source", "sass,cuda").splitlines()))
    cur, h, agg = None, None, []
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
        elif len(r) > 3 and r[0] == "Line No":
            h = r
        elif h and len(r) > 8 and r[0].isdigit():
            try:
                agg.append((int(r[h.index("Instructions Executed")]), int(r[h.index("# Samples")]), cur, int(r[0]), r[1].strip()[:110]))
            except ValueError:
                pass
    if agg:
        ti, ts = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
        lines.append("## hottest source lines (share of warp instructions / of stall samples)")
        for a in sorted(agg, key=lambda a: -a[0])[:int(os.environ.get("NCU_TOP_LINES", "30"))]:
            lines.append(f"{100 * a[0] / ti:5.1f}% inst {100 * a[1] / ts:5.1f}% smp  {a[2]}:{a[3]}  {a[4]}")
    open(out, "w").write("\n".join(lines) + "\n")
    if figures:
        fj = Path(__file__).resolve().parents[1] / "profiles" / "ncu_figures.json"
        allf = json.loads(fj.read_text()) if fj.exists() else {}
        allf[sys.argv[4]] = figures
        fj.write_text(json.dumps(allf, indent=1) + "\n")
    print("\n".join(lines[:60]))


if __name__ == "__main__":
    main()
