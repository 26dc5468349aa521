"""Summarise two `ncu --set full` captures of one kernel (A/B: e.g. a search-heavy and a steady-state frame of
the staged depth-filter kernel) into profiles/<out>.{md,json}.  Runs where ncu is installed (no GPU needed):

    python tools/ncu_summary.py gpurun_out/prof_staged_p5_heavy.ncu-rep gpurun_out/prof_staged_p5_steady.ncu-rep \
        r02_ncu_staged "VGA, 5x5 patch" [algorithmic bytes per launch]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_active.avg", "sm__cycles_elapsed.avg",
    "sm__cycles_active.min", "sm__cycles_active.max", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__issue_active.avg.per_cycle_active",
] + ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % r for r in (
    "wait", "short_scoreboard", "long_scoreboard", "barrier", "math_pipe_throttle", "mio_throttle", "no_instruction",
    "branch_resolving", "dispatch_stall", "not_selected", "membar")]
TO_BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def raw(path):
    txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    names, units, vals = rows[0], rows[1], rows[2]
    return {n: (v, u) for n, u, v in zip(names, units, vals)}


def main():
    heavy, steady = raw(sys.argv[1]), raw(sys.argv[2])
    base = sys.argv[3] if len(sys.argv) > 3 else "r02_ncu_staged"
    what = sys.argv[4] if len(sys.argv) > 4 else "VGA, 5x5 patch"
    algo = float(sys.argv[5]) if len(sys.argv) > 5 else 15974400.0
    names = (sys.argv[6], sys.argv[7]) if len(sys.argv) > 7 else ("heavy", "steady")
    out = {"traffic_bytes_per_launch": {}, "metrics": {}, "what": what, "algorithmic_bytes_per_launch": algo}
    lines = [f"| metric | {names[0]} | {names[1]} |", "|---|---|---|"]
    for m in METRICS:
        if m not in heavy:
            continue
        (hv, hu), (sv, su) = heavy[m], steady[m]
        out["metrics"][m] = {"heavy": hv, "steady": sv, "unit": hu, "unit_steady": su}
        lines.append(f"| `{m}` | {hv} {hu} | {sv} {su} |")
    for key, rep in (("heavy", heavy), ("steady", steady)):
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            v, u = rep[m]
            tot += float(v.replace(",", "")) * TO_BYTES.get(u, 1.0)
        out["traffic_bytes_per_launch"][key] = tot
    with open(os.path.join(ROOT, "profiles", base + ".json"), "w") as f:
        json.dump(out, f, indent=1)
    t = out["traffic_bytes_per_launch"]
    md = f"""# {base} -- ncu `--set full --clock-control none` captures: {what}

Captured with `tools/gpu_trip.sh` (stages ncu_*) on a B200; columns = {names[0]} / {names[1]} launch. Times under ncu are
cold-cache and serialised; `bench.py` reports the CUDA-event times.  Written by `tools/ncu_summary.py`.

""" + "\n".join(lines) + f"""

DRAM traffic per launch (read + write): {names[0]} {t['heavy'] / 1e6:.2f} MB, {names[1]} {t['steady'] / 1e6:.2f} MB; algorithmic
bytes per launch {algo / 1e6:.2f} MB.
"""
    with open(os.path.join(ROOT, "profiles", base + ".md"), "w") as f:
        f.write(md)
    print(md)


if __name__ == "__main__":
    main()
