#!/usr/bin/env python
"""Summarise an .ncu-rep (from `ncu --set full --import-source on`) into a small markdown file
for profiles/: headline counters, DRAM traffic per launch, and the hottest source lines.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_transport.md [--histories N]
"""
import argparse
import csv
import io
import subprocess


def ncu_csv(rep, page, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


RAW = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "smsp__inst_executed_op_global_red.sum", "lts__t_sectors_srcunit_tex_op_red.sum",
    "lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warp_latency_issue_stalled_no_instruction.ratio", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--histories", type=float, default=0, help="histories in the profiled launch")
    ap.add_argument("--title", default="")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--traffic-json", default="", help="update this json with dram bytes/history under --key")
    ap.add_argument("--key", default="c3")
    a = ap.parse_args()

    raw = ncu_csv(a.rep, "raw")
    hdr, units = raw[0], raw[1]
    lines = [f"# ncu summary: {a.title or a.rep}", "",
             "Captured with `ncu --set full --clock-control none --import-source on` (one launch; values under the "
             "profiler are not bench numbers).", ""]
    for li, vals in enumerate(raw[2:]):
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        lines += [f"## launch {li}: `{d.get('Kernel Name', '?')}`", "", "| metric | value | unit |", "|---|---|---|"]
        for k in RAW:
            if k in d:
                lines.append(f"| {k} | {d[k]} | {u[k]} |")
        try:
            scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
            rd = float(d["dram__bytes_read.sum"]) * scale[u["dram__bytes_read.sum"]]
            wr = float(d["dram__bytes_write.sum"]) * scale[u["dram__bytes_write.sum"]]
            lines.append(f"| **traffic = dram read+write per launch** | {rd + wr:.4g} | byte |")
            if a.histories:
                lines.append(f"| traffic per history ({a.histories:.3g} histories) | {(rd + wr) / a.histories:.1f} | byte |")
                if a.traffic_json and li == 0:
                    import json
                    import os
                    d = json.load(open(a.traffic_json)) if os.path.exists(a.traffic_json) else {}
                    d[a.key] = {"dram_bytes_per_history": (rd + wr) / a.histories, "histories_in_capture": a.histories,
                                "dram_read_bytes": rd, "dram_write_bytes": wr, "source": os.path.basename(a.out)}
                    json.dump(d, open(a.traffic_json, "w"), indent=1)
        except Exception:
            pass
        lines.append("")

    src = ncu_csv(a.rep, "source", ["--print-source", "cuda,sass"])
    cur, agg = None, []
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0].isdigit():
            try:
                agg.append((cur, int(r[0]), r[1], int(r[4] or 0), int(r[7] or 0)))
            except (ValueError, IndexError):
                pass
    ts = sum(x[3] for x in agg) or 1
    ti = sum(x[4] for x in agg) or 1
    agg.sort(key=lambda x: -x[3])
    lines += [f"## hottest source lines (warp-stall samples, total {ts}; warp instructions {ti})", "",
              "| file:line | % samples | % instr | source |", "|---|---|---|---|"]
    for f, ln, s, smp, ins in agg[: a.top]:
        s = s.replace("|", "\\|")[:100]
        lines.append(f"| {f}:{ln} | {100 * smp / ts:.1f} | {100 * ins / ti:.1f} | `{s}` |")
    open(a.out, "w").write("\n".join(lines) + "\n")
    print("wrote", a.out)


if __name__ == "__main__":
    main()
