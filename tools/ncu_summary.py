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


def kernel_source_sha():
    """sha256 of the sources of the transport kernel (everything the kernel is made of is in these two headers; the
    host side of the library, gate_b200.cu, is not part of it): a capture only describes the kernel built from
    exactly these files"""
    import hashlib
    import os
    root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "opengate_b200", "csrc")
    h = hashlib.sha256()
    for f in ("transport.cuh", "device_math.cuh"):
        h.update(open(os.path.join(root, f), "rb").read())
    return h.hexdigest()[:16]


def d_kernel_name(hdr, vals):
    """the instantiation as gate_b200_kernel_name() spells it: 'gb2::transport_kernel<unsigned char, 2>'"""
    import re
    nm = dict(zip(hdr, vals)).get("Kernel Name", "")
    m = re.search(r"transport_kernel<[^>]*>", nm)
    return ("gb2::" + m.group(0)) if m else nm


# the phases of the persistent kernel by the device functions inlined into it (transport.cuh): which share of the
# issue slots / stall samples each queue's phase takes
PHASES = {
    "Q_V woodcock": ("volume_gather_labels", "volume_gather_mu", "volume_decide", "route_interaction", "tab_mu", "bin_split", "shell_flight"),
    "Q_C compton": ("phase_compton", "interp_q2", "q2_segment", "rotate_dir"),
    "Q_E end/primary": ("phase_end", "generate_primary", "locate", "cdf_index", "lower_bound_f", "gauss", "flush_buffer", "grid_add", "red_add_f64"),
    "Q_W world": ("phase_world", "ray_box_ax", "ray_box", "snap_to_face", "to_world", "to_volume", "inside_box", "aff_apply", "rot_apply", "aff_apply_inv", "rot_apply_inv", "phsp_store"),
    "Q_R rare": ("phase_rare",),
    "Q_T tle": ("phase_tle", "tle_table_add", "tle_flush_all"),
    "deposit": ("deposit", "ld_dep", "st_dep", "hist_weight"),
    "scheduler (claim/push)": ("transport_kernel", "ld_vol", "st_vol", "ld_cell", "st_cell"),
    "state load/store": ("load_ph", "store_ph", "set_energy", "energy_bin_t"),
    "statistics": ("stat_add64", "stat_add", "stat_reduce_packed", "stat_bump", "stat_overflow", "stat_tle_overflow"),
}


def phase_table(rep, srcf):
    """per-phase share of samples / instructions from the source page (functions are inlined: lines are mapped to
    the device function that contains them)"""
    import bisect
    import re
    src = ncu_csv(rep, "source", ["--print-source", "cuda,sass"])
    hdr, cur, agg = None, None, []
    for r in src:
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r and r[0] == "Line No":
            hdr = r
            continue
        if r and r[0].isdigit() and hdr and len(r) == len(hdr):
            d = {}
            for k, v in zip(hdr, r):
                d.setdefault(k, v)
            g = lambda k: int(d.get(k) or 0)   # noqa: E731
            agg.append((cur, int(r[0]), g("# Samples"), g("Instructions Executed"), g("Thread Instructions Executed"),
                        g("stall_long_sb"), g("stall_no_inst")))
    lines = open(srcf).read().split("\n")
    funcs = []
    for i, l in enumerate(lines):
        if re.match(r"\s*(__device__|__global__)", l) or (re.match(r"\s*template", l) and i + 1 < len(lines) and re.match(r"\s*(__device__|__global__)", lines[i + 1])):
            txt = " ".join(lines[i:i + 3])
            m = re.search(r"\b(\w+)\s*\(", re.sub(r"__launch_bounds__\([^)]*\)(\s*,\s*\w+\))?", "", re.sub(r"template\s*<[^>]*>", "", txt)))
            if m and (not funcs or funcs[-1][1] != m.group(1) or funcs[-1][0] < i):
                funcs.append((i + 1, m.group(1)))
    starts = [f[0] for f in funcs]
    f2p = {fn: ph for ph, fns in PHASES.items() for fn in fns}
    tot = [0] * 5
    by = {}
    for f, ln, smp, ins, tins, lsb, ni in agg:
        if f == "transport.cuh":
            i = bisect.bisect_right(starts, ln) - 1
            fn = funcs[i][1] if i >= 0 else "?"
            ph = f2p.get(fn, "other (" + fn + ")")
        elif f == "device_math.cuh":
            ph = "math: philox/log/exp/sincos"
        else:
            ph = "intrinsics headers (shfl, atomics, ...)"
        a = by.setdefault(ph, [0] * 5)
        for k, v in enumerate((smp, ins, tins, lsb, ni)):
            a[k] += v
            tot[k] += v
    out = ["## by phase (source-level samples; device functions are inlined into the one kernel)", "",
           "| phase | % stall samples | % warp instructions | lanes / instruction | % of samples: long_scoreboard | no_instruction |",
           "|---|---|---|---|---|---|"]
    for ph, v in sorted(by.items(), key=lambda kv: -kv[1][0]):
        if v[0] * 200 < max(tot[0], 1):
            continue
        out.append(f"| {ph} | {100 * v[0] / max(tot[0], 1):.1f} | {100 * v[1] / max(tot[1], 1):.1f} | {v[2] / max(v[1], 1):.1f} | "
                   f"{100 * v[3] / max(tot[0], 1):.1f} | {100 * v[4] / max(tot[0], 1):.1f} |")
    return out + [""]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--histories", type=float, default=0, help="histories in the profiled launch")
    ap.add_argument("--title", default="")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--traffic-json", default="", help="update this json with dram bytes/history under --key")
    ap.add_argument("--key", default="c3")
    ap.add_argument("--source", default="opengate_b200/csrc/transport.cuh", help="transport.cuh of the profiled build")
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
                    kn = d_kernel_name(hdr, vals)
                    d[a.key] = {"dram_bytes_per_history": (rd + wr) / a.histories, "histories_in_capture": a.histories,
                                "dram_read_bytes": rd, "dram_write_bytes": wr, "source": os.path.basename(a.out),
                                "kernel": kn, "kernel_source_sha": kernel_source_sha()}
                    json.dump(d, open(a.traffic_json, "w"), indent=1)
        except Exception:
            pass
        lines.append("")

    try:
        lines += phase_table(a.rep, a.source)
    except Exception as e:   # the phase table is a convenience: never lose the rest of the summary over it
        lines += [f"(phase table failed: {e})", ""]
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
