"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of counters DESIGN.md and
profiles/ quote. usage: ncu_summary.py report.ncu-rep [more.ncu-rep ...]"""
import csv
import io
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"), ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("launch__occupancy_limit_registers", "occ limit regs (blocks)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("dram__bytes_read.sum", "dram read"), ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of peak"),
    ("lts__t_bytes.sum", "L2 bytes"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("lts__t_sectors.sum", "L2 sectors"), ("lts__t_sectors_op_read.sum", "L2 sectors read"), ("lts__t_sectors_op_write.sum", "L2 sectors written"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 sectors read by SMs"),
    ("l1tex__t_bytes.sum", "L1 bytes"), ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "global ld requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "global ld sectors"),
    ("l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "global st requests"),
    ("l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "global st sectors"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "L1 data-pipe wavefronts % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads / warp inst"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
]


def summarise(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    lines = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"## {d.get('Kernel Name', '?')}  (id {d.get('ID')})")
        for key, label in WANT:
            if key in d:
                lines.append(f"  {label:38s} {d[key]:>18s} {u[key]}")
        rq, sc = d.get("l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"), d.get("l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum")
        try:
            lines.append(f"  {'sectors / global ld request':38s} {float(sc.replace(',', '')) / float(rq.replace(',', '')):18.2f}")
        except Exception:
            pass
    return "\n".join(lines)


if __name__ == "__main__":
    for p in sys.argv[1:]:
        print(f"# {p}")
        print(summarise(p))
        print()
