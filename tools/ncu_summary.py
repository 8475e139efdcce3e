"""Summarise an .ncu-rep (raw page) into a small text file for profiles/.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.txt [note]"""
import csv
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.per_cycle_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed.sum",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors_srcunit_tex_op_read.sum",
    "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    lines = [f"# {rep}", f"# {note}", ""]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        lines.append(f"== kernel: {d.get('Kernel Name', '?')}  (id {d.get('ID', '?')})")
        for k in KEYS:
            if k in d:
                lines.append(f"{k:75s} {d[k]:>22s} {units[hdr.index(k)]}")
        stalls = [(float(d[h].replace(',', '')), h[len(STALL):-len('_per_issue_active.ratio')]) for h in hdr
                  if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and d[h] not in ("", "n/a")]
        lines.append("stall reasons (warps per issue-active cycle): " +
                     ", ".join(f"{n}={v:.2f}" for v, n in sorted(stalls, reverse=True)[:8]))
        lines.append("")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()
