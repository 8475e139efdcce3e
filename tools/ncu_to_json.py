"""Numeric per-kernel summary of an .ncu-rep for bench.py's roofline object.

usage: python tools/ncu_to_json.py REP.ncu-rep SAMPLES_PER_LAUNCH SOURCE_NOTE [OUT.json]
Merges into OUT.json (default profiles/dominant_stage_ncu.json): per kernel (short name) the measured DRAM
bytes per launch / per sample and the unit utilisations that name its limiter:
  dram__bytes_read.sum + dram__bytes_write.sum, dram__throughput.avg.pct_of_peak_sustained_elapsed,
  l1tex__throughput.avg.pct_of_peak_sustained_elapsed, lts__throughput.avg.pct_of_peak_sustained_elapsed,
  smsp__issue_active.avg.per_cycle_active (IPC per scheduler), sm__warps_active.avg.pct_of_peak_sustained_active,
  sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active, gpu__time_duration.sum."""
import csv
import json
import os
import re
import subprocess
import sys

M = {"dram_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "l1tex_pct": "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
     "lts_pct": "lts__throughput.avg.pct_of_peak_sustained_elapsed",
     "ipc": "smsp__issue_active.avg.per_cycle_active",
     "occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
     "tensor_pipe_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
     "sm_pct": "sm__throughput.avg.pct_of_peak_sustained_elapsed"}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1e-6, "usecond": 1e-3,
              "msecond": 1.0, "second": 1e3}
STALL = "smsp__average_warps_issue_stalled_"


def num(s):
    try:
        return float(s.replace(",", ""))
    except Exception:
        return None


def main():
    rep, spl, note = sys.argv[1], float(sys.argv[2]), sys.argv[3]
    out_path = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                                    "profiles", "dominant_stage_ncu.json")
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    data = json.load(open(out_path)) if os.path.exists(out_path) else {}
    data["source"] = note
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = re.sub(r"<.*", "", d.get("Kernel Name", "?").replace("void ", "").replace("<unnamed>::", "")).split("(")[0]
        e = {}
        rd = num(d.get("dram__bytes_read.sum", "")); wr = num(d.get("dram__bytes_write.sum", ""))
        if rd is not None and wr is not None:
            b = rd * UNIT_SCALE.get(unit["dram__bytes_read.sum"], 1.0) + wr * UNIT_SCALE.get(unit["dram__bytes_write.sum"], 1.0)
            e["dram_bytes_per_launch"] = b
            e["dram_bytes_per_sample"] = b / spl
        t = num(d.get("gpu__time_duration.sum", ""))
        if t is not None:
            e["ncu_launch_ms"] = t * UNIT_SCALE.get(unit["gpu__time_duration.sum"], 1.0)
        for k, m in M.items():
            v = num(d.get(m, ""))
            if v is not None:
                e[k] = v
        # sector counts of the global loads: the gather kernels are measured against the random-sector ceiling of
        # tools/ubench_gather.cu (profiles/r02_ubench_gather.txt)
        for key, metric in (("l1_sectors_ld", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"),
                            ("l1_requests_ld", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"),
                            ("l1_miss_sectors_ld", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_miss.sum"),
                            ("l1_hit_sectors_ld", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld_lookup_hit.sum"),
                            ("l2_sectors_tex_read", "lts__t_sectors_srcunit_tex_op_read.sum")):
            v = num(d.get(metric, ""))
            if v is not None:
                e[key + "_per_sample"] = v / spl
        for key, metric in (("l1_data_pipe_pct", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
                            ("l1_xbar_req_pct", "l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed"),
                            ("l1_writeback_pct", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed")):
            v = num(d.get(metric, ""))
            if v is not None:
                e[key] = v
        stalls = sorted(((num(d[h]), h[len(STALL):-len("_per_issue_active.ratio")]) for h in hdr
                         if h.startswith(STALL) and h.endswith("_per_issue_active.ratio") and num(d[h]) is not None),
                        reverse=True)[:3]
        e["top_stalls"] = {n: v for v, n in stalls}
        e["samples_per_launch"] = spl
        e["source"] = note
        units_pct = {k: e.get(k, 0.0) for k in ("l1tex_pct", "lts_pct", "dram_pct", "tensor_pipe_pct")}
        top_unit = max(units_pct, key=units_pct.get)
        e["limiter"] = (f"{top_unit.replace('_pct', '')} at {units_pct[top_unit]:.0f} % of peak "
                        f"(l1tex {e.get('l1tex_pct', 0):.0f} %, lts {e.get('lts_pct', 0):.0f} %, dram {e.get('dram_pct', 0):.0f} %, "
                        f"tensor pipe {e.get('tensor_pipe_pct', 0):.0f} %), {e.get('ipc', 0):.2f} IPC per scheduler at "
                        f"{e.get('occupancy_pct', 0):.0f} % occupancy; top stalls " +
                        ", ".join(f"{n} {v:.1f}" for n, v in e["top_stalls"].items()))
        data[name] = e
        print(name, json.dumps(e)[:300])
    json.dump(data, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
