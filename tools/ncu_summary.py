"""Summarise an .ncu-rep (read with the ncu CLI, no GPU needed): the metrics the roofline discussion uses.

  python tools/ncu_summary.py gpurun_out/x.ncu-rep [--md]
  python tools/ncu_summary.py gpurun_out/x.ncu-rep --traffic profiles/integrate_traffic.json bricks
      also records the capture's DRAM bytes per launch (mean over the captured launches) under the given key; bench.py
      reports it as roofline.traffic
"""
import csv
import io
import subprocess
import sys

WANT = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio",
]


def main():
    rep = sys.argv[1]
    md = "--md" in sys.argv
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    if md:
        print("| metric | unit | " + " | ".join(f"launch {i}" for i in range(len(data))) + " |")
        print("|---|---|" + "---|" * len(data))
    for w in WANT:
        if w not in hdr:
            continue
        i = hdr.index(w)
        vals = [r[i][:70] for r in data]
        if md:
            print(f"| `{w}` | {units[i]} | " + " | ".join(vals) + " |")
        else:
            print(f"{w:84s} {units[i]:14s} {vals}")


def traffic(rep, out_json, key):
    import json
    import os
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}

    def col(name):
        i = hdr.index(name)
        return sum(float(r[i].replace(",", "")) * scale[units[i]] for r in data) / len(data)

    rec = {"kernel": data[0][hdr.index("Kernel Name")], "launches": len(data),
           "dram_bytes_read": int(col("dram__bytes_read.sum")), "dram_bytes_write": int(col("dram__bytes_write.sum")),
           "source": f"ncu --set full --clock-control none, {os.path.basename(rep)}"}
    cur = json.load(open(out_json)) if os.path.exists(out_json) else {}
    cur[key] = rec
    json.dump(cur, open(out_json, "w"), indent=1)
    print(json.dumps(rec))


if __name__ == "__main__":
    if "--traffic" in sys.argv:
        k = sys.argv.index("--traffic")
        traffic(sys.argv[1], sys.argv[k + 1], sys.argv[k + 2])
        sys.argv = sys.argv[:k]
    main()
