"""Summarise an .ncu-rep (captured under gpurun) into profiles/<name>.json + .txt: duration, DRAM bytes, pipe
utilisation, issue rate, stall reasons, occupancy.  Usage: python scripts/ncu_summary.py <rep> <out-prefix>"""
import csv
import io
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes_write.sum.per_second",
    "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg.per_second",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__warps_active.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "sm__sass_thread_inst_executed_op_ffma_pred_on.sum", "sm__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "sm__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum",
]


def main(rep, prefix):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        rec = {"kernel": vals[hdr.index("Kernel Name")], "metrics": {}, "stalls_per_issue": {}}
        for h, u, v in zip(hdr, units, vals):
            if h in KEEP:
                rec["metrics"][h] = f"{v} {u}".strip()
            if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                name = h.split("issue_stalled_")[1].split("_per_issue")[0]
                try:
                    rec["stalls_per_issue"][name] = round(float(v), 3)
                except ValueError:
                    pass
        out.append(rec)
    with open(prefix + ".json", "w") as f:
        json.dump(out, f, indent=1)
    with open(prefix + ".txt", "w") as f:
        for rec in out:
            f.write(f"kernel: {rec['kernel']}\n")
            for k, v in rec["metrics"].items():
                f.write(f"  {k:75s} {v}\n")
            top = sorted(rec["stalls_per_issue"].items(), key=lambda kv: -kv[1])[:8]
            f.write("  warp stall reasons (avg warps stalled per issue-active cycle): " + ", ".join(f"{k}={v}" for k, v in top) + "\n\n")
    print(open(prefix + ".txt").read())


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
