"""Condense an ncu report (.ncu-rep, read here without a GPU) into the key numbers we track:
python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/name.json [label]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration",
    "smsp__inst_executed.sum": "warp_instructions",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active": "fp64_pipe_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__shared_mem_per_block_dynamic": "dyn_smem_per_block",
    "launch__occupancy_limit_shared_mem": "occupancy_limit_smem_blocks",
    "launch__occupancy_limit_registers": "occupancy_limit_regs_blocks",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum": "smem_wavefronts",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum": "smem_bank_conflicts",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio": "stall_short_scoreboard",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio": "stall_wait",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio": "stall_math_pipe_throttle",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio": "stall_not_selected",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio": "stall_no_instruction",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum": "dfma_thread_inst",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum": "dmul_thread_inst",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum": "dadd_thread_inst",
}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    label = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"label": label}
        for h, u, v in zip(hdr, units, vals):
            if h == "Kernel Name":
                d["kernel"] = v
            if h in KEYS:
                try:
                    d[KEYS[h]] = {"value": float(v.replace(",", "")), "unit": u}
                except ValueError:
                    d[KEYS[h]] = {"value": v, "unit": u}
        res.append(d)
    json.dump(res, open(out, "w"), indent=1)
    for d in res:
        print(d.get("kernel", "?")[:70])
        for k, v in d.items():
            if isinstance(v, dict):
                print(f"  {k:32s} {v['value']} {v['unit']}")


if __name__ == "__main__":
    main()
