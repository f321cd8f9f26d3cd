"""Summarise .ncu-rep files (read here, no GPU): per launch duration, DRAM bytes, throughput %, occupancy, registers, top stall reasons.
usage: python tools/ncu_summary.py gpurun_out/x.ncu-rep [...]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "l1tex__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_lsu.sum", "smsp__cycles_active.avg",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_drain_per_issue_active.ratio", "smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio"]


def main():
    for path in sys.argv[1:]:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(path, "no data"); continue
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        print(f"== {path}")
        for r in rows[2:]:
            name = r[idx["Kernel Name"]][:60]
            vals = {}
            for k in KEYS:
                if k in idx:
                    vals[k] = (r[idx[k]], units[idx[k]])
            def f(k):
                try:
                    return float(vals[k][0].replace(",", ""))
                except Exception:
                    return float("nan")
            dur = f("gpu__time_duration.sum"); du = vals.get("gpu__time_duration.sum", ("", ""))[1]
            rd, wr = f("dram__bytes_read.sum"), f("dram__bytes_write.sum")
            ru, wu = vals.get("dram__bytes_read.sum", ("", ""))[1], vals.get("dram__bytes_write.sum", ("", ""))[1]
            print(f"  {name}: dur={dur}{du} dram_rd={rd}{ru} wr={wr}{wu} dram%={f('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):.1f} "
                  f"sm%={f('sm__throughput.avg.pct_of_peak_sustained_elapsed'):.1f} warps_active%={f('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} "
                  f"regs={vals.get('launch__registers_per_thread', ('?',))[0]} grid={vals.get('launch__grid_size', ('?',))[0]} "
                  f"L2bytes={vals.get('lts__t_bytes.sum', ('?', ''))[0]}{vals.get('lts__t_bytes.sum', ('', ''))[1]}")
            stalls = sorted(((f(k), k.split('stalled_')[1].split('_per_issue')[0]) for k in KEYS if 'stalled' in k and k in vals), reverse=True)[:5]
            print("     stalls/issue: " + ", ".join(f"{n}={v:.2f}" for v, n in stalls))


if __name__ == "__main__":
    main()
