"""Summarise an .ncu-rep (read with `ncu -i ... --page raw --csv`) into a small JSON for profiles/.
    python -m tools.ncu_summary gpurun_out/prof.ncu-rep profiles/r01_sketch.json [units_per_launch unit_name]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "smsp__inst_executed_op_shared_ld.sum", "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_alu.sum",
    "smsp__inst_executed_pipe_fma.sum", "smsp__inst_executed_pipe_lsu.sum", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__t_bytes.sum", "sm__cycles_active.avg",
    "sm__cycles_elapsed.avg.per_second",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else None
    unit_name = sys.argv[4] if len(sys.argv) > 4 else "units"
    if rep.endswith(".csv"):  # already exported on the GPU box (`ncu -i x.ncu-rep --page raw --csv`)
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, unit = rows[0], rows[1]
    res = []
    for row in rows[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, unit))
        k = {"kernel": d.get("Kernel Name"), "metrics": {}}
        for key in KEYS:
            if key in d and d[key] != "":
                k["metrics"][key] = {"value": d[key], "unit": u[key]}
        if units:
            ms = float(d["gpu__time_duration.sum"].replace(",", "")) * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[u["gpu__time_duration.sum"]]
            inst = float(d["smsp__inst_executed.sum"].replace(",", ""))
            k["derived"] = {f"{unit_name}_per_launch": units, f"{unit_name}_per_s": units / (ms * 1e-3),
                            f"warp_inst_x32_per_{unit_name}": inst * 32 / units,
                            "dram_bytes_per_launch": sum(float(d[x].replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[x]]
                                                         for x in ("dram__bytes_read.sum", "dram__bytes_write.sum"))}
        res.append(k)
    json.dump({"source": rep, "note": "ncu --set full --clock-control none; numbers taken under the profiler are not bench values",
               "kernels": res}, open(out, "w"), indent=1)
    print(json.dumps(res[0].get("derived", {}), indent=1))


if __name__ == "__main__":
    main()
