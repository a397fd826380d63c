"""Summarise an .ncu-rep (read here, no GPU needed) into a small JSON for profiles/.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_two_phase_ncu.json [label]
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tma.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_active.avg.per_cycle_active", "smsp__warps_eligible.avg.per_cycle_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg",
    "smsp__sass_inst_executed_op_tma_ld.sum", "smsp__sass_inst_executed_op_shared_ld.sum",
    "smsp__sass_inst_executed_op_global_ld.sum", "smsp__sass_inst_executed_op_global_st.sum",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    label = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = {"label": label, "source": rep, "kernels": []}
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        k = {"kernel": d.get("Kernel Name"), "metrics": {}}
        for key in KEYS:
            if key in d and d[key] not in (None, ""):
                k["metrics"][key] = {"value": d[key], "unit": u.get(key, "")}
        stalls = {}
        for h in hdr:
            if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
                nm = h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]
                try:
                    v = float(d[h])
                except (ValueError, TypeError):
                    continue
                if v >= 0.01:
                    stalls[nm] = round(v, 3)
        k["stall_warps_per_issue"] = dict(sorted(stalls.items(), key=lambda x: -x[1]))
        res["kernels"].append(k)
    json.dump(res, open(out, "w"), indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
