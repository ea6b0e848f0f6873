"""Reads an `ncu --set full` report of fe_env_step_kernel (one launch) and writes
  profiles/traffic.json      the numbers bench.py quotes (DRAM bytes, warp instructions, active lanes), stamped with the build id
  profiles/<tag>_ncu_summary.txt   the human-readable summary (key metrics + stall reasons)
Usage: python tools/ncu_extract.py gpurun_out/prof.ncu-rep r2a "command line of the capture"
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep, tag = sys.argv[1], sys.argv[2]
cmd = sys.argv[3] if len(sys.argv) > 3 else ""
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
vals = rows[2] if len(rows) > 2 and not rows[1][0].isdigit() else rows[1]
d = dict(zip(hdr, vals))
units = dict(zip(hdr, rows[1])) if vals is not rows[1] else {}


def f(k):
    return float(d[k].replace(",", ""))


def to_bytes(k):
    u = units.get(k, "byte").lower()
    mult = {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}.get(u, 1)
    return f(k) * mult


dram = to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum")
out = {
    "build_id": bench.build_id(),
    "kernel": "fe_env_step_kernel",
    "dram_bytes_per_launch": dram,
    "dram_bytes_read": to_bytes("dram__bytes_read.sum"),
    "dram_bytes_write": to_bytes("dram__bytes_write.sum"),
    "warp_instructions_per_launch": f("smsp__inst_executed.sum"),
    "lanes_active_per_instruction": f("smsp__thread_inst_executed_per_inst_executed.ratio"),
    "duration_ms_under_ncu": f("gpu__time_duration.sum") * {"msecond": 1, "usecond": 1e-3, "second": 1e3, "nsecond": 1e-6}.get(units.get("gpu__time_duration.sum", "msecond"), 1),
    "source": "profiles/%s_env_step_ncu_summary.txt" % tag,
}
json.dump(out, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "sm__icc_request_hit_rate.pct", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__sass_inst_executed_op_local_ld.sum",
        "smsp__sass_inst_executed_op_local_st.sum", "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
with open(os.path.join(ROOT, "profiles", "%s_env_step_ncu_summary.txt" % tag), "w") as fo:
    fo.write("%s\nbuild id %s\n\n" % (cmd, out["build_id"]))
    for k in keys:
        if k in d:
            fo.write("%-70s %s %s\n" % (k, d[k], units.get(k, "")))
    fo.write("warp stall reasons (warps per issue-active cycle):\n")
    st = sorted(((f(k), k) for k in d if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")), reverse=True)
    for v, k in st[:12]:
        fo.write("  %-22s %.3f\n" % (k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")], v))
print(json.dumps(out))
