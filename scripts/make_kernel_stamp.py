#!/usr/bin/env python
"""profiles/kernel_stamp.json from an ncu capture of fly_mlp_f64_kernel (dev tool, no GPU needed):
DRAM traffic of the launch (`--page raw`), executed fp64-pipe instructions per drone-step (`--page source`), stamped
with the capture's workload and the sha256 of csrc/ so that bench.py refuses the numbers once the kernel changes.

    python scripts/make_kernel_stamp.py gpurun_out/r02_fly_survivors.ncu-rep 65536 2000 survivors profiles/r02_fly_survivors
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

rep, drones, H, pop, out_prefix = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
keep = [h for h in hdr if h in ("Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
                                "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__cycles_elapsed.max",
                                "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
                                "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                                "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
                                "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_ld.sum")]
with open(out_prefix + "_ncu_full_summary.csv", "w") as f:
    f.write("metric,unit,value\n")
    for h in keep:
        f.write(f'{h},{m[h][0]},"{m[h][1]}"\n')


def to_bytes(name):
    unit, v = m[name]
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


traffic = int(to_bytes("dram__bytes_read.sum") + to_bytes("dram__bytes_write.sum"))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
srows = list(csv.reader(io.StringIO(src)))
shdr, sdata = srows[1], srows[2:]
col = {h: i for i, h in enumerate(shdr)}
ops = collections.Counter()
for r in sdata:
    t = r[col["Source"]].split()
    op = (t[1] if t[0].startswith("@") else t[0]).split(".")[0]
    ops[op] += int(r[col["Instructions Executed"]])
warp_steps = drones * H / 32.0
fp64 = sum(ops[o] for o in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX")) / warp_steps
stamp = {"csrc_sha16": bench.csrc_hash(), "workload": {"drones": drones, "timesteps": H, "population": pop},
         "traffic_bytes": traffic, "fp64_instr_per_drone_step": round(fp64, 1),
         "fp64_mix_per_drone_step": {o: round(ops[o] / warp_steps, 1) for o in ("DFMA", "DMUL", "DADD", "DSETP")},
         "warp_instr_per_warp_step": round(sum(ops.values()) / warp_steps, 1),
         "kernel_ms_under_ncu": float(m["gpu__time_duration.sum"][1]) * {"ms": 1, "us": 1e-3, "msecond": 1, "usecond": 1e-3, "ns": 1e-6}.get(m["gpu__time_duration.sum"][0], 1),
         "source": os.path.relpath(out_prefix + "_ncu_full_summary.csv", ROOT) + " (ncu --set full --clock-control none --import-source on; source page for the instruction counts)"}
with open(os.path.join(ROOT, "profiles", "kernel_stamp.json"), "w") as f:
    json.dump(stamp, f, indent=1)
print(json.dumps(stamp, indent=1))
