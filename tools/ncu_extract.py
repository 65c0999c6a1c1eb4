"""Pull the headline metrics of every kernel in one or more .ncu-rep files (`ncu -i X --page raw --csv`) into JSON.
Usage: ncu_extract.py out.json rep1 [rep2 ...]"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_ns",
    "dram__bytes_read.sum": "dram_read_bytes",
    "dram__bytes_write.sum": "dram_write_bytes",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__inst_executed_pipe_tensor.sum": "tensor_pipe_inst",
    "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active": "tensor_hmma_active_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_active_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
    "sm__inst_executed.avg.per_cycle_elapsed": "ipc_elapsed",
    "smsp__issue_active.avg.pct": "issue_active_pct",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "launch__registers_per_thread": "registers",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__shared_mem_per_block_dynamic": "dyn_smem",
}
out = {}
for rep in sys.argv[2:]:
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    cols, units = rows[hdr], rows[hdr + 1]
    ki = cols.index("Kernel Name")
    res = []
    for r in rows[hdr + 2:]:
        if len(r) < len(cols):
            continue
        d = {"kernel": r[ki][:120]}
        for c, name in KEYS.items():
            if c in cols:
                v = r[cols.index(c)].replace(",", "")
                try:
                    d[name] = float(v)
                except ValueError:
                    d[name] = v
                if name.endswith("_bytes") or name == "duration_ns":
                    d[name + "_unit"] = units[cols.index(c)]
        res.append(d)
    out[rep.split("/")[-1]] = res
json.dump(out, open(sys.argv[1], "w"), indent=1)
print("wrote", sys.argv[1], {k: len(v) for k, v in out.items()})
