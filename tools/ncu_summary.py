"""Condense an .ncu-rep (ncu --set full) into one JSON line per captured launch -> profiles/*.jsonl.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/x_summary.jsonl "comment line"
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_issued.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    comment = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out, "w") as f:
        if comment:
            f.write("# %s\n" % comment)
        f.write("# one line per captured launch; values as printed by `ncu --page raw --csv` (value unit); stalls = warp-cycles per issued instruction\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            line = {}
            for k in KEYS:
                if k in d and d[k] != "":
                    u = units[hdr.index(k)]
                    line[k] = (d[k] + " " + u).strip()
            stalls = []
            for h in hdr:
                if "issue_stalled" in h and h.endswith("per_issue_active.ratio"):
                    try:
                        stalls.append((h.split("issue_stalled_")[1].replace("_per_issue_active.ratio", ""), float(d[h].replace(",", ""))))
                    except ValueError:
                        pass
            stalls.sort(key=lambda x: -x[1])
            line["top_stalls"] = {k: round(v, 2) for k, v in stalls[:5]}
            f.write(json.dumps(line) + "\n")


if __name__ == "__main__":
    main()
