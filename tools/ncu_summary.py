#!/usr/bin/env python3
"""Summarise an .ncu-rep (ncu --set full) into the small JSON kept under profiles/: per captured launch the duration,
DRAM bytes, pipe / memory utilisation, occupancy, registers and the top stall reasons.
usage: ncu_summary.py report.ncu-rep out.json "capture command" """
import csv, io, json, subprocess, sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fmaheavy.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]
STALL = "smsp__average_warps_issue_stalled_"


def main():
    rep, out, cmd = sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    res = {"capture": cmd, "kernels": []}
    for r in rows[2:]:
        k = {"kernel": r[hdr.index("Kernel Name")][:90], "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")], "metrics": {}, "stalls_per_issue": {}}
        for i, h in enumerate(hdr):
            for key in KEYS:
                if h == key or h.endswith("." + key):
                    try:
                        k["metrics"][key] = {"value": float(r[i]), "unit": units[i]}
                    except ValueError:
                        pass
            if STALL in h and h.endswith("_per_issue_active.ratio"):
                try:
                    v = float(r[i])
                    if v > 0.05:
                        k["stalls_per_issue"][h.split(STALL)[1].replace("_per_issue_active.ratio", "")] = round(v, 3)
                except ValueError:
                    pass
        k["stalls_per_issue"] = dict(sorted(k["stalls_per_issue"].items(), key=lambda kv: -kv[1])[:6])
        m = k["metrics"]
        if "dram__bytes_read.sum" in m and "dram__bytes_write.sum" in m:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            k["dram_bytes_per_launch"] = sum(m[x]["value"] * scale.get(m[x]["unit"], 1) for x in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        res["kernels"].append(k)
    json.dump(res, open(out, "w"), indent=1)
    for k in res["kernels"]:
        m = k["metrics"]
        print(k["kernel"][:60], {x.split(".")[0].replace("__", ":"): round(v["value"], 2) for x, v in m.items() if x in KEYS[:9]}, k["stalls_per_issue"])


if __name__ == "__main__":
    main()
