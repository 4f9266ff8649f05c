"""ncu report -> compact JSON summary (one record per captured launch): python scripts/ncu_summary.py x.ncu-rep out.json"""
import csv
import io
import json
import subprocess
import sys

KEYS = {
    "gpu__time_duration.sum": "duration_us", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct", "launch__registers_per_thread": "regs",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active": "fp64_pipe_pct",
    "sm__inst_executed.avg.pct_of_peak_sustained_elapsed": "issue_active_pct", "smsp__inst_executed.sum": "warp_insts",
    "l1tex__t_sector_hit_rate.pct": "l1_hit_pct", "lts__t_sector_hit_rate.pct": "l2_hit_pct", "launch__grid_size": "grid",
    "launch__occupancy_limit_registers": "occ_limit_regs_blocks", "sm__maximum_warps_per_active_cycle_pct": "theoretical_occ_pct",
}
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e3, "us": 1.0, "ns": 1e-3, "second": 1e6}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    recs = []
    for r in rows[2:]:
        rec = {"kernel": r[idx["Kernel Name"]].split("(")[0]}
        for k, name in KEYS.items():
            if k in idx and r[idx[k]] != "":
                v = float(r[idx[k]].replace(",", ""))
                u = units[idx[k]]
                if name.startswith("dram_r") or name.startswith("dram_w") or name == "duration_us":
                    v *= UNIT.get(u, 1.0)
                rec[name] = v
        if "dram_read" in rec:
            rec["dram_bytes"] = rec["dram_read"] + rec.get("dram_write", 0.0)
            if rec.get("duration_us"):
                rec["dram_gbs"] = rec["dram_bytes"] / rec["duration_us"] / 1e3
        recs.append(rec)
    json.dump(recs, open(out, "w"), indent=1)
    for r in recs:
        print(f"{r['kernel']:32s} {r.get('duration_us', 0):9.1f} us  dram {r.get('dram_bytes', 0) / 1e9:6.3f} GB ({r.get('dram_gbs', 0):6.0f} GB/s)  "
              f"fp64 {r.get('fp64_pipe_pct', 0):5.1f}%  issue {r.get('issue_active_pct', 0):5.1f}%  warps {r.get('warps_active_pct', 0):5.1f}%  regs {int(r.get('regs', 0))}")


if __name__ == "__main__":
    main()
