"""Reads an `ncu --page raw --csv` export of the fused-chain launch and writes profiles-style JSON: DRAM bytes per launch,
duration, pipe utilisation -- keyed by workload and plan hash (bench.py uses it only for the plan it was taken on)."""
import csv, json, sys
raw, out_path, workload, plan_sha1, at_amps, source = sys.argv[1:7]
rows = list(csv.reader(open(raw)))
hdr = rows[0]
data = [r for r in rows[1:] if r and r[0].isdigit()]
ix = {h: i for i, h in enumerate(hdr)}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}
unit_row = rows[1] if len(rows) > 1 and not rows[1][0].isdigit() else None


def num(r, k):
    """Value of metric k in base units (bytes, nanoseconds): ncu scales every column separately."""
    try:
        v = float(r[ix[k]].replace(",", ""))
    except Exception:
        return None
    u = unit_row[ix[k]] if unit_row else ""
    return v * UNIT.get(u, 1.0)


recs = []
for r in data:
    recs.append({"kernel": r[ix["Kernel Name"]], "grid": r[ix["Grid Size"]] if "Grid Size" in ix else None,
                 "duration_ns": num(r, "gpu__time_duration.sum"),
                 "dram_bytes_read": num(r, "dram__bytes_read.sum"), "dram_bytes_write": num(r, "dram__bytes_write.sum"),
                 "fp64_pipe_pct": num(r, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"),
                 "dram_pct": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                 "shared_wavefronts": num(r, "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
                 "shared_bank_conflicts": num(r, "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
                 "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                 "registers": num(r, "launch__registers_per_thread")})
tot = [x["dram_bytes_read"] + x["dram_bytes_write"] for x in recs if x["dram_bytes_read"] is not None]
res = {workload: {"plan_sha1": plan_sha1, "at_amps": int(at_amps), "source": source, "launches": recs,
                  "units": "bytes, nanoseconds, percent",
                  "bytes_per_launch_mean_dominant": sum(tot) / len(tot) if tot else None}}
json.dump(res, open(out_path, "w"), indent=1)
print(json.dumps(res)[:1500])
