"""Turn the raw artefacts a gpurun call leaves in gpurun_out/ into the tracked
summaries under profiles/:

    python scripts/summarize_profiles.py <tag>          e.g.  r1h

reads   gpurun_out/launches_<tag>.csv        (ncu --metrics gpu__time_duration.sum launch list)
        gpurun_out/prof_<tag>_*.ncu-rep      (ncu --set full captures of the top kernels)
        gpurun_out/op_profile_<tag>.json     (per-op CUDA-event table written by bench.py)
        gpurun_out/bench_<tag>*.json         (bench lines)
writes  profiles/<tag>_launches.md, profiles/<tag>_ncu_top_kernels.json,
        profiles/<tag>_ops.md, profiles/<tag>_bench.jsonl, profiles/ncu_traffic.json
"""
import csv
import glob
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
os.makedirs(PROF, exist_ok=True)


def launches():
    path = os.path.join(OUT, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return
    rows = list(csv.reader(open(path)))
    h = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    hdr, data = rows[h], rows[h + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = {}, {}
    for r in data:
        if len(r) <= vi:
            continue
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "us" else v / 1e6 if r[ui] == "ns" else v
        k = r[ki].split("(")[0].replace("void ", "")
        tot[k] = tot.get(k, 0) + v
        cnt[k] = cnt.get(k, 0) + 1
    T = sum(tot.values())
    with open(os.path.join(PROF, f"{tag}_launches.md"), "w") as f:
        f.write(f"# ncu launch list, one step of `bench.py --amps 1024 --no-graph` ({tag})\n\n"
                "`ncu --metrics gpu__time_duration.sum --clock-control none` -- cold-cache, serialised launches: "
                "compare SHARES, not absolutes (in production the step is a CUDA graph whose independent "
                "branches overlap).\n\n"
                f"total {T:.3f} ms over {sum(cnt.values())} launches\n\n| kernel | launches | ms | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda x: -x[1]):
            f.write(f"| `{k}` | {cnt[k]} | {v:.3f} | {100 * v / T:.1f}% |\n")
    print("launch list:", T, "ms")


WANT = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"]


def ncu_reports():
    out = []
    reps = sorted(glob.glob(os.path.join(OUT, f"prof_{tag}_*.ncu-rep")))
    csvs = sorted(glob.glob(os.path.join(OUT, f"prof_{tag}_*.raw.csv")))     # `ncu -i X --page raw --csv` done on the box
    done = {os.path.basename(c)[:-8] for c in csvs}
    for rep in csvs + [r for r in reps if os.path.basename(r)[:-8] not in done]:
        if rep.endswith(".csv"):
            txt = open(rep).read()
        else:
            txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        idx = [hdr.index(w) if w in hdr else None for w in WANT]
        for r in rows[2:]:
            d = {}
            for w, i in zip(WANT, idx):
                if i is None:
                    continue
                d[w] = r[i] + (" " + units[i] if units[i] and w != "Kernel Name" else "")
            d["report"] = os.path.basename(rep)
            out.append(d)
    if out:
        json.dump(out, open(os.path.join(PROF, f"{tag}_ncu_top_kernels.json"), "w"), indent=1)
        # per-launch DRAM traffic of the biggest kernel -> bench.py's roofline.traffic
        def gb(s):
            v, u = s.split()
            return float(v) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u]
        big = [d for d in out if "dram__bytes_write.sum" in d]
        tot = lambda d: gb(d["dram__bytes_read.sum"]) + gb(d["dram__bytes_write.sum"])
        mx = max(tot(d) for d in big)
        top = [d for d in big if tot(d) > 0.1 * mx]          # the dominant launches among the captured ones
        traffic = sum(tot(d) for d in top) / len(top)
        amps = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
        json.dump({"rqc_7x7_d20_c64_s4096": {"bytes_per_launch_mean_dominant": traffic, "at_amps": amps,
                                             "n_launches": len(top),
                                             "source": f"profiles/{tag}_ncu_top_kernels.json"}},
                  open(os.path.join(PROF, "ncu_traffic.json"), "w"), indent=1)
    print("ncu kernels:", len(out))


def ops():
    path = os.path.join(OUT, f"op_profile_{tag}.json")
    if not os.path.exists(path):
        return
    p = json.load(open(path))
    o = [x for v in p["variants"] for x in v["ops"]]
    o.sort(key=lambda x: -x["ms"])
    T = sum(x["ms"] for x in o)
    with open(os.path.join(PROF, f"{tag}_ops.md"), "w") as f:
        f.write(f"# per-op CUDA-event profile ({tag}, dtype {p['dtype']})\n\nserial launches with an event pair per op "
                f"(profile mode); {len(o)} ops, sum {T:.3f} ms\n\n"
                "| op | nC | nK | batch/M/N bits | ms | GB/s (algorithmic) | GFLOP/s |\n|---|---:|---:|---|---:|---:|---:|\n")
        for x in o[:25]:
            f.write(f"| {x['name']} | {x['nC']} | {x['nK']} | {x['batch_bits']}/{x['m_bits']}/{x['n_bits']} | {x['ms']:.3f} | "
                    f"{x['bytes'] / x['ms'] / 1e6:.0f} | {x['flops'] / x['ms'] / 1e6:.0f} |\n")
    print("ops:", len(o))


def bench_lines():
    lines = []
    for path in sorted(glob.glob(os.path.join(OUT, f"bench_{tag}*.json"))):
        for ln in open(path):
            ln = ln.strip()
            if ln.startswith("{"):
                lines.append(json.dumps({"file": os.path.basename(path), **json.loads(ln)}))
    if lines:
        open(os.path.join(PROF, f"{tag}_bench.jsonl"), "w").write("\n".join(lines) + "\n")
    print("bench lines:", len(lines))


launches(); ncu_reports(); ops(); bench_lines()
