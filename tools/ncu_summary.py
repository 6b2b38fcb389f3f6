"""Summarise an .ncu-rep (raw page) into profiles/<tag>_ncu_summary.md and update
profiles/ncu_traffic.json (per-launch DRAM bytes of each hot kernel, read by bench.py)."""
import csv, json, os, re, subprocess, sys, collections
rep, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
     "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
     "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]
def val(d, m):
    v = float(d[idx[m]].replace(",", "")); u = units[idx[m]].split("/")[0]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6,
             "msecond": 1e3, "usecond": 1, "nsecond": 1e-3, "second": 1e6}.get(u, 1)
    return v * scale
agg = collections.OrderedDict()
for d in data:
    name = re.sub(r"^void ", "", d[idx["Kernel Name"]]); name = re.sub(r"\(.*", "", name)
    name = name.replace("ab200::", "")
    agg.setdefault(name, []).append({m: val(d, m) for m in M if m in idx})
lines = [f"# ncu --set full summary ({tag})", "",
         "Source: `ncu --set full --clock-control none --import-source on` on one B200 (sm_100a); per-launch "
         "averages over the captured launches. Durations are cold-cache and serialised (ncu replays each "
         "kernel ~40x): compare shares, not absolutes, with the CUDA-event numbers in BENCH.", "",
         "| kernel | launches | time us | DRAM read MB | DRAM write MB | DRAM % of peak | tensor pipe % | SM % | L2 hit % | regs | grid x block | dyn smem KB |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|"]
traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for name, ls in agg.items():
    a = lambda m: sum(l[m] for l in ls) / len(ls)
    lines.append(f"| `{name}` | {len(ls)} | {a(M[0]):.1f} | {a(M[1])/1e6:.2f} | {a(M[2])/1e6:.2f} | {a(M[3]):.1f} | "
                 f"{a(M[4]):.1f} | {a(M[5]):.1f} | {a(M[6]):.1f} | {a(M[8]):.0f} | {a(M[9]):.0f} x {a(M[10]):.0f} | {a(M[11])/1e3:.1f} |")
    traffic[name] = {"dram_bytes_per_launch": a(M[1]) + a(M[2]), "time_us": a(M[0]), "source": f"{tag}: {os.path.basename(rep)}"}
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(traffic_path, "w"), indent=1)
print("\n".join(lines))
