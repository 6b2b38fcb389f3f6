"""Summarise an .ncu-rep (raw page) into profiles/<tag>_ncu_summary.md and update
profiles/ncu_traffic.json (per-launch DRAM bytes of each hot kernel, read by bench.py)."""
import csv, json, os, re, subprocess, sys, collections
rep, tag = sys.argv[1], sys.argv[2]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
# accepts the report itself or the CSV of its raw page (`ncu -i x.ncu-rep --page raw --csv`, made on the GPU box
# when the report is too large to travel)
raw = open(rep).read() if rep.endswith(".csv") else \
    subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
idx = {h: i for i, h in enumerate(hdr)}
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
     "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
     "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
     "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
     "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
     "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_srcunit_tex_op_read.sum",
     "smsp__issue_active.avg.pct_of_peak_sustained_active", "inst_executed",
     "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "LTS.TriageCompute.lts__throughput.avg.pct_of_peak_sustained_elapsed"]
def val(d, m):
    try:
        v = float(d[idx[m]].replace(",", ""))
    except ValueError:          # "", "no data"
        return float("nan")
    u = units[idx[m]].split("/")[0]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "s": 1e6,
             "msecond": 1e3, "usecond": 1, "nsecond": 1e-3, "second": 1e6}.get(u, 1)
    return v * scale
def kname(d):
    name = re.sub(r"^void ", "", d[idx["Kernel Name"]]); name = re.sub(r"\(.*", "", name)
    name = name.replace("ab200::", "")
    # the 8B gate (147 CTAs) and the MLP proj (128), the cell at 16 384 and at 4 096 rows: different problems
    if name.startswith("gemm_tn_sm100<0") or name.startswith("sum_lstm_cell_bf16x2_kernel<2"):
        name += f" grid {d[idx['launch__grid_size']]}"
    return name
def fnum(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0
agg = collections.OrderedDict()
for d in data:
    agg.setdefault(kname(d), []).append({m: v for m in M if m in idx for v in [val(d, m)] if v == v})
lines = [f"# ncu --set full summary ({tag})", "",
         "Source: `ncu --set full --clock-control none --import-source on` on one B200 (sm_100a); per-launch "
         "averages over the captured launches. Durations are cold-cache and serialised (ncu replays each "
         "kernel ~40x): compare shares, not absolutes, with the CUDA-event numbers in BENCH.", "",
         "| kernel | launches | time us | DRAM read MB | DRAM write MB | DRAM % of peak | tensor pipe % | SM % | L2 hit % | regs | grid x block | dyn smem KB |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|"]
extra = ["", "| kernel | dram__cycles_active % of elapsed | L2 -> SM read MB (lts__t_sectors_srcunit_tex_op_read x 32 B) | issue slots busy % | warp instructions M |",
         "|---|---|---|---|---|"]
traffic_path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
for name, ls in agg.items():
    a = lambda m: (sum(l[m] for l in ls if m in l) / max(1, sum(1 for l in ls if m in l)))
    lines.append(f"| `{name}` | {len(ls)} | {a(M[0]):.1f} | {a(M[1])/1e6:.2f} | {a(M[2])/1e6:.2f} | {a(M[3]):.1f} | "
                 f"{a(M[4]):.1f} | {a(M[5]):.1f} | {a(M[6]):.1f} | {a(M[8]):.0f} | {a(M[9]):.0f} x {a(M[10]):.0f} | {a(M[11])/1e3:.1f} |")
    extra.append(f"| `{name}` | {a(M[16]):.1f} | {a(M[13]) * 32 / 1e6:.1f} | {a(M[14]):.1f} | {a(M[15]) / 1e6:.2f} |")
    traffic[name] = {"dram_bytes_per_launch": a(M[1]) + a(M[2]), "time_us": a(M[0]), "source": f"{tag}: {os.path.basename(rep)}"}
lines += extra
# top warp-stall reasons per kernel (share of the sampled stall cycles)
stall_cols = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
if stall_cols:
    lines += ["", "Top warp-stall reasons (warps stalled per issue-active cycle):", ""]
    for name, ls in agg.items():
        rows_k = [d for d in data if kname(d) == name]
        tot = {c: sum(fnum(d[idx[c]]) for d in rows_k) / len(rows_k) for c in stall_cols}
        top = sorted(tot.items(), key=lambda kv: -kv[1])[:4]
        lines.append(f"* `{name}`: " + ", ".join(f"{c[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]} {v:.2f}" for c, v in top))
if len(sys.argv) > 3:
    lines += ["", f"Binary: libarctic_b200.so sha256 {sys.argv[3]} (printed by tools/ncu_target.py in the same run)."]
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.md"), "w").write("\n".join(lines) + "\n")
json.dump(traffic, open(traffic_path, "w"), indent=1)
print("\n".join(lines))
