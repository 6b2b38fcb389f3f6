"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none -c N --csv --log-file x.csv
python bench.py ...`) into profiles/<tag>_launches.md: per kernel launches, average / minimum device time and the share
of our kernels' total.
    python tools/launch_list_summary.py gpurun_out/r2zz_launches.csv r2zz "<command that was profiled>" """
import collections, csv, os, re, sys
src, tag, cmd = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lines = [l for l in open(src) if l.startswith('"')]
rows = list(csv.DictReader(lines))
ours, others = collections.defaultdict(list), 0
for r in rows:
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", re.sub(r"^void ", "", r["Kernel Name"]))
    if name.startswith("ab200::"):
        ours[name].append(float(r["Metric Value"].replace(",", "")) / 1e3)
    else:
        others += 1
total = sum(sum(v) for v in ours.values())
out = [f"# ncu launch list ({tag})" + (f" (`{cmd}`)" if cmd else ""), "",
       "Per-launch device time of every kernel of the bench command (cold-cache, serialised: shares, not absolutes).", "",
       "| kernel (ours) | launches | avg us | min us | share of our kernels in the list |", "|---|---|---|---|---|"]
for name, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
    out.append(f"| `{name}` | {len(v)} | {sum(v) / len(v):.2f} | {min(v):.2f} | {100 * sum(v) / total:.1f} % |")
out += ["", f"Other kernels in the list ({others} launches) are torch's weight-initialisation / fill / copy kernels.  "
        "One draft step = rms_scale + 3 x (gate GEMM + cell + head GEMM/arg-max)."]
path = os.path.join(ROOT, "profiles", f"{tag}_launches.md")
open(path, "w").write("\n".join(out) + "\n")
print("wrote", path)
