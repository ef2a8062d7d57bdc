"""Turns an ncu launch list (--csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]) of
profiles/profile_step.py into the markdown summary committed under profiles/ and, when the dram metrics are present, the
per-launch gemm_tc traffic JSON bench.py reads.

  python profiles/summarize_launches.py gpurun_out/launches.csv "title" > profiles/rN_launches_summary.md
  python profiles/summarize_launches.py gpurun_out/traffic.csv --traffic profiles/rN_gemm_traffic.json
"""
import collections
import csv
import json
import re
import sys


def load(path):
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    rows = collections.OrderedDict()
    for row in csv.DictReader(lines):
        key = row["ID"]
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        if row["Metric Name"] == "gpu__time_duration.sum":
            v = v / 1000 if u == "ns" else (v * 1000 if u == "ms" else v)  # -> us
        elif u in ("Kbyte", "Mbyte", "Gbyte"):
            v *= {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        rows.setdefault(key, {"name": name, "grid": row["Grid Size"]})[row["Metric Name"]] = v
    return list(rows.values())


def table(seq, title):
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in seq:
        agg[r["name"]][0] += 1
        agg[r["name"]][1] += r["gpu__time_duration.sum"]
    tot = sum(r["gpu__time_duration.sum"] for r in seq)
    out = [f"## {title}: {len(seq)} launches, {tot / 1000:.2f} ms", "", "| kernel | launches | total us | share | avg us |", "|---|---|---|---|---|"]
    for k, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{k}` | {c} | {v:.1f} | {100 * v / tot:.1f}% | {v / c:.1f} |")
    return "\n".join(out)


def main():
    seq = load(sys.argv[1])
    if "--traffic" in sys.argv:
        g = [r for r in seq if "gemm_tc" in r["name"]]
        rd = sum(r.get("dram__bytes_read.sum", 0.0) for r in g)
        wr = sum(r.get("dram__bytes_write.sum", 0.0) for r in g)
        json.dump({"source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:gemm_tc on profiles/profile_step.py",
                   "launches": len(g), "dram_read_bytes": rd, "dram_write_bytes": wr,
                   "traffic_bytes_per_launch": (rd + wr) / max(1, len(g)),
                   "note": "one UNet step (CFG batch 2) + one decode_latent; serialised under ncu"},
                  open(sys.argv[sys.argv.index("--traffic") + 1], "w"), indent=1)
        return
    title = sys.argv[2] if len(sys.argv) > 2 else "launch list"
    ddim = [i for i, r in enumerate(seq) if "cfg_ddim" in r["name"]]
    first = next(i for i, r in enumerate(seq) if "time_embed" in r["name"] or "gemv" in r["name"])
    # with `profile_step.py 2` the list holds two UNet steps: summarise the second (warm weights / instruction caches)
    lo = ddim[-2] + 1 if len(ddim) >= 2 else first
    print(f"# {title}\n")
    print(table(seq[lo:ddim[-1] + 1], "UNet step (CFG batch 2, 64x64 latent, L = 77)"))
    print()
    print(table(seq[ddim[-1] + 1:], "decode_latent + to_rgb8"))
    print("\n## Heaviest single launches\n\n| us | grid | kernel |\n|---|---|---|")
    for r in sorted(seq[lo:], key=lambda r: -r["gpu__time_duration.sum"])[:15]:
        print(f"| {r['gpu__time_duration.sum']:.1f} | {r['grid']} | `{r['name']}` |")


if __name__ == "__main__":
    main()
