"""Joins an ncu launch list (gpu__time_duration.sum, launch order) with the library's launch-order labels (SDB_LABEL_LOG) and
prints the tensor-core GEMM / attention launches of the LAST UNet step + decode with their shapes and ncu durations.

  SDB_LABEL_LOG=gpurun_out/labels.tsv ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv \\
      --log-file gpurun_out/launches.csv python profiles/profile_step.py 2
  python profiles/join_layers.py gpurun_out/launches.csv gpurun_out/labels.tsv > profiles/rN_gemm_layers.md
"""
import collections
import re
import sys

sys.path.insert(0, __file__.rsplit("/", 1)[0])
from summarize_launches import load  # noqa: E402


def main():
    rows = load(sys.argv[1])
    labels = [l.rstrip("\n").split("\t") for l in open(sys.argv[2])]
    for cls, pat in (("gemm_tc", "gemm_tc"), ("attention", "attention_kernel")):
        k = [r for r in rows if pat in r["name"]]
        lb = [l[1] for l in labels if l[0] == cls]
        n = min(len(k), len(lb))
        k, lb = k[-n:], lb[-n:]
        # profile_step.py 2 = two UNet steps + one decode: keep the second step + decode (second half of the list, roughly)
        per = collections.OrderedDict()
        half = n // 2 if cls == "attention" else 0
        for r, l in list(zip(k, lb))[half:]:
            key = re.sub(r" tile=.*", "", l)
            e = per.setdefault(key, [0, 0.0, r["name"], r["grid"]])
            e[0] += 1
            e[1] += r["gpu__time_duration.sum"]
        tot = sum(v[1] for v in per.values())
        print(f"## {cls}: {sum(v[0] for v in per.values())} launches, {tot:.0f} us (ncu, warm caches)\n")
        print("| total us | launches | avg us | shape | kernel | grid |\n|---|---|---|---|---|---|")
        for key, (cnt, us, name, grid) in sorted(per.items(), key=lambda x: -x[1][1]):
            print(f"| {us:.0f} | {cnt} | {us / cnt:.1f} | {key} | `{name}` | {grid} |")
        print()


if __name__ == "__main__":
    main()
