"""One row per `ncu --set full` capture (raw-page CSVs from profiles/collect_r2.sh): duration, tensor-pipe %, DRAM GB/s and % of
the measured HBM peak, L2->SM bytes, MUFU (XU) pipe %, registers, occupancy, top stall reasons.
  python profiles/kernel_table.py gpurun_out/r2f_*.raw.csv > profiles/r2_kernels_ncu.md"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    PEAK = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def load(path):
    rows = list(csv.reader(open(path)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def num(d, key, default=0.0):
    if key not in d:
        return default
    v, u = d[key]
    try:
        x = float(v.replace(",", ""))
    except ValueError:
        return default
    mult = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "us": 1.0, "ns": 1e-3, "ms": 1e3, "Gbyte/s": 1.0, "Tbyte/s": 1e3, "Mbyte/s": 1e-3}.get(u, 1.0)
    return x * mult


def main():
    print("# Round 2 — per-kernel ncu captures (`ncu --set full --clock-control none`, one launch each, second UNet step / decode of "
          "`profiles/profile_step.py 2`; raw pages extracted on the GPU box by `profiles/collect_r2.sh`)\n")
    print(f"Peaks: HBM {PEAK['hbm_gbs']:.0f} GB/s and bf16 {PEAK['bf16_tflops']:.0f} TFLOP/s burst (MEASURED_PEAKS.json, a kernel timed alone). "
          "Durations under `--set full` are cold-cache and serialised; compare shares, not absolutes.\n")
    print("| capture | kernel | grid | us | tensor pipe % | DRAM GB/s (% of measured peak) | DRAM MB (rd+wr) | L2->L1 MB | XU (MUFU) pipe % | FMA pipe % | regs | warps active % | top stalls (per issue) |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    for path in sys.argv[1:]:
        d = load(path)
        name = d.get("Kernel Name", ("?", ""))[0].replace("void sdb::", "")
        name = name.split("(")[0]
        us = num(d, "gpu__time_duration.sum")
        rd, wr = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
        gbs = (rd + wr) / 1e9 / (us * 1e-6) if us else 0.0
        lts = num(d, "lts__t_bytes_srcunit_tex.sum") or num(d, "lts__t_sectors_srcunit_tex.sum") * 32
        stalls = {k.split("issue_stalled_")[1].split("_per_issue")[0]: num(d, k) for k in d if "smsp__average_warps_issue_stalled_" in k and "_per_issue_active" in k and "not_issued" not in k}
        top = sorted(stalls.items(), key=lambda kv: -kv[1])[:3]
        print(f"| {os.path.basename(path).replace('r2f_', '').replace('.raw.csv', '')} | `{name}` | {d.get('Grid Size', ('', ''))[0]} | {us:.1f} | "
              f"{num(d, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {gbs:.0f} ({100 * gbs / PEAK['hbm_gbs']:.0f} %) | {(rd + wr) / 1e6:.1f} | {lts / 1e6:.1f} | "
              f"{num(d, 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active'):.1f} | {num(d, 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active'):.1f} | "
              f"{num(d, 'launch__registers_per_thread'):.0f} | {num(d, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.1f} | "
              + ", ".join(f"{k} {v:.1f}" for k, v in top) + " |")


if __name__ == "__main__":
    main()
