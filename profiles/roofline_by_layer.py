"""Per-layer roofline fractions from the committed per-layer launch table (profiles/rN_gemm_layers.md, ncu warm durations joined with
the library's shape labels) and the measured peaks (MEASURED_PEAKS.json, else the fallback of B200_PROFILING.md).

  python profiles/roofline_by_layer.py profiles/r2_gemm_layers.md > profiles/r2_roofline_by_layer.md

gemm rows: algorithmic FLOPs = 2 M N (K + xk) [x 4 phases for the folded-upsample conv, kind=4], issued = x passes; bytes = the
algorithmic bytes of DESIGN.md §4 (fp16 A, hi + lo when passes >= 2; fp16 W, hi + lo when passes == 3; outputs as the epilogue label
says are unknown here, so the byte column counts operands only and is a LOWER bound on the traffic).
attention rows: 4 nb heads Nq Nk d FLOPs (the split QK^T issues twice the QK^T half)."""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
try:
    P = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    PEAK_BURST, PEAK_SUST, HBM, SRC = P["bf16_tflops"], P["bf16_tflops_sustained"], P["hbm_gbs"], "MEASURED_PEAKS.json"
except Exception:
    PEAK_BURST, PEAK_SUST, HBM, SRC = 1590.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def kv(s):
    return {k: v for k, v in re.findall(r"(\w+)=([^\s|]+)", s)}


def main():
    rows_g, rows_a = [], []
    for line in open(sys.argv[1]):
        c = [x.strip() for x in line.strip().strip("|").split("|")]
        if len(c) < 6 or not c[0].replace(".", "").isdigit():
            continue
        tot, n, avg, shape = float(c[0]), int(c[1]), float(c[2]), c[3]
        d = kv(shape)
        if shape.startswith("gemm"):
            M = int(d["n"]) * int(d["H"]) * int(d["W"])
            N, K, xk, passes, kind = int(d["N"]), int(d["K"]), int(d["xk"]), int(d["passes"]), int(d["kind"])
            ph = 4 if kind == 4 else 1
            fl = 2.0 * M * N * (K + xk) * ph
            taps = 9 if kind in (2, 3) else (4 if kind == 4 else 1)  # K = taps x Cin: every activation element counted once
            by = M * (K + xk) / taps * 2 * (2 if passes >= 2 else 1) + N * (K + xk) * ph * 2 * (2 if passes == 3 else 1)
            rows_g.append((tot, n, avg, M, N, K + xk, passes, kind, fl, by, d.get("epi", "-"), c[5]))
        elif shape.startswith("attention"):
            fl = 4.0 * int(d["nb"]) * int(d["heads"]) * int(d["Nq"]) * int(d["Nk"]) * int(d["d"])
            rows_a.append((tot, n, avg, d, fl, c[4]))
    print(f"# Per-layer roofline fractions (derived from `{os.path.basename(sys.argv[1])}`: ncu warm durations x shape labels)\n")
    print(f"Peaks ({SRC}): bf16/fp16 tensor {PEAK_BURST:.0f} TFLOP/s burst (a kernel timed alone, the denominator used here), "
          f"{PEAK_SUST:.0f} sustained; HBM {HBM:.0f} GB/s. `alg` = algorithmic FLOPs (one product per MAC), `issued` = x passes of the "
          f"split-fp16 product. The byte column counts fp16 operand bytes only (each activation element once, not per tap): a lower "
          f"bound, useful to see which launches are weight-streaming bound.\n")
    print("## gemm_tc\n")
    print("| avg us | launches | M | N | K | passes | conv | alg TFLOP/s | issued TFLOP/s | issued / peak | operand GB/s | / HBM peak | grid |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    tf = ti = tt = 0.0
    for tot, n, avg, M, N, K, p, kind, fl, by, epi, grid in sorted(rows_g, key=lambda r: -r[0]):
        a = fl / (avg * 1e-6) / 1e12
        print(f"| {avg:.1f} | {n} | {M} | {N} | {K} | {p} | {'3x3' if kind in (2, 3) else ('up2' if kind == 4 else '1x1/lin')} | {a:.0f} | {a * p:.0f} | "
              f"{a * p / PEAK_BURST:.2f} | {by / (avg * 1e-6) / 1e9:.0f} | {by / (avg * 1e-6) / 1e9 / HBM:.2f} | {grid} |")
        tf += fl * n
        ti += fl * p * n
        tt += tot
    print(f"\nAll gemm_tc launches of the list: {tf / 1e12:.2f} TFLOP algorithmic, {ti / 1e12:.2f} issued in {tt / 1e3:.2f} ms -> "
          f"{tf / tt / 1e6:.0f} TFLOP/s algorithmic ({tf / tt / 1e6 / PEAK_BURST:.3f} of the burst peak), {ti / tt / 1e6:.0f} issued "
          f"({ti / tt / 1e6 / PEAK_BURST:.3f}).\n")
    print("## attention\n")
    print("| avg us | launches | nb | heads | d | Nq | Nk | alg TFLOP/s | / peak | exponentials per us per SM (MUFU bound: 16 per clk = 31.4 k) |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for tot, n, avg, d, fl, grid in sorted(rows_a, key=lambda r: -r[0]):
        a = fl / (avg * 1e-6) / 1e12
        ex = int(d["nb"]) * int(d["heads"]) * int(d["Nq"]) * max(int(d["Nk"]), 128) / avg / 148
        print(f"| {avg:.1f} | {n} | {d['nb']} | {d['heads']} | {d['d']} | {d['Nq']} | {d['Nk']} | {a:.0f} | {a / PEAK_BURST:.2f} | {ex / 1e3:.1f} k |")


main()
