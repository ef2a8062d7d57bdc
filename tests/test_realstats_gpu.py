"""Stress parity on REALISTIC-STATISTICS weights (the precision policy was budgeted on i.i.d. uniform weights only).

No SD-v1.4 checkpoint exists offline; synth.realistic_stats reshapes the synthetic stream towards a trained checkpoint's
statistics: log-normal per-channel gains (outlier channels 3-5x the rest, residual stream max ~65), GroupNorm / LayerNorm gamma
in [0.4, 1.6], beta +-0.4, attention query / key weights x 1.7 (peaked softmax). Weights go in through sdb_set_tensor (the
product path a dump-dir load takes).

With single fp16 q / k operands in the fused attention, the oracle's operand-rounding emulation of the pass policy predicts
1.17e-3 / 9.0e-4 on these two cases (the attention operands alone inject 8.5e-4): that finding is why the attention of the 3-pass
levels now takes q / k as fp16 hi + lo pairs (3-term split QK^T). Measured with it: 7.9e-4 / 6.4e-4 rel L2 at the default policy,
3.7e-4 / 2.7e-4 with precision = 3 - inside the 1e-3 north-star bar, which this test asserts. DESIGN.md "precision" has the table.
"""
import os

import numpy as np
import pytest

from stable_diffusion_burn_b200 import synth, topology

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b)), float(np.abs(a - b).max() / np.abs(b).max())


def test_unet_realistic_statistics(ctx):
    g = np.load(os.path.join(GOLD, "unet_realstats.npz"))
    params = synth.realistic_stats(synth.make_params(0, topology.unet_params()))
    ctx.init_synthetic(0)
    try:
        for name, arr in params.items():
            ctx.set_tensor(name, arr)
        ctx.finalize_weights()
        res = {}
        for prec in (0, 3):
            ctx.set_option("precision", prec)
            a = rel(ctx.unet_forward(synth.make_latent(2, 32, 32, seed=7), 321, synth.make_context(2, 13, seed=5)), g["b2_32"])
            b = rel(ctx.unet_forward(synth.make_latent(1, 64, 64), 999, synth.make_context(1, 77)), g["n1_64_L77"])
            res[prec] = (a, b)
            print(f"realistic-statistics weights, precision option {prec}: n=2 32x32 rel L2 {a[0]:.3e} max {a[1]:.3e}; "
                  f"n=1 64x64 L=77 rel L2 {b[0]:.3e} max {b[1]:.3e}")
        assert max(max(v) for pair in res[0] for v in [pair]) < 1e-3
        assert max(max(v) for pair in res[3] for v in [pair]) < 6e-4
    finally:
        ctx.set_option("precision", 0)
        ctx.init_synthetic(0)
        ctx.finalize_weights()
