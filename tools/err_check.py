"""Errors of the UNet fixtures under option settings: python tools/err_check.py key=v,key=v [key=v ...]   (GPU box)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from stable_diffusion_burn_b200 import _lib, synth

G = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
variants = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[1:]] or [{}]
c = _lib.Context(0)
c.init_synthetic(0)
c.finalize_weights()


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b)), float(np.abs(a - b).max() / np.abs(b).max())


cases = {
    "kat_zeros": (np.zeros((1, 4, 64, 64), np.float32), 1, synth.kat_context()),
    "sin_ramp": (synth.sin_ramp((1, 4, 64, 64)), 500, synth.make_context(1, 13)),
    "randn_t999": (synth.make_latent(1, 64, 64), 999, synth.make_context(1, 13)),
    "batch2_32": (synth.make_latent(2, 32, 32, seed=7), 321, synth.make_context(2, 5, seed=5)),
    "768px": (synth.make_latent(1, 96, 96, seed=96), 777, synth.make_context(1, 9, seed=96)),
    "b8_64": (synth.make_latent(8, 64, 64, seed=808), 599, synth.make_context(8, 77, seed=88)),
}
for v in variants:
    for k, val in v.items():
        c.set_option(k, int(val))
    out = []
    worst = 0.0
    for name, (x, t, ctx) in cases.items():
        g = np.load(os.path.join(G, f"unet_{name}.npz"))["out"]
        e2, em = rel(c.unet_forward(x, t, ctx), g)
        worst = max(worst, e2, em)
        out.append(f"{name} {e2:.2e}/{em:.2e}")
    g = np.load(os.path.join(G, "cfg_L77.npz"))
    for t in (999, 449):
        _, u, cc = c.forward_diffuser(synth.make_latent(1, 64, 64), t, synth.make_context(1, 77), synth.make_context(1, 2, seed=99)[0], 7.5)
        for nm, a, b in (("u", u, g[f"t{t}:uncond"]), ("c", cc, g[f"t{t}:cond"])):
            e2, em = rel(a, b)
            worst = max(worst, e2, em)
            out.append(f"cfg{t}{nm} {e2:.2e}/{em:.2e}")
    print(v, "WORST", f"{worst:.2e}", "|", "  ".join(out), flush=True)
