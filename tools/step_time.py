"""A/B timing inside one process: 20-step sample_latent (graph replay) and decode, CUDA-event timed on the library's
device entry points, for a list of option settings. Usage: [BATCH=n REPS=7 ROUNDS=2] python tools/step_time.py [key=v,key=v ...]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from stable_diffusion_burn_b200 import _lib, synth

variants = [dict(kv.split("=") for kv in a.split(",") if kv) for a in sys.argv[1:]] or [{}]
c = _lib.Context(0)
c.init_synthetic(0)
c.finalize_weights()
dev = torch.device("cuda:0")
n, H, L = int(os.environ.get("BATCH", 1)), 64, 77
ctx = torch.from_numpy(synth.make_context(n, L)).to(dev)
unc = torch.from_numpy(synth.make_context(1, 2, seed=99)[0]).to(dev)
lat = torch.from_numpy(synth.make_latent(n, H, H)).to(dev)
rgb = torch.empty((n, 8 * H, 8 * H, 3), dtype=torch.uint8, device=dev)
img = torch.empty((n, 3, 8 * H, 8 * H), dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream


def sample(steps):
    c.check(c.lib.sdb_sample_image_dev(c.h, C.c_void_p(ctx.data_ptr()), n, L, C.c_void_p(unc.data_ptr()), 2, 7.5, steps,
                                       C.c_void_p(lat.data_ptr()), H, H, C.c_void_p(rgb.data_ptr()), C.c_void_p(st)))


def decode():
    c.check(c.lib.sdb_decode_latent_dev(c.h, C.c_void_p(lat.data_ptr()), n, H, H, C.c_void_p(img.data_ptr()), C.c_void_p(st)))


def timeit(fn, reps):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


REPS, ROUNDS = int(os.environ.get("REPS", 7)), int(os.environ.get("ROUNDS", 2))
for rnd in range(ROUNDS):
    for v in variants:
        for k, val in v.items():
            c.set_option(k, int(val))
        for _ in range(2):
            sample(20)
        t20 = timeit(lambda: sample(20), REPS)
        td = timeit(decode, REPS)
        print(f"round {rnd} {v}: image min/med {t20[0]:.2f}/{t20[1]:.2f} ms; decode {td[0]:.2f}/{td[1]:.2f} ms; "
              f"unet step ~{(t20[0] - td[0]) / 20:.3f} ms", flush=True)
