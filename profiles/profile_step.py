"""Profiling target for ncu: ONE UNet step exactly as the sampler runs it (batch 2 = uncond + cond, 64x64 latent,
L = 77 / Lu = 2 padded) and ONE decode_latent, eager launches (no CUDA graph) so every kernel is visible.

  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python profiles/profile_step.py
  ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 40 -c 6 -o gpurun_out/gemm python profiles/profile_step.py
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from stable_diffusion_burn_b200 import _lib, synth  # noqa: E402

ctx = _lib.Context(0)
ctx.init_synthetic(0)
ctx.finalize_weights()
ctx.set_option("graphs", 0)
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
c = synth.make_context(1, 77)
u = synth.make_context(1, 2, seed=99)[0]
lat = synth.make_latent(1, 64, 64)
# one DDIM step through the public sampler entry (CFG batch of 2) + decode + u8 pack
rgb = ctx.sample_image(c, u, 7.5, steps, init_latent=lat)
print("launches", ctx.launch_count(), "checksum", int(rgb.astype(np.int64).sum()))
