"""Per-launch table of one UNet step + decode in eager mode with CUDA-event timing (warm caches):
   SDB_PROFILE_DUMP=gpurun_out/layers.tsv python profiles/layer_table.py [cluster]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stable_diffusion_burn_b200 import _lib, synth  # noqa: E402

ctx = _lib.Context(0)
ctx.init_synthetic(0)
ctx.finalize_weights()
ctx.set_option("graphs", 0)
if len(sys.argv) > 1:
    ctx.set_option("cluster", int(sys.argv[1]))
c = synth.make_context(1, 77)
u = synth.make_context(1, 2, seed=99)[0]
lat = synth.make_latent(1, 64, 64)
ctx.sample_image(c, u, 7.5, 2, init_latent=lat)  # warm-up
ctx.profile(True)
ctx.profile_reset()
ctx.sample_image(c, u, 7.5, 1, init_latent=lat)
t = ctx.profile_table()
ctx.profile(False)
print({k: (v["launches"], round(v["ms"], 2)) for k, v in t.items()})
