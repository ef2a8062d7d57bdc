"""Dev tool: budget the tensor-core operand format against the 1e-3 UNet-step tolerance.
Runs the fp32 oracle and operand-rounded variants (fp16 / tf32 / bf16) on one UNet eval."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from stable_diffusion_burn_b200 import synth, topology
from oracle import sd_oracle as O

torch.set_num_threads(os.cpu_count())
t0 = time.time()
arrs = synth.make_params(0, topology.unet_params())
print("params", time.time() - t0, flush=True)
P = O.Params(arrs)
x = torch.from_numpy(synth.make_latent(1, 64, 64))
ctx = torch.from_numpy(synth.make_context(1, 13))
res = {}
for mode in [None, "fp16", "tf32", "bf16"]:
    O.set_emulation(mode)
    t0 = time.time()
    with torch.no_grad():
        taps = {}
        y = O.unet_forward(P, x, 999, ctx, taps=taps)
    res[mode] = (y, taps)
    print(mode, "time", time.time() - t0, "out rms", float(y.pow(2).mean().sqrt()), flush=True)
    if mode is not None:
        ref, rt = res[None]
        print("   rel L2 err", float((y - ref).norm() / ref.norm()), "maxabs/maxref", float((y - ref).abs().max() / ref.abs().max()))
        for k in ["input_blocks/rt2", "input_blocks/r2", "middle_block", "output_blocks/rt2", "output_blocks/rt7"]:
            print("     ", k, float((taps[k] - rt[k]).norm() / rt[k].norm()), "rms", float(rt[k].pow(2).mean().sqrt()))
