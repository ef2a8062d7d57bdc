"""Dev tool (test infrastructure, like the rest of oracle/): CPU design studies on the UNet-step error budget, using the oracle's
operand-rounding emulation with a per-block policy (`set_emulation_policy`). Prints rel-L2 / max errors of one UNet step against the
fp32 oracle for: the pass policy in use (3-term products on levels 0-1, single fp16 pass on levels 2-3), cheaper weight policies on
levels 0-1, and LayerNorm folded into the consuming GEMM as a rank-1 epilogue correction (the GEMM reads the raw, rounded x).
Results of round 1 are quoted in DESIGN.md §4 (findings 1 and 3).   python oracle/policy_study.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from oracle import sd_oracle as O  # noqa: E402
from stable_diffusion_burn_b200 import synth, topology  # noqa: E402

IN = ["conv", "rt1", "rt2", "d1", "rt3", "rt4", "d2", "rt5", "rt6", "d3", "r1", "r2"]
OUT = ["r1", "r2", "ru", "rt1", "rt2", "rtu1", "rt3", "rt4", "rtu2", "rt5", "rt6", "rt7"]
LEVEL = {f"input_blocks/{f}": l for f, l in zip(IN, [0, 0, 0, 0, 1, 1, 1, 2, 2, 2, 3, 3])}
LEVEL["middle_block"] = 3
LEVEL.update({f"output_blocks/{f}": l for f, l in zip(OUT, [3, 3, 3, 2, 2, 2, 1, 1, 1, 0, 0, 0])})


def policy(hi_roles, lo_roles="awqAW"):
    """roles rounded to fp16 on levels 0-1 / levels 2-3 ('q' = attention operands, 'a'/'w' Linear, 'A'/'W' conv)."""
    pol = {b: (hi_roles if l <= 1 else lo_roles) for b, l in LEVEL.items()}
    pol["emb"] = pol["out"] = ""
    return pol


def main():
    torch.set_num_threads(os.cpu_count())
    P = O.Params(synth.make_params(0, topology.unet_params()))
    cases = {"randn_t999": (torch.from_numpy(synth.make_latent(1, 64, 64)), 999, torch.from_numpy(synth.make_context(1, 13))),
             "sin_ramp": (torch.from_numpy(synth.sin_ramp((1, 4, 64, 64))), 500, torch.from_numpy(synth.make_context(1, 13)))}
    variants = [("policy in use (3-term products on levels 0-1)", policy("q"), False),
                ("2-term products on levels 0-1: conv weights in fp16", policy("qW"), False),
                ("2-term products on levels 0-1: all weights in fp16", policy("qwW"), False),
                ("single pass everywhere", policy("awqAW"), False),
                ("policy in use + LayerNorm folded into the consuming GEMMs", policy("q"), True)]
    for cname, (x, t, ctx) in cases.items():
        O.set_emulation(None), O.set_emulation_policy(None)
        O._EMU["ln_fused"] = False
        with torch.no_grad():
            ref = O.unet_forward(P, x, t, ctx)
        O.set_emulation("fp16")
        for vname, pol, fused in variants:
            O.set_emulation_policy(pol)
            O._EMU["ln_fused"], O._EMU["pre_rounded"], O._EMU["keep"] = fused, set(), []
            with torch.no_grad():
                y = O.unet_forward(P, x, t, ctx)
            print(f"{cname:11s} {vname:62s} rel L2 {float((y - ref).norm() / ref.norm()):.3e}  "
                  f"max {float((y - ref).abs().max() / ref.abs().max()):.3e}", flush=True)


if __name__ == "__main__":
    main()
