import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stable_diffusion_burn_b200 import _lib
def rel(a,b): return float(np.linalg.norm(a.astype(np.float64)-b)/np.linalg.norm(b))
c=_lib.Context(0)
rng=np.random.default_rng(0)
for (M,K,N) in [(128,11520,1280),(512,11520,1280),(128,1280,1280),(2048,5760,640),(128,23040,1280)]:
    a=rng.standard_normal((M,K)).astype(np.float16).astype(np.float32)       # fp16-exact operands:
    w=(rng.standard_normal((K,N))*K**-0.5).astype(np.float16).astype(np.float32)  # only accumulation error remains
    ref=a.astype(np.float64)@w.astype(np.float64)
    out={}
    for sk in (1,0):
        c.set_option("splitk",sk)
        out[sk]=c.test_linear(a,w,None,passes=1)
    print(f"M{M} K{K} N{N}: err split-K on {rel(out[1],ref):.3e}  off {rel(out[0],ref):.3e}  on-vs-off {rel(out[1],out[0].astype(np.float64)):.3e}  mean signed err/|ref| off {float((out[0]-ref).mean()/np.abs(ref).mean()):.3e}")
