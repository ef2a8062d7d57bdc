import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from stable_diffusion_burn_b200 import _lib, synth
def rel(a,b): return float(np.linalg.norm(a.astype(np.float64)-b)/np.linalg.norm(b))
c=_lib.Context(0); c.init_synthetic(0); c.finalize_weights()
x=synth.make_latent(4,32,32,seed=51); ctx=synth.make_context(4,11,seed=41)
full=c.unet_forward(x,500,ctx); full2=c.unet_forward(x,500,ctx)
print("unet n=4 repeat identical:", np.array_equal(full,full2))
for i in range(4):
    one=c.unet_forward(x[i:i+1],500,ctx[i:i+1]); print("unet img",i,"batch vs single rel",rel(full[i:i+1],one))
for prec in (3,1):
    c.set_option("precision",prec)
    full=c.unet_forward(x,500,ctx); one=c.unet_forward(x[2:3],500,ctx[2:3]); print("precision",prec,"img2 batch vs single",rel(full[2:3],one))
c.set_option("precision",0)
lat=synth.make_latent(4,16,16,seed=61)
imgs=c.decode_latent(lat); imgs2=c.decode_latent(lat)
print("decode n=4 repeat identical:", np.array_equal(imgs,imgs2))
for i in range(4):
    one=c.decode_latent(lat[i:i+1]); print("decode img",i,"batch vs single rel",rel(imgs[i:i+1],one))
c.set_option("splitk",0)
imgs=c.decode_latent(lat); one=c.decode_latent(lat[1:2]); print("no splitk: decode img1 batch vs single", rel(imgs[1:2],one))
full=c.unet_forward(x,500,ctx); one=c.unet_forward(x[2:3],500,ctx[2:3]); print("no splitk: unet img2 batch vs single",rel(full[2:3],one))
