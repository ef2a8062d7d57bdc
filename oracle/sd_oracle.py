"""CPU ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Op-for-op torch-CPU fp32 restatement of the reference's sampling hot path
(Gadersd/stable-diffusion-burn @ 893fb095). Every function cites the reference
file:line it follows. The reference's arithmetic lives in third-party crates that are
NOT vendored (burn 0.14.0 -> burn-tch 0.14.0 -> tch 0.15.0 -> libtorch), so the burn op
semantics below are restated from their published definitions:
  nn::Linear      y = x @ W[in,out] + b
  nn::conv::Conv2d = at::conv2d (cross-correlation, OIHW weights, zero padding)
  nn::LayerNorm   (x-mean)/sqrt(var_biased+eps)*gamma+beta, eps 1e-5
  nn::Gelu        exact erf form
  activation::softmax  exp(x-max)/sum
  Tensor::repeat(&[..]) per-dimension tiling

PARITY PINNED against the reference's own Python model. The reference holds no golden vector, known-answer test or fixture
for this path (its only test is the tokenizer KAT, src/tokenizer.rs:205-221) and its Rust cannot be compiled here (no
rustc/cargo, crates not vendored) — but it ships the author's tinygrad twin of the model, python/dump.py (the program that
writes the dump-dir the Rust loaders read). tests/ref_shim/ runs that file UNMODIFIED on a minimal tinygrad stand-in
(tinygrad is not installed), saves the model with the reference's own savers (python/save.py, unet.py, autoencoder.py, clip.py,
stablediffusion.py), reads that tree back with this repo's dump-dir reader and compares forwards (tests/test_ref_pin_cpu.py):
    UNetModel.__call__ (dump.py:326-350)      vs unet_forward      2.2e-6 rel L2 (this oracle's GELU switched to tinygrad's
                                                                     tanh form for that comparison only; erf is the Rust form)
    Decoder / post_quant_conv (:76-108,149-150)   vs decode_latent     1.2e-6
    Encoder / quant_conv (:110-148)           vs encode_image      9.4e-7
    AutoencoderKL.__call__ (:144-150)         vs decode(encode)    2.1e-6
    CLIPTextTransformer.__call__ (:452-461)   vs clip_forward      8.3e-7
    timestep_embedding (:273-277, test.py:31-35) vs timestep_embedding 2.6e-6
and the tree the reference's saver writes equals, file for file (names, metadata, small tensors), the tree this repo's writer
produces. tests/golden/ref_python.npz holds the reference model's outputs on the synthetic weights (script:
tests/ref_shim/make_ref_golden.py); the other tests/golden/*.npz are outputs of THIS oracle (erf GELU), which the GPU suite is
held to. What stays unpinned: the DDIM sampler / CFG arithmetic of src/model/stablediffusion/mod.rs:102-192 (no Python twin
exists; restated from the Rust, section "pipeline" below) and burn's erf-GELU / LayerNorm definitions (third-party crates).

`dtype` may be torch.float32 (the parity target) or torch.float64 (truth for tolerance
budgeting). `emu` optionally rounds GEMM-class operands to a tensor-core input format to
budget the precision mode of the CUDA path (dev tool).
"""
from __future__ import annotations

import math
import numpy as np
import torch
import torch.nn.functional as F

_EMU = {"mode": None, "roles": "awq", "blocks": None, "cur": None}


def set_emulation(mode, roles="awq"):
    """None | 'fp16' | 'bf16' | 'tf32' : round GEMM-class operands (RNE), accumulate wide.
    roles: which operand classes are rounded: 'a' activations feeding conv/linear,
    'w' weights, 'q' attention matmul operands (q,k,p,v)."""
    _EMU["mode"] = mode
    _EMU["roles"] = roles


def set_emulation_blocks(blocks):
    """Restrict emulation to UNet blocks whose tap name is in `blocks` (None = everywhere)."""
    _EMU["blocks"] = None if blocks is None else set(blocks)


def _enter(block):
    _EMU["cur"] = block


def set_emulation_policy(policy):
    """Per-block override: {block tap name: roles string}; blocks not listed use the global roles. Roles: 'a'/'w' Linear
    activations/weights, 'A'/'W' conv activations/weights, 'q' attention matmul operands."""
    _EMU["policy"] = policy


def set_emulation_fn(fn):
    """Finest-grained study hook: fn(block tap name, parameter name, role) -> None (operand exact) | 'fp16' | 'bf16' | 'tf32'
    | 'fp16+e4m3' (fp16 value plus an 8-bit-float correction of the rounding residual). Overrides the role/policy tables."""
    _EMU["fn"] = fn


def _round(x, m):
    if m == "fp16":
        return x.to(torch.float16).to(x.dtype)
    if m == "bf16":
        return x.to(torch.bfloat16).to(x.dtype)
    if m == "tf32":
        xi = x.to(torch.float32).view(torch.int32)
        xi = (xi + 0x0FFF + ((xi >> 13) & 1)) & ~0x1FFF
        return xi.view(torch.float32).to(x.dtype)
    if m == "fp16+e4m3":
        hi = x.to(torch.float16).to(x.dtype)
        lo = (x - hi) * 4096.0  # residual <= 2^-11 |x|: scaled into e4m3's normal range
        return hi + lo.to(torch.float8_e4m3fn).to(x.dtype) / 4096.0
    raise ValueError(m)


def _q(x, role="a", name=None):
    if _EMU.get("fn") is not None:
        m = _EMU["fn"](_EMU["cur"], name, role)
        return x if m is None else _round(x, m)
    m = _EMU["mode"]
    pol = _EMU.get("policy")
    roles = _EMU["roles"]
    if pol is not None and _EMU["cur"] in pol:
        roles = pol[_EMU["cur"]]
    elif _EMU["blocks"] is not None and _EMU["cur"] not in _EMU["blocks"]:
        return x
    if not any(ch.isupper() for ch in roles):
        roles = roles + roles.upper()  # "awq" (no conv-specific letters) covers the conv operands too
    if m is None or role not in roles:
        return x
    if m == "fp16":
        return x.to(torch.float16).to(x.dtype)
    if m == "bf16":
        return x.to(torch.bfloat16).to(x.dtype)
    if m == "tf32":
        xi = x.to(torch.float32).view(torch.int32)
        xi = (xi + 0x0FFF + ((xi >> 13) & 1)) & ~0x1FFF
        return xi.view(torch.float32).to(x.dtype)
    raise ValueError(m)


class Params:
    """name -> tensor, names = reference dump-dir paths (src/model/*/load.rs)."""

    def __init__(self, arrays: dict, dtype=torch.float32, norm_eps=None):
        self.dtype = dtype
        self.norm_eps = dict(norm_eps or {})  # norm dir -> eps read from a dump-dir (load_group_norm / load_layer_norm)
        self.t = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dtype) for k, v in arrays.items()}

    def __call__(self, name):
        return self.t[name]

    def has(self, name):
        return name in self.t


# ------------------------------------------------------------------ primitives
def linear(P, name, x):
    """burn nn::Linear: y = x W + b, W stored [in,out] (reference src/model/load.rs:65-76)."""
    xa = x if id(x) in _EMU.get("pre_rounded", ()) else _q(x, "a", name)  # see nn_layer_norm: LN-fused emulation
    y = xa @ _q(P(f"{name}/weight"), "w", name)
    if P.has(f"{name}/bias"):
        y = y + P(f"{name}/bias")
    return y


def conv2d(P, name, x, stride=1, padding=0):
    """burn nn::conv::Conv2d -> at::conv2d, OIHW (reference src/model/load.rs:118-160)."""
    b = P(f"{name}/bias") if P.has(f"{name}/bias") else None
    return F.conv2d(_q(x, "A", name), _q(P(f"{name}/weight"), "W", name), b, stride=stride, padding=padding)


def silu(x):
    """reference src/model/silu.rs:14-16: x * sigmoid(x)."""
    return x * torch.sigmoid(x)


def layernorm_noaffine(x, eps):
    """reference src/model/groupnorm/mod.rs:75-82: u=x-mean; u/sqrt(mean(u*u)+eps) over last dim."""
    u = x - x.mean(dim=-1, keepdim=True)
    return u / ((u * u).mean(dim=-1, keepdim=True) + eps).sqrt()


def group_norm(P, name, x, n_group=32, eps=1e-5):
    """reference src/model/groupnorm/mod.rs:53-73: reshape [N,G,rest] -> layernorm -> *gamma[C] + beta[C]."""
    shape = x.shape
    n = shape[0]
    c = shape[1]
    y = layernorm_noaffine(x.reshape(n, n_group, -1), P.norm_eps.get(name, eps)).reshape(shape)
    aff = [1] * x.dim()
    aff[1] = c
    return y * P(f"{name}/weight").reshape(aff) + P(f"{name}/bias").reshape(aff)


def nn_layer_norm(P, name, x, eps=1e-5):
    """burn nn::LayerNorm (third-party): biased variance, eps inside sqrt, affine."""
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    if _EMU.get("ln_fused") and _EMU["mode"] is not None:
        # design study (not a product path): LayerNorm folded into the consuming GEMM as a rank-1 correction. The GEMM then reads
        # the RAW x rounded to the operand format and the exact row statistics are applied in its epilogue, which equals
        # normalising the rounded x with the exact statistics; the consumer must not round this tensor again.
        h = (_q(x, "a") - mu) / (var + P.norm_eps.get(name, eps)).sqrt() * P(f"{name}/weight") + P(f"{name}/bias")
        _EMU.setdefault("pre_rounded", set()).add(id(h))
        _EMU.setdefault("keep", []).append(h)  # keep the object alive so its id stays unique
        return h
    return (x - mu) / (var + P.norm_eps.get(name, eps)).sqrt() * P(f"{name}/weight") + P(f"{name}/bias")


_GELU = {"form": "erf"}


def set_gelu_form(form):
    """'erf' (burn nn::Gelu, the parity target) | 'tanh' (tinygrad 0.9.2's Tensor.gelu, python/dump.py:203-210).
    'tanh' exists ONLY so that tests/test_ref_pin_cpu.py can compare this oracle with the reference's Python twin, whose GEGLU
    uses the tanh approximation; every fixture the CUDA path is held to is generated with 'erf'."""
    assert form in ("erf", "tanh")
    _GELU["form"] = form


def gelu_erf(x):
    """burn nn::Gelu = exact erf GELU (used at unet/mod.rs:590)."""
    if _GELU["form"] == "tanh":
        return 0.5 * x * (1.0 + torch.tanh(x * 0.7978845608 * (1.0 + 0.044715 * x * x)))
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def attn_decoder_mask(seq_length, dtype=torch.float32):
    """reference src/backend.rs:130-139: zeros with -inf strictly above the diagonal."""
    return torch.full((seq_length, seq_length), float("-inf"), dtype=dtype).triu(1)


def qkv_attention(q, k, v, n_head, mask=None):
    """reference src/model/attention.rs:5-45 (== src/backend.rs:88-128); mask: additive [n_qctx, n_ctx] or None."""
    n_batch, n_qctx, n_state = q.shape
    n_ctx = k.shape[1]
    scale = (n_state / n_head) ** -0.25
    n_hstate = n_state // n_head
    q = q.reshape(n_batch, n_qctx, n_head, n_hstate).transpose(1, 2) * scale
    k = k.reshape(n_batch, n_ctx, n_head, n_hstate).transpose(1, 2).transpose(-1, -2) * scale
    v = v.reshape(n_batch, n_ctx, n_head, n_hstate).transpose(1, 2)
    qk = _q(q, "q") @ _q(k, "q")
    if mask is not None:
        qk = qk + mask[:n_qctx, :n_ctx]
    # burn softmax: exp(x - max) / sum
    w = qk - qk.amax(dim=3, keepdim=True)
    w = w.exp()
    w = w / w.sum(dim=3, keepdim=True)
    o = (_q(w, "q") @ _q(v, "q")).transpose(1, 2).flatten(2, 3)
    return o


def upsample_nearest2x(x):
    """reference unet/mod.rs:390-398: reshape [n,c,h,1,w,1].repeat([1,1,1,2,1,2]).reshape."""
    n, c, h, w = x.shape
    return x.reshape(n, c, h, 1, w, 1).repeat(1, 1, 1, 2, 1, 2).reshape(n, c, 2 * h, 2 * w)


# ------------------------------------------------------------------------ UNet
def timestep_embedding(t, dim=320, max_period=10000, dtype=torch.float32):
    """reference unet/mod.rs:19-30: freqs=exp(-ln(max_period)/half * arange(half)); [cos|sin] -> [1,dim]."""
    half = dim // 2
    freqs = (torch.arange(0, half, dtype=torch.int64).to(dtype) * (-math.log(max_period) / half)).exp()
    args = torch.tensor([t], dtype=torch.int64).to(dtype) * freqs
    return torch.cat([args.cos(), args.sin()], 0).unsqueeze(0)


def res_block(P, name, x, emb):
    """reference unet/mod.rs:712-734."""
    h = group_norm(P, f"{name}/norm_in", x)
    h = silu(h)
    h = conv2d(P, f"{name}/conv_in", h, padding=1)
    e = linear(P, f"{name}/lin_embed", silu(emb))
    h = h + e.reshape(e.shape[0], e.shape[1], 1, 1)
    h = group_norm(P, f"{name}/norm_out", h)
    h = silu(h)
    h = conv2d(P, f"{name}/conv_out", h, padding=1)
    if P.has(f"{name}/skip_connection/weight"):
        return conv2d(P, f"{name}/skip_connection", x) + h
    return x + h


def multi_head_attention(P, name, x, context, n_head=8):
    """reference unet/mod.rs:641-653."""
    xa = x if context is None else context
    q = linear(P, f"{name}/query", x)
    k = linear(P, f"{name}/key", xa)
    v = linear(P, f"{name}/value", xa)
    wv = qkv_attention(q, k, v, n_head)
    return linear(P, f"{name}/out", wv)


def geglu_mlp(P, name, x):
    """reference unet/mod.rs:551-555, 578-592."""
    p = linear(P, f"{name}/geglu/proj", x)
    half = p.shape[-1] // 2
    a, gate = p[..., :half], p[..., half:]
    return linear(P, f"{name}/lin", a * gelu_erf(gate))


def transformer_block(P, name, x, context):
    """reference unet/mod.rs:521-527."""
    x = x + multi_head_attention(P, f"{name}/attn1", nn_layer_norm(P, f"{name}/norm1", x), None)
    x = x + multi_head_attention(P, f"{name}/attn2", nn_layer_norm(P, f"{name}/norm2", x), context)
    return x + geglu_mlp(P, f"{name}/mlp", nn_layer_norm(P, f"{name}/norm3", x))


def spatial_transformer(P, name, x, context):
    """reference unet/mod.rs:461-481."""
    n, c, h, w = x.shape
    x_in = x
    y = group_norm(P, f"{name}/norm", x)
    y = conv2d(P, f"{name}/proj_in", y)
    y = y.reshape(n, c, h * w).transpose(1, 2)
    y = transformer_block(P, f"{name}/transformer", y, context)
    y = y.transpose(1, 2).reshape(n, c, h, w)
    return x_in + conv2d(P, f"{name}/proj_out", y)


def _unet_block(P, name, kind, x, emb, ctx):
    if kind == "conv":
        return conv2d(P, name, x, padding=1)
    if kind == "down":  # unet/mod.rs:412-427: stride 2, pad 1
        return conv2d(P, name, x, stride=2, padding=1)
    if kind == "r":
        return res_block(P, name, x, emb)
    x = res_block(P, f"{name}/res", x, emb)
    if kind in ("rt", "rtu"):
        x = spatial_transformer(P, f"{name}/transformer", x, ctx)
    if kind in ("ru", "rtu"):
        x = conv2d(P, f"{name}/upsample/conv", upsample_nearest2x(x), padding=1)
    return x


def unet_forward(P, x, t, context, taps=None, prefix="unet"):
    """reference unet/mod.rs:109-142. x [n,4,H,W]; t python int (one timestep for the batch);
    context [n,L,768]. `taps`: optional dict receiving the output of every block."""
    from stable_diffusion_burn_b200 import topology as T
    x = x.to(P.dtype)
    context = context.to(P.dtype)
    _enter("emb")
    t_emb = timestep_embedding(int(t), 320, 10000, P.dtype).to(x.device)
    emb = linear(P, f"{prefix}/lin1_time_embed", t_emb)
    emb = silu(emb)
    emb = linear(P, f"{prefix}/lin2_time_embed", emb)
    if taps is not None:
        taps["emb"] = emb
    saved = []
    for f, kind, _, _ in T.UNET_INPUT_BLOCKS:
        _enter(f"input_blocks/{f}")
        x = _unet_block(P, f"{prefix}/input_blocks/{f}", kind, x, emb, context)
        saved.append(x)
        if taps is not None:
            taps[f"input_blocks/{f}"] = x
    m = f"{prefix}/middle_block"
    _enter("middle_block")
    x = res_block(P, f"{m}/res1", x, emb)
    x = spatial_transformer(P, f"{m}/transformer", x, context)
    x = res_block(P, f"{m}/res2", x, emb)
    if taps is not None:
        taps["middle_block"] = x
    for f, kind, _, _ in T.UNET_OUTPUT_BLOCKS:
        x = torch.cat([x, saved.pop()], 1)
        _enter(f"output_blocks/{f}")
        x = _unet_block(P, f"{prefix}/output_blocks/{f}", kind, x, emb, context)
        if taps is not None:
            taps[f"output_blocks/{f}"] = x
    _enter("out")
    x = group_norm(P, f"{prefix}/norm_out", x)
    x = silu(x)
    return conv2d(P, f"{prefix}/conv_out", x, padding=1)


# ------------------------------------------------------------------ CLIP text encoder
def clip_forward(P, tokens, prefix="clip"):
    """reference src/model/clip/mod.rs:56-75 (+ block :109-115, attention :158-180, MLP/QuickGELU :204-227).
    tokens: int64 [n, L] -> [n, L, 768]."""
    from stable_diffusion_burn_b200 import topology as T
    n, L = tokens.shape
    mask = attn_decoder_mask(L, P.dtype)
    x = P(f"{prefix}/token_embedding/weight")[tokens] + P(f"{prefix}/position_embedding/weight")[:L].unsqueeze(0)
    for i in range(T.CLIP_LAYERS):
        b = f"{prefix}/blocks/{i}"
        h = nn_layer_norm(P, f"{b}/attn_ln", x)
        q, k, v = linear(P, f"{b}/attn/query", h), linear(P, f"{b}/attn/key", h), linear(P, f"{b}/attn/value", h)
        x = x + linear(P, f"{b}/attn/out", qkv_attention(q, k, v, T.CLIP_HEADS, mask))
        h = linear(P, f"{b}/mlp/fc1", nn_layer_norm(P, f"{b}/mlp_ln", x))
        h = h * torch.sigmoid(h * 1.702)  # QuickGELU, clip/mod.rs:224-226
        x = x + linear(P, f"{b}/mlp/fc2", h)
    return nn_layer_norm(P, f"{prefix}/layer_norm", x)


# ------------------------------------------------------------------ VAE decoder
def resnet_block(P, name, x):
    """reference autoencoder/mod.rs:513-528."""
    h = conv2d(P, f"{name}/conv1", silu(group_norm(P, f"{name}/norm1", x)), padding=1)
    h = conv2d(P, f"{name}/conv2", silu(group_norm(P, f"{name}/norm2", h)), padding=1)
    if P.has(f"{name}/nin_shortcut/weight"):
        return conv2d(P, f"{name}/nin_shortcut", x) + h
    return x + h


def conv_self_attention_block(P, name, x):
    """reference autoencoder/mod.rs:562-608 (1 head, d = C)."""
    n, c, hh, ww = x.shape
    h = group_norm(P, f"{name}/norm", x)
    q = conv2d(P, f"{name}/q", h).reshape(n, c, hh * ww).transpose(1, 2)
    k = conv2d(P, f"{name}/k", h).reshape(n, c, hh * ww).transpose(1, 2)
    v = conv2d(P, f"{name}/v", h).reshape(n, c, hh * ww).transpose(1, 2)
    wv = qkv_attention(q, k, v, 1).transpose(1, 2).reshape(n, c, hh, ww)
    return x + conv2d(P, f"{name}/proj_out", wv)


def decode_latent(P, latent, taps=None, prefix="autoencoder"):
    """reference autoencoder/mod.rs:68-71 + Decoder::forward :204-217 + DecoderBlock :307-324 + Mid :456-463."""
    from stable_diffusion_burn_b200 import topology as T
    _enter("vae/in")
    x = conv2d(P, f"{prefix}/post_quant_conv", latent.to(P.dtype))
    d = f"{prefix}/decoder"
    x = conv2d(P, f"{d}/conv_in", x, padding=1)
    _enter("vae/mid1")
    x = resnet_block(P, f"{d}/mid/block_1", x)
    _enter("vae/attn")
    x = conv_self_attention_block(P, f"{d}/mid/attn", x)
    _enter("vae/mid2")
    x = resnet_block(P, f"{d}/mid/block_2", x)
    if taps is not None:
        taps["mid"] = x
    nb = len(T.VAE_DECODER_BLOCKS)
    for i in range(nb):
        b = f"{d}/blocks/{i}"
        _enter(f"vae/b{i}/res1")
        x = resnet_block(P, f"{b}/res1", x)
        _enter(f"vae/b{i}/res2")
        x = resnet_block(P, f"{b}/res2", x)
        _enter(f"vae/b{i}/res3")
        x = resnet_block(P, f"{b}/res3", x)
        if i != nb - 1:
            _enter(f"vae/b{i}/up")
            x = conv2d(P, f"{b}/upsampler", upsample_nearest2x(x), padding=1)
        if taps is not None:
            taps[f"blocks/{i}"] = x
    _enter("vae/out")
    return conv2d(P, f"{d}/conv_out", silu(group_norm(P, f"{d}/norm_out", x)), padding=1)


def padded_conv2d_s2(P, name, x):
    """reference autoencoder/mod.rs:340-412 with PaddingCfg(0,1,0,1), stride 2 (:229-236): a conv with symmetric padding
    padding_actual = 2 whose output is sliced from index 1 -- i.e. pad bottom/right by one, no pad top/left."""
    y = conv2d(P, f"{name}/conv", x, stride=2, padding=2)
    hh, ww = x.shape[2] // 2, x.shape[3] // 2  # desired = (0 + 1 + H - 3) / 2 + 1
    return y[:, :, 1:1 + hh, 1:1 + ww]


def encode_image(P, img, taps=None, prefix="autoencoder"):
    """reference autoencoder/mod.rs:60-66 (encode_image) + Encoder::forward :133-145 + EncoderBlock :255-265 + Mid :456-463.
    img [n,3,H,W] -> latent [n,4,H/8,W/8] (the first 4 of quant_conv's 8 channels)."""
    from stable_diffusion_burn_b200 import topology as T
    e = f"{prefix}/encoder"
    _enter("vae_enc/in")
    x = conv2d(P, f"{e}/conv_in", img.to(P.dtype), padding=1)
    nb = len(T.VAE_ENCODER_BLOCKS)
    for i in range(nb):
        b = f"{e}/blocks/{i}"
        _enter(f"vae_enc/b{i}")
        x = resnet_block(P, f"{b}/res1", x)
        x = resnet_block(P, f"{b}/res2", x)
        if i != nb - 1:
            x = padded_conv2d_s2(P, f"{b}/downsampler", x)
        if taps is not None:
            taps[f"blocks/{i}"] = x
    _enter("vae_enc/mid")
    x = resnet_block(P, f"{e}/mid/block_1", x)
    x = conv_self_attention_block(P, f"{e}/mid/attn", x)
    x = resnet_block(P, f"{e}/mid/block_2", x)
    if taps is not None:
        taps["mid"] = x
    _enter("vae_enc/out")
    x = conv2d(P, f"{e}/conv_out", silu(group_norm(P, f"{e}/norm_out", x)), padding=1)
    x = conv2d(P, f"{prefix}/quant_conv", x)
    return x[:, 0:4]


# -------------------------------------------------------------------- pipeline
def forward_diffuser(P, latent, t, context, uncond, scale, taps=None):
    """reference stablediffusion/mod.rs:162-192. `uncond` [Lu,768] is broadcast over the batch
    (the evident intent of `.unsqueeze().repeat(&[0, n_batch])`).
    `taps`: optional dict receiving the two UNet outputs ("uncond", "cond") before the guidance combine."""
    n = latent.shape[0]
    u_ctx = uncond.unsqueeze(0).repeat(n, 1, 1)
    u = unet_forward(P, latent, t, u_ctx)
    c = unet_forward(P, latent, t, context)
    if taps is not None:
        taps["uncond"], taps["cond"] = u, c
    return u + (c - u) * scale


def ddim_timesteps(n_steps, n_train=1000):
    """reference stablediffusion/mod.rs:111,123: (0..1000).rev().step_by(1000 / n_steps)."""
    step = n_train // n_steps
    return list(range(n_train - 1, -1, -step)), step


def sample_latent(P, context, uncond, scale, n_steps, init_latent, taps=None):
    """reference stablediffusion/mod.rs:102-160 (sigma = 0; the initial N(0,1) latent is an input
    because the reference's RNG is unseeded). Alphas are read as f32, widened to f64 (`.to_f64()`),
    and the scalar coefficients are applied as tensor-by-scalar ops."""
    alphas = P("alpha_cumulative_products").to(torch.float32)
    ts, step = ddim_timesteps(n_steps)
    latent = init_latent.to(P.dtype)
    sigma = 0.0
    for i, t in enumerate(ts):
        a_t = float(alphas[t])
        a_prev = float(alphas[t - step]) if t >= step else 1.0
        sqrt_noise = math.sqrt(1.0 - a_t)
        dtaps = {} if taps is not None else None
        if taps is not None:
            taps[f"step{i}/latent_in"] = latent
        pred = forward_diffuser(P, latent, t, context, uncond, scale, taps=dtaps)
        if taps is not None:
            taps[f"step{i}/uncond"], taps[f"step{i}/cond"] = dtaps["uncond"], dtaps["cond"]
        predx0 = (latent - pred * sqrt_noise) / math.sqrt(a_t)
        dir_latent = pred * math.sqrt(1.0 - a_prev - sigma * sigma)
        latent = predx0 * math.sqrt(a_prev) + dir_latent  # + gen_noise()*sigma, sigma == 0
        if taps is not None:
            taps[f"step{i}/pred"] = pred
            taps[f"step{i}/latent"] = latent
    return latent


def latent_to_image_f32(P, latent):
    """reference stablediffusion/mod.rs:69-84 up to (not including) the u8 cast: [n,H,W,3] floats."""
    img = decode_latent(P, latent * (1.0 / 0.18215))
    img = (img + 1.0) / 2.0
    return img.permute(0, 2, 3, 1) * 255.0


def to_u8(img_f):
    """reference stablediffusion/mod.rs:94-97: v.to_f64().min(255.0).max(0.0) as u8 (truncation; NaN -> 255->... `as u8` of NaN is 0 after min/max keeps 255: f64::min(NaN,255)=255)."""
    a = img_f.detach().to(torch.float64).numpy()
    a = np.where(np.isnan(a), 255.0, a)
    a = np.maximum(np.minimum(a, 255.0), 0.0)
    return a.astype(np.uint8)  # truncation toward zero


def sample_image(P, context, uncond, scale, n_steps, init_latent):
    """reference stablediffusion/mod.rs:51-67 -> list of n HWC u8 arrays flattened like Vec<Vec<u8>>."""
    lat = sample_latent(P, context, uncond, scale, n_steps, init_latent)
    return to_u8(latent_to_image_f32(P, lat))
