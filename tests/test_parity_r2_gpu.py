"""GPU parity on the configurations bench.py actually times, through the C ABI.

* the CFG batch at the bench's exact shape (n = 1, 64x64, L = 77, Lu = 2) via sdb_forward_diffuser;
* 20 DDIM steps: every probed step is judged on the ORACLE's latent (UNet-step tolerance 1e-3), the free-running
  result is reported against a stated bound;
* batch 8 at 64x64 (C3 / C5), the 96x96 -> 768x768 decode (C4);
* the VAE decoder / encoder and CLIP against outputs of the REFERENCE'S OWN Python model (tests/golden/ref_python.npz,
  see tests/test_ref_pin_cpu.py) — no oracle in between;
* the GEMM's GEGLU epilogue, in-kernel split-K fold and extra-K operands in isolation (sdb_test_gemm_ex).
"""
import math
import os

import numpy as np
import pytest

from stable_diffusion_burn_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
UNET_TOL = 1.0e-3


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def relmax(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.fixture(scope="module")
def sd(ctx):
    ctx.init_synthetic(0)
    ctx.finalize_weights()
    return ctx


def _unc():
    return synth.make_context(1, 2, seed=99)[0]


# ------------------------------------------------------------------ forward_diffuser, the bench's shape
def test_forward_diffuser_L77(sd):
    g = np.load(os.path.join(GOLD, "cfg_L77.npz"))
    x, c = synth.make_latent(1, 64, 64), synth.make_context(1, 77)
    for t in (999, 449):
        pred, u, cc = sd.forward_diffuser(x, t, c, _unc(), 7.5)
        eu, ec = rel(u, g[f"t{t}:uncond"]), rel(cc, g[f"t{t}:cond"])
        mu, mc = relmax(u, g[f"t{t}:uncond"]), relmax(cc, g[f"t{t}:cond"])
        ep = rel(pred, g[f"t{t}:pred"])
        print(f"forward_diffuser L=77/Lu=2 t={t}: uncond {eu:.3e}/{mu:.3e} cond {ec:.3e}/{mc:.3e} (rel L2 / max), guided pred {ep:.3e}")
        assert max(eu, ec, mu, mc) < UNET_TOL
        # the guidance combine u + 7.5 (c - u) amplifies the two UNet errors by up to |1 - s| + |s| = 14: stated bound 1e-2
        assert ep < 1e-2
        # and the cond half equals a plain UNet::forward on the same inputs (the CFG batch changes nothing per sample)
        assert rel(cc, sd.unet_forward(x, t, c)) < UNET_TOL


# ------------------------------------------------------------------ 20 DDIM steps (config C2 exactly)
def test_sample_20_steps(sd):
    g = np.load(os.path.join(GOLD, "sample_20step.npz"))
    c, init = synth.make_context(1, 77), synth.make_latent(1, 64, 64)
    for i in (0, 9, 19):
        t = int(g[f"step{i}:t"])
        _, u, cc = sd.forward_diffuser(g[f"step{i}:latent_in"], t, c, _unc(), 7.5)
        eu, ec = rel(u, g[f"step{i}:uncond"]), rel(cc, g[f"step{i}:cond"])
        print(f"DDIM step {i} (t={t}) on the oracle's latent: uncond {eu:.3e} cond {ec:.3e}")
        assert eu < UNET_TOL and ec < UNET_TOL
    lat = sd.sample_latent(c, _unc(), 7.5, 20, init_latent=init)
    drift = rel(lat, g["latent"])
    rgb = sd.sample_image(c, _unc(), 7.5, 20, init_latent=init)
    d = np.abs(rgb[:, ::2, ::2, :].astype(np.int16) - g["u8_sub"].astype(np.int16))
    print(f"20 steps free-running: final latent rel L2 {drift:.3e}; u8 within 1 LSB {float((d <= 1).mean()):.4f}, "
          f"within 2 {float((d <= 2).mean()):.4f}, max diff {int(d.max())}")
    # free-running bound: 20 guided steps, each injecting <= 1e-3 per UNet output amplified by the guidance scale and carried
    # through the DDIM recursion (not contractive); stated bound 2e-2 on the latent, 99 % of pixels within 2 LSB
    assert drift < 2e-2
    assert float((d <= 2).mean()) > 0.99


# ------------------------------------------------------------------ batch 8 at 64x64 (C3 / C5)
def test_unet_batch8_64(sd):
    g = np.load(os.path.join(GOLD, "unet_b8_64.npz"))
    out = sd.unet_forward(synth.make_latent(8, 64, 64, seed=808), 599, synth.make_context(8, 77, seed=88))
    errs = [rel(out[i], g["out"][i]) for i in range(8)]
    print("unet batch 8, 64x64, L=77: per-image rel L2", " ".join(f"{e:.2e}" for e in errs))
    assert max(errs) < UNET_TOL and relmax(out, g["out"]) < UNET_TOL


# ------------------------------------------------------------------ 96x96 -> 768x768 decode (C4)
def test_decode_96(sd):
    g = np.load(os.path.join(GOLD, "vae_96.npz"))
    img = sd.decode_latent(synth.make_latent(1, 96, 96, seed=96))
    assert img.shape == (1, 3, 768, 768)
    e2, e3 = rel(img[:, :, ::8, ::8], g["img_sub"]), rel(img[:, :, 380:384, :], g["img_rows"])
    print(f"vae 96x96 -> 768x768: rel L2 sub {e2:.3e} rows {e3:.3e}")
    assert e2 < 1e-3 and e3 < 1e-3
    assert abs(float(img.mean()) - float(g["mean"])) < 1e-3 * max(1.0, abs(float(g["std"])))


# ------------------------------------------------------------------ against the reference's own Python model, directly
def test_vs_reference_python(sd):
    g = np.load(os.path.join(GOLD, "ref_python.npz"))
    e = rel(sd.decode_latent(g["dec16:lat"]), g["dec16:img"])
    img = sd.decode_latent(g["dec64:lat"])
    e2, e3 = rel(img[:, :, ::8, ::8], g["dec64:img_sub"]), rel(img[:, :, 250:254, :], g["dec64:img_rows"])
    ee = rel(sd.encode_image(g["enc64:img"]), g["enc64:lat"])
    ec = max(rel(sd.clip_forward(g["clip:tok"]), g["clip:out"]), rel(sd.clip_forward(g["clip:tok2"]), g["clip:out2"]))
    print(f"vs reference python: decode 16x16 {e:.3e}, decode 64x64 {e2:.3e}/{e3:.3e}, encode {ee:.3e}, clip {ec:.3e}")
    assert max(e, e2, e3, ee, ec) < 1e-3
    # UNet: the Python twin's GEGLU uses tanh-GELU, the Rust model (and this library) erf-GELU; the two forms differ by a few
    # 1e-4 at the UNet output (tests/test_ref_pin_cpu.py), so this check is looser than the UNet-step tolerance
    eu = rel(sd.unet_forward(g["unet64:x"], int(g["unet64:t"]), g["unet64:ctx"]), g["unet64:out"])
    print(f"vs reference python UNet 64x64 L=77 (tanh- vs erf-GELU included): {eu:.3e}")
    assert eu < 2e-3


# ------------------------------------------------------------------ GEMM epilogues / K-loop forms in isolation
def _gelu_erf(x):
    return 0.5 * x * (1.0 + np.vectorize(math.erf)(x / math.sqrt(2.0)))


@pytest.mark.parametrize("M,K,H4,passes", [(256, 320, 1280, 3), (1024, 640, 512, 1), (130, 1280, 256, 3)])
def test_gemm_geglu_epilogue(ctx, M, K, H4, passes):
    rng = np.random.default_rng(M + K)
    a = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((K, 2 * H4)) / math.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(2 * H4).astype(np.float32) * 0.1
    if passes == 1:  # single-pass products see fp16-rounded operands: compare on the same rounding
        a = a.astype(np.float16).astype(np.float32); w = w.astype(np.float16).astype(np.float32)
    p = a.astype(np.float64) @ w.astype(np.float64) + b
    ref = p[:, :H4] * _gelu_erf(p[:, H4:])
    out = ctx.test_gemm_ex(a, w, bias=b, passes=passes, geglu=True)
    e = rel(out, ref)
    print(f"GEGLU epilogue M={M} K={K} H4={H4} passes={passes}: rel L2 {e:.3e}")
    assert e < (5e-5 if passes == 3 else 2e-4)


@pytest.mark.parametrize("M,K,N,passes", [(256, 4096, 320, 3), (128, 11520, 1280, 1), (64, 5120, 640, 3)])
def test_gemm_splitk_fold(ctx, M, K, N, passes):
    """small M x N grid + long K: the library splits K and folds the partial tiles inside the kernel (bias + residual applied
    once, in the fold)."""
    rng = np.random.default_rng(K + N)
    a = rng.standard_normal((M, K)).astype(np.float32)
    w = (rng.standard_normal((K, N)) / math.sqrt(K)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    r = rng.standard_normal((M, N)).astype(np.float32)
    if passes == 1:
        a = a.astype(np.float16).astype(np.float32); w = w.astype(np.float16).astype(np.float32)
    ref = a.astype(np.float64) @ w.astype(np.float64) + b + r
    out = ctx.test_gemm_ex(a, w, bias=b, residual=r, passes=passes)
    out16 = ctx.test_gemm_ex(a, w, bias=b, residual=r, passes=passes, from_f16=True)
    e, e16 = rel(out, ref), rel(out16, ref)
    print(f"split-K fold M={M} K={K} N={N} passes={passes}: rel L2 {e:.3e} (fp16 hi+lo copy {e16:.3e})")
    assert e < 3e-5 and e16 < 3e-5


@pytest.mark.parametrize("M,K,XK,N,passes", [(512, 640, 320, 320, 3), (256, 2560, 1280, 640, 1), (2048, 320, 64, 320, 3)])
def test_gemm_extra_k(ctx, M, K, XK, N, passes):
    """the operands the ResBlock's 1x1 skip conv rides on (appended to the K loop, own weight matrix)."""
    rng = np.random.default_rng(M + XK)
    a = rng.standard_normal((M, K)).astype(np.float32)
    xa = rng.standard_normal((M, XK)).astype(np.float32)
    w = (rng.standard_normal((K, N)) / math.sqrt(K)).astype(np.float32)
    xw = (rng.standard_normal((XK, N)) / math.sqrt(XK)).astype(np.float32)
    b = rng.standard_normal(N).astype(np.float32)
    if passes == 1:
        a, xa, w, xw = (v.astype(np.float16).astype(np.float32) for v in (a, xa, w, xw))
    ref = a.astype(np.float64) @ w.astype(np.float64) + xa.astype(np.float64) @ xw.astype(np.float64) + b
    out = ctx.test_gemm_ex(a, w, bias=b, passes=passes, xa=xa, xw=xw)
    e = rel(out, ref)
    print(f"extra-K M={M} K={K}+{XK} N={N} passes={passes}: rel L2 {e:.3e}")
    assert e < 3e-5


# ------------------------------------------------------------------ LayerNorm folded into the GEMMs around it
@pytest.mark.parametrize("M,K0,C,N,passes,geglu,second", [
    (300, 320, 320, 1152, 3, False, False), (2048, 640, 640, 1920, 3, False, True), (512, 1280, 1280, 1280, 1, False, True),
    (130, 320, 320, 2560, 3, True, True), (1024, 640, 640, 5120, 1, True, False), (64, 1280, 1280, 384, 3, False, False)])
def test_layernorm_folded_into_gemms(ctx, M, K0, C, N, passes, geglu, second):
    """y from a producing GEMM (row statistics in its epilogue, fp16 hi/lo residual pair updated in place), LayerNorm applied by
    the consuming GEMM as rstd * (acc - mean * u) + v: against fp64 LayerNorm + matmul (+ GEGLU)."""
    rng = np.random.default_rng(M + N)
    a = rng.standard_normal((M, K0)).astype(np.float32)
    a2 = rng.standard_normal((M, K0)).astype(np.float32) if second else None
    w0 = (rng.standard_normal((K0, C)) / math.sqrt(K0)).astype(np.float32)
    b0 = (rng.standard_normal(C) * 0.5 + 1.5).astype(np.float32)  # rows with a mean well away from zero: the cancellation case
    g = (1 + 0.1 * rng.standard_normal(C)).astype(np.float32); be = (0.1 * rng.standard_normal(C)).astype(np.float32)
    w1 = (rng.standard_normal((C, N)) / math.sqrt(C)).astype(np.float32)
    b1 = rng.standard_normal(N).astype(np.float32) * 0.3
    y = a.astype(np.float64) @ w0.astype(np.float64) + b0
    if second:
        y = y + a2.astype(np.float64) @ w0.astype(np.float64) + b0
    mu = y.mean(-1, keepdims=True); var = ((y - mu) ** 2).mean(-1, keepdims=True)
    ln = (y - mu) / np.sqrt(var + 1e-5) * g + be
    if passes == 1:  # single-pass consumers see the fp16-rounded raw y and fp16-rounded folded weights: bound, not equality
        tol = 1.5e-3
    else:
        tol = 4e-5
    pre = ln @ w1.astype(np.float64) + b1
    ref = pre[:, :N // 2] * _gelu_erf(pre[:, N // 2:]) if geglu else pre
    out = ctx.test_ln_fold(a, w0, b0, g, be, w1, b1, a2=a2, passes=passes, geglu=geglu)
    e = rel(out, ref)
    print(f"LN folded M={M} C={C} N={N} passes={passes} geglu={geglu} two producers={second}: rel L2 {e:.3e}")
    assert e < tol


# ------------------------------------------------------------------ .mpk model file -> device weights (row f3)
def test_mpk_file_feeds_the_weight_registry(ctx, tmp_path):
    """A NamedMpk-style file holding the VAE decoder (written by mpk.save_mpk: format unverified against burn) replaces the
    decoder weights of a context initialised with a DIFFERENT seed; the decode then matches the seed-0 fixture."""
    from stable_diffusion_burn_b200 import mpk, topology
    params = synth.make_params(0, which=topology.vae_decoder_params())
    f = os.path.join(tmp_path, "decoder.mpk")
    mpk.save_mpk(f, params)
    try:
        ctx.init_synthetic(1)
        n = mpk.load_into(ctx, f)
        assert n == len(params)  # every decoder tensor + the schedule
        ctx.finalize_weights()
        g = np.load(os.path.join(GOLD, "vae_16.npz"))
        assert rel(ctx.decode_latent(synth.make_latent(1, 16, 16, seed=21)), g["img"]) < 1e-3
    finally:
        ctx.init_synthetic(0)
        ctx.finalize_weights()
