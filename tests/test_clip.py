"""CLIP text encoder: oracle vs fixture and vs an independent implementation (CPU), CUDA path vs
fixture through the C ABI (GPU). Tolerance: 1e-3 relative, the same bar as the UNet step (BASELINE.json north_star)."""
import os

import numpy as np
import pytest
import torch

from stable_diffusion_burn_b200 import synth, topology

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "clip.npz")
CASES = ["prompt", "empty", "full77", "pair11"]


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def relmax(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


# ------------------------------------------------------------------------------------------------ CPU
@pytest.fixture(scope="module")
def clipP():
    from oracle import sd_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    raw = synth.make_params(0, which=topology.clip_params())
    return O.Params(raw), raw


def test_oracle_matches_fixture(clipP):
    from oracle import sd_oracle as O
    g = np.load(GOLD)
    with torch.no_grad():
        for case in ("prompt", "pair11"):
            y = O.clip_forward(clipP[0], torch.from_numpy(g["tok:" + case]).long()).numpy()
            assert rel(y, g["out:" + case]) < 1e-5, case


def test_oracle_is_causal_and_batch_independent(clipP):
    """Row l of the output depends on tokens 0..l only (mask of src/backend.rs:130-139), and samples do not mix."""
    from oracle import sd_oracle as O
    g = np.load(GOLD)
    tok = torch.from_numpy(g["tok:full77"]).long()
    with torch.no_grad():
        full = O.clip_forward(clipP[0], tok)
        head = O.clip_forward(clipP[0], tok[:, :20])
        pair = O.clip_forward(clipP[0], torch.cat([tok[:, :11], torch.from_numpy(g["tok:pair11"][1:]).long()]))
    assert rel(head.numpy(), full[:, :20].numpy()) < 1e-5
    assert rel(pair[0].numpy(), full[0, :11].numpy()) < 1e-5
    assert rel(pair[1].numpy(), g["out:pair11"][1]) < 1e-5


def test_oracle_matches_independent_clip_text_model(clipP):
    """Pins the CLIP restatement against a second, independently written implementation of the same published
    architecture (transformers.CLIPTextModel, quick_gelu, causal) loaded with the same synthetic weights."""
    tr = pytest.importorskip("transformers")
    from oracle import sd_oracle as O
    raw = clipP[1]
    cfg = tr.CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12,
                            num_attention_heads=12, max_position_embeddings=77, hidden_act="quick_gelu",
                            layer_norm_eps=1e-5, attn_implementation="eager")
    m = tr.CLIPTextModel(cfg).eval()
    t = lambda name: torch.from_numpy(raw["clip/" + name])
    sdict = {"text_model.embeddings.token_embedding.weight": t("token_embedding/weight"),
             "text_model.embeddings.position_embedding.weight": t("position_embedding/weight"),
             "text_model.final_layer_norm.weight": t("layer_norm/weight"),
             "text_model.final_layer_norm.bias": t("layer_norm/bias")}
    for i in range(12):
        s, d = f"blocks/{i}/", f"text_model.encoder.layers.{i}."
        for a, b in (("attn_ln", "layer_norm1"), ("mlp_ln", "layer_norm2")):
            sdict[d + b + ".weight"], sdict[d + b + ".bias"] = t(s + a + "/weight"), t(s + a + "/bias")
        for a, b in (("attn/query", "self_attn.q_proj"), ("attn/key", "self_attn.k_proj"), ("attn/value", "self_attn.v_proj"),
                     ("attn/out", "self_attn.out_proj"), ("mlp/fc1", "mlp.fc1"), ("mlp/fc2", "mlp.fc2")):
            sdict[d + b + ".weight"] = t(s + a + "/weight").t().contiguous()  # dump-dir Linear is [in,out]
            sdict[d + b + ".bias"] = t(s + a + "/bias")
    missing, unexpected = m.load_state_dict(sdict, strict=False)
    assert not [k for k in missing if "position_ids" not in k] and not unexpected
    g = np.load(GOLD)
    with torch.no_grad():
        for case in ("prompt", "full77"):
            tok = torch.from_numpy(g["tok:" + case]).long()
            want = m(input_ids=tok).last_hidden_state.numpy()
            assert rel(g["out:" + case], want) < 2e-5, case
            assert rel(O.clip_forward(clipP[0], tok).numpy(), want) < 2e-5, case


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def sd(ctx):
    ctx.init_synthetic(0)
    ctx.finalize_weights()
    return ctx


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_clip_forward_golden(sd, case):
    g = np.load(GOLD)
    y = sd.clip_forward(g["tok:" + case])
    want = g["out:" + case]
    assert y.shape == want.shape
    assert rel(y, want) < 1e-3 and relmax(y, want) < 1e-3, (case, rel(y, want), relmax(y, want))


@pytest.mark.gpu
def test_clip_prefix_property_and_errors(sd):
    """Causality at the full 77-token window: a 20-token prefix reproduces the first 20 rows (different row padding,
    different GEMM shapes -> agreement to rounding, not bits)."""
    g = np.load(GOLD)
    tok = g["tok:full77"]
    full, head = sd.clip_forward(tok), sd.clip_forward(tok[:, :20])
    assert rel(head, full[:, :20]) < 1e-3
    with pytest.raises(RuntimeError):
        sd.clip_forward(np.zeros((1, 78), np.int32))
    with pytest.raises(RuntimeError):
        sd.clip_forward(np.full((1, 4), 49408, np.int32))


@pytest.mark.gpu
def test_prompt_to_image_plumbing(sd):
    """tokens -> CLIP -> sample_image, all on the device path: the context the sampler consumes is the CLIP output."""
    g = np.load(GOLD)
    ctx = sd.clip_forward(g["tok:prompt"])
    unc = sd.clip_forward(g["tok:empty"])[0]
    init = synth.make_latent(1, 32, 32, seed=5)
    rgb = sd.sample_image(ctx, unc, 7.5, 1, init_latent=init)
    assert rgb.shape == (1, 256, 256, 3) and rgb.dtype == np.uint8
    want = sd.sample_image(g["out:prompt"], g["out:empty"][0], 7.5, 1, init_latent=init)
    assert (np.abs(rgb.astype(int) - want.astype(int)) <= 1).mean() > 0.999
