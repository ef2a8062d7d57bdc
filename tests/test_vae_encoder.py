"""VAE encoder / Autoencoder::encode_image: oracle vs fixture and vs the explicit-padding form (CPU), CUDA
path vs fixture through the C ABI (GPU). Tolerance 1e-3 relative, like the decoder."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from stable_diffusion_burn_b200 import synth, topology

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden", "vae_enc.npz")


def rel(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


@pytest.fixture(scope="module")
def encP():
    from oracle import sd_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    return O.Params(synth.make_params(0, which=topology.vae_encoder_params()))


def test_oracle_matches_fixture(encP):
    from oracle import sd_oracle as O
    g = np.load(GOLD)
    with torch.no_grad():
        y = O.encode_image(encP, torch.from_numpy(g["img:ramp64"])).numpy()
    assert y.shape == (1, 4, 8, 8)
    assert rel(y, g["lat:ramp64"]) < 1e-5


def test_padded_conv_is_bottom_right_padding(encP):
    """The reference's PaddedConv2d(0,1,0,1) (conv with padding 2, output sliced from 1; autoencoder/mod.rs:340-412) equals an
    explicit zero pad of one row/column at the bottom/right followed by an unpadded stride-2 conv."""
    from oracle import sd_oracle as O
    name = "autoencoder/encoder/blocks/0/downsampler"
    x = torch.from_numpy(np.random.default_rng(0).standard_normal((2, 128, 10, 12)).astype(np.float32))
    with torch.no_grad():
        a = O.padded_conv2d_s2(encP, name, x)
        b = F.conv2d(F.pad(x, (0, 1, 0, 1)), encP(name + "/conv/weight"), encP(name + "/conv/bias"), stride=2)
    assert a.shape == (2, 128, 5, 6)
    assert float((a - b).abs().max()) < 1e-5


def test_topology_has_the_sd_v1_encoder():
    ps = {n: s for n, s, _, _ in topology.vae_encoder_params()}
    assert ps["autoencoder/encoder/conv_in/weight"] == (128, 3, 3, 3)
    assert ps["autoencoder/encoder/conv_out/weight"] == (8, 512, 3, 3)          # mid / out stage at 512 channels
    assert ps["autoencoder/encoder/blocks/2/downsampler/conv/weight"] == (512, 512, 3, 3)
    assert "autoencoder/encoder/blocks/3/downsampler/conv/weight" not in ps
    assert ps["autoencoder/quant_conv/weight"] == (8, 8, 1, 1)
    assert abs(sum(int(np.prod(s)) for s in ps.values()) - 34.16e6) < 0.05e6


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def sd(ctx):
    ctx.init_synthetic(0)
    ctx.finalize_weights()
    return ctx


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["ramp64", "randn128x96"])
def test_encode_image_golden(sd, case):
    g = np.load(GOLD)
    y = sd.encode_image(g["img:" + case])
    want = g["lat:" + case]
    assert y.shape == want.shape
    assert rel(y, want) < 1e-3, (case, rel(y, want))
    assert float(np.abs(y - want).max() / np.abs(want).max()) < 1e-3


@pytest.mark.gpu
def test_encode_batch_and_errors(sd):
    g = np.load(GOLD)
    img = g["img:randn128x96"]
    both = sd.encode_image(img)
    one = sd.encode_image(img[1:2])
    assert rel(both[1:2], one) < 1e-3
    big = np.concatenate([img, img, img[:1]])  # 5 images: two chunks
    y5 = sd.encode_image(big)
    assert np.array_equal(y5[4], y5[0]) or rel(y5[4], y5[0]) < 1e-3
    with pytest.raises(RuntimeError):
        sd.encode_image(np.zeros((1, 3, 60, 64), np.float32))
    with pytest.raises(RuntimeError):
        sd.encode_image(np.zeros((1, 3, 32, 32), np.float32))


@pytest.mark.gpu
def test_autoencoder_forward_plumbing(sd):
    """Autoencoder::forward = decode_latent(encode_image(x)) (autoencoder/mod.rs:56-58): shapes and the decoder's input contract."""
    from oracle import sd_oracle as O
    g = np.load(GOLD)
    lat = sd.encode_image(g["img:ramp64"])
    img = sd.decode_latent(lat)
    assert img.shape == (1, 3, 64, 64) and np.isfinite(img).all()
    P = O.Params(synth.make_params(0, which=topology.vae_decoder_params()))
    with torch.no_grad():
        want = O.decode_latent(P, torch.from_numpy(g["lat:ramp64"])).numpy()
    assert rel(img, want) < 2e-3
