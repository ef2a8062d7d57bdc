"""Dump-dir weight format: the native reader against files written in the reference's layout."""
import ctypes as C
import os

import numpy as np
import pytest

from stable_diffusion_burn_b200 import _lib, dumpdir, synth, topology


def _read(lib, file, ndim):
    dims = (C.c_int64 * 4)()
    n = lib.sdb_read_dump_tensor(os.fsencode(file), ndim, dims, None, 0)
    if n < 0:
        raise RuntimeError(lib.sdb_last_error(None).decode())
    buf = np.empty(n, np.float32)
    assert lib.sdb_read_dump_tensor(os.fsencode(file), ndim, dims, buf.ctypes.data_as(C.POINTER(C.c_float)), n) == n
    return buf.reshape([dims[i] for i in range(ndim)])


def test_native_tensor_reader_matches_writer(tmp_path):
    lib = _lib.load()
    g = np.random.default_rng(0)
    for shape in [(7,), (3, 5), (4, 3, 3, 3), (2, 1, 1, 1)]:
        a = g.standard_normal(shape).astype(np.float32)
        dumpdir.save_tensor(a, "weight", str(tmp_path))
        f = str(tmp_path / "weight.npy")
        assert np.array_equal(_read(lib, f, len(shape)), a)
        assert np.array_equal(dumpdir.read_tensor(f), a)
    dumpdir.save_scalar(1e-5, "eps", str(tmp_path))  # save_scalar = a [1]-shaped tensor (python/save.py:6-8)
    assert _read(lib, str(tmp_path / "eps.npy"), 1)[0] == np.float32(1e-5)


def test_native_reader_rejects_malformed_files(tmp_path):
    lib = _lib.load()
    bad = {
        "f64": np.array([2.0, 1.0, 2.0], np.float64),                 # NpyData<f32> only (load.rs:39)
        "twod": np.zeros((2, 3), np.float32),                          # tensors are 1-D [dims..., values...]
        "count": np.array([2.0, 3.0, 1.0, 2.0], np.float32),           # shape says 6 values, file has 2
        "negdim": np.array([-1.0, 1.0], np.float32),
    }
    for name, arr in bad.items():
        np.save(tmp_path / f"{name}.npy", arr)
        with pytest.raises(RuntimeError):
            _read(lib, str(tmp_path / f"{name}.npy"), 2 if name == "count" else 1)
    (tmp_path / "trunc.npy").write_bytes(open(tmp_path / "count.npy", "rb").read()[:-5])
    with pytest.raises(RuntimeError):
        _read(lib, str(tmp_path / "trunc.npy"), 2)
    (tmp_path / "junk.npy").write_bytes(b"not numpy at all")
    with pytest.raises(RuntimeError):
        _read(lib, str(tmp_path / "junk.npy"), 1)
    with pytest.raises(RuntimeError, match="missing file"):
        _read(lib, str(tmp_path / "absent.npy"), 1)


def test_python_round_trip_and_optional_tensors(tmp_path):
    """Writer -> numpy reader on the CLIP subtree; a dropped GroupNorm affine / bias comes back as ones / zeros."""
    which = topology.clip_params()[:12] + topology.vae_decoder_params()[:4]
    P = synth.make_params(3, which=which)
    drop = {"autoencoder/post_quant_conv/bias"}
    dumpdir.save_dump_dir(str(tmp_path), {k: v for k, v in P.items() if k not in drop})
    assert dumpdir.read_tensor(str(tmp_path / "n_steps.npy"))[0] == 1000
    assert dumpdir.read_tensor(str(tmp_path / "clip/blocks/0/attn/n_head.npy"))[0] == 12
    assert np.array_equal(dumpdir.read_tensor(str(tmp_path / "autoencoder/post_quant_conv/stride.npy")), [1, 1])
    for name, _, _, _ in which:
        f = tmp_path / (name + ".npy")
        if name in drop:
            assert not f.exists()
        else:
            assert np.array_equal(dumpdir.read_tensor(str(f)), P[name]), name


# ------------------------------------------------------------------------------------------------ GPU
@pytest.fixture(scope="module")
def dump(tmp_path_factory):
    """A full SD-v1.4 dump-dir (4.3 GB of .npy) of the synthetic weights, seed 5, with one non-default eps and two
    optional tensors left out."""
    root = str(tmp_path_factory.mktemp("dump"))
    P = synth.make_params(5)
    eps = {"autoencoder/decoder/norm_out": 1e-3, "unet/input_blocks/rt1/transformer/transformer/norm1": 1e-4}
    drop = {"unet/input_blocks/rt1/res/norm_in/weight", "unet/input_blocks/rt1/res/conv_in/bias"}
    dumpdir.save_dump_dir(root, {k: v for k, v in P.items() if k not in drop}, eps=eps)
    for k in drop:
        P[k] = np.ones_like(P[k]) if k.endswith("weight") else np.zeros_like(P[k])
    return root, P, eps


@pytest.mark.gpu
def test_load_dump_dir_matches_set_tensor(ctx, dump):
    root, P, eps = dump
    ctx.load_dump_dir(root)
    names = [t[0] for t in ctx.tensor_list()]
    for name in names[:: max(1, len(names) // 60)] + ["alpha_cumulative_products", "unet/input_blocks/rt1/res/norm_in/weight",
                                                      "unet/input_blocks/rt1/res/conv_in/bias", "clip/token_embedding/weight"]:
        assert np.array_equal(ctx.get_tensor(name, P[name].shape), P[name]), name
    ctx.finalize_weights()
    x, c = synth.make_latent(1, 32, 32, seed=2), synth.make_context(1, 5, seed=4)
    got = ctx.unet_forward(x, 500, c)
    img = ctx.decode_latent(synth.make_latent(1, 8, 8, seed=3))
    # the same weights through sdb_set_tensor, default eps -> differs only through the two overridden eps values
    ctx.init_synthetic(5)
    for k in ("unet/input_blocks/rt1/res/norm_in/weight", "unet/input_blocks/rt1/res/conv_in/bias"):
        ctx.set_tensor(k, P[k])
    ctx.finalize_weights()
    base = ctx.unet_forward(x, 500, c)
    base_img = ctx.decode_latent(synth.make_latent(1, 8, 8, seed=3))
    d_unet = float(np.abs(got - base).max() / np.abs(base).max())
    d_img = float(np.abs(img - base_img).max() / np.abs(base_img).max())
    assert 0 < d_unet < 5e-2 and 1e-4 < d_img < 0.5, (d_unet, d_img)
    # and against the oracle reading the same directory with the same eps table
    import torch
    from oracle import sd_oracle as O
    Po = O.Params(dumpdir.load_dump_dir(root), norm_eps=eps)
    with torch.no_grad():
        want = O.unet_forward(Po, torch.from_numpy(x), 500, torch.from_numpy(c)).numpy()
        want_img = O.decode_latent(Po, torch.from_numpy(synth.make_latent(1, 8, 8, seed=3))).numpy()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) < 1e-3
    assert np.linalg.norm(img - want_img) / np.linalg.norm(want_img) < 1e-3


@pytest.mark.gpu
def test_load_dump_dir_validates_configuration(ctx, dump, tmp_path):
    root, _, _ = dump
    cases = [
        ("unet/input_blocks/d1/stride.npy", np.array([2, 1, 1], np.float32), "compiled"),   # stride 1 on a downsampler
        ("clip/n_layer.npy", np.array([1, 11], np.float32), "compiled"),
        ("unet/input_blocks/rt1/res/norm_in/n_group.npy", np.array([1, 16], np.float32), "compiled"),
        ("unet/lin1_time_embed/weight.npy", np.concatenate([[320, 1279], np.zeros(320 * 1279)]).astype(np.float32), "shape"),
        ("unet/input_blocks/rt1/transformer/transformer/attn1/query/bias.npy", np.array([320] + [0] * 320, np.float32), "no such tensor"),
    ]
    for rel, arr, msg in cases:
        f = os.path.join(root, rel)
        keep = open(f, "rb").read() if os.path.exists(f) else None
        np.save(f, arr)
        try:
            with pytest.raises(RuntimeError, match=msg):
                ctx.load_dump_dir(root)
        finally:
            if keep is None:
                os.remove(f)
            else:
                open(f, "wb").write(keep)
    os.rename(os.path.join(root, "clip/layer_norm/bias.npy"), os.path.join(root, "clip/layer_norm/bias.bak"))
    try:
        with pytest.raises(RuntimeError, match="missing file"):  # LayerNorm affine is not optional (load.rs:93-94)
            ctx.load_dump_dir(root)
    finally:
        os.rename(os.path.join(root, "clip/layer_norm/bias.bak"), os.path.join(root, "clip/layer_norm/bias.npy"))
    with pytest.raises(RuntimeError, match="missing file"):
        ctx.load_dump_dir(str(tmp_path))
    ctx.load_dump_dir(root)  # intact again
