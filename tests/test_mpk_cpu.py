"""`.mpk` reader/writer — format unverified against burn (see stable_diffusion_burn_b200/mpk.py): the tests pin
the module's own round trip, the field-name mapping against the registry, and both tensor encodings."""
import os

import numpy as np
import pytest

from stable_diffusion_burn_b200 import mpk, synth, topology


def _small_params():
    """every registry NAME, tiny random payloads of the right rank (the full model is 4.3 GB: shapes are checked on the GPU side)"""
    rng = np.random.default_rng(0)
    out = {}
    for name, shape, _, _ in topology.all_params():
        out[name] = rng.standard_normal([min(int(d), 3) for d in shape]).astype(np.float32)
    out["alpha_cumulative_products"] = synth.alpha_cumulative_products()
    return out


@pytest.mark.parametrize("legacy", [False, True])
def test_round_trip_all_names(tmp_path, legacy):
    params = _small_params()
    f = os.path.join(tmp_path, "m.mpk")
    mpk.save_mpk(f, params, legacy=legacy)
    back = mpk.load_mpk(f)
    assert set(back) == set(params)
    for k in params:
        assert back[k].shape == params[k].shape and np.array_equal(back[k], params[k]), k


def test_record_tree_uses_the_rust_field_names(tmp_path):
    import msgpack
    f = os.path.join(tmp_path, "m.mpk")
    mpk.save_mpk(f, _small_params())
    doc = msgpack.unpackb(open(f, "rb").read(), raw=False)
    item = doc["item"]
    assert set(item) == {"alpha_cumulative_products", "autoencoder", "diffusion", "clip"}  # StableDiffusion's Param / Module fields
    gn = item["diffusion"]["input_blocks"]["rt1"]["res"]["norm_in"]
    assert set(gn) == {"gamma", "beta"}  # src/model/groupnorm/mod.rs:47-48
    assert isinstance(item["autoencoder"]["decoder"]["blocks"], list) and len(item["autoencoder"]["decoder"]["blocks"]) == 4
    assert isinstance(item["clip"]["blocks"], list) and len(item["clip"]["blocks"]) == 12
    assert doc["metadata"]["version"] == "0.14.0"
