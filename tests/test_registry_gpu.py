"""Weight registry: names/shapes equal the topology mirror of the reference's dump-dir layout, and the device
synthetic generator is bit-identical to the numpy one."""
import numpy as np
import pytest

from stable_diffusion_burn_b200 import synth, topology

pytestmark = pytest.mark.gpu


def test_registry_matches_topology(ctx):
    got = ctx.tensor_list()
    want = [(n, tuple(s)) for (n, s, _, _) in topology.all_params()] + [("alpha_cumulative_products", (1000,))]
    assert got == want


def test_synthetic_bit_identical(ctx):
    ctx.init_synthetic(0)
    plist = topology.all_params()
    picks = [plist[0], plist[1], plist[7], plist[40], plist[-1], plist[-2], plist[len(plist) // 2]]
    for n, s, k, f in picks:
        dev = ctx.get_tensor(n, s)
        host = synth.make_tensor(n, s, k, f, 0)
        assert np.array_equal(dev, host), n
    a = ctx.get_tensor("alpha_cumulative_products", (1000,))
    assert np.array_equal(a, synth.alpha_cumulative_products())


def test_set_get_roundtrip(ctx):
    name, shape, _, _ = topology.all_params()[4]  # a 4-D conv weight
    v = np.random.default_rng(0).standard_normal(shape).astype(np.float32)
    ctx.set_tensor(name, v)
    assert np.array_equal(ctx.get_tensor(name, shape), v)
    with pytest.raises(Exception):
        ctx.set_tensor(name, v.reshape(-1))
    with pytest.raises(Exception):
        ctx.set_tensor("no/such/tensor", v)
