"""`.mpk` model files: reader and writer of burn's NamedMpkFileRecorder format — the README's default model
file `SDv1-4.mpk`, written by the reference's `convert` binary (src/bin/convert/main.rs:32-37: `NamedMpkFileRecorder::<
FullPrecisionSettings>`) and read by `sample` (src/bin/sample/main.rs:27-34).

FORMAT UNVERIFIED AGAINST BURN. The recorder lives in the burn crate (0.14.0), which is not in the reference tree and cannot be
fetched here; what follows restates its published layout from memory and is pinned only by this module's own round trip:
  * the file is one MessagePack document written with named fields (maps keyed by Rust field names):
      {"metadata": {"float": "f32", "int": "i32", "format": ..., "version": "0.14.0", "settings": ...}, "item": <record>}
  * a Module's record is a map field -> record; `Vec<Module>` a list; `Option::None` and constants (usize, f64 ...) nil;
  * a `Param<Tensor>` is {"id": <string>, "param": <tensor>} with <tensor> either
      burn >= 0.14  {"bytes": <bin, little-endian>, "shape": [..], "dtype": "F32"}   (TensorData), or
      burn <  0.14  {"value": [floats...], "shape": [..]}                             (DataSerialize) — both are read.
Field names are the Rust struct fields (src/model/*/mod.rs), which the dump-dir directory names copy, except:
  StableDiffusion.diffusion <-> "unet"; GroupNorm / LayerNorm `gamma`, `beta` <-> "weight", "bias"; `alpha_cumulative_products`
  is the root Param. Linear weights are [in, out] and conv weights OIHW in both formats.
"""
from __future__ import annotations

import numpy as np

from . import topology

_DT = {"F32": np.float32, "F64": np.float64, "F16": np.float16, "I32": np.int32, "I64": np.int64}


def _is_param(node) -> bool:
    return isinstance(node, dict) and "param" in node and "id" in node


def _tensor(node) -> np.ndarray:
    if "bytes" in node:
        dt = node.get("dtype", "F32")
        dt = dt if isinstance(dt, str) else next(iter(dt))  # unit variants may arrive as {"F32": nil}
        a = np.frombuffer(bytes(node["bytes"]), dtype=np.dtype(_DT[dt]).newbyteorder("<"))
    else:
        a = np.asarray(node["value"], np.float32)
    return a.astype(np.float32).reshape([int(d) for d in node["shape"]])


def _walk(node, path, out):
    if _is_param(node):
        out["/".join(path)] = _tensor(node["param"])
    elif isinstance(node, dict):
        for k, v in node.items():
            _walk(v, path + [str(k)], out)
    elif isinstance(node, (list, tuple)):
        for i, v in enumerate(node):
            _walk(v, path + [str(i)], out)


def _registry_name(record_path: str) -> str:
    parts = record_path.split("/")
    if parts[0] == "diffusion":
        parts[0] = "unet"
    if parts[-1] == "gamma":
        parts[-1] = "weight"
    elif parts[-1] == "beta":
        parts[-1] = "bias"
    return "/".join(parts)


def load_mpk(path: str) -> dict:
    """-> registry name (dump-dir path) -> float32 array, for every Param in the file."""
    import msgpack
    with open(path, "rb") as f:
        doc = msgpack.unpackb(f.read(), raw=False, strict_map_key=False)
    item = doc["item"] if isinstance(doc, dict) and "item" in doc else doc
    flat = {}
    _walk(item, [], flat)
    return {_registry_name(k): v for k, v in flat.items()}


def save_mpk(path: str, params: dict, legacy: bool = False) -> None:
    """Writes `params` (registry names) as the record tree described above (legacy = the pre-0.14 {"value","shape"} tensors)."""
    import msgpack
    root: dict = {}
    counter = [0]

    def leaf(a):
        a = np.ascontiguousarray(a, np.float32)
        counter[0] += 1
        t = ({"value": a.reshape(-1).tolist(), "shape": list(a.shape)} if legacy
             else {"bytes": a.astype("<f4").tobytes(), "shape": list(a.shape), "dtype": "F32"})
        return {"id": f"param-{counter[0]}", "param": t}

    for name, arr in params.items():
        parts = name.split("/")
        if parts[0] == "unet":
            parts[0] = "diffusion"
        is_norm = any(k in parts[-2] for k in ("norm", "_ln", "layer_norm")) if len(parts) > 1 else False
        if is_norm:
            parts[-1] = {"weight": "gamma", "bias": "beta"}[parts[-1]]
        node = root
        for p in parts[:-1]:
            node = node.setdefault(p, {})
        node[parts[-1]] = leaf(arr)

    def listify(node):  # {"0": .., "1": ..} -> [..] (Vec<Module> fields: blocks)
        if not isinstance(node, dict) or _is_param(node):
            return node
        node = {k: listify(v) for k, v in node.items()}
        if node and all(k.isdigit() for k in node):
            return [node[str(i)] for i in range(len(node))]
        return node

    doc = {"metadata": {"float": "f32", "int": "i32", "format": "burn_core::record::file::NamedMpkFileRecorder<burn_core::record::settings::FullPrecisionSettings>",
                        "version": "0.14.0", "settings": "FullPrecisionSettings"},
           "item": listify(root)}
    with open(path, "wb") as f:
        f.write(msgpack.packb(doc, use_bin_type=True))


def load_into(ctx, path: str) -> int:
    """load_mpk + sdb_set_tensor for every tensor of the registry found in the file; returns the count. Call
    ctx.finalize_weights() afterwards."""
    arrays = load_mpk(path)
    want = {n: s for n, s, _, _ in topology.all_params()}
    want["alpha_cumulative_products"] = (1000,)
    n = 0
    for name, shape in want.items():
        if name in arrays:
            a = arrays[name]
            if tuple(a.shape) != tuple(shape):
                raise ValueError(f"{name}: shape {a.shape} in the file, {shape} in the registry")
            ctx.set_tensor(name, a)
            n += 1
    return n
