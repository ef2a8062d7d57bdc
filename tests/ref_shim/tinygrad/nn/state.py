def torch_load(fn):  # python/dump.py:21 imports these for its __main__ (checkpoint conversion); no checkpoint exists here
    raise NotImplementedError("ref_shim: torch_load is not available (no checkpoint in this image)")


def load_state_dict(model, state_dict, strict=True):
    raise NotImplementedError("ref_shim: load_state_dict is not available")
