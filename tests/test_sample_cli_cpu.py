"""Argument contract of the `sample` stand-in (tools/sample.py) against the reference binary's (src/bin/sample/main.rs:36-57):
argument count, parse errors and their messages, exit code 1 — everything that needs no GPU."""
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("sample_cli", os.path.join(ROOT, "tools", "sample.py"))
sample = importlib.util.module_from_spec(spec)
spec.loader.exec_module(sample)


def test_usage_and_parse_errors(capsys):
    with pytest.raises(SystemExit) as e:
        sample.parse_args(["sample", "dump", "params"])
    assert e.value.code == 1 and "Usage: sample <model_type(burn or dump)> <model_name>" in capsys.readouterr().err
    with pytest.raises(SystemExit) as e:
        sample.parse_args(["sample", "dump", "params", "x", "20", "a prompt", "img"])
    assert e.value.code == 1 and "Error: Invalid unconditional guidance scale." in capsys.readouterr().err
    with pytest.raises(SystemExit) as e:
        sample.parse_args(["sample", "dump", "params", "7.5", "-3", "a prompt", "img"])
    assert e.value.code == 1 and "Error: Invalid number of diffusion steps." in capsys.readouterr().err
    with pytest.raises(SystemExit) as e:
        sample.parse_args(["sample", "dump", "params", "7.5", "20", "a prompt", "img", "tpu"])
    assert e.value.code == 1 and "Unknown device: tpu" in capsys.readouterr().err
    with pytest.raises(SystemExit) as e:  # accepted by the reference, refused here: no CPU fallback
        sample.parse_args(["sample", "dump", "params", "7.5", "20", "a prompt", "img", "cpu"])
    assert e.value.code == 1


def test_accepted_forms():
    assert sample.parse_args(["sample", "burn", "SDv1-4.mpk", "7.5", "20", "An ancient mossy stone.", "img"]) == \
        ("burn", "SDv1-4.mpk", 7.5, 20, "An ancient mossy stone.", "img", 0)
    assert sample.parse_args(["sample", "dump", "params", "7.5", "1", "p", "o", "cuda"])[-1] == 0
    assert sample.parse_args(["sample", "dump", "params", "7.5", "1", "p", "o", "CUDA3"])[-1] == 3
    assert sample.parse_args(["sample", "dump", "params", "7.5", "1", "p", "o", "cudax"])[-1] == 0  # unwrap_or(0), main.rs:67
