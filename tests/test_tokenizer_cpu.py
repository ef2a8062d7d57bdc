"""The reference's only test, reproduced (src/tokenizer.rs:205-221): the one golden vector the reference holds."""
import os

import pytest

from stable_diffusion_burn_b200 import tokenizer as T

try:
    VOCAB = T.find_vocab()
except FileNotFoundError:
    VOCAB = None

pytestmark = pytest.mark.skipif(VOCAB is None, reason="bpe_simple_vocab_16e6.txt (reference data file) not available")


@pytest.fixture(scope="module")
def tok():
    return T.SimpleTokenizer(VOCAB)


def test_reference_kat_encode_decode(tok):
    text = "Hello world! <|startoftext|>asdf<|startoftext|>"
    assert tok.encode(text) == [3306, 1002, 256, 49406, 587, 10468, 49406]
    assert tok.decode(tok.encode(text)) == "hello world ! <|startoftext|>asdf <|startoftext|>"


def test_special_tokens_and_prompt_framing(tok):
    # StableDiffusion::context frames the prompt as <|startoftext|>{text}<|endoftext|> (stablediffusion/mod.rs:200);
    # the unconditional context is the empty prompt -> exactly [49406, 49407]
    assert tok.encode("<|startoftext|><|endoftext|>") == [49406, 49407]
    ids = tok.encode("<|startoftext|>a photo of a cat<|endoftext|>")
    assert ids[0] == 49406 and ids[-1] == 49407 and len(ids) == 7
    assert len(tok.encoder) == 49408


def test_cleaning_quirks(tok):
    assert tok.encode("  Hello   WORLD!\n") == tok.encode("hello world!")
    # no padding / truncation to 77
    assert len(tok.encode("cat " * 100)) == 100
    # round trip of non-ASCII bytes through the byte<->unicode table
    assert tok.decode(tok.encode("café ☕")).strip() == "café ☕"
