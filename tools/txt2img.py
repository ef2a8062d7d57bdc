"""prompt -> PNG through the device path, the way the reference's `sample` binary is used
(src/bin/sample/main.rs:37-48: sample <model_type> <model> <scale> <n_steps> <prompt> <output_name>), as an example of the
Python mirror. Not part of the measured path (bench.py is) and not a CLI clone: weights come from a dump-dir or the synthetic
stream, images are written as <output_name><i>.png like save_images (:115-122).

  python tools/txt2img.py dump params 7.5 20 "An ancient mossy stone." img
  python tools/txt2img.py synthetic 0 7.5 20 "An ancient mossy stone." img      # random-init weights: noise, but end to end
"""
import os
import struct
import sys
import zlib

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402


def write_png(path, rgb):
    """8-bit RGB PNG, no dependency: rgb [H, W, 3] uint8."""
    h, w, _ = rgb.shape
    raw = b"".join(b"\x00" + rgb[y].tobytes() for y in range(h))

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)) +
                chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def main(argv):
    if len(argv) != 7:
        print(f"Usage: {argv[0]} <model_type(dump or synthetic)> <model_dir or seed> <unconditional_guidance_scale> "
              "<n_diffusion_steps> <prompt> <output_image_name>", file=sys.stderr)
        return 1
    kind, model, scale, n_steps, prompt, out = argv[1], argv[2], float(argv[3]), int(argv[4]), argv[5], argv[6]
    from stable_diffusion_burn_b200 import pipeline, tokenizer
    print("Loading tokenizer...")
    tok = tokenizer.SimpleTokenizer(tokenizer.find_vocab())
    print("Loading model...")
    sd = pipeline.StableDiffusion(0)
    sd = sd.load_dump_dir(model) if kind == "dump" else sd.init_synthetic(int(model))
    unconditional_context = sd.unconditional_context(tok)
    context = sd.context(tok, prompt)
    print("Sampling image...")
    images = sd.sample_image(context, unconditional_context, scale, n_steps)
    for i, img in enumerate(images):
        write_png(f"{out}{i}.png", np.asarray(img, np.uint8).reshape(512, 512, 3))
    sd.close()
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
