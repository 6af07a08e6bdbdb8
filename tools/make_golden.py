"""Generates tests/golden/*.npz from the reference's own test images and the
reference's own CPU generator compiled in place (oracle/_ref/libnvpyr_ref.so).

Run in the build container only (/root/reference does not exist on the GPU box):
    python tools/make_golden.py

Each fixture: a crop of a reference test image (RGBA8, premultiplied through the
reference-equivalent pre-pass when the benchmark would do so,
demo_app/mipmaps_app.cpp:606), chosen to reproduce the schedule class the full
image exercises (docs/test_images.txt), plus
  ref_cpu_sha256   sha256 of the packed chain produced by the REFERENCE's
                   cpuGenerateMipmaps_sRGBA (include/mipmap_storage.hpp:395-414)
  oracle_a_sha256  sha256 of the chain produced by our shader-order Oracle A
  delta_a_vs_ref   worst |Oracle A - reference CPU| (MipmapStorage::compare),
                   to be held against demo_app/rtx3090.json's recorded deltas
"""
import hashlib
import os
import sys

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _oracle  # noqa: E402

IMAGES = "/root/reference/test_images"
OUT = os.path.join(ROOT, "tests", "golden")

# name, file, (x, y, w, h), premultiply, what it exercises
CROPS = [
    ("pow2_256", "4096.jpg", (1024, 1024, 256, 256), False, "fast 6 + fast 2 (4096.jpg class)"),
    ("odd_255", "4095.jpg", (100, 200, 255, 255), False, "all 3x3 general (4095.jpg class)"),
    ("odd_127", "lunch_2047.jpg", (900, 700, 127, 127), False, "all 3x3 general (lunch_2047 class)"),
    ("m3_120x72", "1080p.jpg", (640, 360, 120, 72), False, "fast 3 then general (1080p class)"),
    ("m5_160x96", "1440p.jpg", (800, 480, 160, 96), False, "fast 5 then general (1440p class)"),
    ("m4_240x144", "4k.jpg", (1600, 880, 240, 144), False, "fast 4 then general (4k class)"),
    ("alpha_m2_260", "alpha2052.png", (896, 896, 260, 260), True, "fast 2, general, fast (alpha2052 class)"),
    ("alpha_m3_200x120", "alpha1080p.png", (800, 400, 200, 120), True, "alpha, fast 3 then general"),
    ("tall_136x512", "tall.jpg", (400, 1500, 136, 512), False, "fast, general, general, fast, general (tall class)"),
    ("npot_309x99", "mandelbrots.png", (1200, 400, 309, 99), True, "arbitrary NPOT (mandelbrots class)"),
    ("even_254", "4094.jpg", (2000, 2000, 254, 254), False, "even but not %4: general (4094 class)"),
]


def main():
    o = _oracle.load_oracle()
    r = _oracle.load_ref()
    assert r is not None, "needs the reference tree"
    os.makedirs(OUT, exist_ok=True)
    for name, fn, (x, y, w, h), premul, what in CROPS:
        im = Image.open(os.path.join(IMAGES, fn)).convert("RGBA").crop((x, y, x + w, y + h))
        l0 = np.asarray(im, dtype=np.uint8).reshape(-1).copy()
        if premul:
            l0 = o.premultiply(l0)
        a, stores = o.shader_chain(l0, w, h)
        ref_chain = r.cpu_chain(o.new_chain(l0, w, h), w, h)
        b = o.cpu_chain(l0, w, h)
        assert (b == ref_chain).all(), "Oracle B must equal the reference CPU generator bit for bit"
        delta, where = r.compare(a, ref_chain, w, h)
        plan = [(s.pipeline, s.input_level, s.level_count) for s in o.plan(w, h)]
        np.savez_compressed(
            os.path.join(OUT, name + ".npz"), level0=l0.reshape(h, w, 4), width=w, height=h,
            source=f"{fn} crop x={x} y={y} premultiplied={premul}", what=what,
            ref_cpu_sha256=hashlib.sha256(ref_chain.tobytes()).hexdigest(),
            oracle_a_sha256=hashlib.sha256(a.tobytes()).hexdigest(), delta_a_vs_ref=delta,
            plan=np.array(plan, dtype=np.uint32))
        print(f"{name:18s} {w}x{h} plan={plan} delta(A,ref)={delta} at {where}")


if __name__ == "__main__":
    main()
