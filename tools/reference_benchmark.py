"""The reference's `demo_app -benchmark` run, in the reference's own JSON schema (demo_app/mipmaps_app.cpp:553-858;
published results: demo_app/rtx3090.json), so that the two files can be laid side by side.

  python tools/reference_benchmark.py [--batches 64] [--out profiles/xxx.json] [--images DIR]

Protocol as in the reference: per (image, pipeline alternative) `batches + 1` batches of 8 back-to-back generations
between two timestamps (CUDA events here), the first batch discarded, min / median / max per generation in ns; and
"delta" = the worst code-value difference between the generated chain and the REFERENCE CPU generator
(cpuGenerateMipmaps_sRGBA, the real one from oracle/_ref when it was built, else its restatement) on the same
premultiplied level 0 -- the number the reference records as its own test result (`worstDeltaArray`, :812-821).

Images: with --images DIR (default: tests/golden/test_images when it exists) the reference's 13 test images are used
(needs PIL); otherwise each is replaced by a synthetic image of the same size and alpha class (smooth colour fields plus
noise; alpha images get a varying alpha channel and are premultiplied like mipmaps_app.cpp:606 does).  Alternatives: those of demo_app/pipeline_alternative.cpp that this library offers --
default, generalonly (no fast pipeline), levels_1_5 / levels_1_6 (fast dispatcher <2,5> / <2,6>), f16Shared, srgbShared,
noBilinear (= default here: the software 4-tap first reduction is the only one), blit / generalblit
(NVPYR_FLAG_GENERAL_BLIT).  onelevel / levels_1_3 / levels_3_3 / workgroup1024 are not offered and
are left out.  Timings go through the Python wrapper (10-20 us of host time per call: host-bound for the small images;
tools/bench_native.cpp measures the default alternative from C++).
This tool is bench/test infrastructure: it is the one place outside tests/ and bench.py that calls oracle/.
"""
import argparse, json, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vk_compute_mipmaps_b200 as nv
import _oracle

IMAGES = [  # name, w, h, has alpha   (test_images/, docs/test_images.txt)
    ("1080p.jpg", 1920, 1080, False), ("1440p.jpg", 2560, 1440, False), ("4094.jpg", 4094, 4094, False),
    ("4095.jpg", 4095, 4095, False), ("4096.jpg", 4096, 4096, False), ("4k.jpg", 3840, 2160, False),
    ("alpha1080p.png", 1920, 1080, True), ("alpha2048.png", 2048, 2048, True), ("alpha2052.png", 2052, 2052, True),
    ("lunch_2047.jpg", 2047, 2047, False), ("lunch_with_friend.jpg", 2048, 2048, False),
    ("mandelbrots.png", 3095, 990, False), ("tall.jpg", 1080, 4096, False),
]
ALTERNATIVES = [  # label, flags, fast divisibility, fast max levels
    ("default", nv.FLAG_NONE, 0, 0), ("generalonly", nv.FLAG_FORCE_GENERAL, 0, 0), ("levels_1_5", nv.FLAG_NONE, 2, 5),
    ("levels_1_6", nv.FLAG_NONE, 2, 6), ("srgbShared", nv.FLAG_SRGB_SHARED, 0, 0), ("f16Shared", nv.FLAG_F16_SHARED, 0, 0),
    ("noBilinear", nv.FLAG_NONE, 0, 0), ("blit", nv.FLAG_GENERAL_BLIT | nv.FLAG_FORCE_GENERAL, 0, 0),
    ("generalblit", nv.FLAG_GENERAL_BLIT, 0, 0),
]


def synthetic(w, h, alpha, seed):
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w].astype(np.float32)
    img = np.empty((h, w, 4), dtype=np.float32)
    img[..., 0] = 127.5 + 127.5 * np.sin(x / 37.0 + y / 91.0)
    img[..., 1] = 255.0 * ((x // 64 + y // 64) % 2) * 0.6 + 40.0
    img[..., 2] = 255.0 * y / max(1, h - 1)
    img[..., 3] = 127.5 + 127.5 * np.cos((x - y) / 53.0) if alpha else 255.0
    img[..., :3] += rng.normal(0, 12.0, (h, w, 3)).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8).reshape(-1)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", type=int, default=64)
    ap.add_argument("--out")
    ap.add_argument("--images", help="directory holding the reference's test_images (optional)")
    a = ap.parse_args()
    default_images = os.path.join(ROOT, "tests", "golden", "test_images")
    if not a.images and os.path.isdir(default_images):
        a.images = default_images
    oracle, ref = _oracle.load_oracle(), _oracle.load_ref()
    st = torch.cuda.current_stream()
    lines = ["{"]
    note = ("synthetic stand-ins of the reference's test images (same sizes / alpha classes)" if not a.images else
            "the reference's test images") + "; delta = worst difference vs the reference CPU generator " \
           + ("(oracle/_ref: the real cpuGenerateMipmaps_sRGBA)" if ref else "(restated)") + "; B200, tools/reference_benchmark.py"
    lines.append('"_note": %s,' % json.dumps(note))
    for ii, (name, w, h, alpha) in enumerate(IMAGES):
        if a.images:
            from PIL import Image
            l0 = np.asarray(Image.open(os.path.join(a.images, name)).convert("RGBA"), dtype=np.uint8).reshape(-1)
        else:
            l0 = synthetic(w, h, alpha, ii)
        l0 = oracle.premultiply(l0)  # mipmaps_app.cpp:606 loads every image with doPremultiplyAlpha = true
        chain0 = oracle.new_chain(l0, w, h)
        want = ref.cpu_chain(chain0, w, h) if ref else oracle.cpu_chain(l0, w, h)
        n = chain0.size
        nrot = max(2, min(8, int(400e6 // n) + 1))
        bufs = [torch.from_numpy(chain0).cuda() for _ in range(nrot)]
        lines.append('"%s": {' % name)
        for ai, (label, flags, div, mx) in enumerate(ALTERNATIVES):
            pipes = nv.PyramidPipelines(fast_divisibility=div, fast_max_levels=mx)
            nv.cmd_pyramid_dispatch(st, pipes, w, h, image=bufs[0], flags=flags)
            torch.cuda.synchronize()
            got = bufs[0].cpu().numpy()
            delta = int(oracle.compare(got, want, w, h).worst)
            times = []
            for bi in range(a.batches + 1):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for i in range(8):
                    nv.cmd_pyramid_dispatch(st, pipes, w, h, image=bufs[(bi * 8 + i) % nrot], flags=flags)
                e1.record(st)
                torch.cuda.synchronize()
                if bi:
                    times.append(e0.elapsed_time(e1) * 1e6 / 8)
            times.sort()
            row = '  "%s":%s{"median_ns":%7.0f, "min_ns":%7.0f, "max_ns":%7.0f, "delta":%d}%s' % (
                label, " " * max(0, 18 - len(label)), times[len(times) // 2], times[0], times[-1], delta,
                "}" if ai == len(ALTERNATIVES) - 1 else ",")
            lines.append(row)
            print(name, row, flush=True)
        lines.append("}" if ii == len(IMAGES) - 1 else ",")
        del bufs
        torch.cuda.empty_cache()
    text = "\n".join(lines) + "\n"
    json.loads(text)  # must parse
    if a.out:
        open(a.out, "w").write(text)


if __name__ == "__main__":
    main()
