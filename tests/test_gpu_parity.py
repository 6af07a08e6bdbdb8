"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle.

Bar (BASELINE.json north_star): bit-exact 8-bit codes against the shader-order oracle
(Oracle A) for every pipeline; <= recorded worst delta against the reference's own CPU
generator (2 opaque / 5 alpha, demo_app/rtx3090.json); rgba32f within 1e-6 relative of a
float64 evaluation (and bit-exact against the float32 shader-order oracle)."""
import glob
import hashlib
import os

import numpy as np
import pytest

import _oracle

pytestmark = pytest.mark.gpu
GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
MAGENTA = np.array([255, 0, 255, 254], dtype=np.uint8)  # mipmaps_app.cpp:185-201 (alpha 254: never a valid opaque result)


def gpu_chain(nv, torch, l0, w, h, fmt=0, levels=0, pipelines=None, flags=0, prefill=True):
    pipelines = pipelines or nv.PyramidPipelines(format=fmt)
    dt = torch.uint8 if fmt == 0 else torch.float32
    n = nv.chain_bytes(w, h, levels, fmt) // (1 if fmt == 0 else 4)
    buf = torch.empty(n, dtype=dt, device="cuda")
    if fmt == 0 and prefill:
        buf.view(-1, 4)[:] = torch.from_numpy(MAGENTA).cuda()
    elif prefill:
        buf.fill_(float("nan"))
    buf[:4 * w * h] = torch.from_numpy(np.ascontiguousarray(l0)).cuda().view(-1)
    nv.cmd_pyramid_dispatch(None, pipelines, w, h, levels, image=buf, flags=flags)
    torch.cuda.synchronize()
    return buf.cpu().numpy()


def assert_same(a, b, w, h, oracle, what=""):
    if not (a == b).all():
        c = oracle.compare(a, b, w, h)
        raise AssertionError(f"{what} {w}x{h}: worst delta {c.worst} at x={c.x} y={c.y} level={c.level} "
                             f"channel={c.channel}; {c.mismatched}/{c.compared} texels differ")


FAST_SIZES = [(64, 64), (128, 64), (64, 128), (256, 256), (4, 4), (8, 8), (16, 48), (32, 32), (96, 160), (192, 320),
              (1024, 1024), (16, 16), (128, 4096), (2048, 64), (8, 32), (12, 20), (448, 64)]
GENERAL_SIZES = [(63, 63), (100, 37), (1, 50), (50, 1), (5, 5), (2, 2), (33, 2), (255, 255), (511, 300), (3, 3),
                 (1, 2), (2, 1), (1, 1), (254, 254), (17, 513), (129, 129), (6, 10)]
MIXED_SIZES = [(260, 260), (136, 512), (120, 72), (160, 96), (240, 144), (1920, 1080), (1080, 4096), (2052, 2052),
               (2560, 1440), (3095, 990)]


@pytest.mark.parametrize("size", FAST_SIZES + GENERAL_SIZES + MIXED_SIZES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_srgba8_bit_exact_vs_shader_order_oracle(nv, cuda, oracle, size):
    w, h = size
    for seed, make in ((1, _oracle.random_level0), (2, lambda w, h, s: _oracle.smooth_level0(w, h, s))):
        l0 = make(w, h, seed)
        want, _ = oracle.shader_chain(l0, w, h)
        got = gpu_chain(nv, cuda, l0, w, h)
        assert_same(got, want, w, h, oracle, "srgba8")


@pytest.mark.parametrize("size", [(256, 256), (4096, 64), (192, 320), (136, 512), (260, 260), (63, 63)],
                         ids=lambda s: f"{s[0]}x{s[1]}")
def test_dark_and_sparse_images_bit_exact(nv, cuda, oracle, size):
    """Exact zeros, the smallest non-zero sums (isolated code-1 texels) and saturated texels: the
    edge cases of the clamp-free encode table of the tuned kernel."""
    w, h = size
    rng = np.random.default_rng(w * 7 + h)
    images = [np.zeros(4 * w * h, dtype=np.uint8), np.full(4 * w * h, 255, dtype=np.uint8)]
    sparse = np.zeros(4 * w * h, dtype=np.uint8)
    idx = rng.integers(0, sparse.size, sparse.size // 97)
    sparse[idx] = 1
    images.append(sparse)
    low = rng.integers(0, 3, 4 * w * h, dtype=np.uint8)
    images.append(low)
    spikes = np.zeros(4 * w * h, dtype=np.uint8)
    spikes[rng.integers(0, spikes.size, 40)] = 255
    images.append(spikes)
    for l0 in images:
        want, _ = oracle.shader_chain(l0, w, h)
        got = gpu_chain(nv, cuda, l0, w, h)
        assert_same(got, want, w, h, oracle, "dark/sparse")


def test_generic_functor_kernel_matches_tuned_kernel(nv, cuda, oracle):
    """NVPYR_GENERIC_FAST=1 routes sRGBA8 through fastKernel<Srgba8, M> (the functor-template kernel a user
    instance would get); it must give the same bits as the tuned kernel.  Runs in a subprocess because the
    switch is read once at library load."""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import vk_compute_mipmaps_b200 as nv, _oracle\n"
        "o = _oracle.load_oracle()\n"
        "for (w, h) in [(256, 256), (192, 64), (120, 72), (64, 4096)]:\n"
        "    l0 = _oracle.random_level0(w, h, 3)\n"
        "    buf = torch.zeros(nv.chain_bytes(w, h), dtype=torch.uint8, device='cuda')\n"
        "    buf[:4 * w * h] = torch.from_numpy(l0).cuda()\n"
        "    nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(), w, h, image=buf)\n"
        "    torch.cuda.synchronize()\n"
        "    assert (buf.cpu().numpy() == o.shader_chain(l0, w, h)[0]).all(), (w, h)\n"
        "print('generic ok')\n") % (_oracle.ROOT, os.path.join(_oracle.ROOT, "tests"))
    env = dict(os.environ, NVPYR_GENERIC_FAST="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "generic ok" in out.stdout, out.stderr[-2000:]


def test_standalone_kernels_on_small_levels(nv, cuda, oracle):
    """NVPYR_TAIL_MAX_TEXELS=0 disables the fused tail launch, so the tuned stand-alone kernels (strip-walking
    general kernel, warp-tile fast kernel) also run the tiny and ragged levels: every strip/segment edge case at
    sizes the oracle checks in milliseconds."""
    import subprocess
    import sys
    sizes = [(63, 63), (100, 37), (255, 255), (511, 300), (254, 254), (17, 513), (129, 129), (6, 10), (31, 31),
             (32, 32), (61, 61), (62, 62), (33, 2), (2, 33), (3, 3), (5, 5), (2, 2), (301, 7), (7, 301), (91, 181),
             (260, 260), (136, 512), (120, 72), (1023, 57), (64, 64), (192, 320)]
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import vk_compute_mipmaps_b200 as nv, _oracle\n"
        "o = _oracle.load_oracle()\n"
        "for (w, h) in %r:\n"
        "  for fg in (False, True):\n"
        "    l0 = _oracle.random_level0(w, h, w * 31 + h)\n"
        "    buf = torch.zeros(nv.chain_bytes(w, h), dtype=torch.uint8, device='cuda')\n"
        "    buf[:4 * w * h] = torch.from_numpy(l0).cuda()\n"
        "    nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(fast_pipeline=not fg), w, h, image=buf)\n"
        "    torch.cuda.synchronize()\n"
        "    want = o.shader_chain(l0, w, h, force_general=fg)[0]\n"
        "    got = buf.cpu().numpy()\n"
        "    assert (got == want).all(), (w, h, fg, int((got != want).sum()))\n"
        "print('standalone ok')\n") % (_oracle.ROOT, os.path.join(_oracle.ROOT, "tests"), sizes)
    env = dict(os.environ, NVPYR_TAIL_MAX_TEXELS="0")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "standalone ok" in out.stdout, (out.stdout[-500:], out.stderr[-2000:])


@pytest.mark.parametrize("knobs", [{}, {"NVPYR_CASCADE_AREA_MAX": "40000"}, {"NVPYR_CASCADE_COST_FACTOR": "100", "NVPYR_CASCADE_COST_FLOOR": "100000000"},
                                   {"NVPYR_CASCADE_SOLO_MAX_TEXELS": "17000", "NVPYR_CASCADE_ROUND_COST": "0"}],
                         ids=["default-knobs", "small-area-raw-staging", "deepest-groups", "large-solo-many-rounds"])
def test_cascade_tail_bit_exact(nv, cuda, oracle, knobs):
    """NVPYR_CASCADE=1: the general dispatches of a chain's small levels run as cascades (cascadeRun: several
    dispatches per launch on shared-memory tiles with recomputed halos, then the rest of the chain solo on one CTA;
    off by default because it is no faster).  Same bits as the oracle for sRGBA8 (with and without
    the fast pipeline, both lossy shared types) and rgba32f, with the geometry knobs pushed to their corners: tiny
    areas (raw staging, many tiles per CTA), the deepest groups, whole 127^2 levels solo."""
    import subprocess
    import sys
    sizes = [(1023, 1023), (1022, 766), (511, 300), (513, 513), (255, 255), (254, 254), (127, 129), (100, 37), (63, 63),
             (1920, 1080), (1080, 512), (2052, 1028), (773, 247), (17, 513), (301, 7), (7, 301), (1, 37), (64, 1), (5, 5),
             (3, 3), (2, 2), (260, 260), (640, 360), (2047, 700)]
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import vk_compute_mipmaps_b200 as nv, _oracle\n"
        "o = _oracle.load_oracle()\n"
        "n = 0\n"
        "for (w, h) in %r:\n"
        "  if w * h > %d: continue\n"
        "  l0 = _oracle.random_level0(w, h, w * 13 + h)\n"
        "  for fg in (False, True):\n"
        "    for flags in (0, nv.FLAG_F16_SHARED, nv.FLAG_SRGB_SHARED):\n"
        "      if flags and (fg or w * h > 300000): continue\n"
        "      buf = torch.zeros(nv.chain_bytes(w, h), dtype=torch.uint8, device='cuda')\n"
        "      buf[:4 * w * h] = torch.from_numpy(l0).cuda()\n"
        "      nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(fast_pipeline=not fg), w, h, image=buf, flags=flags)\n"
        "      torch.cuda.synchronize()\n"
        "      want = o.shader_chain(l0, w, h, force_general=fg, f16_shared=bool(flags & nv.FLAG_F16_SHARED),\n"
        "                            srgb_shared=bool(flags & nv.FLAG_SRGB_SHARED))[0]\n"
        "      got = buf.cpu().numpy()\n"
        "      assert (got == want).all(), (w, h, fg, flags, int((got != want).sum()))\n"
        "      n += 1\n"
        "  if w * h <= 300000:\n"
        "    f0 = _oracle.random_level0(w, h, w + h, fmt=1)\n"
        "    fb = torch.zeros(nv.chain_bytes(w, h, 0, 1) // 4, dtype=torch.float32, device='cuda')\n"
        "    fb[:4 * w * h] = torch.from_numpy(np.ascontiguousarray(f0)).cuda().view(-1)\n"
        "    nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(format=1), w, h, image=fb)\n"
        "    torch.cuda.synchronize()\n"
        "    wantf = o.shader_chain(f0, w, h, fmt=1)[0]\n"
        "    assert (fb.cpu().numpy().view(np.uint32) == wantf.view(np.uint32)).all(), ('rgba32f', w, h)\n"
        "    n += 1\n"
        "print('cascade ok', n)\n") % (_oracle.ROOT, os.path.join(_oracle.ROOT, "tests"), sizes, 600000 if knobs else 1 << 30)
    env = dict(os.environ, NVPYR_CASCADE="1", **knobs)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "cascade ok" in out.stdout, (out.stdout[-500:], out.stderr[-2000:])


@pytest.mark.parametrize("size", [(64, 64), (256, 256), (96, 160), (100, 37), (260, 260), (1, 9)],
                         ids=lambda s: f"{s[0]}x{s[1]}")
def test_force_general_bit_exact(nv, cuda, oracle, size):
    """fastPipeline == VK_NULL_HANDLE / -force-no-fast-pipeline."""
    w, h = size
    l0 = _oracle.random_level0(w, h, 3)
    want, _ = oracle.shader_chain(l0, w, h, force_general=True)
    got = gpu_chain(nv, cuda, l0, w, h, pipelines=nv.PyramidPipelines(fast_pipeline=False))
    assert_same(got, want, w, h, oracle, "force-general")
    got2 = gpu_chain(nv, cuda, l0, w, h, flags=nv.FLAG_FORCE_GENERAL)
    assert (got == got2).all()


@pytest.mark.parametrize("div,max_levels", [(2, 6), (2, 5), (2, 3), (8, 3), (4, 4)])
def test_dispatcher_variants_bit_exact(nv, cuda, oracle, div, max_levels):
    """levels_1_6 / levels_1_5 / levels_1_3 / levels_3_3 style <Div, Max> variants change the carry groups."""
    for w, h in [(256, 256), (66, 130), (192, 64), (120, 72)]:
        l0 = _oracle.random_level0(w, h, 5)
        want, _ = oracle.shader_chain(l0, w, h, div=div, max_levels=max_levels)
        got = gpu_chain(nv, cuda, l0, w, h,
                        pipelines=nv.PyramidPipelines(fast_divisibility=div, fast_max_levels=max_levels))
        assert_same(got, want, w, h, oracle, f"variant<{div},{max_levels}>")


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_fixtures(nv, cuda, oracle, path):
    """Crops of the reference's test images: GPU == Oracle A bits; GPU vs the REFERENCE's CPU chain within the
    reference's recorded worst deltas (rtx3090.json: 2 opaque, 5/4/4 alpha)."""
    g = np.load(path)
    w, h = int(g["width"]), int(g["height"])
    l0 = g["level0"].reshape(-1)
    got = gpu_chain(nv, cuda, l0, w, h)
    assert hashlib.sha256(got.tobytes()).hexdigest() == str(g["oracle_a_sha256"])
    ref_cpu = oracle.cpu_chain(l0, w, h)
    assert hashlib.sha256(ref_cpu.tobytes()).hexdigest() == str(g["ref_cpu_sha256"])
    c = oracle.compare(got, ref_cpu, w, h)
    opaque = bool((g["level0"][..., 3] == 255).all())
    assert c.worst == int(g["delta_a_vs_ref"]) and c.worst <= (2 if opaque else 5)


@pytest.mark.parametrize("size", [(64, 64), (256, 128), (96, 160), (63, 63), (100, 37), (260, 260), (1, 7), (2, 2)],
                         ids=lambda s: f"{s[0]}x{s[1]}")
def test_rgba32f(nv, cuda, oracle, size):
    w, h = size
    l0 = _oracle.random_level0(w, h, 8, fmt=1)
    want, _ = oracle.shader_chain(l0, w, h, fmt=1)
    got = gpu_chain(nv, cuda, l0, w, h, fmt=1)
    assert (got.view(np.uint32) == want.view(np.uint32)).all(), "float32 bits differ from shader-order oracle"
    # tolerance stated by north_star: 1e-6 relative against a float64 evaluation of the same weights
    cur = l0.reshape(h, w, 4).astype(np.float64)
    off, cw, ch = w * h, w, h
    for s in oracle.plan(w, h):
        for _ in range(s.level_count):
            cur = downsample_f64(cur)
            ch, cw = cur.shape[:2]
            lvl = got[4 * off:4 * (off + cw * ch)].reshape(ch, cw, 4)
            np.testing.assert_allclose(lvl, cur, rtol=1e-6, atol=1e-12)
            off += cw * ch
        cur = got[4 * (off - cw * ch):4 * off].reshape(ch, cw, 4).astype(np.float64)  # next dispatch re-reads


def downsample_f64(src):
    """Energy-conserving 1/2/3-tap separable reduction in float64 (weights of nvpro_pyramid.glsl:582-586)."""
    def axis(a, ax):
        n_src = a.shape[ax]
        a = np.moveaxis(a, ax, 0)
        if n_src == 1:
            out = a
        elif n_src % 2 == 0:
            out = 0.5 * (a[0::2] + a[1::2])
        else:
            n = n_src // 2
            i = np.arange(n, dtype=np.float64).reshape((-1,) + (1,) * (a.ndim - 1))
            w0, w1, w2 = (n - i) / (2 * n + 1), n / (2 * n + 1), (1 + i) / (2 * n + 1)
            out = w0 * a[0:2 * n:2] + w1 * a[1:2 * n:2] + w2 * a[2:2 * n + 1:2]
        return np.moveaxis(out, 0, ax)
    return axis(axis(src, 0), 1)


def test_premultiply(nv, cuda, oracle):
    w, h = 200, 120
    l0 = _oracle.smooth_level0(w, h, 4)
    want_l0 = oracle.premultiply(l0)
    src = cuda.from_numpy(l0).cuda()
    dst = cuda.empty_like(src)
    nv.premultiply_alpha(None, src, dst, w * h)
    cuda.cuda.synchronize()
    assert (dst.cpu().numpy() == want_l0).all()
    # flag path == pre-pass followed by generation (scoped_image.hpp:233-255 then the dispatch)
    want, _ = oracle.shader_chain(want_l0, w, h)
    got = gpu_chain(nv, cuda, l0, w, h, flags=nv.FLAG_PREMULTIPLY_ALPHA)
    assert_same(got, want, w, h, oracle, "premultiply")
    # opaque texels are unchanged
    op = _oracle.random_level0(64, 64, 1, opaque=True)
    s2 = cuda.from_numpy(op).cuda()
    nv.premultiply_alpha(None, s2, s2, 64 * 64)
    assert (s2.cpu().numpy() == op).all()
    # texel counts that are not multiples of four, buffers that are only 4-byte aligned, in place and not
    big = _oracle.random_level0(257, 129, 8)
    want_big = oracle.premultiply(big)
    for first, count in ((0, 257 * 129), (0, 1002), (1, 1003), (3, 4099), (2, 3)):
        src = cuda.from_numpy(big).cuda()
        dst = cuda.zeros_like(src)
        nv.premultiply_alpha(None, src[4 * first:], dst[4 * first:], count)
        nv.premultiply_alpha(None, src[4 * first:], src[4 * first:], count)
        cuda.cuda.synchronize()
        sl = slice(4 * first, 4 * (first + count))
        assert (dst.cpu().numpy()[sl] == want_big[sl]).all() and (src.cpu().numpy()[sl] == want_big[sl]).all(), (first, count)
        assert (dst.cpu().numpy()[4 * (first + count):] == 0).all() and (dst.cpu().numpy()[:4 * first] == 0).all()


@pytest.mark.parametrize("size", [(1024, 1024), (1920, 1080), (1028, 1028), (768, 2048)])
def test_premultiply_fused_into_level0_read(nv, cuda, oracle, size):
    """Large images whose chain starts with a tuned fast step premultiply level 0 inside that launch
    (scoped_image.hpp:233-255 fused into the read): same level 0 and same chain as premultiplying first.
    Inputs: random alpha; mostly opaque with transparent and zero-colour patches (unchanged blocks are not
    rewritten)."""
    w, h = size
    rnd = _oracle.random_level0(w, h, 31)
    patch = _oracle.random_level0(w, h, 32, opaque=True).reshape(h, w, 4).copy()
    patch[h // 4:h // 2, w // 4:w // 2, 3] = 0          # fully transparent block
    patch[h // 2:h // 2 + 9, :, 3] = 128                 # a semi-transparent stripe
    patch[:7, :11, :3] = 0                               # black, opaque
    patch[-5:, -5:] = [0, 0, 0, 77]                      # black, translucent
    for l0 in (rnd, patch.reshape(-1)):
        pm = oracle.premultiply(l0)
        want, _ = oracle.shader_chain(pm, w, h)
        before = nv.launch_count()
        got = gpu_chain(nv, cuda, l0, w, h, flags=nv.FLAG_PREMULTIPLY_ALPHA)
        launches = nv.launch_count() - before
        assert (got[:4 * w * h] == pm).all(), "level 0 is not the premultiplied image"
        assert_same(got, want, w, h, oracle, "fused premultiply")
        # no separate premultiply launch: as many launches as without the flag
        before = nv.launch_count()
        gpu_chain(nv, cuda, pm, w, h)
        assert nv.launch_count() - before == launches


@pytest.mark.parametrize("size", [(256, 256), (1024, 512), (96, 160), (32, 32), (260, 260), (255, 255), (511, 300),
                                  (136, 512), (1920, 1080), (100, 37), (64, 4096)])
def test_f16_shared_variant_bit_exact(nv, cuda, oracle, size):
    """NVPYR_FLAG_F16_SHARED = the reference's F16_SHARED build of the sRGBA8 shaders (values passing through
    shared memory inside a dispatch are rounded to binary16): bit-exact against Oracle A's restatement, which
    is pinned by executing the reference's shaders with the macro set (tests/test_oracle_pins.py).  Also with
    the fast pipeline disabled, through the host round trip, and rejected for rgba32f."""
    w, h = size
    l0 = _oracle.random_level0(w, h, 21)
    for fg in (False, True):
        want, _ = oracle.shader_chain(l0, w, h, force_general=fg, f16_shared=True)
        got = gpu_chain(nv, cuda, l0, w, h, pipelines=nv.PyramidPipelines(fast_pipeline=not fg), flags=nv.FLAG_F16_SHARED)
        assert_same(got, want, w, h, oracle, f"f16 shared, force_general={fg}")
    want, _ = oracle.shader_chain(l0, w, h, f16_shared=True)
    assert (nv.generate_host(l0, w, h, flags=nv.FLAG_F16_SHARED) == want).all()
    if size == (256, 256):
        assert (want != oracle.shader_chain(l0, w, h)[0]).any()
        with pytest.raises(nv.NvpyrError):
            gpu_chain(nv, cuda, _oracle.random_level0(64, 64, 1, fmt=1), 64, 64, fmt=1, flags=nv.FLAG_F16_SHARED)


@pytest.mark.parametrize("size", [(256, 256), (1024, 512), (96, 160), (32, 32), (260, 260), (255, 255), (511, 300),
                                  (136, 512), (1920, 1080), (100, 37), (64, 4096)])
def test_srgb_shared_variant_bit_exact(nv, cuda, oracle, size):
    """NVPYR_FLAG_SRGB_SHARED = the reference's SRGB_SHARED build of the sRGBA8 shaders (values passing through
    shared memory inside a dispatch are packed to 8-bit sRGB and unpacked again): bit-exact against Oracle A's
    restatement, which is pinned by executing the reference's shaders with the macro set
    (tests/test_oracle_pins.py).  Also with the fast pipeline disabled, through the host round trip, on a smooth
    image, and rejected for rgba32f and together with F16_SHARED."""
    w, h = size
    for l0 in (_oracle.random_level0(w, h, 23), _oracle.smooth_level0(w, h, 5)):
        for fg in (False, True):
            want, _ = oracle.shader_chain(l0, w, h, force_general=fg, srgb_shared=True)
            got = gpu_chain(nv, cuda, l0, w, h, pipelines=nv.PyramidPipelines(fast_pipeline=not fg), flags=nv.FLAG_SRGB_SHARED)
            assert_same(got, want, w, h, oracle, f"srgb shared, force_general={fg}")
    want, _ = oracle.shader_chain(l0, w, h, srgb_shared=True)
    assert (nv.generate_host(l0, w, h, flags=nv.FLAG_SRGB_SHARED) == want).all()
    if size == (256, 256):
        assert (want != oracle.shader_chain(l0, w, h)[0]).any()
        with pytest.raises(nv.NvpyrError):
            gpu_chain(nv, cuda, _oracle.random_level0(64, 64, 1, fmt=1), 64, 64, fmt=1, flags=nv.FLAG_SRGB_SHARED)
        with pytest.raises(nv.NvpyrError):
            gpu_chain(nv, cuda, l0, w, h, flags=nv.FLAG_SRGB_SHARED | nv.FLAG_F16_SHARED)


@pytest.mark.parametrize("size", [(256, 256), (1000, 700), (333, 97), (2052, 1028), (5, 5), (1, 37), (1920, 1080), (260, 260)])
def test_general_blit_variant_bit_exact(nv, cuda, oracle, size):
    """NVPYR_FLAG_GENERAL_BLIT = demo_app's "generalblit" alternative (pipeline_alternative.cpp:16,
    mipmap_pipelines.cpp:350-453): levels the fast pipeline does not take are blitted one at a time with a linear
    filter; with the fast pipeline absent it is the "blit" alternative.  Bit-exact against the oracle's restatement of
    the same loop and of the pinned blit arithmetic (a Vulkan blit's precision is implementation-defined: that part
    of the parity is unpinned), for sRGBA8 and rgba32f, through the host round trip, and rejected
    together with the shared-type flags."""
    w, h = size
    for fmt in (0, 1):
        l0 = _oracle.random_level0(w, h, 31, fmt=fmt)
        for fg in (False, True):
            want, _ = oracle.shader_chain(l0, w, h, fmt=fmt, force_general=fg, general_blit=True)
            got = gpu_chain(nv, cuda, l0, w, h, fmt=fmt, pipelines=nv.PyramidPipelines(format=fmt, fast_pipeline=not fg),
                            flags=nv.FLAG_GENERAL_BLIT)
            if fmt == 0:
                assert_same(got, want, w, h, oracle, f"general blit, force_general={fg}")
            else:
                assert (got.view(np.uint32) == want.view(np.uint32)).all(), (size, fg)
    l0 = _oracle.random_level0(w, h, 31)
    want, _ = oracle.shader_chain(l0, w, h, general_blit=True)
    assert (nv.generate_host(l0, w, h, flags=nv.FLAG_GENERAL_BLIT) == want).all()
    if size == (333, 97):
        assert (want != oracle.shader_chain(l0, w, h)[0]).any()  # not the general pipeline's result
        with pytest.raises(nv.NvpyrError):
            gpu_chain(nv, cuda, l0, w, h, flags=nv.FLAG_GENERAL_BLIT | nv.FLAG_F16_SHARED)


def test_external_memory_fd_import_live(nv, cuda):
    """nvpyrImportExternalMemoryFd with a LIVE file descriptor (the CUDA half of the Vulkan interop):
    device memory created through CUDA's virtual-memory API is exported as a POSIX fd -- the GPU boxes have no Vulkan
    implementation to call vkGetMemoryFdKHR on (profiles/r2_vulkan_probe.txt) -- the fd goes through the library's import
    (cudaImportExternalMemory, opaque fd: the call a Vulkan fd takes), the chain is generated in the imported buffer
    and must equal the oracle's bit for bit; mixed schedule (1920x1080), fast + general + fast (260x260), NPOT."""
    import subprocess, sys
    tool = os.path.join(_oracle.ROOT, "tools", "extmem_probe.py")
    for w, h in ((1920, 1080), (260, 260), (333, 97)):
        r = subprocess.run([sys.executable, tool, str(w), str(h)], capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        assert "imported: device pointer" in r.stdout and "== oracle" in r.stdout, r.stdout


def test_slab_task_handoff_stress(nv, cuda):
    """The slab-task mode hands a tile's level +3 sums from the warps that produce them to the warp that arrives last
    through shared memory ordered by fences and a shared atomic (no barrier), and recycles stash slots through a
    generation counter -- orderings compute-sanitizer's racecheck cannot model.  Evidence instead: thousands of chains
    on varied sizes, in a process where the mode is forced onto every size (up to 8192^2: 110 tiles and four slot
    generations per CTA), every repetition equal to the first, and the first equal to what a tile-mode process
    (no hand-off at all) produces."""
    import subprocess, sys
    tool = os.path.join(_oracle.ROOT, "tools", "stress_slab.py")
    outs = {}
    for name, env in (("slab", {"NVPYR_SLAB_MAX_TILES_PER_WARP_X100": "1000000"}), ("tile", {"NVPYR_NO_SLAB_TASKS": "1"})):
        r = subprocess.run([sys.executable, tool, "--reps", "400" if name == "slab" else "8"], capture_output=True, text=True,
                           timeout=900, env=dict(os.environ, **env))
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
        outs[name] = {l.split()[0]: l.split()[-1] for l in r.stdout.strip().splitlines()}
        assert all(" differing 0 " in l for l in r.stdout.strip().splitlines()), r.stdout
    assert len(outs["slab"]) == 7 and outs["slab"] == outs["tile"], outs


def test_partial_level_count(nv, cuda, oracle):
    w, h = 256, 256
    l0 = _oracle.random_level0(w, h, 6)
    for levels in (2, 3, 7, 8):
        want, _ = oracle.shader_chain(l0, w, h, levels=levels)
        got = gpu_chain(nv, cuda, l0, w, h, levels=levels)
        assert (got == want).all(), levels


def test_pitched_levels(nv, cuda, oracle):
    """Per-level pointer + pitch table (analogue of the reference's per-level image views)."""
    w, h = 192, 128
    l0 = _oracle.random_level0(w, h, 7)
    want, _ = oracle.shader_chain(l0, w, h)
    n = nv.level_count(w, h)
    bufs, ptrs, pitches = [], [], []
    for i in range(n):
        lw, lh = max(1, w >> i), max(1, h >> i)
        pitch = (lw * 4 + 64 + 15) // 16 * 16 if i % 2 == 0 else lw * 4 + 4
        b = cuda.zeros(pitch * lh, dtype=cuda.uint8, device="cuda")
        bufs.append(b), ptrs.append(b.data_ptr()), pitches.append(pitch)
    lvl0 = cuda.from_numpy(l0).cuda().view(h, w * 4)
    bufs[0].view(h, pitches[0])[:, :w * 4] = lvl0
    nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(), w, h, image=None, level_ptrs=ptrs, pitches=pitches)
    cuda.cuda.synchronize()
    off = 0
    for i in range(n):
        lw, lh = max(1, w >> i), max(1, h >> i)
        got = bufs[i].view(lh, pitches[i])[:, :lw * 4].cpu().numpy().reshape(-1)
        assert (got == want[4 * off:4 * (off + lw * lh)]).all(), i
        off += lw * lh


def test_batch_and_determinism(nv, cuda, oracle):
    w, h = 128, 128
    imgs, wants = [], []
    for k in range(5):
        l0 = _oracle.random_level0(w, h, 100 + k)
        wants.append(oracle.shader_chain(l0, w, h)[0])
        buf = cuda.zeros(nv.chain_bytes(w, h), dtype=cuda.uint8, device="cuda")
        buf[:4 * w * h] = cuda.from_numpy(l0).cuda()
        imgs.append(buf)
    nv.dispatch_batch(None, nv.PyramidPipelines(), imgs, w, h)
    cuda.cuda.synchronize()
    for b, want in zip(imgs, wants):
        assert (b.cpu().numpy() == want).all()
    first = [b.clone() for b in imgs]
    nv.dispatch_batch(None, nv.PyramidPipelines(), imgs, w, h)  # idempotent: level 0 untouched
    cuda.cuda.synchronize()
    assert all(bool((a == b).all()) for a, b in zip(first, imgs))


@pytest.mark.parametrize("size,flags", [((1024, 1024), 0), ((2048, 1024), 0), ((1920, 1080), 0), ((1020, 1020), 0),
                                        ((1024, 768), 2)])
def test_fused_batch_bit_exact(nv, cuda, oracle, size, flags):
    """nvpyrDispatchBatch on images of one size whose plan is 'one big fast step, then small steps': TWO
    launches for the whole batch (all tiles of all images through one persistent grid, then one CTA per
    image for the rest), same bits as one dispatch per image.  Sizes: fast6 + fast4; fast6 + fast4 + general;
    fast3 + general x4; fast2 + general ...; premultiplied."""
    w, h = size
    count = 5
    imgs, wants = [], []
    for k in range(count):
        l0 = _oracle.random_level0(w, h, 300 + k)
        wants.append(oracle.shader_chain(oracle.premultiply(l0) if flags & 2 else l0, w, h)[0])
        buf = cuda.empty(nv.chain_bytes(w, h), dtype=cuda.uint8, device="cuda")
        buf.view(-1, 4)[:] = cuda.from_numpy(MAGENTA).cuda()
        buf[:4 * w * h] = cuda.from_numpy(l0).cuda()
        imgs.append(buf)
    before = nv.launch_count()
    nv.dispatch_batch(None, nv.PyramidPipelines(), imgs, w, h, flags=flags)
    cuda.cuda.synchronize()
    launches = nv.launch_count() - before
    for k, (b, want) in enumerate(zip(imgs, wants)):
        assert_same(b.cpu().numpy(), want, w, h, oracle, f"image {k}")
    assert launches == 2, launches  # the premultiply pre-pass rides in the first launch


def test_heterogeneous_batch_falls_back(nv, cuda, oracle):
    """Different streams or sizes cannot share a launch: one dispatch per image, same results."""
    w, h = 1024, 1024
    imgs, wants = [], []
    for k in range(3):
        l0 = _oracle.random_level0(w, h, 400 + k)
        wants.append(oracle.shader_chain(l0, w, h)[0])
        buf = cuda.zeros(nv.chain_bytes(w, h), dtype=cuda.uint8, device="cuda")
        buf[:4 * w * h] = cuda.from_numpy(l0).cuda()
        imgs.append(buf)
    cuda.cuda.synchronize()
    descs = (nv.pyramid.DispatchDesc * 3)()
    streams = [cuda.cuda.Stream() for _ in range(3)]
    for i, img in enumerate(imgs):
        descs[i] = nv.pyramid._make_desc(img, nv.PyramidPipelines(), w, h, 0, 0, streams[i])
    before = nv.launch_count()
    nv.pyramid.check(nv.pyramid.lib.nvpyrDispatchBatch(descs, 3), "nvpyrDispatchBatch")
    cuda.cuda.synchronize()
    assert nv.launch_count() - before == 3 * 2
    for b, want in zip(imgs, wants):
        assert (b.cpu().numpy() == want).all()


def test_generate_host_round_trip(nv, cuda, oracle):
    """minimal_app shape: host level 0 in, packed host chain out."""
    for (w, h) in [(256, 256), (255, 131)]:
        l0 = _oracle.random_level0(w, h, 12)
        want, _ = oracle.shader_chain(l0, w, h)
        got = nv.generate_host(l0, w, h)
        assert (got == want).all()
    l0f = _oracle.random_level0(64, 48, 2, fmt=1)
    wantf, _ = oracle.shader_chain(l0f, 64, 48, fmt=1)
    gotf = nv.generate_host(l0f, 64, 48, fmt=nv.FORMAT_RGBA32F)
    assert (gotf.view(np.uint32) == wantf.view(np.uint32)).all()


def test_generate_host_banded_pipeline(nv, cuda, oracle):
    """nvpyrGenerateHost cuts level 0 into bands and overlaps upload / kernels / download on three streams
    when the chain starts with a fast step.  NVPYR_HOST_BAND_BYTES (read at library load, hence the
    subprocess) shrinks the bands so that sizes the oracle handles exercise many bands, ragged last bands,
    the in-place (staging buffer) mode, the premultiply pre-pass per band and rgba32f."""
    import subprocess
    import sys
    code = r"""
import sys, numpy as np
sys.path.insert(0, %r); sys.path.insert(0, %r)
import _oracle, vk_compute_mipmaps_b200 as nv
o = _oracle.load_oracle()
# (w, h): 6-level first step with 9 bands + ragged last; M = 2 (260 = 4 * 65); M = 3; wide
for (w, h) in [(512, 1088), (260, 780), (1000, 1016), (2048, 256)]:
    l0 = _oracle.random_level0(w, h, w + h)
    want, _ = o.shader_chain(l0, w, h)
    got = nv.generate_host(l0, w, h)
    assert (got == want).all(), ("separate", w, h)
    chain = np.zeros(nv.chain_bytes(w, h), np.uint8)
    chain[:4 * w * h] = l0.ravel()
    nv.generate_host(chain[:4 * w * h], w, h, out=chain)
    assert (chain == want).all(), ("in place", w, h)
    got = nv.generate_host(l0, w, h, mip_levels=3)
    assert (got == o.shader_chain(l0, w, h, levels=3)[0]).all(), ("3 levels", w, h)
# premultiply: level 0 changes, so it must come back even in place
w, h = 256, 1024
l0 = _oracle.random_level0(w, h, 5)
pm = o.premultiply(l0)
want, _ = o.shader_chain(pm, w, h)
chain = np.zeros(nv.chain_bytes(w, h), np.uint8)
chain[:4 * w * h] = l0.ravel()
nv.generate_host(chain[:4 * w * h], w, h, flags=nv.FLAG_PREMULTIPLY_ALPHA, out=chain)
assert (chain == want).all(), "premultiply in place"
# rgba32f
w, h = 128, 512
l0f = _oracle.random_level0(w, h, 2, fmt=1)
wantf, _ = o.shader_chain(l0f, w, h, fmt=1)
gotf = nv.generate_host(l0f, w, h, fmt=nv.FORMAT_RGBA32F)
assert (gotf.view(np.uint32) == wantf.view(np.uint32)).all(), "rgba32f"
print("banded ok", nv.launch_count())
""" % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, NVPYR_HOST_BAND_BYTES=str(128 * 1024))
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout + out.stderr
    # 512x1088 in 128 KB bands = 64-row bands -> 17 launches for step 0 alone: banding really happened
    assert int(out.stdout.split()[-1]) > 60, out.stdout


def test_minimal_app_equivalent(nv, cuda, oracle, tmp_path):
    """examples/minimal_mipmaps = the reference's minimal_app with the Vulkan sequence swapped for libnvpyr:
    same command line, same per-level TGA files.  Checked against the oracle chain of the same input,
    with -do-premultiply-alpha and with -force-no-fast-pipeline."""
    import subprocess
    exe = os.path.join(_oracle.ROOT, "examples", "minimal_mipmaps")
    assert os.path.exists(exe), "examples/minimal_mipmaps has not been built (__graft_entry__.build())"
    w, h = 260, 136
    l0 = _oracle.smooth_level0(w, h, 9)
    src = str(tmp_path / "in.tga")
    nv.write_tga(src, l0, w, h)
    for args, pm, fg in (([], False, False), (["-do-premultiply-alpha"], True, False),
                         (["-premultiplied-alpha", "-force-no-fast-pipeline"], False, True)):
        base = str(tmp_path / ("out_%d%d.tga" % (pm, fg)))
        r = subprocess.run([exe, "-i", src, "-o", base] + args, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr
        want, _ = oracle.shader_chain(oracle.premultiply(l0) if pm else l0, w, h, force_general=fg)
        for level, v in enumerate(nv.level_views(want, w, h)):
            got, gw, gh = nv.read_image(nv.level_filename(base, level))
            assert (gh, gw) == v.shape[:2] and (got == v).all(), (args, level)
        assert ("Wrote " + nv.level_filename(base, 8)) in r.stderr
    r = subprocess.run([exe, "-bogus"], capture_output=True, text=True)
    assert r.returncode != 0 and "unrecognised option '-bogus'" in r.stderr
    r = subprocess.run([exe, "-i"], capture_output=True, text=True)
    assert r.returncode != 0 and "needs a file name" in r.stderr


def test_user_defined_functor_sets(nv, cuda):
    """include/nvpyr.cuh: the CUDA form of the reference's NVPRO_PYRAMID_* macro contract and dispatcher callbacks
    (nvpro_pyramid.glsl:27-120, nvpro_pyramid_dispatch.hpp:99-116).  examples/custom_functors runs a hi-z (max)
    pyramid over R32F images through every schedule variant -- bit-exact vs a CPU loop -- and a user-written copy
    of the RGBA32F instance that must equal the library's own, bit for bit."""
    import subprocess
    exe = os.path.join(_oracle.ROOT, "examples", "custom_functors")
    assert os.path.exists(exe), "examples/custom_functors has not been built (__graft_entry__.build())"
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "custom functor sets ok" in r.stdout
    assert r.stdout.count("0 of") == 36 + 1 + 4 + 21, r.stdout
    assert "-> rejected" in r.stdout
    # the set with its own NVPRO_PYRAMID_LOAD_REDUCE4 (glsl:78-88): the hook runs for the first level of every fast
    # dispatch and nowhere else (minima there, maxima elsewhere, checked against a CPU loop that follows the plan)
    assert r.stdout.count("MinFirst") == 21 and "2 fast dispatches use the hook, 0 of" in r.stdout


@pytest.mark.parametrize("size", [(1024, 512), (333, 201), (1920, 1080)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_recorded_once_replayed_per_frame(nv, cuda, oracle, size):
    """The reference RECORDS the pyramid into a command buffer and submits it every frame
    (nvpro_pyramid_dispatch.hpp:42-53; demo_app/mipmaps_app.cpp:363-369).  The CUDA form: the dispatch is captured
    into a CUDA graph once and the graph is replayed for new contents of level 0 -- programmatic dependent launches,
    the tail kernel's ticket counter and the cached launch configuration must all survive capture and replay."""
    w, h = size
    buf = cuda.zeros(nv.chain_bytes(w, h), dtype=cuda.uint8, device="cuda")
    pipes = nv.PyramidPipelines()
    nv.cmd_pyramid_dispatch(None, pipes, w, h, image=buf)  # warm: per-device context and launch configs exist
    cuda.cuda.synchronize()
    g = cuda.cuda.CUDAGraph()
    with cuda.cuda.graph(g):
        nv.cmd_pyramid_dispatch(None, pipes, w, h, image=buf)
    for frame in range(3):
        l0 = _oracle.random_level0(w, h, 100 + frame)
        buf.zero_()
        buf[:4 * w * h] = cuda.from_numpy(l0).cuda()
        g.replay()
        cuda.cuda.synchronize()
        want, _ = oracle.shader_chain(l0, w, h)
        assert_same(buf.cpu().numpy(), want, w, h, oracle, f"graph replay {frame}")


def _fast_dispatcher(div, max_levels):
    """nvproPyramidDefaultFastDispatcher<div, max_levels> (nvpro_pyramid_dispatch.hpp:195-242) written in Python."""
    def f(state, step):
        if state.currentX % div or state.currentY % div:
            return 0
        n = 0
        while n < min(state.remainingLevels, max_levels) and not ((state.currentX >> n) & 1) and not ((state.currentY >> n) & 1):
            n += 1
        return n
    return f


def test_user_dispatcher_callbacks(nv, cuda, oracle):
    """nvproCmdPyramidDispatch's 7-argument overload (nvpro_pyramid_dispatch.hpp:109-116) through
    nvpyrDispatchWithDispatchers: user callbacks decide pipeline and level count per dispatch.  Callbacks that
    restate the default dispatchers and the <2, 5> alternative must give the oracle's chains for those schedules; a
    general dispatcher that fills one level per dispatch must equal level-by-level generation; bad dispatchers are
    rejected before anything is enqueued."""
    def run(w, h, l0, general=None, fast=None, flags=0):
        buf = cuda.zeros(nv.chain_bytes(w, h), dtype=cuda.uint8, device="cuda")
        buf[:4 * w * h] = cuda.from_numpy(l0).cuda()
        nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(), w, h, 0, general, fast, image=buf, flags=flags)
        cuda.cuda.synchronize()
        return buf.cpu().numpy()

    general2 = lambda state, step: min(2, state.remainingLevels)
    for (w, h) in [(1024, 512), (260, 260), (333, 201), (1920, 1080)]:
        l0 = _oracle.random_level0(w, h, 21)
        want, _ = oracle.shader_chain(l0, w, h)
        assert_same(run(w, h, l0, general2, _fast_dispatcher(4, 6)), want, w, h, oracle, "default dispatchers restated")
        want25, _ = oracle.shader_chain(l0, w, h, div=2, max_levels=5)
        assert_same(run(w, h, l0, None, _fast_dispatcher(2, 5)), want25, w, h, oracle, "fast <2, 5> callback")
        wantg, _ = oracle.shader_chain(l0, w, h, force_general=True)
        assert_same(run(w, h, l0, general2, lambda s, st: 0), wantg, w, h, oracle, "fast callback never eligible")

    # one level per general dispatch == generating level k+1 from level k with separate two-level-count dispatches
    w, h = 333, 201
    l0 = _oracle.random_level0(w, h, 22)
    got = run(w, h, l0, lambda s, st: 1, None, flags=nv.FLAG_FORCE_GENERAL)
    buf = cuda.zeros(nv.chain_bytes(w, h), dtype=cuda.uint8, device="cuda")
    buf[:4 * w * h] = cuda.from_numpy(l0).cuda()
    n = nv.level_count(w, h)
    for k in range(n - 1):
        lw, lh = nv.level_extent(w, h, k)
        ptrs = [buf.data_ptr() + 4 * nv.level_offset_texels(w, h, k + i) for i in range(2)]
        nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(fast_pipeline=False), lw, lh, 2, image=None, level_ptrs=ptrs)
    cuda.cuda.synchronize()
    assert_same(got, buf.cpu().numpy(), w, h, oracle, "one level per dispatch")

    # rejected plans: general fills nothing / too much, fast promises levels the size does not allow, 3-level general
    before = nv.launch_count()
    for general, fast in ((lambda s, st: 0, None), (lambda s, st: s.remainingLevels + 1, None),
                          (None, lambda s, st: min(3, s.remainingLevels)), (lambda s, st: min(3, s.remainingLevels), lambda s, st: 0)):
        with pytest.raises(nv.NvpyrError) as e:
            run(333, 201, l0, general, fast)
        assert e.value.status == 1  # NVPYR_ERROR_INVALID_VALUE
    assert nv.launch_count() == before


def test_other_stream(nv, cuda, oracle):
    w, h = 320, 192
    l0 = _oracle.random_level0(w, h, 13)
    want, _ = oracle.shader_chain(l0, w, h)
    s = cuda.cuda.Stream()
    buf = cuda.zeros(nv.chain_bytes(w, h), dtype=cuda.uint8, device="cuda")
    buf[:4 * w * h] = cuda.from_numpy(l0).cuda()
    cuda.cuda.synchronize()
    with cuda.cuda.stream(s):
        nv.cmd_pyramid_dispatch(s, nv.PyramidPipelines(), w, h, image=buf)
    s.synchronize()
    assert (buf.cpu().numpy() == want).all()


def test_reentrant_from_several_host_threads(nv, cuda, oracle):
    """SURVEY 8b threading contract: the entry points are reentrant and safe from several host threads on
    different streams (dispatch, batch and the host round trip share per-device state: ticket pool, base
    pointer ring, staging chain).  Eight threads, each with its own stream and images, many rounds."""
    import threading
    sizes = [(1024, 512), (260, 260), (255, 131), (1024, 1024)]
    jobs = []
    for t in range(8):
        w, h = sizes[t % len(sizes)]
        l0 = _oracle.random_level0(w, h, 900 + t)
        jobs.append((w, h, l0, oracle.shader_chain(l0, w, h)[0]))
    errors = []

    def worker(t):
        try:
            w, h, l0, want = jobs[t]
            stream = cuda.cuda.Stream()
            bufs = [cuda.zeros(nv.chain_bytes(w, h), dtype=cuda.uint8, device="cuda") for _ in range(3)]
            for rnd in range(6):
                for b in bufs:
                    b.zero_()
                    b[:4 * w * h] = cuda.from_numpy(l0).cuda()
                cuda.cuda.synchronize()
                nv.cmd_pyramid_dispatch(stream, nv.PyramidPipelines(), w, h, image=bufs[0])
                nv.dispatch_batch(stream, nv.PyramidPipelines(), bufs[1:], w, h)
                host = nv.generate_host(l0, w, h)
                stream.synchronize()
                for b in bufs:
                    if not (b.cpu().numpy() == want).all():
                        errors.append((t, rnd, "device"))
                if not (host == want).all():
                    errors.append((t, rnd, "host"))
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:5]


def test_negative_control_is_detected(nv, cuda, oracle):
    """The harness must see a wrong generator (reference's `null` / `baseline` alternatives)."""
    w, h = 128, 128
    l0 = _oracle.random_level0(w, h, 14, opaque=True)
    want, _ = oracle.shader_chain(l0, w, h)
    got = gpu_chain(nv, cuda, l0, w, h, levels=3)  # only two levels filled ...
    full = np.concatenate([got, np.tile(MAGENTA, (want.size - got.size) // 4)])  # ... the rest stays magenta
    assert oracle.compare(full, want, w, h).worst > 100


@pytest.mark.parametrize("size", [(1024, 1024), (2048, 1024), (512, 2048), (1024, 768), (1040, 528), (1056, 1056)])
def test_slab_tasks_of_the_fast_kernel(nv, cuda, oracle, size):
    """Images with few 64 x 2^M tiles for the resident warps run the tuned fast kernel in slab-task mode (a warp
    takes one 64x8 slab; the last warp to arrive at a tile finishes levels +4..+M): M = 6, 5 (1056 = 32 * 33)
    and 4 (1040 = 16 * 65, 528 = 16 * 33); same bits as the tile mode (NVPYR_NO_SLAB_TASKS=1 covers that one in
    every other test of sizes >= 4096 tiles)."""
    w, h = size
    l0 = _oracle.random_level0(w, h, 55)
    want, _ = oracle.shader_chain(l0, w, h)
    for _ in range(3):  # arrival order varies from run to run
        got = gpu_chain(nv, cuda, l0, w, h)
        assert_same(got, want, w, h, oracle, "slab tasks")


@pytest.mark.parametrize("staged", ["1", "0"], ids=["cp.async-staged-rows", "per-lane-loads"])
def test_four_column_strip_kernel_on_every_size(nv, cuda, oracle, staged):
    """generalStrip4Kernel normally takes only large levels; NVPYR_GEN_STRIP4_MIN_TEXELS=0 (with the tail fusion
    off so that small levels reach the stand-alone kernels) runs all its variants -- 1 / 2 levels, 2 or 3 taps
    per axis on either level, ragged strips and segments -- on sizes the oracle handles quickly.  Both ways of
    fetching the source rows: staged in shared memory by 4-byte cp.async copies (rows whose pitch is not a multiple
    of 16 bytes are realigned on the way; the default) and with per-lane 4-byte loads (NVPYR_GEN_STAGED=0)."""
    import subprocess
    import sys
    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "import vk_compute_mipmaps_b200 as nv, _oracle\n"
        "o = _oracle.load_oracle()\n"
        "sizes = [(63, 63), (100, 37), (255, 255), (511, 300), (254, 254), (17, 513), (129, 129), (6, 10), (1023, 511),\n"
        "         (1022, 766), (777, 1031), (1200, 900), (125, 3), (126, 126), (127, 2), (249, 251), (250, 250), (2047, 700)]\n"
        "for (w, h) in sizes:\n"
        "    for fg in (True, False):\n"
        "        l0 = _oracle.random_level0(w, h, w * 7 + h)\n"
        "        buf = torch.zeros(nv.chain_bytes(w, h), dtype=torch.uint8, device='cuda')\n"
        "        buf[:4 * w * h] = torch.from_numpy(l0).cuda()\n"
        "        nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(fast_pipeline=not fg), w, h, image=buf)\n"
        "        torch.cuda.synchronize()\n"
        "        assert (buf.cpu().numpy() == o.shader_chain(l0, w, h, force_general=fg)[0]).all(), (w, h, fg)\n"
        "print('strip4 ok')\n") % (_oracle.ROOT, os.path.join(_oracle.ROOT, "tests"))
    env = dict(os.environ, NVPYR_GEN_STRIP4_MIN_TEXELS="0", NVPYR_TAIL_MAX_TEXELS="0", NVPYR_GEN_STAGED=staged)
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "strip4 ok" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.parametrize("size", [(4096, 4096), (4095, 4095), (2047, 2047), (4094, 2049)], ids=lambda s: f"{s[0]}x{s[1]}")
def test_baseline_config_sizes_bit_exact(nv, cuda, oracle, size):
    """BASELINE configs 1 and 2 at full size (oracle runs ~1 s)."""
    w, h = size
    l0 = _oracle.random_level0(w, h, 21)
    want, _ = oracle.shader_chain(l0, w, h)
    got = gpu_chain(nv, cuda, l0, w, h)
    assert_same(got, want, w, h, oracle, "full size")


def test_headline_16384_properties(nv, cuda, oracle):
    """BASELINE config 3 (16384^2, 1.43 GB chain): too big for the oracle in seconds, so check through
    size-independent properties: (a) levels 1..6 of random 64x64-aligned tiles equal the oracle chain of the
    crop (a 6-level fast step never looks outside its tile); (b) levels 7..14 equal the oracle chain of the
    GPU's own level 6 taken as a 256x256 image (same carry groups: fast 6 + fast 2); (c) no magenta left;
    (d) a constant image stays constant at every level."""
    torch = cuda
    w = h = 16384
    n = nv.chain_bytes(w, h)
    buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    buf.view(-1, 4)[:] = torch.from_numpy(MAGENTA).cuda()
    g = torch.Generator(device="cuda").manual_seed(0)
    buf[:4 * w * h] = torch.randint(0, 256, (4 * w * h,), dtype=torch.uint8, device="cuda", generator=g)
    nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(), w, h, image=buf)
    torch.cuda.synchronize()
    views = nv.level_views(buf, w, h)
    rng = np.random.default_rng(0)
    for _ in range(12):
        tx, ty = int(rng.integers(0, w // 64)), int(rng.integers(0, h // 64))
        crop = views[0][ty * 64:(ty + 1) * 64, tx * 64:(tx + 1) * 64].cpu().numpy().reshape(-1)
        want, _ = oracle.shader_chain(crop, 64, 64)
        off = 64 * 64
        for lvl in range(1, 7):
            e = 64 >> lvl
            got = views[lvl][ty * e:(ty + 1) * e, tx * e:(tx + 1) * e].cpu().numpy().reshape(-1)
            assert (got == want[4 * off:4 * (off + e * e)]).all(), (tx, ty, lvl)
            off += e * e
    l6 = views[6].cpu().numpy().reshape(-1)
    want_tail, _ = oracle.shader_chain(l6, 256, 256)
    off6 = nv.level_offset_texels(w, h, 6)
    got_tail = buf[4 * off6:].cpu().numpy()
    assert (got_tail == want_tail).all()
    rest = buf[4 * w * h:].view(-1, 4)
    assert not bool((rest == torch.from_numpy(MAGENTA).cuda()).all(dim=1).any())
    del views, rest
    buf.view(-1, 4)[:w * h] = torch.tensor([10, 128, 250, 77], dtype=torch.uint8, device="cuda")
    nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(), w, h, image=buf)
    torch.cuda.synchronize()
    assert bool((buf.view(-1, 4) == torch.tensor([10, 128, 250, 77], dtype=torch.uint8, device="cuda")).all())


# ------------------------------------------------------------------------------------------------------------------
# Round 2: the configurations the numbers are quoted on, checked in full.

def _crop_chains(oracle, level0_hw4, edge, levels, workers=None):
    """Oracle A on every edge x edge crop of level 0 (numpy [H, W, 4]), `levels` levels each, across host threads
    (the C oracle releases the GIL).  A fast-pipeline step of M levels never looks outside its 2^M-aligned tile, so
    for edge % 64 == 0 the crops of a 6-level step are exact.  Returns {(cy, cx): chain bytes}."""
    import concurrent.futures as cf
    h, w = level0_hw4.shape[:2]
    jobs = [(cy, cx) for cy in range(0, h, edge) for cx in range(0, w, edge)]

    def run(job):
        cy, cx = job
        crop = np.ascontiguousarray(level0_hw4[cy:cy + edge, cx:cx + edge]).reshape(-1)
        return job, oracle.shader_chain(crop, edge, edge, levels=levels)[0]
    with cf.ThreadPoolExecutor(workers or max(1, len(os.sched_getaffinity(0)))) as ex:
        return dict(ex.map(run, jobs))


@pytest.mark.parametrize("content", ["julia", "random", "gradient"])
def test_headline_16384_every_tile(nv, cuda, oracle, content):
    """BASELINE config 3 in full, on the three inputs bench.py times (the Julia set of the reference demo, uniform
    random bytes, the smooth gradient): levels 1..6 of EVERY 64x64 tile against Oracle A (64 crops of 2048^2, each
    a 7-level chain), then levels 7..14 against the oracle chain of the GPU's own level 6 (same carry groups:
    fast 6 | fast 6 | fast 2)."""
    import bench
    torch = cuda
    w = h = 16384
    n = nv.chain_bytes(w, h)
    buf = torch.empty(n, dtype=torch.uint8, device="cuda")
    buf.view(-1, 4)[w * h:] = torch.from_numpy(MAGENTA).cuda()
    l0 = buf[:4 * w * h].view(h, w, 4)
    if content == "random":
        g = torch.Generator(device="cuda").manual_seed(7)
        buf[:4 * w * h] = torch.randint(0, 256, (4 * w * h,), dtype=torch.uint8, device="cuda", generator=g)
    else:
        {"julia": bench.fill_julia, "gradient": bench.fill_gradient}[content](l0, w, h)
    host_l0 = l0.cpu().numpy()
    nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(), w, h, image=buf)
    torch.cuda.synchronize()
    assert bool((buf[:4 * w * h].view(h, w, 4).cpu() == torch.from_numpy(host_l0)).all()), "level 0 was modified"
    edge = 2048
    chains = _crop_chains(oracle, host_l0, edge, 7)
    views = [v.cpu().numpy() for v in nv.level_views(buf, w, h)[:7]]
    for (cy, cx), want in chains.items():
        off = edge * edge
        for lvl in range(1, 7):
            e = edge >> lvl
            got = views[lvl][cy >> lvl:(cy >> lvl) + e, cx >> lvl:(cx >> lvl) + e].reshape(-1)
            assert (got == want[4 * off:4 * (off + e * e)]).all(), (content, cx, cy, lvl)
            off += e * e
    want_tail, _ = oracle.shader_chain(views[6].reshape(-1), 256, 256)
    off6 = nv.level_offset_texels(w, h, 6)
    assert (buf[4 * off6:].cpu().numpy() == want_tail).all(), content


IMAGES_DIR = os.path.join(os.path.dirname(__file__), "golden", "test_images")
# file, recorded worst delta of the reference's GPU "default" pipeline vs its CPU generator (demo_app/rtx3090.json),
# worst delta of Oracle A vs the real CPU generator on the PIL-decoded, premultiplied image (computed in the build
# container: the alpha images reproduce the recorded 5 / 4 / 4, the opaque ones stay at or below the recorded 2)
REFERENCE_IMAGES = [("1080p.jpg", 2, 1), ("1440p.jpg", 2, 1), ("4094.jpg", 2, 2), ("4095.jpg", 2, 2), ("4096.jpg", 2, 2),
                    ("4k.jpg", 2, 1), ("alpha1080p.png", 5, 5), ("alpha2048.png", 4, 4), ("alpha2052.png", 4, 4),
                    ("lunch_2047.jpg", 2, 2), ("lunch_with_friend.jpg", 2, 2), ("mandelbrots.png", 2, 2), ("tall.jpg", 2, 1)]


@pytest.mark.parametrize("name,recorded,expected", REFERENCE_IMAGES, ids=[r[0] for r in REFERENCE_IMAGES])
def test_reference_test_images_full_size(nv, cuda, oracle, name, recorded, expected):
    """BASELINE configs 1, 2 and 4 on the reference's own test images at FULL size (test_images/*, shipped under
    tests/golden/test_images): level 0 = the decoded image, premultiplied by the library as the reference's loader
    does (scoped_image.hpp:233-255, mipmaps_app.cpp:606).  Bit-exact against Oracle A, and the worst delta against
    the reference's own CPU generator (oracle/_ref when present, else Oracle B, its pinned restatement) equals the
    known answer."""
    from PIL import Image
    im = Image.open(os.path.join(IMAGES_DIR, name)).convert("RGBA")
    w, h = im.size
    raw = np.asarray(im, dtype=np.uint8).reshape(-1).copy()
    l0 = oracle.premultiply(raw)
    want, _ = oracle.shader_chain(l0, w, h)
    got = gpu_chain(nv, cuda, raw, w, h, flags=2)  # NVPYR_FLAG_PREMULTIPLY_ALPHA: the pre-pass runs on the GPU
    assert_same(got, want, w, h, oracle, name)
    ref = _oracle.load_ref()
    cpu = ref.cpu_chain(oracle.new_chain(l0, w, h), w, h) if ref is not None else oracle.cpu_chain(l0, w, h)
    worst = oracle.compare(got, cpu, w, h).worst
    assert worst == expected and worst <= recorded, (name, worst, expected, recorded)


def test_fused_batch_at_the_benchmarked_size(nv, cuda, oracle):
    """BASELINE config 5's unit: 4096^2 textures through nvpyrDispatchBatch (fastSrgba8Kernel<6, batch> +
    tailBatchKernel, two launches), every texture bit-exact against Oracle A."""
    import concurrent.futures as cf
    w = h = 4096
    count = 8
    l0s = [_oracle.random_level0(w, h, 900 + k) if k % 2 == 0 else _oracle.smooth_level0(w, h, 900 + k) for k in range(count)]
    with cf.ThreadPoolExecutor(max(1, len(os.sched_getaffinity(0)))) as ex:
        wants = list(ex.map(lambda l0: oracle.shader_chain(l0, w, h)[0], l0s))
    imgs = []
    for l0 in l0s:
        buf = cuda.empty(nv.chain_bytes(w, h), dtype=cuda.uint8, device="cuda")
        buf.view(-1, 4)[:] = cuda.from_numpy(MAGENTA).cuda()
        buf[:4 * w * h] = cuda.from_numpy(l0).cuda()
        imgs.append(buf)
    before = nv.launch_count()
    nv.dispatch_batch(None, nv.PyramidPipelines(), imgs, w, h)
    cuda.cuda.synchronize()
    assert nv.launch_count() - before == 2
    for k, (b, want) in enumerate(zip(imgs, wants)):
        assert_same(b.cpu().numpy(), want, w, h, oracle, f"texture {k}")


def test_batch_recorded_into_a_graph(nv, cuda, oracle):
    """nvpyrDispatchBatch under stream capture must not record its base-pointer upload (a copy from a
    host vector that is gone at replay time): it falls back to one dispatch per image while capturing, and the
    replayed graph produces the oracle's bits for new level-0 contents, also while live batches keep running."""
    w, h = 1024, 512
    count = 4
    nv.init()
    bufs = [cuda.zeros(nv.chain_bytes(w, h), dtype=cuda.uint8, device="cuda") for _ in range(count)]
    live = [cuda.zeros(nv.chain_bytes(w, h), dtype=cuda.uint8, device="cuda") for _ in range(count)]
    pipes = nv.PyramidPipelines()
    nv.dispatch_batch(None, pipes, bufs, w, h)  # warm: launch configurations cached
    cuda.cuda.synchronize()
    g = cuda.cuda.CUDAGraph()
    with cuda.cuda.graph(g):
        nv.dispatch_batch(None, pipes, bufs, w, h)
    side = cuda.cuda.Stream()
    for frame in range(3):
        l0s = [_oracle.random_level0(w, h, 40 + 10 * frame + k) for k in range(count)]
        for b, lb, l0 in zip(bufs, live, l0s):
            b.zero_()
            lb.zero_()
            b[:4 * w * h] = cuda.from_numpy(l0).cuda()
            lb[:4 * w * h] = b[:4 * w * h]
        cuda.cuda.synchronize()
        g.replay()
        nv.dispatch_batch(side, pipes, live, w, h)  # a live fused batch on another stream at the same time
        cuda.cuda.synchronize()
        for k, l0 in enumerate(l0s):
            want, _ = oracle.shader_chain(l0, w, h)
            assert_same(bufs[k].cpu().numpy(), want, w, h, oracle, f"replayed batch, frame {frame}, image {k}")
            assert_same(live[k].cpu().numpy(), want, w, h, oracle, f"live batch, frame {frame}, image {k}")


def test_ticket_counters_are_never_shared_between_unordered_launches(nv, cuda, oracle):
    """the tail kernel's 'last CTA' counter belongs to ONE stream (or to one captured launch).  Many
    streams, two graphs replayed concurrently with live dispatches: every chain must still be complete."""
    w, h = 1920, 1080  # fast 3, then a tail launch with a grid step and solo steps
    l0 = _oracle.random_level0(w, h, 77)
    want, _ = oracle.shader_chain(l0, w, h)
    pipes = nv.PyramidPipelines()
    nv.init()
    streams = [cuda.cuda.Stream() for _ in range(24)]
    bufs = [cuda.zeros(nv.chain_bytes(w, h), dtype=cuda.uint8, device="cuda") for _ in range(len(streams) + 2)]
    nv.cmd_pyramid_dispatch(None, pipes, w, h, image=bufs[0])
    cuda.cuda.synchronize()
    graphs = []
    for b in bufs[-2:]:
        g = cuda.cuda.CUDAGraph()
        with cuda.cuda.graph(g):
            nv.cmd_pyramid_dispatch(None, pipes, w, h, image=b)
        graphs.append(g)
    for rnd in range(5):
        for b in bufs:
            b.zero_()
            b[:4 * w * h] = cuda.from_numpy(l0).cuda()
        cuda.cuda.synchronize()
        for g in graphs:
            g.replay()
        for s, b in zip(streams, bufs):
            nv.cmd_pyramid_dispatch(s, pipes, w, h, image=b)
        for g in graphs:
            g.replay()
        cuda.cuda.synchronize()
        for k, b in enumerate(bufs):
            assert_same(b.cpu().numpy(), want, w, h, oracle, f"round {rnd}, buffer {k}")


def test_slab_task_hand_off_stress(nv, cuda, oracle):
    """The slab-task mode hands a tile's level +3 sums from the warps that made them to the warp that arrives last
    (store -> fence -> shared atomic -> fence -> load, no barrier).  compute-sanitizer's racecheck models barriers
    only, so the evidence is volume: thousands of launches over sizes with different tile / slab interleavings,
    every result compared ON THE DEVICE with the oracle's chain."""
    torch = cuda
    sizes = [(1024, 1024), (2048, 1024), (512, 2048), (1040, 528), (1056, 1056), (1024, 64), (64, 1024), (1984, 1088)]
    pipes = nv.PyramidPipelines()
    for (w, h) in sizes:
        l0 = _oracle.random_level0(w, h, w ^ h)
        want = torch.from_numpy(oracle.shader_chain(l0, w, h)[0]).cuda()
        n = nv.chain_bytes(w, h)
        bufs = [torch.zeros(n, dtype=torch.uint8, device="cuda") for _ in range(4)]
        for b in bufs:
            b[:4 * w * h] = want[:4 * w * h]
        bad = torch.zeros((), dtype=torch.int64, device="cuda")
        for it in range(400):
            b = bufs[it & 3]
            b[4 * w * h:].zero_()
            nv.cmd_pyramid_dispatch(None, pipes, w, h, image=b)
            bad += (b != want).sum()
        assert int(bad.item()) == 0, (w, h, int(bad.item()))


def test_concurrent_host_round_trips_and_pageable_callers(nv, cuda, oracle):
    """nvpyrGenerateHost takes a pipeline (scratch chain, streams, staging) from a per-device pool instead of
    holding one lock for the whole call: four host threads at once, pinned and PAGEABLE buffers (the latter pass
    through the pinned staging chain band by band), in place and with separate buffers, at a size that is banded
    with the default 32 MB bands (4096^2: 64 MB of level 0)."""
    import threading
    torch = cuda
    w = h = 4096
    n = nv.chain_bytes(w, h)
    l0 = _oracle.random_level0(w, h, 31)
    want, _ = oracle.shader_chain(l0, w, h)
    errors = []

    def worker(t):
        try:
            for rnd in range(2):
                if t % 2 == 0:  # pageable, in place
                    chain = np.zeros(n, dtype=np.uint8)
                    chain[:4 * w * h] = l0
                    nv.generate_host(chain[:4 * w * h], w, h, out=chain)
                    got = chain
                else:  # pinned in, pageable out
                    pin = torch.from_numpy(l0.copy()).pin_memory()
                    got = nv.generate_host(pin.numpy(), w, h)
                if not (got == want).all():
                    errors.append((t, rnd))
        except Exception as e:  # noqa: BLE001
            errors.append((t, repr(e)))
    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    [t.start() for t in threads]
    [t.join() for t in threads]
    assert not errors, errors
    # pageable premultiplied in place: level 0 comes back changed
    raw = _oracle.random_level0(2048, 8192, 5)
    pm = oracle.premultiply(raw)
    want2, _ = oracle.shader_chain(pm, 2048, 8192)
    chain = np.zeros(nv.chain_bytes(2048, 8192), dtype=np.uint8)
    chain[:raw.size] = raw
    nv.generate_host(chain[:raw.size], 2048, 8192, flags=nv.FLAG_PREMULTIPLY_ALPHA, out=chain)
    assert (chain == want2).all()
