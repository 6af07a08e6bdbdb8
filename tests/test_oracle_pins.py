"""Pins the CPU oracle (oracle/nvpyr_oracle.c) to the reference: its own code compiled in
place (oracle/_ref), its recorded known answers (demo_app/rtx3090.json), and the golden
fixtures generated from its test images (tools/make_golden.py)."""
import ctypes as C
import glob
import hashlib
import os

import numpy as np
import pytest

import _oracle

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))

# demo_app/rtx3090.json "default" deltas (GPU shader vs reference CPU generator):
# 2 on every opaque image, 5/4/4 on alpha1080p/alpha2048/alpha2052.
KNOWN_DELTA_OPAQUE, KNOWN_DELTA_ALPHA = 2, 5


def test_transfer_tables_match_formula(oracle):
    """Pinned tables == srgb.h formulas evaluated with this platform's powf."""
    lib = oracle.lib
    for c in range(256):
        assert lib.nvo_linear_from_srgb(c) == lib.nvo_linear_from_srgb_formula(c)
    rng = np.random.default_rng(0)
    xs = np.concatenate([rng.random(20000, dtype=np.float32), rng.random(5000, dtype=np.float32) * 0.01,
                         np.array([0.0, 1.0, 0.0031308, 0.5, 1.5, 2.0], dtype=np.float32)])
    for x in xs:
        assert lib.nvo_srgb_from_linear(float(x)) == lib.nvo_srgb_from_linear_formula(float(x))


def test_transfer_thresholds_are_tight(oracle):
    """Each pinned threshold is the FIRST float of its code (formula on both sides of it)."""
    import re
    src = open(os.path.join(_oracle.ROOT, "vk_compute_mipmaps_b200", "csrc", "srgb_tables.inc")).read()
    body = src.split("NVPYR_SRGB_ENCODE_THRESHOLD_BITS[255]")[1].split("};")[0]
    thr = [int(t, 16) for t in re.findall(r"0x([0-9a-f]{8})u", body)]
    assert len(thr) == 255 and thr == sorted(thr)
    f = oracle.lib.nvo_srgb_from_linear_formula
    for c, bits in enumerate(thr, start=1):
        at = np.array([bits], dtype=np.uint32).view(np.float32)[0]
        below = np.array([bits - 1], dtype=np.uint32).view(np.float32)[0]
        assert f(float(at)) == c and f(float(below)) == c - 1


def test_committed_tables_are_the_generator_output(tmp_path):
    """vk_compute_mipmaps_b200/csrc/srgb_tables.inc is exactly what tools/gen_srgb_tables.c prints on this platform
    (the pinned bit patterns were not edited by hand; the exhaustive monotonicity pass is the generator's
    --exhaustive option, ~10 s, not run here)."""
    import shutil, subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "gen")
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", os.path.join(_oracle.ROOT, "tools", "gen_srgb_tables.c"), "-lm", "-o", exe])
    out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout
    assert out == open(os.path.join(_oracle.ROOT, "vk_compute_mipmaps_b200", "csrc", "srgb_tables.inc")).read()


def test_transfer_matches_reference_header(oracle, ref):
    for c in range(256):
        assert oracle.lib.nvo_linear_from_srgb(c) == ref.lib.ref_linear_from_srgb(c)
    rng = np.random.default_rng(1)
    for x in rng.random(30000, dtype=np.float32):
        assert oracle.lib.nvo_srgb_from_linear(float(x)) == ref.lib.ref_srgb_from_linear(float(x))
    # round trip identity: premultiply is a no-op on opaque texels
    for c in range(256):
        assert oracle.lib.nvo_srgb_from_linear(oracle.lib.nvo_linear_from_srgb(c)) == c


@pytest.mark.parametrize("size", [(64, 64), (63, 63), (17, 5), (1, 9), (9, 1), (100, 37), (255, 256), (2, 1), (1, 2),
                                  (333, 1), (1, 1), (3, 3), (2, 2), (129, 65), (300, 199)])
def test_oracle_b_equals_reference_cpu_generator(oracle, ref, size):
    w, h = size
    l0 = _oracle.random_level0(w, h, w * 1000 + h)
    ours = oracle.cpu_chain(l0, w, h)
    theirs = ref.cpu_chain(oracle.new_chain(l0, w, h), w, h)
    assert (ours == theirs).all()


def test_comparator_equals_reference(oracle, ref):
    w, h = 97, 61
    a = oracle.cpu_chain(_oracle.random_level0(w, h, 1), w, h)
    b = a.copy()
    rng = np.random.default_rng(2)
    idx = rng.integers(4 * w * h, a.size, 50)
    b[idx] = rng.integers(0, 256, 50, dtype=np.uint8)
    c = oracle.compare(a, b, w, h)
    d, (x, y, lvl, ch) = ref.compare(a, b, w, h)
    assert (c.worst, c.x, c.y, c.level, c.channel) == (d, x, y, lvl, ch)
    b2 = a.copy()
    b2[:4 * w * h] ^= 0xFF  # level 0 is skipped by the comparator (mipmap_storage.hpp:170)
    assert oracle.compare(a, b2, w, h).worst == 0


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_fixtures(oracle, path):
    g = np.load(path)
    w, h = int(g["width"]), int(g["height"])
    l0 = g["level0"].reshape(-1)
    b = oracle.cpu_chain(l0, w, h)
    assert hashlib.sha256(b.tobytes()).hexdigest() == str(g["ref_cpu_sha256"]), "Oracle B != reference CPU chain"
    a, stores = oracle.shader_chain(l0, w, h)
    assert hashlib.sha256(a.tobytes()).hexdigest() == str(g["oracle_a_sha256"])
    assert [(s.pipeline, s.input_level, s.level_count) for s in oracle.plan(w, h)] == [tuple(p) for p in g["plan"]]
    c = oracle.compare(a, b, w, h)
    assert c.worst == int(g["delta_a_vs_ref"])
    opaque = bool((g["level0"][..., 3] == 255).all())
    assert c.worst <= (KNOWN_DELTA_OPAQUE if opaque else KNOWN_DELTA_ALPHA)


@pytest.mark.parametrize("size,opaque", [((256, 256), True), ((255, 255), True), ((260, 260), False),
                                         ((136, 512), True), ((120, 72), False), ((64, 64), True)])
def test_oracle_a_within_recorded_deltas(oracle, size, opaque):
    """Known-answer bound of rtx3090.json on smooth synthetic images (premultiplied when not opaque)."""
    w, h = size
    l0 = _oracle.smooth_level0(w, h, 5)
    if opaque:
        l0[3::4] = 255
    else:
        l0 = oracle.premultiply(l0)
    a, _ = oracle.shader_chain(l0, w, h)
    b = oracle.cpu_chain(l0, w, h)
    assert oracle.compare(a, b, w, h).worst <= (KNOWN_DELTA_OPAQUE if opaque else KNOWN_DELTA_ALPHA)


@pytest.mark.parametrize("size", [(64, 64), (128, 64), (256, 256), (4, 4), (8, 8), (16, 48), (32, 32), (96, 160),
                                  (63, 63), (100, 37), (1, 50), (50, 1), (5, 5), (2, 2), (33, 2), (260, 260), (72, 520)])
def test_oracle_a_coverage(oracle, size):
    """Every texel of every level >= 1 is stored (magenta pre-fill, mipmaps_app.cpp:345-361), and the number
    of stores equals texels + the general pipeline's duplicated halo stores."""
    w, h = size
    l0 = _oracle.random_level0(w, h, 9, opaque=True)
    buf = oracle.new_chain(l0, w, h)
    magenta = np.tile(np.array([255, 0, 255, 254], dtype=np.uint8), buf.size // 4 - w * h)
    buf[4 * w * h:] = magenta
    import ctypes as C
    st = C.c_uint64()
    oracle.lib.nvo_shader_chain(0, buf.ctypes.data, w, h, 0, 0, 4, 6, C.byref(st))
    rest = buf[4 * w * h:].reshape(-1, 4)
    assert not (rest == np.array([255, 0, 255, 254], dtype=np.uint8)).all(axis=1).any()
    assert st.value >= rest.shape[0]
    # alpha of an opaque image stays 255 in shader order (round, not truncate)
    assert (rest[:, 3] == 255).all()


def test_oracle_a_force_general_and_fast_agree_on_even_levels(oracle):
    """Level 1 of a pow2 image: the general pipeline's REDUCE2 o REDUCE2 equals the fast pipeline's
    0.25*((a+b)+(c+d)) with vertical pairing bit for bit (SURVEY section 8a note 1)."""
    w = h = 64
    l0 = _oracle.random_level0(w, h, 4)
    fast, _ = oracle.shader_chain(l0, w, h)
    gen, _ = oracle.shader_chain(l0, w, h, force_general=True)
    n1 = 4 * (w * h + (w // 2) * (h // 2))
    assert (fast[:n1] == gen[:n1]).all()


def test_oracle_a_rgba32f_close_to_float64(oracle):
    w, h = 96, 80
    l0 = _oracle.random_level0(w, h, 3, fmt=1)
    a, _ = oracle.shader_chain(l0, w, h, fmt=1)
    # float64 box filter, level by level
    cur = l0.reshape(h, w, 4).astype(np.float64)
    off = w * h
    cw, ch = w, h
    while cw % 2 == 0 and ch % 2 == 0 and cw > 1:
        cur = 0.25 * (cur[0::2, 0::2] + cur[1::2, 0::2] + cur[0::2, 1::2] + cur[1::2, 1::2])
        cw //= 2
        ch //= 2
        got = a[4 * off:4 * (off + cw * ch)].reshape(ch, cw, 4)
        np.testing.assert_allclose(got, cur, rtol=1e-6, atol=0)
        off += cw * ch


# ---------------------------------------------------------------------------------------------------------
# Oracle A against the reference's OWN shader sources, executed (oracle/glsl_emu): nvproCmdPyramidDispatch from
# nvpro_pyramid_dispatch.hpp drives nvpro_pyramid.glsl + srgba8_mipmap_preamble.glsl compiled as C++, one fiber
# per invocation, shuffles and barriers as scheduling points.

EMU_SIZES = [(64, 64), (256, 256), (128, 64), (64, 128), (16, 48), (32, 32), (96, 160), (8, 8), (4, 4), (192, 320),
             (63, 63), (100, 37), (260, 260), (136, 512), (120, 72), (160, 96), (240, 144), (1, 9), (9, 1), (5, 5),
             (2, 2), (33, 2), (255, 255), (254, 254), (17, 129), (1, 1 << 5), (3, 3), (6, 10), (129, 129), (448, 64),
             (512, 512), (511, 300)]


@pytest.mark.parametrize("size", EMU_SIZES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_oracle_a_equals_executed_reference_shaders(oracle, emu, size):
    w, h = size
    for seed, make in ((1, _oracle.random_level0), (2, lambda w, h, s: _oracle.smooth_level0(w, h, s))):
        l0 = make(w, h, seed)
        for have_fast in (1, 0):
            want, stores = oracle.shader_chain(l0, w, h, force_general=not have_fast)
            got, dispatches, emu_stores = emu.run_chain(oracle.new_chain(l0, w, h), w, h, 0, have_fast)
            assert (got == want).all(), (size, have_fast)
            assert emu_stores == stores  # same number of imageStore calls, duplicates included
            assert dispatches == len(oracle.plan(w, h, 0, have_fast))


def test_oracle_a_equals_executed_reference_shaders_partial_levels(oracle, emu):
    w, h = 256, 128
    l0 = _oracle.random_level0(w, h, 3)
    for levels in (2, 3, 5, 8):
        want, _ = oracle.shader_chain(l0, w, h, levels=levels)
        got, _, _ = emu.run_chain(oracle.new_chain(l0, w, h, levels=levels), w, h, levels, 1)
        assert (got == want).all(), levels


F16_SIZES = [(64, 64), (256, 256), (128, 64), (16, 48), (32, 32), (96, 160), (192, 320), (63, 63), (100, 37), (260, 260),
             (136, 512), (120, 72), (255, 255), (17, 129), (129, 129), (511, 300)]


@pytest.mark.parametrize("size", F16_SIZES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_oracle_a_f16_shared_equals_executed_reference_shaders(oracle, emu, size):
    """The F16_SHARED build (srgba8_mipmap_preamble.glsl:103-108): the reference's shaders compiled with the macro
    set, f16vec4 = IEEE binary16 round-to-nearest-even, against Oracle A's restatement of it.  The variant must
    differ from the default build somewhere (else the test would prove nothing)."""
    w, h = size
    differs = False
    for seed, make in ((1, _oracle.random_level0), (2, lambda w, h, s: _oracle.smooth_level0(w, h, s))):
        l0 = make(w, h, seed)
        for have_fast in (1, 0):
            want, stores = oracle.shader_chain(l0, w, h, force_general=not have_fast, f16_shared=True)
            got, _, emu_stores = emu.run_chain(oracle.new_chain(l0, w, h), w, h, 0, have_fast, f16_shared=1)
            assert (got == want).all(), (size, have_fast)
            assert emu_stores == stores
            differs |= bool((want != oracle.shader_chain(l0, w, h, force_general=not have_fast)[0]).any())
    if max(w, h) >= 64:
        assert differs


@pytest.mark.parametrize("size", F16_SIZES, ids=lambda s: f"{s[0]}x{s[1]}")
def test_oracle_a_srgb_shared_equals_executed_reference_shaders(oracle, emu, size):
    """The SRGB_SHARED build (srgba8_mipmap_preamble.glsl:60-101, the demo's "srgbShared" alternative): values that
    pass through shared memory are packed to 8-bit sRGB (packUnorm4x8 of srgbComponentFromLinear) and unpacked again.
    The reference's shaders compiled with the macro set and executed, against Oracle A's restatement through the
    pinned SHARED tables (tools/gen_srgb_tables.c).  Must differ from the default build somewhere."""
    w, h = size
    differs = False
    for seed, make in ((1, _oracle.random_level0), (2, lambda w, h, s: _oracle.smooth_level0(w, h, s))):
        l0 = make(w, h, seed)
        for have_fast in (1, 0):
            want, stores = oracle.shader_chain(l0, w, h, force_general=not have_fast, srgb_shared=True)
            got, _, emu_stores = emu.run_chain(oracle.new_chain(l0, w, h), w, h, 0, have_fast, srgb_shared=1)
            assert (got == want).all(), (size, have_fast)
            assert emu_stores == stores
            differs |= bool((want != oracle.shader_chain(l0, w, h, force_general=not have_fast)[0]).any())
    if max(w, h) >= 64:
        assert differs


def test_srgb_shared_tables_equal_the_glsl_functions(oracle, emu):
    """Every code's unpack value and every pack threshold (and its predecessor) of the pinned SHARED tables against the
    preamble's own linearFromSrgbComponent / srgbComponentFromLinear + packUnorm4x8 / unpackUnorm4x8 as executed."""
    lib = oracle.lib
    lib.nvo_srgb_shared_unpack.restype = C.c_float
    lib.nvo_srgb_shared_unpack.argtypes = [C.c_uint32]
    lib.nvo_srgb_shared_pack.argtypes = [C.c_float]
    emu.lib.emu_srgb_shared_round_trip.restype = C.c_float
    emu.lib.emu_srgb_shared_round_trip.argtypes = [C.c_float, C.POINTER(C.c_uint32)]
    import re
    src = open(os.path.join(_oracle.ROOT, "vk_compute_mipmaps_b200", "csrc", "srgb_tables.inc")).read()
    thr = [int(t, 16) for t in re.findall(r"0x([0-9a-f]{8})u", src.split("NVPYR_SRGB_SHARED_ENCODE_THRESHOLD_BITS[255]")[1].split("};")[0])]
    assert len(thr) == 255
    code = C.c_uint32()
    for c, bits in enumerate(thr, start=1):
        for b, want in ((bits, c), (bits - 1, c - 1)):
            x = float(np.array([b], dtype=np.uint32).view(np.float32)[0])
            back = emu.lib.emu_srgb_shared_round_trip(x, C.byref(code))
            assert code.value == want == lib.nvo_srgb_shared_pack(x)
            assert np.float32(back).tobytes() == np.float32(lib.nvo_srgb_shared_unpack(want)).tobytes()


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(p)[:-4] for p in GOLDEN])
def test_golden_fixtures_through_executed_shaders(oracle, emu, path):
    g = np.load(path)
    w, h = int(g["width"]), int(g["height"])
    l0 = g["level0"].reshape(-1)
    got, _, _ = emu.run_chain(oracle.new_chain(l0, w, h), w, h)
    assert hashlib.sha256(got.tobytes()).hexdigest() == str(g["oracle_a_sha256"])


def test_glsl_encode_equals_pinned_thresholds(emu):
    """srgbFromLinear of srgba8_mipmap_preamble.glsl:110-121, as compiled by the shim, has exactly the pinned
    thresholds (so the GLSL twin and the C++ twin in shaders/srgb.h agree on every float)."""
    import re
    src = open(os.path.join(_oracle.ROOT, "vk_compute_mipmaps_b200", "csrc", "srgb_tables.inc")).read()
    thr = [int(t, 16) for t in re.findall(r"0x([0-9a-f]{8})u", src.split("NVPYR_SRGB_ENCODE_THRESHOLD_BITS[255]")[1].split("};")[0])]
    assert len(thr) == 255
    f = emu.lib.emu_glsl_srgb_from_linear
    for c, bits in enumerate(thr, start=1):
        at = np.array([bits], dtype=np.uint32).view(np.float32)[0]
        below = np.array([bits - 1], dtype=np.uint32).view(np.float32)[0]
        assert f(float(at)) == c and f(float(below)) == c - 1
