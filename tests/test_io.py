"""Image files either side of the path: the TGA writer that replaces stbi_write_tga / writeMipmapsTga
(include/mipmap_storage.hpp:441-479) and the reader that stands in for stbi_load (scoped_image.hpp:217-218).
Host-only code of libnvpyr.so: no GPU needed.  stb itself is not available here, so byte parity with stb is
unpinned; what IS pinned: header, texel order, row order and packet semantics through two independent decoders
(ours and PIL), plus hand-worked packetisations of the restated stb algorithm."""
import os

import numpy as np
import pytest

PIL_Image = pytest.importorskip("PIL.Image")


def images():
    rng = np.random.default_rng(5)
    out = {"random": rng.integers(0, 256, (37, 53, 4), dtype=np.uint8)}
    runs = np.zeros((9, 300, 4), np.uint8)
    runs[:, :, 3] = 255
    runs[1, :, 0] = 7                                   # one 300-texel run: 128 + 128 + 44
    runs[2, ::2, 1] = 9                                 # alternating texels
    runs[3, 100:229, 2] = 200                           # run of 129 inside a row
    runs[4] = rng.integers(0, 2, (300, 1), dtype=np.uint8) * np.array([255, 128, 1, 77], np.uint8)  # short runs
    runs[5, -1] = 1                                     # change in the last texel
    runs[6, -2:] = 3                                    # run of two at the end of a row
    out["runs"] = runs
    out["1x1"] = rng.integers(0, 256, (1, 1, 4), dtype=np.uint8)
    out["1xN"] = np.repeat(rng.integers(0, 256, (1, 1, 4), dtype=np.uint8), 200, axis=1)
    out["Nx1"] = rng.integers(0, 3, (130, 1, 4), dtype=np.uint8)
    return out


@pytest.mark.parametrize("name", sorted(images()))
def test_tga_round_trip_and_pil(nv, tmp_path, name):
    img = images()[name]
    h, w = img.shape[:2]
    path = str(tmp_path / f"{name}.tga")
    nv.write_tga(path, img, w, h)
    raw = open(path, "rb").read()
    # header of stbi_write_tga(..., comp = 4) with RLE: "111 221 2222 11"
    assert list(raw[:18]) == [0, 0, 10, 0, 0, 0, 0, 0, 0, 0, 0, 0, w & 255, w >> 8, h & 255, h >> 8, 32, 8]
    back, bw, bh = nv.read_image(path)
    assert (bw, bh) == (w, h) and (back == img).all()
    if len(raw) >= 26:  # PIL looks for a TGA 2.0 footer 26 bytes before the end and trips over smaller files
        pil = np.asarray(PIL_Image.open(path).convert("RGBA"))
        assert pil.shape == img.shape and (pil == img).all()
    assert len(raw) <= 18 + w * h * 4 + (w * h + 127) // 128 + h  # never much worse than raw


def test_packets_follow_the_restated_stb_algorithm(nv, tmp_path):
    """Hand-worked rows (see stbi_write_tga_core: runs first, literal packets stop one texel before px[k-2] == px[k])."""
    A, B, Cc = [1, 2, 3, 4], [5, 6, 7, 8], [9, 10, 11, 12]
    bgra = lambda t: [t[2], t[1], t[0], t[3]]
    cases = [
        ([A, A, A, B, Cc, Cc], [0x82] + bgra(A) + [0x02] + bgra(B) + bgra(Cc) + bgra(Cc)),
        ([A, B, A, B], [0x00] + bgra(A) + [0x00] + bgra(B) + [0x01] + bgra(A) + bgra(B)),
        ([A], [0x00] + bgra(A)),
        ([A] * 130, [0xFF] + bgra(A) + [0x81] + bgra(A)),
    ]
    for k, (row, want) in enumerate(cases):
        path = str(tmp_path / f"p{k}.tga")
        nv.write_tga(path, np.array(row, np.uint8), len(row), 1)
        assert list(open(path, "rb").read()[18:]) == want, k


def test_rows_are_stored_bottom_up(nv, tmp_path):
    img = np.zeros((2, 1, 4), np.uint8)
    img[0, 0], img[1, 0] = [10, 20, 30, 40], [50, 60, 70, 80]
    path = str(tmp_path / "rows.tga")
    nv.write_tga(path, img, 1, 2)
    assert list(open(path, "rb").read()[18:]) == [0, 70, 60, 50, 80, 0, 30, 20, 10, 40]


def test_level_filenames(nv):
    """mipmap_storage.hpp:447-460: level 0 keeps the base name, level n goes before the last dot."""
    assert nv.level_filename("out.tga", 0) == "out.tga"
    assert nv.level_filename("out.tga", 3) == "out.3.tga"
    assert nv.level_filename("a.b.tga", 12) == "a.b.12.tga"
    assert nv.level_filename("noext", 2) == "noext2"
    assert nv.level_filename("./vk_compute_mipmaps_minimal.tga", 1) == "./vk_compute_mipmaps_minimal.1.tga"


def test_write_mipmaps_tga(nv, tmp_path):
    w, h = 20, 12
    rng = np.random.default_rng(1)
    chain = rng.integers(0, 256, nv.chain_bytes(w, h), dtype=np.uint8)
    base = str(tmp_path / "mips.tga")
    nv.write_mipmaps_tga(chain, w, h, base)
    views = nv.level_views(chain, w, h)
    assert len(views) == 5
    for level, v in enumerate(views):
        back, bw, bh = nv.read_image(nv.level_filename(base, level))
        assert (bh, bw) == v.shape[:2] and (back == v).all()
    assert not os.path.exists(nv.level_filename(base, 5))
    with pytest.raises(nv.NvpyrError):
        nv.write_mipmaps_tga(chain, w, h, str(tmp_path / "no_such_dir" / "x.tga"))


def test_reader_formats(nv, tmp_path):
    rng = np.random.default_rng(2)
    rgb = rng.integers(0, 256, (7, 5, 3), dtype=np.uint8)
    grey = rng.integers(0, 256, (4, 9), dtype=np.uint8)
    opaque = np.dstack([rgb, np.full((7, 5), 255, np.uint8)])
    # binary PPM / PGM, with a comment in the header
    (tmp_path / "a.ppm").write_bytes(b"P6\n# made by a test\n5 7\n255\n" + rgb.tobytes())
    (tmp_path / "a.pgm").write_bytes(b"P5 9 4 255\n" + grey.tobytes())
    got, w, h = nv.read_image(str(tmp_path / "a.ppm"))
    assert (w, h) == (5, 7) and (got == opaque).all()
    got, w, h = nv.read_image(str(tmp_path / "a.pgm"))
    assert (w, h) == (9, 4) and (got[..., 0] == grey).all() and (got[..., 1] == grey).all() and (got[..., 3] == 255).all()
    # TGA as PIL writes it: raw and RLE, 24 and 32 bits (PIL stores top-down or bottom-up as it likes)
    rgba = rng.integers(0, 256, (6, 11, 4), dtype=np.uint8)
    for mode, arr, want in (("RGB", rgb, opaque), ("RGBA", rgba, rgba)):
        for comp in (None, "tga_rle"):
            path = str(tmp_path / f"pil_{mode}_{comp}.tga")
            PIL_Image.fromarray(arr, mode).save(path, compression=comp)
            got, w, h = nv.read_image(path)
            assert (h, w) == want.shape[:2] and (got == want).all(), (mode, comp)
    # truncated and missing files are errors, not crashes
    data = open(str(tmp_path / "pil_RGBA_tga_rle.tga"), "rb").read()
    (tmp_path / "cut.tga").write_bytes(data[: len(data) // 2])
    for bad in ("cut.tga", "missing.tga"):
        with pytest.raises(nv.NvpyrError):
            nv.read_image(str(tmp_path / bad))
