"""ctypes access to the CPU checkers (oracle/).  TEST INFRASTRUCTURE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libnvpyr_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libnvpyr_ref.so")
EMU_SO = os.path.join(ORACLE_DIR, "_ref", "libnvpyr_glsl_emu.so")
P = C.c_void_p


class Step(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "pipeline", "input_level", "level_count", "src_w", "src_h", "workgroups", "push_constant", "bind",
        "barrier_after")]

    def key(self, with_src=False):
        k = (self.pipeline, self.input_level, self.level_count, self.workgroups, self.push_constant, self.bind,
             self.barrier_after)
        return k + ((self.src_w, self.src_h) if with_src else ())


class Cmp(C.Structure):
    _fields_ = [("worst", C.c_uint32), ("x", C.c_uint32), ("y", C.c_uint32), ("level", C.c_uint32),
                ("channel", C.c_uint32), ("mismatched", C.c_uint64), ("compared", C.c_uint64)]


def build_oracle():
    src = os.path.join(ORACLE_DIR, "nvpyr_oracle.c")
    if (not os.path.exists(ORACLE_SO)) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "oracle"])
    if os.path.isdir("/root/reference/nvpro_pyramid") and not (os.path.exists(REF_SO) and os.path.exists(EMU_SO)):
        subprocess.check_call(["make", "-s", "-C", ORACLE_DIR, "ref"])


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        lib.nvo_plan.argtypes = [C.c_uint32] * 6 + [C.POINTER(Step), C.c_uint32]
        lib.nvo_linear_from_srgb.restype = C.c_float
        lib.nvo_linear_from_srgb_formula.restype = C.c_float
        lib.nvo_srgb_from_linear.argtypes = [C.c_float]
        lib.nvo_srgb_from_linear_formula.argtypes = [C.c_float]
        lib.nvo_chain_texels.restype = C.c_uint64
        lib.nvo_level_offset.restype = C.c_uint64
        lib.nvo_shader_chain.argtypes = [C.c_int, P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32,
                                         C.c_uint32, C.POINTER(C.c_uint64)]
        lib.nvo_cpu_chain.argtypes = [C.c_int, P, C.c_uint32, C.c_uint32]
        lib.nvo_compare_srgba8.argtypes = [P, P, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(Cmp)]
        lib.nvo_premultiply_srgba8.argtypes = [P, P, C.c_uint64]
        lib.nvo_julia_srgba8.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int]

    def level_count(self, w, h):
        return self.lib.nvo_level_count(w, h)

    def chain_texels(self, w, h, levels=0):
        return self.lib.nvo_chain_texels(w, h, levels or self.level_count(w, h))

    def plan(self, w, h, levels=0, have_fast=1, div=4, max_levels=6):
        s = (Step * 40)()
        n = self.lib.nvo_plan(w, h, levels, have_fast, div, max_levels, s, 40)
        assert n >= 0
        return list(s[:n])

    def new_chain(self, level0, w, h, fmt=0, levels=0):
        dt = np.uint8 if fmt == 0 else np.float32
        buf = np.zeros(4 * self.chain_texels(w, h, levels), dtype=dt)
        buf[:4 * w * h] = np.asarray(level0, dtype=dt).reshape(-1)
        return buf

    def shader_chain(self, level0, w, h, fmt=0, levels=0, force_general=False, div=4, max_levels=6, f16_shared=False,
                     srgb_shared=False, general_blit=False):
        """Oracle A: chain in shader order. Returns (chain, stores).  f16_shared / srgb_shared: the F16_SHARED /
        SRGB_SHARED builds of the shaders.  general_blit: the demo's blit fallback instead of the general pipeline."""
        buf = self.new_chain(level0, w, h, fmt, levels)
        st = C.c_uint64()
        n = self.lib.nvo_shader_chain(fmt, buf.ctypes.data, w, h, levels,
                                      (1 if force_general else 0) | (2 if f16_shared else 0) | (4 if srgb_shared else 0)
                                      | (8 if general_blit else 0),
                                      div, max_levels, C.byref(st))
        assert n >= 0
        return buf, st.value

    def cpu_chain(self, level0, w, h, fmt=0):
        """Oracle B: the reference's own CPU generator restated."""
        buf = self.new_chain(level0, w, h, fmt)
        assert self.lib.nvo_cpu_chain(fmt, buf.ctypes.data, w, h) > 0
        return buf

    def compare(self, a, b, w, h, levels=0):
        c = Cmp()
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.uint8)
        self.lib.nvo_compare_srgba8(a.ctypes.data, b.ctypes.data, w, h, levels, C.byref(c))
        return c

    def premultiply(self, level0):
        a = np.ascontiguousarray(level0, dtype=np.uint8).reshape(-1)
        out = np.empty_like(a)
        self.lib.nvo_premultiply_srgba8(a.ctypes.data, out.ctypes.data, a.size // 4)
        return out

    def julia(self, w, h, alpha_normalized=2109710467, max_iterations=64):
        out = np.empty(4 * w * h, dtype=np.uint8)
        self.lib.nvo_julia_srgba8(out.ctypes.data, w, h, alpha_normalized, max_iterations)
        return out


class Ref:
    """The reference's own headers compiled in place (oracle/ref_harness.cpp)."""

    def __init__(self, lib):
        self.lib = lib
        lib.ref_record_dispatch.argtypes = [C.c_uint32] * 4 + [C.POINTER(Step), C.c_uint32]
        lib.ref_record_dispatch_variant.argtypes = [C.c_uint32] * 5 + [C.POINTER(Step), C.c_uint32]
        lib.ref_linear_from_srgb.restype = C.c_float
        lib.ref_srgb_from_linear.argtypes = [C.c_float]
        lib.ref_cpu_generate_srgba8.argtypes = [P, C.c_uint32, C.c_uint32]
        lib.ref_compare.argtypes = [P, P, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32)]
        lib.ref_layout.argtypes = [C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32),
                                   C.POINTER(C.c_uint32), C.c_uint32]
        lib.ref_storage_create.restype = P
        lib.ref_storage_create.argtypes = [C.c_uint32, C.c_uint32, P]
        lib.ref_storage_generate.argtypes = [P]
        lib.ref_storage_read.argtypes = [P, P]
        lib.ref_storage_destroy.argtypes = [P]
        lib.ref_storage_bytes.restype = C.c_uint64
        lib.ref_storage_bytes.argtypes = [P]

    def plan(self, w, h, levels=0, have_fast=1):
        s = (Step * 40)()
        n = self.lib.ref_record_dispatch(w, h, levels, have_fast, s, 40)
        assert n >= 0, n
        return list(s[:n])

    def plan_variant(self, w, h, levels, div, max_levels):
        s = (Step * 40)()
        n = self.lib.ref_record_dispatch_variant(w, h, levels, div, max_levels, s, 40)
        assert n >= 0, n
        return list(s[:n])

    def cpu_chain(self, chain, w, h):
        buf = np.array(chain, dtype=np.uint8, copy=True)
        self.lib.ref_cpu_generate_srgba8(buf.ctypes.data, w, h)
        return buf

    def compare(self, a, b, w, h):
        xyzc = (C.c_uint32 * 4)()
        a = np.ascontiguousarray(a, dtype=np.uint8)
        b = np.ascontiguousarray(b, dtype=np.uint8)
        return self.lib.ref_compare(a.ctypes.data, b.ctypes.data, w, h, xyzc), tuple(xyzc)

    def layout(self, w, h):
        off = (C.c_uint64 * 40)()
        ws = (C.c_uint32 * 40)()
        hs = (C.c_uint32 * 40)()
        n = self.lib.ref_layout(w, h, off, ws, hs, 40)
        return [(off[i], ws[i], hs[i]) for i in range(n)]


class GlslEmu:
    """The reference's GLSL shaders + dispatch header executed on the CPU (oracle/glsl_emu)."""

    def __init__(self, lib):
        self.lib = lib
        lib.emu_run_chain.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
        lib.emu_run_chain_ex.argtypes = [P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint64)]
        lib.emu_glsl_srgb_from_linear.argtypes = [C.c_float]

    def run_chain(self, chain, w, h, levels=0, have_fast=1, f16_shared=0, srgb_shared=0):
        """f16_shared / srgb_shared: run the F16_SHARED / SRGB_SHARED build of the shaders."""
        buf = np.array(chain, dtype=np.uint8, copy=True)
        st = C.c_uint64()
        n = self.lib.emu_run_chain_ex(buf.ctypes.data, w, h, levels, have_fast, 2 if srgb_shared else (1 if f16_shared else 0),
                                      C.byref(st))
        assert n >= 0, n
        return buf, n, st.value


def load_emu():
    build_oracle()
    if not os.path.exists(EMU_SO):
        return None
    return GlslEmu(C.CDLL(EMU_SO))


def load_oracle():
    build_oracle()
    return Oracle(C.CDLL(ORACLE_SO))


def load_ref():
    build_oracle()
    if not os.path.exists(REF_SO):
        return None
    return Ref(C.CDLL(REF_SO))


def random_level0(w, h, seed, opaque=False, fmt=0):
    rng = np.random.default_rng(seed)
    if fmt == 1:
        return rng.random(4 * w * h, dtype=np.float32)
    a = rng.integers(0, 256, 4 * w * h, dtype=np.uint8)
    if opaque:
        a[3::4] = 255
    return a


def smooth_level0(w, h, seed=0):
    """Low-entropy image (gradients + a little noise), alpha varying."""
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:h, 0:w]
    img = np.empty((h, w, 4), dtype=np.float64)
    img[..., 0] = 255.0 * x / max(1, w - 1)
    img[..., 1] = 255.0 * y / max(1, h - 1)
    img[..., 2] = 127.5 + 127.5 * np.sin(x / 17.0) * np.cos(y / 23.0)
    img[..., 3] = 255.0 * (0.5 + 0.5 * np.cos((x + y) / 31.0))
    img += rng.normal(0, 1.5, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8).reshape(-1)
