"""Host-side mirror of the reference's nvpro_pyramid interface, over the C ABI.

Reference interface (nvpro_pyramid/nvpro_pyramid_dispatch.hpp):

    struct NvproPyramidPipelines { generalPipeline, fastPipeline, layout, pushConstantOffset };  (:29-35)
    nvproCmdPyramidDispatch(cmdBuf, pipelines, baseWidth, baseHeight, mipLevels = 0);            (:54-59)

Here a "pipeline" is a shipped functor instance compiled into libnvpyr.so, the
command buffer is a CUDA stream, and the image is a packed linear chain in
device memory (layout of include/mipmap_storage.hpp:53-76).
"""
import ctypes as C
import functools
import os
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import (FLAG_F16_SHARED, FLAG_SRGB_SHARED, FLAG_GENERAL_BLIT, FLAG_FORCE_GENERAL, FLAG_NONE, FLAG_PREMULTIPLY_ALPHA, FORMAT_RGBA32F, FORMAT_SRGBA8, Dispatcher,
                   DispatchDesc, Extent2D, NvpyrError, PlanOptions, PlanStep, check, lib)

__all__ = [
    "PyramidPipelines", "cmd_pyramid_dispatch", "dispatch_batch", "level_count", "level_extent", "level_offset_texels",
    "chain_bytes", "chain_texels", "get_plan", "generate_host", "premultiply_alpha", "level_views", "launch_count", "init",
    "write_tga", "level_filename", "write_mipmaps_tga", "read_image",
    "FORMAT_SRGBA8", "FORMAT_RGBA32F", "FLAG_NONE", "FLAG_FORCE_GENERAL", "FLAG_PREMULTIPLY_ALPHA", "FLAG_F16_SHARED", "FLAG_SRGB_SHARED", "FLAG_GENERAL_BLIT", "NvpyrError",
]


@dataclass(frozen=True)
class PyramidPipelines:
    """Analogue of NvproPyramidPipelines (dispatch.hpp:29-35).

    general_pipeline is mandatory in the reference; fast_pipeline may be
    VK_NULL_HANDLE (here: False), which forces the general pipeline for every
    level (minimal_app -force-no-fast-pipeline).
    """
    format: int = FORMAT_SRGBA8
    fast_pipeline: bool = True
    fast_divisibility: int = 0  # template arg of nvproPyramidDefaultFastDispatcher, 0 = 4
    fast_max_levels: int = 0    # 0 = 6


def _texel_bytes(fmt):
    return 4 if fmt == FORMAT_SRGBA8 else 16


def level_count(width, height):
    return lib.nvpyrGetLevelCount(Extent2D(width, height))


def level_extent(width, height, level):
    out = Extent2D()
    check(lib.nvpyrGetLevelExtent(Extent2D(width, height), level, C.byref(out)), "nvpyrGetLevelExtent")
    return out.width, out.height


def level_offset_texels(width, height, level):
    out = C.c_uint64()
    check(lib.nvpyrGetLevelOffsetTexels(Extent2D(width, height), level, C.byref(out)), "nvpyrGetLevelOffsetTexels")
    return out.value


@functools.lru_cache(maxsize=4096)
def chain_bytes(width, height, levels=0, fmt=FORMAT_SRGBA8):
    out = C.c_uint64()
    check(lib.nvpyrGetChainBytes(Extent2D(width, height), levels, fmt, C.byref(out)), "nvpyrGetChainBytes")
    return out.value


def chain_texels(width, height, levels=0):
    return chain_bytes(width, height, levels, FORMAT_SRGBA8) // 4


def get_plan(width, height, levels=0, flags=0, fast_divisibility=0, fast_max_levels=0):
    """The reference-equivalent dispatch sequence as a list of dicts."""
    steps = (PlanStep * _lib.NVPYR_MAX_STEPS)()
    n = C.c_uint32()
    opt = PlanOptions(flags, fast_divisibility, fast_max_levels)
    check(lib.nvpyrGetPlan(Extent2D(width, height), levels, C.byref(opt), steps, _lib.NVPYR_MAX_STEPS, C.byref(n)),
          "nvpyrGetPlan")
    return [{f: getattr(s, f) for f, _ in PlanStep._fields_} for s in steps[:n.value]]


def _device_ptr(image):
    if isinstance(image, int):
        return image
    if hasattr(image, "data_ptr"):  # torch tensor
        if not image.is_cuda:
            raise ValueError("image tensor must live in CUDA memory (use generate_host for host buffers)")
        if not image.is_contiguous():
            raise ValueError("image tensor must be contiguous")
        return image.data_ptr()
    raise TypeError("image must be a CUDA tensor or an integer device pointer")


def _stream_ptr(stream):
    if stream is None:
        import torch
        return torch.cuda.current_stream().cuda_stream
    if isinstance(stream, int):
        return stream
    return stream.cuda_stream


def _make_desc(image, pipelines, base_width, base_height, mip_levels, flags, stream, level_ptrs=None, pitches=None):
    d = DispatchDesc()
    d.structSize = C.sizeof(DispatchDesc)
    d.format = pipelines.format
    d.flags = flags | (0 if pipelines.fast_pipeline else FLAG_FORCE_GENERAL)
    d.extent = Extent2D(base_width, base_height)
    d.levelCount = mip_levels
    d.base = None if image is None else _device_ptr(image)
    if level_ptrs is not None:
        for i, p in enumerate(level_ptrs):
            d.levels[i] = p
            d.rowPitchBytes[i] = 0 if pitches is None else pitches[i]
    d.fastDivisibility = pipelines.fast_divisibility
    d.fastMaxLevels = pipelines.fast_max_levels
    d.stream = _stream_ptr(stream)
    return d


def cmd_pyramid_dispatch(stream, pipelines, base_width, base_height, mip_levels=0, general_dispatcher=None,
                         fast_dispatcher=None, *, image, flags=FLAG_NONE, level_ptrs=None, pitches=None):
    """nvproCmdPyramidDispatch(cmdBuf, pipelines, baseWidth, baseHeight, mipLevels[, generalDispatcher,
    fastDispatcher]) -- both overloads (dispatch.hpp:54-59 and :109-116).

    Enqueues on ``stream`` (None = torch's current stream) the generation of mip
    levels 1..mip_levels-1 of ``image`` (packed chain, level 0 filled) from level 0.
    general_dispatcher / fast_dispatcher: Python callables ``f(state, step) -> levels filled`` with the contract of
    nvpro_pyramid_dispatcher_t (state: currentLevel, remainingLevels, currentX, currentY; fast may return 0).
    """
    if image is not None and hasattr(image, "numel"):
        need = chain_bytes(base_width, base_height, mip_levels, pipelines.format)
        have = image.numel() * image.element_size()
        if have < need:
            raise ValueError(f"image buffer holds {have} bytes, the chain needs {need}")
    d = _make_desc(image, pipelines, base_width, base_height, mip_levels, flags, stream, level_ptrs, pitches)
    if general_dispatcher is None and fast_dispatcher is None:
        check(lib.nvpyrDispatchEx(C.byref(d)), "nvpyrDispatchEx")
        return

    def wrap(f):
        if f is None:
            return Dispatcher()  # NULL: the default
        return Dispatcher(lambda state, step, _user: int(f(state.contents, step.contents)))
    g, f = wrap(general_dispatcher), wrap(fast_dispatcher)  # kept alive until the call returns
    check(lib.nvpyrDispatchWithDispatchers(C.byref(d), g, f, None), "nvpyrDispatchWithDispatchers")


def dispatch_batch(stream, pipelines, images, base_width, base_height, mip_levels=0, flags=FLAG_NONE):
    """nvpyrDispatchBatch over independent images of one size."""
    descs = (DispatchDesc * len(images))()
    for i, img in enumerate(images):
        descs[i] = _make_desc(img, pipelines, base_width, base_height, mip_levels, flags, stream)
    check(lib.nvpyrDispatchBatch(descs, len(images)), "nvpyrDispatchBatch")


def premultiply_alpha(stream, src, dst, texels):
    check(lib.nvpyrPremultiplyAlpha(_device_ptr(src), _device_ptr(dst), texels, _stream_ptr(stream)),
          "nvpyrPremultiplyAlpha")


def generate_host(level0, width, height, mip_levels=0, fmt=FORMAT_SRGBA8, flags=FLAG_NONE, out=None):
    """minimal_app's round trip with host buffers: upload, generate, download.

    level0: numpy array (uint8 HxWx4 or float32 HxWx4, C-contiguous, may be pinned).
    Returns the packed chain as a flat numpy array (level 0 included).
    """
    dt = np.uint8 if fmt == FORMAT_SRGBA8 else np.float32
    level0 = np.ascontiguousarray(level0, dtype=dt)
    if level0.size != width * height * 4:
        raise ValueError("level0 has the wrong number of elements")
    n = chain_bytes(width, height, mip_levels, fmt) // np.dtype(dt).itemsize
    if out is None:
        out = np.empty(n, dtype=dt)
    elif not (isinstance(out, np.ndarray) and out.dtype == dt and out.flags.c_contiguous and out.flags.writeable
              and out.size >= n):
        raise ValueError(f"out must be a writable C-contiguous {np.dtype(dt).name} array of at least {n} elements "
                         "(the packed chain, level 0 included)")
    check(lib.nvpyrGenerateHost(level0.ctypes.data, out.ctypes.data, Extent2D(width, height), mip_levels, fmt, flags),
          "nvpyrGenerateHost")
    return out


def write_tga(filename, rgba8, width, height):
    """stbi_write_tga(filename, w, h, 4, data) as the reference calls it (mipmap_storage.hpp:462)."""
    a = np.ascontiguousarray(rgba8, dtype=np.uint8)
    if a.size != 4 * width * height:
        raise ValueError("rgba8 has the wrong number of elements")
    check(lib.nvpyrWriteTga(os.fsencode(filename), a.ctypes.data, Extent2D(width, height)), "nvpyrWriteTga")


def level_filename(base_filename, level):
    buf = C.create_string_buffer(4096)
    check(lib.nvpyrGetLevelFilename(os.fsencode(base_filename), level, buf, len(buf)), "nvpyrGetLevelFilename")
    return os.fsdecode(buf.value)


def write_mipmaps_tga(chain, width, height, base_filename, mip_levels=0):
    """writeMipmapsTga (mipmap_storage.hpp:441-479) for a packed sRGBA8 host chain."""
    a = np.ascontiguousarray(chain, dtype=np.uint8)
    if a.size != chain_bytes(width, height, mip_levels):
        raise ValueError("chain has the wrong size")
    check(lib.nvpyrWriteChainTga(a.ctypes.data, Extent2D(width, height), mip_levels, os.fsencode(base_filename)),
          "nvpyrWriteChainTga")


def read_image(filename):
    """Stands where the reference calls stbi_load(..., 4): returns (HxWx4 uint8 array, width, height)."""
    ptr, ext = C.c_void_p(), Extent2D()
    check(lib.nvpyrReadImage(os.fsencode(filename), C.byref(ptr), C.byref(ext)), "nvpyrReadImage")
    try:
        n = 4 * ext.width * ext.height
        a = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n,)).copy()
    finally:
        lib.nvpyrFree(ptr)
    return a.reshape(ext.height, ext.width, 4), ext.width, ext.height


def level_views(chain, width, height, mip_levels=0):
    """Split a flat packed chain (numpy or torch, 4 scalars per texel) into per-level [H, W, 4] views."""
    n = mip_levels or level_count(width, height)
    views, off = [], 0
    for i in range(n):
        w, h = max(1, width >> i), max(1, height >> i)
        views.append(chain[4 * off:4 * (off + w * h)].reshape(h, w, 4))
        off += w * h
    return views


def init():
    """nvpyrInit: create the per-device state for the current device now (required before a CUDA-graph capture
    if no dispatch has run on the device yet)."""
    check(lib.nvpyrInit(), "nvpyrInit")


def launch_count():
    return lib.nvpyrGetLaunchCount()
