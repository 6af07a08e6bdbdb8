"""The C-ABI library loads and exports every symbol include/nvpyr.h declares; argument validation
that needs no GPU.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "nvpyr.h")).read()
    return sorted(set(re.findall(r"NVPYR_API\s+[\w\s\*]+?\b(nvpyr\w+)\s*\(", hdr)))


def test_header_declares_expected_entry_points():
    syms = declared_symbols()
    for s in ["nvpyrDispatch", "nvpyrDispatchEx", "nvpyrDispatchBatch", "nvpyrGetPlan", "nvpyrGetLevelCount",
              "nvpyrGetLevelOffsetTexels", "nvpyrGetChainBytes", "nvpyrGenerateHost", "nvpyrPremultiplyAlpha",
              "nvpyrImportExternalMemoryFd", "nvpyrGetErrorString", "nvpyrShutdown", "nvpyrWriteTga",
              "nvpyrWriteChainTga", "nvpyrGetLevelFilename", "nvpyrReadImage", "nvpyrFree"]:
        assert s in syms
    assert len(syms) >= 22


def test_library_exports_every_declared_symbol(nv):
    lib = C.CDLL(nv.pyramid._lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), s


def test_header_is_plain_c(tmp_path):
    """The header compiles as C99 (no torch/CUDA types in the signatures)."""
    import subprocess
    src = tmp_path / "t.c"
    src.write_text('#include "nvpyr.h"\nint main(void){ nvpyrExtent2D e = {4, 4}; (void)e; return 0; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-fsyntax-only",
                           "-I", os.path.join(ROOT, "include"), str(src)])


def test_struct_layout_matches_ctypes(nv, tmp_path):
    import subprocess
    src = tmp_path / "s.c"
    src.write_text('#include <stdio.h>\n#include "nvpyr.h"\nint main(void){ printf("%zu %zu %zu %zu\\n", '
                   'sizeof(nvpyrDispatchDesc), sizeof(nvpyrPlanStep), offsetof(nvpyrDispatchDesc, levels), '
                   'offsetof(nvpyrDispatchDesc, stream)); return 0; }\n')
    exe = tmp_path / "s"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).split()
    L = nv.pyramid._lib
    assert int(out[0]) == C.sizeof(L.DispatchDesc) and int(out[1]) == C.sizeof(L.PlanStep)
    assert int(out[2]) == L.DispatchDesc.levels.offset and int(out[3]) == L.DispatchDesc.stream.offset


def test_error_codes_without_gpu(nv):
    L = nv.pyramid._lib
    lib = L.lib
    E = L.Extent2D
    assert lib.nvpyrGetVersion() == 100
    assert lib.nvpyrGetLevelCount(E(0, 5)) == 0
    out64 = C.c_uint64()
    assert lib.nvpyrGetChainBytes(E(0, 0), 0, 0, C.byref(out64)) == L.ERROR_INVALID_VALUE
    assert lib.nvpyrGetChainBytes(E(4, 4), 9, 0, C.byref(out64)) == L.ERROR_INVALID_VALUE
    assert lib.nvpyrGetChainBytes(E(4, 4), 0, 7, C.byref(out64)) == L.ERROR_UNSUPPORTED
    assert lib.nvpyrGetChainBytes(E(4, 4), 0, 0, None) == L.ERROR_INVALID_VALUE
    # dispatch validation happens before any CUDA call
    assert lib.nvpyrDispatch(None, 0, E(64, 64), None) == L.ERROR_INVALID_VALUE
    assert lib.nvpyrDispatch(C.c_void_p(0x1000), 0, E(0, 64), None) == L.ERROR_INVALID_VALUE
    assert lib.nvpyrDispatch(C.c_void_p(0x1000), 99, E(64, 64), None) == L.ERROR_INVALID_VALUE
    assert lib.nvpyrDispatch(C.c_void_p(0x1004), 0, E(64, 64), None) == L.ERROR_INVALID_VALUE  # misaligned base
    d = L.DispatchDesc()
    assert lib.nvpyrDispatchEx(C.byref(d)) == L.ERROR_INVALID_VALUE  # structSize == 0
    d.structSize = C.sizeof(L.DispatchDesc)
    d.extent = E(8, 8)
    d.base = 0x1000
    d.format = 5
    assert lib.nvpyrDispatchEx(C.byref(d)) == L.ERROR_UNSUPPORTED
    d.format = L.FORMAT_RGBA32F
    d.flags = L.FLAG_PREMULTIPLY_ALPHA
    assert lib.nvpyrDispatchEx(C.byref(d)) == L.ERROR_UNSUPPORTED
    d.flags = 1 << 9
    assert lib.nvpyrDispatchEx(C.byref(d)) == L.ERROR_UNSUPPORTED
    assert lib.nvpyrDispatchEx(None) == L.ERROR_INVALID_VALUE
    assert lib.nvpyrDispatchBatch(None, 3) == L.ERROR_INVALID_VALUE
    assert lib.nvpyrDispatchBatch(None, 0) == L.SUCCESS
    assert lib.nvpyrPremultiplyAlpha(None, None, 4, None) == L.ERROR_INVALID_VALUE
    assert lib.nvpyrGenerateHost(None, None, E(4, 4), 0, 0, 0) == L.ERROR_INVALID_VALUE
    h, p = C.c_void_p(), C.c_void_p()
    assert lib.nvpyrImportExternalMemoryFd(-1, 16, 0, 16, C.byref(h), C.byref(p)) == L.ERROR_INVALID_VALUE
    assert lib.nvpyrImportExternalMemoryFd(3, 16, 8, 16, C.byref(h), C.byref(p)) == L.ERROR_INVALID_VALUE
    assert lib.nvpyrReleaseExternalMemory(None) == L.ERROR_INVALID_VALUE
    assert lib.nvpyrGetErrorString(L.ERROR_UNSUPPORTED) == b"NVPYR_ERROR_UNSUPPORTED"
    steps = (L.PlanStep * 4)()
    n = C.c_uint32()
    assert lib.nvpyrGetPlan(E(4095, 4095), 0, None, steps, 4, C.byref(n)) == L.ERROR_INVALID_VALUE  # maxSteps too small
    opt = L.PlanOptions(0, 6, 6)
    assert lib.nvpyrGetPlan(E(64, 64), 0, C.byref(opt), steps, 4, C.byref(n)) == L.ERROR_UNSUPPORTED  # odd divisibility


def test_product_package_does_not_touch_oracle():
    """The shipped package and the CUDA sources never reference oracle/ (no CPU fallback)."""
    pkg = os.path.join(ROOT, "vk_compute_mipmaps_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".hpp", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle/" not in txt.replace("oracle/ ", "") or f in ("_lib.py", "__init__.py"), f
                assert "nvpyr_oracle" not in txt and "libnvpyr_ref" not in txt, f


def test_cxx_template_header_compiles_for_sm100a(tmp_path):
    """include/nvpyr.cuh (user-defined functor sets, SURVEY 8b "Extensibility") is header-only: a user translation
    unit with its own functor set must compile for sm_100a with nvcc alone."""
    import shutil, subprocess
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    src = tmp_path / "user.cu"
    src.write_text("""
#include "nvpyr.cuh"
struct LumaMin : nvpyr::PyramidFunctors<LumaMin>
{
  using Value = float2;
  static constexpr int kTexelBytes = 8;
  __device__ static Value load(const Params*, const void* t) { return *static_cast<const float2*>(t); }
  __device__ static void store(const Params*, void* t, Value v) { *static_cast<float2*>(t) = v; }
  __device__ static Value reduce(float a0, Value v0, float a1, Value v1, float a2, Value v2)
  { return make_float2(a0 * v0.x + a1 * v1.x + a2 * v2.x, fminf(v0.y, fminf(v1.y, v2.y))); }
};
nvpyrStatus run(const nvpyrDispatchDesc& d) { return nvpyr::dispatch<LumaMin>(d); }
""")
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-c", "-I",
                        os.path.join(ROOT, "include"), str(src), "-o", str(tmp_path / "user.o")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr


def test_generate_host_rejects_bad_out_buffers(nv):
    """generate_host(out=...) must not hand an undersized / mistyped / strided buffer to the C ABI (the download
    would overflow it).  Checked before the library is called: no GPU needed."""
    import numpy as np
    l0 = np.zeros((8, 8, 4), dtype=np.uint8)
    n = nv.chain_bytes(8, 8)
    for bad in (np.zeros(n - 1, dtype=np.uint8), np.zeros(n, dtype=np.float32), np.zeros(2 * n, dtype=np.uint8)[::2],
                bytearray(n)):
        with pytest.raises(ValueError):
            nv.generate_host(l0, 8, 8, out=bad)
    ro = np.zeros(n, dtype=np.uint8)
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        nv.generate_host(l0, 8, 8, out=ro)
