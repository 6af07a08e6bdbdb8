"""Hands libnvpyr a LIVE file descriptor: device memory created with the CUDA virtual-memory API
(cuMemCreate, requestedHandleTypes = POSIX file descriptor) is exported with cuMemExportToShareableHandle and the fd
goes through nvpyrImportExternalMemoryFd (cudaImportExternalMemory, opaque fd) -- the same call path a Vulkan
application takes with the fd of vkGetMemoryFdKHR (VK_KHR_external_memory_fd).  The GPU boxes have no Vulkan
implementation (profiles/r2_vulkan_probe.txt), so the exporting side is CUDA's own allocator instead of a VkDeviceMemory;
the importing side -- the library's half of the interop -- is the real one.  The chain generated in the imported
buffer is compared with the oracle bit for bit.
usage: python tools/extmem_probe.py [W H]"""
import ctypes as C
import os
import sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import vk_compute_mipmaps_b200 as nv
from vk_compute_mipmaps_b200._lib import lib
try:
    from cuda.bindings import driver as cu
except ImportError:
    from cuda import cuda as cu


def ck(res, what):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"{what}: {err}")
    return res[1] if len(res) == 2 else res[1:]


def run(w, h):
    torch.cuda.init()
    torch.zeros(1, device="cuda")  # primary context current
    prop = cu.CUmemAllocationProp()
    prop.type = cu.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
    prop.location.type = cu.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
    prop.location.id = torch.cuda.current_device()
    prop.requestedHandleTypes = cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR
    gran = ck(cu.cuMemGetAllocationGranularity(prop, cu.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM),
              "cuMemGetAllocationGranularity")
    need = nv.chain_bytes(w, h)
    size = (need + gran - 1) // gran * gran
    handle = ck(cu.cuMemCreate(size, prop, 0), "cuMemCreate")
    fd = int(ck(cu.cuMemExportToShareableHandle(handle, cu.CUmemAllocationHandleType.CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
                "cuMemExportToShareableHandle"))
    print(f"exported fd {fd} of a {size}-byte device allocation (chain needs {need})")
    ext, ptr = C.c_void_p(), C.c_void_p()
    st = lib.nvpyrImportExternalMemoryFd(fd, size, 0, need, C.byref(ext), C.byref(ptr))
    if st != 0:
        print(f"nvpyrImportExternalMemoryFd -> {lib.nvpyrGetErrorString(st).decode()} (cudaError {lib.nvpyrGetLastCudaError()})")
        ck(cu.cuMemRelease(handle), "cuMemRelease")
        return 2
    print(f"imported: device pointer {ptr.value:#x}")
    import _oracle
    o = _oracle.load_oracle()
    l0 = _oracle.random_level0(w, h, 77)
    ck(cu.cuMemcpyHtoD(ptr.value, l0.ctypes.data, l0.nbytes), "cuMemcpyHtoD")
    nv.cmd_pyramid_dispatch(None, nv.PyramidPipelines(), w, h, image=ptr.value)
    torch.cuda.synchronize()
    got = np.empty(need, dtype=np.uint8)
    ck(cu.cuMemcpyDtoH(got.ctypes.data, ptr.value, need), "cuMemcpyDtoH")
    want = o.shader_chain(l0, w, h)[0]
    same = bool((got == want).all())
    print(f"{w}x{h}: chain generated in the imported buffer {'==' if same else '!='} oracle ({need} bytes)")
    assert lib.nvpyrReleaseExternalMemory(ext) == 0
    ck(cu.cuMemRelease(handle), "cuMemRelease")
    return 0 if same else 1


if __name__ == "__main__":
    w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1920, 1080)
    sys.exit(run(w, h))
