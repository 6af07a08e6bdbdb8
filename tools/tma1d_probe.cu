// tools/tma1d_probe.cu -- does a ONE-dimensional tensor map accept a box that starts at any 4-byte element?
// (The general pipeline stages rows of NPOT levels -- pitch not a multiple of 16 bytes -- this way.)
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tma1d_probe tools/tma1d_probe.cu && tools/tma1d_probe
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <vector>

__global__ void probe(const __grid_constant__ CUtensorMap map, uint32_t start, uint32_t* out)
{
  __shared__ __align__(128) uint32_t buf[128];
  __shared__ unsigned long long bar;
  const uint32_t b = uint32_t(__cvta_generic_to_shared(&bar)), d = uint32_t(__cvta_generic_to_shared(buf));
  if(threadIdx.x == 0)
  {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(b) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], 512;" ::"r"(b) : "memory");
    asm volatile("cp.async.bulk.tensor.1d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2}], [%3];" ::"r"(d),
                 "l"(&map), "r"(start), "r"(b)
                 : "memory");
  }
  __syncthreads();
  asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(b) : "memory");
  out[threadIdx.x] = buf[threadIdx.x];
}

int main()
{
  using Fn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void*                           f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaFree(0);
  if(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || !f)
    return printf("no cuTensorMapEncodeTiled\n"), 1;
  const uint32_t        n = 100000;
  std::vector<uint32_t> h(n);
  for(uint32_t i = 0; i < n; ++i)
    h[i] = i * 2654435761u;
  uint32_t *dIn, *dOut;
  cudaMalloc(&dIn, n * 4), cudaMalloc(&dOut, 128 * 4);
  cudaMemcpy(dIn, h.data(), n * 4, cudaMemcpyHostToDevice);
  const cuuint64_t dims[1] = {n};
  const cuuint64_t dummyStrides[1] = {0};
  const cuuint32_t box[1] = {128}, estr[1] = {1};
  for(int variant = 0; variant < 2; ++variant)
  {
    CUtensorMap map;
    CUresult    r = reinterpret_cast<Fn>(f)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 1, dIn, dims, variant ? dummyStrides : nullptr, box, estr,
                                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode rank 1, strides %s -> CUresult %d\n", variant ? "dummy array" : "nullptr", int(r));
    if(r != CUDA_SUCCESS)
      continue;
    for(uint32_t start : {0u, 1u, 2u, 3u, 5u, 4095u, 16383u, n - 128u, n - 100u, n - 1u})
    {
      probe<<<1, 128>>>(map, start, dOut);
      uint32_t    got[128];
      cudaError_t e = cudaMemcpy(got, dOut, sizeof got, cudaMemcpyDeviceToHost);
      if(e != cudaSuccess)
        return printf("start %u: %s\n", start, cudaGetErrorString(e)), 1;
      int bad = 0, zeros = 0;
      for(uint32_t i = 0; i < 128; ++i)
      {
        const uint32_t want = start + i < n ? h[start + i] : 0u;
        bad += got[i] != want;
        zeros += start + i >= n;
      }
      printf("  start %6u: %s (%d out-of-range elements zero-filled)\n", start, bad ? "MISMATCH" : "ok", zeros);
    }
  }
  return 0;
}
