// tmu_probe.cu -- design probe, not product code.  Answers on a real B200:
//  (1) what float does the texture unit return for each sRGB8 code (cudaTextureDesc::sRGB)?
//  (2) is a bilinear fetch at the 2x2 centre a deterministic function of those values, and which?
//  (3) how fast is "1 bilinear fetch per 2x2 quad" over 16384^2 against a plain streaming kernel?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmu_probe tools/tmu_probe.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../vk_compute_mipmaps_b200/csrc/srgb_tables.inc"

#define CK(x) do { cudaError_t e = (x); if(e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while(0)

__global__ void pointFetch(cudaTextureObject_t t, float4* out, int w, int h)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if(x < w && y < h) out[y * w + x] = tex2D<float4>(t, x + 0.5f, y + 0.5f);
}
__global__ void quadFetch(cudaTextureObject_t t, float4* out, int w2, int h2)
{
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if(x < w2 && y < h2) out[y * w2 + x] = tex2D<float4>(t, 2 * x + 1.0f, 2 * y + 1.0f);
}
// throughput: one bilinear fetch per quad, cheap pack, 4-byte store
__global__ void __launch_bounds__(256) quadStream(cudaTextureObject_t t, uint32_t* out, int w2, int h2)
{
  int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if(x < w2 && y < h2)
  {
    float4 v = tex2D<float4>(t, 2 * x + 1.0f, 2 * y + 1.0f);
    uint32_t p = (uint32_t)(v.x * 255.f) | (uint32_t)(v.y * 255.f) << 8 | (uint32_t)(v.z * 255.f) << 16 | (uint32_t)(v.w * 255.f) << 24;
    out[(size_t)y * w2 + x] = p;
  }
}
// streaming floor: read level 0 with 16-byte loads, write 1/4 of it
__global__ void __launch_bounds__(256) streamProbe(const uint4* in, uint32_t* out, size_t n4)
{
  for(size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
  {
    uint4 v = in[i];
    out[i] = v.x ^ v.y ^ v.z ^ v.w;
  }
}

static float b2f(uint32_t b) { float f; memcpy(&f, &b, 4); return f; }
static uint32_t f2b(float f) { uint32_t b; memcpy(&b, &f, 4); return b; }

int main()
{
  setvbuf(stdout, NULL, _IONBF, 0);
  // ---------- (1) decode table ----------
  const int w = 256, h = 4;
  uint8_t* himg = (uint8_t*)malloc(w * h * 4);
  for(int y = 0; y < h; ++y) for(int x = 0; x < w; ++x) { uint8_t* p = himg + 4 * (y * w + x); p[0] = x; p[1] = 255 - x; p[2] = (x * 7) & 255; p[3] = x; }
  uint8_t* dimg; size_t pitch;
  CK(cudaMallocPitch(&dimg, &pitch, w * 4, h));
  CK(cudaMemcpy2D(dimg, pitch, himg, w * 4, w * 4, h, cudaMemcpyHostToDevice));
  cudaResourceDesc rd; memset(&rd, 0, sizeof rd);
  rd.resType = cudaResourceTypePitch2D; rd.res.pitch2D.devPtr = dimg; rd.res.pitch2D.desc = cudaCreateChannelDesc<uchar4>();
  rd.res.pitch2D.width = w; rd.res.pitch2D.height = h; rd.res.pitch2D.pitchInBytes = pitch;
  cudaTextureDesc td; memset(&td, 0, sizeof td);
  td.addressMode[0] = td.addressMode[1] = cudaAddressModeClamp; td.filterMode = cudaFilterModePoint;
  td.readMode = cudaReadModeNormalizedFloat; td.sRGB = 1; td.normalizedCoords = 0;
  cudaTextureObject_t tp; CK(cudaCreateTextureObject(&tp, &rd, &td, nullptr));
  float4* dout; CK(cudaMalloc(&dout, w * h * sizeof(float4)));
  pointFetch<<<dim3(1, h), 256>>>(tp, dout, w, h); CK(cudaDeviceSynchronize());
  float4* hout = (float4*)malloc(w * h * sizeof(float4)); CK(cudaMemcpy(hout, dout, w * h * sizeof(float4), cudaMemcpyDeviceToHost));
  float hwlut[256]; int equal = 0; double maxrel = 0; int maxulp = 0;
  for(int c = 0; c < 256; ++c)
  {
    hwlut[c] = hout[c].x; float ref = b2f(NVPYR_SRGB_DECODE_BITS[c]);
    if(f2b(hwlut[c]) == NVPYR_SRGB_DECODE_BITS[c]) ++equal;
    int ulp = abs((int)f2b(hwlut[c]) - (int)NVPYR_SRGB_DECODE_BITS[c]); if(ulp > maxulp) maxulp = ulp;
    if(ref > 0) { double r = fabs((double)hwlut[c] - ref) / ref; if(r > maxrel) maxrel = r; }
  }
  printf("(1) HW sRGB decode vs pinned table: %d/256 bit-equal, max ulp %d, max rel %.3e\n", equal, maxulp, maxrel);
  printf("    samples: c=1 hw=%.9g (%08x) ref=%.9g | c=128 hw=%.9g (%08x) ref=%.9g | c=255 hw=%.9g | alpha(77)=%.9g (77/255=%.9g)\n",
         hwlut[1], f2b(hwlut[1]), b2f(NVPYR_SRGB_DECODE_BITS[1]), hwlut[128], f2b(hwlut[128]), b2f(NVPYR_SRGB_DECODE_BITS[128]), hwlut[255], hout[77].w, 77.f / 255.f);
  // how many mantissa bits does the HW value use?
  int maxbits = 0; for(int c = 1; c < 256; ++c) { uint32_t m = f2b(hwlut[c]) & 0x7FFFFF; int tz = m ? __builtin_ctz(m) : 23; if(23 - tz > maxbits) maxbits = 23 - tz; }
  printf("    HW decode values use at most %d mantissa bits; as k/2^n? ", maxbits);
  for(int n = 8; n <= 24; ++n) { int ok = 1; for(int c = 0; c < 256 && ok; ++c) { double v = (double)hwlut[c] * (double)(1u << n); if(v != floor(v)) ok = 0; } if(ok) { printf("multiples of 2^-%d", n); break; } }
  printf("\n");
  // G channel decode must equal table of (255-x): consistency
  int chanok = 1; for(int c = 0; c < 256; ++c) if(hout[c].y != hwlut[255 - c]) chanok = 0;
  printf("    channels consistent: %d\n", chanok);

  // ---------- (2) bilinear at quad centres ----------
  const int W = 1024, Hh = 1024;
  uint8_t* big = (uint8_t*)malloc((size_t)W * Hh * 4); srand(1);
  for(size_t i = 0; i < (size_t)W * Hh * 4; ++i) big[i] = rand() & 255;
  // make half of the image smooth-ish
  for(int y = 0; y < Hh / 2; ++y) for(int x = 0; x < W; ++x) for(int c = 0; c < 4; ++c) big[4 * ((size_t)y * W + x) + c] = (uint8_t)((x / 4 + y / 3 + c * 40 + (rand() & 3)) & 255);
  uint8_t* dbig; size_t bp; CK(cudaMallocPitch(&dbig, &bp, W * 4, Hh)); CK(cudaMemcpy2D(dbig, bp, big, W * 4, W * 4, Hh, cudaMemcpyHostToDevice));
  rd.res.pitch2D.devPtr = dbig; rd.res.pitch2D.width = W; rd.res.pitch2D.height = Hh; rd.res.pitch2D.pitchInBytes = bp;
  td.filterMode = cudaFilterModeLinear;
  cudaTextureObject_t tl; CK(cudaCreateTextureObject(&tl, &rd, &td, nullptr));
  float4* dq; CK(cudaMalloc(&dq, (size_t)(W / 2) * (Hh / 2) * sizeof(float4)));
  quadFetch<<<dim3(W / 2 / 256, Hh / 2), 256>>>(tl, dq, W / 2, Hh / 2); CK(cudaDeviceSynchronize());
  float4* hq = (float4*)malloc((size_t)(W / 2) * (Hh / 2) * sizeof(float4)); CK(cudaMemcpy(hq, dq, (size_t)(W / 2) * (Hh / 2) * sizeof(float4), cudaMemcpyDeviceToHost));
  long n = 0, eqV = 0, eqH = 0, eqD = 0, eqPinned = 0; double maxd = 0, maxdPinned = 0;
  for(int y = 0; y < Hh / 2; ++y) for(int x = 0; x < W / 2; ++x) for(int c = 0; c < 3; ++c)
  {
    const uint8_t* p00 = big + 4 * ((size_t)(2 * y) * W + 2 * x); const uint8_t* p10 = p00 + 4; const uint8_t* p01 = p00 + 4 * W; const uint8_t* p11 = p01 + 4;
    float a = hwlut[p00[c]], b = hwlut[p10[c]], cc = hwlut[p01[c]], d = hwlut[p11[c]];
    float got = ((float*)&hq[(size_t)y * (W / 2) + x])[c];
    float v = 0.25f * ((a + cc) + (b + d)), hh = 0.25f * ((a + b) + (cc + d));
    double ex = 0.25 * ((double)a + b + cc + d);
    ++n; eqV += got == v; eqH += got == hh; eqD += got == (float)ex;
    double dd = fabs(got - ex); if(dd > maxd) maxd = dd;
    float pa = b2f(NVPYR_SRGB_DECODE_BITS[p00[c]]), pb = b2f(NVPYR_SRGB_DECODE_BITS[p10[c]]), pc = b2f(NVPYR_SRGB_DECODE_BITS[p01[c]]), pd = b2f(NVPYR_SRGB_DECODE_BITS[p11[c]]);
    float pin = 0.25f * ((pa + pc) + (pb + pd)); eqPinned += got == pin; double dp = fabs(got - pin); if(dp > maxdPinned) maxdPinned = dp;
  }
  printf("(2) bilinear @ quad centre, %ld rgb samples: == f32 vertical-pair avg of HW LUT %.4f%%, == horizontal-pair %.4f%%, == exact-sum rounded %.4f%%, max |diff| vs exact avg of HW LUT %.3e\n",
         n, 100.0 * eqV / n, 100.0 * eqH / n, 100.0 * eqD / n, maxd);
  printf("    vs pinned-table software path: bit-equal %.4f%%, max |diff| %.3e\n", 100.0 * eqPinned / n, maxdPinned);
  for(int i = 0; i < 4; ++i) { int x = 37 + i * 91, y = 300 + i * 13; float got = hq[(size_t)y * (W / 2) + x].x; const uint8_t* p00 = big + 4 * ((size_t)(2 * y) * W + 2 * x);
    double ex = 0.25 * ((double)hwlut[p00[0]] + hwlut[p00[4]] + hwlut[p00[4 * W]] + hwlut[p00[4 * W + 4]]); printf("    sample: got %.10g (%08x) exact-avg %.10g\n", got, f2b(got), ex); }

  // ---------- (3) throughput ----------
  const int BW = 16384, BH = 16384; uint8_t* l0; size_t lp = (size_t)BW * 4;
  CK(cudaMalloc(&l0, lp * BH)); CK(cudaMemset(l0, 0x5A, lp * BH));
  uint32_t* l1; CK(cudaMalloc(&l1, (size_t)(BW / 2) * (BH / 2) * 4));
  rd.res.pitch2D.devPtr = l0; rd.res.pitch2D.width = BW; rd.res.pitch2D.height = BH; rd.res.pitch2D.pitchInBytes = lp;
  cudaTextureObject_t tb; CK(cudaCreateTextureObject(&tb, &rd, &td, nullptr));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1); float ms;
  for(int rep = 0; rep < 2; ++rep)
  {
    cudaEventRecord(e0);
    for(int i = 0; i < 5; ++i) quadStream<<<dim3(BW / 2 / 32, BH / 2 / 8), 256>>>(tb, l1, BW / 2, BH / 2);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
  }
  double bytes = (double)BW * BH * 4 * 1.25;
  printf("(3) bilinear fetch per quad over 16384^2 (+L1 store): %.1f us  -> %.0f GB/s (L0 read + L1 write)\n", ms / 5 * 1e3, bytes / (ms / 5 * 1e-3) / 1e9);
  for(int rep = 0; rep < 2; ++rep)
  {
    cudaEventRecord(e0);
    for(int i = 0; i < 5; ++i) streamProbe<<<148 * 8, 256>>>((const uint4*)l0, l1, (size_t)BW * BH / 4);
    cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1);
  }
  printf("    plain streaming probe (LDG.128 level 0, write 1/4): %.1f us -> %.0f GB/s\n", ms / 5 * 1e3, bytes / (ms / 5 * 1e-3) / 1e9);
  return 0;
}
