// tools/bench_native.cpp -- the config table of tools/bench_configs.py measured from C++: the reference's benchmark
// protocol (demo_app/mipmaps_app.cpp:621-622,712-728: batches of 8 back-to-back generations between two timestamps,
// first batch discarded, min / median per generation) with nvpyrDispatch called the way a C++ application calls it.
// Python's ctypes wrapper costs 10-20 us per call, which is more than the device needs for a chain of <= 2048^2: the
// Python table is host-bound there, this one is not.
// build: g++ -O2 -std=c++17 -I include -I /usr/local/cuda/include tools/bench_native.cpp -o tools/bench_native
//            -L vk_compute_mipmaps_b200 -lnvpyr -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/vk_compute_mipmaps_b200
// usage: tools/bench_native [--batches 30] [--json out.json] [--peak GBps] [--only substr] [--opaque]
//        --opaque: alpha = 255 in every sRGBA8 texel (an image that came from a JPEG: the fast kernel's opaque path)
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <chrono>
#include <vector>

#include "nvpyr.h"

struct Config
{
  const char* name;
  uint32_t    w, h;
  nvpyrFormat fmt;
};
static const Config kConfigs[] = {
    {"4096.jpg class", 4096, 4096, NVPYR_FORMAT_SRGBA8},   {"4095.jpg class", 4095, 4095, NVPYR_FORMAT_SRGBA8},
    {"4094.jpg class", 4094, 4094, NVPYR_FORMAT_SRGBA8},   {"lunch_2047 class", 2047, 2047, NVPYR_FORMAT_SRGBA8},
    {"2048 class", 2048, 2048, NVPYR_FORMAT_SRGBA8},       {"1080p class", 1920, 1080, NVPYR_FORMAT_SRGBA8},
    {"1440p class", 2560, 1440, NVPYR_FORMAT_SRGBA8},      {"4k class", 3840, 2160, NVPYR_FORMAT_SRGBA8},
    {"tall class", 1080, 4096, NVPYR_FORMAT_SRGBA8},       {"alpha2052 class", 2052, 2052, NVPYR_FORMAT_SRGBA8},
    {"mandelbrots class", 3095, 990, NVPYR_FORMAT_SRGBA8}, {"16384 synthetic", 16384, 16384, NVPYR_FORMAT_SRGBA8},
    {"8192 synthetic", 8192, 8192, NVPYR_FORMAT_SRGBA8},   {"4096 rgba32f", 4096, 4096, NVPYR_FORMAT_RGBA32F},
    {"4095 rgba32f", 4095, 4095, NVPYR_FORMAT_RGBA32F},
};

#define CK(x)                                                                                                     \
  do                                                                                                              \
  {                                                                                                               \
    cudaError_t e_ = (x);                                                                                         \
    if(e_ != cudaSuccess)                                                                                         \
    {                                                                                                             \
      fprintf(stderr, "%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e_));                                   \
      exit(2);                                                                                                    \
    }                                                                                                             \
  } while(0)

int main(int argc, char** argv)
{
  int         batches = 30;
  double      peak    = 6446.9;
  std::string json, only;
  bool        opaque = false;
  for(int i = 1; i < argc; ++i)
  {
    if(!strcmp(argv[i], "--batches") && i + 1 < argc)
      batches = atoi(argv[++i]);
    else if(!strcmp(argv[i], "--json") && i + 1 < argc)
      json = argv[++i];
    else if(!strcmp(argv[i], "--peak") && i + 1 < argc)
      peak = atof(argv[++i]);
    else if(!strcmp(argv[i], "--only") && i + 1 < argc)
      only = argv[++i];
    else if(!strcmp(argv[i], "--opaque"))
      opaque = true;
  }
  if(nvpyrInit() != NVPYR_SUCCESS)
  {
    fprintf(stderr, "nvpyrInit failed (no sm_100 device?)\n");
    return 2;
  }
  cudaStream_t stream;
  CK(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  FILE* jf = json.empty() ? nullptr : fopen(json.c_str(), "w");
  if(jf)
    fprintf(jf, "[\n");
  bool first = true;
  for(const Config& c : kConfigs)
  {
    if(!only.empty() && !strstr(c.name, only.c_str()))
      continue;
    uint64_t bytes = 0;
    nvpyrGetChainBytes({c.w, c.h}, 0, c.fmt, &bytes);
    const int nrot = int(std::max<uint64_t>(2, std::min<uint64_t>(8, 400000000ull / bytes + 1)));  // rotate: small chains must not stay in L2
    const uint64_t       l0 = uint64_t(c.w) * c.h * (c.fmt == NVPYR_FORMAT_SRGBA8 ? 4 : 16);
    std::vector<uint8_t> host(l0);
    uint32_t             s = 12345u + c.w;
    if(c.fmt == NVPYR_FORMAT_SRGBA8)
      for(uint64_t i = 0; i < l0; i += 4)
      {
        s = s * 1664525u + 1013904223u;
        const uint32_t v = (s ^ (s >> 13)) | (opaque ? 0xFF000000u : 0u);
        memcpy(&host[i], &v, 4);
      }
    else
      for(uint64_t i = 0; i < l0; i += 4)
      {
        s = s * 1664525u + 1013904223u;
        const float v = float(s >> 8) * (1.0f / 16777216.0f);
        memcpy(&host[i], &v, 4);
      }
    std::vector<void*> bufs(nrot);
    for(void*& b : bufs)
    {
      CK(cudaMalloc(&b, bytes));
      CK(cudaMemcpy(b, host.data(), l0, cudaMemcpyHostToDevice));
    }
    nvpyrDispatchDesc d;
    memset(&d, 0, sizeof d);
    d.structSize = sizeof d;
    d.format     = c.fmt;
    d.extent     = {c.w, c.h};
    d.stream     = reinterpret_cast<nvpyrStream>(stream);
    const uint64_t launches0 = nvpyrGetLaunchCount();
    d.base                   = bufs[0];
    if(nvpyrDispatchEx(&d) != NVPYR_SUCCESS)
      return 3;
    const uint64_t launches = nvpyrGetLaunchCount() - launches0;
    CK(cudaStreamSynchronize(stream));
    std::vector<double> ns, hostNs;  // device time per chain (events) and host time to enqueue one chain
    for(int b = 0; b <= batches; ++b)
    {
      CK(cudaEventRecord(e0, stream));
      const auto h0 = std::chrono::steady_clock::now();
      for(int i = 0; i < 8; ++i)
      {
        d.base = bufs[(b * 8 + i) % nrot];
        nvpyrDispatchEx(&d);
      }
      const auto h1 = std::chrono::steady_clock::now();
      if(b)
        hostNs.push_back(std::chrono::duration<double, std::nano>(h1 - h0).count() / 8);
      CK(cudaEventRecord(e1, stream));
      CK(cudaStreamSynchronize(stream));
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      if(b)
        ns.push_back(ms * 1e6 / 8);
    }
    std::sort(ns.begin(), ns.end());
    std::sort(hostNs.begin(), hostNs.end());
    const double mn = ns.front(), med = ns[ns.size() / 2], hostMed = hostNs[hostNs.size() / 2];
    printf("%-20s %5ux%-5u %-8s launches %2llu  min %9.1f us  median %9.1f us  %8.1f GB/s  (%.1f%% of HBM peak)  host enqueue %5.1f us\n", c.name, c.w,
           c.h, c.fmt == NVPYR_FORMAT_SRGBA8 ? "srgba8" : "rgba32f", (unsigned long long)launches, mn / 1e3, med / 1e3, bytes / med,
           100.0 * bytes / med / peak, hostMed / 1e3);
    fflush(stdout);
    if(jf)
    {
      fprintf(jf, "%s {\"config\": \"%s\", \"w\": %u, \"h\": %u, \"format\": \"%s\", \"launches\": %llu, \"min_ns\": %.0f, \"median_ns\": %.0f, "
                  "\"algorithmic_bytes\": %llu, \"GBps_at_median\": %.1f, \"frac_of_hbm_peak\": %.3f, \"rotating_buffers\": %d, "
                  "\"content\": \"%s\", \"harness\": \"C++ (tools/bench_native.cpp), 8 back-to-back nvpyrDispatchEx calls between two events\"}",
              first ? "" : ",\n", c.name, c.w, c.h, c.fmt == NVPYR_FORMAT_SRGBA8 ? "srgba8" : "rgba32f", (unsigned long long)launches, mn, med,
              (unsigned long long)bytes, bytes / med, bytes / med / peak, nrot, opaque ? "uniform random colours, alpha 255" : "uniform random bytes");
      first = false;
    }
    for(void* b : bufs)
      cudaFree(b);
  }
  if(jf)
  {
    fprintf(jf, "\n]\n");
    fclose(jf);
  }
  return 0;
}
