// minimal_mipmaps -- the reference's minimal_app (minimal_app/minimal_mipmaps.cpp) with the Vulkan
// upload / nvproCmdPyramidDispatch / download sequence (:134-217) replaced by one libnvpyr call.
// Same command line (:300-380), same output files (writeMipmapsTga naming and TGA flavour).
//
//   minimal_mipmaps -i input.{tga,ppm,pgm} -o out.tga [-force-no-fast-pipeline]
//                   [-premultiplied-alpha | -do-premultiply-alpha]
//
// Differences: inputs are TGA or binary PPM/PGM (no JPEG/PNG decoder is bundled, the reference uses
// stb_image); there is no device-capability probe -- the fast pipeline needs nothing optional on sm_100.
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "nvpyr.h"

// Command line: the option NAMES are the reference tool's (so that scripts written for it keep working); the parser
// is a plain table walk.
struct Options
{
  std::string input  = "4096.tga";
  std::string output = "./vk_compute_mipmaps_minimal.tga";
  uint32_t    flags  = NVPYR_FLAG_NONE;
};

static void usage(const char* program, FILE* to)
{
  fprintf(to,
          "usage: %s [-i image] [-o level-file-template] [-force-no-fast-pipeline]\n"
          "          [-premultiplied-alpha | -do-premultiply-alpha]\n\n"
          "Builds the full mip chain of one sRGBA8 image on the GPU and writes one TGA per level.\n"
          "  -i FILE    level 0: TGA, or binary PPM / PGM (default 4096.tga)\n"
          "  -o FILE    output name; the level number is inserted before the extension\n"
          "  -force-no-fast-pipeline   general (NPOT) pipeline for every level\n"
          "  -do-premultiply-alpha     level 0 has straight alpha: premultiply it first\n"
          "  -premultiplied-alpha      level 0 is premultiplied already (default)\n"
          "Either way the files written hold premultiplied alpha.\n",
          program);
}

static Options parseOptions(int argc, char** argv)
{
  struct Switch
  {
    const char* name;
    uint32_t    set, clear;
  };
  static const Switch switches[] = {{"-force-no-fast-pipeline", NVPYR_FLAG_FORCE_GENERAL, 0u},
                                    {"-do-premultiply-alpha", NVPYR_FLAG_PREMULTIPLY_ALPHA, 0u},
                                    {"-premultiplied-alpha", 0u, NVPYR_FLAG_PREMULTIPLY_ALPHA}};
  Options opt;
  int     at = 1;
  while(at < argc)
  {
    const std::string word = argv[at++];
    if(word == "-h" || word == "--help" || word == "/?")
    {
      usage(argv[0], stdout);
      exit(EXIT_SUCCESS);
    }
    std::string* value = word == "-i" ? &opt.input : word == "-o" ? &opt.output : nullptr;
    if(value != nullptr)
    {
      if(at == argc)
      {
        fprintf(stderr, "%s: option %s needs a file name\n", argv[0], word.c_str());
        exit(EXIT_FAILURE);
      }
      *value = argv[at++];
      continue;
    }
    bool known = false;
    for(const Switch& sw : switches)
      if(word == sw.name)
      {
        opt.flags = (opt.flags | sw.set) & ~sw.clear;
        known     = true;
      }
    if(!known)
    {
      fprintf(stderr, "%s: unrecognised option '%s'\n\n", argv[0], word.c_str());
      usage(argv[0], stderr);
      exit(EXIT_FAILURE);
    }
  }
  return opt;
}

static void check(nvpyrStatus st, const char* what)
{
  if(st == NVPYR_SUCCESS)
    return;
  fprintf(stderr, "%s: %s (cudaError %d)\n", what, nvpyrGetErrorString(st), nvpyrGetLastCudaError());
  exit(EXIT_FAILURE);
}

int main(int argc, char** argv)
{
  const Options opt = parseOptions(argc, argv);

  // Load image from file (ScopedImage::stageImage, scoped_image.hpp:210-262).
  void*         pixels = nullptr;
  nvpyrExtent2D extent{};
  fprintf(stderr, "Loading: '%s'...", opt.input.c_str());
  check(nvpyrReadImage(opt.input.c_str(), &pixels, &extent), "nvpyrReadImage");
  fprintf(stderr, " done (%u x %u)\n", extent.width, extent.height);

  // One staging chain, level 0 in place -- the reference's staging buffer (scoped_image.hpp:436-453).
  uint64_t chainBytes = 0;
  check(nvpyrGetChainBytes(extent, 0, NVPYR_FORMAT_SRGBA8, &chainBytes), "nvpyrGetChainBytes");
  std::vector<unsigned char> chain(chainBytes);
  memcpy(chain.data(), pixels, size_t(extent.width) * extent.height * 4u);
  nvpyrFree(pixels);

  // Upload, premultiply (optional), generate every level, download: replaces
  // cmdReallocUploadImage + nvproCmdPyramidDispatch + cmdDownloadImage (minimal_mipmaps.cpp:134-217).
  // NVPYR_FLAG_FORCE_GENERAL stands for pipelines.fastPipeline = VK_NULL_HANDLE (:192-196).
  check(nvpyrGenerateHost(chain.data(), chain.data(), extent, 0, NVPYR_FORMAT_SRGBA8, opt.flags), "nvpyrGenerateHost");

  // Write to disk (writeMipmapsTga, mipmap_storage.hpp:441-479).
  check(nvpyrWriteChainTga(chain.data(), extent, 0, opt.output.c_str()), "nvpyrWriteChainTga");
  char name[4096];
  for(uint32_t level = 0; level < nvpyrGetLevelCount(extent); ++level)
    if(nvpyrGetLevelFilename(opt.output.c_str(), level, name, sizeof name) == NVPYR_SUCCESS)
      fprintf(stderr, "Wrote %s\n", name);
  nvpyrShutdown();
  return EXIT_SUCCESS;
}
