// minimal_mipmaps -- the reference's minimal_app (minimal_app/minimal_mipmaps.cpp) with the Vulkan
// upload / nvproCmdPyramidDispatch / download sequence (:134-217) replaced by one libnvpyr call.
// Same command line (:300-380), same output files (writeMipmapsTga naming and TGA flavour).
//
//   minimal_mipmaps -i input.{tga,ppm,pgm} -o out.tga [-force-no-fast-pipeline]
//                   [-premultiplied-alpha | -do-premultiply-alpha]
//
// Differences: inputs are TGA or binary PPM/PGM (no JPEG/PNG decoder is bundled, the reference uses
// stb_image); there is no device-capability probe -- the fast pipeline needs nothing optional on sm_100.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "nvpyr.h"

struct Config
{
  bool        forceDisableFastPipeline = false;  // minimal_mipmaps.cpp:42-43
  bool        doPremultiplyAlpha       = false;  // :45-47
  std::string rawInputFilename         = "4096.tga";
  std::string outputFilenameTemplate   = "./vk_compute_mipmaps_minimal.tga";  // :53
  Config(int argc, char** argv);
};

static const char helpString[] =
    "%s:\n    Generates mipmaps for an input image and exports as TGA.\n"
    "\n"
    "    ** Arguments **\n"
    "-i [input filename] (TGA or binary PPM/PGM)\n"
    "-o [output filename] (will be annotated with mip level numbers)\n"
    "-force-no-fast-pipeline: debug tool, never use the fast pipeline.\n"
    "-premultiplied-alpha: indicate input image has premultiplied alpha.\n"
    "-do-premultiply-alpha: indicate input image does not have premultiplied\n"
    "    alpha, so the program must do this itself.\n"
    "Note that output images have premultiplied alpha in either case.\n";

Config::Config(int argc, char** argv)
{
  for(int i = 1; i < argc; ++i)
  {
    const char* arg    = argv[i];
    const char* param0 = argv[i + 1];  // argv[argc] is NULL
    auto        needed = [&] {
      if(param0 == nullptr)
      {
        fprintf(stderr, "%s: %s missing parameter\n", argv[0], arg);
        exit(EXIT_FAILURE);
      }
    };
    if(strcmp(arg, "-h") == 0 || strcmp(arg, "/?") == 0)
    {
      printf(helpString, argv[0]);
      exit(EXIT_SUCCESS);
    }
    else if(strcmp(arg, "-i") == 0)
      needed(), rawInputFilename = param0, ++i;
    else if(strcmp(arg, "-o") == 0)
      needed(), outputFilenameTemplate = param0, ++i;
    else if(strcmp(arg, "-force-no-fast-pipeline") == 0)
      forceDisableFastPipeline = true;
    else if(strcmp(arg, "-premultiplied-alpha") == 0)
      doPremultiplyAlpha = false;
    else if(strcmp(arg, "-do-premultiply-alpha") == 0)
      doPremultiplyAlpha = true;
    else
    {
      fprintf(stderr, "%s: Unknown argument '%s'\n", argv[0], arg);
      exit(EXIT_FAILURE);
    }
  }
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
  const Config config(argc, argv);

  // Load image from file (ScopedImage::stageImage, scoped_image.hpp:210-262).
  void*         pixels = nullptr;
  nvpyrExtent2D extent{};
  fprintf(stderr, "Loading: '%s'...", config.rawInputFilename.c_str());
  check(nvpyrReadImage(config.rawInputFilename.c_str(), &pixels, &extent), "nvpyrReadImage");
  fprintf(stderr, " done (%u x %u)\n", extent.width, extent.height);

  // One staging chain, level 0 in place -- the reference's staging buffer (scoped_image.hpp:436-453).
  uint64_t chainBytes = 0;
  check(nvpyrGetChainBytes(extent, 0, NVPYR_FORMAT_SRGBA8, &chainBytes), "nvpyrGetChainBytes");
  std::vector<unsigned char> chain(chainBytes);
  memcpy(chain.data(), pixels, size_t(extent.width) * extent.height * 4u);
  nvpyrFree(pixels);

  // Upload, premultiply (optional), generate every level, download: replaces
  // cmdReallocUploadImage + nvproCmdPyramidDispatch + cmdDownloadImage (minimal_mipmaps.cpp:134-217).
  uint32_t flags = NVPYR_FLAG_NONE;
  if(config.forceDisableFastPipeline)
    flags |= NVPYR_FLAG_FORCE_GENERAL;  // pipelines.fastPipeline = VK_NULL_HANDLE (:192-196)
  if(config.doPremultiplyAlpha)
    flags |= NVPYR_FLAG_PREMULTIPLY_ALPHA;
  check(nvpyrGenerateHost(chain.data(), chain.data(), extent, 0, NVPYR_FORMAT_SRGBA8, flags), "nvpyrGenerateHost");

  // Write to disk (writeMipmapsTga, mipmap_storage.hpp:441-479).
  check(nvpyrWriteChainTga(chain.data(), extent, 0, config.outputFilenameTemplate.c_str()), "nvpyrWriteChainTga");
  char name[4096];
  for(uint32_t level = 0; level < nvpyrGetLevelCount(extent); ++level)
    if(nvpyrGetLevelFilename(config.outputFilenameTemplate.c_str(), level, name, sizeof name) == NVPYR_SUCCESS)
      fprintf(stderr, "Wrote %s\n", name);
  nvpyrShutdown();
  return EXIT_SUCCESS;
}
