// Declaration-only stand-in for stb_image_write.h (un-vendored third-party
// dependency of the reference); the one function the reference's
// include/mipmap_storage.hpp calls is defined in oracle/ref_harness.cpp.
#pragma once
extern "C" int stbi_write_tga(char const* filename, int w, int h, int comp, const void* data);
