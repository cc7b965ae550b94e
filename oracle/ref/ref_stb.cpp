// Shim over the reference's texture decoder, stb_image.h as vendored under /root/reference/ext/stb, compiled where it lies
// (oracle/ref/Makefile).  TEST INFRASTRUCTURE: the reference loads every texture with stbi_load(path, &w, &h, &ch, 4)
// (zvk/core/HostImage.cpp:70-75, src/Resource.cpp:26).  Used to (1) decode the VeachAjar textures into the side-car files the
// host loads (tools/prepare_assets.py), so the measured texels are the reference decoder's, and (2) check host/Image.cpp.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"   // the reference's copy (-I$(REF)/ext/stb)

extern "C" __attribute__((visibility("default")))
uint8_t* ref_stbi_load_rgba8(const char* path, int* width, int* height) {
	int channels = 0;
	return stbi_load(path, width, height, &channels, 4);
}
extern "C" __attribute__((visibility("default")))
void ref_stbi_free(uint8_t* p) { stbi_image_free(p); }
