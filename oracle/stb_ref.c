/* ORACLE — TEST INFRASTRUCTURE ONLY.  Compiles the reference's vendored image decoder (external/stb_image.h) where it
 * lies under $(REF) into oracle/_ref/libstb_ref.so, so the tests can compare the product's own JPEG / PNG decoder with
 * what the reference uploads as texture bytes (src/SceneLoader.cpp:198-207).  Nothing is copied into this repository. */
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"

unsigned char *stb_ref_load(const char *path, int *w, int *h) { int channels; return stbi_load(path, w, h, &channels, 4); }
void stb_ref_free(void *p) { stbi_image_free(p); }
