/*
 * stbref.c — TEST INFRASTRUCTURE ONLY. Compiles the reference's OWN vendored image libraries in place
 * (/root/reference/deps/include/stb_image.h, stb_image_resize2.h; nothing is copied into this repo) and
 * exposes the two calls the reference's texture bake makes:
 *   tinygltf LoadImageData -> stbi_load_from_memory(bytes, size, &w, &h, &comp, 4)   (deps/include/tiny_gltf.h:2603-2638)
 *   ImageManager::upload_image -> stbir_resize_uint8_srgb(..., 512, 512, 0, STBIR_RGBA) (src/image_manager.hpp:52-62)
 * Used by tests/tools/make_golden_images.py to produce tests/golden/images.npz and by the tests (when
 * /root/reference is present) to check the host image codecs of sycl-ray-tracer_b200/host/image_codecs.hpp.
 * Built by `make -C oracle ref` into oracle/_ref/libstbref.so (git-ignored).
 */
#define STB_IMAGE_IMPLEMENTATION
#define STBI_NO_STDIO
#include "stb_image.h"
#define STB_IMAGE_RESIZE_IMPLEMENTATION
#include "stb_image_resize2.h"

#include <string.h>

/* returns 1 and fills out (w*h*4 bytes, caller-allocated with capacity cap) or 0 on failure / too small */
int stbref_load_rgba(const unsigned char *bytes, int size, int *w, int *h, int *comp, unsigned char *out, int cap) {
    unsigned char *d = stbi_load_from_memory(bytes, size, w, h, comp, 4);
    if (!d) return 0;
    const int need = *w * *h * 4;
    if (need > cap) {
        stbi_image_free(d);
        return 0;
    }
    memcpy(out, d, (size_t)need);
    stbi_image_free(d);
    return 1;
}

int stbref_is_16_bit(const unsigned char *bytes, int size) { return stbi_is_16_bit_from_memory(bytes, size); }

const char *stbref_failure_reason(void) { return stbi_failure_reason(); }

int stbref_resize_srgb_rgba(const unsigned char *in, int w, int h, unsigned char *out, int ow, int oh) {
    return stbir_resize_uint8_srgb(in, w, h, 0, out, ow, oh, 0, STBIR_RGBA) == out;
}
