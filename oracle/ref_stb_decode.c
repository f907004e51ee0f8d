/* ref_stb_decode.c — decodes an image with the stb_image.h the REFERENCE vendors
 * (libs/zstbi/libs/stbi/stb_image.h, v2.28), compiled from where it lies under /root/reference
 * (-I on the command line; nothing is copied into this repo).  zstbi.Image.loadFromFile(path, 4)
 * (libs/zstbi/src/zstbi.zig:118-152) is stbi_load(path, &w, &h, &c, 4); this tool writes the same
 * RGBA8 bytes as "w h\n" + raw data so tools/make_earthmap_fixture.py can commit them as a fixture.
 * Built into oracle/_ref/ (git-ignored).  Test infrastructure only. */
#define STB_IMAGE_IMPLEMENTATION
#include "stb_image.h"
#include <stdio.h>

int main(int argc, char** argv) {
    if (argc != 3) {
        fprintf(stderr, "usage: %s in.jpg out.rgba\n", argv[0]);
        return 2;
    }
    int w, h, c;
    unsigned char* data = stbi_load(argv[1], &w, &h, &c, 4);
    if (!data) {
        fprintf(stderr, "decode failed: %s\n", stbi_failure_reason());
        return 1;
    }
    FILE* f = fopen(argv[2], "wb");
    if (!f) return 1;
    fprintf(f, "%d %d %d\n", w, h, c);
    fwrite(data, 1, (size_t)w * h * 4, f);
    fclose(f);
    stbi_image_free(data);
    return 0;
}
