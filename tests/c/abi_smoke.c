/* abi_smoke.c — a plain C consumer of libjpgpu.so: nothing but include/jpgpu.h, gcc and the shared library.
 * What a Rust/Go/Java binding does, minus the language: parse + decode one file through jpgpu_decode_file
 * (JPEGImage::parse, reference src/jpeg/mod.rs:202), then the same file through the batch and the host-pipeline
 * entry points, and print one line the calling test compares with the oracle:
 *     <width> <height> <bytes_read> <fnv1a64 of the W*H*3 output bytes> <batch == single> <pipeline == single>
 * usage: abi_smoke <file.jpg> [ext_flags] [layout]        (exit code = JPGPU status) */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jpgpu.h"

static unsigned long long fnv1a(const unsigned char *p, size_t n) {
    unsigned long long h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= p[i]; h *= 1099511628211ull; }
    return h;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: %s file.jpg [ext] [layout]\n", argv[0]); return 2; }
    const unsigned ext = argc > 2 ? (unsigned)atoi(argv[2]) : JPGPU_EXT_NONE;
    const unsigned layout = argc > 3 ? (unsigned)atoi(argv[3]) : JPGPU_LAYOUT_REF;
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END);
    const long len = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char *file = malloc((size_t)len);
    if (fread(file, 1, (size_t)len, f) != (size_t)len) { fprintf(stderr, "short read\n"); return 2; }
    fclose(f);

    if (jpgpu_abi_version() != JPGPU_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 2; }
    jpgpu_image_desc desc;
    int st = jpgpu_parse(file, (size_t)len, ext, layout, &desc);
    if (st != JPGPU_OK) { fprintf(stderr, "parse: %s\n", jpgpu_status_string(st)); return st; }
    const size_t nbytes = (size_t)desc.width * desc.height * 3;
    unsigned char *rgb = malloc(nbytes), *rgb2 = malloc(nbytes);

    jpgpu_ctx *ctx = NULL;
    st = jpgpu_create(0, &ctx);
    if (st != JPGPU_OK) { fprintf(stderr, "create: %s\n", jpgpu_status_string(st)); return st; }

    /* 1. the reference's call: one file in, pixels out */
    uint32_t w = 0, h = 0;
    size_t bytes_read = 0;
    st = jpgpu_decode_file(ctx, file, (size_t)len, ext, layout, rgb, nbytes, &w, &h, &bytes_read);
    if (st != JPGPU_OK) { fprintf(stderr, "decode_file: %s (%s)\n", jpgpu_status_string(st), jpgpu_last_error(ctx)); return st; }

    /* 2. the batch entry points on the same image */
    jpgpu_batch *b = NULL;
    int32_t bst = -1;
    uint64_t bbr = 0;
    uint8_t *outs[1] = {rgb2};
    memset(rgb2, 0, nbytes);
    st = jpgpu_batch_create(ctx, &desc, 1, &b);
    if (st == JPGPU_OK) st = jpgpu_batch_upload(b);
    if (st == JPGPU_OK) st = jpgpu_batch_decode(b);
    if (st == JPGPU_OK) st = jpgpu_batch_download(b, outs);
    if (st == JPGPU_OK) st = jpgpu_batch_results(b, &bst, &bbr);
    if (st != JPGPU_OK || bst != JPGPU_OK) { fprintf(stderr, "batch: %d / %d\n", st, (int)bst); return st ? st : bst; }
    const int batch_same = memcmp(rgb, rgb2, nbytes) == 0 && bbr == bytes_read;
    jpgpu_batch_destroy(b);

    /* 3. host buffers in, host buffers out through the pipeline (single copies each way) */
    size_t off = 0;
    int32_t pst = -1;
    uint64_t pbr = 0;
    unsigned char *pout = malloc(nbytes + 4096);
    st = jpgpu_decode_batch_host(0, &desc, 1, file, (size_t)len, pout, nbytes + 4096, &off, &pst, &pbr);
    if (st != JPGPU_OK || pst != JPGPU_OK) { fprintf(stderr, "pipeline: %d / %d\n", st, (int)pst); return st ? st : pst; }
    const int pipe_same = memcmp(rgb, pout + off, nbytes) == 0 && pbr == bytes_read;

    printf("%u %u %zu %016llx %d %d\n", w, h, bytes_read, fnv1a(rgb, nbytes), batch_same, pipe_same);
    jpgpu_destroy(ctx);
    free(pout); free(rgb2); free(rgb); free(file);
    return 0;
}
