/*
 * A C host for libcuburn_b200: the filter chain and the pixel-format output of a
 * frame, driven through the C ABI alone (no Python, no torch) -- the calls a
 * maintainer's binding would make where cuburn/filters.py and cuburn/output.py
 * launch their kernels (recipes: cuburn/filters.py:46-163, output.py:21-26).
 *
 *   gcc -O2 -Iinclude examples/c_host.c -Lcuburn_b200/csrc -lcuburn_b200 -lm \
 *       -Wl,-rpath,$PWD/cuburn_b200/csrc -o examples/c_host
 *   examples/c_host out.ppm        (exit status 77: no usable GPU)
 *
 * The histogram is synthetic (three soft blobs in the (sum Y, sum U, sum V, count)
 * format the iterate kernel writes); the per-genome iterate module itself needs the
 * generated source of cuburn_b200/code/itergen.py and is not part of this example.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "cuburn_b200.h"

#define CHECK(call)                                                          \
    do {                                                                     \
        int rc__ = (call);                                                   \
        if (rc__ != CB_OK) {                                                 \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc__, cb_last_error()); \
            exit(1);                                                         \
        }                                                                    \
    } while (0)

static void gauss7(float stdev, float c[7]) {
    float sum = 0;
    for (int i = 0; i < 7; i++) {
        c[i] = expf((float)((i - 3) * (i - 3)) / (-2.0f * stdev * stdev));
        sum += c[i];
    }
    for (int i = 0; i < 7; i++) c[i] /= sum;
}

int main(int argc, char **argv) {
    const int w = 320, h = 180, spp = 200, nstreams = 4096;
    if (cb_init(0) != CB_OK) {
        fprintf(stderr, "no usable GPU: %s\n", cb_last_error());
        return 77;
    }
    cb_dims dim;
    CHECK(cb_calc_dim(w, h, &dim));
    const size_t nbins = (size_t)dim.aheight * dim.astride;

    /* synthetic histogram: density blobs, colour = per-sample YUV (+0.5 chroma bias) */
    float *hist = (float *)calloc(nbins * 4, sizeof(float));
    const float blob[3][5] = {{0.30f, 0.40f, 0.9f, 0.45f, 0.60f},
                              {0.65f, 0.55f, 0.5f, 0.70f, 0.35f},
                              {0.50f, 0.75f, 0.7f, 0.30f, 0.30f}};
    for (int y = 0; y < dim.aheight; y++)
        for (int x = 0; x < dim.astride; x++)
            for (int b = 0; b < 3; b++) {
                float dx = (x - blob[b][0] * dim.awidth) / 30.0f;
                float dy = (y - blob[b][1] * dim.aheight) / 22.0f;
                float den = 4.0f * spp * expf(-(dx * dx + dy * dy));
                float *p = hist + 4 * ((size_t)y * dim.astride + x);
                p[0] += den * blob[b][2];
                p[1] += den * blob[b][3];
                p[2] += den * blob[b][4];
                p[3] += den;
            }

    cb_stream s;
    CHECK(cb_stream_create(&s));
    cb_dptr front, back, left, seeds;
    CHECK(cb_malloc(16 * nbins, &front));
    CHECK(cb_malloc(16 * nbins, &back));
    CHECK(cb_malloc(16 * nbins, &left));
    CHECK(cb_memcpy_h2d(front, hist, 16 * nbins, s));

    /* MWC streams for the output dither: {multiplier, state, carry} per stream */
    uint32_t *h_seeds = (uint32_t *)malloc(12 * nstreams);
    for (int i = 0; i < nstreams; i++) {
        h_seeds[3 * i + 0] = 0xffffff4eu;            /* first entry of the multiplier table */
        h_seeds[3 * i + 1] = 12345u + 7919u * (uint32_t)i;
        h_seeds[3 * i + 2] = 1u + 104729u * (uint32_t)i % 0x7fffffffu;
    }
    CHECK(cb_malloc(12 * nstreams, &seeds));
    CHECK(cb_memcpy_h2d(seeds, h_seeds, 12 * nstreams, s));

#define SWAP() do { cb_dptr t__ = front; front = back; back = t__; } while (0)
    /* yuv (filters.py:46-54) */
    CHECK(cb_yuv_to_rgb(back, front, &dim, s));
    SWAP();
    /* bilateral: eight directions, each consuming the previous output (filters.py:56-95) */
    float c1[7];
    gauss7(1.0f, c1);
    for (int pattern = 0; pattern < 8; pattern++) {
        CHECK(cb_bilateral_direction(back, front, left, pattern, 15, c1, 6.0f * w / 1920.0f,
                                     0.05f, 1.5f, 0.8f, 4.0f, &dim, s));
        SWAP();
    }
    /* logscale (filters.py:100-108): k1 = brightness 268/256, k2 = 1 / (area spp) */
    const float scale = 0.5f, brightness = 4.0f;
    const float area = (float)h / (scale * scale * (float)w);
    CHECK(cb_logscale(front, front, brightness * 268.0f / 256.0f, 1.0f / (area * spp), &dim, s));
    /* smearclip (filters.py:138-163) */
    const float gam = 1.0f / 4.0f, lin = 0.01f, lingam = powf(lin, gam - 1.0f);
    float cw[7];
    gauss7(0.7f, cw);
    CHECK(cb_apply_gamma_full_hi(left, front, gam - 1.0f, &dim, s));
    CHECK(cb_full_blur(back, left, 2, 0, cw, &dim, s));
    CHECK(cb_full_blur(left, back, 3, 0, cw, &dim, s));
    CHECK(cb_full_blur(back, left, 0, 0, cw, &dim, s));
    CHECK(cb_full_blur(left, back, 1, 0, cw, &dim, s));
    CHECK(cb_smearclip(front, left, gam - 1.0f, lin, lingam, &dim, s));
    /* output (output.py:21-26, 81-88): crop, dither, RGBA8 */
    size_t bytes = 0;
    CHECK(cb_convert_size(CB_FMT_RGBA_U8, &dim, &bytes));
    CHECK(cb_convert(CB_FMT_RGBA_U8, back, front, 12, &dim, seeds, nstreams, s));
    void *frame = NULL;
    CHECK(cb_host_alloc(bytes, &frame));
    CHECK(cb_memcpy_d2h(frame, back, bytes, s));
    CHECK(cb_stream_sync(s));

    const uint8_t *px = (const uint8_t *)frame;
    unsigned long sum = 0;
    int peak = 0;
    for (size_t i = 0; i < (size_t)w * h; i++)
        for (int ch = 0; ch < 3; ch++) {
            sum += px[4 * i + ch];
            if (px[4 * i + ch] > peak) peak = px[4 * i + ch];
        }
    printf("%s: %dx%d RGBA8, %zu bytes, mean level %.2f, peak %d\n", cb_version(), w, h, bytes,
           (double)sum / (3.0 * w * h), peak);
    if (argc > 1) {
        FILE *fp = fopen(argv[1], "wb");
        if (!fp) { perror(argv[1]); return 1; }
        fprintf(fp, "P6\n%d %d\n255\n", w, h);
        for (size_t i = 0; i < (size_t)w * h; i++) fwrite(px + 4 * i, 1, 3, fp);
        fclose(fp);
    }
    CHECK(cb_host_free(frame));
    CHECK(cb_free(front)); CHECK(cb_free(back)); CHECK(cb_free(left)); CHECK(cb_free(seeds));
    CHECK(cb_stream_destroy(s));
    free(hist); free(h_seeds);
    return peak > 32 ? 0 : 2;
}
