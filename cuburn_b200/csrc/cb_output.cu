// Pixel-format conversion: crop the gutter, apply the colour matrix, dither and
// quantise the filtered float4 buffer.
//
// Restates the f32_to_* kernels of the reference (cuburn/code/output.py:7-225)
// with a different work layout: instead of one thread per pixel grabbing an RNG
// block from a ring buffer, `nstreams` persistent threads each own one MWC
// stream and walk pixels i = stream, stream+nstreams, ... in order.  Output is
// therefore a pure function of (input, seeds), which lets the CPU oracle
// reproduce it bit for bit.  All arithmetic is explicit single-rounded IEEE
// (no FMA contraction), compiled without fast-math.
#include "cb_common.h"
#include "device/det_math.cuh"
#include "device/mwc.cuh"

// dclampf (code/output.py:7-13): true black stays black, everything else gets
// +[0, 0.99) of dither before truncation and is clamped to the peak.
__device__ __forceinline__ float dither_clamp(mwc_st &rng, float peak, float v) {
    float r = mwc_next_01(rng);       // drawn unconditionally: the stream position
                                      // must not depend on pixel content
    if (v > 0.0f) return fminf(peak, FA(FM(v, peak), FM(0.99f, r)));
    return 0.0f;
}

// JPEG full-range matrix (code/output.py:90-92)
__device__ __forceinline__ float jpeg_y(float4 p) {
    return FA(FA(FM(0.299f, p.x), FM(0.587f, p.y)), FM(0.114f, p.z));
}
__device__ __forceinline__ float jpeg_cb(float4 p) {
    return FA(FS(FM(-0.168736f, p.x), FM(0.331264f, p.y)), FM(0.5f, p.z));
}
__device__ __forceinline__ float jpeg_cr(float4 p) {
    return FS(FS(FM(0.5f, p.x), FM(0.418688f, p.y)), FM(0.081312f, p.z));
}

#define CV_BATCH 4

// Rows [row0, row1) of the output are converted; for the pixels of other rows a stream
// only draws the random numbers it would have used (RGBA formats: four per pixel), so
// that a frame converted band by band -- on one GPU or one band per GPU -- is identical
// to the frame converted at once, and leaves the same RNG state behind.
template <int FMT>
__global__ void __launch_bounds__(256)
k_convert(void *dstv, const float4 *src, int gutter, cb_dims dim, mwc_st *seeds,
          int nstreams, int row0, int row1) {
    int sid = blockIdx.x * blockDim.x + threadIdx.x;
    if (sid >= nstreams) return;
    mwc_st rng = seeds[sid];
    const int w = dim.width, h = dim.height;
    const int npix = w * h;

    // A stream's pixels are produced in order (its RNG state is carried along), but
    // their loads are independent: fetch CV_BATCH of them before converting any.
    for (int i0 = sid; i0 < npix; i0 += CV_BATCH * nstreams) {
      float4 batch[CV_BATCH];
#pragma unroll
      for (int k = 0; k < CV_BATCH; k++) {
          int i = i0 + k * nstreams;
          if (i < npix) {
              int y = i / w, x = i - y * w;
              if (y >= row0 && y < row1)
                  batch[k] = src[(y + gutter) * dim.astride + x + gutter];
          }
      }
#pragma unroll
      for (int k = 0; k < CV_BATCH; k++) {
        int i = i0 + k * nstreams;
        if (i >= npix) break;
        int y = i / w, x = i - y * w;
        if (y < row0 || y >= row1) {        // only reachable for the RGBA formats
            mwc_next(rng); mwc_next(rng); mwc_next(rng); mwc_next(rng);
            continue;
        }
        float4 p = batch[k];
        if (FMT == CB_FMT_RGBA_U8) {
            uchar4 o;
            o.x = (unsigned char)dither_clamp(rng, 255.0f, p.x);
            o.y = (unsigned char)dither_clamp(rng, 255.0f, p.y);
            o.z = (unsigned char)dither_clamp(rng, 255.0f, p.z);
            o.w = (unsigned char)dither_clamp(rng, 255.0f, p.w);
            reinterpret_cast<uchar4 *>(dstv)[i] = o;
        } else if (FMT == CB_FMT_RGBA_U16) {
            ushort4 o;
            o.x = (unsigned short)dither_clamp(rng, 65535.0f, p.x);
            o.y = (unsigned short)dither_clamp(rng, 65535.0f, p.y);
            o.z = (unsigned short)dither_clamp(rng, 65535.0f, p.z);
            o.w = (unsigned short)dither_clamp(rng, 65535.0f, p.w);
            reinterpret_cast<ushort4 *>(dstv)[i] = o;
        } else if (FMT == CB_FMT_YUV444P) {
            unsigned char *d = reinterpret_cast<unsigned char *>(dstv);
            d[i] = (unsigned char)dither_clamp(rng, 255.0f, jpeg_y(p));
            d[i + npix] = (unsigned char)dither_clamp(rng, 255.0f, FA(jpeg_cb(p), 0.5f));
            d[i + 2 * npix] = (unsigned char)dither_clamp(rng, 255.0f, FA(jpeg_cr(p), 0.5f));
        } else if (FMT == CB_FMT_YUV444P10) {
            unsigned short *d = reinterpret_cast<unsigned short *>(dstv);
            d[i] = (unsigned short)dither_clamp(rng, 1023.0f, jpeg_y(p));
            float cb = FA(jpeg_cb(p), 0.5f);
            (void)dither_clamp(rng, 1023.0f, cb);
            // the reference stores this plane without dither (code/output.py:129)
            d[i + npix] = (unsigned short)fminf(1023.0f, fmaxf(0.0f, FM(1023.0f, cb)));
            d[i + 2 * npix] = (unsigned short)dither_clamp(rng, 1023.0f, FA(jpeg_cr(p), 0.5f));
        } else if (FMT == CB_FMT_YUV420P10) {
            unsigned short *d = reinterpret_cast<unsigned short *>(dstv);
            d[i] = (unsigned short)dither_clamp(rng, 1023.0f, jpeg_y(p));
            // pixel (x, y) also owns chroma site (x, y) of the half-size planes
            if (2 * x < w && 2 * y < h) {
                const float4 *q = src + (2 * y + gutter) * dim.astride + 2 * x + gutter;
                float4 a = q[0], b = q[1], c = q[dim.astride], e = q[dim.astride + 1];
                float sum = (float)((double)a.w + 1e-12);
                float cb = FM(a.w, jpeg_cb(a)), cr = FM(a.w, jpeg_cr(a));
                sum = FA(sum, b.w); cb = FA(cb, FM(b.w, jpeg_cb(b))); cr = FA(cr, FM(b.w, jpeg_cr(b)));
                sum = FA(sum, c.w); cb = FA(cb, FM(c.w, jpeg_cb(c))); cr = FA(cr, FM(c.w, jpeg_cr(c)));
                sum = FA(sum, e.w); cb = FA(cb, FM(e.w, jpeg_cb(e))); cr = FA(cr, FM(e.w, jpeg_cr(e)));
                int ci = npix + (w / 2) * y + x;
                d[ci] = (unsigned short)dither_clamp(rng, 1023.0f, FA(FD(cb, sum), 0.5f));
                d[ci + npix / 4] = (unsigned short)dither_clamp(rng, 1023.0f, FA(FD(cr, sum), 0.5f));
            }
        } else {    // CB_FMT_YUV444P12: Rec.709, studio swing (code/output.py:194-225)
            unsigned short *d = reinterpret_cast<unsigned short *>(dstv);
            p.x = fminf(1.0f, fmaxf(0.0f, p.x));
            p.y = fminf(1.0f, fmaxf(0.0f, p.y));
            p.z = fminf(1.0f, fmaxf(0.0f, p.z));
            float yy = FA(FA(FM(0.2126f, p.x), FM(0.7152f, p.y)), FM(0.0722f, p.z));
            float cb = FA(FA(FS(FM(-0.11457f, p.x), FM(0.38543f, p.y)), FM(0.5f, p.z)), 0.5f);
            float cr = FA(FS(FS(FM(0.5f, p.x), FM(0.45416f, p.y)), FM(0.04585f, p.z)), 0.5f);
            d[i] = (unsigned short)FA(dither_clamp(rng, 3504.0f, yy), 256.0f);
            d[i + npix] = (unsigned short)FA(dither_clamp(rng, 3584.0f, cb), 256.0f);
            d[i + 2 * npix] = (unsigned short)FA(dither_clamp(rng, 3584.0f, cr), 256.0f);
        }
      }
    }
    seeds[sid] = rng;
}

extern "C" {

int cb_convert_size(cb_pixfmt fmt, const cb_dims *dim, size_t *bytes) {
    CB_REQUIRE(dim && bytes, "null argument");
    size_t npix = (size_t)dim->width * dim->height;
    switch (fmt) {
    case CB_FMT_RGBA_U8: *bytes = npix * 4; break;
    case CB_FMT_RGBA_U16: *bytes = npix * 8; break;
    case CB_FMT_YUV444P: *bytes = npix * 3; break;
    case CB_FMT_YUV444P10: *bytes = npix * 6; break;
    case CB_FMT_YUV420P10: *bytes = npix * 3; break;
    case CB_FMT_YUV444P12: *bytes = npix * 6; break;
    default: cb_set_error("unknown pixel format %d", (int)fmt); return CB_ERR_INVALID;
    }
    return CB_OK;
}

int cb_convert_rows(cb_pixfmt fmt, cb_dptr dst, cb_dptr src, int gutter,
                    const cb_dims *dim, cb_dptr seeds, int nstreams, int row0, int row1,
                    cb_stream s) {
    CB_REQUIRE(dim && nstreams > 0, "bad convert arguments");
    CB_REQUIRE(0 <= row0 && row0 <= row1 && row1 <= dim->height, "rows outside the frame");
    CB_REQUIRE((row0 == 0 && row1 == dim->height) ||
               fmt == CB_FMT_RGBA_U8 || fmt == CB_FMT_RGBA_U16,
               "row bands are supported for the RGBA formats");
    if (fmt == CB_FMT_YUV420P10)
        CB_REQUIRE(dim->width % 4 == 0 && dim->height % 2 == 0,
                   "yuv420p10 needs width % 4 == 0 and even height");
    int grid = (nstreams + 255) / 256;
    void *d = cb_ptr<void>(dst);
    const float4 *sp = cb_ptr<const float4>(src);
    mwc_st *sd = cb_ptr<mwc_st>(seeds);
#define GO(F) k_convert<F><<<grid, 256, 0, cb_cs(s)>>>(d, sp, gutter, *dim, sd, nstreams, row0, row1)
    switch (fmt) {
    case CB_FMT_RGBA_U8: GO(CB_FMT_RGBA_U8); break;
    case CB_FMT_RGBA_U16: GO(CB_FMT_RGBA_U16); break;
    case CB_FMT_YUV444P: GO(CB_FMT_YUV444P); break;
    case CB_FMT_YUV444P10: GO(CB_FMT_YUV444P10); break;
    case CB_FMT_YUV420P10: GO(CB_FMT_YUV420P10); break;
    case CB_FMT_YUV444P12: GO(CB_FMT_YUV444P12); break;
    default: cb_set_error("unknown pixel format %d", (int)fmt); return CB_ERR_INVALID;
    }
#undef GO
    CB_LAUNCH_CHECK();
    return CB_OK;
}

int cb_convert(cb_pixfmt fmt, cb_dptr dst, cb_dptr src, int gutter,
               const cb_dims *dim, cb_dptr seeds, int nstreams, cb_stream s) {
    CB_REQUIRE(dim, "bad convert arguments");
    return cb_convert_rows(fmt, dst, src, gutter, dim, seeds, nstreams, 0, dim->height, s);
}

}  // extern "C"
