/*
 * cuburn_b200 -- C ABI of the B200 (sm_100a) render hot path.
 *
 * The reference renderer has no FFI of its own: its host code
 * (cuburn/render.py, cuburn/filters.py, cuburn/output.py) drives named CUDA
 * kernels through PyCUDA.  This library sits exactly where PyCUDA sits: device
 * memory / streams / events, run-time compilation of the per-genome iterate
 * module (NVRTC, sm_100a), and one entry point per kernel of the path.  Every
 * entry point below names the reference call site or kernel it replaces
 * (paths relative to the reference tree).
 *
 * Conventions
 *   - every function returns CB_OK (0) or a negative cb_status; the message of
 *     the last failure on the calling thread is available from cb_last_error()
 *   - device pointers are passed as uint64_t (cb_dptr); host pointers are plain
 *   - all launches are asynchronous on the given stream; only cb_stream_sync,
 *     cb_event_sync and cb_device_sync block
 *   - not thread-safe per device, like the reference (render.py:401-402)
 */
#ifndef CUBURN_B200_H
#define CUBURN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint64_t cb_dptr;
typedef struct cb_stream_s *cb_stream;
typedef struct cb_event_s *cb_event;
typedef struct cb_module_s *cb_module;
typedef struct cb_comm_s *cb_comm;

typedef enum {
    CB_OK = 0,
    CB_ERR_CUDA = -1,        /* CUDA runtime / driver failure */
    CB_ERR_NVRTC = -2,       /* run-time compilation failed; log in cb_last_error */
    CB_ERR_INVALID = -3,     /* bad argument */
    CB_ERR_NOMEM = -4,       /* device or pinned host allocation failed
                                (pycuda MemoryError, render.py:140-147) */
    CB_ERR_NOT_READY = 1     /* cb_event_query: work still in flight */
} cb_status;

/* Accumulation-buffer geometry (render.py:80-89 Framebuffers.calc_dim):
 * gutter 12, awidth = width + 24, aheight = 16*ceil((height+24)/16),
 * astride = 32*ceil(awidth/32). */
typedef struct {
    int32_t width, height, awidth, aheight, astride;
} cb_dims;

/* ---- library / device --------------------------------------------------- */
const char *cb_last_error(void);
const char *cb_version(void);
int cb_device_count(int *count);                          /* main.py:96-107 */
int cb_device_info(int device, char *name, size_t name_len, int *cc_major,
                   int *cc_minor, int *sm_count, size_t *total_mem,
                   size_t *l2_bytes);
int cb_init(int device);                                  /* main.py:38-41 */
int cb_device_sync(void);
/* Number of kernels this library has launched in this process (its own kernels and the
 * NVRTC-built ones; memsets, copies and NCCL collectives are not counted). */
int cb_launch_count(uint64_t *count);
int cb_calc_dim(int width, int height, cb_dims *out);     /* render.py:80-89 */

/* ---- memory, streams, events (PyCUDA driver calls, SURVEY 8b index) ------- */
int cb_malloc(size_t bytes, cb_dptr *out);                /* render.py:133-138 */
int cb_free(cb_dptr p);
int cb_host_alloc(size_t bytes, void **out);              /* pinned; render.py:93 */
int cb_host_free(void *p);
/* Page-lock memory the caller owns (e.g. a shared-memory frame several GPU processes
 * copy their bands into) so that device->host copies into it are asynchronous. */
int cb_host_register(void *p, size_t bytes);
int cb_host_unregister(void *p);
int cb_stream_create(cb_stream *out);                     /* render.py:92,261 */
int cb_stream_destroy(cb_stream s);
int cb_stream_sync(cb_stream s);
int cb_stream_wait_event(cb_stream s, cb_event e);        /* render.py:359,372 */
int cb_event_create(cb_event *out);
int cb_event_destroy(cb_event e);
int cb_event_record(cb_event e, cb_stream s);
int cb_event_query(cb_event e);                           /* CB_OK | CB_ERR_NOT_READY */
int cb_event_sync(cb_event e);
int cb_event_elapsed_ms(cb_event start, cb_event stop, float *ms);
int cb_memcpy_h2d(cb_dptr dst, const void *src, size_t bytes, cb_stream s);
int cb_memcpy_d2h(void *dst, cb_dptr src, size_t bytes, cb_stream s);
int cb_memcpy_d2d(cb_dptr dst, cb_dptr src, size_t bytes, cb_stream s);
/* 32-bit fill usable in a stream (code/util.py:240-263 fill_dptr) */
int cb_fill32(cb_dptr dst, size_t nwords, uint32_t value, cb_stream s);

/* ---- MWC RNG (code/mwc.py) ---------------------------------------------- */
/* test_mwc (code/mwc.py:81-87): every stream advances `rounds` steps, the sum
 * of its outputs goes to sums[i] (u64) and the state is written back. */
int cb_mwc_test(cb_dptr seeds, int nstreams, int rounds, cb_dptr sums,
                cb_stream s);

/* ---- genome interpolation (code/interp.py) ------------------------------- */
/* Catmull-Rom evaluation of every packed knot row at every temporal sample
 * (catmull_rom / catmull_rom_mag, interp.py:295-366).  vals[ts][row] =
 * spline_row(tstart + ts*tstep); row_mag[row] != 0 selects the magnitude
 * domain.  times/knots are [nrows][32] float. */
int cb_interp_rows(cb_dptr vals, cb_dptr times, cb_dptr knots, cb_dptr row_mag,
                   int nrows, float tstart, float tstep, int nts, cb_stream s);
/* The precalc stage of interp_iter_params (interp.py:235-271 + the precalc
 * hunks in code/iter.py:12-30,56-95 and code/variations.py): runs the packed
 * precalc program (int32 [nops][12], see cuburn_b200/code/packer.py) for each
 * temporal sample, reading vals[ts][*] and writing params[ts][0..nslots). */
int cb_interp_params(cb_dptr params, int param_stride, cb_dptr vals, int nrows,
                     cb_dptr program, int nops, const cb_dims *dim, int nts,
                     cb_stream s);
/* interp_palette_flat (interp.py:372-433): palette row r (of nrows_out) is the
 * blend in YUV of the two palettes around t = tstart + r*tstep, biased by +0.5
 * in U,V, dithered by +-0.49 and quantised to 8 bits.  Output is float4
 * [nrows_out][256] = (Y,U,V)/255 and 1.0 (the unit the float4 histogram
 * accumulates, iter.py:395-406).  Entry (r, c) draws from RNG stream r*256+c. */
int cb_interp_palette(cb_dptr palette_out, cb_dptr seeds, cb_dptr ptimes,
                      cb_dptr pals, float tstart, float tstep, int nrows_out,
                      cb_stream s);

/* ---- per-genome iterate module (code/iter.py, render.py:232-246) --------- */
/* Compile CUDA source for sm_100a with NVRTC and load it.  `headers` /
 * `header_names` are in-memory include files.  On failure the NVRTC log is the
 * cb_last_error() text. */
int cb_module_build(const char *source, const char *name,
                    const char *const *headers, const char *const *header_names,
                    int nheaders, const char *const *options, int noptions,
                    cb_module *out);
int cb_module_destroy(cb_module m);
int cb_module_get_cubin(cb_module m, const void **cubin, size_t *size);
/* registers / static shared memory / max resident CTAs per SM of a kernel */
int cb_module_kernel_info(cb_module m, const char *kernel, int block_threads,
                          int *num_regs, int *static_smem, int *ctas_per_sm);
/* Bytes of local memory per thread (register spills, stack): the renderer lowers the
 * occupancy target of a module whose kernel spills at 32 registers. */
int cb_module_kernel_local_bytes(cb_module m, const char *kernel, int *local_bytes);
/* Generic launch of a kernel of a built module: args is an array of pointers
 * to the argument values (cuLaunchKernel convention). */
int cb_module_launch(cb_module m, const char *kernel, int gx, int gy, int gz,
                     int bx, int by, int bz, int dyn_smem, void **args,
                     cb_stream s);

/* Stream-ordered device-to-device copy into a __constant__ / __device__ variable
 * of the module (pycuda module.get_global + memcpy, render.py:290-293).  The
 * still variant of the iterate module keeps its parameter block in the
 * constant `c_params`. */
int cb_module_set_global(cb_module m, const char *symbol, cb_dptr src,
                         size_t bytes, cb_stream s);

typedef struct {
    cb_dptr hist;        /* float4 [aheight][astride] accumulation buffer.  The float4
                            modules add integer palette levels (sum Y, sum U, sum V of
                            8-bit levels, count): cb_hist_finish scales them to the
                            (sum/255, count) the filters consume */
    cb_dptr seeds;       /* mwc_st [nstreams] */
    cb_dptr points;      /* float4 [nstreams] trajectory state (x, y, color, -) */
    cb_dptr params;      /* float [nts][param_stride] from cb_interp_params */
    cb_dptr palette;     /* float4 [pal_rows][256] from cb_interp_palette */
    cb_dims dim;
    int32_t param_stride;
    int32_t nts;         /* temporal samples (1024, render.py:207) */
    int32_t pal_rows;    /* palette rows (64, render.py:202) */
    int32_t fuse_rounds; /* >0: reseed all points and run this many unrecorded
                            rounds first (iter.py:211-216) */
    int32_t swizzle_bins; /* 0: hist is linear.  Otherwise a multiple of 65536:
                            bins below it are stored in the slice-balancing
                            layout that cb_hist_unswizzle undoes */
    uint64_t first_sample;   /* global index of the first sample of this call */
    uint64_t nsamples;       /* samples (recorded xform applications) to run */
    uint64_t total_samples;  /* samples of the whole frame, all calls/GPUs */
    cb_dptr cells;           /* packed-accumulator module only: u64 [aheight][astride]
                                cells (count:10 | Y:18 | U:18 | V:18, iter.py:334-407) */
    cb_dptr palette_packed;  /* packed-accumulator and hot-bin modules: u64 [pal_rows][256]
                                from cb_palette_pack */
    cb_dptr hot_tags;        /* hot-bin module only: int32 [1025] from cb_hot_scan */
    int32_t first_round;     /* rounds every CTA has run in earlier calls of this frame
                                (fuse rounds included): a frame split over several calls
                                then draws the same samples as one call would */
    cb_dptr spill;           /* float4 modules: 0, or a zeroed float4 grid (layout of hist)
                                that receives the bins the in-kernel sweep finds full, so
                                that the float sums in hist stay below 2^24 and exact
                                (the reference's spill, iter.py:359-407, done by sweeping) */
    int32_t spill_bins;      /* bins every unit of 32768 samples examines */
    float spill_count;       /* a bin holding >= this many samples is moved to spill */
    cb_dptr tickets;         /* uint32 [2] scratch, zeroed by cb_iterate: unit and sweep counters */
    int32_t dynamic;         /* 0: CTA b runs units b, b + grid_ctas, ... (the sample set is a
                                pure function of the seeds); 1: CTAs claim units from a
                                counter as they become ready (balances the 2x spread in CTA
                                speed of a persistent grid; which stream draws which unit
                                then depends on timing) */
} cb_iter_args;
/* The chaos game (iter kernel, code/iter.py:157-418): nsamples iterations
 * accumulated into hist.  grid_ctas persistent CTAs of 256 threads; work is
 * split in units of 32768 samples and first_sample must be unit aligned. */
int cb_iterate(cb_module m, const cb_iter_args *args, int grid_ctas,
               cb_stream s);

/* Packed accumulation, for grids far larger than L2 (the reference's scheme,
 * iter.py:334-544).  cb_palette_pack turns the float4 palette table into the u64
 * addends ((1<<54) | Y<<36 | U<<18 | V, interp.py:428-429); the ACC_PACKED variant of
 * the iterate module adds them into `cells` and spills full cells into `hist`;
 * cb_flush_packed is flush_atom (iter.py:420-479): hist[i] += unpack(cells[i]). */
int cb_palette_pack(cb_dptr palette_packed, cb_dptr palette4, int nrows, cb_stream s);
int cb_flush_packed(cb_dptr hist4, cb_dptr cells, const cb_dims *dim, cb_stream s);

/* Hot bins.  A single histogram address absorbs ~6.5e8 reductions/s, so flames with very
 * bright bins are bound by those bins; the reference thins them (hotspot flags written
 * by flush_atom, iter.py:481-526, consumed at iter.py:319-329).  Here: after a short
 * pilot pass of cb_iterate, cb_hot_scan enters every bin of `hist4` (layout as
 * accumulated: swizzle_bins as in cb_iter_args) holding >= threshold samples into a
 * 1024-slot hash table: `tags` = int32 [1025], bin index or -1 per slot (the hotter bin
 * wins a shared slot), then the hash multiplier chosen for this frame.  `count` = int32
 * [4] (zeroed once): count[0] receives the number of bins holding >= trigger samples,
 * count[2] the number of table entries.  `scratch` = 8 x 1024 zeroed u64 (left zeroed).
 * The HOT_BINS variant of the iterate module accumulates the listed bins in shared
 * memory as integer level sums and folds them into the histogram itself.  `spill4`: 0 or
 * the spill grid of the pilot pass (a bin's count is the sum of both). */
int cb_hot_scan(cb_dptr tags, cb_dptr count, cb_dptr scratch, cb_dptr hist4,
                cb_dptr spill4, int swizzle_bins, float threshold, float trigger,
                const cb_dims *dim, cb_stream s);

/* Undo the accumulation layout: dst[i] = src[swizzle(i)] for i < swizzle_bins,
 * dst[i] = src[i] above; dst is the linear float4 [aheight][astride] histogram the
 * filters consume (iter.py:395-406 semantics). */
int cb_hist_unswizzle(cb_dptr dst4, cb_dptr src4, int swizzle_bins,
                      const cb_dims *dim, cb_stream s);
/* End of the float4 accumulation: dst[i] = (hist[j] + spill[j]) * (k, k, k, 1) with
 * j = swizzle(i) as above and k = level_scale (1/255: level sums -> the (sum Y/255,
 * sum U/255, sum V/255, count) of iter.py:395-406).  spill4 may be 0; dst4 may equal
 * hist4 only when swizzle_bins == 0. */
int cb_hist_finish(cb_dptr dst4, cb_dptr hist4, cb_dptr spill4, int swizzle_bins,
                   float level_scale, const cb_dims *dim, cb_stream s);

/* ---- sorting (cuburn/code/sort.py:385-520, helpers/sortbench.cu) ----------
 * One stable radix pass over 32-bit keys: dst receives the n keys of src grouped by the
 * `bits` (1..8) bits above lo_bit, equal digits in their original order, so that
 * least-significant-digit passes compose into a full sort (the reference's pass is
 * unstable and its multi-pass sort marked broken, sort.py:437-441).  ignore_max drops
 * keys equal to 0xffffffff (Sorter.sort's flag of the same name).  scratch: at least
 * cb_sort_scratch_words(n, bits) 32-bit words; afterwards scratch[d * groups], with
 * groups = ceil(n / 8192) (1 if n == 0), is the index in dst of the first key with digit
 * d, and scratch[words - 8] the number of keys kept. */
int cb_sort_scratch_words(uint64_t max_keys, int bits, uint64_t *words);
int cb_sort_pass(cb_dptr dst, cb_dptr src, uint64_t n, int lo_bit, int bits,
                 int ignore_max, cb_dptr scratch, cb_stream s);

/* ---- filters (code/filters.py; host recipes in cuburn/filters.py) -------- */
int cb_yuv_to_rgb(cb_dptr dst, cb_dptr src, const cb_dims *dim, cb_stream s);
int cb_den_blur(cb_dptr dst1, cb_dptr src4, int pattern, int upsample,
                const float coefs[7], const cb_dims *dim, cb_stream s);
int cb_den_blur_1c(cb_dptr dst1, cb_dptr src1, int pattern, int upsample,
                   const float coefs[7], const cb_dims *dim, cb_stream s);
int cb_full_blur(cb_dptr dst4, cb_dptr src4, int pattern, int upsample,
                 const float coefs[7], const cb_dims *dim, cb_stream s);
/* The reference's `bilateral` launch (cuburn/filters.py:86-94, code/filters.py:166-264):
 * blur1 is the density plane after cb_den_blur + cb_den_blur_1c(upsample 1).  Runs the
 * same restructured 31-tap kernels as cb_bilateral_direction; the per-pixel records
 * (8 bytes per bin) live in stream-ordered scratch memory for the duration of the call. */
int cb_bilateral(cb_dptr dst4, cb_dptr src4, cb_dptr blur1, int pattern,
                 int radius, float sstd, float cstd, float dstd, float dpow,
                 float gspeed, const cb_dims *dim, cb_stream s);
/* One whole direction pass of the bilateral recipe (cuburn/filters.py:80-94:
 * den_blur -> den_blur_1c(up=1) -> bilateral) in restructured form: per-pixel
 * terms (w^dpow, 1/(blur+1e-6)) are hoisted into two prologue kernels and the tap
 * weight is one exp2 of summed log2-domain terms.  scratch4: float4-sized. */
int cb_bilateral_direction(cb_dptr dst4, cb_dptr src4, cb_dptr scratch4,
                           int pattern, int radius, const float coefs[7],
                           float sstd, float cstd, float dstd, float dpow,
                           float gspeed, const cb_dims *dim, cb_stream s);
int cb_logscale(cb_dptr dst4, cb_dptr src4, float k1, float k2,
                const cb_dims *dim, cb_stream s);
int cb_apply_gamma(cb_dptr dst1, cb_dptr src4, float gamma, const cb_dims *dim,
                   cb_stream s);
int cb_haloclip(cb_dptr pix4, cb_dptr den1, float gamma_m_1, const cb_dims *dim,
                cb_stream s);
int cb_apply_gamma_full_hi(cb_dptr dst4, cb_dptr src4, float gamma_m_1,
                           const cb_dims *dim, cb_stream s);
int cb_smearclip(cb_dptr pix4, cb_dptr smear4, float gamma_m_1, float linrange,
                 float lingam, const cb_dims *dim, cb_stream s);
int cb_plainclip(cb_dptr pix4, float gamma_m_1, float linrange, float lingam,
                 float brightness, const cb_dims *dim, cb_stream s);
int cb_colorclip(cb_dptr pix4, float vibrance, float highpow, float gamma,
                 float linrange, float lingam, const cb_dims *dim, cb_stream s);
int cb_logencode(cb_dptr dst4, cb_dptr src4, float degamma, const cb_dims *dim,
                 cb_stream s);

/* ---- pixel-format output (code/output.py; launchC in output.py:21-26) ----- */
typedef enum {
    CB_FMT_RGBA_U8 = 0,     /* f32_to_rgba_u8   code/output.py:20-44   */
    CB_FMT_RGBA_U16 = 1,    /* f32_to_rgba_u16  code/output.py:47-71   */
    CB_FMT_YUV444P = 2,     /* f32_to_yuv444p   code/output.py:75-102  */
    CB_FMT_YUV444P10 = 3,   /* f32_to_yuv444p10 code/output.py:106-134 */
    CB_FMT_YUV420P10 = 4,   /* f32_to_yuv420p10 code/output.py:138-190 */
    CB_FMT_YUV444P12 = 5    /* f32_to_yuv444p12 code/output.py:194-225 */
} cb_pixfmt;
/* Crop the gutter, convert and dither-quantise src (float4 [ah][astride]) into
 * dst.  Pixel i (row-major over width x height) is produced by RNG stream
 * i % nstreams, each stream walking its pixels in increasing order. */
int cb_convert(cb_pixfmt fmt, cb_dptr dst, cb_dptr src, int gutter,
               const cb_dims *dim, cb_dptr seeds, int nstreams, cb_stream s);
int cb_convert_size(cb_pixfmt fmt, const cb_dims *dim, size_t *bytes);
/* The same for output rows [row0, row1) only (RGBA formats): every stream still draws
 * the random numbers of the rows it skips, so bands converted separately -- e.g. one per
 * GPU of a still whose filter chain is sharded by rows -- assemble into exactly the
 * frame cb_convert produces, and the seeds end in the same state. */
int cb_convert_rows(cb_pixfmt fmt, cb_dptr dst, cb_dptr src, int gutter,
                    const cb_dims *dim, cb_dptr seeds, int nstreams, int row0,
                    int row1, cb_stream s);

/* ---- multi-GPU exchange (one process per GPU, NCCL over NVLink) ----------
 * The reference ships whole frames between worker processes
 * (distribute.py:131-248) and has no device-side exchange.  A still split over
 * GPUs by samples needs one: the sum of the per-GPU histograms, and, when the
 * filter chain is sharded by rows too, the gather of the filtered bands.  NCCL
 * is bound at run time; without it these return CB_ERR_INVALID with a message.
 * Calls are stream-ordered on `s` like everything else. */
#define CB_COMM_ID_BYTES 128
int cb_comm_version(int *version);                       /* NCCL_VERSION_CODE */
/* One rank creates the id, every rank receives it out of band (file, socket,
 * torch.distributed ...) and joins with its rank. */
int cb_comm_unique_id(uint8_t id[CB_COMM_ID_BYTES]);
int cb_comm_create(const uint8_t id[CB_COMM_ID_BYTES], int rank, int world,
                   cb_comm *comm);
int cb_comm_destroy(cb_comm comm);
/* hist4 (float4 [aheight][astride]) += the other ranks' histograms, in place:
 * on `root` only, or on every rank when root = -1. */
int cb_hist_reduce(cb_comm comm, cb_dptr hist4, const cb_dims *dim, int root,
                   cb_stream s);
/* Rows [row0, row1) of the accumulation grid that `rank` filters: equally tall
 * bands, multiples of 16 rows; the last bands slide up (and overlap their
 * neighbour) when aheight does not divide. */
int cb_band_rows(int aheight, int rank, int world, int *row0, int *row1);
/* Collect every rank's band of frame4 (float4 [aheight][astride]) in root's
 * copy; a rank's own band must already be in place. */
int cb_band_gather(cb_comm comm, cb_dptr frame4, const cb_dims *dim, int root,
                   cb_stream s);

#ifdef __cplusplus
}
#endif
#endif /* CUBURN_B200_H */
