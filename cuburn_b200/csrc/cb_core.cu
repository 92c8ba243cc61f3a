// Device services of the C ABI: init, memory, streams, events, fills, and the
// MWC self-test kernel.  Replaces the PyCUDA driver calls indexed in
// SURVEY.md section 8(b).
#include <stdarg.h>
#include <string.h>

#include "cb_common.h"
#include "device/mwc.cuh"

static thread_local char g_err[16384] = "";
static int g_sm_count = 148;

void cb_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int cb_sm_count() { return g_sm_count; }

unsigned long long g_cb_launches = 0;

extern "C" {

const char *cb_last_error(void) { return g_err; }
const char *cb_version(void) { return "cuburn_b200 0.1.0 (sm_100a)"; }

int cb_device_count(int *count) {
    CB_REQUIRE(count, "count is null");
    CB_CUDA(cudaGetDeviceCount(count));
    return CB_OK;
}

int cb_device_info(int device, char *name, size_t name_len, int *cc_major,
                   int *cc_minor, int *sm_count, size_t *total_mem,
                   size_t *l2_bytes) {
    cudaDeviceProp p;
    CB_CUDA(cudaGetDeviceProperties(&p, device));
    if (name && name_len) {
        strncpy(name, p.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (total_mem) *total_mem = p.totalGlobalMem;
    if (l2_bytes) *l2_bytes = (size_t)p.l2CacheSize;
    return CB_OK;
}

int cb_init(int device) {
    CB_CUDA(cudaSetDevice(device));
    CB_CUDA(cudaFree(0));
    int sms = 0, major = 0;
    CB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    CB_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device));
    if (major != 10) {
        cb_set_error("device %d has compute capability %d.x; this library is "
                     "built for sm_100a (B200) only", device, major);
        return CB_ERR_INVALID;
    }
    g_sm_count = sms;
    return CB_OK;
}

int cb_device_sync(void) {
    CB_CUDA(cudaDeviceSynchronize());
    return CB_OK;
}

int cb_calc_dim(int width, int height, cb_dims *out) {
    CB_REQUIRE(out && width > 0 && height > 0, "bad dimensions");
    const int gutter = 12;
    out->width = width;
    out->height = height;
    out->awidth = width + 2 * gutter;
    out->aheight = 16 * ((height + 2 * gutter + 15) / 16);
    out->astride = 32 * ((out->awidth + 31) / 32);
    return CB_OK;
}

int cb_malloc(size_t bytes, cb_dptr *out) {
    CB_REQUIRE(out, "out is null");
    void *p = nullptr;
    CB_CUDA(cudaMalloc(&p, bytes ? bytes : 1));
    *out = (cb_dptr)p;
    return CB_OK;
}

int cb_free(cb_dptr p) {
    if (p) CB_CUDA(cudaFree((void *)p));
    return CB_OK;
}

int cb_host_alloc(size_t bytes, void **out) {
    CB_REQUIRE(out, "out is null");
    CB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
    return CB_OK;
}

int cb_host_free(void *p) {
    if (p) CB_CUDA(cudaFreeHost(p));
    return CB_OK;
}

int cb_launch_count(uint64_t *count) {
    CB_REQUIRE(count, "null argument");
    *count = g_cb_launches;
    return CB_OK;
}

int cb_host_register(void *p, size_t bytes) {
    CB_REQUIRE(p && bytes, "null argument");
    CB_CUDA(cudaHostRegister(p, bytes, cudaHostRegisterPortable));
    return CB_OK;
}

int cb_host_unregister(void *p) {
    if (p) CB_CUDA(cudaHostUnregister(p));
    return CB_OK;
}

int cb_stream_create(cb_stream *out) {
    CB_REQUIRE(out, "out is null");
    cudaStream_t s;
    CB_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *out = (cb_stream)s;
    return CB_OK;
}

int cb_stream_destroy(cb_stream s) {
    if (s) CB_CUDA(cudaStreamDestroy(cb_cs(s)));
    return CB_OK;
}

int cb_stream_sync(cb_stream s) {
    CB_CUDA(cudaStreamSynchronize(cb_cs(s)));
    return CB_OK;
}

int cb_stream_wait_event(cb_stream s, cb_event e) {
    CB_CUDA(cudaStreamWaitEvent(cb_cs(s), (cudaEvent_t)e, 0));
    return CB_OK;
}

int cb_event_create(cb_event *out) {
    CB_REQUIRE(out, "out is null");
    cudaEvent_t e;
    CB_CUDA(cudaEventCreateWithFlags(&e, cudaEventBlockingSync));
    *out = (cb_event)e;
    return CB_OK;
}

int cb_event_destroy(cb_event e) {
    if (e) CB_CUDA(cudaEventDestroy((cudaEvent_t)e));
    return CB_OK;
}

int cb_event_record(cb_event e, cb_stream s) {
    CB_CUDA(cudaEventRecord((cudaEvent_t)e, cb_cs(s)));
    return CB_OK;
}

int cb_event_query(cb_event e) {
    cudaError_t r = cudaEventQuery((cudaEvent_t)e);
    if (r == cudaSuccess) return CB_OK;
    if (r == cudaErrorNotReady) return CB_ERR_NOT_READY;
    cb_set_error("cudaEventQuery failed: %s", cudaGetErrorString(r));
    return CB_ERR_CUDA;
}

int cb_event_sync(cb_event e) {
    CB_CUDA(cudaEventSynchronize((cudaEvent_t)e));
    return CB_OK;
}

int cb_event_elapsed_ms(cb_event start, cb_event stop, float *ms) {
    CB_REQUIRE(ms, "ms is null");
    CB_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return CB_OK;
}

int cb_memcpy_h2d(cb_dptr dst, const void *src, size_t bytes, cb_stream s) {
    CB_CUDA(cudaMemcpyAsync((void *)dst, src, bytes, cudaMemcpyHostToDevice, cb_cs(s)));
    return CB_OK;
}

int cb_memcpy_d2h(void *dst, cb_dptr src, size_t bytes, cb_stream s) {
    CB_CUDA(cudaMemcpyAsync(dst, (const void *)src, bytes, cudaMemcpyDeviceToHost, cb_cs(s)));
    return CB_OK;
}

int cb_memcpy_d2d(cb_dptr dst, cb_dptr src, size_t bytes, cb_stream s) {
    CB_CUDA(cudaMemcpyAsync((void *)dst, (const void *)src, bytes, cudaMemcpyDeviceToDevice, cb_cs(s)));
    return CB_OK;
}

}  // extern "C"

// ---- fill ---------------------------------------------------------------------
// Each thread issues FILL_PER independent 16-byte stores per trip (a CTA covers
// FILL_PER contiguous 4 KiB runs), scalar tail for the last nwords % 4.
#define FILL_PER 8
__global__ void __launch_bounds__(256)
k_fill32(uint32_t *dst, size_t nwords, uint32_t value) {
    const size_t nvec = nwords >> 2;
    const uint4 v = make_uint4(value, value, value, value);
    uint4 *d4 = reinterpret_cast<uint4 *>(dst);
    const size_t chunk = (size_t)256 * FILL_PER;
    for (size_t base = (size_t)blockIdx.x * chunk; base < nvec; base += (size_t)gridDim.x * chunk) {
#pragma unroll
        for (int k = 0; k < FILL_PER; k++) {
            size_t i = base + (size_t)k * 256 + threadIdx.x;
            if (i < nvec) d4[i] = v;
        }
    }
    size_t t = (nvec << 2) + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t < nwords) dst[t] = value;
}

extern "C" int cb_fill32(cb_dptr dst, size_t nwords, uint32_t value, cb_stream s) {
    if (nwords == 0) return CB_OK;
    CB_REQUIRE((dst & 15) == 0, "fill destination must be 16-byte aligned");
    size_t nvec = nwords >> 2;
    size_t want = (nvec + 256 * FILL_PER - 1) / (256 * FILL_PER);
    size_t cap = (size_t)cb_sm_count() * 8;
    int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    k_fill32<<<grid, 256, 0, cb_cs(s)>>>(cb_ptr<uint32_t>(dst), nwords, value);
    CB_LAUNCH_CHECK();
    return CB_OK;
}

// ---- MWC self test -----------------------------------------------------------
__global__ void __launch_bounds__(256)
k_mwc_test(mwc_st *states, int nstreams, int rounds, unsigned long long *sums) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nstreams) return;
    mwc_st r = states[i];
    unsigned long long acc = 0;
    for (int k = 0; k < rounds; k++) acc += mwc_next(r);
    sums[i] = acc;
    states[i] = r;
}

extern "C" int cb_mwc_test(cb_dptr seeds, int nstreams, int rounds, cb_dptr sums,
                           cb_stream s) {
    CB_REQUIRE(nstreams > 0 && rounds >= 0, "bad stream/round count");
    k_mwc_test<<<(nstreams + 255) / 256, 256, 0, cb_cs(s)>>>(
        cb_ptr<mwc_st>(seeds), nstreams, rounds, cb_ptr<unsigned long long>(sums));
    CB_LAUNCH_CHECK();
    return CB_OK;
}
