// Shared internals of libcuburn_b200: error plumbing and small helpers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/cuburn_b200.h"

void cb_set_error(const char *fmt, ...);

#define CB_CUDA(expr)                                                         \
    do {                                                                      \
        cudaError_t err__ = (expr);                                           \
        if (err__ != cudaSuccess) {                                           \
            (void)cudaGetLastError(); /* do not leave it for a later check */ \
            cb_set_error("%s failed: %s (%s:%d)", #expr,                      \
                         cudaGetErrorString(err__), __FILE__, __LINE__);      \
            return err__ == cudaErrorMemoryAllocation ? CB_ERR_NOMEM          \
                                                      : CB_ERR_CUDA;          \
        }                                                                     \
    } while (0)

#define CB_REQUIRE(cond, msg)                                                 \
    do {                                                                      \
        if (!(cond)) {                                                        \
            cb_set_error("invalid argument: %s (%s)", msg, #cond);            \
            return CB_ERR_INVALID;                                            \
        }                                                                     \
    } while (0)

// kernels of this library launched so far (cb_launch_count)
extern unsigned long long g_cb_launches;

#define CB_LAUNCH_CHECK()                                                     \
    do {                                                                      \
        g_cb_launches++;                                                      \
        cudaError_t err__ = cudaGetLastError();                               \
        if (err__ != cudaSuccess) {                                           \
            cb_set_error("kernel launch failed: %s (%s:%d)",                  \
                         cudaGetErrorString(err__), __FILE__, __LINE__);      \
            return CB_ERR_CUDA;                                               \
        }                                                                     \
    } while (0)

static inline cudaStream_t cb_cs(cb_stream s) { return (cudaStream_t)s; }

template <typename T>
static inline T *cb_ptr(cb_dptr p) { return reinterpret_cast<T *>(p); }

// SM count of the device selected by cb_init (148 on B200).
int cb_sm_count();
