// Where the packed parameter block of the current temporal sample lives.
// Included by the generated source after NSLOTS / PARAMS_CONST are defined and
// before any function that reads P[slot].
#pragma once

#ifndef ITER_THREADS
#define ITER_THREADS 256
#endif
#ifndef UNIT_ROUNDS
#define UNIT_ROUNDS 128
#endif
#ifndef ITER_MIN_CTAS
#define ITER_MIN_CTAS 8        // 32 registers, 64 warps / SM: measured fastest (profiles/r01_iter_variants.md)
#endif

#ifndef POINTS
#define POINTS 1
#endif
// The trajectories one thread carries through a round.
struct point_set {
    float x[POINTS], y[POINTS], c[POINTS];
    int last[POINTS];           // previous xform of each trajectory (xaos)
};

#if PARAMS_CONST
// Stills: one block for the whole launch.  Constant indices make every P[slot] a
// constant-bank operand of the consuming instruction (no load, no register).
__constant__ float c_params[NSLOTS > 0 ? NSLOTS : 1];
#define P c_params
#else
// Motion blur: the CTA stages the block of its unit's temporal sample here; generated
// code reads it as aligned float4s (P4) where slots are contiguous.
__shared__ __align__(16) float s_params[(NSLOTS + 3) / 4 * 4 + 4];
#define P s_params
#define P4 (reinterpret_cast<const float4 *>(s_params))
#endif

// xform opacity (genome/specs.py:17): the share of the xform's points that are drawn.
__device__ __forceinline__ bool opacity_visible(float opacity, mwc_st &rng) {
    return opacity >= 1.0f || mwc_next_01(rng) < opacity;
}
