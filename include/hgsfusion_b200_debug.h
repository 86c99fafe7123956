/* hgsfusion_b200_debug.h -- measurement hooks of libhgsfusion_b200.so.  NOT part of the drop-in boundary: nothing a
 * reference-side binding needs is declared here, and these are the only entry points of the library that synchronise.
 * Used by bench.py (roofline leg) and scripts/. */
#ifndef HGSFUSION_B200_DEBUG_H
#define HGSFUSION_B200_DEBUG_H

#include "hgsfusion_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Measurement hook (off by default): after hgsf_emit_timing_begin(capacity > 0), every hgsf_points_to_bev /
 * hgsf_pillarize call made by this host thread records a CUDA event pair on its stream around the launch of
 * the path's dominant kernel (k_emit), up to `capacity` calls.  hgsf_emit_timing_collect synchronises those
 * events, writes the per-launch durations (ms) and returns how many (<0: -cudaError_t), and rearms the ring.
 * hgsf_emit_timing_begin(0) switches it off.  These two are the only entry points that synchronise. */
HGSF_API int hgsf_emit_timing_begin(int capacity);
HGSF_API int hgsf_emit_timing_collect(float *ms, int n);

#ifdef __cplusplus
}
#endif

#endif /* HGSFUSION_B200_DEBUG_H */
