/*
 * pillar_oracle.c -- CPU ORACLE for the radar pillarization hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in hgsfusion_b200/ may import, link or
 * execute this file; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker
 * or as the timed CPU baseline, never as the shipped path.
 *
 * It restates, in plain scalar fp32 C, what the reference computes on the CPU
 * for   points -> pillars -> PillarVFE -> PointPillarScatter.
 * Citations are file:line under /root/reference.
 *
 * Pinning status
 *   - orc_pillar_vfe / orc_pointpillar_scatter: PINNED.  Checked bit-for-bit
 *     against the reference's own pcdet/models/backbones_3d/vfe/pillar_vfe.py
 *     and pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py imported
 *     by file path (tests/golden/make_golden.py wrote the tests/golden npz fixtures).
 *   - orc_voxelize: PARITY UNPINNED.  The arithmetic lives in spconv
 *     (spconv.utils.Point2VoxelCPU3d, called at
 *     pcdet/datasets/processor/data_processor.py:37-43,55-60); spconv/cumm are
 *     neither vendored under /root/reference nor installed, and the reference
 *     pins no version (setup.py:48 comments it out; docs/INSTALL.md allows
 *     v1.0/v1.2/v2.x).  This function restates spconv v2.x's published
 *     point_to_voxel loop and is pinned only by hand-computed cases
 *     (tests/golden/voxelize_cases.json) and by structural invariants.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -mfma).
 * -ffp-contract=off matters: every fp32 operation below is rounded on its own
 * unless it is written as fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <pthread.h>

#define ORC_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------- */
/* Minimal static-partition parallel-for on pthreads (libgomp is not in the    */
/* image).  g_threads == 1 (default) runs inline: the scalar port.             */
static int g_threads = 1;
typedef void (*orc_range_fn)(int64_t lo, int64_t hi, void *ctx);
typedef struct { orc_range_fn fn; void *ctx; int64_t lo, hi; } orc_job;
static void *orc_job_main(void *p)
{
    orc_job *j = (orc_job *)p;
    j->fn(j->lo, j->hi, j->ctx);
    return NULL;
}
static void orc_parallel_for(int64_t n, orc_range_fn fn, void *ctx)
{
    int T = g_threads;
    if (T > n) T = (int)(n > 0 ? n : 1);
    if (T <= 1) { fn(0, n, ctx); return; }
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)T);
    orc_job *jobs = (orc_job *)malloc(sizeof(orc_job) * (size_t)T);
    for (int t = 0; t < T; ++t) {
        jobs[t].fn = fn; jobs[t].ctx = ctx;
        jobs[t].lo = n * t / T; jobs[t].hi = n * (t + 1) / T;
        if (t == T - 1) orc_job_main(&jobs[t]);
        else pthread_create(&th[t], NULL, orc_job_main, &jobs[t]);
    }
    for (int t = 0; t < T - 1; ++t) pthread_join(th[t], NULL);
    free(th);
    free(jobs);
}
ORC_API int orc_num_threads(void) { return g_threads; }
ORC_API void orc_set_num_threads(int n) { g_threads = n > 0 ? n : 1; }

/* ------------------------------------------------------------------------- */
/* a1: pcdet/utils/common_utils.py:78-81  mask_points_by_range                */
/* keep = x >= xmin && x <= xmax && y >= ymin && y <= ymax (no z test)        */
/* points: [n, stride] fp32, x at column xcol.  keep: [n] uint8.  returns kept */
ORC_API int64_t orc_mask_points_by_range(const float *points, int64_t n, int stride, int xcol,
                                         const float *range6, uint8_t *keep)
{
    int64_t kept = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float x = points[i * stride + xcol], y = points[i * stride + xcol + 1];
        const int k = (x >= range6[0]) && (x <= range6[3]) && (y >= range6[1]) && (y <= range6[4]);
        keep[i] = (uint8_t)k;
        kept += k;
    }
    return kept;
}

/* ------------------------------------------------------------------------- */
/* data_processor.py:135-136:  grid = round((range[3:6]-range[0:3]) / vsize)  */
/* range is np.float32 (subtracted in fp32), VOXEL_SIZE a python list -> the   */
/* divide happens in float64; np.round is round-half-even.                     */
ORC_API void orc_grid_size(const float *range6, const double *vsize3, int32_t *grid3)
{
    for (int j = 0; j < 3; ++j) {
        const float span = range6[3 + j] - range6[j];
        grid3[j] = (int32_t)nearbyint((double)span / vsize3[j]);
    }
}

/* ------------------------------------------------------------------------- */
/* a2: spconv v2 Point2VoxelCPU.point_to_voxel (called at                      */
/* data_processor.py:55).  One frame.                                          */
/*   points   [n, stride] fp32, x/y/z at columns xcol..xcol+2, the F features   */
/*            copied into voxels are columns xcol..xcol+F-1                     */
/*   lookup   caller scratch int32[nz*ny*nx], must be all -1 on entry, is       */
/*            restored to -1 on exit (as spconv resets its grid)                */
/*   voxels   [max_voxels, P, F] fp32, zero filled here for the rows returned   */
/*   coords   [max_voxels, 3] int32 (z, y, x)                                   */
/*   num      [max_voxels] int32                                                */
/*   point_pillar (optional, may be NULL) [n] int32: pillar id the point was    */
/*            stored in, -1 if dropped (out of range, overflow or truncated)    */
/* returns the number of pillars.                                              */
/* overflow_break = 0: spconv 2.x (Point2VoxelCPU3d): a point that would open pillar number max_voxels + 1 is skipped      */
/*                      (`continue`), later points still fill the pillars that exist.                                    */
/* overflow_break = 1: spconv 1.x (VoxelGenerator / points_to_voxel, which data_processor.py:16-26 prefers when it       */
/*                      imports): the loop `break`s at that point -- every later point of the frame is dropped.          */
ORC_API int32_t orc_voxelize_mode(const float *points, int64_t n, int stride, int xcol, int F,
                                  const float *range6, const float *vsize3, const int32_t *grid3,
                                  int P, int max_voxels, int32_t *lookup,
                                  float *voxels, int32_t *coords, int32_t *num, int32_t *point_pillar, int overflow_break);
ORC_API int32_t orc_voxelize(const float *points, int64_t n, int stride, int xcol, int F,
                             const float *range6, const float *vsize3, const int32_t *grid3,
                             int P, int max_voxels, int32_t *lookup,
                             float *voxels, int32_t *coords, int32_t *num, int32_t *point_pillar)
{
    return orc_voxelize_mode(points, n, stride, xcol, F, range6, vsize3, grid3, P, max_voxels, lookup, voxels, coords, num,
                             point_pillar, 0);
}
ORC_API int32_t orc_voxelize_mode(const float *points, int64_t n, int stride, int xcol, int F,
                                  const float *range6, const float *vsize3, const int32_t *grid3,
                                  int P, int max_voxels, int32_t *lookup,
                                  float *voxels, int32_t *coords, int32_t *num, int32_t *point_pillar, int overflow_break)
{
    const int nx = grid3[0], ny = grid3[1];
    int32_t voxel_num = 0;
    int stop = 0;
    for (int64_t i = 0; i < n; ++i) {
        const float *pt = points + i * stride + xcol;
        int c[3];
        int ok = 1;
        for (int j = 0; j < 3; ++j) {
            /* fp32 subtract, fp32 true divide, floor; upper bound exclusive */
            const float q = (pt[j] - range6[j]) / vsize3[j];
            const float f = floorf(q);
            if (!(f >= 0.0f) || !(f < (float)grid3[j])) { ok = 0; break; }
            c[j] = (int)f;
        }
        if (point_pillar) point_pillar[i] = -1;
        if (!ok) continue;
        if (stop) continue;                          /* spconv 1.x broke out of the loop earlier */
        const int64_t cell = ((int64_t)c[2] * ny + c[1]) * nx + c[0];
        int32_t v = lookup[cell];
        if (v == -1) {
            if (voxel_num >= max_voxels) {           /* later NEW pillars are dropped ... */
                if (overflow_break) stop = 1;        /* ... and with spconv 1.x everything after this point */
                continue;
            }
            v = voxel_num++;
            lookup[cell] = v;
            coords[3 * v + 0] = c[2];
            coords[3 * v + 1] = c[1];
            coords[3 * v + 2] = c[0];
            num[v] = 0;
            memset(voxels + (size_t)v * P * F, 0, sizeof(float) * (size_t)P * F);
        }
        if (num[v] < P) {
            memcpy(voxels + ((size_t)v * P + num[v]) * F, pt, sizeof(float) * (size_t)F);
            num[v] += 1;
            if (point_pillar) point_pillar[i] = v;
        }
    }
    for (int32_t v = 0; v < voxel_num; ++v) {
        const int64_t cell = ((int64_t)coords[3 * v] * ny + coords[3 * v + 1]) * nx + coords[3 * v + 2];
        lookup[cell] = -1;
    }
    return voxel_num;
}

/* ------------------------------------------------------------------------- */
/* a5-a8: PillarVFE.forward, eval mode, single last-layer PFN                  */
/* (pillar_vfe.py:94-123 with PFNLayer.forward :29-49).                        */
/* Operation order is the one that reproduces torch CPU bit for bit            */
/* (tests/golden): see comments at each step.                                  */
typedef struct {
    int32_t F;                 /* point features in voxels[..., :F]                        */
    int32_t P;                 /* slots per pillar                                         */
    int32_t C;                 /* PFN output channels                                      */
    int32_t use_absolute_xyz;  /* USE_ABSLOTE_XYZ (pillar_vfe.py:58,105-108)               */
    int32_t with_distance;     /* WITH_DISTANCE (:61,110-112)                              */
    int32_t use_norm;          /* USE_NORM: Linear(no bias)+BN, else Linear(bias) (:21-25) */
    float vx, vy, vz;          /* voxel size as fp32 scalars (:76-78, python float -> f32) */
    float x_off, y_off, z_off; /* vx/2+xmin ... evaluated by the caller as python does (:79-81) */
    float eps;                 /* 1e-3 (:23)                                               */
} orc_vfe_cfg;

static inline int orc_cin(const orc_vfe_cfg *g)
{
    return (g->use_absolute_xyz ? g->F : g->F - 3) + 6 + (g->with_distance ? 1 : 0);
}

/* torch CPU sum(dim=1) over [M,P,3]: four interleaved partial sums over the
 * leading 4*floor(P/4) slots, the tail added into partial 0, then
 * ((a0+a1)+a2)+a3.  (pillar_vfe.py:97) */
static inline float orc_slot_sum(const float *v, int P, int stride)
{
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    const int P4 = (P / 4) * 4;
    for (int s = 0; s < P4; ++s) a[s & 3] = a[s & 3] + v[(size_t)s * stride];
    for (int s = P4; s < P; ++s) a[0] = a[0] + v[(size_t)s * stride];
    return ((a[0] + a[1]) + a[2]) + a[3];
}

/* one decorated row -> Cin features (pillar_vfe.py:97-118) */
static inline void orc_decorate(const orc_vfe_cfg *g, const float *pt, const float mean[3],
                                const float centre[3], float *feat)
{
    int k = 0;
    for (int j = g->use_absolute_xyz ? 0 : 3; j < g->F; ++j) feat[k++] = pt[j];
    for (int j = 0; j < 3; ++j) feat[k++] = pt[j] - mean[j];
    for (int j = 0; j < 3; ++j) feat[k++] = pt[j] - centre[j];
    if (g->with_distance) {
        /* torch.norm(xyz, 2, dim=2) on the CPU: sqrt(fma(z,z, fma(y,y, x*x)))  (probed, tests/golden) */
        feat[k++] = sqrtf(fmaf(pt[2], pt[2], fmaf(pt[1], pt[1], pt[0] * pt[0])));
    }
}

/*   voxels [M,P,F], coords [M,4] (b,z,y,x) as fp32 (the reference moves them   */
/*   to the GPU as float, pcdet/models/__init__.py:36), num [M] fp32             */
/*   W [C,Cin]; bias [C] (use_norm=0) ; gamma,beta,rmean,rvar [C] (use_norm=1)   */
/*   out [M,C]                                                                   */
typedef struct {
    const orc_vfe_cfg *g;
    const float *voxels, *coords, *num, *W, *bias, *gamma, *beta, *rmean, *invstd, *padv;
    float *out;
} orc_vfe_ctx;

static void orc_vfe_range(int64_t lo, int64_t hi, void *p)
{
    const orc_vfe_ctx *x = (const orc_vfe_ctx *)p;
    const orc_vfe_cfg *g = x->g;
    const int F = g->F, P = g->P, C = g->C, Cin = orc_cin(g);
    for (int64_t m = lo; m < hi; ++m) {
        const float *vox = x->voxels + (size_t)m * P * F;
        const float cnt_f = x->num[m];
        int cnt = (int)cnt_f;             /* mask = num.int() > arange(P)  (:87-91) */
        if (cnt > P) cnt = P;
        float mean[3], centre[3], feat[64];
        for (int j = 0; j < 3; ++j) mean[j] = orc_slot_sum(vox + j, P, F) / cnt_f;
        /* two roundings, no FMA: fl(fl(c*v)+off)  (:101-103) */
        centre[0] = x->coords[4 * m + 3] * g->vx + g->x_off;
        centre[1] = x->coords[4 * m + 2] * g->vy + g->y_off;
        centre[2] = x->coords[4 * m + 1] * g->vz + g->z_off;
        float *o = x->out + (size_t)m * C;
        /* a zeroed (padded) row still runs through BN+ReLU and joins the max (:37-42) */
        for (int c = 0; c < C; ++c) o[c] = (cnt < P) ? x->padv[c] : 0.0f;   /* ReLU output >= 0 */
        for (int s = 0; s < cnt; ++s) {
            orc_decorate(g, vox + (size_t)s * F, mean, centre, feat);
            for (int c = 0; c < C; ++c) {
                const float *w = x->W + (size_t)c * Cin;
                float acc = 0.0f;
                for (int k = 0; k < Cin; ++k) acc = fmaf(feat[k], w[k], acc);   /* sequential FMA, k order */
                float y;
                if (g->use_norm) y = ((acc - x->rmean[c]) * x->invstd[c]) * x->gamma[c] + x->beta[c];   /* 4 rounded ops */
                else             y = acc + x->bias[c];
                y = (y > 0.f || y != y) ? y : 0.f;      /* relu; NaN propagates as in torch */
                if (y > o[c] || y != y) o[c] = y;       /* max; NaN propagates as in torch  */
            }
        }
    }
}

ORC_API void orc_pillar_vfe(const orc_vfe_cfg *g, int64_t M,
                            const float *voxels, const float *coords, const float *num,
                            const float *W, const float *bias,
                            const float *gamma, const float *beta,
                            const float *rmean, const float *rvar, const float *invstd_override,
                            float *out)
{
    /* invstd_override (may be NULL): torch evaluates 1/sqrt(var+eps) on the CPU through MKL VML's
     * vsSqrt, which is not correctly rounded (observed 1 ulp off on ~1 channel in 64).  The oracle's
     * own value is the IEEE one; the golden test passes torch's vector here to pin everything else
     * bit for bit. */
    const int C = g->C;
    float *invstd = (float *)malloc(sizeof(float) * (size_t)C);
    float *padv = (float *)malloc(sizeof(float) * (size_t)C);
    for (int c = 0; c < C; ++c) {
        float y;
        if (g->use_norm) {
            invstd[c] = invstd_override ? invstd_override[c] : 1.0f / sqrtf(rvar[c] + g->eps);
            y = ((0.0f - rmean[c]) * invstd[c]) * gamma[c] + beta[c];
        } else {
            invstd[c] = 0.f;
            y = 0.0f + bias[c];
        }
        padv[c] = y > 0.f ? y : 0.f;
    }
    orc_vfe_ctx x = {g, voxels, coords, num, W, bias, gamma, beta, rmean, invstd, padv, out};
    orc_parallel_for(M, orc_vfe_range, &x);
    free(invstd);
    free(padv);
}

/* ------------------------------------------------------------------------- */
/* a10: PointPillarScatter.forward (pointpillar_scatter.py:14-41).             */
/*   canvas [B, C, ny, nx] is zeroed here; idx = z + y*nx + x (nz == 1).        */
/*   Duplicate coords: the last pillar in list order wins (index_put order on   */
/*   the CPU); the voxelizer never produces duplicates.                         */
typedef struct { float *canvas; size_t chunk, total; } orc_zero_ctx;
static void orc_zero_range(int64_t lo, int64_t hi, void *p)
{
    const orc_zero_ctx *z = (const orc_zero_ctx *)p;
    size_t a = (size_t)lo * z->chunk, b = (size_t)hi * z->chunk;
    if (b > z->total) b = z->total;
    if (a < b) memset(z->canvas + a, 0, sizeof(float) * (b - a));
}

ORC_API void orc_pointpillar_scatter(int64_t M, int C, int B, int ny, int nx,
                                     const float *pillar_features, const float *coords, float *canvas)
{
    const size_t plane = (size_t)ny * nx;
    orc_zero_ctx z = {canvas, (size_t)1 << 20, (size_t)B * C * plane};
    orc_parallel_for((int64_t)((z.total + z.chunk - 1) / z.chunk), orc_zero_range, &z);
    for (int64_t m = 0; m < M; ++m) {
        const int b = (int)coords[4 * m];
        if (b < 0 || b >= B) continue;
        /* index arithmetic in fp32 like the reference (:31-32); exact below 2^24 */
        const float fidx = coords[4 * m + 1] + coords[4 * m + 2] * (float)nx + coords[4 * m + 3];
        const size_t idx = (size_t)(int64_t)fidx;
        float *dst = canvas + (size_t)b * C * plane + idx;
        const float *src = pillar_features + (size_t)m * C;
        for (int c = 0; c < C; ++c) dst[(size_t)c * plane] = src[c];
    }
}

/* ------------------------------------------------------------------------- */
/* The whole path for a batch, as the reference runs it when Path A is          */
/* configured: per-frame voxelize (DataLoader worker side; frames run on         */
/* separate threads here as they run in separate worker processes there,         */
/* tools/train.py:27), collate (dataset.py:232-244), PillarVFE,                  */
/* PointPillarScatter.  Used by tests as the end-to-end checker and by bench.py  */
/* as the timed CPU baseline.                                                    */
/*   points [n_total, stride]; frame_offsets [B+1]; F features from column xcol  */
/*   outputs sized by the caller with cap = sum_b min(n_b, max_voxels):          */
/*   voxels_out [cap,P,F], coords_out [cap,4] int32 (b,z,y,x), num_out [cap]     */
/*   int32, feat_out [cap,C], canvas [B,C,ny,nx] (may be NULL),                  */
/*   frame_pillars [B] (may be NULL).  returns total pillars M.                  */
typedef struct {
    const float *points; const int64_t *frame_offsets; int stride, xcol, F, P, max_voxels;
    const float *range6, *vsize3; const int32_t *grid3;
    float *voxels; int32_t *coords, *num, *count; const int64_t *cap_off;
} orc_vox_ctx;

static void orc_vox_range(int64_t lo, int64_t hi, void *p)
{
    const orc_vox_ctx *x = (const orc_vox_ctx *)p;
    const size_t cells = (size_t)x->grid3[0] * x->grid3[1] * x->grid3[2];
    int32_t *lookup = (int32_t *)malloc(sizeof(int32_t) * cells);
    for (size_t i = 0; i < cells; ++i) lookup[i] = -1;
    int32_t *c3 = (int32_t *)malloc(sizeof(int32_t) * 3 * (size_t)(x->max_voxels > 0 ? x->max_voxels : 1));
    for (int64_t b = lo; b < hi; ++b) {
        const int64_t n = x->frame_offsets[b + 1] - x->frame_offsets[b];
        const int64_t o = x->cap_off[b];
        const int32_t m = orc_voxelize(x->points + x->frame_offsets[b] * x->stride, n, x->stride, x->xcol,
                                       x->F, x->range6, x->vsize3, x->grid3, x->P, x->max_voxels, lookup,
                                       x->voxels + (size_t)o * x->P * x->F, c3, x->num + o, NULL);
        for (int32_t v = 0; v < m; ++v) {
            x->coords[4 * (o + v) + 0] = (int32_t)b;
            x->coords[4 * (o + v) + 1] = c3[3 * v + 0];
            x->coords[4 * (o + v) + 2] = c3[3 * v + 1];
            x->coords[4 * (o + v) + 3] = c3[3 * v + 2];
        }
        x->count[b] = m;
    }
    free(lookup);
    free(c3);
}

ORC_API int64_t orc_voxelize_batch(const float *points, const int64_t *frame_offsets, int B,
                                   int stride, int xcol, int F, int P,
                                   const float *range6, const float *vsize3, const int32_t *grid3,
                                   int max_voxels,
                                   float *voxels_out, int32_t *coords_out, int32_t *num_out,
                                   int32_t *frame_pillars)
{
    int64_t *cap_off = (int64_t *)malloc(sizeof(int64_t) * (size_t)(B + 1));
    int32_t *count = (int32_t *)malloc(sizeof(int32_t) * (size_t)(B > 0 ? B : 1));
    cap_off[0] = 0;
    for (int b = 0; b < B; ++b) {
        int64_t n = frame_offsets[b + 1] - frame_offsets[b];
        cap_off[b + 1] = cap_off[b] + (n < max_voxels ? n : max_voxels);
    }
    orc_vox_ctx x = {points, frame_offsets, stride, xcol, F, P, max_voxels, range6, vsize3, grid3,
                     voxels_out, coords_out, num_out, count, cap_off};
    orc_parallel_for(B, orc_vox_range, &x);
    /* collate: np.concatenate of the per-frame arrays (dataset.py:232-244) */
    int64_t M = 0;
    for (int b = 0; b < B; ++b) {
        const int64_t o = cap_off[b], m = count[b];
        if (o != M && m > 0) {
            memmove(voxels_out + (size_t)M * P * F, voxels_out + (size_t)o * P * F, sizeof(float) * (size_t)m * P * F);
            memmove(coords_out + 4 * M, coords_out + 4 * o, sizeof(int32_t) * 4 * (size_t)m);
            memmove(num_out + M, num_out + o, sizeof(int32_t) * (size_t)m);
        }
        if (frame_pillars) frame_pillars[b] = (int32_t)m;
        M += m;
    }
    free(cap_off);
    free(count);
    return M;
}

ORC_API int64_t orc_points_to_bev(const float *points, const int64_t *frame_offsets, int B,
                                  int stride, int xcol, const orc_vfe_cfg *g,
                                  const float *range6, const float *vsize3, const int32_t *grid3,
                                  int max_voxels,
                                  const float *W, const float *bias,
                                  const float *gamma, const float *beta,
                                  const float *rmean, const float *rvar,
                                  float *voxels_out, int32_t *coords_out, int32_t *num_out,
                                  float *feat_out, float *canvas, int32_t *frame_pillars)
{
    const int64_t M = orc_voxelize_batch(points, frame_offsets, B, stride, xcol, g->F, g->P, range6, vsize3,
                                         grid3, max_voxels, voxels_out, coords_out, num_out, frame_pillars);
    /* load_data_to_gpu turns coords and counts into fp32 (pcdet/models/__init__.py:36) */
    float *coords_f = (float *)malloc(sizeof(float) * 4 * (size_t)(M > 0 ? M : 1));
    float *num_f = (float *)malloc(sizeof(float) * (size_t)(M > 0 ? M : 1));
    for (int64_t i = 0; i < 4 * M; ++i) coords_f[i] = (float)coords_out[i];
    for (int64_t i = 0; i < M; ++i) num_f[i] = (float)num_out[i];
    orc_pillar_vfe(g, M, voxels_out, coords_f, num_f, W, bias, gamma, beta, rmean, rvar, NULL, feat_out);
    if (canvas) orc_pointpillar_scatter(M, g->C, B, grid3[1], grid3[0], feat_out, coords_f, canvas);
    free(coords_f);
    free(num_f);
    return M;
}
