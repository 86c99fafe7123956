/*
 * hgsfusion_b200.h -- C ABI of the B200-native radar pillarization hot path.
 *
 * One shared library (hgsfusion_b200/libhgsfusion_b200.so), sm_100a only.  Every entry
 * point takes raw DEVICE pointers, sizes, a geometry struct and a cudaStream_t; writes only
 * caller-provided buffers; never allocates, never synchronises the host, never calls exit();
 * and returns an int status: 0 ok, <0 invalid use (HGSF_ERR_*), >0 a cudaError_t.  Safe to
 * call concurrently on different streams / devices with different workspaces.
 *
 * Each entry names the reference interface it replaces (file:line under the HGSFusion repo).
 * The reference binds its native code through pybind11 (pcdet/ops/pillar_ops/src/pillar_api.cpp:10-22:
 * `int f(at::Tensor...)`, outputs pre-allocated by the Python caller, legacy default stream,
 * exit(-1) on kernel failure); this header is the torch-free equivalent of that surface.
 */
#ifndef HGSFUSION_B200_H
#define HGSFUSION_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HGSF_ABI_VERSION 1

#if defined(__GNUC__)
#define HGSF_API __attribute__((visibility("default")))
#else
#define HGSF_API
#endif

/* status codes (<0); positive values are cudaError_t */
#define HGSF_OK                 0
#define HGSF_ERR_INVALID_ARG   (-1)   /* null pointer, negative size, inconsistent geometry          */
#define HGSF_ERR_UNSUPPORTED   (-2)   /* a configuration outside the compiled kernel set             */
#define HGSF_ERR_WORKSPACE     (-3)   /* workspace too small / misaligned                            */
#define HGSF_ERR_DRIVER        (-4)   /* cuTensorMapEncodeTiled unavailable or failed                */

typedef void *hgsf_stream_t;          /* a cudaStream_t */

/* Voxel grid.  Mirrors what DataProcessor / PillarVFE derive from the YAML:
 *   grid        = round((range[3:6]-range[0:3]) / VOXEL_SIZE)   data_processor.py:135-136
 *   voxel_size  = VOXEL_SIZE as fp32 scalars                    pillar_vfe.py:76-78
 *   centre_off  = voxel/2 + range_min, evaluated by the host exactly as pillar_vfe.py:79-81 does */
typedef struct hgsf_geometry {
    float   pc_range[6];      /* xmin ymin zmin xmax ymax zmax */
    float   voxel_size[3];    /* vx vy vz */
    int32_t grid[3];          /* nx ny nz */
    float   centre_off[3];    /* x_offset y_offset z_offset */
} hgsf_geometry;

/* A batch of points as DatasetTemplate.collate_batch lays it out (pcdet/datasets/dataset.py:237-244):
 * [n, stride] fp32 rows, frames contiguous and in batch order.  Either `frame_offsets` (device
 * int32[batch_size+1]) or `batch_col` (column holding the batch index as a float, normally 0) says
 * where frames start; when both are given frame_offsets wins. */
typedef struct hgsf_points {
    const float   *data;          /* device */
    int64_t        n;             /* rows */
    int32_t        stride;        /* floats per row */
    int32_t        xyz_col;       /* column of x (y, z follow); the F features are columns xyz_col .. xyz_col+F-1 */
    int32_t        num_features;  /* F */
    int32_t        batch_col;     /* -1 if absent */
    const int32_t *frame_offsets; /* device int32[batch_size+1] or NULL */
    int32_t        batch_size;    /* B */
    int32_t        flags;         /* HGSF_POINTS_* below, 0 = the default path */
} hgsf_points;

/* hgsf_points.flags */
#define HGSF_POINTS_SPCONV1_BREAK 1   /* spconv 1.x overflow semantics (VoxelGenerator.generate, which VoxelGeneratorWrapper prefers
                                       * when it imports: data_processor.py:16-26,47-52): the voxelization loop STOPS at the first
                                       * point that would open pillar number max_voxels + 1, dropping every later point of the frame;
                                       * default = spconv 2.x (Point2VoxelCPU3d): only new pillars are refused                       */
#define HGSF_POINTS_FLIP_X        2   /* voxelize (-x, y, z, ...): DataProcessor.double_flip's x flip (data_processor.py:116-130,
                                       * 161-178); the stored features carry the flipped sign, as the reference's copies do          */
#define HGSF_POINTS_FLIP_Y        4   /* likewise (x, -y, z, ...)                                                                     */

/* The single (last) PFN layer of PillarVFE (pillar_vfe.py:8-49,63-74), eval mode.
 * weight = pfn_layers.0.linear.weight [C, Cin]; USE_NORM=True: BatchNorm1d(eps=1e-3) running stats;
 * USE_NORM=False: bias = pfn_layers.0.linear.bias and the four bn_* are NULL. */
typedef struct hgsf_pfn {
    const float *weight;      /* device [C, Cin] row-major */
    const float *bias;        /* device [C] or NULL */
    const float *bn_weight;   /* device [C] or NULL */
    const float *bn_bias;
    const float *bn_mean;
    const float *bn_var;
    float        bn_eps;
    int32_t      in_channels;       /* Cin = (F or F-3) + 6 (+1) */
    int32_t      out_channels;      /* C */
    int32_t      use_absolute_xyz;  /* USE_ABSLOTE_XYZ */
    int32_t      with_distance;     /* WITH_DISTANCE */
} hgsf_pfn;

/* Outputs of the pillar path.  All device pointers; any of voxels / pillar_features /
 * spatial_features may be NULL to skip that output.  Rows are in the reference's first-seen
 * order, frames concatenated (collate_batch).  `pillar_capacity` rows are available in each
 * per-pillar buffer; the number actually written is num_pillars[0] (<= capacity, see
 * hgsf_pillar_capacity). */
typedef struct hgsf_pillar_outputs {
    int32_t *voxel_coords;       /* [cap, 4] (b, z, y, x)                                   */
    int32_t *voxel_num_points;   /* [cap]                                                   */
    float   *voxels;             /* [cap, P, F] zero padded, or NULL                        */
    float   *pillar_features;    /* [cap, C] or NULL                                        */
    float   *spatial_features;   /* [B, C*nz, ny, nx] or NULL (requires nz == 1)            */
    int32_t *num_pillars;        /* [1 + B]: total, then per frame                          */
    int64_t  pillar_capacity;
} hgsf_pillar_outputs;

HGSF_API int hgsf_abi_version(void);
HGSF_API const char *hgsf_status_string(int status);

/* Upper bound on the number of pillars: min(n, B*min(max_voxels, nx*ny*nz)). */
HGSF_API int64_t hgsf_pillar_capacity(const hgsf_geometry *geom, int64_t n_points, int32_t batch_size, int32_t max_voxels);

/* Bytes of device workspace hgsf_pillarize / hgsf_points_to_bev need (256-byte aligned base). */
HGSF_API int hgsf_workspace_size(const hgsf_geometry *geom, int64_t n_points, int32_t batch_size, int32_t num_features,
                        size_t *bytes);

/* points -> voxels, voxel_coords, voxel_num_points.
 * Replaces DataProcessor.transform_points_to_voxels + VoxelGeneratorWrapper.generate
 * (pcdet/datasets/processor/data_processor.py:16-61,133-183; spconv Point2VoxelCPU3d.point_to_voxel)
 * and the voxel part of collate_batch (pcdet/datasets/dataset.py:232-244): first-seen pillar order,
 * first `max_points_per_voxel` points per pillar in input order, at most `max_voxels` pillars per frame.
 * Uses out->voxel_coords, voxel_num_points, voxels (optional), num_pillars. */
HGSF_API int hgsf_pillarize(const hgsf_geometry *geom, const hgsf_points *points,
                   int32_t max_points_per_voxel, int32_t max_voxels,
                   void *workspace, size_t workspace_bytes,
                   const hgsf_pillar_outputs *out, hgsf_stream_t stream);

/* points -> (voxel_coords, voxel_num_points, [voxels]) + pillar_features + spatial_features in one pass.
 * Replaces, fused: transform_points_to_voxels (above), PillarVFE.forward
 * (pcdet/models/backbones_3d/vfe/pillar_vfe.py:94-123 incl. PFNLayer.forward :29-49) and
 * PointPillarScatter.forward (pcdet/models/backbones_2d/map_to_bev/pointpillar_scatter.py:14-41). */
HGSF_API int hgsf_points_to_bev(const hgsf_geometry *geom, const hgsf_points *points, const hgsf_pfn *pfn,
                       int32_t max_points_per_voxel, int32_t max_voxels,
                       void *workspace, size_t workspace_bytes,
                       const hgsf_pillar_outputs *out, hgsf_stream_t stream);

/* batch_dict contract mode: voxels [M,P,F], voxel_coords [M,4], voxel_num_points [M] -> pillar_features [M,C].
 * Replaces PillarVFE.forward (pillar_vfe.py:94-123).  coords / num_points arrive as float32 in the
 * reference (pcdet/models/__init__.py:36); `coords_are_float` / `num_are_float` select fp32 or int32. */
HGSF_API int hgsf_pillar_vfe(const hgsf_geometry *geom, const hgsf_pfn *pfn,
                    const float *voxels, const void *voxel_coords, const void *voxel_num_points,
                    int32_t coords_are_float, int32_t num_are_float,
                    int64_t num_pillars, int32_t max_points_per_voxel, int32_t num_features,
                    float *pillar_features, hgsf_stream_t stream);

/* The same for a STACKED PFN of two layers (model_cfg NUM_FILTERS = [2H, C1]; PillarVFE.__init__ pillar_vfe.py:63-74 builds
 * PFNLayer(Cin, 2H, last_layer=False) -> Linear(Cin, H) and PFNLayer(2H, C1, last_layer=True); PFNLayer.forward :29-49 concatenates
 * each slot's features with the pillar's max).  pfn0->out_channels = H in {32, 64}, pfn1->in_channels = 2H,
 * pfn1->out_channels = C1 in {32, 64, 128}; use_absolute_xyz / with_distance are read from pfn0.  Eval mode. */
HGSF_API int hgsf_pillar_vfe_stacked(const hgsf_geometry *geom, const hgsf_pfn *pfn0, const hgsf_pfn *pfn1,
                            const float *voxels, const void *voxel_coords, const void *voxel_num_points,
                            int32_t coords_are_float, int32_t num_are_float,
                            int64_t num_pillars, int32_t max_points_per_voxel, int32_t num_features,
                            float *pillar_features, hgsf_stream_t stream);

/* Bytes of workspace hgsf_pointpillar_scatter needs. */
HGSF_API int hgsf_scatter_workspace_size(const hgsf_geometry *geom, int32_t batch_size, size_t *bytes);

/* pillar_features [M,C] + voxel_coords [M,4] -> spatial_features [B,C,ny,nx] (zero elsewhere).
 * Replaces PointPillarScatter.forward (pointpillar_scatter.py:14-41).  Rows whose batch index is
 * outside [0,B) are ignored; duplicate cells resolve to the last row (CPU index_put order). */
HGSF_API int hgsf_pointpillar_scatter(const hgsf_geometry *geom, const float *pillar_features,
                             const void *voxel_coords, int32_t coords_are_float,
                             int64_t num_pillars, int32_t channels, int32_t batch_size,
                             void *workspace, size_t workspace_bytes,
                             float *spatial_features, hgsf_stream_t stream);

/* Number of kernels (memsets excluded) the last call on this host thread enqueued (bench.py's gpu_launches). */
HGSF_API int hgsf_last_launch_count(void);

/* (Measurement hooks that synchronise -- bench.py's roofline leg -- live in hgsfusion_b200_debug.h, not in the product ABI.) */

/* ---------------------------------------------------------------------------------------------------------------
 * Path B -- the PillarNet reader the shipped HGSFusion YAMLs run.  These replace the reference's pybind module
 * `pillar_cuda` (pcdet/ops/pillar_ops/src/pillar_api.cpp:10-22) and the Python glue around it.
 * --------------------------------------------------------------------------------------------------------------- */

/* Bytes of workspace hgsf_pillarnet_indices needs. */
HGSF_API int hgsf_pillarnet_workspace_size(int64_t n_points, size_t *bytes);

/* gen_indice_pairs + flatten_indices (pcdet/ops/pillar_ops/pillar_utils.py:84-132, group_utils.py:12-31; kernels
 * create_pillar_indices_stack / create_pillar_indices / create_pillar_indice_pairs_stack / flatten_indice_pairs,
 * pillar_ops_gpu.cu:13-117, group_ops_gpu.cu:9-24) in ONE launch, without the reference's two cumsum + .item() syncs.
 *   xyz [N,3] fp32 relative coordinates, xyz_batch_cnt [B] int32; H = Ny, W = Nx
 *   pillar_bev_indices [B,H,W] int32 (pillar id per cell, -1 none); pillars [>= min(N, B*H*W), 3] int32 (b, y, x),
 *   raster order; indice_pairs [N] int32 (pillar id per point or -1; may be NULL); point_idx / pillar_idx [N] int32
 *   (the points that have a pillar, in input order, and their pillar ids); counts [2] int32 = {M, L} on the device. */
HGSF_API int hgsf_pillarnet_indices(float bev_size, const float *xyz, const int32_t *xyz_batch_cnt, int64_t n_points,
                                    int32_t batch_size, int32_t H, int32_t W, int32_t *pillar_bev_indices,
                                    int32_t *pillars, int32_t *indice_pairs, int32_t *point_idx, int32_t *pillar_idx,
                                    int32_t *counts, void *workspace, size_t workspace_bytes, hgsf_stream_t stream);

/* out[l, :] = features[set_indices[l], :]   (gather_feature_wrapper, group_ops_gpu.cu:42-55) */
HGSF_API int hgsf_gather_feature(const int32_t *set_indices, const float *features, int64_t L, int32_t C, float *out,
                                 hgsf_stream_t stream);
/* grad_features[set_indices[l], :] += grad_out[l, :]; grad_features pre-zeroed by the caller
 * (gather_feature_grad_wrapper, group_ops_gpu.cu:57-70) */
HGSF_API int hgsf_gather_feature_grad(const int32_t *set_indices, const float *grad_out, int64_t L, int32_t C,
                                      float *grad_features, hgsf_stream_t stream);

/* out[c, index[p]] = max(0, max_p src[c, p]); arg[c, m] = a flat id c*L+p whose src is within 1e-5 of out[c, m], -1 if
 * none (scatter_max_wrapper, scatter_ops_gpu.cu:13-46).  src [C,L], out/arg [C,M]; both are initialised here.
 * arg may be NULL (forward only). */
HGSF_API int hgsf_scatter_max(const int32_t *index, const float *src, int32_t C, int64_t L, int64_t M, int32_t *arg,
                              float *out, hgsf_stream_t stream);
/* grad_src.flat[arg[c, m]] = grad_out[c, m] where arg >= 0; grad_src [C,L] pre-zeroed by the caller
 * (scatter_max_grad_wrapper, scatter_ops_gpu.cu:48-58) */
HGSF_API int hgsf_scatter_max_grad(const int32_t *arg, const float *grad_out, int32_t C, int64_t M, float *grad_src,
                                   hgsf_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Path B input prep: PillarNet.forward's per-frame split (pcdet/models/backbones_3d/vfe/pillarnet.py:51-58) and
 * DynamicPillarFeatureNet.forward's feature encoding (pillarnet_modules/dynamic_pillar_encoder.py:55-118) in ONE launch
 * instead of the reference's Python loop of boolean-mask assigns per frame.
 *   points [L, 1+Fin] fp32 collated points, column 0 = frame index (pcdet/datasets/dataset.py:237-244)
 *   encoding  HGSF_ENCODE_SPLIT : USE_VIRTUAL_POINT + ENCODING_TYPE 'split' (:65-86): pt_features[:, :3] = xyz;
 *                                 rows whose column Fin-2 is >= 0.5 (real) copy columns 3..3+n into 3..3+n, the others
 *                                 (virtual) into 3+n..3+2n; the last two columns are the two flag columns; everything
 *                                 else 0.  VoD: Fin 17, n 12, Fout 29 (:72-73); TJ4D: Fin 18, n 13, Fout 31 (:75-76)
 *             HGSF_ENCODE_COPY  : 'mixed' or no virtual points (:87-91, :100-104): pt_features = the Fin columns
 *             HGSF_ENCODE_DIRECT: 'direct' (:92-96): pt_features = the first Fin-2 columns
 *   pc_min[3] range minimum; xyz [L,3] = points[:, 1:4] - pc_min (one fp32 subtract each, :46-53)
 *   pt_features [L, Fout]; xyz_batch_cnt [B] int32 = rows per frame; info [2] int32 on the device:
 *   info[1] = rows kept (rows whose frame index is not an integer in [0, B) match no `points[:,0] == i` mask and are
 *   dropped, as in the reference); info[0] bit 0 set = rows were not grouped by ascending frame (or a dropped row
 *   precedes a kept one), so the emitted order is not the reference's: call again with `order` [L] int32 = a stable
 *   ordering of the rows by frame with the dropped rows last (NULL = input order; the collated batch is always grouped).
 *   Rows [info[1], L) of the outputs are not written. */
#define HGSF_ENCODE_SPLIT  0
#define HGSF_ENCODE_COPY   1
#define HGSF_ENCODE_DIRECT 2
HGSF_API int hgsf_split_encode(const float *points, int64_t n_rows, int32_t Fin, int32_t Fout, int32_t n_split,
                               int32_t batch_size, int32_t encoding, const float *pc_min, const int32_t *order,
                               float *xyz, float *pt_features, int32_t *xyz_batch_cnt, int32_t *info,
                               hgsf_stream_t stream);

/* The PillarNet reader's forward in ONE launch, eval mode: PillarQueryAndGroup's gathers and centre offsets
 * (pcdet/ops/pillar_ops/pillar_utils.py:31-54), the shared MLP Linear(no bias) + BatchNorm1d + ReLU
 * (pillar_modules.py:19-26,76) and scatter_max (scatter_ops_gpu.cu:13-25), without materialising group_features, the
 * MLP output or its transpose.  point_idx / pillar_idx / pillars come from hgsf_pillarnet_indices; pfn->weight is
 * shared_mlps.0.weight [32, Cf+6] (out_channels must be 32, in_channels Cf+6 <= 40), bn_* are shared_mlps.1.*;
 * z_center = (zmax + zmin) / 2 of the range (pillar_utils.py:28); pillar_features [M, 32] is initialised here. */
HGSF_API int hgsf_pillarnet_reader(const float *xyz, const float *pt_features, int32_t num_point_features,
                                   const int32_t *point_idx, const int32_t *pillar_idx, int64_t L,
                                   const int32_t *pillars, int64_t M, float bev_size, float z_center,
                                   const hgsf_pfn *pfn, float *pillar_features, hgsf_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Training through the path.  The reference trains PillarVFE with torch autograd over
 * PFNLayer.forward (pillar_vfe.py:29-49): BatchNorm1d in train mode normalises with the statistics of ALL M*P rows of a
 * channel (zero-padded rows included) and updates the running statistics (momentum 0.01, unbiased variance); backward
 * goes through max -> ReLU -> batch-norm -> Linear.  Three calls on the contract layout (voxels, voxel_coords,
 * voxel_num_points as in hgsf_pillar_vfe):
 *
 *   forward, train mode:  hgsf_pillar_vfe_batch_stats -> batch_mean / batch_var (and the running-stat update), then
 *                         hgsf_pillar_vfe with pfn.bn_mean = batch_mean, pfn.bn_var = batch_var
 *   backward:             hgsf_pillar_vfe_backward with the same pfn the forward used
 *
 * `stats` is device scratch of hgsf_train_stats_doubles(C, Cin) doubles, written by batch_stats and read by backward
 * (keep it between the two); `scratch` of backward is hgsf_train_scratch_doubles(C, Cin) doubles. */
HGSF_API int64_t hgsf_train_stats_doubles(int32_t out_channels, int32_t in_channels);
HGSF_API int64_t hgsf_train_scratch_doubles(int32_t out_channels, int32_t in_channels);

/* running_mean / running_var (device [C], may be NULL) are updated in place as torch.nn.BatchNorm1d does;
 * batch_mean / batch_var (device [C]) receive the biased batch statistics the forward normalises with. */
HGSF_API int hgsf_pillar_vfe_batch_stats(const hgsf_geometry *geom, const hgsf_pfn *pfn, const float *voxels,
                                         const void *voxel_coords, const void *voxel_num_points,
                                         int32_t coords_are_float, int32_t num_are_float, int64_t num_pillars,
                                         int32_t max_points_per_voxel, int32_t num_features, float momentum,
                                         float *running_mean, float *running_var, float *batch_mean, float *batch_var,
                                         double *stats, hgsf_stream_t stream);

/* grad_pillar_features [M,C] -> grad_weight [C,Cin], grad_bn_weight [C], grad_bn_bias [C] (USE_NORM False: grad_bn_bias
 * receives the Linear bias gradient, grad_bn_weight is ignored).  `stats` = the buffer hgsf_pillar_vfe_batch_stats
 * filled when the forward ran on batch statistics, NULL when it ran on the running statistics (eval-mode BN: the
 * statistics are constants). */
HGSF_API int hgsf_pillar_vfe_backward(const hgsf_geometry *geom, const hgsf_pfn *pfn, const float *voxels,
                                      const void *voxel_coords, const void *voxel_num_points,
                                      int32_t coords_are_float, int32_t num_are_float, int64_t num_pillars,
                                      int32_t max_points_per_voxel, int32_t num_features,
                                      const float *grad_pillar_features, const double *stats, double *scratch,
                                      float *grad_weight, float *grad_bn_weight, float *grad_bn_bias,
                                      hgsf_stream_t stream);

/* Train mode from points, without the contract tensors in between and without a host round trip (replaces the reference's
 * train-mode chain data_processor.py:140-178 -> pillar_vfe.py:29-49,84-123 -> pointpillar_scatter.py:17-41).
 *
 * hgsf_points_to_bev_train: three launches -- the pillarization front end, a statistics pass (second moments of the decorated
 * features over every kept point; from them batch_mean / batch_var [C], running_mean / running_var updated in place with
 * `momentum`, either may be NULL, and the sums the backward needs), and the fused pass of hgsf_points_to_bev normalising with
 * those batch statistics.  pfn->bn_mean / bn_var are ignored.  `stats`: hgsf_train_stats_doubles(C, Cin) doubles of device
 * memory, written here and read by the backward (keep it between the two).  out->spatial_features and out->pillar_features are
 * required; pass out->voxels when a backward follows.  HGSF_ERR_UNSUPPORTED outside the fused kernel's domain (BatchNorm,
 * 64 channels, max_points_per_voxel <= 32, spconv-2 overflow rule, grid[0] % 4 == 0): use the three contract calls above.
 *
 * hgsf_points_to_bev_train_backward: gradients of linear.weight [C,Cin], norm.weight [C], norm.bias [C] from the cotangents
 * of spatial_features (may be NULL) and of pillar_features ([capacity,C], may be NULL).  voxels / voxel_coords /
 * voxel_num_points are the forward's outputs as they are (capacity rows); the pillar count is read on the device from
 * num_pillars[0].  pfn->bn_mean / bn_var = the forward's batch_mean / batch_var; stats = the forward's buffer.  grad_rows:
 * device scratch [capacity, C]; scratch: hgsf_train_scratch_doubles doubles.  Three launches. */
HGSF_API int hgsf_points_to_bev_train(const hgsf_geometry *geom, const hgsf_points *points, const hgsf_pfn *pfn,
                                      int32_t max_points_per_voxel, int32_t max_voxels, void *workspace, size_t workspace_bytes,
                                      const hgsf_pillar_outputs *out, float momentum, float *running_mean, float *running_var,
                                      float *batch_mean, float *batch_var, double *stats, hgsf_stream_t stream);
HGSF_API int hgsf_points_to_bev_train_backward(const hgsf_geometry *geom, const hgsf_pfn *pfn, const float *voxels,
                                               const int32_t *voxel_coords, const int32_t *voxel_num_points,
                                               int64_t pillar_capacity, const int32_t *num_pillars,
                                               int32_t max_points_per_voxel, int32_t num_features, int32_t batch_size,
                                               const float *grad_spatial_features, const float *grad_pillar_features,
                                               float *grad_rows, const double *stats, double *scratch, float *grad_weight,
                                               float *grad_bn_weight, float *grad_bn_bias, hgsf_stream_t stream);

/* PointPillarScatter backward (autograd of pointpillar_scatter.py:33-35): grad_pillar_features[m, :] =
 * grad_spatial_features[b, :, y, x]; rows whose coordinates fall outside the canvas get zeros. */
HGSF_API int hgsf_pointpillar_scatter_backward(const hgsf_geometry *geom, const float *grad_spatial_features,
                                               const void *voxel_coords, int32_t coords_are_float, int64_t num_pillars,
                                               int32_t C, int32_t batch_size, float *grad_pillar_features,
                                               hgsf_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * Hybrid point assembly on the device -- the step in front of the path.  Replaces, per batch and
 * without a host sync, what the reference does per sample in float64 numpy inside DataLoader workers:
 *   pcdet/datasets/kitti/vod_dataset.py:498-522 / tj4d_dataset.py:588-610   raw sweep + mask points + virtual points ->
 *        [N, W+2] with the two flag columns: sweep rows (1,1) and 1 in the 8 label columns, mask rows (0,0), virtual rows
 *        (0,1); a frame without mask points keeps only its sweep; with mask points but no virtual points every row gets 1
 *        in the last column (the reference's points[-0:, -1] = 1)
 *   vod_dataset.py:13-19,511-514     NO_DUP: sweep rows whose squared distance to a mask point is <= dup_threshold dropped
 *   vod_dataset.py:181-197,525-528 + pcdet/utils/calibration_kitti.py:68-88   FOV_POINTS_ONLY
 *   pcdet/utils/common_utils.py:78-81 (data_processor.py:83-85)               x/y range mask, both ends inclusive
 *   pcdet/datasets/dataset.py:237-244, pcdet/models/__init__.py:23-36         batch-index column, float32
 * Inputs are float32 device arrays, the frames concatenated, with [batch_size+1] int32 device offsets each:
 *   real [sum Nr, real_features] (7 VoD, 8 TJ4D); gt_real / virt [.., hybrid_features] with hybrid_features =
 *   real_features + 8, or hybrid_features = 0 and NULL arrays for USE_VIRTUAL_POINTS False (output rows = the sweep).
 *   calib: NULL (FOV_POINTS_ONLY False) or [batch_size, 26] float32 per frame: the [4,3] lidar->rect matrix V2C^T.R0^T
 *   (row major, the float32 product calibration_kitti.py:74 forms), P2 [3,4] row major, image height, image width.
 *   range_xy: NULL (no range mask) or {xmin, ymin, xmax, ymax} as DOUBLES -- the reference compares float64 points with
 *   the Python floats of POINT_CLOUD_RANGE, which differs from a float32 compare for points on the boundary.
 * The filters are evaluated in double in the reference's order; the kept set and the row order are identical.
 * points_out [n_candidates, 1 + F] (F = hybrid_features + 2, or real_features) receives the kept rows, frame after frame
 * in input order, column 0 = frame index -- the collated layout hgsf_points_to_bev / hgsf_split_encode take;
 * frame_offsets_out [batch_size+1] int32 (device) = row offset of every frame, last entry = rows kept. */
typedef struct hgsf_hybrid_inputs {
    const float   *real;
    const float   *gt_real;
    const float   *virt;
    const int32_t *real_offsets;
    const int32_t *gt_offsets;
    const int32_t *virt_offsets;
    int64_t        n_candidates;     /* sum Nr + sum Ng + sum Nv */
    int32_t        real_features;
    int32_t        hybrid_features;
    int32_t        batch_size;
    int32_t        no_dup;
    double         dup_threshold;    /* 0.001 in the reference */
} hgsf_hybrid_inputs;
HGSF_API int hgsf_hybrid_workspace_size(int64_t n_candidates, size_t *bytes);
HGSF_API int hgsf_assemble_hybrid_points(const hgsf_hybrid_inputs *in, const float *calib, const double *range_xy,
                                         void *workspace, size_t workspace_bytes, float *points_out,
                                         int32_t *frame_offsets_out, hgsf_stream_t stream);

/* SparseConvTensor.dense() of the PillarNet branch (pcdet/models/backbones_3d/vfe/pillarnet_modules/lss_fpn.py:111-113,
 * rpn.py:243-247; spconv's scatter of [M,C] features at int32 indices [M,3] = (b, y, x) into zeros [B, C, ny, nx]):
 * SURVEY.md 8(a) row a16.  The same tile writer as hgsf_pointpillar_scatter (one pass over the dense tensor, zeros
 * included, TMA tile stores); rows whose index falls outside are ignored; duplicate indices resolve to the last row
 * (spconv's indices are unique).  C must be a multiple of 32, <= 256.  dense is written completely. */
HGSF_API int hgsf_sparse_to_dense_workspace_size(int32_t batch_size, int32_t ny, int32_t nx, size_t *bytes);
HGSF_API int hgsf_sparse_to_dense(const float *features, const int32_t *indices, int64_t num_rows, int32_t C,
                                  int32_t batch_size, int32_t ny, int32_t nx, void *workspace, size_t workspace_bytes,
                                  float *dense, hgsf_stream_t stream);

/* ------------------------------------------------------------------------------------------------------------
 * The pillar-list consumer: the submanifold 3x3 convolutions of SpMiddlePillarEncoder18.conv1
 * (pcdet/models/backbones_3d/vfe/pillarnet_modules/pcnres18.py:82-95 conv2D3x3 with stride 1 -> spconv.SubMConv2d,
 * :108-151 Sparse2DBasicBlockV, :154-188 Sparse2DBasicBlock, :212-215 conv1) evaluated on the reader's pillar list
 * (pillar_modules.py:82) instead of on a dense canvas.
 *
 * hgsf_subm_neighbors: the rule book of one indice_key -- neighbors [num_rows, 9] int32, entry tap = ky*3 + kx holds the
 *   pillar id at (y + ky - 1, x + kx - 1) of the same frame or -1 -- from pillar_bev_indices [B,H,W] and pillars [M,3]
 *   (b, y, x), both outputs of hgsf_pillarnet_indices.  Built once and reused by every convolution sharing the key.
 * hgsf_subm_conv3x3: out[m] = act( BN( bias + sum_tap W[tap] . in[neighbors[m][tap]] ) + residual[m] ), absent
 *   neighbours contributing nothing: the cross-correlation spconv.SubMConv2d(kernel 3, padding 1) computes on the active
 *   set.  fp32 FMA in (tap, input channel) order.  weight_layout HGSF_WEIGHT_KRSC: [Cout, 3, 3, Cin] (spconv 2.x),
 *   HGSF_WEIGHT_RSCK: [3, 3, Cin, Cout] (spconv 1.x).  bias, the four bn_* (BatchNorm1d eval, all or none) and residual
 *   [num_rows, Cout] may be NULL.  (Cin, Cout) in {(32,32), (32,64), (64,64)}.  out must not alias features;
 *   out == residual is allowed (each element is read and written by the same thread).
 * num_rows_dev: optional device int32 holding the row count (counts[0] of hgsf_pillarnet_indices) so that no host sync is
 *   needed; rows >= min(num_rows, *num_rows_dev) are left untouched. */
#define HGSF_WEIGHT_KRSC 0
#define HGSF_WEIGHT_RSCK 1
typedef struct hgsf_subm_conv {
    const float *weight;
    int32_t      weight_layout;
    const float *bias;          /* [Cout] or NULL */
    const float *bn_weight;     /* [Cout] or NULL */
    const float *bn_bias;
    const float *bn_mean;
    const float *bn_var;
    float        bn_eps;
    int32_t      in_channels;
    int32_t      out_channels;
    int32_t      relu;
} hgsf_subm_conv;
HGSF_API int hgsf_subm_neighbors(const int32_t *pillar_bev_indices, const int32_t *pillars, int64_t num_rows,
                                 const int32_t *num_rows_dev, int32_t batch_size, int32_t H, int32_t W,
                                 int32_t *neighbors, hgsf_stream_t stream);
HGSF_API int hgsf_subm_conv3x3(const hgsf_subm_conv *conv, const float *features, const int32_t *neighbors,
                               int64_t num_rows, const int32_t *num_rows_dev, const float *residual, float *out,
                               hgsf_stream_t stream);

/* The indices of SparseConv2d(kernel 3, stride 2, padding 1) -- the first layer of SpMiddlePillarEncoder18.conv2 / conv3 /
 * conv4 (pcnres18.py:217-221): the output active set (a cell (yo, xo) of the [Ho, Wo] = [(H-1)/2+1, (W-1)/2+1] grid is
 * active iff an input pillar lies in rows 2yo-1..2yo+1, columns 2xo-1..2xo+1), as out_pillars [Mo, 3] (b, y, x) in raster
 * order with its cell table out_bev [B, Ho, Wo] (-1 none) -- spconv's own output order is an implementation detail of its
 * hash table; raster order is what hgsf_pillarnet_indices gives the reader -- and the rule book neighbors [out_capacity, 9]:
 * input pillar id at (2yo + ky - 1, 2xo + kx - 1) or -1.  hgsf_subm_conv3x3 then evaluates the convolution from that rule
 * book (features = the INPUT rows, num_rows = out_capacity, num_rows_dev = out_counts).  out_counts [2] int32 (device):
 * {Mo, scratch}.  out_capacity >= min(4 * num_rows, B * Ho * Wo).  pillars must list the frames contiguously (the reader's
 * order).  No host sync. */
HGSF_API int hgsf_sparse_conv_s2_workspace_size(int64_t num_rows, int32_t batch_size, size_t *bytes);
HGSF_API int hgsf_sparse_conv_s2_indices(const int32_t *pillar_bev_indices, const int32_t *pillars, int64_t num_rows,
                                         const int32_t *num_rows_dev, int32_t batch_size, int32_t H, int32_t W,
                                         int32_t *out_bev, int32_t *out_pillars, int32_t *out_counts, int32_t *neighbors,
                                         int64_t out_capacity, void *workspace, size_t workspace_bytes, hgsf_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HGSFUSION_B200_H */
