// train_ops.cuh -- launchers of the training kernels (train_ops.cu).
#pragma once
#include "contract_ops.cuh"

namespace hgsf {

// doubles in the statistics buffer: Sx [C], Sxx [C], T [C][Cin], s [Cin]
inline size_t train_stats_len(int C, int cin) { return (size_t)2 * C + (size_t)C * cin + cin; }
// doubles in the backward accumulator: A [C][Cin], dGammaRaw [C], dBeta [C]
inline size_t train_acc_len(int C, int cin) { return (size_t)C * cin + 2 * (size_t)C; }

int launch_vfe_stats(const VfeParams &q, bool abs_xyz, bool dist, double *stats, cudaStream_t s);
int launch_bn_finalize(const double *stats, double n_rows, int C, float momentum, float *running_mean, float *running_var,
                       float *batch_mean, float *batch_var, cudaStream_t s);
int launch_vfe_backward(const VfeParams &q, bool abs_xyz, bool dist, const float *grad_out, const double *stats, int mode,
                        double *acc, float *grad_weight, float *grad_gamma, float *grad_beta, cudaStream_t s, int *launches);
int launch_scatter_grad(const float *grad_canvas, const void *coords, int coords_float, long long M, int C, int B, int ny, int nx,
                        float *grad_feats, cudaStream_t s, const int32_t *M_dev = nullptr, const float *add = nullptr);

}  // namespace hgsf
