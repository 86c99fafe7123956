// ref_pillar_ops_shim.cu -- TEST INFRASTRUCTURE.  extern "C" doors onto the REFERENCE's own Path B kernel launchers
// (pcdet/ops/pillar_ops/src/{pillar,group,scatter}_ops_gpu.cu, compiled unmodified from /root/reference by
// oracle/build_ref_pillar_ops.sh into oracle/_ref/libref_pillar_ops.so).  Only the launcher prototypes are declared
// here; no reference source is copied.  The launchers run on the legacy default stream and exit(-1) on a kernel error.
#include <cuda_runtime.h>

// prototypes as in pillar_ops_gpu.h:11-17, group_ops_gpu.h:8-9,13-14, scatter_ops_gpu.h:8-9
void create_pillar_indices_stack_kernel_launcher(int N, int B, int H, int W, float bev_size, const float *xyz,
                                                 const int *xyz_batch_cnt, bool *pillar_mask);
void create_pillar_indices_kernel_launcher(int B, int H, int W, const int *bevIndices, int *pillarIndices);
void create_pillar_indice_pairs_stack_kernel_launcher(int N, int B, int H, int W, float bev_size, const float *xyz,
                                                      const int *xyz_batch_cnt, const int *pillar_bev_indices, int *indice_pairs);
void flatten_indice_paris_kernel_launcher(int N, int K, const int *indicePairs, const int *position, int *firstIndices,
                                          int *secondIndices);
void gather_feature_kernel_launcher(int L, int C, const int *set_indices, const float *features, float *out);
void gather_feature_grad_kernel_launcher(int L, int C, const int *set_indices, const float *outGrad, float *inGrad);
void scatter_max_kernel_launcher(int C, int L, int M, const int *index, const float *src, int *arg, float *out);
void scatter_max_grad_kernel_launcher(int C, int M, const int *arg, const float *grad_out, float *grad_src);

#define REF_API extern "C" __attribute__((visibility("default")))

REF_API void ref_create_pillar_indices_stack(int N, int B, int H, int W, float bev_size, const float *xyz, const int *cnt, bool *mask) {
    create_pillar_indices_stack_kernel_launcher(N, B, H, W, bev_size, xyz, cnt, mask);
}
REF_API void ref_create_pillar_indices(int B, int H, int W, const int *bev, int *pillars) {
    create_pillar_indices_kernel_launcher(B, H, W, bev, pillars);
}
REF_API void ref_create_pillar_indice_pairs_stack(int N, int B, int H, int W, float bev_size, const float *xyz, const int *cnt,
                                                  const int *bev, int *pairs) {
    create_pillar_indice_pairs_stack_kernel_launcher(N, B, H, W, bev_size, xyz, cnt, bev, pairs);
}
REF_API void ref_flatten_indice_pairs(int N, int K, const int *pairs, const int *position, int *first, int *second) {
    flatten_indice_paris_kernel_launcher(N, K, pairs, position, first, second);
}
REF_API void ref_gather_feature(int L, int C, const int *idx, const float *features, float *out) {
    gather_feature_kernel_launcher(L, C, idx, features, out);
}
REF_API void ref_gather_feature_grad(int L, int C, const int *idx, const float *out_grad, float *in_grad) {
    gather_feature_grad_kernel_launcher(L, C, idx, out_grad, in_grad);
}
REF_API void ref_scatter_max(int C, int L, int M, const int *index, const float *src, int *arg, float *out) {
    scatter_max_kernel_launcher(C, L, M, index, src, arg, out);
}
REF_API void ref_scatter_max_grad(int C, int M, const int *arg, const float *grad_out, float *grad_src) {
    scatter_max_grad_kernel_launcher(C, M, arg, grad_out, grad_src);
}
