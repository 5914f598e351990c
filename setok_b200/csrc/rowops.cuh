// rowops.cuh — host launch declarations for the row kernels (rowops.cu) and clustering kernels (dpc.cu).
#pragma once
#include "common.cuh"

namespace setok {
int launch_im2col(const void* images, int image_dtype, void* A, int B, int H, int W, int patch, int Kp, int split, cudaStream_t stream);
int launch_im2col_u8(const uint8_t* images, const setok_u8_norm* norm, void* A, int B, int H, int W, int patch, int Kp, int split, cudaStream_t stream);
int launch_cls_rows(float* emb, const float* cls, const float* pos, int B, int T, int C, cudaStream_t stream);
int launch_select_rows(const void* x, int x_dtype, void* out, int out_dtype, int B, int T, int skip, int C, const float* pos, cudaStream_t stream);
int launch_sort_by_cluster(const int64_t* idx_cluster, const int32_t* num_clusters, const int32_t* offsets, int B, int N,
                           int32_t* perm, int32_t* row_seg, int32_t* seg_off, cudaStream_t stream);
int launch_gather_rows(const float* in, float* out, const int32_t* perm, int rows, int C, cudaStream_t stream);
int launch_segment_mean(const float* x, const int32_t* seg_off, const int32_t* n_seg_dev, int cap, int C, float* out,
                        float* out2, cudaStream_t stream);
int launch_image_segments(const int32_t* offsets, int B, int32_t* row_seg, cudaStream_t stream);
int launch_convert(const void* in, int in_dtype, void* out, int out_dtype, long long n, cudaStream_t stream);
int launch_convert_rows(const void* in, int in_dtype, void* out, int out_dtype, int rows, int C, int act, const int32_t* m_dev,
                        cudaStream_t stream);
int launch_masked_softmax(const float* S, void* P, const int32_t* seg_off, const int32_t* row_seg, int rows, int N, int ldS, int ldP,
                          float scale, cudaStream_t stream);
int launch_iota_mod(int32_t* idx, int rows, int Q, cudaStream_t stream);
int launch_add_pos_rows(const float* in, const float* pos, void* out_bf16, int rows, int Q, int C, cudaStream_t stream);
// zeroes rows [*n_live, min(*n_live + n_rows, cap)) of a bf16 matrix (the rows a 64-key TMA box reads past the last live key)
int launch_zero_tail_rows(void* buf_bf16, long long ld, const int32_t* n_live_dev, int n_rows, int cap, int cols, cudaStream_t stream);
// first LayerNorm-fold record of the residual stream x [rows, C] f32 (records of 2 + 2*ceil(C/128) floats) and its xhat
int launch_ln_fold_init(const float* x, void* xhat_bf16, float* rec, float eps, int rows, int C, cudaStream_t stream);
int launch_preln_fold_init(const float* emb, float* x, void* xhat_bf16, float* rec, const float* gamma, const float* beta, float eps, int rows,
                           int C, cudaStream_t stream);
int get_pos_table(int h, int w, int C, const float** out, cudaStream_t stream);
}  // namespace setok
