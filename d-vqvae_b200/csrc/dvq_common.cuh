// Shared internals of libdvq_sm100.so (not part of the public ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "dvq.h"

namespace dvq {

// thread-local last-error text behind dvq_last_error()
void set_error(const char* fmt, ...);
int fail(int code, const char* fmt, ...);

#define DVQ_CUDA_CHECK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess)                                                                \
      return ::dvq::fail(DVQ_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                         __FILE__, __LINE__);                                             \
  } while (0)

struct DeviceProps {
  int sm_count;
  int cc_major;
  int cc_minor;
  int max_smem_optin;
};
int device_props(DeviceProps* out);  // cached per device

// instrumentation (dvq_api.cu)
void count_launch(int n = 1);
void profile_mark(int stage, bool begin, cudaStream_t s);  // no-op unless dvq_profile_enable(1)

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// ---- workspace layout of dvq_vq_forward (all offsets 256-byte aligned) -------------------
struct VqWorkspace {
  size_t off_ee;        // float [K]            ||e_k||^2
  size_t off_counters;  // int   [8]            refine-list length, overflow flags
  size_t off_rowlist;   // int   [2][N]         rows the tensor-core filter could not decide + their candidate masks
  size_t off_bop;       // operand image of the codebook for the tcgen05 path
  size_t off_binned;    // bin tables + (row, sub-chunk) pair list of the binned refine (vq_refine_binned.cu)
  size_t total;
};
VqWorkspace vq_workspace_layout(int64_t N, int K, int D, int flags);

// ---- kernels launched by the API layer ---------------------------------------------------
// ee[k] = sum_d E[k,d]^2
int launch_code_norms(const float* E, int K, int D, float* ee, cudaStream_t s);

// Exact FP32 CUDA-core path.  row_list == nullptr: rows [0,N).  Otherwise rows
// row_list[0 .. *n_list) (device-side count), used as the refine stage of the tcgen05 path.
int launch_vq_simt(const float* z, const float* E, const float* ee, int64_t N, int K, int D, int train,
                   float* z_q, int64_t* idx, unsigned long long* hist, double* sse,
                   const int* row_list, const int* n_list, int list_stride, cudaStream_t s);

// tcgen05 filter kernel (vq_tc_sm100.cu); supported(K,D) says whether the shape is handled.
bool vq_tc_supported(int64_t N, int K, int D);
void vq_tc_layout_info(int K, int D, int* out8);   // dvq_debug_tc_layout
bool vq_tc_pair_selected(int64_t N, int K, int D);   // CTA-pair (cta_group::2) kernel for this shape?
void vq_tc_pair_layout_info(int64_t N, int K, int D, int* out8);   // dvq_debug_tc_pair_layout
long long vq_tc_image_offset(int k, int d, int K, int D, int pair, long long* image_bytes);   // dvq_debug_tc_image_offset
size_t vq_tc_operand_bytes(int K, int D);
int launch_vq_tc(const float* z, const float* E, const float* ee, int64_t N, int K, int D, int train,
                 float* z_q, int64_t* idx, unsigned long long* hist, double* sse,
                 void* bop, int* counters, int* row_list, int* cand_list, int* zero_ints, int zero_n,
                 bool codebook_cached, cudaStream_t s);

// candidate-restricted exact refine (vq_refine.cu); cand_list[i] = bit mask of 32*2^gshift-code groups
bool vq_refine_supported(int K, int D);
bool vq_refine_codebook_in_smem(int K, int D);
int vq_tc_cand_gshift(int K);
int launch_vq_refine(const float* z, const float* E, const float* ee, int K, int D, int train, float* z_q, int64_t* idx,
                     unsigned long long* hist, double* sse, const int* row_list, const int* cand_list,
                     const int* n_list, int gshift, const int* ovf_last, const int* n_ovf, cudaStream_t s);

// binned exact refine (vq_refine_binned.cu): (row, sub-chunk) pairs bucketed by sub-chunk; hands the list to
// launch_vq_refine through counters[5] / counters[6] when the pairs do not fit its workspace
long long refine_pair_cap_override();   // dvq_vq_set_refine (dvq_api.cu); 0 = default
size_t vq_refine_binned_bytes(int64_t N);
size_t vq_refine_binned_zero_bytes();
int launch_vq_refine_binned(const float* z, const float* E, const float* ee, int64_t N, int K, int D, int train, float* z_q,
                            int64_t* idx, unsigned long long* hist, double* sse, const int* row_list, const int* cand_list,
                            int* counters, int list_mode, void* ws, cudaStream_t s);

int launch_onehot(const int64_t* idx, int64_t N, int K, float* onehot, cudaStream_t s);
int launch_finalize(const unsigned long long* hist, const double* sse, int64_t N, int K, int D, float al,
                    float beta, float* loss, float* ppl, cudaStream_t s);
int launch_vq_backward(const float* z, const float* E, const int64_t* idx, const float* g_zq, const float* g_loss,
                       const float* rows, int64_t N, int D, float al, float beta, float* dz, float* dE, cudaStream_t s);
int launch_gather_multi(const float* const* E, int G, const int64_t* codes, int64_t N, int K, int D, float* out, int64_t out_stride,
                        int* oob, cudaStream_t s);
int launch_vq_code_sums(const float* z, const int64_t* idx, int64_t N, int D, float* sums, cudaStream_t s);
int launch_gather(const float* E, const int64_t* idx, int64_t N, int K, int D, float* out, int* oob,
                  cudaStream_t s);

int launch_umma_probe(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes, int ksteps,
                      const uint32_t* strides, uint32_t idesc, int n_cols, float* out, int* err, cudaStream_t s);

// GatedPixelCNN sampler (pcnn_sm100.cu)
int launch_pcnn_gemm(const DvqPcnnGemm* g, cudaStream_t s);
int launch_pcnn_embed(const int64_t* x, int x_stride, int W, int B, int Bp, const float* emb, int n_emb, int d, void* img16,
                      float* img32, cudaStream_t s);
int launch_pcnn_rows_to_image(const int64_t* label, int B, int Bp, const float* table, int n_rows, int kd, void* img16, cudaStream_t s);

// MANO hand layer (mano_sm100.cu)
int launch_mano(const DvqManoModel* m, const float* betas, const float* global_orient, const float* hand_pose, const float* transl, int B,
                float* vertices, float* joints, cudaStream_t s);

// PointNet
size_t pointnet_workspace_bytes(int B, int C, int P, int flags);
int launch_pointnet(const float* x, const DvqPointNetWeights* w, int B, int C, int P, int flags, float* feat,
                    float* trans, void* ws, size_t ws_bytes, cudaStream_t s);
size_t pointnet_tc_image_bytes();
int launch_pointnet_tc_trunk(const float* x, const float* trans, const float* w1, const float* b1, const float* w2,
                             const float* b2, const float* w3, int B, int C, int P, float* maxbuf, void* images,
                             bool main_trunk, cudaStream_t s);

// ---- small device helpers ----------------------------------------------------------------
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace dvq
