// Fused PointNet object encoder (eval mode): replaces network/pointnet_encoder.py:27-45 (STN3d)
// and :140-166 (PointNetEncoder.forward, global_feat=True, feature_transform=False).
//
// One "trunk" kernel runs the whole shared MLP  C -> 64 -> 128 -> 1024  for a tile of 128 points
// of one cloud without ever writing an activation to HBM: the [B,64,P], [B,128,P] and three
// [B,1024,P] tensors the reference materialises (conv, bn, relu) live in shared memory /
// registers, and only the running per-channel maximum leaves the SM (one atomic max per
// channel per tile).  BatchNorm is pre-folded into the conv weights by the caller; the
// per-channel bias and the (monotone) ReLU commute with the max and are applied after it.
//
// Launch sequence per forward:  init -> trunk<STN> -> stn_head (fc 1024-512-256-9, +I)
//                               -> trunk<MAIN> (3x3 input transform fused into the load)
//                               -> decode.
// Loads of x[b, c, p0:p0+128] are point-major and coalesced; the max over the point set is a
// thread-local max over the register tile, a 16-lane shuffle tree, then the atomic.
//
// Roofline: compute (2*139 520 flop per point per trunk; 16*P bytes in + 4 KB out per cloud);
// this version runs the contractions on the FP32 FMA pipe with sequential-k accumulation.
#include <float.h>

#include "dvq_common.cuh"

namespace dvq {
namespace {

constexpr int PT = 128;         // points per CTA
constexpr int LDP = PT + 4;     // padded leading dim of activation tiles (floats)
constexpr int WK = 16;          // k-slice of the weight tile staged per step
constexpr int NTHREADS = 256;
constexpr int C1 = 64, C2 = 128, C3 = 1024;

// order-preserving float max through integer atomics (buffer initialised to -inf)
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f)
    atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else
    atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// acc[8][8] += W[m0 + rows, 0:KD] . act[0:KD, points]   (rows: ty*4.., 64+ty*4..; points: tx*4.., 64+tx*4..)
// W is row-major [M][KD] in global memory (L2-resident); act is k-major in shared memory.
template <int KD>
__device__ __forceinline__ void gemm_128x128(const float* __restrict__ W, int m0, const float* __restrict__ act,
                                             float (*wt)[LDP], float acc[8][8], int tid) {
  const int tx = tid & 15, ty = tid >> 4;
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  const float* w0 = W + (int64_t)(m0 + lrow) * KD + lk;
  const float* w1 = W + (int64_t)(m0 + lrow + 64) * KD + lk;
  float4 p0 = ldg4(w0), p1 = ldg4(w1);
#pragma unroll 1
  for (int k0 = 0; k0 < KD; k0 += WK) {
    __syncthreads();
    wt[lk + 0][lrow] = p0.x; wt[lk + 1][lrow] = p0.y; wt[lk + 2][lrow] = p0.z; wt[lk + 3][lrow] = p0.w;
    wt[lk + 0][lrow + 64] = p1.x; wt[lk + 1][lrow + 64] = p1.y; wt[lk + 2][lrow + 64] = p1.z; wt[lk + 3][lrow + 64] = p1.w;
    __syncthreads();
    if (k0 + WK < KD) { p0 = ldg4(w0 + k0 + WK); p1 = ldg4(w1 + k0 + WK); }
#pragma unroll
    for (int kk = 0; kk < WK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&wt[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&wt[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(act + (k0 + kk) * LDP + tx * 4);
      const float4 b1 = *reinterpret_cast<const float4*>(act + (k0 + kk) * LDP + 64 + tx * 4);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
}

struct TrunkWeights {
  const float *w1, *b1, *w2, *b2, *w3;
};

// dynamic smem: xs[4][LDP] | h1[64][LDP] | h2[128][LDP] | wt[16][LDP]
constexpr size_t kTrunkSmem = sizeof(float) * (size_t)LDP * (4 + C1 + C2 + WK);

template <bool MAIN>
__global__ void __launch_bounds__(NTHREADS, 2)
pointnet_trunk_kernel(const float* __restrict__ x, TrunkWeights w, const float* __restrict__ trans, int B, int C,
                      int P, int ptiles, float* __restrict__ maxbuf /* [B,1024], -inf initialised */) {
  extern __shared__ __align__(16) float smem[];
  float* xs = smem;
  float* h1 = xs + 4 * LDP;
  float* h2 = h1 + C1 * LDP;
  float(*wt)[LDP] = reinterpret_cast<float(*)[LDP]>(h2 + C2 * LDP);

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.x / ptiles;
  const int p0 = (blockIdx.x - b * ptiles) * PT;
  const int npts = min(PT, P - p0);

  // ---- load the point tile, point-major/coalesced; fuse xyz' = xyz . trans (:143-149) -------
  for (int i = tid; i < 4 * PT; i += NTHREADS) {
    const int c = i / PT, p = i - c * PT;
    float v = 0.f;
    if (c < C && p < npts) v = __ldg(x + ((int64_t)b * C + c) * P + p0 + p);
    xs[c * LDP + p] = v;
  }
  __syncthreads();
  if (MAIN) {
    if (tid < PT) {
      const float* t = trans + (int64_t)b * 9;
      const float x0 = xs[0 * LDP + tid], x1 = xs[1 * LDP + tid], x2 = xs[2 * LDP + tid];
      // bmm([P,3],[3,3]): out_j = sum_i x_i * T[i][j], sequential-i fmaf
      const float y0 = fmaf(x2, __ldg(t + 6), fmaf(x1, __ldg(t + 3), x0 * __ldg(t + 0)));
      const float y1 = fmaf(x2, __ldg(t + 7), fmaf(x1, __ldg(t + 4), x0 * __ldg(t + 1)));
      const float y2 = fmaf(x2, __ldg(t + 8), fmaf(x1, __ldg(t + 5), x0 * __ldg(t + 2)));
      xs[0 * LDP + tid] = y0; xs[1 * LDP + tid] = y1; xs[2 * LDP + tid] = y2;
    }
    __syncthreads();
  }

  // ---- layer 1: C -> 64, ReLU -----------------------------------------------------------------
  for (int i = tid; i < C1 * PT; i += NTHREADS) {
    const int c = i / PT, p = i - c * PT;
    float s = 0.f;
    for (int k = 0; k < C; ++k) s = fmaf(__ldg(w.w1 + c * C + k), xs[k * LDP + p], s);
    s += __ldg(w.b1 + c);
    h1[c * LDP + p] = fmaxf(s, 0.f);
  }
  // (gemm_128x128 starts with a __syncthreads)

  // ---- layer 2: 64 -> 128, ReLU ---------------------------------------------------------------
  {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    gemm_128x128<C1>(w.w2, 0, h1, wt, acc, tid);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int ch = (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
      const float bias = __ldg(w.b2 + ch);
      float4 lo, hi;
      lo.x = fmaxf(acc[i][0] + bias, 0.f); lo.y = fmaxf(acc[i][1] + bias, 0.f);
      lo.z = fmaxf(acc[i][2] + bias, 0.f); lo.w = fmaxf(acc[i][3] + bias, 0.f);
      hi.x = fmaxf(acc[i][4] + bias, 0.f); hi.y = fmaxf(acc[i][5] + bias, 0.f);
      hi.z = fmaxf(acc[i][6] + bias, 0.f); hi.w = fmaxf(acc[i][7] + bias, 0.f);
      *reinterpret_cast<float4*>(h2 + ch * LDP + tx * 4) = lo;
      *reinterpret_cast<float4*>(h2 + ch * LDP + 64 + tx * 4) = hi;
    }
  }

  // ---- layer 3: 128 -> 1024 in 8 chunks of 128 channels, fused max over the point set ----------
  bool valid[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) valid[j] = (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4)) < npts;
#pragma unroll 1
  for (int m0 = 0; m0 < C3; m0 += 128) {
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    gemm_128x128<C2>(w.w3, m0, h2, wt, acc, tid);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float m = -INFINITY;
#pragma unroll
      for (int j = 0; j < 8; ++j) m = valid[j] ? fmaxf(m, acc[i][j]) : m;
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (tx == 0) {
        const int ch = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        atomic_max_float(maxbuf + (int64_t)b * C3 + ch, m);
      }
    }
  }
}

__global__ void fill_kernel(float* p, int64_t n, float v) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// feat[b,c] = maxbuf[b,c] + bias[c]  (optionally ReLU) — bias / ReLU commute with the max
__global__ void decode_kernel(const float* __restrict__ maxbuf, const float* __restrict__ bias, int64_t n, int relu,
                              float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = maxbuf[i] + __ldg(bias + (i & (C3 - 1)));
    out[i] = relu ? fmaxf(v, 0.f) : v;
  }
}

// STN head (pointnet_encoder.py:33-45): g[1024] -> fc1+bn4+relu (512) -> fc2+bn5+relu (256) -> fc3 (9) + I.
// One CTA per HB clouds; one warp per output neuron, lanes stride the input (coalesced weight rows).
constexpr int HB = 4;
__global__ void __launch_bounds__(256)
stn_head_kernel(const float* __restrict__ g /* [B,1024] = relu(max+b3) */, DvqPointNetWeights w, int B,
                float* __restrict__ trans) {
  __shared__ float s0[HB][1024];
  __shared__ float s1[HB][512];
  __shared__ float s2[HB][256];
  const int b0 = blockIdx.x * HB;
  const int nb = min(HB, B - b0);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
  for (int i = tid; i < HB * 1024; i += blockDim.x) {
    const int bb = i >> 10;
    s0[bb][i & 1023] = bb < nb ? g[(int64_t)(b0 + bb) * 1024 + (i & 1023)] : 0.f;
  }
  __syncthreads();
  for (int o = warp; o < 512; o += nwarps) {
    float acc[HB] = {0.f, 0.f, 0.f, 0.f};
    const float* wr = w.stn_fc1_w + (int64_t)o * 1024;
    for (int k = lane; k < 1024; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int bb = 0; bb < HB; ++bb) acc[bb] = fmaf(wv, s0[bb][k], acc[bb]);
    }
#pragma unroll
    for (int bb = 0; bb < HB; ++bb) acc[bb] = warp_sum(acc[bb]);
    if (lane == 0) {
      const float bias = __ldg(w.stn_fc1_b + o);
#pragma unroll
      for (int bb = 0; bb < HB; ++bb) s1[bb][o] = fmaxf(acc[bb] + bias, 0.f);
    }
  }
  __syncthreads();
  for (int o = warp; o < 256; o += nwarps) {
    float acc[HB] = {0.f, 0.f, 0.f, 0.f};
    const float* wr = w.stn_fc2_w + (int64_t)o * 512;
    for (int k = lane; k < 512; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int bb = 0; bb < HB; ++bb) acc[bb] = fmaf(wv, s1[bb][k], acc[bb]);
    }
#pragma unroll
    for (int bb = 0; bb < HB; ++bb) acc[bb] = warp_sum(acc[bb]);
    if (lane == 0) {
      const float bias = __ldg(w.stn_fc2_b + o);
#pragma unroll
      for (int bb = 0; bb < HB; ++bb) s2[bb][o] = fmaxf(acc[bb] + bias, 0.f);
    }
  }
  __syncthreads();
  for (int o = warp; o < 9; o += nwarps) {
    float acc[HB] = {0.f, 0.f, 0.f, 0.f};
    const float* wr = w.stn_fc3_w + (int64_t)o * 256;
    for (int k = lane; k < 256; k += 32) {
      const float wv = __ldg(wr + k);
#pragma unroll
      for (int bb = 0; bb < HB; ++bb) acc[bb] = fmaf(wv, s2[bb][k], acc[bb]);
    }
#pragma unroll
    for (int bb = 0; bb < HB; ++bb) acc[bb] = warp_sum(acc[bb]);
    if (lane == 0) {
      const float bias = __ldg(w.stn_fc3_b + o);
      const float iden = (o == 0 || o == 4 || o == 8) ? 1.f : 0.f;
#pragma unroll
      for (int bb = 0; bb < HB; ++bb)
        if (bb < nb) trans[(int64_t)(b0 + bb) * 9 + o] = (acc[bb] + bias) + iden;
    }
  }
}

}  // namespace

// workspace: maxbuf_stn [B,1024] | maxbuf_main [B,1024] | g [B,1024] | (tensor-core path) FP16 weight images
size_t pointnet_workspace_bytes(int B, int C, int P, int flags) {
  (void)C; (void)P;
  return align_up(sizeof(float) * 1024 * (size_t)B, 256) * 3 + ((flags & DVQ_PN_FP16_TC) ? align_up(pointnet_tc_image_bytes(), 256) : 0);
}

int launch_pointnet(const float* x, const DvqPointNetWeights* w, int B, int C, int P, int flags, float* feat, float* trans,
                    void* ws, size_t ws_bytes, cudaStream_t s) {
  (void)ws_bytes;
  DeviceProps dp;
  int rc = device_props(&dp);
  if (rc) return rc;
  DVQ_CUDA_CHECK(cudaFuncSetAttribute(pointnet_trunk_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrunkSmem));
  DVQ_CUDA_CHECK(cudaFuncSetAttribute(pointnet_trunk_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTrunkSmem));
  const size_t stride = align_up(sizeof(float) * 1024 * (size_t)B, 256);
  float* max_stn = reinterpret_cast<float*>(static_cast<char*>(ws));
  float* max_main = reinterpret_cast<float*>(static_cast<char*>(ws) + stride);
  float* g = reinterpret_cast<float*>(static_cast<char*>(ws) + 2 * stride);
  void* images = static_cast<char*>(ws) + 3 * stride;
  const bool tcp = (flags & DVQ_PN_FP16_TC) != 0;
  const int64_t nfeat = (int64_t)B * 1024;
  const int ptiles = (P + PT - 1) / PT;
  const int dec_blocks = (int)((nfeat + 255) / 256 < (int64_t)dp.sm_count * 8 ? (nfeat + 255) / 256 : (int64_t)dp.sm_count * 8);
  if (tcp) {
    rc = launch_pointnet_tc_trunk(x, nullptr, w->stn_w1, w->stn_b1, w->stn_w2, w->stn_b2, w->stn_w3, B, C, P, max_stn, images, false, s);
    if (rc) return rc;
  } else {
    const int fill_blocks = (int)((2 * stride / 4 + 255) / 256 < (size_t)dp.sm_count * 8 ? (2 * stride / 4 + 255) / 256 : (size_t)dp.sm_count * 8);
    fill_kernel<<<fill_blocks, 256, 0, s>>>(max_stn, (int64_t)(2 * stride / 4), -INFINITY);
    DVQ_CUDA_CHECK(cudaGetLastError());
    count_launch();
    TrunkWeights ts = {w->stn_w1, w->stn_b1, w->stn_w2, w->stn_b2, w->stn_w3};
    pointnet_trunk_kernel<false><<<(unsigned)((int64_t)B * ptiles), NTHREADS, kTrunkSmem, s>>>(x, ts, nullptr, B, C, P, ptiles, max_stn);
    DVQ_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  decode_kernel<<<dec_blocks, 256, 0, s>>>(max_stn, w->stn_b3, nfeat, 1, g);
  DVQ_CUDA_CHECK(cudaGetLastError());
  count_launch();
  stn_head_kernel<<<(B + HB - 1) / HB, 256, 0, s>>>(g, *w, B, trans);
  DVQ_CUDA_CHECK(cudaGetLastError());
  count_launch();
  if (tcp) {
    rc = launch_pointnet_tc_trunk(x, trans, w->w1, w->b1, w->w2, w->b2, w->w3, B, C, P, max_main, images, true, s);
    if (rc) return rc;
  } else {
    TrunkWeights tm = {w->w1, w->b1, w->w2, w->b2, w->w3};
    pointnet_trunk_kernel<true><<<(unsigned)((int64_t)B * ptiles), NTHREADS, kTrunkSmem, s>>>(x, tm, trans, B, C, P, ptiles, max_main);
    DVQ_CUDA_CHECK(cudaGetLastError());
    count_launch();
  }
  decode_kernel<<<dec_blocks, 256, 0, s>>>(max_main, w->b3, nfeat, 0, feat);
  DVQ_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return DVQ_OK;
}

}  // namespace dvq
