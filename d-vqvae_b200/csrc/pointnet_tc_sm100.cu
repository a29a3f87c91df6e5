// tcgen05 PointNet trunk: the shared MLP C -> 64 -> 128 -> 1024 + max-pool over the point set with the
// 64->128 and 128->1024 layers (99.8 % of the flops) on the 5th-generation tensor cores, FP16 operands
// (the same 10-bit mantissa as the TF32 cuDNN path the reference takes on a GPU), FP32 accumulation in
// TMEM.  Replaces network/pointnet_encoder.py:29-32 (STN trunk) and :150-164 (main trunk).
//
// Work split: CTA = (group of clouds, quarter of the 1024 output channels).  The CTA keeps its quarter
// of the layer-3 weights (256 x 128 fp16 = 64 KB) and all of layer 2 (128 x 80 fp16 incl. the bias
// fold column) resident in shared memory as UMMA operand images and walks over 256-point tiles:
//   L1  (CUDA cores, thread = point): coalesced point-major load of x[b, :, p0:p0+256], optional 3x3
//       input transform, 4 -> 64 in FP32, ReLU fused into the FP16 pack (cvt.rn.relu.f16x2.f32),
//       rows written straight into the K-major A image of layer 2 (+ a 1.0 column for the bias).
//   L2  tcgen05.mma: D2[point, channel] = h1 . W2^T (M = 128 points x 2 halves, N = 128, K = 80);
//       epilogue: thread = point reads its 128 channels from TMEM, ReLU + FP16 pack, writes its row of
//       the K-major B image of layer 3 (16 conflict-free 128-bit stores).
//   L3  tcgen05.mma: D3[channel, point] = W3 . h2^T (M = 128 channels, N = 128 points, K = 128), so the
//       max over the point set is a THREAD-LOCAL 3-input max along the TMEM columns of each lane; the
//       per-channel running maximum lives in registers across all tiles of the cloud and is written
//       once (no atomics).  Points beyond P replicate the last valid point, which leaves the max intact.
// Bias (and the STN's ReLU) of layer 3 commute with the max and are applied by decode_kernel.
// Layers 1-2 are recomputed by the 4 channel-quarter CTAs (6 % of the flops) — cheaper than streaming
// 256 KB of layer-3 weights per tile from L2.
//
// Roofline: tensor pipe; algorithmic 2 * 139 520 flop per point per trunk, 16 * P bytes in + 4 KB out
// per cloud.  Inside a CTA the CUDA-core layer 1 of tile t+1 runs under the layer-3 MMAs of tile t; the
// layer-2 MMA + epilogue and the layer-3 epilogues are still exposed (one CTA per SM), see DESIGN.md.
#include <cuda_fp16.h>
#include <type_traits>
#include <stdlib.h>

#include "dvq_common.cuh"
#include "tc_prims.cuh"

namespace dvq {
namespace {

constexpr int PT = 256;                 // points per tile
constexpr int NC = 256;                 // compute threads (thread <-> point in L1 / L2 epilogue)
constexpr int NT = NC + 32;             // + one warp whose lane 0 only issues the MMAs (keeps the compute warps free)
constexpr int K1 = 80;                  // layer-2 contraction length: 64 channels + 16 fold columns
constexpr int KC1 = K1 / 8;             // 8-wide k-chunks of the layer-2 operands
constexpr int KC2 = 128 / 8;            // k-chunks of the layer-3 operands

// shared memory (bytes); every operand image is [k-chunk][row][16 B] (K-major, no swizzle)
constexpr uint32_t W3Q_BYTES = KC2 * 256 * 16;   // 64 KB: 256 channels of this quarter
constexpr uint32_t W2_BYTES = KC1 * 128 * 16;    // 20 KB
constexpr uint32_t H1_BYTES = KC1 * PT * 16;     // 40 KB
constexpr uint32_t H2_BYTES = KC2 * PT * 16;     // 64 KB
constexpr uint32_t OFF_W3Q = 0;
constexpr uint32_t OFF_W2 = OFF_W3Q + W3Q_BYTES;
constexpr uint32_t OFF_H1 = OFF_W2 + W2_BYTES;
constexpr uint32_t OFF_H2 = OFF_H1 + H1_BYTES;
constexpr uint32_t OFF_X = OFF_H2 + H2_BYTES;            // float xs[4][PT]
constexpr uint32_t OFF_W1 = OFF_X + 4 * PT * 4;          // float4 w1[64] + float b1[64]
constexpr uint32_t OFF_MAX = OFF_W1 + 64 * 16 + 64 * 4;  // float mx[2][128] (warps 4-7 -> warps 0-3)
constexpr uint32_t SMEM_TOTAL = OFF_MAX + 2 * 128 * 4;

struct TcTrunkParams {
  const float* x;        // [B,C,P]
  const float* trans;    // [B,9] or nullptr (STN trunk)
  const float* w1;       // [64,C] folded fp32
  const float* b1;       // [64]
  const uint8_t* w2img;  // FP16 operand image of layer 2 (W2_BYTES)
  const uint8_t* w3img;  // FP16 operand images of layer 3: 4 quarters x W3Q_BYTES
  float* maxbuf;         // [B,1024] raw maxima (bias / ReLU applied by decode_kernel)
  int B, C, P;
};

// folded fp32 weights -> FP16 UMMA operand images (runs per forward; ~300 KB)
__global__ void pointnet_pack_kernel(const float* __restrict__ w2, const float* __restrict__ b2,
                                     const float* __restrict__ w3, uint8_t* __restrict__ w2img, uint8_t* __restrict__ w3img) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int stride = gridDim.x * blockDim.x;
  for (int i = gid; i < 128 * K1; i += stride) {          // W2 image: row = channel, K = 64 + fold
    const int c = i / K1, k = i - c * K1;
    float v = 0.f;
    if (k < 64) v = w2[c * 64 + k];
    else if (k == 64) v = b2[c];                           // multiplied by the 1.0 column of h1
    *reinterpret_cast<__half*>(w2img + (size_t)(k >> 3) * 128 * 16 + (size_t)c * 16 + (k & 7) * 2) = __float2half_rn(v);
  }
  for (int i = gid; i < 1024 * 128; i += stride) {        // W3 images: 4 quarters of 256 channels
    const int c = i >> 7, k = i & 127;
    const int q = c >> 8, r = c & 255;
    *reinterpret_cast<__half*>(w3img + (size_t)q * W3Q_BYTES + (size_t)(k >> 3) * 256 * 16 + (size_t)r * 16 + (k & 7) * 2) =
        __float2half_rn(w3[i]);
  }
}

__device__ __forceinline__ uint32_t pack_relu_f16x2(float lo, float hi) {   // {relu(lo), relu(hi)} as fp16 pair
  uint32_t r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float max3f(float a, float b, float c) {
  float r;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

template <bool MAIN>
__global__ void __launch_bounds__(NT, 1) pointnet_trunk_tc_kernel(const TcTrunkParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_l2, bar_l3[2];
  __shared__ uint32_t tmem_slot;
  __shared__ int serr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool compute = warp < NC / 32;                // warps 0-7; warp 8 issues the tensor-core work
  const bool issuer = (tid == NC);
  const int q = blockIdx.x & 3;                       // channel quarter
  const int group = blockIdx.x >> 2, ngroups = gridDim.x >> 2;
  const int C = p.C, P = p.P;
  float4* w1s = reinterpret_cast<float4*>(smem + OFF_W1);
  float* b1s = reinterpret_cast<float*>(smem + OFF_W1 + 64 * 16);
  float* mxs = reinterpret_cast<float*>(smem + OFF_MAX);

  // ---- one-time setup: weights of this CTA into shared memory, TMEM, barrier ---------------------
  {
    const uint4* src3 = reinterpret_cast<const uint4*>(p.w3img + (size_t)q * W3Q_BYTES);
    uint4* dst3 = reinterpret_cast<uint4*>(smem + OFF_W3Q);
    for (uint32_t i = tid; i < W3Q_BYTES / 16; i += NT) dst3[i] = __ldg(src3 + i);
    const uint4* src2 = reinterpret_cast<const uint4*>(p.w2img);
    uint4* dst2 = reinterpret_cast<uint4*>(smem + OFF_W2);
    for (uint32_t i = tid; i < W2_BYTES / 16; i += NT) dst2[i] = __ldg(src2 + i);
    if (tid < 64) {
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      w.x = __ldg(p.w1 + tid * C + 0); w.y = __ldg(p.w1 + tid * C + 1); w.z = __ldg(p.w1 + tid * C + 2);
      if (C > 3) w.w = __ldg(p.w1 + tid * C + 3);
      w1s[tid] = w;
      b1s[tid] = __ldg(p.b1 + tid);
    }
    // fold k-chunks of the h1 image (columns 64..79): column 64 = 1.0 for every point, rest 0 — constant
    for (int r = tid; r < PT; r += NT) {
      uint4 one = make_uint4(0x00003c00u, 0u, 0u, 0u);   // fp16 1.0 in element 0
      *reinterpret_cast<uint4*>(smem + OFF_H1 + 8 * PT * 16 + r * 16) = one;
      *reinterpret_cast<uint4*>(smem + OFF_H1 + 9 * PT * 16 + r * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (tid == 0) {
    serr = 0;
    tc::mbar_init(&bar_l2, 1);
    tc::mbar_init(&bar_l3[0], 1);
    tc::mbar_init(&bar_l3[1], 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  volatile int* errw = &serr;
  uint32_t phase = 0;   // every barrier completes exactly once per tile

  const uint32_t idesc = tc::make_idesc_f16(128, 128, 0);
  const uint64_t d_h1 = tc::make_smem_desc(tc::smem_u32(smem + OFF_H1), PT * 16, 128);
  const uint64_t d_w2 = tc::make_smem_desc(tc::smem_u32(smem + OFF_W2), 128 * 16, 128);
  const uint64_t d_w3 = tc::make_smem_desc(tc::smem_u32(smem + OFF_W3Q), 256 * 16, 128);
  const uint64_t d_h2 = tc::make_smem_desc(tc::smem_u32(smem + OFF_H2), PT * 16, 128);
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  const int ntiles = (P + PT - 1) / PT;

  for (int b = group; b < p.B; b += ngroups) {
    float run0 = -INFINITY, run1 = -INFINITY;   // running max of channels (q*256 + mc*128 + (warp&3)*32 + lane), mc = 0,1
    float t00 = 1.f, t01 = 0.f, t02 = 0.f, t10 = 0.f, t11 = 1.f, t12 = 0.f, t20 = 0.f, t21 = 0.f, t22 = 1.f;
    if (MAIN) {
      const float* t = p.trans + (size_t)b * 9;
      t00 = __ldg(t + 0); t01 = __ldg(t + 1); t02 = __ldg(t + 2); t10 = __ldg(t + 3); t11 = __ldg(t + 4);
      t12 = __ldg(t + 5); t20 = __ldg(t + 6); t21 = __ldg(t + 7); t22 = __ldg(t + 8);
    }
    // L1 of one tile: thread = point; points past the end replicate the last valid one
    auto layer1 = [&](int tile) {
      const int p0 = tile * PT;
      const int npts = min(PT, P - p0);
      const int pp = p0 + min(tid, npts - 1);
      const float* xb = p.x + (size_t)b * C * P + pp;
      float x0 = __ldg(xb), x1 = __ldg(xb + P), x2 = __ldg(xb + 2 * (size_t)P);
      const float x3 = C > 3 ? __ldg(xb + 3 * (size_t)P) : 0.f;
      if (MAIN) {   // bmm([P,3],[3,3]) (:146): out_j = sum_i x_i T[i][j], sequential-i fmaf
        const float y0 = fmaf(x2, t20, fmaf(x1, t10, x0 * t00));
        const float y1 = fmaf(x2, t21, fmaf(x1, t11, x0 * t01));
        const float y2 = fmaf(x2, t22, fmaf(x1, t12, x0 * t02));
        x0 = y0; x1 = y1; x2 = y2;
      }
      uint8_t* row = smem + OFF_H1 + tid * 16;
#pragma unroll
      for (int kc = 0; kc < 8; ++kc) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c0 = kc * 8 + e * 2;
          const float4 wa = w1s[c0], wb = w1s[c0 + 1];
          const float va = fmaf(wa.w, x3, fmaf(wa.z, x2, fmaf(wa.y, x1, wa.x * x0))) + b1s[c0];
          const float vb = fmaf(wb.w, x3, fmaf(wb.z, x2, fmaf(wb.y, x1, wb.x * x0))) + b1s[c0 + 1];
          pk[e] = pack_relu_f16x2(va, vb);
        }
        *reinterpret_cast<uint4*>(row + kc * PT * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
      tc::fence_proxy_async_smem();
    };
    // L3 epilogue of one 128-channel chunk: lane = channel, columns = points, thread-local 3-input max
    auto chunk_max = [&](uint32_t col_base) {
      float m = -INFINITY;
#pragma unroll
      for (int c0 = 0; c0 < 128; c0 += 32) {
        uint32_t v[32];
        tc::tmem_ld32(tmem + lane_addr + col_base + (uint32_t)(warp >> 2) * 128u + (uint32_t)c0, v);
        tc::tmem_ld_wait();
        float a[11];
#pragma unroll
        for (int i = 0; i < 10; ++i) a[i] = max3f(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
        a[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
        m = max3f(m, max3f(a[0], a[1], a[2]), max3f(a[3], a[4], a[5]));
        m = max3f(m, max3f(a[6], a[7], a[8]), fmaxf(a[9], a[10]));
      }
      return m;
    };

    if (compute) layer1(0);
    __syncthreads();
    bool ok = true;
    for (int tile = 0; ok && tile < ntiles; ++tile) {
      // ---- L2: two M=128 halves of points, N = 128 channels, K = 80 -> TMEM columns [0,256) ----------
      if (issuer) {
        tc::tc_fence_after();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint64_t ad = d_h1 + (uint64_t)((h * 128 * 16) >> 4), bd = d_w2;
#pragma unroll
          for (int j = 0; j < K1 / 16; ++j) {
            tc::umma_f16(tmem + (uint32_t)h * 128u, ad, bd, idesc, j > 0 ? 1u : 0u);
            ad += (uint64_t)((2 * PT * 16) >> 4); bd += (uint64_t)((2 * 128 * 16) >> 4);
          }
        }
        tc::umma_commit(&bar_l2);
      }
      if (!tc::mbar_wait(&bar_l2, phase, errw, 1)) { ok = false; break; }
      tc::tc_fence_after();
      // L2 epilogue: thread = point (warp w: half = w / 4, TMEM lane quarter = w % 4)
      if (compute) {
        const int half = warp >> 2;
        const int prow = half * 128 + (warp & 3) * 32 + lane;
        uint8_t* row = smem + OFF_H2 + prow * 16;
#pragma unroll
        for (int c0 = 0; c0 < 128; c0 += 32) {
          uint32_t v[32];
          tc::tmem_ld32(tmem + lane_addr + (uint32_t)half * 128u + (uint32_t)c0, v);
          tc::tmem_ld_wait();
#pragma unroll
          for (int kc = 0; kc < 4; ++kc) {
            uint32_t pk[4];
#pragma unroll
            for (int e = 0; e < 4; ++e)
              pk[e] = pack_relu_f16x2(__uint_as_float(v[kc * 8 + e * 2]), __uint_as_float(v[kc * 8 + e * 2 + 1]));
            *reinterpret_cast<uint4*>(row + (c0 / 8 + kc) * PT * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
        }
      }
      tc::tc_fence_before();
      tc::fence_proxy_async_smem();
      __syncthreads();
      // ---- L3: both 128-channel chunks issued back to back (chunk 0 -> columns [256,512), chunk 1 ->
      //      the columns [0,256) the L2 epilogue just drained); the CUDA-core layer 1 of the NEXT tile
      //      runs underneath these 2048 tensor-pipe cycles ---------------------------------------------
      if (issuer) {
        tc::tc_fence_after();
#pragma unroll
        for (int mc = 0; mc < 2; ++mc) {
#pragma unroll
          for (int ph = 0; ph < 2; ++ph) {
            uint64_t ad = d_w3 + (uint64_t)((mc * 128 * 16) >> 4), bd = d_h2 + (uint64_t)((ph * 128 * 16) >> 4);
#pragma unroll
            for (int j = 0; j < 128 / 16; ++j) {
              tc::umma_f16(tmem + (mc == 0 ? 256u : 0u) + (uint32_t)ph * 128u, ad, bd, idesc, j > 0 ? 1u : 0u);
              ad += (uint64_t)((2 * 256 * 16) >> 4); bd += (uint64_t)((2 * PT * 16) >> 4);
            }
          }
          tc::umma_commit(&bar_l3[mc]);
        }
      }
      if (compute && tile + 1 < ntiles) layer1(tile + 1);   // h1 is free: the L2 MMAs of this tile completed above
      if (!tc::mbar_wait(&bar_l3[0], phase, errw, 2)) { ok = false; break; }
      tc::tc_fence_after();
      if (compute) run0 = fmaxf(run0, chunk_max(256u));
      if (!tc::mbar_wait(&bar_l3[1], phase, errw, 3)) { ok = false; break; }
      tc::tc_fence_after();
      if (compute) run1 = fmaxf(run1, chunk_max(0u));
      phase ^= 1u;
      tc::tc_fence_before();
      __syncthreads();   // accumulators drained, h1(tile+1) complete, h2 free for the next L2 epilogue
    }
    // ---- cloud done: combine the two point-half partials and write this quarter's 256 maxima ------
    if (compute && warp >= 4) { mxs[(warp & 3) * 32 + lane] = run0; mxs[128 + (warp & 3) * 32 + lane] = run1; }
    __syncthreads();
    if (warp < 4) {
      const int c = warp * 32 + lane;
      float* out = p.maxbuf + (size_t)b * 1024 + q * 256;
      out[c] = fmaxf(run0, mxs[c]);
      out[128 + c] = fmaxf(run1, mxs[128 + c]);
    }
    __syncthreads();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}


// ------------------------------------------------------------------------------------------------------------------
// Pipelined trunk (the default): all three layers on the tensor pipe, back to back.
//
// 128-point tiles.  Layer 1 (C -> 64) is a tcgen05.mma too: the point coordinates enter as an FP16 pair per value
// (x = xh + xl, 22 bits) in a 16-column A row  [xh(4) | 1 | xl(4) | xh(4) | 1 | 0 0]  against the weight row
// [wh(4) | bh | wh(4) | wl(4) | bl | 0 0]  (w = wh + wl, b = bh + bl), i.e. x.w + b up to the xl.wl terms (2^-22) —
// the CUDA-core version of this layer was 490 instructions per point, two thirds of everything the compute warps
// issued, and kept a fully overlapped pipeline compute-bound (the 128-point variant with a CUDA-core layer 1 measured
// 4 % slower than the 256-point kernel above).  With it a round costs the compute warps ~150 instructions per thread.
//
// Tensor queue of round r (one issuing thread, in order):
//     L1(r+2) 32 clk | L2(r+1) 320 clk | L3 chunk 0 (r) 512 clk | wait "chunk 1 drained" | L3 chunk 1 (r) 512 clk
// Two groups of eight warps work underneath, each step gated by the commit barrier of the MMA it consumes:
//     feed :  coordinates of tile r+3 -> x16 (those of r+4 requested);  layer-2 epilogue (r+1) -> h2[(r+1) & 1];
//             layer-1 epilogue (r+2) -> h1
//     drain:  max-epilogue of chunk 1 (r-1) -> arrive "chunk 1 drained";  max-epilogue of chunk 0 (r)
// (as ONE group doing all five steps in sequence the round took 1.8x its tensor time: 869 TFLOP/s).
// TMEM: layer-2 accumulator [0,128), layer-3 chunk 0 [128,256), chunk 1 [256,384), layer-1 accumulator [384,448).
// The tile stream runs across the clouds of the CTA's group without draining the pipeline; running maxima are
// written when a cloud ends.
// ------------------------------------------------------------------------------------------------------------------
constexpr int PT2 = 128;
constexpr int CG2 = 2;                            // column groups: warps of a role per TMEM lane quarter
constexpr int NF2 = CG2 * 128;                    // feed warps (epilogues of layers 1 / 2, coordinates) ...
constexpr int NC2 = 2 * NF2;                      // ... + as many drain warps (max-epilogues of layer 3): the two chains run side by side
constexpr int NT2 = NC2 + 32;
constexpr uint32_t H1B = KC1 * PT2 * 16;          // 20 KB
constexpr uint32_t H2B = KC2 * PT2 * 16;          // 32 KB
constexpr uint32_t X16B = 2 * PT2 * 16;           // 4 KB: layer-1 A image, 16 columns
constexpr uint32_t W1B = 2 * 64 * 16;             // 2 KB: layer-1 B image, 64 channels x 16 columns
constexpr uint32_t O2_W3Q = 0;
constexpr uint32_t O2_W2 = O2_W3Q + W3Q_BYTES;
constexpr uint32_t O2_H1 = O2_W2 + W2_BYTES;
constexpr uint32_t O2_H2 = O2_H1 + H1B;           // two buffers
constexpr uint32_t O2_X16 = O2_H2 + 2 * H2B;
constexpr uint32_t O2_W1 = O2_X16 + X16B;
constexpr uint32_t O2_MAX = O2_W1 + W1B;
constexpr uint32_t SMEM2_TOTAL = O2_MAX + (CG2 - 1) * 256 * 4;   // partial maxima of the column groups 1..CG2-1
constexpr uint32_t TM_L2 = 0, TM_C0 = 128, TM_C1 = 256, TM_L1 = 384;

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {   // 32 lanes x 16 consecutive columns
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ float f16_lo_as_float(uint32_t pair) { return __half2float(__ushort_as_half((unsigned short)(pair & 0xffffu))); }
__device__ __forceinline__ float f16_hi_as_float(uint32_t pair) { return __half2float(__ushort_as_half((unsigned short)(pair >> 16))); }

template <bool MAIN>
__global__ void __launch_bounds__(NT2, 1) pointnet_trunk_tc2_kernel(const TcTrunkParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ uint64_t bar_l1, bar_l2, bar_l3[2], bar_c1free;
  __shared__ uint32_t tmem_slot;
  __shared__ int serr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool compute = warp < NC2 / 32;
  const int q = blockIdx.x & 3;
  const int group = blockIdx.x >> 2, ngroups = gridDim.x >> 2;
  const int C = p.C, P = p.P;
  float* mxs = reinterpret_cast<float*>(smem + O2_MAX);
  {
    const uint4* src3 = reinterpret_cast<const uint4*>(p.w3img + (size_t)q * W3Q_BYTES);
    uint4* dst3 = reinterpret_cast<uint4*>(smem + O2_W3Q);
    for (uint32_t i = tid; i < W3Q_BYTES / 16; i += NT2) dst3[i] = __ldg(src3 + i);
    const uint4* src2 = reinterpret_cast<const uint4*>(p.w2img);
    uint4* dst2 = reinterpret_cast<uint4*>(smem + O2_W2);
    for (uint32_t i = tid; i < W2_BYTES / 16; i += NT2) dst2[i] = __ldg(src2 + i);
    if (tid < 64) {   // layer-1 weight row of channel tid: [wh(4) | bh | wh(4) | wl(4) | bl | 0 0], two 8-wide k-chunks
      float w[4] = {0.f, 0.f, 0.f, 0.f};
      for (int c = 0; c < C && c < 4; ++c) w[c] = __ldg(p.w1 + tid * C + c);
      const float bias = __ldg(p.b1 + tid);
      const uint32_t wh01 = pack_f16x2(w[0], w[1]), wh23 = pack_f16x2(w[2], w[3]);
      const uint32_t wl01 = pack_f16x2(w[0] - f16_lo_as_float(wh01), w[1] - f16_hi_as_float(wh01));
      const uint32_t wl23 = pack_f16x2(w[2] - f16_lo_as_float(wh23), w[3] - f16_hi_as_float(wh23));
      const uint32_t bh = pack_f16x2(bias, 0.f);
      const uint32_t bl = pack_f16x2(bias - f16_lo_as_float(bh), 0.f);
      // columns: 0-3 wh | 4 bh | 5-7 wh0..2   ||   8 wh3 | 9-12 wl | 13 bl | 14-15 0
      const uint32_t c45 = (bh & 0xffffu) | (wh01 << 16);                        // {bh, wh0}
      const uint32_t c67 = (wh01 >> 16) | (wh23 << 16);                          // {wh1, wh2}
      const uint32_t c89 = (wh23 >> 16) | (wl01 << 16);                          // {wh3, wl0}
      const uint32_t cab = (wl01 >> 16) | (wl23 << 16);                          // {wl1, wl2}
      const uint32_t ccd = (wl23 >> 16) | (bl << 16);                            // {wl3, bl}
      *reinterpret_cast<uint4*>(smem + O2_W1 + tid * 16) = make_uint4(wh01, wh23, c45, c67);
      *reinterpret_cast<uint4*>(smem + O2_W1 + 64 * 16 + tid * 16) = make_uint4(c89, cab, ccd, 0u);
    }
    for (int r = tid; r < PT2; r += NT2) {   // fold k-chunks of h1: column 64 = 1.0 (layer-2 bias), rest 0 — constant
      *reinterpret_cast<uint4*>(smem + O2_H1 + 8 * PT2 * 16 + r * 16) = make_uint4(0x00003c00u, 0u, 0u, 0u);
      *reinterpret_cast<uint4*>(smem + O2_H1 + 9 * PT2 * 16 + r * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  if (tid == 0) {
    serr = 0;
    tc::mbar_init(&bar_l1, 1);
    tc::mbar_init(&bar_l2, 1);
    tc::mbar_init(&bar_l3[0], 1);
    tc::mbar_init(&bar_l3[1], 1);
    tc::mbar_init(&bar_c1free, NF2 / 32);   // one arrival per drain warp
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  tc::fence_proxy_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  volatile int* errw = &serr;

  const uint32_t idesc = tc::make_idesc_f16(128, 128, 0);
  const uint32_t idesc_l1 = tc::make_idesc_f16(128, 64, 0);
  const uint64_t d_x16 = tc::make_smem_desc(tc::smem_u32(smem + O2_X16), PT2 * 16, 128);
  const uint64_t d_w1 = tc::make_smem_desc(tc::smem_u32(smem + O2_W1), 64 * 16, 128);
  const uint64_t d_h1 = tc::make_smem_desc(tc::smem_u32(smem + O2_H1), PT2 * 16, 128);
  const uint64_t d_w2 = tc::make_smem_desc(tc::smem_u32(smem + O2_W2), 128 * 16, 128);
  const uint64_t d_w3 = tc::make_smem_desc(tc::smem_u32(smem + O2_W3Q), 256 * 16, 128);
  const uint32_t lane_addr = (uint32_t)((warp & 3) * 32) << 16;
  const int ntiles = (P + PT2 - 1) / PT2;
  const int nclouds = p.B > group ? (p.B - 1 - group) / ngroups + 1 : 0;
  const int S = nclouds * ntiles;                 // tile stream of this CTA

  // ---- roles of one round -----------------------------------------------------------------------------------------
  // x16 of stream tile s: threads 0..127 = points; points past the end replicate the last valid one.  The coordinates
  // are requested one round ahead (load_point) so that their global-memory latency is not on the round's path.
  // (the streams advance by one tile per call: cloud / tile counters instead of a division per round)
  float px0 = 0.f, px1 = 0.f, px2 = 0.f, px3 = 0.f;
  int lp_b = group, lp_tile = 0;   // cloud and tile of the next load_point
  int px_b = group;                // cloud of the coordinates held in px*
  auto load_point = [&]() {
    if (tid < PT2) {
      const int b = lp_b, tile = lp_tile;
      px_b = b;
      if (++lp_tile == ntiles) { lp_tile = 0; lp_b += ngroups; }
      const int p0 = tile * PT2;
      const int npts = min(PT2, P - p0);
      const int pp = p0 + min(tid, npts - 1);
      const float* xb = p.x + (size_t)b * C * P + pp;
      px0 = __ldg(xb); px1 = __ldg(xb + P); px2 = __ldg(xb + 2 * (size_t)P);
      px3 = C > 3 ? __ldg(xb + 3 * (size_t)P) : 0.f;
    }
  };
  auto write_x16 = [&]() {
    if (tid < PT2) {
      float x0 = px0, x1 = px1, x2 = px2;
      const float x3 = px3;
      if (MAIN) {   // bmm([P,3],[3,3]) (:146): out_j = sum_i x_i T[i][j], sequential-i fmaf
        const float* t = p.trans + (size_t)px_b * 9;
        const float y0 = fmaf(x2, __ldg(t + 6), fmaf(x1, __ldg(t + 3), x0 * __ldg(t + 0)));
        const float y1 = fmaf(x2, __ldg(t + 7), fmaf(x1, __ldg(t + 4), x0 * __ldg(t + 1)));
        const float y2 = fmaf(x2, __ldg(t + 8), fmaf(x1, __ldg(t + 5), x0 * __ldg(t + 2)));
        x0 = y0; x1 = y1; x2 = y2;
      }
      const uint32_t h01 = pack_f16x2(x0, x1), h23 = pack_f16x2(x2, x3);
      const uint32_t l01 = pack_f16x2(x0 - f16_lo_as_float(h01), x1 - f16_hi_as_float(h01));
      const uint32_t l23 = pack_f16x2(x2 - f16_lo_as_float(h23), x3 - f16_hi_as_float(h23));
      const uint32_t one = 0x3c00u;
      // columns: 0-3 xh | 4 one | 5-7 xl0..2   ||   8 xl3 | 9-12 xh | 13 one | 14-15 0
      *reinterpret_cast<uint4*>(smem + O2_X16 + tid * 16) = make_uint4(h01, h23, one | (l01 << 16), (l01 >> 16) | (l23 << 16));
      *reinterpret_cast<uint4*>(smem + O2_X16 + PT2 * 16 + tid * 16) =
          make_uint4((l23 >> 16) | (h01 << 16), (h01 >> 16) | (h23 << 16), (h23 >> 16) | (one << 16), 0u);
    }
  };
  // column group of this warp: the compute warps of a TMEM lane quarter split the accumulator columns evenly
  const bool feed = warp < NF2 / 32, drain = compute && !feed;
  const int cg = (warp >> 2) & (CG2 - 1);
  const int prow = (warp & 3) * 32 + lane;   // tile point (layer 1 / 2 epilogues) or chunk channel (layer 3) = TMEM lane
  // `NCOL` accumulator columns starting at `col` -> ReLU -> FP16 pairs -> k-chunks kc0.. of row `prow` of an operand image
  auto drain_relu = [&](uint32_t col, uint8_t* img, int kc0, auto ncol_tag) {
    constexpr int NCOL = decltype(ncol_tag)::value;
    uint8_t* row = img + prow * 16;
#pragma unroll
    for (int c0 = 0; c0 < NCOL; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem + lane_addr + col + (uint32_t)c0, v);
#pragma unroll
      for (int kc = 0; kc < 2; ++kc) {
        uint32_t pk[4];
#pragma unroll
        for (int e = 0; e < 4; ++e)
          pk[e] = pack_relu_f16x2(__uint_as_float(v[kc * 8 + e * 2]), __uint_as_float(v[kc * 8 + e * 2 + 1]));
        *reinterpret_cast<uint4*>(row + (kc0 + c0 / 8 + kc) * PT2 * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      }
    }
  };
  // layer-1 epilogue: thread = (point, 64 / CG2 channels) -> its h1 row
  auto l1_epilogue = [&]() {
    drain_relu(TM_L1 + (uint32_t)(cg * (64 / CG2)), smem + O2_H1, cg * (8 / CG2), std::integral_constant<int, 64 / CG2>());
  };
  // layer-2 epilogue of stream tile s: thread = (point, 128 / CG2 channels) -> its row of h2[s & 1]
  auto l2_epilogue = [&](int s) {
    drain_relu(TM_L2 + (uint32_t)(cg * (128 / CG2)), smem + O2_H2 + (uint32_t)(s & 1) * H2B, cg * (16 / CG2), std::integral_constant<int, 128 / CG2>());
  };
  // layer-3 max epilogue: lane = channel of the chunk, this warp's 128 / CG2 point columns
  auto chunk_max = [&](uint32_t col_base) {
    float m = -INFINITY;
#pragma unroll
    for (int c0 = 0; c0 < 128 / CG2; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld32(tmem + lane_addr + col_base + (uint32_t)(cg * (128 / CG2) + c0), v);
      tc::tmem_ld_wait_dep32(v);
      float a[11];
#pragma unroll
      for (int i = 0; i < 10; ++i) a[i] = max3f(__uint_as_float(v[3 * i]), __uint_as_float(v[3 * i + 1]), __uint_as_float(v[3 * i + 2]));
      a[10] = fmaxf(__uint_as_float(v[30]), __uint_as_float(v[31]));
      m = max3f(m, max3f(a[0], a[1], a[2]), max3f(a[3], a[4], a[5]));
      m = max3f(m, max3f(a[6], a[7], a[8]), fmaxf(a[9], a[10]));
    }
    return m;
  };

  float run0 = -INFINITY, run1 = -INFINITY;   // running maxima of channel q*256 + mc*128 + prow over this warp's column group
  bool ok = true;
  if (compute && S > 0) { load_point(); write_x16(); }
  if (compute && S > 1) load_point();
  int c1_b = group, c1_tile = 0;   // cloud and tile of the next chunk-1 epilogue
  tc::fence_proxy_async_smem();
  __syncthreads();
  for (int r = -2; ok && r <= S; ++r) {
    // ---- the round's tensor work, in queue order ----
    if (tid == NC2) {
      tc::tc_fence_after();
      if (r + 2 < S) {
        tc::umma_f16(tmem + TM_L1, d_x16, d_w1, idesc_l1, 0u);
        tc::umma_commit(&bar_l1);
      }
      if (r + 1 >= 0 && r + 1 < S) {
        uint64_t ad = d_h1, bd = d_w2;
#pragma unroll
        for (int j = 0; j < K1 / 16; ++j) {
          tc::umma_f16(tmem + TM_L2, ad, bd, idesc, j > 0 ? 1u : 0u);
          ad += (uint64_t)((2 * PT2 * 16) >> 4); bd += (uint64_t)((2 * 128 * 16) >> 4);
        }
        tc::umma_commit(&bar_l2);
      }
      if (r >= 0 && r < S) {
        const uint64_t d_h2 = tc::make_smem_desc(tc::smem_u32(smem + O2_H2 + (uint32_t)(r & 1) * H2B), PT2 * 16, 128);
#pragma unroll
        for (int mc = 0; mc < 2; ++mc) {
          if (mc == 1 && r >= 1) {   // chunk 1's accumulator is single-buffered: the maxima of tile r - 1 must be out
            if (!tc::mbar_wait(&bar_c1free, (uint32_t)((r - 1) & 1), errw, 5)) break;
            tc::tc_fence_after();
          }
          uint64_t ad = d_w3 + (uint64_t)((mc * 128 * 16) >> 4), bd = d_h2;
#pragma unroll
          for (int j = 0; j < 128 / 16; ++j) {
            tc::umma_f16(tmem + (mc == 0 ? TM_C0 : TM_C1), ad, bd, idesc, j > 0 ? 1u : 0u);
            ad += (uint64_t)((2 * 256 * 16) >> 4); bd += (uint64_t)((2 * PT2 * 16) >> 4);
          }
          tc::umma_commit(&bar_l3[mc]);
        }
      }
    }
    if (drain) {
      // (i) max-epilogue of chunk 1 of tile r-1 (its MMAs were the last of the previous round)
      if (r >= 1) {
        const int s = r - 1;
        if (!tc::mbar_wait(&bar_l3[1], (uint32_t)(s & 1), errw, 3)) { ok = false; }
        tc::tc_fence_after();
#ifndef DVQ_PN_KO_DRAIN
        if (ok) run1 = fmaxf(run1, chunk_max(TM_C1));
#endif
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bar_c1free);
        const int b = c1_b;
        const bool cloud_done = ++c1_tile == ntiles;
        if (cloud_done) { c1_tile = 0; c1_b += ngroups; }
        if (ok && cloud_done) {
          // the cloud of tile s is complete: combine the column groups, write this quarter's 256 maxima
          if (cg > 0) { mxs[(cg - 1) * 256 + prow] = run0; mxs[(cg - 1) * 256 + 128 + prow] = run1; }
          asm volatile("bar.sync 2, %0;" ::"n"(NF2) : "memory");
          if (cg == 0) {
            float m0 = run0, m1 = run1;
#pragma unroll
            for (int g2 = 0; g2 < CG2 - 1; ++g2) { m0 = fmaxf(m0, mxs[g2 * 256 + prow]); m1 = fmaxf(m1, mxs[g2 * 256 + 128 + prow]); }
            float* out = p.maxbuf + (size_t)b * 1024 + q * 256;
            out[prow] = m0;
            out[128 + prow] = m1;
          }
          asm volatile("bar.sync 2, %0;" ::"n"(NF2) : "memory");
          run0 = -INFINITY; run1 = -INFINITY;
        }
      }
      // (ii) max-epilogue of chunk 0 of tile r
      if (ok && r >= 0 && r < S) {
        if (!tc::mbar_wait(&bar_l3[0], (uint32_t)(r & 1), errw, 2)) { ok = false; }
        tc::tc_fence_after();
#ifndef DVQ_PN_KO_DRAIN
        if (ok) run0 = fmaxf(run0, chunk_max(TM_C0));
#endif
      }
    }
    if (feed) {
      // (a) coordinates of tile r+3 -> x16 (free once the layer-1 MMA of tile r+2, first in this round's queue, completed);
      //     runs while the layer-2 MMAs of tile r+1 execute
      if (r + 2 < S) {
        if (!tc::mbar_wait(&bar_l1, (uint32_t)((r + 2) & 1), errw, 4)) { ok = false; }
        tc::tc_fence_after();
#ifndef DVQ_PN_KO_FEED
        if (ok && r + 3 < S) { write_x16(); if (r + 4 < S) load_point(); }
#endif
      }
      // (b) layer-2 epilogue of tile r+1 -> h2[(r+1)&1]
      if (ok && r + 1 >= 0 && r + 1 < S) {
        if (!tc::mbar_wait(&bar_l2, (uint32_t)((r + 1) & 1), errw, 1)) { ok = false; }
        tc::tc_fence_after();
#ifndef DVQ_PN_KO_FEED
        if (ok) l2_epilogue(r + 1);
#endif
      }
      // (c) layer-1 epilogue of tile r+2 -> h1 (free: the layer-2 MMAs of tile r+1 completed above)
#ifndef DVQ_PN_KO_FEED
      if (ok && r + 2 < S) l1_epilogue();
#endif
    }
    tc::tc_fence_before();
    tc::fence_proxy_async_smem();
    ok = __syncthreads_and(ok ? 1 : 0) != 0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

}  // namespace

size_t pointnet_tc_image_bytes() { return align_up((size_t)W2_BYTES, 256) + 4 * (size_t)W3Q_BYTES; }

int launch_pointnet_tc_trunk(const float* x, const float* trans, const float* w1, const float* b1, const float* w2,
                             const float* b2, const float* w3, int B, int C, int P, float* maxbuf, void* images,
                             bool main_trunk, cudaStream_t s) {
  DeviceProps dp;
  int rc = device_props(&dp);
  if (rc) return rc;
  uint8_t* w2img = static_cast<uint8_t*>(images);
  uint8_t* w3img = w2img + align_up((size_t)W2_BYTES, 256);
  pointnet_pack_kernel<<<64, 256, 0, s>>>(w2, b2, w3, w2img, w3img);
  DVQ_CUDA_CHECK(cudaGetLastError());
  count_launch();
  TcTrunkParams p;
  p.x = x; p.trans = trans; p.w1 = w1; p.b1 = b1; p.w2img = w2img; p.w3img = w3img; p.maxbuf = maxbuf; p.B = B; p.C = C; p.P = P;
  int groups = (dp.sm_count + 3) / 4;
  if (groups > B) groups = B;
  static const bool v1 = getenv("DVQ_PN_TC_V1") != nullptr;   // the 256-point kernel with a CUDA-core layer 1 (no cross-tile overlap), kept for A/B runs
  if (!v1) {
    const size_t smem2 = SMEM2_TOTAL + 128;
    if (main_trunk) {
      DVQ_CUDA_CHECK(cudaFuncSetAttribute(pointnet_trunk_tc2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      pointnet_trunk_tc2_kernel<true><<<groups * 4, NT2, smem2, s>>>(p);
    } else {
      DVQ_CUDA_CHECK(cudaFuncSetAttribute(pointnet_trunk_tc2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
      pointnet_trunk_tc2_kernel<false><<<groups * 4, NT2, smem2, s>>>(p);
    }
    DVQ_CUDA_CHECK(cudaGetLastError());
    count_launch();
    return DVQ_OK;
  }
  const size_t smem = SMEM_TOTAL + 128;
  if (main_trunk) {
    DVQ_CUDA_CHECK(cudaFuncSetAttribute(pointnet_trunk_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pointnet_trunk_tc_kernel<true><<<groups * 4, NT, smem, s>>>(p);
  } else {
    DVQ_CUDA_CHECK(cudaFuncSetAttribute(pointnet_trunk_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pointnet_trunk_tc_kernel<false><<<groups * 4, NT, smem, s>>>(p);
  }
  DVQ_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return DVQ_OK;
}

}  // namespace dvq
