// Exact FP32 re-evaluation of the rows the tcgen05 filter could not decide, restricted to the
// candidate code groups the filter recorded (a bit per group of 32*2^gshift codes that held a key
// inside the row's error band).  One warp per undecided row; lane = code inside a 32-code group.
//
// Arithmetic is bit-identical to vq_simt_fp32.cu: per (row, code) the dot product is a sequential
// fmaf over d = 0..D-1, zz = lane-strided fmaf partials + xor-shuffle tree, d = fma(-2, dot,
// fl(zz + ee_k)), winner = lexicographic minimum of (distance, index) — so a row refined here gets
// exactly the index the all-FP32 kernel would give it (the true minimiser is always inside the
// candidate groups: every key outside the band is provably farther).
//
// The codebook is staged once per CTA in shared memory with a (D+4)-float row pitch: 16-byte aligned
// rows whose 128-bit reads are conflict-free for lane = code (a quarter-warp covers all 32 banks).  The
// z row stays in shared memory too (broadcast 128-bit reads), so a thread needs ~40 registers and 32
// warps per SM hide the latency of the sequential FMA chain; each undecided row then costs ~2 groups
// x (32 LDS.128 + 64 FMAs) per lane instead of the K x D of the full kernel.
#include "dvq_common.cuh"

namespace dvq {
namespace {

constexpr int RTHREADS = 1024;

template <int DT, bool TRAIN, bool SMEM_E>
__global__ void __launch_bounds__(RTHREADS, 1)
vq_refine_kernel(const float* __restrict__ z, const float* __restrict__ E, const float* __restrict__ ee, int K,
                 float* __restrict__ zq, int64_t* __restrict__ idx_out, unsigned long long* __restrict__ hist,
                 double* __restrict__ sse, const int* __restrict__ row_list, const int* __restrict__ cand_list,
                 const int* __restrict__ n_list, int gshift) {
  extern __shared__ __align__(16) float smem_f[];   // zbuf[warps][2][DT] | es[K][DT+1] when SMEM_E
  constexpr int PITCH = DT + 4;
  constexpr int NW = RTHREADS / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* zbuf = smem_f + warp * 2 * DT;             // double-buffered z row of this warp
  float* es = smem_f + NW * 2 * DT;
  if (SMEM_E) {
    for (int i = tid; i < K * (DT / 4); i += RTHREADS) {
      const int k = i / (DT / 4), d = (i - k * (DT / 4)) * 4;
      *reinterpret_cast<float4*>(es + k * PITCH + d) = ldg4(E + (size_t)k * DT + d);
    }
    __syncthreads();
  }
  const int n = *n_list;
  const int wglobal = blockIdx.x * NW + warp;
  const int wtotal = gridDim.x * NW;
  const int groups = ((K + 31) / 32 + (1 << gshift) - 1) >> gshift;   // candidate bits in use
  double lsse = 0.0;
  // cp.async prefetch of the next undecided row of this warp hides its (random-access) global latency
  auto prefetch = [&](int i, int buf) {
    if (i < n && lane < DT / 4) {
      const float* src = z + (int64_t)row_list[i] * DT + lane * 4;
      const uint32_t dst = (uint32_t)__cvta_generic_to_shared(zbuf + buf * DT + lane * 4);
      asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  prefetch(wglobal, 0);
  int buf = 0;
  for (int i = wglobal; i < n; i += wtotal, buf ^= 1) {
    prefetch(i + wtotal, buf ^ 1);
    asm volatile("cp.async.wait_group 1;" ::: "memory");
    __syncwarp();
    const int64_t row = row_list[i];
    const unsigned cand = (unsigned)cand_list[i];
    const float* zb = zbuf + buf * DT;
    float s2 = 0.f;   // ||z||^2 exactly as vq_simt_fp32.cu computes it
#pragma unroll
    for (int c = 0; c < DT; c += 32) {
      if (c + lane < DT) { const float v = zb[c + lane]; s2 = fmaf(v, v, s2); }
    }
    const float zz = warp_sum(s2);
    float best = INFINITY;
    int bidx = 0;
    unsigned todo = cand ? cand : 0xffffffffu;
    if (groups < 32) todo &= (1u << groups) - 1u;
    while (todo) {                       // set bits in ascending order = groups in ascending code order
      const int g = __ffs(todo) - 1;
      todo &= todo - 1u;
      for (int sc = g << gshift; sc < ((g + 1) << gshift) && sc * 32 < K; ++sc) {
        const int code = sc * 32 + lane;
        float dist = INFINITY;
        if (code < K) {
          float acc = 0.f;
          if (SMEM_E) {
            const float* er = es + code * PITCH;
#pragma unroll
            for (int d = 0; d < DT; d += 4) {
              const float4 e4 = *reinterpret_cast<const float4*>(er + d);
              const float4 z4 = *reinterpret_cast<const float4*>(zb + d);   // broadcast read
              acc = fmaf(z4.x, e4.x, acc); acc = fmaf(z4.y, e4.y, acc);
              acc = fmaf(z4.z, e4.z, acc); acc = fmaf(z4.w, e4.w, acc);
            }
          } else {
            const float* er = E + (int64_t)code * DT;
#pragma unroll
            for (int d = 0; d < DT; d += 4) {
              const float4 e4 = ldg4(er + d);
              const float4 z4 = *reinterpret_cast<const float4*>(zb + d);
              acc = fmaf(z4.x, e4.x, acc); acc = fmaf(z4.y, e4.y, acc);
              acc = fmaf(z4.z, e4.z, acc); acc = fmaf(z4.w, e4.w, acc);
            }
          }
          dist = __fmaf_rn(-2.0f, acc, __fadd_rn(zz, __ldg(ee + code)));
        }
        int k = code;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float od = __shfl_xor_sync(0xffffffffu, dist, o);
          const int ok = __shfl_xor_sync(0xffffffffu, k, o);
          if (od < dist || (od == dist && ok < k)) { dist = od; k = ok; }
        }
        if (dist < best) { best = dist; bidx = k; }   // groups visited in ascending code order
      }
    }
    // outputs for this row (vq_simt_fp32.cu's epilogue)
    float* orow = zq + row * DT;
    const float* erow = E + (int64_t)bidx * DT;
    for (int c = lane * 4; c < DT; c += 128) {
      const float4 e4 = ldg4(erow + c);
      float4 o4 = e4;
      if (TRAIN) {
        const float4 z4 = *reinterpret_cast<const float4*>(zbuf + buf * DT + c);
        const float dx = __fsub_rn(e4.x, z4.x), dy = __fsub_rn(e4.y, z4.y);
        const float dz = __fsub_rn(e4.z, z4.z), dw = __fsub_rn(e4.w, z4.w);
        lsse += (double)dx * dx + (double)dy * dy + (double)dz * dz + (double)dw * dw;
        o4 = make_float4(__fadd_rn(z4.x, dx), __fadd_rn(z4.y, dy), __fadd_rn(z4.z, dz), __fadd_rn(z4.w, dw));
      }
      *reinterpret_cast<float4*>(orow + c) = o4;
    }
    if (lane == 0) {
      idx_out[row] = (int64_t)bidx;
      if (TRAIN) atomicAdd(hist + bidx, 1ull);
    }
    __syncwarp();   // everyone is done with zbuf[buf] before the next-but-one prefetch overwrites it
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  if (TRAIN) {
    lsse = warp_sum(lsse);
    if (lane == 0 && lsse != 0.0) atomicAdd(sse, lsse);
  }
}

}  // namespace

bool vq_refine_supported(int K, int D) { return D == 64 && K >= 32; }

int launch_vq_refine(const float* z, const float* E, const float* ee, int K, int D, int train, float* z_q, int64_t* idx,
                     unsigned long long* hist, double* sse, const int* row_list, const int* cand_list,
                     const int* n_list, int gshift, cudaStream_t s) {
  DeviceProps dp;
  int rc = device_props(&dp);
  if (rc) return rc;
  if (D != 64) return fail(DVQ_ERR_BAD_SHAPE, "candidate refine kernel is instantiated for e_dim 64 only");
  const size_t zbuf_bytes = (size_t)(RTHREADS / 32) * 2 * 64 * sizeof(float);
  const size_t smem_e = (size_t)K * (64 + 4) * sizeof(float);
  const bool in_smem = smem_e + zbuf_bytes <= 200 * 1024;
#define DVQ_LAUNCH_REFINE(TR_, SM_)                                                                                      \
  do {                                                                                                                   \
    const size_t bytes = zbuf_bytes + (SM_ ? smem_e : 0);                                                                            \
    DVQ_CUDA_CHECK(cudaFuncSetAttribute(vq_refine_kernel<64, TR_, SM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); \
    vq_refine_kernel<64, TR_, SM_><<<dp.sm_count, RTHREADS, bytes, s>>>(z, E, ee, K, z_q, idx, hist, sse, row_list,      \
                                                                       cand_list, n_list, gshift);                      \
  } while (0)
  if (train) { if (in_smem) DVQ_LAUNCH_REFINE(true, true); else DVQ_LAUNCH_REFINE(true, false); }
  else       { if (in_smem) DVQ_LAUNCH_REFINE(false, true); else DVQ_LAUNCH_REFINE(false, false); }
#undef DVQ_LAUNCH_REFINE
  DVQ_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return DVQ_OK;
}

}  // namespace dvq
