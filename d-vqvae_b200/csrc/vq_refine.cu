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
                 const int* __restrict__ n_list, int gshift, const int* __restrict__ ovf_last, const int* __restrict__ n_ovf) {
  extern __shared__ __align__(16) float smem_f[];   // zbuf[warps][2][DT] | es[K][DT+1] when SMEM_E
  constexpr int PITCH = DT + 4;
  constexpr int NW = RTHREADS / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* zbuf = smem_f + warp * 2 * DT;             // double-buffered z row of this warp
  float* es = smem_f + NW * 2 * DT;
  // nothing listed (the usual case when the binned kernels ran first): leave before staging the codebook
  if (*n_list == 0 && (n_ovf == nullptr || *n_ovf == 0)) return;
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
  const bool list_mode = gshift < 0;   // candidates as up to three (sub-chunk index + 1) entries of 10 bits, see vq_tc_sm100.cu
  const int nsc = (K + 31) / 32;
  const int groups = list_mode ? 0 : (nsc + (1 << gshift) - 1) >> gshift;   // candidate bits in use
  double lsse = 0.0;
  // cp.async prefetch of the next undecided row of this warp hides its (random-access) global latency
  auto prefetch = [&](int i, int buf) {
    if (i < n) {
      const float* src = z + (int64_t)row_list[i] * DT;
#pragma unroll
      for (int v = lane; v < DT / 4; v += 32) {
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(zbuf + buf * DT + v * 4);
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src + v * 4) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  // outputs of one row (vq_simt_fp32.cu's epilogue), executed by one warp
  auto emit_row = [&](int64_t row, int bidx, const float* zsrc) {
    float* orow = zq + row * DT;
    const float* erow = E + (int64_t)bidx * DT;
    for (int c = lane * 4; c < DT; c += 128) {
      const float4 e4 = ldg4(erow + c);
      float4 o4 = e4;
      if (TRAIN) {
        const float4 z4 = *reinterpret_cast<const float4*>(zsrc + c);
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
    // candidate sub-chunks (32 codes each) in ascending code order, two at a time: one broadcast read of z
    // feeds both dot products, and the two sequential FMA chains interleave
    unsigned todo = 0u;
    int g_sc = 0, g_end = 0, l0 = -1, l1 = -1, l2 = -1;
    if (list_mode) {
      if (cand == 0u || cand == 0x3fffffffu) { g_end = nsc; }   // overflow / degenerate: every code
      else {
        l0 = (int)(cand & 1023u) - 1; l1 = (int)((cand >> 10) & 1023u) - 1; l2 = (int)((cand >> 20) & 1023u) - 1;
        // ascending order, empty entries (-1) last
        const int big = 1 << 30;
        int a0 = l0 < 0 ? big : l0, a1 = l1 < 0 ? big : l1, a2 = l2 < 0 ? big : l2;
        int t;
        if (a0 > a1) { t = a0; a0 = a1; a1 = t; }
        if (a1 > a2) { t = a1; a1 = a2; a2 = t; }
        if (a0 > a1) { t = a0; a0 = a1; a1 = t; }
        l0 = a0 == big ? -1 : a0; l1 = a1 == big ? -1 : a1; l2 = a2 == big ? -1 : a2;
      }
    } else {
      todo = cand ? cand : 0xffffffffu;
      if (groups < 32) todo &= (1u << groups) - 1u;
    }
    auto next_sc = [&]() -> int {   // warp-uniform; -1 when exhausted
      for (;;) {
        if (g_sc < g_end) {
          const int r = g_sc++;
          if (r * 32 < K) return r;
          g_sc = g_end;
          continue;
        }
        if (list_mode) {
          const int r = l0;
          l0 = l1; l1 = l2; l2 = -1;
          return (r >= 0 && r * 32 < K) ? r : -1;
        }
        if (!todo) return -1;
        const int g = __ffs(todo) - 1;
        todo &= todo - 1u;
        g_sc = g << gshift;
        g_end = (g + 1) << gshift;
      }
    };
    // lexicographic (distance, code) minimum over the warp: monotone integer image of the distance, one
    // REDUX.MIN, then the lowest lane holding the minimum (= lowest code: exact ties keep the first index)
    auto warp_argmin = [&](float dist, int sc, float& best_, int& bidx_) {
      const uint32_t b = __float_as_uint(dist);
      const uint32_t key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);   // NaN-free input: order-preserving
      const uint32_t kmin = __reduce_min_sync(0xffffffffu, key);
      const unsigned who = __ballot_sync(0xffffffffu, key == kmin);
      const int src = __ffs(who) - 1;
      const float dmin = __shfl_sync(0xffffffffu, dist, src);
      if (dmin < best_) { best_ = dmin; bidx_ = sc * 32 + src; }
    };
    for (;;) {
      const int sa = next_sc();
      if (sa < 0) break;
      const int sb = next_sc();
      const int code_a = sa * 32 + lane, code_b = (sb < 0 ? sa : sb) * 32 + lane;
      const bool va = code_a < K, vb = sb >= 0 && code_b < K;
      float acc_a = 0.f, acc_b = 0.f;
      if (SMEM_E && sb < 0) {
        // a single sub-chunk left (about half of the rows have an odd number of candidate sub-chunks): one FMA chain, half the
        // shared-memory reads of the paired pass
        const float* ea = es + (va ? code_a : 0) * PITCH;
#pragma unroll
        for (int d = 0; d < DT; d += 4) {
          const float4 z4 = *reinterpret_cast<const float4*>(zb + d);   // broadcast read
          const float4 a4 = *reinterpret_cast<const float4*>(ea + d);
          acc_a = fmaf(z4.x, a4.x, acc_a); acc_a = fmaf(z4.y, a4.y, acc_a);
          acc_a = fmaf(z4.z, a4.z, acc_a); acc_a = fmaf(z4.w, a4.w, acc_a);
        }
      } else if (SMEM_E) {
        const float* ea = es + (va ? code_a : 0) * PITCH;
        const float* eb = es + (vb ? code_b : 0) * PITCH;
#pragma unroll
        for (int d = 0; d < DT; d += 4) {
          const float4 z4 = *reinterpret_cast<const float4*>(zb + d);   // broadcast read
          const float4 a4 = *reinterpret_cast<const float4*>(ea + d);
          const float4 b4 = *reinterpret_cast<const float4*>(eb + d);
          acc_a = fmaf(z4.x, a4.x, acc_a); acc_b = fmaf(z4.x, b4.x, acc_b);
          acc_a = fmaf(z4.y, a4.y, acc_a); acc_b = fmaf(z4.y, b4.y, acc_b);
          acc_a = fmaf(z4.z, a4.z, acc_a); acc_b = fmaf(z4.z, b4.z, acc_b);
          acc_a = fmaf(z4.w, a4.w, acc_a); acc_b = fmaf(z4.w, b4.w, acc_b);
        }
      } else {
        const float* ea = E + (int64_t)(va ? code_a : 0) * DT;
        const float* eb = E + (int64_t)(vb ? code_b : 0) * DT;
#pragma unroll
        for (int d = 0; d < DT; d += 4) {
          const float4 z4 = *reinterpret_cast<const float4*>(zb + d);
          const float4 a4 = ldg4(ea + d);
          const float4 b4 = ldg4(eb + d);
          acc_a = fmaf(z4.x, a4.x, acc_a); acc_b = fmaf(z4.x, b4.x, acc_b);
          acc_a = fmaf(z4.y, a4.y, acc_a); acc_b = fmaf(z4.y, b4.y, acc_b);
          acc_a = fmaf(z4.z, a4.z, acc_a); acc_b = fmaf(z4.z, b4.z, acc_b);
          acc_a = fmaf(z4.w, a4.w, acc_a); acc_b = fmaf(z4.w, b4.w, acc_b);
        }
      }
      const float dist_a = va ? __fmaf_rn(-2.0f, acc_a, __fadd_rn(zz, __ldg(ee + code_a))) : INFINITY;
      warp_argmin(dist_a, sa, best, bidx);
      if (sb >= 0) {
        const float dist_b = vb ? __fmaf_rn(-2.0f, acc_b, __fadd_rn(zz, __ldg(ee + code_b))) : INFINITY;
        warp_argmin(dist_b, sb, best, bidx);
      }
    }
    emit_row(row, bidx, zbuf + buf * DT);
    __syncwarp();   // everyone is done with zbuf[buf] before the next-but-one prefetch overwrites it
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  // ---- overflow rows (list mode: more than three candidate sub-chunks, degenerate rows): one CTA per row, the
  //      warps split the sub-chunks of the whole codebook, lexicographic (distance, code) minimum across warps ----
  if (n_ovf != nullptr) {
    __shared__ float red_d[NW];
    __shared__ int red_k[NW];
    const int n2 = *n_ovf;
    float* zrow = smem_f;   // every warp is done with its z buffers after the barrier below
    for (int j = blockIdx.x; j < n2; j += gridDim.x) {
      __syncthreads();
      const int64_t row = ovf_last[-(int64_t)j];   // the list is stored downwards from the end of the buffer
      for (int v = tid; v < DT / 4; v += RTHREADS) reinterpret_cast<float4*>(zrow)[v] = ldg4(z + row * DT + v * 4);
      __syncthreads();
      float s2 = 0.f;
#pragma unroll
      for (int c = 0; c < DT; c += 32) {
        if (c + lane < DT) { const float v = zrow[c + lane]; s2 = fmaf(v, v, s2); }
      }
      const float zz = warp_sum(s2);
      float best = INFINITY;
      int bidx = 0x7fffffff;
      for (int sc = warp; sc < nsc; sc += NW) {
        const int code = sc * 32 + lane;
        float dist = INFINITY;
        if (code < K) {
          const float* er = SMEM_E ? es + code * PITCH : E + (int64_t)code * DT;
          float acc = 0.f;
#pragma unroll 8
          for (int d = 0; d < DT; d += 4) {
            const float4 z4 = *reinterpret_cast<const float4*>(zrow + d);
            const float4 e4 = SMEM_E ? *reinterpret_cast<const float4*>(er + d) : ldg4(er + d);
            acc = fmaf(z4.x, e4.x, acc); acc = fmaf(z4.y, e4.y, acc);
            acc = fmaf(z4.z, e4.z, acc); acc = fmaf(z4.w, e4.w, acc);
          }
          dist = __fmaf_rn(-2.0f, acc, __fadd_rn(zz, __ldg(ee + code)));
        }
        const uint32_t b = __float_as_uint(dist);
        const uint32_t key = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
        const uint32_t kmin = __reduce_min_sync(0xffffffffu, key);
        const int src = __ffs(__ballot_sync(0xffffffffu, key == kmin)) - 1;
        const float dmin = __shfl_sync(0xffffffffu, dist, src);
        if (dmin < best) { best = dmin; bidx = sc * 32 + src; }
      }
      if (lane == 0) { red_d[warp] = best; red_k[warp] = bidx; }
      __syncthreads();
      if (warp == 0) {
        float d = red_d[lane];
        int k = red_k[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const float od = __shfl_xor_sync(0xffffffffu, d, o);
          const int ok = __shfl_xor_sync(0xffffffffu, k, o);
          if (od < d || (od == d && ok < k)) { d = od; k = ok; }
        }
        if (k == 0x7fffffff) k = 0;
        emit_row(row, k, zrow);
      }
    }
  }
  if (TRAIN) {
    lsse = warp_sum(lsse);
    if (lane == 0 && lsse != 0.0) atomicAdd(sse, lsse);
  }
}

}  // namespace

// true when the per-row kernel can keep the whole FP32 codebook in shared memory (its fast regime)
bool vq_refine_codebook_in_smem(int K, int D) {
  return (size_t)K * (D + 4) * sizeof(float) + (size_t)(RTHREADS / 32) * 2 * D * sizeof(float) <= 200 * 1024;
}

bool vq_refine_supported(int K, int D) { return (D == 16 || D == 32 || D == 64 || D == 128 || D == 256 || D == 512) && K >= 32; }

template <int DT>
static int launch_refine_dt(const float* z, const float* E, const float* ee, int K, int train, float* z_q, int64_t* idx,
                            unsigned long long* hist, double* sse, const int* row_list, const int* cand_list,
                            const int* n_list, int gshift, const int* ovf_last, const int* n_ovf, int sm_count, cudaStream_t s) {
  const size_t zbuf_bytes = (size_t)(RTHREADS / 32) * 2 * DT * sizeof(float);
  const size_t smem_e = (size_t)K * (DT + 4) * sizeof(float);
  const bool in_smem = vq_refine_codebook_in_smem(K, DT);
#define DVQ_LAUNCH_REFINE(TR_, SM_)                                                                                      \
  do {                                                                                                                   \
    const size_t bytes = zbuf_bytes + (SM_ ? smem_e : 0);                                                                \
    DVQ_CUDA_CHECK(cudaFuncSetAttribute(vq_refine_kernel<DT, TR_, SM_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes)); \
    vq_refine_kernel<DT, TR_, SM_><<<sm_count, RTHREADS, bytes, s>>>(z, E, ee, K, z_q, idx, hist, sse, row_list,         \
                                                                      cand_list, n_list, gshift, ovf_last, n_ovf);      \
  } while (0)
  if (train) { if (in_smem) DVQ_LAUNCH_REFINE(true, true); else DVQ_LAUNCH_REFINE(true, false); }
  else       { if (in_smem) DVQ_LAUNCH_REFINE(false, true); else DVQ_LAUNCH_REFINE(false, false); }
#undef DVQ_LAUNCH_REFINE
  DVQ_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return DVQ_OK;
}

int launch_vq_refine(const float* z, const float* E, const float* ee, int K, int D, int train, float* z_q, int64_t* idx,
                     unsigned long long* hist, double* sse, const int* row_list, const int* cand_list,
                     const int* n_list, int gshift, const int* ovf_last, const int* n_ovf, cudaStream_t s) {
  DeviceProps dp;
  int rc = device_props(&dp);
  if (rc) return rc;
  switch (D) {
    case 16: return launch_refine_dt<16>(z, E, ee, K, train, z_q, idx, hist, sse, row_list, cand_list, n_list, gshift, ovf_last, n_ovf, dp.sm_count, s);
    case 32: return launch_refine_dt<32>(z, E, ee, K, train, z_q, idx, hist, sse, row_list, cand_list, n_list, gshift, ovf_last, n_ovf, dp.sm_count, s);
    case 64: return launch_refine_dt<64>(z, E, ee, K, train, z_q, idx, hist, sse, row_list, cand_list, n_list, gshift, ovf_last, n_ovf, dp.sm_count, s);
    case 128: return launch_refine_dt<128>(z, E, ee, K, train, z_q, idx, hist, sse, row_list, cand_list, n_list, gshift, ovf_last, n_ovf, dp.sm_count, s);
    case 256: return launch_refine_dt<256>(z, E, ee, K, train, z_q, idx, hist, sse, row_list, cand_list, n_list, gshift, ovf_last, n_ovf, dp.sm_count, s);
    case 512: return launch_refine_dt<512>(z, E, ee, K, train, z_q, idx, hist, sse, row_list, cand_list, n_list, gshift, ovf_last, n_ovf, dp.sm_count, s);
    default: return fail(DVQ_ERR_BAD_SHAPE, "candidate refine kernel is instantiated for e_dim 16..512 (powers of two) only");
  }
}

}  // namespace dvq
