// Binned exact refine of the rows the tcgen05 filter could not decide.
//
// vq_refine.cu gives every undecided row its own warp (lane = code of a candidate sub-chunk), so each
// (row, sub-chunk) pair re-reads 32 code rows from shared memory: 2 x 512-byte reads per 8 FMAs, bound by
// the shared-memory pipe (and by L2 when the codebook does not fit).  Here the (row, sub-chunk) pairs are
// bucketed by sub-chunk first; a warp then keeps the 32 code rows of ONE sub-chunk in registers and streams
// the z rows of that bucket past them (broadcast shared-memory reads, 1 read per 4 FMAs): the FP32 FMA pipe
// is the bound and a code row is fetched once per bucket slice instead of once per row.
//
//   refine_prep_kernel     one warp per listed row: ||z||^2 (stashed in the row's still unwritten z_q slot),
//                          idx[row] := ~0 (the 64-bit (distance, code) minimum is taken in place), pair counts
//                          per sub-chunk (shared-memory histogram per CTA, one global atomic per bin per CTA)
//   refine_scatter_kernel  exclusive scan of the counts (every CTA, in shared memory), CTA-local counting +
//                          one reservation per (CTA, bin), pairs[pos] = listed-row index; work-item table
//   refine_pairs_kernel    one work item = (sub-chunk, slice of its bucket): exact distances of 16 rows at a
//                          time against the warp's 32 codes, warp argmin, atomicMin on idx[row]
//   refine_emit_kernel     one warp per listed row: winner -> idx, z_q, SSE, histogram
//
// Arithmetic is bit-identical to vq_simt_fp32.cu / vq_refine.cu: per (row, code) a sequential fmaf over
// d = 0..D-1, zz = lane-strided fmaf partials + xor-shuffle tree, dist = fma(-2, dot, fl(zz + ee_k)), winner =
// lexicographic minimum of (distance, code) — here as an unsigned 64-bit minimum of
// (order-preserving image of the distance) << 32 | code, so the order in which pairs are processed is irrelevant.
//
// Rows whose candidate record says "every code" (degenerate rows; list mode: more than three candidate
// sub-chunks) are paired with every sub-chunk.  If the pairs do not fit the workspace (2 N entries) the
// scatter kernel hands the whole list to vq_refine.cu instead (device-side switch, no host sync).
#include "dvq_common.cuh"

namespace dvq {
namespace {

constexpr int PREP_THREADS = 1024;
constexpr int PAIR_THREADS = 384;
constexpr int RB = 16;                 // rows per batch of the pairs kernel
constexpr int MAX_BINS = 1024;         // K <= 32 768
constexpr uint32_t CAND_ALL_LIST = 0x3fffffffu;

// counters[] slots used here (zeroed per call by tc_cb_stats_kernel): [0] listed rows, [2] overflow rows,
// [4] total pairs, [5] / [6] rows / overflow rows handed to the fallback kernel
enum { C_N = 0, C_NOVF = 2, C_PAIRS = 4, C_FB_N = 5, C_FB_NOVF = 6 };

struct BinTables {        // int arrays in the workspace
  int* count;             // [MAX_BINS]   pairs per sub-chunk            (zeroed per call)
  int* cursor;            // [MAX_BINS]   scatter reservations           (zeroed per call)
  int* start;             // [MAX_BINS+1] exclusive scan of count
  int* item_start;        // [MAX_BINS+1] exclusive scan of ceil(count / ch); [MAX_BINS+1] = ch
};

__device__ __forceinline__ int64_t listed_row(int i, int n, const int* __restrict__ row_list, const int* __restrict__ ovf_last) {
  return i < n ? row_list[i] : ovf_last[-(int64_t)(i - n)];   // the overflow list is stored downwards from the end
}

// candidate sub-chunks of a listed row, spread over a group of `gs` lanes (gl = lane inside the group)
template <typename F>
__device__ __forceinline__ void for_each_candidate(unsigned cand, bool all, bool list_mode, int nbins, int gl, int gs, F&& f) {
  if (all) {
    for (int b = gl; b < nbins; b += gs) f(b);
  } else if (list_mode) {
    for (int k = gl; k < 3; k += gs) {
      const int e = (int)((cand >> (10 * k)) & 1023u) - 1;
      if (e >= 0 && e < nbins) f(e);
    }
  } else {
    for (int b = gl; b < 31 && b < nbins; b += gs)
      if ((cand >> b) & 1u) f(b);
  }
}
__device__ __forceinline__ bool cand_is_all(unsigned cand, bool listed_normal, bool list_mode) {
  if (!listed_normal) return true;                       // overflow list
  if (list_mode) return cand == 0u || cand == CAND_ALL_LIST;
  return cand == 0u || cand == 0x7fffffffu;
}

__global__ void __launch_bounds__(PREP_THREADS, 1)
refine_prep_kernel(const float* __restrict__ z, int D, int K, float* __restrict__ zq, unsigned long long* __restrict__ idx64,
                   const int* __restrict__ row_list, const int* __restrict__ cand_list, const int* __restrict__ ovf_last,
                   int* __restrict__ counters, BinTables bt, int list_mode) {
  __shared__ int s_cnt[MAX_BINS];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = counters[C_N], n2 = list_mode ? counters[C_NOVF] : 0;
  if (n + n2 == 0) return;
  const int nbins = (K + 31) / 32;
  for (int b = tid; b < nbins; b += PREP_THREADS) s_cnt[b] = 0;
  __syncthreads();
  // Four rows per warp (8 lanes each) so that four independent row fetches are in flight per warp.  ||z||^2 is
  // bit-identical to vq_simt_fp32.cu's (lane l of a 32-lane warp accumulates elements l, l + 32, ... with fmaf,
  // then an xor-shuffle tree): lane j of a group plays the virtual lanes j, j + 8, j + 16, j + 24; the first two
  // tree levels (xor 16, xor 8) are in-thread sums of the same operand pairs, the last three are shuffles.
  const int wtotal = gridDim.x * (PREP_THREADS / 32);
  const int g = lane >> 3, j = lane & 7;
  int my_pairs = 0;
  for (int base = (blockIdx.x * (PREP_THREADS / 32) + warp) * 4; base < n + n2; base += wtotal * 4) {
    const int i = base + g;
    const bool active = i < n + n2;
    const int64_t row = active ? listed_row(i, n, row_list, ovf_last) : 0;
    const unsigned cand = (active && i < n) ? (unsigned)cand_list[i] : 0u;
    const float* zr = z + row * D;
    float pm[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < D; c += 32) {
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        const int d = c + j + 8 * m;
        if (d < D) { const float v = __ldg(zr + d); pm[m] = fmaf(v, v, pm[m]); }
      }
    }
    float zz = (pm[0] + pm[2]) + (pm[1] + pm[3]);
    zz += __shfl_xor_sync(0xffffffffu, zz, 4);
    zz += __shfl_xor_sync(0xffffffffu, zz, 2);
    zz += __shfl_xor_sync(0xffffffffu, zz, 1);
    if (active) {
      if (j == 0) {
        zq[row * D] = zz;              // the row's z_q is rewritten by the emit kernel
        idx64[row] = ~0ull;
      }
      const bool all = cand_is_all(cand, i < n, list_mode != 0);
      for_each_candidate(cand, all, list_mode != 0, nbins, j, 8, [&](int b) { atomicAdd(&s_cnt[b], 1); ++my_pairs; });
    }
  }
  __syncthreads();
  for (int b = tid; b < nbins; b += PREP_THREADS) {
    const int c = s_cnt[b];
    if (c) atomicAdd(bt.count + b, c);
  }
  my_pairs = (int)warp_sum((float)my_pairs);   // < 2^24 per warp: exact in FP32
  if (lane == 0 && my_pairs) atomicAdd(counters + C_PAIRS, my_pairs);
}

// exclusive scan of v[0..n) (n <= MAX_BINS) by a 1024-thread block; out[n] = total
__device__ __forceinline__ void block_scan_1024(int v, int n, int* out, int* s_warp) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    int w = s_warp[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    s_warp[lane] = w;   // inclusive over warps
  }
  __syncthreads();
  const int base = warp ? s_warp[warp - 1] : 0;
  if (tid < n) out[tid] = base + inc - v;
  if (tid == n - 1) out[n] = base + inc;
  __syncthreads();
}

__global__ void __launch_bounds__(PREP_THREADS, 1)
refine_scatter_kernel(int K, const int* __restrict__ cand_list, int* __restrict__ counters, BinTables bt, int* __restrict__ pairs,
                      long long pair_cap, int list_mode, int pair_warps) {
  __shared__ int s_start[MAX_BINS + 1];
  __shared__ int s_items[MAX_BINS + 1];
  __shared__ int s_cnt[MAX_BINS];
  __shared__ int s_base[MAX_BINS];
  __shared__ int s_warp[32];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = counters[C_N], n2 = list_mode ? counters[C_NOVF] : 0;
  if (n + n2 == 0) return;
  const long long total = counters[C_PAIRS];
  if (total > pair_cap || total >= (1ll << 30)) {   // does not fit: the per-row kernel takes the whole list
    if (blockIdx.x == 0 && tid == 0) { counters[C_FB_N] = n; counters[C_FB_NOVF] = n2; }
    return;
  }
  const int nbins = (K + 31) / 32;
  const int cnt = tid < nbins ? bt.count[tid] : 0;
  block_scan_1024(cnt, nbins, s_start, s_warp);
  // bucket slice per work item: about two items per warp of the pairs kernel, a multiple of the batch size
  int ch = (int)((total + 2ll * pair_warps - 1) / (2ll * pair_warps));
  ch = (ch + RB - 1) / RB * RB;
  ch = ch < RB ? RB : (ch > 4096 ? 4096 : ch);
  block_scan_1024((cnt + ch - 1) / ch, nbins, s_items, s_warp);
  if (blockIdx.x == 0) {
    for (int b = tid; b <= nbins; b += PREP_THREADS) { bt.start[b] = s_start[b]; bt.item_start[b] = s_items[b]; }
    if (tid == 0) bt.item_start[MAX_BINS + 1] = ch;
  }
  for (int b = tid; b < nbins; b += PREP_THREADS) s_cnt[b] = 0;
  __syncthreads();
  // this CTA's contiguous share of the listed rows: count, reserve, place.  One thread per row (a row has two or
  // three candidates); rows paired with every sub-chunk are spread over the warp.
  const int per = (n + n2 + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per, i1 = min(n + n2, i0 + per);
  auto sweep = [&](auto&& f) {
    for (int ib = i0 + warp * 32; ib < i1; ib += PREP_THREADS) {
      const int i = ib + lane;
      const bool active = i < i1;
      const unsigned cand = (active && i < n) ? (unsigned)cand_list[i] : 0u;
      const bool all = active && cand_is_all(cand, i < n, list_mode != 0);
      if (active && !all) for_each_candidate(cand, false, list_mode != 0, nbins, 0, 1, [&](int b) { f(b, i); });
      unsigned am = __ballot_sync(0xffffffffu, all);
      while (am) {
        const int ii = ib + __ffs(am) - 1;
        am &= am - 1u;
        for (int b = lane; b < nbins; b += 32) f(b, ii);
      }
    }
  };
  sweep([&](int b, int) { atomicAdd(&s_cnt[b], 1); });
  __syncthreads();
  for (int b = tid; b < nbins; b += PREP_THREADS) {
    const int c = s_cnt[b];
    s_base[b] = c ? s_start[b] + atomicAdd(bt.cursor + b, c) : 0;
    s_cnt[b] = 0;
  }
  __syncthreads();
  sweep([&](int b, int i) { pairs[s_base[b] + atomicAdd(&s_cnt[b], 1)] = i; });
}

// order-preserving unsigned image of a float (NaN-free input)
__device__ __forceinline__ uint32_t float_key(float d) {
  const uint32_t b = __float_as_uint(d);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

template <int DS>
__global__ void __launch_bounds__(PAIR_THREADS, DS <= 32 ? 2 : 1)
refine_pairs_kernel(const float* __restrict__ z, const float* __restrict__ E, const float* __restrict__ ee, int K, int D,
                    const float* __restrict__ zq, unsigned long long* __restrict__ idx64, const int* __restrict__ row_list,
                    const int* __restrict__ ovf_last, const int* __restrict__ pairs, const int* __restrict__ counters,
                    BinTables bt, long long pair_cap, int list_mode) {
  extern __shared__ __align__(16) float smem_f[];   // zs[warps][2][RB][DS]
  __shared__ int s_start[MAX_BINS + 1];
  __shared__ int s_items[MAX_BINS + 1];
  constexpr int NW = PAIR_THREADS / 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = counters[C_N], n2 = list_mode ? counters[C_NOVF] : 0;
  const long long total = counters[C_PAIRS];
  if (n + n2 == 0 || total == 0 || total > pair_cap || total >= (1ll << 30)) return;
  const int nbins = (K + 31) / 32;
  for (int b = tid; b <= nbins; b += PAIR_THREADS) { s_start[b] = bt.start[b]; s_items[b] = bt.item_start[b]; }
  __syncthreads();
  const int ch = bt.item_start[MAX_BINS + 1];
  const int items = s_items[nbins];
  float* zs = smem_f + warp * 2 * RB * DS;   // two halves: one being read, one being filled
  const int ns = D / DS;
  // sliced e_dim: the 32 code rows' slices are staged too ([32][DS + 4] floats, single-buffered: filled by coalesced
  // cp.async for the next step as soon as every lane has copied its own row into registers — conflict-free 128-bit loads —,
  // i.e. under this step's FMAs).
  // Each lane fetching its own row straight from global memory — 32 different lines per load instruction — took two
  // thirds of the LSU data pipe, which the ncu capture showed to be the kernel's bound (74 % busy, FMA pipe 24 %).
  constexpr int EP = DS + 4;
  float* es = smem_f + NW * 2 * RB * DS + warp * 32 * EP;
  for (int item = blockIdx.x * NW + warp; item < items; item += gridDim.x * NW) {
    // bucket of this item: the last b with item_start[b] <= item
    int lo = 0, hi = nbins;   // invariant: s_items[lo] <= item < s_items[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (s_items[mid] <= item) lo = mid; else hi = mid;
    }
    const int b = lo;
    const int p0 = s_start[b] + (item - s_items[b]) * ch;
    const int p1 = min(p0 + ch, s_start[b + 1]);
    const int code = b * 32 + lane;
    const bool valid = code < K;
    const float* erow = E + (size_t)(valid ? code : 0) * D;
    const float eek = valid ? __ldg(ee + code) : 0.f;
    float e[DS];
    if (ns == 1) {
#pragma unroll
      for (int d = 0; d < DS; d += 4) {
        const float4 v = ldg4(erow + d);
        e[d] = v.x; e[d + 1] = v.y; e[d + 2] = v.z; e[d + 3] = v.w;
      }
    }
    // Software pipeline over the (batch, slice) steps of the item: the z rows of the next step are copied into the
    // other half of the warp's staging buffer with cp.async (zero-filled for missing rows), and the
    // (row, ||z||^2) records of the next batch are read, while the current step is computed.
    constexpr int ZL = RB * (DS / 4) / 32;   // 16-byte pieces per lane of one staged slice
    auto fetch_rows = [&](int p, int64_t& row_, float& zz_) {
      row_ = -1; zz_ = 0.f;
      if (p < p1 && lane < min(RB, p1 - p)) {
        row_ = listed_row(pairs[p + lane], n, row_list, ovf_last);
        zz_ = zq[row_ * D];
      }
    };
    auto stage_e = [&](int sl) {   // (no commit of its own)
#pragma unroll
      for (int k = 0; k < DS / 4; ++k) {
        const int f = k * 32 + lane;
        const int cl = f / (DS / 4), c4 = f - cl * (DS / 4);
        const bool cv = b * 32 + cl < K;
        const float* src = E + (size_t)(cv ? b * 32 + cl : 0) * D + sl * DS + c4 * 4;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(es + cl * EP + c4 * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(cv ? 16 : 0) : "memory");
      }
    };
    auto stage_z = [&](int64_t row_, int sl, int buf) {
#pragma unroll
      for (int k = 0; k < ZL; ++k) {
        const int f = k * 32 + lane;
        const int r = f / (DS / 4), c4 = f - r * (DS / 4);
        const long long rr = __shfl_sync(0xffffffffu, (long long)row_, r);
        const float* src = rr >= 0 ? z + rr * D + sl * DS + c4 * 4 : z;
        const uint32_t dst = (uint32_t)__cvta_generic_to_shared(zs + buf * RB * DS + r * DS + c4 * 4);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(rr >= 0 ? 16 : 0) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    int64_t row, row_n;
    float zz, zz_n;
    int buf = 0;
    fetch_rows(p0, row, zz);
    if (ns > 1) stage_e(0);   // same group as the first z rows
    stage_z(row, 0, 0);
    for (int p = p0; p < p1; p += RB) {
      const int nb = min(RB, p1 - p);
      fetch_rows(p + RB, row_n, zz_n);
      float acc[RB];
#pragma unroll
      for (int r = 0; r < RB; ++r) acc[r] = 0.f;
      for (int sl = 0; sl < ns; ++sl) {
        __syncwarp();   // the previous step's reads of the other half are done
        if (sl + 1 < ns) stage_z(row, sl + 1, buf ^ 1);
        else if (p + RB < p1) stage_z(row_n, 0, buf ^ 1);
        else asm volatile("cp.async.commit_group;" ::: "memory");   // keep one group per step
        asm volatile("cp.async.wait_group 1;" ::: "memory");         // this step's rows (and code slices) have landed
        __syncwarp();
        if (ns > 1) {
          const float* eb = es + lane * EP;
#pragma unroll
          for (int d = 0; d < DS; d += 4) {
            const float4 v = *reinterpret_cast<const float4*>(eb + d);
            e[d] = v.x; e[d + 1] = v.y; e[d + 2] = v.z; e[d + 3] = v.w;
          }
          __syncwarp();   // every lane has its row: the buffer takes the next step's slices (a group of its own, older than
                          // the next step's z group, so that step's wait_group 1 covers it)
          if (sl + 1 < ns || p + RB < p1) {
            stage_e(sl + 1 < ns ? sl + 1 : 0);
            asm volatile("cp.async.commit_group;" ::: "memory");
          }
        }
        const float* zb = zs + buf * RB * DS;
        buf ^= 1;
#pragma unroll
        for (int d = 0; d < DS; d += 4) {
#pragma unroll
          for (int r = 0; r < RB; ++r) {
            const float4 z4 = *reinterpret_cast<const float4*>(zb + r * DS + d);   // broadcast read
            acc[r] = fmaf(z4.x, e[d], acc[r]);
            acc[r] = fmaf(z4.y, e[d + 1], acc[r]);
            acc[r] = fmaf(z4.z, e[d + 2], acc[r]);
            acc[r] = fmaf(z4.w, e[d + 3], acc[r]);
          }
        }
      }
#pragma unroll
      for (int r = 0; r < RB; ++r) {
        if (r < nb) {   // warp-uniform
          const float zzr = __shfl_sync(0xffffffffu, zz, r);
          float dist = __fmaf_rn(-2.0f, acc[r], __fadd_rn(zzr, eek));
          if (!valid || !(dist == dist)) dist = INFINITY;
          const uint32_t key = float_key(dist);
          const uint32_t kmin = __reduce_min_sync(0xffffffffu, key);
          const int src = __ffs(__ballot_sync(0xffffffffu, key == kmin)) - 1;   // lowest lane = lowest code
          const long long rr = __shfl_sync(0xffffffffu, (long long)row, r);
          if (lane == 0) atomicMin(idx64 + rr, ((unsigned long long)kmin << 32) | (unsigned)(b * 32 + src));
        }
      }
      row = row_n; zz = zz_n;
    }
  }
}

template <bool TRAIN>
__global__ void __launch_bounds__(PREP_THREADS, 1)
refine_emit_kernel(const float* __restrict__ z, const float* __restrict__ E, int D, float* __restrict__ zq,
                   unsigned long long* __restrict__ idx64, unsigned long long* __restrict__ hist, double* __restrict__ sse,
                   const int* __restrict__ row_list, const int* __restrict__ ovf_last, const int* __restrict__ counters,
                   long long pair_cap, int list_mode) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = counters[C_N], n2 = list_mode ? counters[C_NOVF] : 0;
  const long long total = counters[C_PAIRS];
  if (n + n2 == 0 || total > pair_cap || total >= (1ll << 30)) return;
  const int wtotal = gridDim.x * (PREP_THREADS / 32);
  const int g = lane >> 3, j = lane & 7;   // four rows per warp, 8 lanes each: four independent fetch chains in flight
  double lsse = 0.0;
  for (int base = (blockIdx.x * (PREP_THREADS / 32) + warp) * 4; base < n + n2; base += wtotal * 4) {
    const int i = base + g;
    const bool active = i < n + n2;
    const int64_t row = active ? listed_row(i, n, row_list, ovf_last) : 0;
    const unsigned long long best = active ? idx64[row] : 0ull;
    const int bidx = best == ~0ull ? 0 : (int)(best & 0xffffffffull);
    float* orow = zq + row * D;
    const float* erow = E + (int64_t)bidx * D;
    const float* zsrc = z + row * D;
    if (active) {
      for (int c = j * 4; c < D; c += 32) {
        const float4 e4 = ldg4(erow + c);
        float4 o4 = e4;
        if (TRAIN) {
          const float4 z4 = ldg4(zsrc + c);
          const float dx = __fsub_rn(e4.x, z4.x), dy = __fsub_rn(e4.y, z4.y);
          const float dz = __fsub_rn(e4.z, z4.z), dw = __fsub_rn(e4.w, z4.w);
          lsse += (double)dx * dx + (double)dy * dy + (double)dz * dz + (double)dw * dw;
          o4 = make_float4(__fadd_rn(z4.x, dx), __fadd_rn(z4.y, dy), __fadd_rn(z4.z, dz), __fadd_rn(z4.w, dw));
        }
        *reinterpret_cast<float4*>(orow + c) = o4;
      }
    }
    __syncwarp();   // every lane of the group has read idx64[row] before it is overwritten
    if (active && j == 0) {
      idx64[row] = (unsigned long long)bidx;
      if (TRAIN) atomicAdd(hist + bidx, 1ull);
    }
  }
  if (TRAIN) {
    lsse = warp_sum(lsse);
    if (lane == 0 && lsse != 0.0) atomicAdd(sse, lsse);
  }
}

}  // namespace

size_t vq_refine_binned_bytes(int64_t N) {
  return align_up(sizeof(int) * (size_t)(4 * MAX_BINS + 8), 256) + align_up(sizeof(int) * 2 * (size_t)N, 256);
}

// the bin tables are zeroed by the caller (count, cursor: the first 2 * MAX_BINS ints of `ws`)
size_t vq_refine_binned_zero_bytes() { return sizeof(int) * 2 * MAX_BINS; }

int launch_vq_refine_binned(const float* z, const float* E, const float* ee, int64_t N, int K, int D, int train, float* z_q,
                            int64_t* idx, unsigned long long* hist, double* sse, const int* row_list, const int* cand_list,
                            int* counters, int list_mode, void* ws, cudaStream_t s) {
  DeviceProps dp;
  int rc = device_props(&dp);
  if (rc) return rc;
  if (K > 32 * MAX_BINS) return fail(DVQ_ERR_BAD_SHAPE, "binned refine handles K <= %d", 32 * MAX_BINS);
  int* wi = static_cast<int*>(ws);
  BinTables bt;
  bt.count = wi; bt.cursor = wi + MAX_BINS; bt.start = wi + 2 * MAX_BINS; bt.item_start = wi + 3 * MAX_BINS + 1;
  int* pairs = reinterpret_cast<int*>(static_cast<char*>(ws) + align_up(sizeof(int) * (size_t)(4 * MAX_BINS + 8), 256));
  const long long cap_override = refine_pair_cap_override();
  const long long pair_cap = (cap_override > 0 && cap_override < 2 * (long long)N) ? cap_override : 2 * (long long)N;
  const int* ovf_last = row_list + (N - 1);
  unsigned long long* idx64 = reinterpret_cast<unsigned long long*>(idx);
  // slice width of the pairs kernel: a lane keeps DS floats of its code row in registers.  DS = 64 (168 registers, one CTA
  // of 12 warps per SM) re-reads a code row half as often, DS = 32 (80 registers) fits two CTAs per SM: measured at
  // N = 2M-4M, refine stage: e_dim 128 / 256 / 512 at K = 16 384 / 16 384 / 4096: 0.60 -> 0.44, 1.85 -> 1.37, 2.32 -> 1.71 ms
  // with DS = 32; e_dim 64 (one slice at DS = 64: the code row stays in registers for the whole work item): 0.15 -> 0.19 ms at
  // K = 2048.  Hence 32 from e_dim 128 on.  DVQ_REFINE_DS=32|64 overrides (A/B runs).
  static const char* ds_env = getenv("DVQ_REFINE_DS");
  const int ds_want = ds_env ? atoi(ds_env) : (D >= 128 ? 32 : 64);
  const int DS = D < 64 ? D : (ds_want == 32 ? 32 : 64);
  const int pair_grid = dp.sm_count * (DS <= 32 && D >= 64 ? 2 : 1);
  const int pair_warps = pair_grid * (PAIR_THREADS / 32);

  refine_prep_kernel<<<dp.sm_count, PREP_THREADS, 0, s>>>(z, D, K, z_q, idx64, row_list, cand_list, ovf_last, counters, bt, list_mode);
  DVQ_CUDA_CHECK(cudaGetLastError());
  refine_scatter_kernel<<<dp.sm_count, PREP_THREADS, 0, s>>>(K, cand_list, counters, bt, pairs, pair_cap, list_mode, pair_warps);
  DVQ_CUDA_CHECK(cudaGetLastError());
  const size_t smem = (size_t)(PAIR_THREADS / 32) * (2 * RB * DS + (D > DS ? 32 * (DS + 4) : 0)) * sizeof(float);
#define DVQ_LAUNCH_PAIRS(DS_)                                                                                              \
  do {                                                                                                                     \
    DVQ_CUDA_CHECK(cudaFuncSetAttribute(refine_pairs_kernel<DS_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    refine_pairs_kernel<DS_><<<pair_grid, PAIR_THREADS, smem, s>>>(z, E, ee, K, D, z_q, idx64, row_list, ovf_last, pairs,  \
                                                                   counters, bt, pair_cap, list_mode);                     \
  } while (0)
  if (DS == 64) DVQ_LAUNCH_PAIRS(64);
  else if (DS == 32) DVQ_LAUNCH_PAIRS(32);
  else if (DS == 16) DVQ_LAUNCH_PAIRS(16);
  else return fail(DVQ_ERR_BAD_SHAPE, "binned refine needs e_dim 16, 32 or a multiple of 64");
#undef DVQ_LAUNCH_PAIRS
  DVQ_CUDA_CHECK(cudaGetLastError());
  if (train)
    refine_emit_kernel<true><<<dp.sm_count, PREP_THREADS, 0, s>>>(z, E, D, z_q, idx64, hist, sse, row_list, ovf_last, counters, pair_cap, list_mode);
  else
    refine_emit_kernel<false><<<dp.sm_count, PREP_THREADS, 0, s>>>(z, E, D, z_q, idx64, hist, sse, row_list, ovf_last, counters, pair_cap, list_mode);
  DVQ_CUDA_CHECK(cudaGetLastError());
  count_launch(4);
  return DVQ_OK;
}

}  // namespace dvq
