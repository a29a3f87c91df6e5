// TMEM read-out microbenchmark (sm_100a): bytes per cycle and SM that tcgen05.ld delivers to registers, as a function of
// the number of warps reading (1..4 per TMEM lane quarter = 4..16 per SM), the load width (x32 / x64 / x128 columns per
// instruction, shape 32x32b) and the shape 16x256b.  The VQ filter reads every FP32 accumulator once (128 rows x K codes x
// 4 bytes per row tile), so this rate is a roofline of its own for short contractions (e_dim 64).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/tmem_read scripts/ubench/tmem_read.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define R32(v, o) "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]), \
                  "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15]), \
                  "=r"(v[o + 16]), "=r"(v[o + 17]), "=r"(v[o + 18]), "=r"(v[o + 19]), "=r"(v[o + 20]), "=r"(v[o + 21]), "=r"(v[o + 22]), "=r"(v[o + 23]), \
                  "=r"(v[o + 24]), "=r"(v[o + 25]), "=r"(v[o + 26]), "=r"(v[o + 27]), "=r"(v[o + 28]), "=r"(v[o + 29]), "=r"(v[o + 30]), "=r"(v[o + 31])

__device__ __forceinline__ void ld_x32(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : R32(v, 0)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld_x64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, %50, %51, "
      "%52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : R32(v, 0), R32(v, 32)
      : "r"(taddr)
      : "memory");
}
// 16x256b.x8: 16 lanes x 64 columns... per thread 32 registers (8 repeats x 4 registers)
__device__ __forceinline__ void ld_16x256_x8(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : R32(v, 0)
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// MODE 0: 32x32b.x32, one load in flight; 1: 32x32b.x32, two in flight; 2: 32x32b.x64; 3: 16x256b.x8 (both lane halves)
template <int MODE>
__global__ void k(uint32_t* out, long long* cyc, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&slot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16);
  const int wq = warp >> 2, nwq = blockDim.x >> 7;   // this warp's share of the 512 columns
  uint32_t v[64];
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    // every warp of a lane quarter reads its share of the 512 columns once per iteration
    if (MODE == 0) {
      for (int c = wq * 32; c < 512; c += 32 * nwq) { ld_x32(base + c, v); ld_wait(); acc += v[0] ^ v[31]; }
    } else if (MODE == 1) {
      for (int c = wq * 64; c < 512; c += 64 * nwq) {
        uint32_t w[64];
        ld_x32(base + c, v); ld_x32(base + c + 32, w); ld_wait(); acc += v[0] ^ v[31] ^ w[0] ^ w[31];
      }
    } else if (MODE == 2) {
      for (int c = wq * 64; c < 512; c += 64 * nwq) { ld_x64(base + c, v); ld_wait(); acc += v[0] ^ v[63]; }
    } else {
      // 16x256b: one instruction covers 16 lanes x (8 x 8) columns; two per 32-lane quarter
      for (int c = wq * 64; c < 512; c += 64 * nwq) {
        ld_16x256_x8(base + c, v); ld_wait(); acc += v[0] ^ v[31];
        ld_16x256_x8(base + ((uint32_t)16 << 16) + c, v); ld_wait(); acc += v[0] ^ v[31];
      }
    }
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(slot), "r"(512u) : "memory");
}

template <int MODE>
void run(uint32_t* out, long long* cyc, const char* name) {
  for (int wq = 1; wq <= 4; ++wq) {
    const int threads = 128 * wq, iters = 2000;
    k<MODE><<<148, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    k<MODE><<<148, threads>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < 148; ++i) avg += h[i];
    avg /= 148;
    const double bytes = (double)iters * 128 * 512 * 4;   // the whole TMEM once per iteration
    printf("%-34s warps/quarter %d  %.1f cycles per 256 KB  %.1f B/clk/SM  %s\n", name, wq, avg / iters, bytes / avg, e ? cudaGetErrorString(e) : "");
  }
}

int main() {
  uint32_t* out;
  long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4);
  cudaMalloc(&cyc, 148 * 8);
  run<0>(out, cyc, "32x32b.x32, 1 in flight");
  run<1>(out, cyc, "32x32b.x32, 2 in flight");
  run<2>(out, cyc, "32x32b.x64");
  run<3>(out, cyc, "16x256b.x8 (two per quarter)");
  return 0;
}
