// Diagnostic: one CTA, one accumulator tile.  Loads pre-built operand images (exact shared-memory
// byte images) and issues `ksteps` tcgen05.mma (M=128, kind::f16) with caller-supplied descriptor
// strides, then dumps the TMEM accumulator.  tests/test_tc_probe_gpu.py uses it to pin the
// no-swizzle K-major descriptor convention the VQ kernel relies on against a CPU matmul.
#include "dvq_common.cuh"
#include "tc_prims.cuh"

namespace dvq {
namespace {
__global__ void __launch_bounds__(128, 1)
umma_probe_kernel(const uint4* __restrict__ a_img, uint32_t a_bytes, const uint4* __restrict__ b_img, uint32_t b_bytes,
                  int ksteps, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_kstep, uint32_t b_lbo, uint32_t b_sbo,
                  uint32_t b_kstep, uint32_t idesc, int n_cols, float* __restrict__ out, int* __restrict__ err) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ int serr;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* a_s = smem;
  uint8_t* b_s = smem + ((a_bytes + 1023u) & ~1023u);
  for (uint32_t i = tid; i < a_bytes / 16; i += 128) reinterpret_cast<uint4*>(a_s)[i] = a_img[i];
  for (uint32_t i = tid; i < b_bytes / 16; i += 128) reinterpret_cast<uint4*>(b_s)[i] = b_img[i];
  tc::fence_proxy_async_smem();
  if (tid == 0) {
    serr = 0;
    tc::mbar_init(&bar, 1);
    tc::fence_barrier_init();
  }
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t taddr = tmem_slot;
  if (tid == 0) {
    const uint32_t a0 = tc::smem_u32(a_s), b0 = tc::smem_u32(b_s);
    for (int j = 0; j < ksteps; ++j)
      tc::umma_f16(taddr, tc::make_smem_desc(a0 + j * a_kstep, a_lbo, a_sbo), tc::make_smem_desc(b0 + j * b_kstep, b_lbo, b_sbo),
                   idesc, j > 0 ? 1u : 0u);
    tc::umma_commit(&bar);
  }
  const bool ok = tc::mbar_wait(&bar, 0, &serr, 1);
  tc::tc_fence_after();
  if (ok) {
    for (int c0 = 0; c0 < n_cols; c0 += 32) {
      uint32_t v[32];
      tc::tmem_ld32(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) out[(size_t)(warp * 32 + lane) * n_cols + c0 + j] = __uint_as_float(v[j]);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (tid == 0 && serr) *err = serr;
  if (warp == 0) tc::tmem_dealloc(taddr, 256);
}
}  // namespace

int launch_umma_probe(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes, int ksteps,
                      const uint32_t* strides /*a_lbo,a_sbo,a_kstep,b_lbo,b_sbo,b_kstep*/, uint32_t idesc, int n_cols,
                      float* out, int* err, cudaStream_t s) {
  const size_t smem = ((a_bytes + 1023u) & ~1023u) + b_bytes + 1024;
  DVQ_CUDA_CHECK(cudaFuncSetAttribute(umma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_probe_kernel<<<1, 128, smem, s>>>(static_cast<const uint4*>(a_img), a_bytes, static_cast<const uint4*>(b_img), b_bytes,
                                         ksteps, strides[0], strides[1], strides[2], strides[3], strides[4], strides[5],
                                         idesc, n_cols, out, err);
  DVQ_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return DVQ_OK;
}
}  // namespace dvq
