// Pipe-rate microbenchmark (sm_100a): cycles per warp-instruction per SM sub-partition for the instruction
// mixes of the VQ filter epilogue (FMNMX3 / FFMA.SAT imm / FFMA2 / scalar FFMA ...), with 1..4 warps per
// sub-partition.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench/pipes scripts/ubench/pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

template <int MIX>
__global__ void k(float* out, long long* cyc, int iters, float seed) {
  float x[8], y[8];
  uint64_t p[8];
  unsigned u[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { x[i] = seed + i + threadIdx.x; y[i] = seed * 0.5f + i; p[i] = pk(x[i], y[i]); u[i] = threadIdx.x * 7 + i; }
  const float c = seed * 3.f;
  const uint64_t two = pk(2.f, 2.f);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MIX == 0) {        // FFMA.SAT imm
#define X(i) asm volatile("fma.rn.sat.f32 %0, %0, -1048576.0, %1;" : "+f"(x[i]) : "f"(c));
      REP8(X) REP8(X)
#undef X
    } else if (MIX == 1) { // FFMA reg
#define X(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(c), "f"(y[i]));
      REP8(X) REP8(X)
#undef X
    } else if (MIX == 2) { // FFMA2
#define X(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(two), "l"(p[(i + 1) & 7]));
      REP8(X) REP8(X)
#undef X
    } else if (MIX == 3) { // FMNMX3
#define X(i) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(x[i]) : "f"(y[i]), "f"(c));
      REP8(X) REP8(X)
#undef X
    } else if (MIX == 4) { // FMNMX
#define X(i) asm volatile("min.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(y[i]));
      REP8(X) REP8(X)
#undef X
    } else if (MIX == 5) { // FMNMX3 + FFMA.SAT imm interleaved 1:1
#define X(i) asm volatile("min.f32 %0, %0, %2, %3;\n\tfma.rn.sat.f32 %1, %1, -1048576.0, %3;" : "+f"(x[i]), "+f"(y[i]) : "f"(seed), "f"(c));
      REP8(X)
#undef X
    } else if (MIX == 6) { // FMNMX3 + FFMA2 interleaved 1:1
#define X(i) asm volatile("min.f32 %0, %0, %2, %3;\n\tfma.rn.f32x2 %1, %1, %4, %1;" : "+f"(x[i]), "+l"(p[i]) : "f"(seed), "f"(c), "l"(two));
      REP8(X)
#undef X
    } else if (MIX == 7) { // FFMA imm (non-sat)
#define X(i) asm volatile("fma.rn.f32 %0, %0, 2.0, %1;" : "+f"(x[i]) : "f"(y[i]));
      REP8(X) REP8(X)
#undef X
    } else if (MIX == 8) { // IADD3-ish
#define X(i) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      REP8(X) REP8(X)
#undef X
    } else if (MIX == 9) { // set.lt (FSET)
#define X(i) asm volatile("set.lt.f32.f32 %0, %0, %1;" : "+f"(x[i]) : "f"(y[i]));
      REP8(X) REP8(X)
#undef X
    } else if (MIX == 10) { // filter ratio: 4 FMNMX3 : 8 FFMA.SAT : 4 FFMA2 (+ FFMA2 extra ~ 19/16)
#define X(i) asm volatile("min.f32 %0, %0, %3, %4;\n\tfma.rn.sat.f32 %1, %1, -1048576.0, %4;\n\tfma.rn.sat.f32 %3, %3, -1048576.0, %4;\n\tfma.rn.f32x2 %2, %2, %5, %2;" \
                          : "+f"(x[i]), "+f"(y[i]), "+l"(p[i]), "+f"(seed) : "f"(c), "l"(two));
      REP8(X)
#undef X
    } else if (MIX == 11) { // shf funnel + sub (integer indicator fold)
#define X(i) asm volatile("sub.u32 %1, %2, %1;\n\tshf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(u[i]), "+r"(u[(i + 3) & 7]) : "r"(u[(i + 5) & 7]));
      REP8(X)
#undef X
    } else if (MIX == 12) { // HFMA2-ish packed half min (HMNMX2)
#define X(i) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
      REP8(X) REP8(X)
#undef X
    } else if (MIX == 13) { // cvt.rn.f16x2.f32 (F2FP)
#define X(i) asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u[i]) : "f"(x[i]), "f"(y[i]));
      REP8(X) REP8(X)
#undef X
    } else if (MIX == 14) { // 2 FFMA.SAT imm : 1 FFMA2
#define X(i) asm volatile("fma.rn.sat.f32 %0, %0, -1048576.0, %3;\n\tfma.rn.sat.f32 %1, %1, -1048576.0, %3;\n\tfma.rn.f32x2 %2, %2, %4, %2;" \
                          : "+f"(x[i]), "+f"(y[i]), "+l"(p[i]) : "f"(c), "l"(two));
      REP8(X)
#undef X
    } else if (MIX == 15) { // FMNMX3 + scalar FFMA imm interleaved 1:2
#define X(i) asm volatile("min.f32 %0, %0, %2, %3;\n\tfma.rn.f32 %1, %1, 2.0, %3;\n\tfma.rn.sat.f32 %2, %2, -1048576.0, %3;" : "+f"(x[i]), "+f"(y[i]), "+f"(seed) : "f"(c));
      REP8(X)
#undef X
    } else if (MIX == 16) { // vimnmx3 u32 (integer 3-input min)
#define X(i) asm volatile("min.u32 %0, %0, %1;\n\tmin.u32 %0, %0, %2;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(u[(i + 2) & 7]));
      REP8(X)
#undef X
    }
  }
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(p[i])); s += x[i] + y[i] + a + b + (float)u[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + seed;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

static const char* names[] = {"FFMA.SAT imm", "FFMA reg", "FFMA2", "FMNMX3", "FMNMX", "FMNMX3+FFMA.SAT 1:1", "FMNMX3+FFMA2 1:1", "FFMA imm",
                              "IADD", "FSET", "filter mix 1:2:1", "ISUB+SHF funnel", "HMNMX2", "F2FP", "SAT,SAT,FFMA2", "FMNMX3+FFMAimm+SAT", "VIMNMX x2"};
static const int per_iter[] = {16, 16, 16, 16, 16, 16, 16, 16, 16, 16, 32, 16, 16, 16, 24, 24, 16};

template <int MIX>
void run(float* out, long long* cyc) {
  for (int wps = 1; wps <= 4; ++wps) {
    const int threads = 128 * wps, iters = 2000;
    k<MIX><<<148, threads>>>(out, cyc, iters, 1.25f);
    cudaDeviceSynchronize();
    k<MIX><<<148, threads>>>(out, cyc, iters, 1.25f);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
    const double winstr = (double)iters * per_iter[MIX] * wps;   // warp-instructions per sub-partition
    printf("%-24s warps/SMSP %d  cycles/warp-instr/SMSP %.3f  (IPC %.3f) %s\n", names[MIX], wps, avg / winstr, winstr / avg, e ? cudaGetErrorString(e) : "");
  }
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  run<0>(out, cyc); run<1>(out, cyc); run<2>(out, cyc); run<3>(out, cyc); run<4>(out, cyc); run<5>(out, cyc); run<6>(out, cyc); run<7>(out, cyc);
  run<8>(out, cyc); run<9>(out, cyc); run<10>(out, cyc); run<11>(out, cyc); run<12>(out, cyc); run<13>(out, cyc); run<14>(out, cyc); run<15>(out, cyc); run<16>(out, cyc);
  return 0;
}
