// MANO hand layer (linear blend skinning): 10 shape + 3 global-orientation + PCA pose parameters -> 778 vertices.
// Replaces the third-party call at network/gen_net.py:116-118
//     self.rh_mano(betas=recon[:, :10], global_orient=0, hand_pose=recon[:, 10:55], transl=0).vertices
// (package `mano`, created at gen_diverse_grasp_obman.py:355-360 with use_pca=True, num_pca_comps=45,
// flat_hand_mean=True; the arithmetic is smplx/lbs.py's, restated in oracle/mano_oracle.py) so that the decoder's 55
// parameters reach the 778-point PointNet without leaving the device or the stream.
//
// One CTA per hand, everything but the model tables in shared memory:
//   1  full pose = [global_orient | coeffs @ components] + pose_mean; Rodrigues per joint (angle = |r + 1e-8|)
//   2  v_shaped = v_template + shapedirs . betas   (thread = vertex coordinate, coalesced table rows)
//   3  J = J_regressor @ v_shaped                  (warp per joint, shuffle reduction)
//   4  kinematic chain: the five fingers are independent three-joint chains off the root (a lane per finger)
//   5  v_posed = v_shaped + posedirs . vec(R_1..15 - I)   (135 coalesced table rows per coordinate: the bulk of the work,
//      1.26 MB of L2-resident table per hand)
//   6  skinning: thread = vertex, T = sum_j w_vj A_j, out = T [v_posed; 1] + transl
// FP32 with fmaf throughout; agreement with the float64 oracle ~1e-7 of the hand size.
#include "dvq_common.cuh"

namespace dvq {
namespace {

constexpr int NV = 778, NJ = 16, NB = 10, NP = 45, NPF = 9 * (NJ - 1), NC = NV * 3;
constexpr int MT = 256;

__global__ void __launch_bounds__(MT) mano_lbs_kernel(const DvqManoModel m, const float* __restrict__ betas, const float* __restrict__ global_orient,
                                                      const float* __restrict__ hand_pose, const float* __restrict__ transl, int B,
                                                      float* __restrict__ vertices, float* __restrict__ joints) {
  __shared__ float pose[3 * NJ], R[NJ][9], pf[NPF], beta_s[NB], J[NJ][3], G[NJ][12], A[NJ][12];
  __shared__ float vs[NC];          // v_shaped, then v_posed
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (b >= B) return;
  // ---- 1: pose ----
  if (tid < 3) pose[tid] = (global_orient ? global_orient[(size_t)b * 3 + tid] : 0.f) + m.pose_mean[tid];
  if (tid >= 32 && tid < 32 + NP) {
    const int k = tid - 32;
    float v;
    if (m.ncomps > 0) {   // PCA coefficients -> axis-angle: sum over the first ncomps components, in order
      v = 0.f;
      for (int i = 0; i < m.ncomps; ++i) v = fmaf(hand_pose[(size_t)b * m.ncomps + i], m.hands_components[i * NP + k], v);
    } else {
      v = hand_pose[(size_t)b * NP + k];
    }
    pose[3 + k] = v + m.pose_mean[3 + k];
  }
  if (tid >= 96 && tid < 96 + NB) beta_s[tid - 96] = betas[(size_t)b * NB + tid - 96];
  __syncthreads();
  if (tid < NJ) {
    const float rx = pose[3 * tid], ry = pose[3 * tid + 1], rz = pose[3 * tid + 2];
    const float ex = rx + 1e-8f, ey = ry + 1e-8f, ez = rz + 1e-8f;
    const float angle = sqrtf(fmaf(ez, ez, fmaf(ey, ey, ex * ex)));
    const float dx = rx / angle, dy = ry / angle, dz = rz / angle;
    float s, c;
    sincosf(angle, &s, &c);
    const float t = 1.f - c;
    // R = I + s K + t K^2,  K = [[0,-dz,dy],[dz,0,-dx],[-dy,dx,0]],  K^2 = d d^T - |d|^2 I
    const float n2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    float* r = R[tid];
    r[0] = 1.f + t * (dx * dx - n2); r[1] = fmaf(t, dx * dy, -s * dz);  r[2] = fmaf(t, dx * dz, s * dy);
    r[3] = fmaf(t, dx * dy, s * dz);  r[4] = 1.f + t * (dy * dy - n2); r[5] = fmaf(t, dy * dz, -s * dx);
    r[6] = fmaf(t, dx * dz, -s * dy); r[7] = fmaf(t, dy * dz, s * dx);  r[8] = 1.f + t * (dz * dz - n2);
  }
  __syncthreads();
  if (tid < NPF) { const int j = tid / 9 + 1, e = tid % 9; pf[tid] = R[j][e] - ((e == 0 || e == 4 || e == 8) ? 1.f : 0.f); }
  // ---- 2: shape blend ----
  for (int o = tid; o < NC; o += MT) {
    float v = __ldg(m.v_template + o);
#pragma unroll
    for (int l = 0; l < NB; ++l) v = fmaf(beta_s[l], __ldg(m.shapedirs + l * NC + o), v);
    vs[o] = v;
  }
  __syncthreads();
  // ---- 3: joints ----
  for (int j = warp; j < NJ; j += MT / 32) {
    float ax = 0.f, ay = 0.f, az = 0.f;
    for (int v = lane; v < NV; v += 32) {
      const float w = __ldg(m.j_regressor + j * NV + v);
      ax = fmaf(w, vs[3 * v], ax); ay = fmaf(w, vs[3 * v + 1], ay); az = fmaf(w, vs[3 * v + 2], az);
    }
    ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
    if (lane == 0) { J[j][0] = ax; J[j][1] = ay; J[j][2] = az; }
  }
  __syncthreads();
  // ---- 4: kinematic chain (G = [R | t], 3 x 4 row-major) ----
  if (tid == 0) {
#pragma unroll
    for (int e = 0; e < 9; ++e) G[0][(e / 3) * 4 + e % 3] = R[0][e];
    G[0][3] = J[0][0]; G[0][7] = J[0][1]; G[0][11] = J[0][2];
  }
  __syncthreads();
  if (tid < 5) {   // finger tid: joints 3 tid + 1 .. 3 tid + 3 (MANO's tree: parents 0,1,2 / 0,4,5 / ...; checked on the host)
    for (int j = 3 * tid + 1; j <= 3 * tid + 3; ++j) {
      const int pj = m.parents[j];
      const float tx = J[j][0] - J[pj][0], ty = J[j][1] - J[pj][1], tz = J[j][2] - J[pj][2];
      const float* gp = G[pj];
      const float* r = R[j];
#pragma unroll
      for (int i = 0; i < 3; ++i) {
        const float g0 = gp[4 * i], g1 = gp[4 * i + 1], g2 = gp[4 * i + 2];
        G[j][4 * i + 0] = fmaf(g2, r[6], fmaf(g1, r[3], g0 * r[0]));
        G[j][4 * i + 1] = fmaf(g2, r[7], fmaf(g1, r[4], g0 * r[1]));
        G[j][4 * i + 2] = fmaf(g2, r[8], fmaf(g1, r[5], g0 * r[2]));
        G[j][4 * i + 3] = fmaf(g2, tz, fmaf(g1, ty, g0 * tx)) + gp[4 * i + 3];
      }
    }
  }
  __syncthreads();
  if (tid < NJ) {
    const float* g = G[tid];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      A[tid][4 * i] = g[4 * i]; A[tid][4 * i + 1] = g[4 * i + 1]; A[tid][4 * i + 2] = g[4 * i + 2];
      A[tid][4 * i + 3] = g[4 * i + 3] - fmaf(g[4 * i + 2], J[tid][2], fmaf(g[4 * i + 1], J[tid][1], g[4 * i] * J[tid][0]));
    }
    if (joints) {
      const float tx = transl ? transl[(size_t)b * 3] : 0.f, ty = transl ? transl[(size_t)b * 3 + 1] : 0.f, tz = transl ? transl[(size_t)b * 3 + 2] : 0.f;
      float* jo = joints + ((size_t)b * NJ + tid) * 3;
      jo[0] = g[3] + tx; jo[1] = g[7] + ty; jo[2] = g[11] + tz;
    }
  }
  // ---- 5: pose blend (in place: vs becomes v_posed) ----
  for (int o = tid; o < NC; o += MT) {
    float v = 0.f;
#pragma unroll 9
    for (int k = 0; k < NPF; ++k) v = fmaf(pf[k], __ldg(m.posedirs + k * NC + o), v);
    vs[o] += v;
  }
  __syncthreads();
  // ---- 6: skinning ----
  const float tx = transl ? transl[(size_t)b * 3] : 0.f, ty = transl ? transl[(size_t)b * 3 + 1] : 0.f, tz = transl ? transl[(size_t)b * 3 + 2] : 0.f;
  for (int v = tid; v < NV; v += MT) {
    float T[12];
#pragma unroll
    for (int e = 0; e < 12; ++e) T[e] = 0.f;
    const float4* wv = reinterpret_cast<const float4*>(m.weights + (size_t)v * NJ);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 w4 = __ldg(wv + q);
      const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float* a = A[4 * q + jj];
#pragma unroll
        for (int e = 0; e < 12; ++e) T[e] = fmaf(w[jj], a[e], T[e]);
      }
    }
    const float x = vs[3 * v], y = vs[3 * v + 1], z = vs[3 * v + 2];
    float* o = vertices + ((size_t)b * NV + v) * 3;
    o[0] = fmaf(T[2], z, fmaf(T[1], y, T[0] * x)) + T[3] + tx;
    o[1] = fmaf(T[6], z, fmaf(T[5], y, T[4] * x)) + T[7] + ty;
    o[2] = fmaf(T[10], z, fmaf(T[9], y, T[8] * x)) + T[11] + tz;
  }
}

}  // namespace

int launch_mano(const DvqManoModel* m, const float* betas, const float* global_orient, const float* hand_pose, const float* transl, int B,
                float* vertices, float* joints, cudaStream_t s) {
  if (B == 0) return DVQ_OK;
  mano_lbs_kernel<<<B, MT, 0, s>>>(*m, betas, global_orient, hand_pose, transl, B, vertices, joints);
  DVQ_CUDA_CHECK(cudaGetLastError());
  count_launch();
  return DVQ_OK;
}

}  // namespace dvq
