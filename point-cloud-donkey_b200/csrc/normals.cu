// normals.cu — normal estimation + consistent orientation for clouds that arrive without normals.
//
// Replaces ImplicitShapeModel::computeNormals for unorganized clouds (implicit_shape_model.cpp:940-1037):
//   pcl::NormalEstimationOMPWithEigVals::computeFeature   third_party/pcl_normal_3d_omp_with_eigenvalues/
//                                                          normal_3d_omp_with_eigenvalues.hpp:63-141
//   computePointNormalMod / eigen33Mod / flipNormalTowardsViewpointMod            ...with_eigenvalues.h:60-181
//   NormalOrientation::processSHOTLRF                     utils/normal_orientation.cpp:48-110
// ConsistentNormalsMethod 0: PCA normal flipped towards the origin; 1: centroid removed, flipped, inverted;
// 2 (code default): inverted z axis of the SHOT reference frame of every point (radius = NormalRadius), PCA normal
// where that frame is undefined.  Method 2 reuses the fused LRF kernel (shot.cu) with "keypoints = all points".
//
// The radius search runs on the same uniform grid as the descriptor stage (9 binary searches per 27-cell
// neighbourhood).  One warp per point accumulates the nine moment sums in fp64 (the reference: fp32 in search order —
// its covariance is only good to ~1e-3 relative at these scales, so the parity bar for PCA normals is an angle, see
// tests), then the reference's closed-form float eigen-solve (pcl::computeRoots + cross products) is evaluated as
// written.
#include "common.cuh"
#include "stages.h"

namespace {

__device__ __forceinline__ void roots2(float b, float c, float r[3]) {
  r[0] = 0.0f;
  float d = (float)((double)(b * b) - 4.0 * (double)c);
  if (d < 0.0f) d = 0.0f;
  float sd = sqrtf(d);
  r[2] = 0.5f * (b + sd);
  r[1] = 0.5f * (b - sd);
}

// pcl::computeRoots (PCL 1.10 common/impl/eigen.hpp), float
__device__ void roots3(const float m[3][3], float r[3]) {
  float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] -
             m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] -
             m[1][2] * m[1][2];
  float c2 = m[0][0] + m[1][1] + m[2][2];
  if (fabsf(c0) < 1.1920929e-07f) {
    roots2(c2, c1, r);
    return;
  }
  const float s_inv3 = (float)(1.0 / 3.0);
  const float s_sqrt3 = sqrtf(3.0f);
  float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0f) a_over_3 = 0.0f;
  float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0f) q = 0.0f;
  float rho = sqrtf(-a_over_3);
  float theta = atan2f(sqrtf(-q), half_b) * s_inv3;
  float cos_theta = cosf(theta), sin_theta = sinf(theta);
  r[0] = c2_over_3 + 2.0f * rho * cos_theta;
  r[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  r[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (r[0] >= r[1]) { float t = r[0]; r[0] = r[1]; r[1] = t; }
  if (r[1] >= r[2]) {
    float t = r[1]; r[1] = r[2]; r[2] = t;
    if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
  }
  if (r[0] <= 0) roots2(c2, c1, r);
}

// eigen33Mod (..._with_eigenvalues.h:60-95): smallest eigenvalue's vector, eigenvalues ascending
__device__ void eigen33_mod(const float mat[3][3], float evals[3], float evec[3]) {
  float scale = 0.f;
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) scale = fmaxf(scale, fabsf(mat[i][j]));
  if (scale <= 1.17549435e-38f) scale = 1.0f;
  float sm[3][3];
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) sm[i][j] = __fdiv_rn(mat[i][j], scale);
  float rt[3];
  roots3(sm, rt);
#pragma unroll
  for (int i = 0; i < 3; ++i) evals[i] = rt[i] * scale;
#pragma unroll
  for (int i = 0; i < 3; ++i) sm[i][i] -= rt[0];
  float v1[3] = {sm[0][1] * sm[1][2] - sm[0][2] * sm[1][1], sm[0][2] * sm[1][0] - sm[0][0] * sm[1][2],
                 sm[0][0] * sm[1][1] - sm[0][1] * sm[1][0]};
  float v2[3] = {sm[0][1] * sm[2][2] - sm[0][2] * sm[2][1], sm[0][2] * sm[2][0] - sm[0][0] * sm[2][2],
                 sm[0][0] * sm[2][1] - sm[0][1] * sm[2][0]};
  float v3[3] = {sm[1][1] * sm[2][2] - sm[1][2] * sm[2][1], sm[1][2] * sm[2][0] - sm[1][0] * sm[2][2],
                 sm[1][0] * sm[2][1] - sm[1][1] * sm[2][0]};
  float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const float* v = v3;
  float l = l3;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
  else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
  const float s = sqrtf(l);
#pragma unroll
  for (int a = 0; a < 3; ++a) evec[a] = __fdiv_rn(v[a], s);
}

// pcl::compute3DCentroid (float accumulators, index order): one thread per cloud
__global__ void k_centroid(const float4* __restrict__ pts, const long long* __restrict__ off, int B, float4* cen) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  const long long lo = off[b], hi = off[b + 1];
  for (long long i = lo; i < hi; ++i) {
    float4 p = pts[i];
    sx = __fadd_rn(sx, p.x);
    sy = __fadd_rn(sy, p.y);
    sz = __fadd_rn(sz, p.z);
  }
  const float n = (float)(hi - lo);
  cen[b] = hi > lo ? make_float4(__fdiv_rn(sx, n), __fdiv_rn(sy, n), __fdiv_rn(sz, n), 0.f) : make_float4(0, 0, 0, 0);
}

__global__ void k_shift(float4* pts, const int* __restrict__ cloud, long long n, const float4* __restrict__ cen) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  float4 p = pts[i];
  const float4 c = cen[cloud[i]];
  pts[i] = make_float4(__fsub_rn(p.x, c.x), __fsub_rn(p.y, c.y), __fsub_rn(p.z, c.z), p.w);
}

// computePointNormalMod for every point: one warp per point, result (nx, ny, nz, curvature) NOT yet flipped
__global__ void __launch_bounds__(256) k_normals_pca(const float4* __restrict__ pts, const int* __restrict__ cloud,
                                                     long long n, const float4* __restrict__ surfS,
                                                     const unsigned long long* __restrict__ skeys,
                                                     const long long* __restrict__ surf_off,
                                                     const CloudInfo* __restrict__ ci, float inv_cell, float r2,
                                                     float4* out) {
  const long long i = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (i >= n) return;
  const float4 p = pts[i];
  const int b = cloud[i];
  const CloudInfo c = ci[b];
  const int cx = grid_coord(p.x, c.mn[0], inv_cell), cy = grid_coord(p.y, c.mn[1], inv_cell),
            cz = grid_coord(p.z, c.mn[2], inv_cell);
  long long beg = 0;
  int len = 0;
  if (lane < 9) {
    const int y = cy + lane % 3 - 1, z = cz + lane / 3 - 1;
    if (y >= 0 && y <= 65535 && z >= 0 && z <= 65535) {
      const long long lo = surf_off[b], hi = surf_off[b + 1];
      beg = lower_bound_u64(skeys, lo, hi, grid_key((unsigned)b, max(cx - 1, 0), y, z));
      const long long end = lower_bound_u64(skeys, beg, hi, grid_key((unsigned)b, min(cx + 1, 65535), y, z) + 1ull);
      len = (int)(end - beg);
    }
  }
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0, a4 = 0, a5 = 0, a6 = 0, a7 = 0, a8 = 0;
  int cnt = 0;
  for (int r = 0; r < 9; ++r) {
    const long long rb = __shfl_sync(0xffffffffu, beg, r);
    const int rl = __shfl_sync(0xffffffffu, len, r);
    for (int e = lane; e < rl; e += 32) {
      const float4 q = surfS[rb + e];
      if (sqdist3_rn(p.x, p.y, p.z, q.x, q.y, q.z) < r2) {
        const double x = q.x, y = q.y, z = q.z;
        a0 += x * x; a1 += x * y; a2 += x * z; a3 += y * y; a4 += y * z; a5 += z * z;
        a6 += x; a7 += y; a8 += z;
        ++cnt;
      }
    }
  }
  a0 = warp_sum(a0); a1 = warp_sum(a1); a2 = warp_sum(a2); a3 = warp_sum(a3); a4 = warp_sum(a4);
  a5 = warp_sum(a5); a6 = warp_sum(a6); a7 = warp_sum(a7); a8 = warp_sum(a8);
  cnt = warp_sum(cnt);
  if (lane != 0) return;
  const float qnan = __int_as_float(0x7fc00000);
  if (cnt < 3) {
    out[i] = make_float4(qnan, qnan, qnan, qnan);
    return;
  }
  const double inv = 1.0 / (double)cnt;
  const double mx = a6 * inv, my = a7 * inv, mz = a8 * inv;
  float cov[3][3];
  cov[0][0] = (float)(a0 * inv - mx * mx);
  cov[0][1] = (float)(a1 * inv - mx * my);
  cov[0][2] = (float)(a2 * inv - mx * mz);
  cov[1][1] = (float)(a3 * inv - my * my);
  cov[1][2] = (float)(a4 * inv - my * mz);
  cov[2][2] = (float)(a5 * inv - mz * mz);
  cov[1][0] = cov[0][1];
  cov[2][0] = cov[0][2];
  cov[2][1] = cov[1][2];
  float ev[3], nv[3];
  eigen33_mod(cov, ev, nv);
  const float eig_sum = cov[0][0] + cov[1][1] + cov[2][2];
  const float curv = eig_sum != 0.f ? fabsf(__fdiv_rn(ev[0], eig_sum)) : 0.f;
  out[i] = make_float4(nv[0], nv[1], nv[2], curv);
}

// points whose SHOT frame is undefined, per cloud (normal_orientation.cpp:63-82)
__global__ void k_invalid_count(const float* __restrict__ lrf, const int* __restrict__ cloud, long long n, int* inv) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* f = lrf + 9 * i;
  if (!(isfinite(f[0]) && isfinite(f[3]) && isfinite(f[6]))) atomicAdd(&inv[cloud[i]], 1);
}

// orientation + scatter back to the caller's point order (NaN for non-finite input points)
__global__ void k_normals_finish(long long P, const int* __restrict__ flag_pt, const int* __restrict__ pos_pt,
                                 const float4* __restrict__ pts, const int* __restrict__ cloud,
                                 const long long* __restrict__ pts_off, const float4* __restrict__ pca,
                                 const float* __restrict__ lrf, const int* __restrict__ n_invalid, int method,
                                 float* nrm_out, float* curv_out) {
  long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (j >= P) return;
  const float qnan = __int_as_float(0x7fc00000);
  float nx = qnan, ny = qnan, nz = qnan, cv = qnan;
  if (flag_pt[j]) {
    const int i = pos_pt[j];
    const float4 a = pca[i];
    const float4 p = pts[i];
    nx = a.x; ny = a.y; nz = a.z; cv = a.w;
    // flipNormalTowardsViewpointMod with the viewpoint at the origin (a NaN normal compares false and stays NaN)
    const float vx = __fsub_rn(0.f, p.x), vy = __fsub_rn(0.f, p.y), vz = __fsub_rn(0.f, p.z);
    const float cos_theta = __fadd_rn(__fadd_rn(__fmul_rn(vx, nx), __fmul_rn(vy, ny)), __fmul_rn(vz, nz));
    if (cos_theta < 0) { nx = -nx; ny = -ny; nz = -nz; }
    if (method == 1) { nx = -nx; ny = -ny; nz = -nz; }
    if (method == 2) {
      const float* f = lrf + 9 * (size_t)i;
      if (isfinite(f[0]) && isfinite(f[3]) && isfinite(f[6])) {
        nx = -f[6]; ny = -f[7]; nz = -f[8];
      }
      // normal_orientation.cpp:85-107 as written: the repair loop recomputes points 0..n_invalid-1 of the cloud
      // (not the invalid ones), without viewpoint flip; pcl::Normal(x,y,z) zeroes their curvature
      const int b = cloud[i];
      if ((long long)i - pts_off[b] < (long long)n_invalid[b]) { nx = a.x; ny = a.y; nz = a.z; cv = 0.f; }
    }
  }
  nrm_out[3 * j] = nx;
  nrm_out[3 * j + 1] = ny;
  nrm_out[3 * j + 2] = nz;
  if (curv_out) curv_out[j] = cv;
}

}  // namespace

// Inputs: ws.in_xyz (P x 3), ws.cloud_off.  Output: ws.in_nrm (P x 3) and, if curv_out_d, the curvature (P).
// Leaves the compaction / grid workspaces in an undefined state (the feature pipeline rebuilds them).  Syncs.
int stage_normals(pcdb_ctx* ctx, int B, int64_t P, float* curv_out_d) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const pcdb_params& p = ctx->prm;
  const int method = p.consistent_normals_method;
  if (method < 0 || method > 2)
    return ctx->fail(PCDB_E_INVALID, "ConsistentNormalsMethod %d is not supported (0, 1 or 2)", method);
  if (!(p.normal_radius > 0)) return ctx->fail(PCDB_E_INVALID, "NormalRadius must be positive");
  PCDB_CUDA(w.in_nrm.ensure(sizeof(float) * 3 * (P + 1)));
  if (P == 0) return PCDB_OK;
  PCDB_TRY(stage_compact(ctx, B, P, false, false));
  int n_pts_i = 0;
  PCDB_CUDA(cudaMemcpyAsync(&n_pts_i, w.pos_pt.as<int>() + P, sizeof(int), cudaMemcpyDeviceToHost, st));
  PCDB_CUDA(cudaStreamSynchronize(st));
  const int64_t n = n_pts_i;
  const double radius = (double)p.normal_radius;  // float member -> setRadiusSearch(double)
  PCDB_CUDA(w.nrm_pca.ensure(sizeof(float4) * (n + 1)));
  PCDB_CUDA(w.nrm_cen.ensure(sizeof(float4) * (B + 1)));
  PCDB_CUDA(w.nrm_inv.ensure(sizeof(int) * (B + 1)));
  PCDB_CUDA(w.kp4.ensure(sizeof(float4) * (n + 1)));
  PCDB_CUDA(w.kp_cloud.ensure(sizeof(int) * (n + 1)));
  if (n > 0) {
    if (method == 1) {
      k_centroid<<<cdiv(B, 64), 64, 0, st>>>(w.surf4.as<float4>(), w.surf_off.as<long long>(), B,
                                             w.nrm_cen.as<float4>());
      PCDB_LAUNCH_CHECK();
      k_shift<<<cdiv(n, 256), 256, 0, st>>>(w.surf4.as<float4>(), w.surf_cloud.as<int>(), n, w.nrm_cen.as<float4>());
      PCDB_LAUNCH_CHECK();
    }
    // bounding boxes over the (possibly shifted) points only; the voxel-grid part of the set-up is unused here
    PCDB_TRY(stage_cloud_setup(ctx, B, 0, w.surf4.as<float4>(), w.surf_cloud.as<int>(), n, p.leaf_size, radius));
    int err[4] = {0, 0, 0, 0};
    PCDB_CUDA(cudaMemcpyAsync(err, w.err_flag.p, sizeof(err), cudaMemcpyDeviceToHost, st));
    PCDB_CUDA(cudaStreamSynchronize(st));
    if (err[0] & 2) return ctx->fail(PCDB_E_INVALID, "NormalRadius too small for the cloud extent");
    PCDB_CUDA(cudaMemcpyAsync(w.kp4.p, w.surf4.p, sizeof(float4) * n, cudaMemcpyDeviceToDevice, st));
    PCDB_CUDA(cudaMemcpyAsync(w.kp_cloud.p, w.surf_cloud.p, sizeof(int) * n, cudaMemcpyDeviceToDevice, st));
    PCDB_TRY(stage_grid(ctx, B, n, n, false));
    if (method == 2) {
      PCDB_CUDA(w.lrf.ensure(sizeof(float) * 9 * n));
      PCDB_TRY(stage_shot(ctx, n, n, false, radius, radius, true, false, nullptr, w.lrf.as<float>(), nullptr));
      PCDB_CUDA(cudaMemsetAsync(w.nrm_inv.p, 0, sizeof(int) * (B + 1), st));
      k_invalid_count<<<cdiv(n, 256), 256, 0, st>>>(w.lrf.as<float>(), w.surf_cloud.as<int>(), n, w.nrm_inv.as<int>());
      PCDB_LAUNCH_CHECK();
    }
    k_normals_pca<<<cdiv(n * 32, 256), 256, 0, st>>>(w.surf4.as<float4>(), w.surf_cloud.as<int>(), n,
                                                     w.surfS4.as<float4>(), w.gkeys2.as<unsigned long long>(),
                                                     w.surf_off.as<long long>(), w.cinfo.as<CloudInfo>(),
                                                     ctx->grid_inv_cell, (float)(radius * radius),
                                                     w.nrm_pca.as<float4>());
    PCDB_LAUNCH_CHECK();
  }
  k_normals_finish<<<cdiv(P, 256), 256, 0, st>>>(P, w.flag_pt.as<int>(), w.pos_pt.as<int>(), w.surf4.as<float4>(),
                                                 w.surf_cloud.as<int>(), w.surf_off.as<long long>(),
                                                 w.nrm_pca.as<float4>(), w.lrf.as<float>(), w.nrm_inv.as<int>(), method,
                                                 w.in_nrm.as<float>(), curv_out_d);
  PCDB_LAUNCH_CHECK();
  return PCDB_OK;
}
