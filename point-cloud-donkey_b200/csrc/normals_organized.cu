// normals_organized.cu — normals of an ORGANIZED cloud (a Kinect-style height x width grid with NaN holes).
//
// Replaces the organized branch of ImplicitShapeModel::computeNormals (implicit_shape_model.cpp:948-966):
// pcl::IntegralImageNormalEstimation with AVERAGE_3D_GRADIENT, MaxDepthChangeFactor 0.02, NormalSmoothingSize 10,
// border policy IGNORE, no depth-dependent smoothing, normals flipped towards the sensor origin.
//   1. k_org_gradients  3D gradients right - left / down - up of the interior pixels (0 on the image border), written as
//                       eight fp64 channels per pixel: dx.xyz, finite(dx), dy.xyz, finite(dy) — non-finite gradients add 0
//   2. k_org_rowscan / k_org_colscan   the two (W + 1) x (H + 1) integral images as a row scan then a column scan (PCL
//                       runs the 2D recurrence; the fp64 sums agree to rounding)
//   3. k_org_change     depth-change map -> initial distances (0 at a depth jump / hole, W + H elsewhere)
//   4. k_org_chamfer    PCL's two raster passes of the 1 / 1.4 chamfer transform, order dependent as written: one CTA
//                       walks the anti-diagonal wavefronts t = 2 row + column (every pixel of a wavefront only depends
//                       on earlier ones), flat-array indexing as PCL's pointer arithmetic incl. its row-end reads
//   5. k_org_normals    box sums of both gradients over an int(min(distance, 10))^2 rectangle, n = gy x gx normalised
#include <cmath>

#include "common.cuh"
#include "stages.h"

namespace {

__global__ void k_org_gradients(const float* __restrict__ xyz, int W, int H, double* __restrict__ ch) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)W * H) return;
  const int r = (int)(i / W), c = (int)(i % W);
  float dx[3] = {0.f, 0.f, 0.f}, dy[3] = {0.f, 0.f, 0.f};
  if (r >= 1 && r < H - 1 && c >= 1 && c < W - 1) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      dx[a] = __fsub_rn(xyz[3 * (i + 1) + a], xyz[3 * (i - 1) + a]);
      dy[a] = __fsub_rn(xyz[3 * (i + W) + a], xyz[3 * (i - W) + a]);
    }
  }
  const bool fx = isfinite(__fadd_rn(__fadd_rn(dx[0], dx[1]), dx[2]));  // pcl_isfinite(element->sum())
  const bool fy = isfinite(__fadd_rn(__fadd_rn(dy[0], dy[1]), dy[2]));
  double* o = ch + 8 * i;
  o[0] = fx ? (double)dx[0] : 0.0; o[1] = fx ? (double)dx[1] : 0.0; o[2] = fx ? (double)dx[2] : 0.0; o[3] = fx ? 1.0 : 0.0;
  o[4] = fy ? (double)dy[0] : 0.0; o[5] = fy ? (double)dy[1] : 0.0; o[6] = fy ? (double)dy[2] : 0.0; o[7] = fy ? 1.0 : 0.0;
}

// one warp per image row: inclusive prefix of the 8 channels along the row into I[(r + 1)][c + 1]; row 0 / column 0 = 0
__global__ void k_org_rowscan(const double* __restrict__ ch, int W, int H, double* __restrict__ I) {
  const int r = (int)((blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (r > H) return;
  const size_t IW = (size_t)W + 1;
  if (r == H) {  // integral row 0
    for (size_t c = lane; c < IW * 8; c += 32) I[c] = 0.0;
    return;
  }
  double* out = I + (size_t)(r + 1) * IW * 8;
  if (lane < 8) out[lane] = 0.0;
  double carry[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int c0 = 0; c0 < W; c0 += 32) {
    const int c = c0 + lane;
    double v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = c < W ? ch[8 * ((size_t)r * W + c) + k] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const double t = __shfl_up_sync(0xffffffffu, v[k], o);
        if (lane >= o) v[k] += t;
      }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      v[k] += carry[k];
      if (c < W) out[8 * (size_t)(c + 1) + k] = v[k];
      carry[k] = __shfl_sync(0xffffffffu, v[k], 31);
    }
  }
}

// one thread per (integral column, channel): running sum down the rows
__global__ void k_org_colscan(int W, int H, double* __restrict__ I) {
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const size_t IW = (size_t)W + 1;
  if (t >= (long long)IW * 8) return;
  double acc = 0.0;
  for (int r = 1; r <= H; ++r) {
    double* p = I + (size_t)r * IW * 8 + t;
    acc += *p;
    *p = acc;
  }
}

__device__ __forceinline__ bool depth_jump(float d, float o) {
  const float lim = __fmul_rn(__fmul_rn(0.02f, __fadd_rn(fabsf(d), 1.0f)), 2.0f);
  return fabsf(__fsub_rn(d, o)) > lim || !isfinite(d) || !isfinite(o);
}

__global__ void k_org_change(const float* __restrict__ xyz, int W, int H, float* __restrict__ dist) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)W * H) return;
  const int r = (int)(i / W), c = (int)(i % W);
  auto z = [&](long long j) { return xyz[3 * j + 2]; };
  bool zero = false;
  if (r < H - 1 && c < W - 1) zero |= depth_jump(z(i), z(i + 1)) || depth_jump(z(i), z(i + W));
  if (c >= 1 && r < H - 1) zero |= depth_jump(z(i - 1), z(i));          // the pixel to the left marked its right neighbour
  if (r >= 1 && c < W - 1) zero |= depth_jump(z(i - W), z(i));          // the pixel above marked its lower neighbour
  dist[i] = zero ? 0.0f : (float)(W + H);
}

// both raster passes of PCL's chamfer transform; single CTA, global memory (the image does not fit shared memory)
__global__ void __launch_bounds__(1024) k_org_chamfer(int W, int H, float* dist) {
  volatile float* d = dist;
  // forward: rows 1..H-1, columns 1..W-1; (r, c) reads (r-1, c-1), (r-1, c), (r-1, c+1) and (r, c-1)
  for (int t = 3; t <= 2 * (H - 1) + (W - 1); ++t) {
    const int r_lo = max(1, (t - (W - 1) + 1) / 2), r_hi = min(H - 1, (t - 1) / 2);
    for (int r = r_lo + (int)threadIdx.x; r <= r_hi; r += (int)blockDim.x) {
      const int c = t - 2 * r;
      const long long i = (long long)r * W + c;
      const float upLeft = d[i - W - 1] + 1.4f, up = d[i - W] + 1.0f, upRight = d[i - W + 1] + 1.4f;
      const float left = d[i - 1] + 1.0f, center = d[i];
      const float mn = fminf(fminf(upLeft, up), fminf(left, upRight));
      if (mn < center) d[i] = mn;
    }
    __syncthreads();
  }
  // backward: rows H-2..0, columns W-2..0; (r, c) reads (r+1, c-1), (r+1, c), (r+1, c+1) and (r, c+1)
  for (int t = 3; t <= 2 * (H - 1) + (W - 1); ++t) {
    const int q_lo = max(1, (t - (W - 1) + 1) / 2), q_hi = min(H - 1, (t - 1) / 2);  // q = H-1-r, p = W-1-c = t - 2q
    for (int q = q_lo + (int)threadIdx.x; q <= q_hi; q += (int)blockDim.x) {
      const int r = H - 1 - q, c = W - 1 - (t - 2 * q);
      const long long i = (long long)r * W + c;
      const float lowerLeft = d[i + W - 1] + 1.4f, lower = d[i + W] + 1.0f, lowerRight = d[i + W + 1] + 1.4f;
      const float right = d[i + 1] + 1.0f, center = d[i];
      const float mn = fminf(fminf(lowerLeft, lower), fminf(right, lowerRight));
      if (mn < center) d[i] = mn;
    }
    __syncthreads();
  }
}

__global__ void k_org_normals(const float* __restrict__ xyz, int W, int H, const double* __restrict__ I,
                              const float* __restrict__ dist, float* __restrict__ nrm) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)W * H) return;
  const int r = (int)(i / W), c = (int)(i % W);
  const float nanf_ = __int_as_float(0x7fc00000);
  float nx = nanf_, ny = nanf_, nz = nanf_;
  const int border = 10;  // int(normal_smoothing_size_)
  if (r >= border && r < H - border && c >= border && c < W - border && isfinite(xyz[3 * i + 2])) {
    const float smoothing = fminf(dist[i], 10.0f);
    if (smoothing > 2.0f) {
      const int rw = (int)smoothing;
      const int sx = c - rw / 2, sy = r - rw / 2;
      const size_t IW = (size_t)W + 1;
      const double* ul = I + 8 * ((size_t)sy * IW + sx);
      const double* ur = ul + 8 * rw;
      const double* ll = I + 8 * ((size_t)(sy + rw) * IW + sx);
      const double* lr = ll + 8 * rw;
      auto box = [&](int k) { return ((lr[k] + ul[k]) - ur[k]) - ll[k]; };
      if (box(3) != 0.0 && box(7) != 0.0) {
        const double gx0 = box(0), gx1 = box(1), gx2 = box(2), gy0 = box(4), gy1 = box(5), gy2 = box(6);
        const double v0 = gy1 * gx2 - gy2 * gx1, v1 = gy2 * gx0 - gy0 * gx2, v2 = gy0 * gx1 - gy1 * gx0;
        const double len2 = v0 * v0 + v1 * v1 + v2 * v2;
        if (len2 != 0.0) {
          const double len = sqrt(len2);
          nx = (float)(v0 / len); ny = (float)(v1 / len); nz = (float)(v2 / len);
          // pcl::flipNormalTowardsViewpoint, viewpoint = sensor origin
          const float vx = __fsub_rn(0.f, xyz[3 * i]), vy = __fsub_rn(0.f, xyz[3 * i + 1]), vz = __fsub_rn(0.f, xyz[3 * i + 2]);
          const float cs = __fadd_rn(__fadd_rn(__fmul_rn(vx, nx), __fmul_rn(vy, ny)), __fmul_rn(vz, nz));
          if (cs < 0.f) { nx = -nx; ny = -ny; nz = -nz; }
        }
      }
    }
  }
  nrm[3 * i] = nx;
  nrm[3 * i + 1] = ny;
  nrm[3 * i + 2] = nz;
}

}  // namespace

// Input: ws.in_xyz (H x W x 3).  Output: ws.in_nrm (H x W x 3), NaN where PCL yields NaN.
int stage_normals_organized(pcdb_ctx* ctx, int W, int H) {
  Workspace& w = ctx->ws;
  cudaStream_t st = ctx->stream;
  const int64_t n = (int64_t)W * H;
  PCDB_CUDA(w.in_nrm.ensure(sizeof(float) * 3 * (n + 1)));
  if (n == 0) return PCDB_OK;
  float* nrm = w.in_nrm.as<float>();
  const float* xyz = w.in_xyz.as<float>();
  PCDB_CUDA(w.org_ch.ensure(sizeof(double) * 8 * (size_t)n));
  PCDB_CUDA(w.org_int.ensure(sizeof(double) * 8 * (size_t)(W + 1) * (H + 1)));
  PCDB_CUDA(w.org_dist.ensure(sizeof(float) * (size_t)(n + 1)));
  const unsigned g = cdiv(n, 256);
  if (W < 3 || H < 3) {  // no interior pixel: every normal is NaN
    PCDB_CUDA(cudaMemsetAsync(nrm, 0xff, sizeof(float) * 3 * n, st));
    return PCDB_OK;
  }
  k_org_gradients<<<g, 256, 0, st>>>(xyz, W, H, w.org_ch.as<double>());
  PCDB_LAUNCH_CHECK();
  k_org_rowscan<<<cdiv(((int64_t)H + 1) * 32, 128), 128, 0, st>>>(w.org_ch.as<double>(), W, H, w.org_int.as<double>());
  PCDB_LAUNCH_CHECK();
  k_org_colscan<<<cdiv(((int64_t)W + 1) * 8, 128), 128, 0, st>>>(W, H, w.org_int.as<double>());
  PCDB_LAUNCH_CHECK();
  k_org_change<<<g, 256, 0, st>>>(xyz, W, H, w.org_dist.as<float>());
  PCDB_LAUNCH_CHECK();
  k_org_chamfer<<<1, 1024, 0, st>>>(W, H, w.org_dist.as<float>());
  PCDB_LAUNCH_CHECK();
  k_org_normals<<<g, 256, 0, st>>>(xyz, W, H, w.org_int.as<double>(), w.org_dist.as<float>(), nrm);
  PCDB_LAUNCH_CHECK();
  return PCDB_OK;
}
