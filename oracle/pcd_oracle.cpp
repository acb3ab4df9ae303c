/*
 * pcd_oracle.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (C++17 + OpenMP, no dependencies) of the classification hot path of
 * vseib/point-cloud-donkey.  It is the parity checker for the CUDA library in
 * point-cloud-donkey_b200/csrc and the `cpu_baseline` / `--impl reference` arm of bench.py.
 * Nothing under point-cloud-donkey_b200/ may include, link or call it.
 *
 * PARITY STATUS: the reference cannot be built here (needs PCL/FLANN/Eigen/Boost) and ships no
 * tests or golden vectors, so most of this file is "parity unpinned": it follows the in-repo
 * orchestration line by line and restates the PCL 1.10 / FLANN 1.9 arithmetic.
 * Two pieces ARE pinned:
 *   - RGB->CIELab + colour distance: against the reference's own
 *     third_party/pcl_color_conversion/color_conversion.cpp compiled into oracle/_ref
 *     (tests/test_oracle_pins.py, tests/golden/lab_golden.npz);
 *   - the L2 / chi^2 functor values and exact-kNN order: against a real FLANN build
 *     (cv2.flann linear index; tests/golden/flann_golden.npz).
 *
 * Reference paths are relative to /root/reference/src/implicit_shape_model/.
 * Build: see oracle/Makefile (-ffp-contract=off: the reference is an x86-64 SSE build without FMA).
 */
#include "../include/pcdb200.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <random>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

const float kNaNf = std::numeric_limits<float>::quiet_NaN();

/* ------------------------------------------------------------------------------------------- */
/* A.2  radius search: pcl::search::KdTree -> FLANN KDTreeSingleIndex<L2_Simple<float>>, exact  */
/* ------------------------------------------------------------------------------------------- */
struct Nbr {
  float d2;
  int idx;
};
inline bool nbr_less(const Nbr& a, const Nbr& b) { return a.d2 < b.d2 || (a.d2 == b.d2 && a.idx < b.idx); }

/* flann L2_Simple: result += diff*diff, x then y then z, float, no FMA (compiled -ffp-contract=off). */
inline float sqdist3(const float* a, const float* b) {
  float d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
  float r = d0 * d0;
  r += d1 * d1;
  r += d2 * d2;
  return r;
}

/* pcl::KdTreeFLANN::radiusSearch squares the double radius and narrows: float(radius*radius). */
inline float radius_sq(double radius) { return static_cast<float>(radius * radius); }

/* Uniform grid used only to make the brute-force membership test cheap on big clouds; the set it
 * returns is the same as testing every point (cells are wider than the radius). */
struct CloudGrid {
  const float* xyz = nullptr;
  int n = 0;
  bool brute = true;
  float mn[3] = {0, 0, 0};
  float inv = 0;
  int dim[3] = {1, 1, 1};
  std::vector<int> start, order;

  void build(const float* pts, int count, double radius) {
    xyz = pts;
    n = count;
    brute = true;
    if (n < 512 || !(radius > 0)) return;
    float mx[3];
    for (int a = 0; a < 3; ++a) mn[a] = mx[a] = pts[a];
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a) {
        mn[a] = std::min(mn[a], pts[3 * i + a]);
        mx[a] = std::max(mx[a], pts[3 * i + a]);
      }
    float cell = static_cast<float>(radius * 1.001);
    inv = 1.0f / cell;
    double total = 1;
    for (int a = 0; a < 3; ++a) {
      double d = std::floor((double(mx[a]) - double(mn[a])) * inv) + 1;
      if (!(d >= 1) || d > 4096) return;
      dim[a] = int(d);
      total *= d;
    }
    if (total > 8e6) return;
    int ncell = dim[0] * dim[1] * dim[2];
    start.assign(ncell + 1, 0);
    std::vector<int> cid(n);
    for (int i = 0; i < n; ++i) {
      cid[i] = cell_of(pts + 3 * i);
      start[cid[i] + 1]++;
    }
    for (int c = 0; c < ncell; ++c) start[c + 1] += start[c];
    order.resize(n);
    std::vector<int> fill(start.begin(), start.end() - 1);
    for (int i = 0; i < n; ++i) order[fill[cid[i]]++] = i;
    brute = false;
  }
  int coord(float v, int a) const {
    int c = int(std::floor((v - mn[a]) * inv));
    return std::max(0, std::min(dim[a] - 1, c));
  }
  int cell_of(const float* p) const { return (coord(p[2], 2) * dim[1] + coord(p[1], 1)) * dim[0] + coord(p[0], 0); }

  void query(const float* q, float r2, std::vector<Nbr>& out) const {
    out.clear();
    if (brute) {
      for (int i = 0; i < n; ++i) {
        float d2 = sqdist3(q, xyz + 3 * i);
        if (d2 < r2) out.push_back({d2, i});
      }
    } else {
      int c[3];
      for (int a = 0; a < 3; ++a) c[a] = int(std::floor((q[a] - mn[a]) * inv));
      for (int z = std::max(0, c[2] - 1); z <= std::min(dim[2] - 1, c[2] + 1); ++z)
        for (int y = std::max(0, c[1] - 1); y <= std::min(dim[1] - 1, c[1] + 1); ++y)
          for (int x = std::max(0, c[0] - 1); x <= std::min(dim[0] - 1, c[0] + 1); ++x) {
            int cell = (z * dim[1] + y) * dim[0] + x;
            for (int s = start[cell]; s < start[cell + 1]; ++s) {
              int i = order[s];
              float d2 = sqdist3(q, xyz + 3 * i);
              if (d2 < r2) out.push_back({d2, i});
            }
          }
    }
    std::sort(out.begin(), out.end(), nbr_less); /* sorted_results_ = true; DistanceIndex::operator< */
  }
};

/* ------------------------------------------------------------------------------------------- */
/* A.1  pcl::VoxelGrid<PointXYZRGB> (keypoints/keypoints_voxel_grid.cpp:30-46)                  */
/* ------------------------------------------------------------------------------------------- */
struct Keypoints {
  std::vector<float> xyz;
  std::vector<uint32_t> rgb;
};

bool finite3(const float* p) { return std::isfinite(p[0]) && std::isfinite(p[1]) && std::isfinite(p[2]); }

/* returns false when PCL would warn "leaf size too small" (output = input copy) */
bool voxel_keypoints(const float* xyz, const uint32_t* rgb, int n, float leaf, Keypoints& out) {
  out.xyz.clear();
  out.rgb.clear();
  if (n <= 0) return true;
  float inv = 1.0f / leaf;
  float mn[3], mx[3];
  bool any = false;
  for (int i = 0; i < n; ++i) {
    const float* p = xyz + 3 * i;
    if (!finite3(p)) continue;
    if (!any) {
      for (int a = 0; a < 3; ++a) mn[a] = mx[a] = p[a];
      any = true;
    } else
      for (int a = 0; a < 3; ++a) {
        mn[a] = std::min(mn[a], p[a]);
        mx[a] = std::max(mx[a], p[a]);
      }
  }
  if (!any) return true;
  int64_t dx = static_cast<int64_t>((mx[0] - mn[0]) * inv) + 1;
  int64_t dy = static_cast<int64_t>((mx[1] - mn[1]) * inv) + 1;
  int64_t dz = static_cast<int64_t>((mx[2] - mn[2]) * inv) + 1;
  if ((dx * dy * dz) > static_cast<int64_t>(std::numeric_limits<int32_t>::max())) {
    for (int i = 0; i < n; ++i) {
      out.xyz.insert(out.xyz.end(), xyz + 3 * i, xyz + 3 * i + 3);
      out.rgb.push_back(rgb ? rgb[i] : 0u);
    }
    return false;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int a = 0; a < 3; ++a) {
    min_b[a] = static_cast<int>(std::floor(mn[a] * inv));
    max_b[a] = static_cast<int>(std::floor(mx[a] * inv));
    div_b[a] = max_b[a] - min_b[a] + 1;
  }
  int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
  std::vector<std::pair<int, int>> iv; /* (voxel idx, point idx) */
  iv.reserve(n);
  for (int i = 0; i < n; ++i) {
    const float* p = xyz + 3 * i;
    if (!finite3(p)) continue;
    int ijk0 = static_cast<int>(std::floor(p[0] * inv) - static_cast<float>(min_b[0]));
    int ijk1 = static_cast<int>(std::floor(p[1] * inv) - static_cast<float>(min_b[1]));
    int ijk2 = static_cast<int>(std::floor(p[2] * inv) - static_cast<float>(min_b[2]));
    iv.push_back({ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2], i});
  }
  /* PCL: std::sort on idx (unstable); the oracle DEFINES the order as stable by point index. */
  std::stable_sort(iv.begin(), iv.end(),
                   [](const std::pair<int, int>& a, const std::pair<int, int>& b) { return a.first < b.first; });
  size_t s = 0;
  while (s < iv.size()) {
    size_t e = s;
    float sx = 0, sy = 0, sz = 0, sr = 0, sg = 0, sb = 0;
    while (e < iv.size() && iv[e].first == iv[s].first) {
      int i = iv[e].second;
      sx += xyz[3 * i];
      sy += xyz[3 * i + 1];
      sz += xyz[3 * i + 2];
      uint32_t c = rgb ? rgb[i] : 0u;
      sr += static_cast<float>((c >> 16) & 0xFF);
      sg += static_cast<float>((c >> 8) & 0xFF);
      sb += static_cast<float>(c & 0xFF);
      ++e;
    }
    float cnt = static_cast<float>(e - s);
    out.xyz.push_back(sx / cnt);
    out.xyz.push_back(sy / cnt);
    out.xyz.push_back(sz / cnt);
    uint32_t r = static_cast<uint32_t>(sr / cnt), g = static_cast<uint32_t>(sg / cnt),
             b = static_cast<uint32_t>(sb / cnt);
    out.rgb.push_back((r << 16) | (g << 8) | b);
    s = e;
  }
  return true;
}

/* ------------------------------------------------------------------------------------------- */
/* symmetric 3x3 eigen (double).  The reference uses Eigen::SelfAdjointEigenSolver<Matrix3d>     */
/* (third_party/pcl_shot_na_lrf/shot_na_lrf.hpp:95); restated as cyclic Jacobi, ascending.       */
/* ------------------------------------------------------------------------------------------- */
void eig3_sym(const double Ain[3][3], double w[3], double V[3][3]) {
  double A[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      A[i][j] = Ain[i][j];
      V[i][j] = (i == j) ? 1.0 : 0.0;
    }
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = std::fabs(A[0][1]) + std::fabs(A[0][2]) + std::fabs(A[1][2]);
    if (off == 0.0) break;
    for (int p = 0; p < 2; ++p)
      for (int q = p + 1; q < 3; ++q) {
        double apq = A[p][q];
        if (apq == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        if (!std::isfinite(theta)) t = 0.0;
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 3; ++k) { /* A <- A * J */
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 3; ++k) { /* A <- J^T * A */
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        A[p][q] = A[q][p] = 0.0;
        for (int k = 0; k < 3; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int ord[3] = {0, 1, 2};
  double d[3] = {A[0][0], A[1][1], A[2][2]};
  std::sort(ord, ord + 3, [&](int a, int b) { return d[a] < d[b] || (d[a] == d[b] && a < b); });
  double Vs[3][3];
  for (int j = 0; j < 3; ++j) {
    w[j] = d[ord[j]];
    for (int i = 0; i < 3; ++i) Vs[i][j] = V[i][ord[j]];
  }
  std::memcpy(V, Vs, sizeof(Vs));
}

/* ------------------------------------------------------------------------------------------- */
/* A.3  SHOT local reference frame (features/features.cpp:238-252 ->                             */
/*      pcl::SHOTLocalReferenceFrameEstimation::getLocalRF; anchor shot_na_lrf.hpp:48-178)      */
/* ------------------------------------------------------------------------------------------- */
void shot_lrf(const float* surf, const std::vector<Nbr>& nb, const float* kp, double radius, float rf[9]) {
  std::vector<double> vij;
  vij.reserve(nb.size() * 3);
  double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  double sum = 0.0;
  int valid = 0;
  for (const Nbr& nbr : nb) {
    const float* pt = surf + 3 * nbr.idx;
    if (pt[0] == kp[0] && pt[1] == kp[1] && pt[2] == kp[2]) continue; /* shot_na_lrf.hpp:69 */
    double v[3] = {double(pt[0] - kp[0]), double(pt[1] - kp[1]), double(pt[2] - kp[2])};
    double distance = radius - std::sqrt(double(nbr.d2)); /* :76 */
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) cov[i][j] += distance * (v[i] * v[j]);
    sum += distance;
    vij.insert(vij.end(), v, v + 3);
    valid++;
  }
  if (valid < 5) { /* :85-91 */
    for (int i = 0; i < 9; ++i) rf[i] = kNaNf;
    return;
  }
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) cov[i][j] /= sum;
  double w[3], V[3][3];
  eig3_sym(cov, w, V);
  if (!std::isfinite(w[0]) || !std::isfinite(w[1]) || !std::isfinite(w[2])) {
    for (int i = 0; i < 9; ++i) rf[i] = kNaNf;
    return;
  }
  double v1[3] = {V[0][2], V[1][2], V[2][2]}; /* largest eigenvalue -> x */
  double v3[3] = {V[0][0], V[1][0], V[2][0]}; /* smallest -> z */
  int plusNormal = 0, plusTangent = 0;
  for (int ne = 0; ne < valid; ++ne) {
    const double* r = &vij[3 * ne];
    double dp = r[0] * v1[0] + r[1] * v1[1] + r[2] * v1[2];
    if (dp >= 0) plusTangent++;
    dp = r[0] * v3[0] + r[1] * v3[1] + r[2] * v3[2];
    if (dp >= 0) plusNormal++;
  }
  auto disambiguate = [&](int plus, double* v) {
    plus = 2 * plus - valid;
    if (plus == 0) {
      const int points = 5;
      int medianIndex = valid / 2;
      for (int i = -points / 2; i <= points / 2; ++i) {
        const double* r = &vij[3 * (medianIndex - i)];
        if (r[0] * v[0] + r[1] * v[1] + r[2] * v[2] > 0) plus++;
      }
      if (plus < points / 2 + 1)
        for (int a = 0; a < 3; ++a) v[a] *= -1;
    } else if (plus < 0)
      for (int a = 0; a < 3; ++a) v[a] *= -1;
  };
  disambiguate(plusTangent, v1);
  disambiguate(plusNormal, v3);
  float x[3] = {float(v1[0]), float(v1[1]), float(v1[2])};
  float z[3] = {float(v3[0]), float(v3[1]), float(v3[2])};
  float y[3] = {z[1] * x[2] - z[2] * x[1], z[2] * x[0] - z[0] * x[2], z[0] * x[1] - z[1] * x[0]};
  for (int a = 0; a < 3; ++a) {
    rf[a] = x[a];
    rf[3 + a] = y[a];
    rf[6 + a] = z[a];
  }
}

/* ------------------------------------------------------------------------------------------- */
/* Normals — ImplicitShapeModel::computeNormals implicit_shape_model.cpp:940-1037, */
/* unorganized clouds.  NormalEstimationOMPWithEigVals::computeFeature                           */
/* third_party/pcl_normal_3d_omp_with_eigenvalues/normal_3d_omp_with_eigenvalues.hpp:63-141,     */
/* computePointNormalMod / eigen33Mod / flipNormalTowardsViewpointMod .h:60-181,                 */
/* NormalOrientation::processSHOTLRF utils/normal_orientation.cpp:48-110.                        */
/* pcl::computeMeanAndCovarianceMatrix and pcl::computeRoots are PCL 1.10 (not in the reference  */
/* tree): restated from the published sources -> "parity unpinned".                              */
/* ------------------------------------------------------------------------------------------- */
inline void compute_roots2(float b, float c, float roots[3]) { /* pcl/common/impl/eigen.hpp computeRoots2 */
  roots[0] = 0.0f;
  float d = float(b * b - 4.0 * c);
  if (d < 0.0) d = 0.0f;
  float sd = std::sqrt(d);
  roots[2] = 0.5f * (b + sd);
  roots[1] = 0.5f * (b - sd);
}

inline void compute_roots(const float m[3][3], float roots[3]) { /* pcl/common/impl/eigen.hpp computeRoots */
  float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] -
             m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] -
             m[1][2] * m[1][2];
  float c2 = m[0][0] + m[1][1] + m[2][2];
  if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) {
    compute_roots2(c2, c1, roots);
    return;
  }
  const float s_inv3 = float(1.0 / 3.0);
  const float s_sqrt3 = std::sqrt(3.0f);
  float c2_over_3 = c2 * s_inv3;
  float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
  if (a_over_3 > 0.0f) a_over_3 = 0.0f;
  float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
  float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
  if (q > 0.0f) q = 0.0f;
  float rho = std::sqrt(-a_over_3);
  float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
  float cos_theta = std::cos(theta);
  float sin_theta = std::sin(theta);
  roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
  roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
  roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
  if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  if (roots[1] >= roots[2]) {
    std::swap(roots[1], roots[2]);
    if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
  }
  if (roots[0] <= 0) compute_roots2(c2, c1, roots);
}

/* eigen33Mod (.h:60-95): eigenvalues ascending + eigenvector of the smallest one */
inline void eigen33_mod(const float mat[3][3], float evals[3], float evec[3]) {
  float scale = 0.f;
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(mat[i][j]));
  if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
  float sm[3][3];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) sm[i][j] = mat[i][j] / scale;
  float rt[3];
  compute_roots(sm, rt);
  for (int i = 0; i < 3; ++i) evals[i] = rt[i] * scale;
  for (int i = 0; i < 3; ++i) sm[i][i] -= rt[0];
  auto cross = [](const float* a, const float* b, float* o) {
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
  };
  float v1[3], v2[3], v3[3];
  cross(sm[0], sm[1], v1);
  cross(sm[0], sm[2], v2);
  cross(sm[1], sm[2], v3);
  float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];
  float l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2];
  float l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
  const float* v;
  float l;
  if (l1 >= l2 && l1 >= l3) { v = v1; l = l1; }
  else if (l2 >= l1 && l2 >= l3) { v = v2; l = l2; }
  else { v = v3; l = l3; }
  float inv = std::sqrt(l);
  for (int a = 0; a < 3; ++a) evec[a] = v[a] / inv;
}

/* computePointNormalMod (.h:112-150) over the neighbour list in radius-search order; float accumulators as
 * pcl::computeMeanAndCovarianceMatrix<PointT, float> has them */
bool pca_normal(const float* pts, const std::vector<Nbr>& nb, float n[3], float* curvature) {
  if (nb.size() < 3) {
    n[0] = n[1] = n[2] = *curvature = kNaNf;
    return false;
  }
  float accu[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (const Nbr& e : nb) {
    const float* p = pts + 3 * e.idx;
    accu[0] += p[0] * p[0];
    accu[1] += p[0] * p[1];
    accu[2] += p[0] * p[2];
    accu[3] += p[1] * p[1];
    accu[4] += p[1] * p[2];
    accu[5] += p[2] * p[2];
    accu[6] += p[0];
    accu[7] += p[1];
    accu[8] += p[2];
  }
  const float cnt = float(nb.size());
  for (float& a : accu) a /= cnt;
  float cov[3][3];
  cov[0][0] = accu[0] - accu[6] * accu[6];
  cov[0][1] = accu[1] - accu[6] * accu[7];
  cov[0][2] = accu[2] - accu[6] * accu[8];
  cov[1][1] = accu[3] - accu[7] * accu[7];
  cov[1][2] = accu[4] - accu[7] * accu[8];
  cov[2][2] = accu[5] - accu[8] * accu[8];
  cov[1][0] = cov[0][1];
  cov[2][0] = cov[0][2];
  cov[2][1] = cov[1][2];
  float ev[3];
  eigen33_mod(cov, ev, n);
  float eig_sum = cov[0][0] + cov[1][1] + cov[2][2];
  *curvature = eig_sum != 0 ? std::fabs(ev[0] / eig_sum) : 0.f;
  return true;
}

inline void flip_towards_viewpoint(const float* p, float vx, float vy, float vz, float n[3]) { /* .h:162-181 */
  vx -= p[0];
  vy -= p[1];
  vz -= p[2];
  float cos_theta = (vx * n[0] + vy * n[1] + vz * n[2]);
  if (cos_theta < 0) {
    n[0] *= -1;
    n[1] *= -1;
    n[2] *= -1;
  }
}

/* pts: the cloud after removeNaNFromPointCloud (n finite points).  normals_out n x 3, curvature_out n (may be NULL). */
void compute_normals_cloud(const pcdb_params& P, const float* pts, int n, float* normals_out, float* curvature_out) {
  const double radius = double(P.normal_radius); /* float member -> setRadiusSearch(double) */
  const float r2 = radius_sq(radius);
  const int method = P.consistent_normals_method;
  std::vector<float> shifted;
  const float* src = pts;
  if (method == 1) { /* :987-1014: remove the centroid (pcl::compute3DCentroid, float accumulators), flip, invert */
    float c[3] = {0, 0, 0};
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a) c[a] += pts[3 * i + a];
    for (int a = 0; a < 3; ++a) c[a] /= float(n);
    shifted.resize(size_t(n) * 3);
    for (int i = 0; i < n; ++i)
      for (int a = 0; a < 3; ++a) shifted[3 * i + a] = pts[3 * i + a] - c[a];
    src = shifted.data();
  }
  CloudGrid g;
  g.build(src, n, radius);
  std::vector<float> curv(n);
  std::vector<unsigned char> lrf_bad(n, 0);
#pragma omp parallel for schedule(dynamic, 64)
  for (int i = 0; i < n; ++i) {
    std::vector<Nbr> nb;
    g.query(src + 3 * i, r2, nb);
    float* no = normals_out + 3 * i;
    if (pca_normal(src, nb, no, &curv[i])) flip_towards_viewpoint(src + 3 * i, 0.f, 0.f, 0.f, no);
    if (method == 1)
      for (int a = 0; a < 3; ++a) no[a] *= -1;
    if (method == 2) { /* :1015-1019 + normal_orientation.cpp:56-82: normal = inverted z axis of the SHOT LRF */
      float rf[9];
      shot_lrf(src, nb, src + 3 * i, radius, rf);
      if (std::isfinite(rf[0]) && std::isfinite(rf[3]) && std::isfinite(rf[6])) {
        no[0] = -rf[6];
        no[1] = -rf[7];
        no[2] = -rf[8];
      } else
        lrf_bad[i] = 1;
    }
  }
  if (method == 2) {
    /* normal_orientation.cpp:85-107, kept as written: the loop recomputes the normals of points 0..n_invalid-1
     * (pointCloud->at(idx), orientedNormals->at(idx)) instead of the invalid ones, without viewpoint flip, and the
     * pcl::Normal(x,y,z) constructor zeroes their curvature */
    int n_invalid = 0;
    for (int i = 0; i < n; ++i) n_invalid += lrf_bad[i];
    for (int idx = 0; idx < n_invalid; ++idx) {
      std::vector<Nbr> nb;
      g.query(src + 3 * idx, r2, nb);
      float c;
      pca_normal(src, nb, normals_out + 3 * idx, &c);
      curv[idx] = 0.f;
    }
  }
  if (curvature_out) std::memcpy(curvature_out, curv.data(), sizeof(float) * n);
}

/* ------------------------------------------------------------------------------------------- */
/* A.5  RGB -> CIELab (features/features_cshot.cpp:52-71 LUTs,                                   */
/*      third_party/pcl_color_conversion/color_conversion.cpp:19-94)                             */
/* ------------------------------------------------------------------------------------------- */
struct LabLut {
  float srgb[256];
  float sxyz[4000];
  LabLut() {
    for (int i = 0; i < 256; i++) {
      float f = static_cast<float>(i) / 255.0f;
      if (f > 0.04045)
        srgb[i] = powf((f + 0.055f) / 1.055f, 2.4f);
      else
        srgb[i] = f / 12.92f;
    }
    for (int i = 0; i < 4000; i++) {
      float f = static_cast<float>(i) / 4000.0f;
      if (f > 0.008856)
        sxyz[i] = static_cast<float>(powf(f, 0.3333f));
      else
        sxyz[i] = static_cast<float>((7.787 * f) + (16.0 / 116.0));
    }
  }
};
const LabLut& lab_lut() {
  static LabLut l;
  return l;
}

/* L in [0,100], A,B in [-120,120] (PCL RGB2CIELAB); normalisation by 100/120/120 done by the caller */
void rgb2lab(uint32_t rgb, float& L, float& A, float& B2) {
  const LabLut& lut = lab_lut();
  float fr = lut.srgb[(rgb >> 16) & 0xFF];
  float fg = lut.srgb[(rgb >> 8) & 0xFF];
  float fb = lut.srgb[rgb & 0xFF];
  const float x = fr * 0.412453f + fg * 0.357580f + fb * 0.180423f;
  const float y = fr * 0.212671f + fg * 0.715160f + fb * 0.072169f;
  const float z = fr * 0.019334f + fg * 0.119193f + fb * 0.950227f;
  float vx = x / 0.95047f;
  float vy = y;
  float vz = z / 1.08883f;
  vx = lut.sxyz[int(vx * 4000)];
  vy = lut.sxyz[int(vy * 4000)];
  vz = lut.sxyz[int(vz * 4000)];
  L = 116.0f * vy - 16.0f;
  if (L > 100) L = 100.0f;
  A = 500.0f * (vx - vy);
  if (A > 120)
    A = 120.0f;
  else if (A < -120)
    A = -120.0f;
  B2 = 200.0f * (vy - vz);
  if (B2 > 120)
    B2 = 120.0f;
  else if (B2 < -120)
    B2 = -120.0f;
}

/* colour distance of features_short_cshot.cpp:194-198 / color_conversion.cpp:85-93 (float fabs overloads) */
double color_distance(float L, float a, float b, float LRef, float aRef, float bRef) {
  double cd = (std::fabs(LRef - L) + ((std::fabs(aRef - a) + std::fabs(bRef - b)) / 2)) / 3;
  if (cd > 1.0) cd = 1.0;
  if (cd < 0.0) cd = 0.0;
  return cd;
}

/* ------------------------------------------------------------------------------------------- */
/* A.4/A.5  SHOT-352 / CSHOT-1344 (features/features_shot.cpp:28-81, features_cshot.cpp:28-103  */
/*          -> pcl::SHOTEstimation::computePointSHOT / interpolateSingle/DoubleChannel)         */
/* ------------------------------------------------------------------------------------------- */
const double PST_PI = 3.1415926535897932384626433832795;
const double PST_RAD_45 = 0.78539816339744830961566084581988;
const double PST_RAD_90 = 1.5707963267948966192313216916398;
const double PST_RAD_135 = 2.3561944901923449288469825374596;
const double PST_RAD_PI_7_8 = 2.7488935718910690836548129603691;

inline float dot3f(const float* a, const float* b) {
  float r = a[0] * b[0];
  r += a[1] * b[1];
  r += a[2] * b[2];
  return r;
}

void shot_describe(bool color, const float* surf, const float* normals, const uint32_t* surf_rgb,
                   const std::vector<Nbr>& nb, const float* kp, uint32_t kp_rgb, const float* rf, double radius,
                   float* shot) {
  const int nr_shape = 10, nr_color = 30, D = color ? PCDB_CSHOT_DIM : PCDB_SHOT_DIM;
  const int stride_color = 32 * (nr_shape + 1);
  const double radius3_4 = (radius * 3) / 4, radius1_4 = radius / 4, radius1_2 = radius / 2;
  if (nb.size() < 5 || !std::isfinite(rf[0]) || !std::isfinite(rf[3]) || !std::isfinite(rf[6])) {
    for (int i = 0; i < D; ++i) shot[i] = kNaNf;
    return;
  }
  const float* fx = rf;
  const float* fy = rf + 3;
  const float* fz = rf + 6;
  std::vector<double> bdS(nb.size()), bdC;
  for (size_t i = 0; i < nb.size(); ++i) {
    const float* nrm = normals + 3 * nb[i].idx;
    if (!finite3(nrm))
      bdS[i] = std::numeric_limits<double>::quiet_NaN();
    else {
      double cosineDesc = dot3f(nrm, fz);
      if (cosineDesc > 1.0) cosineDesc = 1.0;
      if (cosineDesc < -1.0) cosineDesc = -1.0;
      bdS[i] = ((1.0 + cosineDesc) * nr_shape) / 2;
    }
  }
  if (color) {
    bdC.resize(nb.size());
    float LRef, aRef, bRef;
    rgb2lab(kp_rgb, LRef, aRef, bRef);
    LRef /= 100.0f;
    aRef /= 120.0f;
    bRef /= 120.0f;
    for (size_t i = 0; i < nb.size(); ++i) {
      float L, a, b;
      rgb2lab(surf_rgb ? surf_rgb[nb[i].idx] : 0u, L, a, b);
      L /= 100.0f;
      a /= 120.0f;
      b /= 120.0f;
      bdC[i] = color_distance(L, a, b, LRef, aRef, bRef) * nr_color;
    }
  }
  for (int i = 0; i < D; ++i) shot[i] = 0.0f;
  for (size_t i = 0; i < nb.size(); ++i) {
    if (!std::isfinite(bdS[i])) continue;
    const float* p = surf + 3 * nb[i].idx;
    float delta[3] = {p[0] - kp[0], p[1] - kp[1], p[2] - kp[2]};
    double distance = std::sqrt(double(nb[i].d2));
    if (std::fabs(distance - 0.0) < 1E-15) continue;
    double xIn = dot3f(delta, fx), yIn = dot3f(delta, fy), zIn = dot3f(delta, fz);
    if (std::fabs(yIn) < 1E-30) yIn = 0;
    if (std::fabs(xIn) < 1E-30) xIn = 0;
    if (std::fabs(zIn) < 1E-30) zIn = 0;
    unsigned char bit4 = ((yIn > 0) || ((yIn == 0.0) && (xIn < 0))) ? 1 : 0;
    unsigned char bit3 = static_cast<unsigned char>(((xIn > 0) || ((xIn == 0.0) && (yIn > 0))) ? !bit4 : bit4);
    int desc_index = (bit4 << 3) + (bit3 << 2);
    desc_index = desc_index << 1;
    if ((xIn * yIn > 0) || (xIn == 0.0))
      desc_index += (std::fabs(xIn) >= std::fabs(yIn)) ? 0 : 4;
    else
      desc_index += (std::fabs(xIn) > std::fabs(yIn)) ? 4 : 0;
    desc_index += zIn > 0 ? 1 : 0;
    desc_index += (distance > radius1_2) ? 2 : 0;

    int stepS = static_cast<int>(std::floor(bdS[i] + 0.5));
    int volS = desc_index * (nr_shape + 1);
    double bS = bdS[i] - stepS;
    double wS = (1 - std::fabs(bS));
    if (bS > 0)
      shot[volS + ((stepS + 1) % nr_shape)] += static_cast<float>(bS);
    else
      shot[volS + ((stepS - 1 + nr_shape) % nr_shape)] -= static_cast<float>(bS);
    int stepC = 0, volC = 0;
    double wC = 0;
    if (color) {
      stepC = static_cast<int>(std::floor(bdC[i] + 0.5));
      volC = stride_color + desc_index * (nr_color + 1);
      double bC = bdC[i] - stepC;
      wC = (1 - std::fabs(bC));
      if (bC > 0)
        shot[volC + ((stepC + 1) % nr_color)] += static_cast<float>(bC);
      else
        shot[volC + ((stepC - 1 + nr_color) % nr_color)] -= static_cast<float>(bC);
    }
    auto add_nb = [&](int di, double v) { /* neighbouring volume `di` receives v in both channels */
      shot[di * (nr_shape + 1) + stepS] += static_cast<float>(v);
      if (color) shot[stride_color + di * (nr_color + 1) + stepC] += static_cast<float>(v);
    };
    /* radial */
    if (distance > radius1_2) {
      double rd = (distance - radius3_4) / radius1_2;
      if (distance > radius3_4) {
        wS += 1 - rd;
        wC += 1 - rd;
      } else {
        wS += 1 + rd;
        wC += 1 + rd;
        add_nb(desc_index - 2, -rd);
      }
    } else {
      double rd = (distance - radius1_4) / radius1_2;
      if (distance < radius1_4) {
        wS += 1 + rd;
        wC += 1 + rd;
      } else {
        wS += 1 - rd;
        wC += 1 - rd;
        add_nb(desc_index + 2, rd);
      }
    }
    /* elevation */
    double inclinationCos = zIn / distance;
    if (inclinationCos < -1.0) inclinationCos = -1.0;
    if (inclinationCos > 1.0) inclinationCos = 1.0;
    double inclination = std::acos(inclinationCos);
    if (inclination > PST_RAD_90 || (std::fabs(inclination - PST_RAD_90) < 1e-30 && zIn <= 0)) {
      double id = (inclination - PST_RAD_135) / PST_RAD_90;
      if (inclination > PST_RAD_135) {
        wS += 1 - id;
        wC += 1 - id;
      } else {
        wS += 1 + id;
        wC += 1 + id;
        add_nb(desc_index + 1, -id);
      }
    } else {
      double id = (inclination - PST_RAD_45) / PST_RAD_90;
      if (inclination < PST_RAD_45) {
        wS += 1 + id;
        wC += 1 + id;
      } else {
        wS += 1 - id;
        wC += 1 - id;
        add_nb(desc_index - 1, id);
      }
    }
    /* azimuth */
    if (yIn != 0.0 || xIn != 0.0) {
      double azimuth = std::atan2(yIn, xIn);
      int sel = desc_index >> 2;
      double ad = (azimuth - (-PST_RAD_PI_7_8 + PST_RAD_45 * sel)) / PST_RAD_45;
      ad = (std::max)(-0.5, std::min(ad, 0.5));
      if (ad > 0) {
        wS += 1 - ad;
        wC += 1 - ad;
        add_nb((desc_index + 4) % 32, ad);
      } else {
        wS += 1 + ad;
        wC += 1 + ad;
        add_nb((desc_index - 4 + 32) % 32, -ad);
      }
    }
    shot[volS + stepS] += static_cast<float>(wS);
    if (color) shot[volC + stepC] += static_cast<float>(wC);
  }
  double acc = 0.0;
  for (int j = 0; j < D; ++j) acc += shot[j] * shot[j];
  acc = std::sqrt(acc);
  for (int j = 0; j < D; ++j) shot[j] /= static_cast<float>(acc);
  (void)PST_PI;
}

/* ------------------------------------------------------------------------------------------- */
/* A.6  FLANN distance functors (utils/distance.h:42-75 -> flann::L2 / flann::ChiSquareDistance)*/
/* ------------------------------------------------------------------------------------------- */
inline float dist_l2(const float* a, const float* b, int n) {
  float result = 0;
  int i = 0;
  for (; i + 3 < n; i += 4) {
    float d0 = a[i] - b[i], d1 = a[i + 1] - b[i + 1], d2 = a[i + 2] - b[i + 2], d3 = a[i + 3] - b[i + 3];
    result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
  }
  for (; i < n; ++i) {
    float d0 = a[i] - b[i];
    result += d0 * d0;
  }
  return result;
}
inline float dist_chi2(const float* a, const float* b, int n) {
  float result = 0;
  for (int i = 0; i < n; ++i) {
    float sum = a[i] + b[i];
    if (sum > 0) {
      float diff = a[i] - b[i];
      result += diff * diff / sum;
    }
  }
  return result;
}
inline float dist_fn(int type, const float* a, const float* b, int n) {
  return type == PCDB_DIST_CHISQUARED ? dist_chi2(a, b, n) : dist_l2(a, b, n);
}

/* four codewords at once (independent accumulator chains; each chain keeps FLANN's order) */
inline void dist_l2_x4(const float* q, const float* c0, const float* c1, const float* c2, const float* c3, int n,
                       float out[4]) {
  float r0 = 0, r1 = 0, r2 = 0, r3 = 0;
  int i = 0;
  for (; i + 3 < n; i += 4) {
    float a0 = q[i], a1 = q[i + 1], a2 = q[i + 2], a3 = q[i + 3];
    {
      float d0 = a0 - c0[i], d1 = a1 - c0[i + 1], d2 = a2 - c0[i + 2], d3 = a3 - c0[i + 3];
      r0 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    {
      float d0 = a0 - c1[i], d1 = a1 - c1[i + 1], d2 = a2 - c1[i + 2], d3 = a3 - c1[i + 3];
      r1 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    {
      float d0 = a0 - c2[i], d1 = a1 - c2[i + 1], d2 = a2 - c2[i + 2], d3 = a3 - c2[i + 3];
      r2 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
    {
      float d0 = a0 - c3[i], d1 = a1 - c3[i + 1], d2 = a2 - c3[i + 2], d3 = a3 - c3[i + 3];
      r3 += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
    }
  }
  for (; i < n; ++i) {
    float d;
    d = q[i] - c0[i];
    r0 += d * d;
    d = q[i] - c1[i];
    r1 += d * d;
    d = q[i] - c2[i];
    r2 += d * d;
    d = q[i] - c3[i];
    r3 += d * d;
  }
  out[0] = r0;
  out[1] = r1;
  out[2] = r2;
  out[3] = r3;
}

/* exact kNN, ascending (distance, row) — FLANN's tie order is tree dependent */
struct Cand {
  float d;
  int idx;
};
inline bool cand_less(const Cand& a, const Cand& b) { return a.d < b.d || (a.d == b.d && a.idx < b.idx); }

void knn_one(const float* q, const float* words, int64_t N, int D, int k, int type, Cand* best /* k */, int& found) {
  found = 0;
  auto push = [&](float d, int idx) {
    Cand c{d, idx};
    if (found < k) {
      int p = found++;
      while (p > 0 && cand_less(c, best[p - 1])) {
        best[p] = best[p - 1];
        --p;
      }
      best[p] = c;
    } else if (cand_less(c, best[k - 1])) {
      int p = k - 1;
      while (p > 0 && cand_less(c, best[p - 1])) {
        best[p] = best[p - 1];
        --p;
      }
      best[p] = c;
    }
  };
  int64_t j = 0;
  if (type == PCDB_DIST_EUCLIDEAN) {
    for (; j + 3 < N; j += 4) {
      float o[4];
      dist_l2_x4(q, words + j * D, words + (j + 1) * D, words + (j + 2) * D, words + (j + 3) * D, D, o);
      for (int t = 0; t < 4; ++t)
        if (found < k || o[t] <= best[k - 1].d) push(o[t], int(j + t));
    }
  }
  for (; j < N; ++j) push(dist_fn(type, q, words + j * D, D), int(j));
}

/* ------------------------------------------------------------------------------------------- */
/* quaternion helpers restating utils/utils.cpp:136-178,342-394,560-574 with                    */
/* boost::math::quaternion<float> operator*= evaluation order                                   */
/* ------------------------------------------------------------------------------------------- */
struct Quat {
  float a, b, c, d; /* w x y z */
};
inline Quat qmul(const Quat& l, const Quat& r) {
  Quat o;
  o.a = +l.a * r.a - l.b * r.b - l.c * r.c - l.d * r.d;
  o.b = +l.a * r.b + l.b * r.a + l.c * r.d - l.d * r.c;
  o.c = +l.a * r.c - l.b * r.d + l.c * r.a + l.d * r.b;
  o.d = +l.a * r.d + l.b * r.c - l.c * r.b + l.d * r.a;
  return o;
}
inline Quat qconj(const Quat& q) { return Quat{q.a, -q.b, -q.c, -q.d}; }

/* Utils::getRotQuaternion: Eigen's column-major matrix with columns = axes is read row-major by
 * matrix2Quat, i.e. rows = axes (utils.cpp:136-152,384-394). */
Quat lrf_quat(const float* rf) {
  float m[3][3] = {{rf[0], rf[1], rf[2]}, {rf[3], rf[4], rf[5]}, {rf[6], rf[7], rf[8]}};
  float quat[4] = {0, 0, 0, 0}; /* x y z w */
  float trace = m[0][0] + m[1][1] + m[2][2];
  float root;
  if (trace > 0.0f) {
    root = sqrtf(trace + 1.0f);
    quat[3] = 0.5f * root;
    root = 0.5f / root;
    quat[0] = (m[2][1] - m[1][2]) * root;
    quat[1] = (m[0][2] - m[2][0]) * root;
    quat[2] = (m[1][0] - m[0][1]) * root;
  } else {
    static const size_t next[3] = {1, 2, 0};
    size_t i = 0;
    if (m[1][1] > m[0][0]) i = 1;
    if (m[2][2] > m[i][i]) i = 2;
    size_t j = next[i];
    size_t k = next[j];
    root = sqrtf(m[i][i] - m[j][j] - m[k][k] + 1.0);
    quat[i] = 0.5f * root;
    root = 0.5f / root;
    quat[3] = (m[k][j] - m[j][k]) * root;
    quat[j] = (m[j][i] + m[i][j]) * root;
    quat[k] = (m[k][i] + m[i][k]) * root;
  }
  return Quat{quat[3], quat[0], quat[1], quat[2]};
}
inline void quat_rotate(const Quat& q, const float* p, float* out) { /* q p q*  (rotateInto) */
  Quat t = qmul(qmul(q, Quat{0, p[0], p[1], p[2]}), qconj(q));
  out[0] = t.b;
  out[1] = t.c;
  out[2] = t.d;
}
inline void quat_rotate_inv(const Quat& q, const float* p, float* out) { /* q* p q  (rotateBack) */
  Quat t = qmul(qmul(qconj(q), Quat{0, p[0], p[1], p[2]}), q);
  out[0] = t.b;
  out[1] = t.c;
  out[2] = t.d;
}

/* ------------------------------------------------------------------------------------------- */
/* Approximate activation: FLANN's randomized kd-forest (KDTreeIndexParams(FLANNNumKDTrees = 4),  */
/* SearchParams(checks = 128); utils/flann_helper.cpp:59-65, activation_strategy_knn.h:66-70).   */
/* FLANN 1.9.1 is not in the reference tree: this restates the published algorithm               */
/* (flann/algorithms/kdtree_index.h: meanSplit on a 100-point sample, cut dimension drawn from   */
/* the 5 highest-variance dimensions, best-bin-first search over all trees with one shared heap). */
/* It is the HONEST COST STAND-IN for what the reference executes by default; its neighbour sets */
/* depend on the random draws and are never used for parity.             */
/* ------------------------------------------------------------------------------------------- */
struct KdForest {
  struct Node {
    int divfeat = -1; /* -1: leaf */
    float divval = 0.f;
    int child1 = -1, child2 = -1; /* leaf: child1 = point index */
  };
  int D = 0;
  const float* data = nullptr;
  std::vector<std::vector<Node>> trees;
  std::vector<int> roots;
  int checks = 128;

  static constexpr int SAMPLE_MEAN = 100, RAND_DIM = 5;

  int divide(std::vector<Node>& nodes, int* ind, int count, std::mt19937& rng) {
    int id = int(nodes.size());
    nodes.emplace_back();
    if (count == 1) {
      nodes[id].child1 = ind[0];
      return id;
    }
    /* meanSplit */
    std::vector<double> mean(D, 0.0), var(D, 0.0);
    int cnt = std::min(int(SAMPLE_MEAN) + 1, count);
    for (int j = 0; j < cnt; ++j) {
      const float* v = data + size_t(ind[j]) * D;
      for (int k = 0; k < D; ++k) mean[k] += v[k];
    }
    for (int k = 0; k < D; ++k) mean[k] /= cnt;
    for (int j = 0; j < cnt; ++j) {
      const float* v = data + size_t(ind[j]) * D;
      for (int k = 0; k < D; ++k) {
        double d = v[k] - mean[k];
        var[k] += d * d;
      }
    }
    /* selectDivision: one of the RAND_DIM dimensions of highest variance */
    int topind[RAND_DIM];
    int num = 0;
    for (int i = 0; i < D; ++i) {
      if (num < RAND_DIM || var[i] > var[topind[num - 1]]) {
        if (num < RAND_DIM) topind[num++] = i;
        else topind[num - 1] = i;
        int j = num - 1;
        while (j > 0 && var[topind[j]] > var[topind[j - 1]]) {
          std::swap(topind[j], topind[j - 1]);
          --j;
        }
      }
    }
    int cutfeat = topind[std::uniform_int_distribution<int>(0, num - 1)(rng)];
    float cutval = float(mean[cutfeat]);
    /* planeSplit */
    int left = 0, right = count - 1;
    for (;;) {
      while (left <= right && data[size_t(ind[left]) * D + cutfeat] < cutval) ++left;
      while (left <= right && data[size_t(ind[right]) * D + cutfeat] >= cutval) --right;
      if (left > right) break;
      std::swap(ind[left], ind[right]);
      ++left;
      --right;
    }
    int lim1 = left;
    right = count - 1;
    for (;;) {
      while (left <= right && data[size_t(ind[left]) * D + cutfeat] <= cutval) ++left;
      while (left <= right && data[size_t(ind[right]) * D + cutfeat] > cutval) --right;
      if (left > right) break;
      std::swap(ind[left], ind[right]);
      ++left;
      --right;
    }
    int lim2 = left;
    int index;
    if (lim1 > count / 2) index = lim1;
    else if (lim2 < count / 2) index = lim2;
    else index = count / 2;
    if (lim1 == count || lim2 == 0) index = count / 2;
    nodes[id].divfeat = cutfeat;
    nodes[id].divval = cutval;
    int c1 = divide(nodes, ind, index, rng);
    int c2 = divide(nodes, ind + index, count - index, rng);
    nodes[id].child1 = c1;
    nodes[id].child2 = c2;
    return id;
  }

  void build(const float* words, int64_t N, int dim, int n_trees, unsigned seed) {
    data = words;
    D = dim;
    trees.assign(n_trees, {});
    roots.assign(n_trees, -1);
#pragma omp parallel for schedule(dynamic, 1)
    for (int t = 0; t < n_trees; ++t) {
      std::mt19937 rng(seed + 7919u * unsigned(t));
      std::vector<int> ind(N);
      for (int64_t i = 0; i < N; ++i) ind[i] = int(i);
      std::shuffle(ind.begin(), ind.end(), rng);
      trees[t].reserve(size_t(2 * N));
      roots[t] = divide(trees[t], ind.data(), int(N), rng);
    }
  }

  struct Branch {
    float mind;
    int tree, node;
    bool operator<(const Branch& o) const { return mind > o.mind; } /* min-heap */
  };

  /* L2 functor with FLANN's early exit on worst_dist (dist.h: checked every 4 elements) */
  static float l2_bounded(const float* a, const float* b, int n, float worst) {
    float result = 0.f;
    int i = 0;
    for (; i + 3 < n; i += 4) {
      float d0 = a[i] - b[i], d1 = a[i + 1] - b[i + 1], d2 = a[i + 2] - b[i + 2], d3 = a[i + 3] - b[i + 3];
      result += d0 * d0 + d1 * d1 + d2 * d2 + d3 * d3;
      if (worst > 0 && result > worst) return result;
    }
    for (; i < n; ++i) {
      float d = a[i] - b[i];
      result += d * d;
    }
    return result;
  }

  void search_level(const float* q, int t, int node, float mind, int& check_count, std::vector<Branch>& heap,
                    std::vector<unsigned char>& checked, std::vector<int>& touched, Cand* best, int kk,
                    int& found) const {
    for (;;) {
      const Node& nd = trees[t][node];
      if (nd.divfeat < 0) {
        const int idx = nd.child1;
        if (checked[idx] || (check_count >= checks && found >= kk)) return;
        checked[idx] = 1;
        touched.push_back(idx);
        ++check_count;
        const float worst = found >= kk ? best[kk - 1].d : -1.f;
        const float d = l2_bounded(data + size_t(idx) * D, q, D, worst);
        Cand c{d, idx};
        if (found < kk || cand_less(c, best[kk - 1])) {
          int pos = std::min(found, kk - 1);
          best[pos] = c;
          if (found < kk) ++found;
          while (pos > 0 && cand_less(best[pos], best[pos - 1])) {
            std::swap(best[pos], best[pos - 1]);
            --pos;
          }
        }
        return;
      }
      if (found >= kk && best[kk - 1].d < mind) return; /* result_set.worstDist() < mindist */
      const float diff = q[nd.divfeat] - nd.divval;
      const int best_child = diff < 0 ? nd.child1 : nd.child2;
      const int other = diff < 0 ? nd.child2 : nd.child1;
      const float new_d = mind + diff * diff;
      if (found < kk || new_d < best[kk - 1].d) {
        heap.push_back({new_d, t, other});
        std::push_heap(heap.begin(), heap.end());
      }
      node = best_child;
    }
  }

  void knn(const float* q, int kk, Cand* best, int& found, std::vector<unsigned char>& checked,
           std::vector<int>& touched) const {
    found = 0;
    int check_count = 0;
    std::vector<Branch> heap;
    heap.reserve(512);
    touched.clear();
    for (int t = 0; t < int(trees.size()); ++t)
      search_level(q, t, roots[t], 0.f, check_count, heap, checked, touched, best, kk, found);
    while (!heap.empty() && (check_count < checks || found < kk)) {
      std::pop_heap(heap.begin(), heap.end());
      Branch b = heap.back();
      heap.pop_back();
      search_level(q, b.tree, b.node, b.mind, check_count, heap, checked, touched, best, kk, found);
    }
    for (int idx : touched) checked[idx] = 0; /* the visited marks are reset through the list of touched leaves */
  }
};

/* ------------------------------------------------------------------------------------------- */
/* model held by the oracle                                                                     */
/* ------------------------------------------------------------------------------------------- */
struct Model {
  pcdb_params prm;
  int64_t N = 0;
  int D = 0;
  std::vector<float> words;
  std::vector<int64_t> vote_off;
  std::vector<float> vote_xyz, vote_weight, vote_bbox, vote_class_weight;
  std::vector<uint32_t> vote_class, vote_instance;
  std::vector<float> kp_train, codeword_weight;
  std::vector<int32_t> codeword_ids;
  std::vector<float> sigma2;
  std::vector<float> dim_first, dim_second; /* Voting::m_dimensions_map (voting.cpp:497-557,619-650), by class id */
  std::shared_ptr<KdForest> forest; /* set: activation runs FLANN-style approximate search (cost stand-in only) */
  double forest_build_ms = 0;
};

/* A.7  CodewordDistribution::castVotes / castVote (codebook/codeword_distribution.cpp:73-167) */
void cast_votes_one(const Model& m, const float* kp, const float* rf, const float* /*desc*/, int row, float dist,
                    std::vector<pcdb_vote>& out) {
  const pcdb_params& P = m.prm;
  Quat rq = lrf_quat(rf);
  for (int64_t v = m.vote_off[row]; v < m.vote_off[row + 1]; ++v) {
    uint32_t classId = m.vote_class[v];
    float classWeight = m.vote_class_weight.empty() ? 1.0f : m.vote_class_weight[v];
    float classSigma = classId < m.sigma2.size() ? m.sigma2[classId] : 1.0f;
    float matching = float((1 / std::sqrt(2 * M_PI * classSigma)) * std::exp(-std::pow(dist, 2) / (2 * classSigma)));
    float voteWeight = m.vote_weight[v];
    float weight = 1.0f;
    weight = P.use_class_weight ? weight * classWeight : weight;
    weight = P.use_vote_weight ? weight * voteWeight : weight;
    weight = P.use_matching_weight ? weight * matching : weight;
    weight = P.use_codeword_weight ? weight * m.codeword_weight[row] : weight;
    if (P.filter_abs_is_int) {
      if (float(std::abs(int(dist))) > 2 * classSigma) continue;
    } else {
      if (std::fabs(dist) > 2 * classSigma) continue; /* :131-135 */
    }
    if (weight < std::numeric_limits<float>::epsilon()) continue;
    pcdb_vote o;
    float rot[3];
    quat_rotate_inv(rq, &m.vote_xyz[3 * v], rot);
    for (int a = 0; a < 3; ++a) {
      o.position[a] = kp[a] + rot[a];
      o.keypoint[a] = kp[a];
      o.keypoint_training[a] = m.kp_train[3 * row + a];
      o.bbox_size[a] = m.vote_bbox[7 * v + 4 + a];
    }
    Quat bq{m.vote_bbox[7 * v], m.vote_bbox[7 * v + 1], m.vote_bbox[7 * v + 2], m.vote_bbox[7 * v + 3]};
    Quat nb = qmul(bq, rq);
    o.bbox_quat[0] = nb.a;
    o.bbox_quat[1] = nb.b;
    o.bbox_quat[2] = nb.c;
    o.bbox_quat[3] = nb.d;
    o.weight = weight;
    o.class_id = classId;
    o.instance_id = m.vote_instance[v];
    o.codeword_id = m.codeword_ids.empty() ? row : m.codeword_ids[row];
    out.push_back(o);
  }
}

/* ------------------------------------------------------------------------------------------- */
/* A.8  VotingMeanShift::iFindMaxima (voting/voting_mean_shift.cpp:39-177) for one class        */
/* ------------------------------------------------------------------------------------------- */
struct ClassVotes {
  std::vector<int64_t> gidx; /* index into the cloud's vote array */
  std::vector<float> pos;    /* 3 per vote */
  std::vector<float> w;      /* working copy (re-weighted in place, voting.cpp:95 copies by value) */
};

void ms_radius(const ClassVotes& cv, const float* c, float r2, std::vector<Nbr>& out) {
  out.clear();
  int n = int(cv.w.size());
  for (int i = 0; i < n; ++i) {
    float d2 = sqdist3(c, &cv.pos[3 * i]);
    if (d2 < r2) out.push_back({d2, i});
  }
  std::sort(out.begin(), out.end(), nbr_less);
}

inline float ms_kernel(int type, float x) {
  if (type == PCDB_KERNEL_GAUSSIAN) {
    float profile = std::exp(-0.5 * x); /* :396-400, double exp narrowed */
    return profile;
  }
  return 1;
}
inline float ms_kernel_derivative(int type, float x) {
  if (type == PCDB_KERNEL_GAUSSIAN) {
    float profile = std::exp(-0.5 * x);
    float derivative = -0.5f * profile;
    return derivative;
  }
  return 1;
}

struct Vec3 {
  float v[3];
};
inline float norm3(const float* a, const float* b) {
  float d0 = a[0] - b[0], d1 = a[1] - b[1], d2 = a[2] - b[2];
  return std::sqrt(d0 * d0 + d1 * d1 + d2 * d2);
}

float ms_density(const pcdb_params& P, float h, ClassVotes& cv, const float* pos, std::vector<int>& members,
                 bool reweight) {
  std::vector<Nbr> nb;
  ms_radius(cv, pos, radius_sq(double(h)), nb);
  members.clear();
  if (nb.empty()) return 0;
  float density = 0;
  for (const Nbr& n : nb) {
    float u = n.d2 / (h * h);
    float weight = ms_kernel(P.ms_kernel, u) * cv.w[n.idx];
    if (reweight) cv.w[n.idx] = weight;
    members.push_back(n.idx);
    density += weight;
  }
  return density;
}

/* MaximaHandler::getSearchDistForClass (maxima_handler.cpp:509-521); m_radius = the value MaximaHandler::setRadius saw */
float search_dist_for_class(const pcdb_params& P, const std::vector<float>& dim_first,
                            const std::vector<float>& dim_second, unsigned class_id, float m_radius) {
  if (P.radius_type == PCDB_RADIUS_FIRST_DIM)
    return (class_id < dim_first.size() ? dim_first[class_id] : 0.f) * P.radius_factor;
  if (P.radius_type == PCDB_RADIUS_SECOND_DIM)
    return (class_id < dim_second.size() ? dim_second[class_id] : 0.f) * P.radius_factor;
  return m_radius;
}

/* single-object mode with a non-default max type (voting_mean_shift.cpp:124-155,161-176): no mean shift, one maximum at
 * `centre`, votes collected (and re-weighted) within h */
void ms_single_maximum(const pcdb_params& P, float h, ClassVotes& cv, const float* centre, std::vector<Vec3>& maxima,
                       std::vector<std::vector<int>>& members, std::vector<std::vector<float>>& member_w) {
  maxima.assign(1, Vec3{{centre[0], centre[1], centre[2]}});
  members.clear();
  member_w.clear();
  std::vector<int> mem;
  ms_density(P, h, cv, centre, mem, true);
  members.push_back(mem);
  std::vector<float> w;
  for (int i : mem) w.push_back(cv.w[i]);
  member_w.push_back(w);
}

void ms_find_maxima(const pcdb_params& P, float h, ClassVotes& cv, std::vector<Vec3>& maxima,
                    std::vector<std::vector<int>>& members, std::vector<std::vector<float>>& member_w) {
  const float r2 = radius_sq(double(h));
  const int n = int(cv.w.size());
  /* createSeeds :431-481 */
  const float binSize = (h * 2.0f) / sqrtf(2);
  struct Key {
    int x, y, z;
    bool operator<(const Key& o) const {
      if (z < o.z) return true;
      if ((z == o.z) && (y < o.y)) return true;
      if ((z == o.z) && (y == o.y) && (x < o.x)) return true;
      return false;
    }
  };
  std::vector<Vec3> seeds;
  if (binSize == 0) {
    for (int i = 0; i < n; ++i) seeds.push_back(Vec3{{cv.pos[3 * i], cv.pos[3 * i + 1], cv.pos[3 * i + 2]}});
  } else {
    std::map<Key, int> bins;
    for (int i = 0; i < n; ++i) {
      Key k{(int)std::floor((cv.pos[3 * i] / binSize) + 0.5), (int)std::floor((cv.pos[3 * i + 1] / binSize) + 0.5),
            (int)std::floor((cv.pos[3 * i + 2] / binSize) + 0.5)};
      bins[k]++;
    }
    for (auto& kv : bins) seeds.push_back(Vec3{{kv.first.x * binSize, kv.first.y * binSize, kv.first.z * binSize}});
  }
  /* iDoMeanShift :201-244 / computeMeanShift :331-376 */
  std::vector<Vec3> centers;
  std::vector<Nbr> nb;
  for (const Vec3& seed : seeds) {
    float cur[3] = {seed.v[0], seed.v[1], seed.v[2]};
    int iter = 0;
    float diff = 0;
    bool skip = false;
    do {
      ms_radius(cv, cur, r2, nb);
      if (nb.empty()) {
        skip = true;
        break;
      }
      float shifted[3] = {0, 0, 0};
      double totalWeight = 0;
      for (const Nbr& q : nb) {
        float u = q.d2 / (h * h);
        float g = -ms_kernel_derivative(P.ms_kernel, u) * cv.w[q.idx];
        for (int a = 0; a < 3; ++a) shifted[a] += g * cv.pos[3 * q.idx + a];
        totalWeight += g;
      }
      if (totalWeight != 0) {
        float tw = static_cast<float>(totalWeight); /* Eigen Vector3f /= takes the scalar as float */
        for (int a = 0; a < 3; ++a) shifted[a] /= tw;
      }
      diff = norm3(cur, shifted);
      for (int a = 0; a < 3; ++a) cur[a] = shifted[a];
      iter++;
    } while (diff > P.ms_threshold && iter <= P.ms_max_iter);
    if (!skip) centers.push_back(Vec3{{cur[0], cur[1], cur[2]}});
  }
  /* densities :90-97 */
  std::vector<float> dens;
  std::vector<int> mem;
  for (const Vec3& c : centers) dens.push_back(ms_density(P, h, cv, c.v, mem, false));
  if (P.maxima_suppression == PCDB_SUPPRESS_AVERAGE) { /* maxima_handler.cpp:94-157 */
    const int M = int(centers.size());
    std::vector<std::vector<int>> dup(M);
    for (int i = 0; i < M; ++i) dup[i].push_back(i);
    std::vector<bool> isdup(M, false);
    for (int k = 0; k < M; ++k) {
      if (isdup[k]) continue;
      for (int j = k + 1; j < M; ++j) {
        if (isdup[j]) continue;
        if (norm3(centers[k].v, centers[j].v) < h) {
          isdup[j] = true;
          dup[k].push_back(j);
        }
      }
    }
    std::vector<Vec3> avg;
    for (int i = 0; i < M; ++i) {
      if (dup[i].size() == 1)
        avg.push_back(centers[dup[i][0]]);
      else {
        float a[3] = {0, 0, 0};
        float sd = 0;
        for (int j : dup[i]) {
          for (int t = 0; t < 3; ++t) a[t] += centers[j].v[t] * dens[j];
          sd += dens[j];
        }
        for (int t = 0; t < 3; ++t) a[t] /= sd;
        avg.push_back(Vec3{{a[0], a[1], a[2]}});
      }
    }
    dens.clear();
    for (const Vec3& c : avg) dens.push_back(ms_density(P, h, cv, c.v, mem, false));
    centers = avg;
  }
  /* suppressNeighborMaxima maxima_handler.cpp:51-92 */
  maxima.clear();
  {
    std::vector<float> work(dens);
    while (true) {
      auto it = std::max_element(work.begin(), work.end());
      float mx = -1;
      if (it != work.end()) mx = *it;
      if (mx != -1) {
        size_t mi = it - work.begin();
        Vec3 c = centers[mi];
        maxima.push_back(c);
        work[mi] = -1;
        for (size_t i = 0; i < centers.size(); ++i)
          if (norm3(c.v, centers[i].v) < h) work[i] = -1;
      } else
        break;
    }
  }
  /* estimateDensityAndReweightVotes per maximum, cumulative :161-176,:289-328 */
  members.clear();
  member_w.clear();
  for (const Vec3& mxp : maxima) {
    ms_density(P, h, cv, mxp.v, mem, true);
    members.push_back(mem);
    std::vector<float> w;
    for (int i : mem) w.push_back(cv.w[i]);
    member_w.push_back(w);
  }
}

/* symmetric 4x4 Jacobi for Utils::quatWeightedAverage (utils/utils.cpp:617-665; the reference uses
 * Eigen::EigenSolver<Matrix4f>, eigenvector sign is solver dependent -> canonical w >= 0 here) */
void quat_average(const std::vector<Quat>& qs, const std::vector<float>& ws, Quat& out) {
  float S[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      float value = 0;
      for (size_t k = 0; k < qs.size(); ++k) {
        const float qv[4] = {qs[k].a, qs[k].b, qs[k].c, qs[k].d};
        value += ws[k] * qv[i] * qv[j];
      }
      S[i][j] = value;
    }
  double A[4][4], V[4][4];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      A[i][j] = S[i][j];
      V[i][j] = i == j;
    }
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0;
    for (int p = 0; p < 4; ++p)
      for (int q = p + 1; q < 4; ++q) off += std::fabs(A[p][q]);
    if (off == 0) break;
    for (int p = 0; p < 3; ++p)
      for (int q = p + 1; q < 4; ++q) {
        double apq = A[p][q];
        if (apq == 0.0) continue;
        double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        if (!std::isfinite(theta)) t = 0.0;
        double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < 4; ++k) {
          double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < 4; ++k) {
          double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        A[p][q] = A[q][p] = 0.0;
        for (int k = 0; k < 4; ++k) {
          double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
  int best = 0;
  float maxEv = 0;
  for (int i = 0; i < 4; ++i)
    if (float(A[i][i]) > maxEv) {
      maxEv = float(A[i][i]);
      best = i;
    }
  double sgn = V[0][best] < 0 ? -1.0 : 1.0;
  out = Quat{float(sgn * V[0][best]), float(sgn * V[1][best]), float(sgn * V[2][best]), float(sgn * V[3][best])};
}

/* Voting::findMaxima (voting/voting.cpp:79-328) for one cloud */
/* ---- Voting::filterVotesWithRansac (voting/voting.cpp:356-433) ------------------------------------------------------
 * pcl::registration::CorrespondenceRejectorSampleConsensus over the member votes of one maximum: source = the votes'
 * training keypoints, target = their scene keypoints, model = rigid transform from 3 correspondences, RANSAC with
 * max 10000 iterations, probability 0.99 and PCL's adaptive stop (RandomSampleConsensus::computeModel), sample
 * goodness by SampleConsensusModelRegistration (pairwise source distances above the PCA-derived threshold, at most
 * 1000 draws per iteration).  PCL is not available here and its sample sequence comes from a Boost RNG with internal
 * shuffle state, so the SEQUENCE is defined by this restatement (PARITY UNPINNED for this stage): sample `it` is a
 * pure function of (key, it, attempt); model fit and residuals are evaluated in double (PCL: float, JacobiSVD).
 * Returns false when the maximum is to be dropped; `keep` marks the inlier votes otherwise. */
inline uint32_t ransac_hash(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  uint32_t h = 0x811C9DC5u;
  const uint32_t v[4] = {a, b, c, d};
  for (int i = 0; i < 4; ++i) {
    h ^= v[i];
    h *= 0x01000193u;
    h ^= h >> 15;
  }
  h ^= h >> 16;
  h *= 0x85EBCA6Bu;
  h ^= h >> 13;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}

/* cyclic Jacobi for a symmetric n x n matrix (n <= 4): eigenvalues on the diagonal of A, eigenvectors in the columns of V */
template <int N>
void jacobi_sym(double A[N][N], double V[N][N]) {
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) V[i][j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 64; ++sweep) {
    double off = 0;
    for (int p = 0; p < N; ++p)
      for (int q = p + 1; q < N; ++q) off += std::fabs(A[p][q]);
    if (off == 0.0) break;
    for (int p = 0; p < N - 1; ++p)
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p][q];
        if (apq == 0.0) continue;
        const double theta = (A[q][q] - A[p][p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        if (!std::isfinite(theta)) t = 0.0;
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < N; ++k) {
          const double akp = A[k][p], akq = A[k][q];
          A[k][p] = c * akp - s * akq;
          A[k][q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {
          const double apk = A[p][k], aqk = A[q][k];
          A[p][k] = c * apk - s * aqk;
          A[q][k] = s * apk + c * aqk;
        }
        A[p][q] = A[q][p] = 0.0;
        for (int k = 0; k < N; ++k) {
          const double vkp = V[k][p], vkq = V[k][q];
          V[k][p] = c * vkp - s * vkq;
          V[k][q] = s * vkp + c * vkq;
        }
      }
  }
}

/* rigid transform (R row-major, t) that maps the three source points onto the three target points in the least-squares
 * sense: Horn's closed form — the rotation is the eigenvector of the largest eigenvalue of the 4 x 4 matrix built from
 * the cross-covariance (equals the SVD solution PCL's TransformationEstimationSVD returns) */
void rigid_from_three(const double s[3][3], const double g[3][3], double R[9], double t[3]) {
  double cs[3], cg[3];
  for (int a = 0; a < 3; ++a) {
    cs[a] = ((s[0][a] + s[1][a]) + s[2][a]) / 3.0;
    cg[a] = ((g[0][a] + g[1][a]) + g[2][a]) / 3.0;
  }
  double M[3][3];
  for (int a = 0; a < 3; ++a)
    for (int b = 0; b < 3; ++b) {
      double acc = 0;
      for (int k = 0; k < 3; ++k) acc += (s[k][a] - cs[a]) * (g[k][b] - cg[b]);
      M[a][b] = acc;
    }
  double Nm[4][4] = {
      {(M[0][0] + M[1][1]) + M[2][2], M[1][2] - M[2][1], M[2][0] - M[0][2], M[0][1] - M[1][0]},
      {0, (M[0][0] - M[1][1]) - M[2][2], M[0][1] + M[1][0], M[2][0] + M[0][2]},
      {0, 0, (M[1][1] - M[0][0]) - M[2][2], M[1][2] + M[2][1]},
      {0, 0, 0, (M[2][2] - M[0][0]) - M[1][1]}};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < i; ++j) Nm[i][j] = Nm[j][i];
  double V[4][4];
  jacobi_sym<4>(Nm, V);
  int best = 0;
  for (int i = 1; i < 4; ++i)
    if (Nm[i][i] > Nm[best][best]) best = i;
  double q[4] = {V[0][best], V[1][best], V[2][best], V[3][best]};
  const double qn = std::sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= qn;
  const double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = 1.0 - 2.0 * (y * y + z * z);
  R[1] = 2.0 * (x * y - w * z);
  R[2] = 2.0 * (x * z + w * y);
  R[3] = 2.0 * (x * y + w * z);
  R[4] = 1.0 - 2.0 * (x * x + z * z);
  R[5] = 2.0 * (y * z - w * x);
  R[6] = 2.0 * (x * z - w * y);
  R[7] = 2.0 * (y * z + w * x);
  R[8] = 1.0 - 2.0 * (x * x + y * y);
  for (int a = 0; a < 3; ++a) t[a] = cg[a] - ((R[3 * a] * cs[0] + R[3 * a + 1] * cs[1]) + R[3 * a + 2] * cs[2]);
}

inline double ransac_residual2(const double R[9], const double t[3], const float* s, const float* g) {
  double d2 = 0;
  for (int a = 0; a < 3; ++a) {
    const double p = ((R[3 * a] * (double)s[0] + R[3 * a + 1] * (double)s[1]) + R[3 * a + 2] * (double)s[2]) + t[a];
    const double d = p - (double)g[a];
    d2 += d * d;
  }
  return d2;
}

bool ransac_filter_votes(const pcdb_vote* votes, const std::vector<int64_t>& idx, float inlier_threshold, uint32_t key,
                         std::vector<char>& keep) {
  const int n = (int)idx.size();
  keep.assign((size_t)n, 0);
  if (n < 3) return false;
  /* SampleConsensusModelRegistration::computeSampleDistanceThreshold: ((sum of sqrt eigenvalues of the source
   * covariance) / 3)^2 */
  /* nine moment sums as 256 strided partial sums combined in order (the summation tree of the device kernel) */
  double tot[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int t = 0; t < 256; ++t) {
    double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = t; i < n; i += 256) {
      const float* s = votes[idx[i]].keypoint_training;
      const double x = s[0], y = s[1], z = s[2];
      acc[0] += x; acc[1] += y; acc[2] += z;
      acc[3] += x * x; acc[4] += x * y; acc[5] += x * z; acc[6] += y * y; acc[7] += y * z; acc[8] += z * z;
    }
    for (int j = 0; j < 9; ++j) tot[j] += acc[j];
  }
  const double inv_n = 1.0 / (double)n;
  for (int j = 0; j < 9; ++j) tot[j] *= inv_n;
  double C[3][3] = {{tot[3] - tot[0] * tot[0], tot[4] - tot[0] * tot[1], tot[5] - tot[0] * tot[2]},
                    {0, tot[6] - tot[1] * tot[1], tot[7] - tot[1] * tot[2]},
                    {0, 0, tot[8] - tot[2] * tot[2]}};
  C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
  double V3[3][3];
  jacobi_sym<3>(C, V3);
  double sdt = 0;
  for (int a = 0; a < 3; ++a) sdt += std::sqrt(C[a][a] > 0 ? C[a][a] : 0.0);
  sdt /= 3.0;
  sdt *= sdt;
  const double thr2 = (double)inlier_threshold * (double)inlier_threshold;
  const double kLogP = -4.605170185988091;   /* log(1 - 0.99) */
  const double kEps = 2.220446049250313e-16;
  double k = 1e300;
  int best = -1, bs[3] = {0, 0, 0};
  for (int it = 0; (double)it < k;) {
    int smp[3] = {-1, -1, -1};
    for (uint32_t at = 0; at < 1000; ++at) { /* max_sample_checks_ */
      int i0 = (int)(ransac_hash(key, (uint32_t)it, at, 0) % (uint32_t)n);
      int i1 = (int)(ransac_hash(key, (uint32_t)it, at, 1) % (uint32_t)(n - 1));
      if (i1 >= i0) ++i1;
      int i2 = (int)(ransac_hash(key, (uint32_t)it, at, 2) % (uint32_t)(n - 2));
      const int lo = i0 < i1 ? i0 : i1, hi = i0 < i1 ? i1 : i0;
      if (i2 >= lo) ++i2;
      if (i2 >= hi) ++i2;
      const float* a = votes[idx[i0]].keypoint_training;
      const float* b = votes[idx[i1]].keypoint_training;
      const float* c = votes[idx[i2]].keypoint_training;
      auto d2 = [](const float* p, const float* q) {
        const double dx = (double)p[0] - (double)q[0], dy = (double)p[1] - (double)q[1], dz = (double)p[2] - (double)q[2];
        return (dx * dx + dy * dy) + dz * dz;
      };
      if (d2(b, a) > sdt && d2(c, a) > sdt && d2(c, b) > sdt) {
        smp[0] = i0; smp[1] = i1; smp[2] = i2;
        break;
      }
    }
    if (smp[0] < 0) break; /* no good sample: RandomSampleConsensus stops */
    double S[3][3], G[3][3], R[9], t[3];
    for (int j = 0; j < 3; ++j)
      for (int a = 0; a < 3; ++a) {
        S[j][a] = votes[idx[smp[j]]].keypoint_training[a];
        G[j][a] = votes[idx[smp[j]]].keypoint[a];
      }
    rigid_from_three(S, G, R, t);
    int cnt = 0;
    for (int i = 0; i < n; ++i)
      if (ransac_residual2(R, t, votes[idx[i]].keypoint_training, votes[idx[i]].keypoint) < thr2) ++cnt;
    if (cnt > best) {
      best = cnt;
      bs[0] = smp[0]; bs[1] = smp[1]; bs[2] = smp[2];
      const double w = (double)best * inv_n;
      double p_no = 1.0 - (w * w) * w;
      p_no = p_no > kEps ? p_no : kEps;
      p_no = p_no < 1.0 - kEps ? p_no : 1.0 - kEps;
      k = kLogP / std::log(p_no);
    }
    ++it;
    if (it > 10000) break;
  }
  if (best < 3) return false; /* no model, or fewer than 3 inliers: best_transformation_ stays the identity */
  double S[3][3], G[3][3], R[9], t[3];
  for (int j = 0; j < 3; ++j)
    for (int a = 0; a < 3; ++a) {
      S[j][a] = votes[idx[bs[j]]].keypoint_training[a];
      G[j][a] = votes[idx[bs[j]]].keypoint[a];
    }
  rigid_from_three(S, G, R, t);
  /* Eigen::Matrix4f::isIdentity(1e-4): diagonal isApprox(1), everything else isMuchSmallerThan(1) */
  bool identity = true;
  for (int a = 0; a < 3 && identity; ++a) {
    for (int b = 0; b < 3; ++b) {
      const float m = (float)R[3 * a + b];
      if (a == b) {
        const float am = std::fabs(m);
        if (!(std::fabs(m - 1.0f) <= (am < 1.0f ? am : 1.0f) * 1e-4f)) identity = false;
      } else if (!(std::fabs(m) <= 1e-4f)) {
        identity = false;
      }
    }
    if (!(std::fabs((float)t[a]) <= 1e-4f)) identity = false;
  }
  if (identity) return false;
  for (int i = 0; i < n; ++i)
    keep[(size_t)i] = ransac_residual2(R, t, votes[idx[i]].keypoint_training, votes[idx[i]].keypoint) < thr2 ? 1 : 0;
  return true;
}

struct ClassDims {
  const std::vector<float>* first;
  const std::vector<float>* second;
};

/* surf / n_surf: the cloud handed to Voting::findMaxima (pointsWithoutNaN, implicit_shape_model.cpp:689) — only the
 * single-object max types read it (centroid, model radius); may be null otherwise. */
void find_maxima_cloud(const pcdb_params& P, const ClassDims& dims, const float* surf, int64_t n_surf,
                       const pcdb_vote* votes, int64_t nv, std::vector<pcdb_maximum>& out,
                       std::vector<int64_t>& member_idx, std::vector<float>& member_w) {
  out.clear();
  member_idx.clear();
  member_w.clear();
  std::map<unsigned, ClassVotes> by_class;
  for (int64_t i = 0; i < nv; ++i) {
    ClassVotes& cv = by_class[votes[i].class_id];
    cv.gidx.push_back(i);
    cv.pos.insert(cv.pos.end(), votes[i].position, votes[i].position + 3);
    cv.w.push_back(votes[i].weight);
  }
  struct Tmp {
    pcdb_maximum m;
    std::vector<int64_t> idx;
    std::vector<float> w;
  };
  std::vector<Tmp> all;
  static const std::vector<float> kNoDims;
  const std::vector<float>& d1 = dims.first ? *dims.first : kNoDims;
  const std::vector<float>& d2 = dims.second ? *dims.second : kNoDims;
  /* VotingMeanShift::m_bandwidth / MaximaHandler::m_radius as iFindMaxima updates them class after class
   * (voting_mean_shift.cpp:47-49).  The reference keeps m_bandwidth across detect() calls; here every cloud starts from
   * the configured Voting.Bandwidth (what a freshly loaded model does). */
  float m_bandwidth = P.bandwidth, m_radius = P.bandwidth;
  const bool single_max = P.single_object_mode && P.single_object_max_type != PCDB_SOMAX_DEFAULT;
  float centre[3] = {0, 0, 0};
  if (single_max) { /* pcl::compute3DCentroid into an Eigen::Vector4f: sequential float sums, one division */
    for (int64_t i = 0; i < n_surf; ++i)
      for (int a = 0; a < 3; ++a) centre[a] += surf[3 * i + a];
    for (int a = 0; a < 3; ++a) centre[a] /= static_cast<float>(n_surf);
  }
  for (auto& kv : by_class) {
    ClassVotes& cv = kv.second;
    std::vector<Vec3> maxima;
    std::vector<std::vector<int>> members;
    std::vector<std::vector<float>> mw;
    m_radius = m_bandwidth;
    m_bandwidth = search_dist_for_class(P, d1, d2, kv.first, m_radius);
    if (!single_max) {
      ms_find_maxima(P, m_bandwidth, cv, maxima, members, mw);
    } else {
      if (P.single_object_max_type == PCDB_SOMAX_BANDWIDTH)
        m_bandwidth = search_dist_for_class(P, d1, d2, kv.first, m_radius);
      if (P.single_object_max_type == PCDB_SOMAX_MODEL_RADIUS) { /* SingleObjectHelper::getModelRadius */
        float r = 0;
        for (int64_t i = 0; i < n_surf; ++i) {
          float d = norm3(surf + 3 * i, centre);
          if (d > r) r = d;
        }
        m_bandwidth = r;
      }
      if (P.single_object_max_type == PCDB_SOMAX_VOTING_SPACE) { /* SingleObjectHelper::getVotingSpaceSize */
        float mx = 0;
        for (size_t i = 0; i < cv.w.size(); ++i) {
          float dx = cv.pos[3 * i] - centre[0], dy = cv.pos[3 * i + 1] - centre[1], dz = cv.pos[3 * i + 2] - centre[2];
          float d = dx * dx + dy * dy + dz * dz;
          mx = mx > d ? mx : d;
        }
        m_bandwidth = std::sqrt(mx);
      }
      ms_single_maximum(P, m_bandwidth, cv, centre, maxima, members, mw);
    }
    if (P.ransac_vote_filtering) { /* voting.cpp:110-127: before the maxima are reduced; dropped maxima vanish */
      float thr = P.ransac_inlier_threshold;
      if (P.ransac_threshold_type == PCDB_RANSAC_OBJECT_RADIUS && kv.first < d1.size()) thr *= d1[kv.first];
      if (P.ransac_threshold_type == PCDB_RANSAC_BBOX_MEDIAN && kv.first < d2.size()) thr *= d2[kv.first];
      for (size_t i = 0; i < maxima.size(); ++i) {
        std::vector<int>& mem = members[i];
        if ((int)mem.size() < P.min_votes_threshold || mem.empty()) { /* voting.cpp:373: not even tried */
          mem.clear();
          mw[i].clear();
          continue;
        }
        std::vector<int64_t> gi(mem.size());
        for (size_t t = 0; t < mem.size(); ++t) gi[t] = cv.gidx[mem[t]];
        std::vector<char> keep;
        const bool ok = ransac_filter_votes(votes, gi, thr, kv.first * 65536u + (uint32_t)i, keep);
        std::vector<int> m2;
        std::vector<float> w2;
        if (ok)
          for (size_t t = 0; t < mem.size(); ++t)
            if (keep[t]) {
              m2.push_back(mem[t]);
              w2.push_back(mw[i][t]);
            }
        mem.swap(m2);
        mw[i].swap(w2);
      }
    }
    for (size_t i = 0; i < maxima.size(); ++i) {
      const std::vector<int>& mem = members[i];
      if ((int)mem.size() < P.min_votes_threshold || mem.empty()) continue;
      std::map<unsigned, float> inst_w;
      for (size_t t = 0; t < mem.size(); ++t) {
        unsigned inst = votes[cv.gidx[mem[t]]].instance_id;
        auto it = inst_w.find(inst);
        if (it != inst_w.end())
          it->second += mw[i][t];
        else
          inst_w.insert({inst, mw[i][t]});
      }
      unsigned max_id = 0; /* uninitialised in the reference when no weight > 0 */
      float best_weight = 0;
      for (auto& iw : inst_w)
        if (iw.second > best_weight) {
          best_weight = iw.second;
          max_id = iw.first;
        }
      Tmp t;
      std::memset(&t.m, 0, sizeof(t.m));
      t.m.class_id = kv.first;
      t.m.instance_id = max_id;
      t.m.instance_weight = inst_w[max_id];
      for (int a = 0; a < 3; ++a) t.m.position[a] = maxima[i].v[a];
      std::vector<Quat> quats;
      std::vector<float> weights;
      float maxWeight = 0;
      float size[3] = {0, 0, 0};
      for (size_t j = 0; j < mem.size(); ++j) {
        const pcdb_vote& v = votes[cv.gidx[mem[j]]];
        float nw = mw[i][j];
        quats.push_back(Quat{v.bbox_quat[0], v.bbox_quat[1], v.bbox_quat[2], v.bbox_quat[3]});
        weights.push_back(nw);
        for (int a = 0; a < 3; ++a) size[a] += nw * v.bbox_size[a];
        maxWeight += nw;
      }
      t.m.weight = maxWeight;
      t.m.raw_weight = maxWeight;
      for (float& w : weights) w /= maxWeight;
      for (int a = 0; a < 3; ++a) t.m.bbox_size[a] = size[a] / maxWeight;
      t.m.bbox_quat[0] = 1; /* VotingMaximum ctor: identity */
      if (P.average_rotation) {
        Quat q;
        quat_average(quats, weights, q);
        t.m.bbox_quat[0] = q.a;
        t.m.bbox_quat[1] = q.b;
        t.m.bbox_quat[2] = q.c;
        t.m.bbox_quat[3] = q.d;
      }
      t.m.n_votes = int(mem.size());
      for (size_t j = 0; j < mem.size(); ++j) {
        t.idx.push_back(cv.gidx[mem[j]]);
        t.w.push_back(mw[i][j]);
      }
      all.push_back(std::move(t));
    }
  }
  /* cross-class filterMaxima (voting.cpp:265-268) */
  if (!P.single_object_mode && P.max_filter_type == PCDB_MAXFILTER_SIMPLE) {
    /* MaximaHandler::suppressNeighborMaxima2 (maxima_handler.cpp:227-268) with radius = Voting.Bandwidth
     * (voting_mean_shift.cpp:48): repeatedly keep the heaviest pending maximum (std::max_element: the first of equal
     * ones) and drop every maximum closer than the radius, whatever its class */
    std::vector<float> work(all.size());
    for (size_t i = 0; i < all.size(); ++i) work[i] = all[i].m.weight;
    std::vector<Tmp> kept_max;
    for (;;) {
      int best = -1;
      for (size_t i = 0; i < all.size(); ++i)
        if (best < 0 || work[best] < work[i]) best = int(i);
      if (best < 0 || work[best] == -1.f) break;
      kept_max.push_back(all[best]);
      work[best] = -1.f;
      const float* c = all[best].m.position;
      for (size_t i = 0; i < all.size(); ++i) {
        const float* q = all[i].m.position;
        float dx = c[0] - q[0], dy = c[1] - q[1], dz = c[2] - q[2];
        if (std::sqrt(dx * dx + dy * dy + dz * dz) < m_radius) work[i] = -1.f;
      }
    }
    all = std::move(kept_max);
  }
  if (!P.single_object_mode && P.max_filter_type == PCDB_MAXFILTER_MERGE) {
    /* MaximaHandler::mergeAndFilterMaxima(maxima, true) + mergeMaxima (maxima_handler.cpp:296-383,386-443): a maximum
     * subsumes later maxima closer than its class' search distance whose own search distance is not larger; the group
     * (neighbours first, the maximum itself last) is merged per class (std::map: ascending class id) by running
     * weighted averages, and the heaviest merged maximum survives. */
    auto dist_for = [&](unsigned cls) { return search_dist_for_class(P, d1, d2, cls, m_radius); };
    std::vector<Tmp> filtered;
    std::vector<bool> dirty(all.size(), false);
    for (size_t i = 0; i < all.size(); ++i) {
      if (dirty[i]) continue;
      const float sd = dist_for(all[i].m.class_id);
      std::vector<size_t> close;
      for (size_t j = i + 1; j < all.size(); ++j) {
        if (dirty[j]) continue;
        const float dist = norm3(all[j].m.position, all[i].m.position);
        const float od = dist_for(all[j].m.class_id);
        if (dist < sd && od <= sd) {
          close.push_back(j);
          dirty[j] = true;
        }
      }
      if (close.empty()) {
        filtered.push_back(all[i]);
        continue;
      }
      close.push_back(i);
      std::map<unsigned, std::vector<size_t>> same;
      for (size_t c : close) same[all[c].m.class_id].push_back(c);
      Tmp best;
      std::memset(&best.m, 0, sizeof(best.m)); /* VotingMaximum(): weight 0, class -1, instance max, identity quat */
      best.m.class_id = 0xffffffffu;
      best.m.instance_id = 0xffffffffu;
      best.m.bbox_quat[0] = 1;
      for (auto& kv : same) {
        Tmp r;
        std::memset(&r.m, 0, sizeof(r.m));
        r.m.bbox_quat[0] = 1;
        std::map<unsigned, float> inst_w;
        for (size_t c : kv.second) {
          const pcdb_maximum& m = all[c].m;
          const float rw = r.m.weight, mw_ = m.weight;
          for (int a = 0; a < 3; ++a) {
            r.m.position[a] = r.m.position[a] * rw + m.position[a] * mw_;
            r.m.position[a] /= (rw + mw_);
            r.m.bbox_size[a] = r.m.bbox_size[a] * rw + m.bbox_size[a] * mw_;
            r.m.bbox_size[a] /= (rw + mw_);
          }
          Quat q;
          quat_average({Quat{r.m.bbox_quat[0], r.m.bbox_quat[1], r.m.bbox_quat[2], r.m.bbox_quat[3]},
                        Quat{m.bbox_quat[0], m.bbox_quat[1], m.bbox_quat[2], m.bbox_quat[3]}},
                       {rw, mw_}, q);
          r.m.bbox_quat[0] = q.a;
          r.m.bbox_quat[1] = q.b;
          r.m.bbox_quat[2] = q.c;
          r.m.bbox_quat[3] = q.d;
          r.m.class_id = m.class_id;
          r.m.weight += m.weight;
          r.idx.insert(r.idx.end(), all[c].idx.begin(), all[c].idx.end());
          r.w.insert(r.w.end(), all[c].w.begin(), all[c].w.end());
          auto it = inst_w.find(m.instance_id);
          if (it != inst_w.end())
            it->second += m.instance_weight;
          else
            inst_w.insert({m.instance_id, m.instance_weight});
          unsigned max_id = 0; /* uninitialised in the reference when no weight > 0 */
          float best_weight = 0;
          for (auto& iw : inst_w)
            if (iw.second > best_weight) {
              best_weight = iw.second;
              max_id = iw.first;
            }
          r.m.instance_id = max_id;
          r.m.instance_weight = inst_w[max_id];
        }
        r.m.raw_weight = r.m.weight;
        r.m.n_votes = int(r.idx.size());
        if (r.m.weight > best.m.weight) best = r;
      }
      filtered.push_back(best);
    }
    all = std::move(filtered);
  }
  std::stable_sort(all.begin(), all.end(), [](const Tmp& a, const Tmp& b) { return a.m.weight > b.m.weight; });
  float sum = 0, sum_inst = 0; /* normalizeWeights :441-462 */
  for (const Tmp& t : all) {
    sum += t.m.weight;
    sum_inst += t.m.instance_weight;
  }
  for (Tmp& t : all) {
    t.m.weight = sum != 0 ? t.m.weight / sum : 0;
    t.m.instance_weight = sum_inst != 0 ? t.m.instance_weight / sum_inst : 0;
  }
  float thr = P.min_threshold;
  if (thr < 0) {
    float mw = all.size() > 0 ? all.front().m.weight : 0.0f;
    thr = -thr * mw;
  }
  std::vector<Tmp> kept;
  for (Tmp& t : all)
    if (t.m.weight >= thr) kept.push_back(std::move(t));
  if (P.best_k > 0 && (int)kept.size() >= P.best_k) kept.resize(P.best_k);
  for (Tmp& t : kept) {
    t.m.vote_begin = (int64_t)member_idx.size();
    member_idx.insert(member_idx.end(), t.idx.begin(), t.idx.end());
    member_w.insert(member_w.end(), t.w.begin(), t.w.end());
    out.push_back(t.m);
  }
}

/* ------------------------------------------------------------------------------------------- */
/* per-cloud feature extraction: Features::operator() + removeNaNFeatures                        */
/* (features/features.cpp:40-116, implicit_shape_model.cpp:733-927,1276-1308)                    */
/* ------------------------------------------------------------------------------------------- */
struct CloudFeatures {
  std::vector<float> xyz, lrf, desc;
  std::vector<float> surf; /* pointsWithoutNaN: finite points with finite normals (what Voting::findMaxima is handed) */
  int64_t n_kp = 0, n_lrf_nb = 0, n_shot_nb = 0;
};

void compute_features_cloud(const pcdb_params& P, const float* xyz, const float* normals, const uint32_t* rgb,
                            int n, CloudFeatures& out, double* t_keypoints_ms) {
  out.xyz.clear();
  out.lrf.clear();
  out.desc.clear();
  const bool color = P.feature_type == PCDB_FEATURE_CSHOT;
  const int D = color ? PCDB_CSHOT_DIM : PCDB_SHOT_DIM;
  /* removeNaNFromPointCloud (implicit_shape_model.cpp:611), filterNormals (:1040-1068) */
  std::vector<float> pts, sxyz, snrm, est;
  std::vector<uint32_t> prgb, srgb;
  if (!normals) { /* hasNormals == false: computeNormals on the NaN-free cloud (implicit_shape_model.cpp:852-858) */
    for (int i = 0; i < n; ++i)
      if (finite3(xyz + 3 * i)) pts.insert(pts.end(), xyz + 3 * i, xyz + 3 * i + 3);
    est.resize(pts.size());
    compute_normals_cloud(P, pts.data(), int(pts.size() / 3), est.data(), nullptr);
    pts.clear();
  }
  int fi = 0;
  for (int i = 0; i < n; ++i) {
    if (!finite3(xyz + 3 * i)) continue;
    const float* nrm_i = normals ? normals + 3 * i : &est[3 * size_t(fi)];
    ++fi;
    pts.insert(pts.end(), xyz + 3 * i, xyz + 3 * i + 3);
    prgb.push_back(rgb ? rgb[i] : 0u);
    if (!finite3(nrm_i)) continue;
    sxyz.insert(sxyz.end(), xyz + 3 * i, xyz + 3 * i + 3);
    snrm.insert(snrm.end(), nrm_i, nrm_i + 3);
    srgb.push_back(rgb ? rgb[i] : 0u);
  }
  auto t0 = std::chrono::steady_clock::now();
  Keypoints kps;
  voxel_keypoints(pts.data(), prgb.data(), int(prgb.size()), P.leaf_size, kps);
  if (t_keypoints_ms)
    *t_keypoints_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  const int Q = int(kps.rgb.size());
  const int ns = int(srgb.size());
  out.n_kp = Q;
  out.surf = sxyz;
  CloudGrid g_lrf, g_shot;
  g_lrf.build(sxyz.data(), ns, P.lrf_radius);
  g_shot.build(sxyz.data(), ns, P.feature_radius);
  std::vector<float> lrf(size_t(Q) * 9), desc(size_t(Q) * D);
  int64_t n1 = 0, n2 = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : n1, n2)
  for (int q = 0; q < Q; ++q) {
    std::vector<Nbr> nb;
    g_lrf.query(&kps.xyz[3 * q], radius_sq(P.lrf_radius), nb);
    n1 += (int64_t)nb.size();
    shot_lrf(sxyz.data(), nb, &kps.xyz[3 * q], P.lrf_radius, &lrf[9 * q]);
    bool ok = std::isfinite(lrf[9 * q]) && std::isfinite(lrf[9 * q + 3]) && std::isfinite(lrf[9 * q + 6]);
    if (!ok) {
      desc[size_t(q) * D] = kNaNf;
      continue;
    }
    g_shot.query(&kps.xyz[3 * q], radius_sq(P.feature_radius), nb);
    n2 += (int64_t)nb.size();
    shot_describe(color, sxyz.data(), snrm.data(), srgb.data(), nb, &kps.xyz[3 * q], kps.rgb[q], &lrf[9 * q],
                  P.feature_radius, &desc[size_t(q) * D]);
  }
  out.n_lrf_nb = n1;
  out.n_shot_nb = n2;
  for (int q = 0; q < Q; ++q) {
    bool ok = std::isfinite(lrf[9 * q]) && std::isfinite(lrf[9 * q + 3]) && std::isfinite(lrf[9 * q + 6]);
    if (!ok) continue;
    bool nan = false;
    for (int j = 0; j < D && !nan; ++j) nan = std::isnan(desc[size_t(q) * D + j]);
    if (nan) continue;
    out.xyz.insert(out.xyz.end(), &kps.xyz[3 * q], &kps.xyz[3 * q] + 3);
    out.lrf.insert(out.lrf.end(), &lrf[9 * q], &lrf[9 * q] + 9);
    out.desc.insert(out.desc.end(), &desc[size_t(q) * D], &desc[size_t(q) * D] + D);
  }
}

/* ActivationStrategyKNN::activateKNN for a block of queries (activation_strategy_knn.h:41-126) */
void activate_block(const Model& m, const float* queries, int64_t Q, int k, int dist_type, bool use_ratio,
                    float ratio_thr, int32_t* idx_out, float* dist_out, int32_t* count_out) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int64_t q = 0; q < Q; ++q) {
    const float* qv = queries + q * m.D;
    int32_t* io = idx_out + q * k;
    float* dout = dist_out + q * k;
    for (int j = 0; j < k; ++j) {
      io[j] = -1;
      dout[j] = kNaNf;
    }
    if (m.N <= k) { /* :50-54 */
      for (int64_t j = 0; j < m.N; ++j) {
        io[j] = int(j);
        dout[j] = dist_fn(dist_type, m.words.data() + j * m.D, qv, m.D);
      }
      count_out[q] = int(m.N);
      continue;
    }
    int kk = use_ratio ? k + 1 : k;
    Cand best[PCDB_MAX_K + 1];
    int found = 0;
    if (m.forest && dist_type == PCDB_DIST_EUCLIDEAN) {
      static thread_local std::vector<unsigned char> checked;
      static thread_local std::vector<int> touched;
      if ((int64_t)checked.size() != m.N) checked.assign(m.N, 0);
      m.forest->knn(qv, kk, best, found, checked, touched);
    } else
      knn_one(qv, m.words.data(), m.N, m.D, kk, dist_type, best, found);
    int use = std::min(found, k);
    if (use_ratio && k == 1 && found >= 2) {
      if (best[0].d / best[1].d > ratio_thr) use = 0; /* :75-85 */
    }
    for (int j = 0; j < use; ++j) {
      io[j] = best[j].idx;
      dout[j] = best[j].d;
    }
    count_out[q] = use;
  }
}

std::string g_err;

/* trained model arrays kept between orc_train and orc_train_fetch */
struct Trained {
  int64_t N = 0, V = 0;
  int D = 0;
  std::vector<float> words, vote_xyz, vote_weight, vote_bbox, vote_class_weight, kp_train, sigma2;
  std::vector<int64_t> vote_off;
  std::vector<uint32_t> vote_class, vote_instance;
  std::vector<int32_t> ids;
} g_trained;

}  // namespace

/* ============================================================================================ */
/*                                         C interface                                           */
/* ============================================================================================ */
extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }
int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

void orc_default_params(pcdb_params* p) {
  std::memset(p, 0, sizeof(*p));
  p->feature_type = PCDB_FEATURE_SHOT;
  p->feature_radius = 0.1;
  p->lrf_radius = double(0.2f);
  p->leaf_size = 0.1f;
  p->distance_type = PCDB_DIST_EUCLIDEAN;
  p->knn_k = 1;
  p->distance_ratio_threshold = 0.95f;
  p->bandwidth = 0.2f;
  p->ms_threshold = 1e-3f;
  p->ms_max_iter = 1000;
  p->ms_kernel = PCDB_KERNEL_GAUSSIAN;
  p->maxima_suppression = PCDB_SUPPRESS_AVERAGE;
  p->min_votes_threshold = 1;
  p->best_k = -1;
  p->normal_radius = 0.05f;
  p->consistent_normals_method = 2;
  p->max_filter_type = PCDB_MAXFILTER_NONE;
  p->radius_type = PCDB_RADIUS_CONFIG;
  p->radius_factor = 1.0f;
  p->single_object_max_type = PCDB_SOMAX_DEFAULT;
  p->ransac_vote_filtering = 0;
  p->ransac_inlier_threshold = 0.1f;
  p->ransac_threshold_type = PCDB_RANSAC_FIXED;
  p->ransac_refine_model = 0;
}

int orc_voxel_keypoints(const float* xyz, const uint32_t* rgb, const int64_t* cloud_off, int32_t B, float leaf,
                        float* kp_xyz_out, uint32_t* kp_rgb_out, int64_t* kp_off_out, int64_t kp_capacity) {
  int64_t total = 0;
  kp_off_out[0] = 0;
  for (int b = 0; b < B; ++b) {
    Keypoints k;
    int64_t s = cloud_off[b];
    voxel_keypoints(xyz + 3 * s, rgb ? rgb + s : nullptr, int(cloud_off[b + 1] - s), leaf, k);
    int64_t q = (int64_t)k.rgb.size();
    if (total + q > kp_capacity) {
      g_err = "kp_capacity too small";
      return PCDB_E_CAPACITY;
    }
    std::memcpy(kp_xyz_out + 3 * total, k.xyz.data(), sizeof(float) * 3 * q);
    if (kp_rgb_out) std::memcpy(kp_rgb_out + total, k.rgb.data(), sizeof(uint32_t) * q);
    total += q;
    kp_off_out[b + 1] = total;
  }
  return PCDB_OK;
}

int orc_radius_neighbours(const float* surf_xyz, const int64_t* surf_off, const float* kp_xyz, const int64_t* kp_off,
                          int32_t B, double radius, int64_t* nbr_off_out, int32_t* nbr_idx_out, float* nbr_d2_out,
                          int64_t capacity) {
  int64_t total = 0;
  nbr_off_out[0] = 0;
  for (int b = 0; b < B; ++b) {
    CloudGrid g;
    g.build(surf_xyz + 3 * surf_off[b], int(surf_off[b + 1] - surf_off[b]), radius);
    std::vector<Nbr> nb;
    for (int64_t q = kp_off[b]; q < kp_off[b + 1]; ++q) {
      g.query(kp_xyz + 3 * q, radius_sq(radius), nb);
      if (total + (int64_t)nb.size() > capacity) {
        g_err = "neighbour capacity too small";
        return PCDB_E_CAPACITY;
      }
      for (const Nbr& n : nb) {
        nbr_idx_out[total] = n.idx;
        nbr_d2_out[total] = n.d2;
        ++total;
      }
      nbr_off_out[q + 1] = total;
    }
  }
  return PCDB_OK;
}

int orc_shot_lrf(const float* surf_xyz, const int64_t* surf_off, const float* kp_xyz, const int64_t* kp_off,
                 int32_t B, double radius, float* lrf9_out) {
  for (int b = 0; b < B; ++b) {
    CloudGrid g;
    const float* surf = surf_xyz + 3 * surf_off[b];
    g.build(surf, int(surf_off[b + 1] - surf_off[b]), radius);
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t q = kp_off[b]; q < kp_off[b + 1]; ++q) {
      std::vector<Nbr> nb;
      g.query(kp_xyz + 3 * q, radius_sq(radius), nb);
      shot_lrf(surf, nb, kp_xyz + 3 * q, radius, lrf9_out + 9 * q);
    }
  }
  return PCDB_OK;
}

int orc_shot_describe(int32_t feature_type, const float* surf_xyz, const float* surf_normals,
                      const uint32_t* surf_rgb, const int64_t* surf_off, const float* kp_xyz, const uint32_t* kp_rgb,
                      const float* kp_lrf9, const int64_t* kp_off, int32_t B, double radius, float* desc_out) {
  const bool color = feature_type == PCDB_FEATURE_CSHOT;
  const int D = color ? PCDB_CSHOT_DIM : PCDB_SHOT_DIM;
  for (int b = 0; b < B; ++b) {
    CloudGrid g;
    const int64_t s = surf_off[b];
    g.build(surf_xyz + 3 * s, int(surf_off[b + 1] - s), radius);
#pragma omp parallel for schedule(dynamic, 8)
    for (int64_t q = kp_off[b]; q < kp_off[b + 1]; ++q) {
      std::vector<Nbr> nb;
      g.query(kp_xyz + 3 * q, radius_sq(radius), nb);
      shot_describe(color, surf_xyz + 3 * s, surf_normals + 3 * s, surf_rgb ? surf_rgb + s : nullptr, nb,
                    kp_xyz + 3 * q, kp_rgb ? kp_rgb[q] : 0u, kp_lrf9 + 9 * q, radius, desc_out + q * D);
    }
  }
  return PCDB_OK;
}

/* normals_out P x 3 aligned with the input (NaN for non-finite points); curvature_out P or NULL */
int orc_compute_normals(const pcdb_params* prm, const float* xyz, const int64_t* cloud_off, int32_t B,
                        float* normals_out, float* curvature_out) {
  for (int b = 0; b < B; ++b) {
    const int64_t s = cloud_off[b], e = cloud_off[b + 1];
    std::vector<float> pts, nrm, cv;
    std::vector<int64_t> where;
    for (int64_t i = s; i < e; ++i) {
      for (int a = 0; a < 3; ++a) normals_out[3 * i + a] = kNaNf;
      if (curvature_out) curvature_out[i] = kNaNf;
      if (!finite3(xyz + 3 * i)) continue;
      pts.insert(pts.end(), xyz + 3 * i, xyz + 3 * i + 3);
      where.push_back(i);
    }
    nrm.resize(pts.size());
    cv.resize(where.size());
    compute_normals_cloud(*prm, pts.data(), int(where.size()), nrm.data(), cv.data());
    for (size_t j = 0; j < where.size(); ++j) {
      std::memcpy(normals_out + 3 * where[j], &nrm[3 * j], sizeof(float) * 3);
      if (curvature_out) curvature_out[where[j]] = cv[j];
    }
  }
  return PCDB_OK;
}

/* ImplicitShapeModel::computeNormals, organized branch (implicit_shape_model.cpp:948-966):
 * pcl::IntegralImageNormalEstimation, AVERAGE_3D_GRADIENT, MaxDepthChangeFactor 0.02, NormalSmoothingSize 10, border
 * policy IGNORE, no depth-dependent smoothing, viewpoint = sensor origin (0,0,0).  Restated from PCL's published
 * algorithm (features/integral_image_normal.hpp, integral_image2D.hpp) — PARITY UNPINNED (no PCL here):
 *   1. 3D gradients: dx = right - left, dy = down - up for the interior pixels, 0 on the image border;
 *   2. integral images of both (fp64 sums, non-finite gradients skipped and counted out);
 *   3. depth-change map (|dz| to the right / lower neighbour > 0.02 (|z| + 1) 2, or a non-finite depth) -> chamfer
 *      distance transform (1 / 1.4 weights, two raster passes over the flat array exactly as PCL indexes it);
 *   4. per pixel at least 10 inside the image: s = min(distance, 10); s <= 2 -> NaN; else box sums of both gradients
 *      over an int(s) x int(s) rectangle, normal = gy x gx normalised, flipped towards the viewpoint.
 * xyz: height x width x 3 row-major; normals_out the same shape (NaN where PCL yields NaN). */
int orc_compute_normals_organized(const float* xyz, int32_t width, int32_t height, float* normals_out) {
  const int W = width, H = height;
  const size_t n = (size_t)W * H;
  for (size_t i = 0; i < 3 * n; ++i) normals_out[i] = kNaNf;
  if (W < 3 || H < 3) return PCDB_OK;
  std::vector<float> dx(3 * n, 0.f), dy(3 * n, 0.f);
  for (int r = 1; r < H - 1; ++r)
    for (int c = 1; c < W - 1; ++c) {
      const size_t i = (size_t)r * W + c;
      for (int a = 0; a < 3; ++a) {
        dx[3 * i + a] = xyz[3 * (i + 1) + a] - xyz[3 * (i - 1) + a];
        dy[3 * i + a] = xyz[3 * (i + W) + a] - xyz[3 * (i - W) + a];
      }
    }
  /* integral images, (W + 1) x (H + 1), IntegralImage2D<float, 3>::computeIntegralImages */
  const size_t IW = (size_t)W + 1;
  std::vector<double> Ix(3 * IW * (H + 1), 0.0), Iy(3 * IW * (H + 1), 0.0);
  std::vector<unsigned> Cx(IW * (H + 1), 0u), Cy(IW * (H + 1), 0u);
  auto integrate = [&](const std::vector<float>& d, std::vector<double>& I, std::vector<unsigned>& Cn) {
    for (int r = 0; r < H; ++r)
      for (int c = 0; c < W; ++c) {
        const size_t cur = (size_t)(r + 1) * IW + (c + 1), up = (size_t)r * IW + (c + 1), lf = (size_t)(r + 1) * IW + c,
                     ul = (size_t)r * IW + c;
        for (int a = 0; a < 3; ++a) I[3 * cur + a] = I[3 * up + a] + I[3 * lf + a] - I[3 * ul + a];
        Cn[cur] = Cn[up] + Cn[lf] - Cn[ul];
        const float* e = &d[3 * ((size_t)r * W + c)];
        if (std::isfinite(e[0] + e[1] + e[2])) {
          for (int a = 0; a < 3; ++a) I[3 * cur + a] += (double)e[a];
          ++Cn[cur];
        }
      }
  };
  integrate(dx, Ix, Cx);
  integrate(dy, Iy, Cy);
  /* depth-change map and chamfer distance transform */
  std::vector<unsigned char> change(n, 255);
  for (int r = 0; r < H - 1; ++r)
    for (int c = 0; c < W - 1; ++c) {
      const size_t i = (size_t)r * W + c;
      const float depth = xyz[3 * i + 2], depthR = xyz[3 * (i + 1) + 2], depthD = xyz[3 * (i + W) + 2];
      const float lim = (0.02f * (std::fabs(depth) + 1.0f) * 2.0f);
      if (std::fabs(depth - depthR) > lim || !std::isfinite(depth) || !std::isfinite(depthR)) {
        change[i] = 0;
        change[i + 1] = 0;
      }
      if (std::fabs(depth - depthD) > lim || !std::isfinite(depth) || !std::isfinite(depthD)) {
        change[i] = 0;
        change[i + W] = 0;
      }
    }
  std::vector<float> dist(n);
  for (size_t i = 0; i < n; ++i) dist[i] = change[i] == 0 ? 0.0f : static_cast<float>(W + H);
  for (int r = 1; r < H; ++r) {
    float* prev = &dist[(size_t)(r - 1) * W];
    float* cur = &dist[(size_t)r * W];
    for (int c = 1; c < W; ++c) {
      const float upLeft = prev[c - 1] + 1.4f, up = prev[c] + 1.0f, upRight = prev[c + 1] + 1.4f; /* prev[W] = cur[0] */
      const float left = cur[c - 1] + 1.0f, center = cur[c];
      const float mn = std::min(std::min(upLeft, up), std::min(left, upRight));
      if (mn < center) cur[c] = mn;
    }
  }
  for (int r = H - 2; r >= 0; --r) {
    float* next = &dist[(size_t)(r + 1) * W];
    float* cur = &dist[(size_t)r * W];
    for (int c = W - 2; c >= 0; --c) {
      const float lowerLeft = next[c - 1] + 1.4f, lower = next[c] + 1.0f, lowerRight = next[c + 1] + 1.4f; /* next[-1] = cur[W-1] */
      const float right = cur[c + 1] + 1.0f, center = cur[c];
      const float mn = std::min(std::min(lowerLeft, lower), std::min(right, lowerRight));
      if (mn < center) cur[c] = mn;
    }
  }
  const int border = 10; /* int(normal_smoothing_size_) */
  for (int r = border; r < H - border; ++r)
    for (int c = border; c < W - border; ++c) {
      const size_t i = (size_t)r * W + c;
      if (!std::isfinite(xyz[3 * i + 2])) continue;
      const float smoothing = std::min(dist[i], 10.0f);
      if (!(smoothing > 2.0f)) continue;
      const int rw = static_cast<int>(smoothing), rh = rw;
      const int sx = c - rw / 2, sy = r - rh / 2;
      const size_t ul = (size_t)sy * IW + sx, ur = ul + rw, ll = (size_t)(sy + rh) * IW + sx, lr = ll + rw;
      const unsigned cx = Cx[lr] + Cx[ul] - Cx[ur] - Cx[ll], cy = Cy[lr] + Cy[ul] - Cy[ur] - Cy[ll];
      if (cx == 0 || cy == 0) continue;
      double gx[3], gy[3];
      for (int a = 0; a < 3; ++a) {
        gx[a] = Ix[3 * lr + a] + Ix[3 * ul + a] - Ix[3 * ur + a] - Ix[3 * ll + a];
        gy[a] = Iy[3 * lr + a] + Iy[3 * ul + a] - Iy[3 * ur + a] - Iy[3 * ll + a];
      }
      double nv[3] = {gy[1] * gx[2] - gy[2] * gx[1], gy[2] * gx[0] - gy[0] * gx[2], gy[0] * gx[1] - gy[1] * gx[0]};
      const double len2 = nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2];
      if (len2 == 0.0) continue;
      const double len = std::sqrt(len2);
      float nx = static_cast<float>(nv[0] / len), ny = static_cast<float>(nv[1] / len), nz = static_cast<float>(nv[2] / len);
      /* pcl::flipNormalTowardsViewpoint with the viewpoint at the sensor origin */
      const float vx = 0.f - xyz[3 * i], vy = 0.f - xyz[3 * i + 1], vz = 0.f - xyz[3 * i + 2];
      if (vx * nx + vy * ny + vz * nz < 0) {
        nx *= -1;
        ny *= -1;
        nz *= -1;
      }
      normals_out[3 * i] = nx;
      normals_out[3 * i + 1] = ny;
      normals_out[3 * i + 2] = nz;
    }
  return PCDB_OK;
}

int orc_compute_features(const pcdb_params* prm, const float* xyz, const float* normals, const uint32_t* rgb,
                         const int64_t* cloud_off, int32_t B, float* feat_xyz_out, float* feat_lrf9_out,
                         float* feat_desc_out, int64_t* feat_off_out, int64_t feat_capacity) {
  const int D = prm->feature_type == PCDB_FEATURE_CSHOT ? PCDB_CSHOT_DIM : PCDB_SHOT_DIM;
  int64_t total = 0;
  feat_off_out[0] = 0;
  for (int b = 0; b < B; ++b) {
    CloudFeatures f;
    int64_t s = cloud_off[b];
    compute_features_cloud(*prm, xyz + 3 * s, normals ? normals + 3 * s : nullptr, rgb ? rgb + s : nullptr,
                           int(cloud_off[b + 1] - s), f, nullptr);
    int64_t q = (int64_t)f.xyz.size() / 3;
    if (total + q > feat_capacity) {
      g_err = "feat_capacity too small";
      return PCDB_E_CAPACITY;
    }
    std::memcpy(feat_xyz_out + 3 * total, f.xyz.data(), sizeof(float) * 3 * q);
    std::memcpy(feat_lrf9_out + 9 * total, f.lrf.data(), sizeof(float) * 9 * q);
    std::memcpy(feat_desc_out + D * total, f.desc.data(), sizeof(float) * size_t(D) * q);
    total += q;
    feat_off_out[b + 1] = total;
  }
  return PCDB_OK;
}

/* functor values for arbitrary pairs: rows a[i], b[i] */
int orc_distance(const float* a, const float* b, int64_t n, int32_t D, int32_t dist_type, float* out) {
  for (int64_t i = 0; i < n; ++i) out[i] = dist_fn(dist_type, a + i * D, b + i * D, D);
  return PCDB_OK;
}

/* RGB->Lab (normalised L/100, a/120, b/120) and colour distance, for the reference pin */
int orc_rgb_to_lab_normalized(const uint32_t* rgb, int64_t n, float* lab_out) {
  for (int64_t i = 0; i < n; ++i) {
    float L, a, b;
    rgb2lab(rgb[i], L, a, b);
    lab_out[3 * i] = L / 100.0f;
    lab_out[3 * i + 1] = a / 120.0f;
    lab_out[3 * i + 2] = b / 120.0f;
  }
  return PCDB_OK;
}
int orc_color_distance(const float* lab, const float* lab_ref, int64_t n, float* out) {
  for (int64_t i = 0; i < n; ++i)
    out[i] = float(color_distance(lab[3 * i], lab[3 * i + 1], lab[3 * i + 2], lab_ref[3 * i], lab_ref[3 * i + 1],
                                  lab_ref[3 * i + 2]));
  return PCDB_OK;
}
int orc_lab_luts(float* srgb256, float* sxyz4000) {
  std::memcpy(srgb256, lab_lut().srgb, sizeof(float) * 256);
  std::memcpy(sxyz4000, lab_lut().sxyz, sizeof(float) * 4000);
  return PCDB_OK;
}

/* ---- model ---------------------------------------------------------------------------------- */
void* orc_model_create(const pcdb_params* prm, const float* words, int64_t N, int32_t D, const int64_t* vote_off,
                       const float* vote_xyz, const float* vote_weight, const uint32_t* vote_class,
                       const uint32_t* vote_instance, const float* vote_bbox, const float* vote_class_weight,
                       const float* kp_train, const int32_t* codeword_ids, const float* codeword_weight,
                       const float* class_sigma2, int32_t n_classes) {
  Model* m = new Model();
  m->prm = *prm;
  m->N = N;
  m->D = D;
  m->words.assign(words, words + N * D);
  m->vote_off.assign(vote_off, vote_off + N + 1);
  int64_t V = vote_off[N];
  m->vote_xyz.assign(vote_xyz, vote_xyz + 3 * V);
  m->vote_weight.assign(vote_weight, vote_weight + V);
  m->vote_class.assign(vote_class, vote_class + V);
  m->vote_instance.assign(vote_instance, vote_instance + V);
  m->vote_bbox.assign(vote_bbox, vote_bbox + 7 * V);
  if (vote_class_weight) m->vote_class_weight.assign(vote_class_weight, vote_class_weight + V);
  m->kp_train.assign(kp_train, kp_train + 3 * N);
  if (codeword_ids) m->codeword_ids.assign(codeword_ids, codeword_ids + N);
  if (codeword_weight)
    m->codeword_weight.assign(codeword_weight, codeword_weight + N);
  else
    m->codeword_weight.assign(N, 1.0f);
  m->sigma2.assign(class_sigma2, class_sigma2 + n_classes);
  return m;
}
/* trees > 0: build the FLANN-like forest and use it for activation (approximate; cost stand-in); trees == 0: exact */
double orc_model_set_approximate(void* model, int32_t trees, int32_t checks, uint32_t seed) {
  Model* m = static_cast<Model*>(model);
  if (trees <= 0) {
    m->forest.reset();
    return 0.0;
  }
  auto t0 = std::chrono::steady_clock::now();
  auto f = std::make_shared<KdForest>();
  f->checks = checks;
  f->build(m->words.data(), m->N, m->D, trees, seed);
  m->forest = f;
  m->forest_build_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return m->forest_build_ms;
}
void orc_model_set_params(void* model, const pcdb_params* prm) { static_cast<Model*>(model)->prm = *prm; }
void orc_model_set_class_dimensions(void* model, const float* first_dim, const float* second_dim, int32_t n_classes) {
  Model* m = static_cast<Model*>(model);
  m->dim_first.assign(first_dim, first_dim + n_classes);
  m->dim_second.assign(second_dim, second_dim + n_classes);
}
void orc_model_destroy(void* model) { delete static_cast<Model*>(model); }

int orc_knn(void* model, const float* queries, int64_t Q, int32_t k, int32_t dist_type, int32_t /*mode*/,
            int32_t* idx_out, float* dist_out, int32_t* count_out) {
  const Model& m = *static_cast<Model*>(model);
  if (k < 1 || k > PCDB_MAX_K) {
    g_err = "k out of range";
    return PCDB_E_INVALID;
  }
  activate_block(m, queries, Q, k, dist_type, m.prm.use_distance_ratio != 0, m.prm.distance_ratio_threshold, idx_out,
                 dist_out, count_out);
  return PCDB_OK;
}

int orc_cast_votes(void* model, const float* feat_xyz, const float* feat_lrf9, const int64_t* feat_off, int32_t B,
                   const int32_t* knn_idx, const float* knn_dist, const int32_t* knn_count, int32_t k,
                   pcdb_vote* votes_out, int64_t* vote_off_out, int64_t vote_capacity) {
  const Model& m = *static_cast<Model*>(model);
  int64_t total = 0;
  vote_off_out[0] = 0;
  std::vector<pcdb_vote> tmp;
  for (int b = 0; b < B; ++b) {
    for (int64_t q = feat_off[b]; q < feat_off[b + 1]; ++q) {
      for (int j = 0; j < knn_count[q]; ++j) {
        tmp.clear();
        cast_votes_one(m, feat_xyz + 3 * q, feat_lrf9 + 9 * q, nullptr, knn_idx[q * k + j], knn_dist[q * k + j], tmp);
        if (total + (int64_t)tmp.size() > vote_capacity) {
          g_err = "vote_capacity too small";
          return PCDB_E_CAPACITY;
        }
        for (const pcdb_vote& v : tmp) votes_out[total++] = v;
      }
    }
    vote_off_out[b + 1] = total;
  }
  return PCDB_OK;
}

static std::vector<int64_t> g_member_idx;
static std::vector<float> g_member_w;

int orc_find_maxima(void* model, const pcdb_vote* votes, const int64_t* vote_off, int32_t B,
                    pcdb_maximum* maxima_out, int64_t* maxima_off_out, int64_t maxima_capacity) {
  const Model& m = *static_cast<Model*>(model);
  g_member_idx.clear();
  g_member_w.clear();
  int64_t total = 0;
  maxima_off_out[0] = 0;
  for (int b = 0; b < B; ++b) {
    std::vector<pcdb_maximum> mx;
    std::vector<int64_t> mi;
    std::vector<float> mw;
    if (m.prm.single_object_mode && m.prm.single_object_max_type != PCDB_SOMAX_DEFAULT) {
      g_err = "SingleObjectMaxType other than Default needs the cloud: use the fused classify entry";
      return PCDB_E_UNSUPPORTED;
    }
    find_maxima_cloud(m.prm, ClassDims{&m.dim_first, &m.dim_second}, nullptr, 0, votes + vote_off[b],
                      vote_off[b + 1] - vote_off[b], mx, mi, mw);
    if (total + (int64_t)mx.size() > maxima_capacity) {
      g_err = "maxima_capacity too small";
      return PCDB_E_CAPACITY;
    }
    for (pcdb_maximum& x : mx) {
      x.vote_begin += (int64_t)g_member_idx.size();
      maxima_out[total++] = x;
    }
    for (int64_t i : mi) g_member_idx.push_back(i + vote_off[b]);
    g_member_w.insert(g_member_w.end(), mw.begin(), mw.end());
    maxima_off_out[b + 1] = total;
  }
  return PCDB_OK;
}

int orc_get_maximum_votes(int64_t* vote_index_out, float* vote_weight_out, int64_t capacity, int64_t* n_out) {
  *n_out = (int64_t)g_member_idx.size();
  if (capacity < *n_out) return PCDB_E_CAPACITY;
  std::memcpy(vote_index_out, g_member_idx.data(), sizeof(int64_t) * g_member_idx.size());
  std::memcpy(vote_weight_out, g_member_w.data(), sizeof(float) * g_member_w.size());
  return PCDB_OK;
}

/* ImplicitShapeModel::detect for B clouds, sequential over clouds as eval_tool does
 * (src/eval_tool/eval_classification.cpp:347-356); OpenMP inside the stages as the reference:
 * over keypoints (PCL *OMP estimators), over features (codebook.cpp:483); mean-shift sequential. */
int orc_classify_batch(void* model, const float* xyz, const float* normals, const uint32_t* rgb,
                       const int64_t* cloud_off, int32_t B, int32_t* label_out, pcdb_maximum* maxima_out,
                       int64_t* maxima_off_out, int64_t maxima_capacity, double* times_ms_out, int64_t* counts_out) {
  const Model& m = *static_cast<Model*>(model);
  const pcdb_params& P = m.prm;
  double t_feat = 0, t_kp = 0, t_vote = 0, t_max = 0;
  int64_t total_max = 0;
  int64_t c_kp = 0, c_feat = 0, c_nl = 0, c_ns = 0, c_votes = 0;
  if (maxima_off_out) maxima_off_out[0] = 0;
  auto T0 = std::chrono::steady_clock::now();
  for (int b = 0; b < B; ++b) {
    auto t0 = std::chrono::steady_clock::now();
    CloudFeatures f;
    int64_t s = cloud_off[b];
    double kp_ms = 0;
    compute_features_cloud(P, xyz + 3 * s, normals ? normals + 3 * s : nullptr, rgb ? rgb + s : nullptr,
                           int(cloud_off[b + 1] - s), f, &kp_ms);
    auto t1 = std::chrono::steady_clock::now();
    t_kp += kp_ms;
    t_feat += std::chrono::duration<double, std::milli>(t1 - t0).count() - kp_ms;
    int64_t Q = (int64_t)f.xyz.size() / 3;
    c_kp += f.n_kp;
    c_feat += Q;
    c_nl += f.n_lrf_nb;
    c_ns += f.n_shot_nb;
    const int k = P.knn_k;
    std::vector<int32_t> idx(size_t(Q) * k), cnt(Q);
    std::vector<float> dist(size_t(Q) * k);
    std::vector<pcdb_vote> votes;
    if (m.N > 0 && Q > 0) {
      activate_block(m, f.desc.data(), Q, k, P.distance_type, P.use_distance_ratio != 0, P.distance_ratio_threshold,
                     idx.data(), dist.data(), cnt.data());
      for (int64_t q = 0; q < Q; ++q)
        for (int j = 0; j < cnt[q]; ++j) {
          /* castVotes recomputes the functor value (codeword_distribution.cpp:87) — same value */
          cast_votes_one(m, &f.xyz[3 * q], &f.lrf[9 * q], nullptr, idx[q * k + j], dist[q * k + j], votes);
        }
    }
    c_votes += (int64_t)votes.size();
    auto t2 = std::chrono::steady_clock::now();
    t_vote += std::chrono::duration<double, std::milli>(t2 - t1).count();
    std::vector<pcdb_maximum> mx;
    std::vector<int64_t> mi;
    std::vector<float> mw;
    find_maxima_cloud(P, ClassDims{&m.dim_first, &m.dim_second}, f.surf.data(), (int64_t)f.surf.size() / 3,
                      votes.data(), (int64_t)votes.size(), mx, mi, mw);
    auto t3 = std::chrono::steady_clock::now();
    t_max += std::chrono::duration<double, std::milli>(t3 - t2).count();
    label_out[b] = mx.empty() ? -1 : int32_t(mx[0].class_id); /* eval_classification.cpp:412-417 */
    if (maxima_out) {
      if (total_max + (int64_t)mx.size() > maxima_capacity) {
        g_err = "maxima_capacity too small";
        return PCDB_E_CAPACITY;
      }
      for (const pcdb_maximum& x : mx) maxima_out[total_max++] = x;
      maxima_off_out[b + 1] = total_max;
    }
  }
  if (times_ms_out) {
    times_ms_out[0] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - T0).count();
    times_ms_out[1] = t_feat;
    times_ms_out[2] = t_kp;
    times_ms_out[3] = 0;
    times_ms_out[4] = 0;
    times_ms_out[5] = t_vote;
    times_ms_out[6] = t_max;
  }
  if (counts_out) {
    counts_out[0] = c_kp;
    counts_out[1] = c_feat;
    counts_out[2] = c_nl;
    counts_out[3] = c_ns;
    counts_out[4] = c_votes;
  }
  return PCDB_OK;
}

/* ---- training: ImplicitShapeModel::train + Codebook::activate ------------------------------- */
/* (implicit_shape_model.cpp:252-500, codebook/codebook.cpp:64-368,                              */
/*  codebook/codeword_distribution.cpp:37-71,171-243).  Clustering "None", ranking "Uniform".    */
/* Features must be given class-major (ascending class id, then model, then feature) — the       */
/* iteration order of the reference's std::map.  cloud_* arrays have one entry per training      */
/* cloud: class, instance, bbox (pos3, quat wxyz 4, size3).                                      */
int orc_aabb(const float* xyz, int64_t n, float* bbox10) { /* Utils::computeAABB utils.cpp:221-233 */
  float mn[3], mx[3];
  bool any = false;
  for (int64_t i = 0; i < n; ++i) {
    const float* p = xyz + 3 * i;
    if (!finite3(p)) continue;
    for (int a = 0; a < 3; ++a) {
      mn[a] = any ? std::min(mn[a], p[a]) : p[a];
      mx[a] = any ? std::max(mx[a], p[a]) : p[a];
    }
    any = true;
  }
  if (!any) return PCDB_E_INVALID;
  for (int a = 0; a < 3; ++a) {
    float size = mx[a] - mn[a];
    bbox10[7 + a] = size;
    bbox10[a] = mn[a] + (size / 2);
  }
  bbox10[3] = 1;
  bbox10[4] = bbox10[5] = bbox10[6] = 0;
  return PCDB_OK;
}

int orc_train(const pcdb_params* prm, const float* feat_xyz, const float* feat_lrf9, const float* feat_desc,
              const int64_t* feat_off, int32_t n_clouds, const uint32_t* cloud_class, const uint32_t* cloud_instance,
              const float* cloud_bbox10, int32_t n_classes, int64_t* N_out, int64_t* V_out) {
  const pcdb_params& P = *prm;
  const int D = P.feature_type == PCDB_FEATURE_CSHOT ? PCDB_CSHOT_DIM : PCDB_SHOT_DIM;
  const int64_t F = feat_off[n_clouds];
  const int k = P.knn_k;
  for (int c = 1; c < n_clouds; ++c)
    if (cloud_class[c] < cloud_class[c - 1]) {
      g_err = "training clouds must be class-major";
      return PCDB_E_INVALID;
    }
  /* codewords = features (clustering None): ids 0..F-1 (codeword.cpp:28-29) */
  Model cw;
  cw.prm = P;
  cw.N = F;
  cw.D = D;
  cw.words.assign(feat_desc, feat_desc + F * D);
  std::vector<int32_t> idx(size_t(F) * k), cnt(F);
  std::vector<float> dist(size_t(F) * k);
  /* training-time activation: distance ratio is detection-only (activation_strategy_knn.h:67) */
  activate_block(cw, feat_desc, F, k, P.distance_type, false, 0.f, idx.data(), dist.data(), cnt.data());
  struct Entry {
    std::vector<float> votes, bbox;
    std::vector<uint32_t> cls, inst;
    std::vector<int64_t> feat; /* activating feature (for computeWeights) */
    std::vector<int> cloud;
  };
  std::map<int, Entry> distribution;
  std::vector<float> sigma2(n_classes, 1.0f);
  std::vector<int> feat_cloud(F);
  for (int c = 0; c < n_clouds; ++c)
    for (int64_t f = feat_off[c]; f < feat_off[c + 1]; ++f) feat_cloud[f] = c;
  int c0 = 0;
  while (c0 < n_clouds) {
    int c1 = c0;
    unsigned cls = cloud_class[c0];
    while (c1 < n_clouds && cloud_class[c1] == cls) ++c1;
    int64_t num_features = feat_off[c1] - feat_off[c0];
    int max_elements = int(std::sqrt(double(num_features)));
    std::vector<int64_t> allModelFeatures;
    std::vector<int> allActivated;
    for (int c = c0; c < c1; ++c) {
      const float* bb = cloud_bbox10 + 10 * c;
      for (int64_t f = feat_off[c]; f < feat_off[c + 1]; ++f) {
        Quat rq = lrf_quat(feat_lrf9 + 9 * f);
        for (int j = 0; j < cnt[f]; ++j) {
          Entry& e = distribution[idx[f * k + j]];
          /* CodewordDistribution::addCodeword :37-71 */
          float vote[3] = {bb[0] - feat_xyz[3 * f], bb[1] - feat_xyz[3 * f + 1], bb[2] - feat_xyz[3 * f + 2]};
          float rot[3];
          quat_rotate(rq, vote, rot);
          e.votes.insert(e.votes.end(), rot, rot + 3);
          e.cls.push_back(cls);
          e.inst.push_back(cloud_instance[c]);
          Quat bq{bb[3], bb[4], bb[5], bb[6]};
          Quat nq = qmul(bq, qconj(rq));
          float nb[7] = {nq.a, nq.b, nq.c, nq.d, bb[7], bb[8], bb[9]};
          e.bbox.insert(e.bbox.end(), nb, nb + 7);
          e.feat.push_back(f);
          e.cloud.push_back(c);
        }
        if ((int)allActivated.size() < max_elements)
          for (int j = 0; j < cnt[f]; ++j) allActivated.push_back(idx[f * k + j]);
      }
      if ((int)allModelFeatures.size() < max_elements)
        for (int64_t f = feat_off[c]; f < feat_off[c + 1]; ++f) allModelFeatures.push_back(f);
    }
    /* class variance codebook.cpp:166-193 */
    float sum = 0;
    std::vector<float> distances;
    for (int64_t f : allModelFeatures)
      for (int w : allActivated) {
        float d = dist_fn(P.distance_type, feat_desc + f * D, feat_desc + int64_t(w) * D, D);
        sum += d;
        distances.push_back(d);
      }
    int num = int(allModelFeatures.size() * allActivated.size());
    float mean = sum / num;
    float variance = 0;
    for (float d : distances) {
      float diff = d - mean;
      variance += diff * diff;
    }
    variance /= num - 1;
    if (cls < (unsigned)n_classes) sigma2[cls] = variance;
    c0 = c1;
  }
  /* clean-up for KNN k == 1 (codebook.cpp:201-224) */
  if (k == 1) {
    for (auto it = distribution.begin(); it != distribution.end();)
      if (it->second.cls.size() != 1)
        it = distribution.erase(it);
      else
        ++it;
  }
  Trained& T = g_trained;
  T = Trained();
  T.D = D;
  T.sigma2 = sigma2;
  T.vote_off.push_back(0);
  for (auto& kv : distribution) {
    const Entry& e = kv.second;
    const int nv = int(e.cls.size());
    T.ids.push_back(kv.first);
    T.words.insert(T.words.end(), feat_desc + int64_t(kv.first) * D, feat_desc + int64_t(kv.first + 1) * D);
    T.kp_train.insert(T.kp_train.end(), feat_xyz + 3 * int64_t(kv.first), feat_xyz + 3 * int64_t(kv.first) + 3);
    /* computeWeights codeword_distribution.cpp:171-243 */
    for (int i = 0; i < nv; ++i) {
      std::vector<float> lw;
      const float* bbi = cloud_bbox10 + 10 * e.cloud[i];
      for (int j = 0; j < nv; ++j) {
        int64_t f = e.feat[j];
        Quat rq = lrf_quat(feat_lrf9 + 9 * f);
        float rot[3];
        quat_rotate_inv(rq, &e.votes[3 * i], rot);
        float center[3] = {feat_xyz[3 * f] + rot[0], feat_xyz[3 * f + 1] + rot[1], feat_xyz[3 * f + 2] + rot[2]};
        float d = norm3(center, bbi);
        const float sigma = 0.5f;
        lw.push_back(float(std::exp((-1 * (d * d)) / (sigma * sigma))));
      }
      std::sort(lw.begin(), lw.end());
      float median = lw.size() % 2 == 0 ? (lw[lw.size() / 2 - 1] + lw[lw.size() / 2]) / 2 : lw[lw.size() / 2];
      T.vote_weight.push_back(median);
    }
    T.vote_xyz.insert(T.vote_xyz.end(), e.votes.begin(), e.votes.end());
    T.vote_bbox.insert(T.vote_bbox.end(), e.bbox.begin(), e.bbox.end());
    T.vote_class.insert(T.vote_class.end(), e.cls.begin(), e.cls.end());
    T.vote_instance.insert(T.vote_instance.end(), e.inst.begin(), e.inst.end());
    T.vote_off.push_back((int64_t)T.vote_class.size());
  }
  T.N = (int64_t)T.ids.size();
  T.V = (int64_t)T.vote_class.size();
  T.vote_class_weight.assign(T.V, 1.0f); /* statistical weights (codebook.cpp:226-366): UseClassWeight=false on this path */
  *N_out = T.N;
  *V_out = T.V;
  return PCDB_OK;
}

int orc_train_fetch(float* words, int64_t* vote_off, float* vote_xyz, float* vote_weight, uint32_t* vote_class,
                    uint32_t* vote_instance, float* vote_bbox, float* vote_class_weight, float* kp_train,
                    int32_t* codeword_ids, float* class_sigma2, int32_t n_classes) {
  const Trained& T = g_trained;
  std::memcpy(words, T.words.data(), sizeof(float) * T.words.size());
  std::memcpy(vote_off, T.vote_off.data(), sizeof(int64_t) * T.vote_off.size());
  std::memcpy(vote_xyz, T.vote_xyz.data(), sizeof(float) * T.vote_xyz.size());
  std::memcpy(vote_weight, T.vote_weight.data(), sizeof(float) * T.vote_weight.size());
  std::memcpy(vote_class, T.vote_class.data(), sizeof(uint32_t) * T.vote_class.size());
  std::memcpy(vote_instance, T.vote_instance.data(), sizeof(uint32_t) * T.vote_instance.size());
  std::memcpy(vote_bbox, T.vote_bbox.data(), sizeof(float) * T.vote_bbox.size());
  std::memcpy(vote_class_weight, T.vote_class_weight.data(), sizeof(float) * T.vote_class_weight.size());
  std::memcpy(kp_train, T.kp_train.data(), sizeof(float) * T.kp_train.size());
  std::memcpy(codeword_ids, T.ids.data(), sizeof(int32_t) * T.ids.size());
  for (int c = 0; c < n_classes; ++c) class_sigma2[c] = c < (int)T.sigma2.size() ? T.sigma2[c] : 1.0f;
  return PCDB_OK;
}

/* Per-shard top-k merge */
int orc_merge_topk(const int32_t* cand_idx, const float* cand_dist, int32_t S, int64_t Q, int32_t k, int32_t* idx_out,
                   float* dist_out) {
  for (int64_t q = 0; q < Q; ++q) {
    std::vector<Cand> c;
    for (int s = 0; s < S; ++s)
      for (int j = 0; j < k; ++j) {
        int64_t o = (int64_t(s) * Q + q) * k + j;
        if (cand_idx[o] >= 0) c.push_back({cand_dist[o], cand_idx[o]});
      }
    std::sort(c.begin(), c.end(), cand_less);
    for (int j = 0; j < k; ++j) {
      idx_out[q * k + j] = j < (int)c.size() ? c[j].idx : -1;
      dist_out[q * k + j] = j < (int)c.size() ? c[j].d : kNaNf;
    }
  }
  return PCDB_OK;
}

} /* extern "C" */
