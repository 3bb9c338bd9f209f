// Fundamental-matrix RANSAC gate of the point tracker — the one sequential step of perform_matching.
//
// Replaces cv::findFundamentalMat(pts0_n, pts1_n, cv::FM_RANSAC, 2.0 / max_focal, 0.999, mask) at
// TrackKLT.cpp:869-873.  OpenCV is not linkable here, so this is an independent C++ implementation of the
// published algorithm: RANSACPointSetRegistrator with the 7-point solver, OpenCV's
// deterministic multiply-with-carry RNG seeded with 2^64-1 per call, the collinearity test on the 7th sample,
// symmetric epipolar distance, adaptive iteration count.  The null-space basis of the 7x9 system is the one
// OpenCV's own SVD produces (two vectors completed from a fixed pseudo-random start by double Gram-Schmidt), so
// that the pencil basis — and therefore the order of the cubic's roots, which breaks inlier-count ties — matches
// the library.  n < 15 switches to LMedS exactly as
// OpenCV silently does.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

namespace plviwo {

namespace {

struct CvRng {  // cv::RNG: state = (state & 0xffffffff) * 4164903690 + (state >> 32)
  uint64_t state;
  explicit CvRng(uint64_t s) : state(s ? s : 0xffffffffu) {}
  unsigned next() {
    state = (uint64_t)(unsigned)state * 4164903690u + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

inline int cv_round(double v) { return (int)std::nearbyint(v); }

// Null-space basis of the 7 x 9 system exactly as cv::SVDecomp(A, W, U, Vt, MODIFY_A | FULL_UV) delivers it in rows
// 7 and 8 of Vt.  OpenCV's JacobiSVD (no LAPACK in the build) orthogonalises the 7 rows, then produces the two
// missing right singular vectors by starting from fixed pseudo-random +-1/9 vectors (cv::RNG(0x12345678)) and
// projecting out every earlier row twice (with an L1 renormalisation after each projection).  That completion is a
// projection onto the orthogonal complement of the row space, so it does not depend on WHICH orthonormal basis of
// the row space is used: a re-orthogonalised Gram-Schmidt basis (1 us) gives the same two vectors as the Jacobi
// sweeps (10+ us) to rounding error, and the root order of the cubic — which breaks inlier-count ties — with it.
// Returns false when the 7 rows are numerically rank deficient (the caller then treats the sample as degenerate).
bool null_space_7x9(double At[9][9]) {
  const int m = 9, n = 7, n1 = 9;
  const double eps = DBL_EPSILON * 10, minval = DBL_MIN;
  for (int i = 0; i < n; i++) {
    double norm0 = 0;
    for (int k = 0; k < m; k++) norm0 += At[i][k] * At[i][k];
    for (int pass = 0; pass < 2; pass++)
      for (int j = 0; j < i; j++) {
        double d = 0;
        for (int k = 0; k < m; k++) d += At[i][k] * At[j][k];
        for (int k = 0; k < m; k++) At[i][k] -= d * At[j][k];
      }
    double norm = 0;
    for (int k = 0; k < m; k++) norm += At[i][k] * At[i][k];
    if (!(norm > 1e-24 * norm0) || !(norm > 0)) return false;
    double s = 1 / std::sqrt(norm);
    for (int k = 0; k < m; k++) At[i][k] *= s;
  }
  CvRng rng(0x12345678);
  for (int i = n; i < n1; i++) {
    double sd = 0;
    for (int ii = 0; ii < 100 && sd <= minval; ii++) {
      const double val0 = 1. / m;
      for (int k = 0; k < m; k++) At[i][k] = (rng.next() & 256) != 0 ? val0 : -val0;
      for (int iter = 0; iter < 2; iter++) {
        for (int j = 0; j < i; j++) {
          sd = 0;
          for (int k = 0; k < m; k++) sd += At[i][k] * At[j][k];
          double asum = 0;
          for (int k = 0; k < m; k++) {
            double t = At[i][k] - sd * At[j][k];
            At[i][k] = t;
            asum += std::abs(t);
          }
          asum = asum > eps * 100 ? 1 / asum : 0;
          for (int k = 0; k < m; k++) At[i][k] *= asum;
        }
      }
      sd = 0;
      for (int k = 0; k < m; k++) sd += At[i][k] * At[i][k];
      sd = std::sqrt(sd);
    }
    double s = sd > minval ? 1 / sd : 0.;
    for (int k = 0; k < m; k++) At[i][k] *= s;
  }
  return true;
}

// cv::solveCubic for a0 x^3 + a1 x^2 + a2 x + a3 (double coefficients)
int solve_cubic(const double c[4], double r[3]) {
  double a0 = c[0], a1 = c[1], a2 = c[2], a3 = c[3];
  double x0 = 0, x1 = 0, x2 = 0;
  int n = 0;
  if (a0 == 0) {
    if (a1 == 0) {
      if (a2 == 0) {
        n = a3 == 0 ? -1 : 0;
      } else {
        x0 = -a3 / a2;
        n = 1;
      }
    } else {
      double d = a2 * a2 - 4 * a1 * a3;
      if (d >= 0) {
        d = std::sqrt(d);
        double q1 = (-a2 + d) * 0.5;
        double q2 = (a2 + d) * -0.5;
        if (std::fabs(q1) > std::fabs(q2)) {
          x0 = q1 / a1;
          x1 = a3 / q1;
        } else {
          x0 = q2 / a1;
          x1 = a3 / q2;
        }
        n = d > 0 ? 2 : 1;
      }
    }
  } else {
    a0 = 1. / a0;
    a1 *= a0;
    a2 *= a0;
    a3 *= a0;
    double Q = (a1 * a1 - 3 * a2) * (1. / 9);
    double R = (2 * a1 * a1 * a1 - 9 * a1 * a2 + 27 * a3) * (1. / 54);
    double Qcubed = Q * Q * Q;
    double d = Qcubed - R * R;
    const double kPi = 3.1415926535897932384626433832795;
    if (d > 0) {
      double theta = std::acos(R / std::sqrt(Qcubed));
      double sqrtQ = std::sqrt(Q);
      double t0 = -2 * sqrtQ;
      double t1 = theta * (1. / 3);
      double t2 = a1 * (1. / 3);
      x0 = t0 * std::cos(t1) - t2;
      x1 = t0 * std::cos(t1 + (2. * kPi / 3)) - t2;
      x2 = t0 * std::cos(t1 + (4. * kPi / 3)) - t2;
      n = 3;
    } else if (d == 0) {
      if (R >= 0) {
        x0 = -2 * std::pow(R, 1. / 3) - a1 / 3;
        x1 = std::pow(R, 1. / 3) - a1 / 3;
      } else {
        x0 = 2 * std::pow(-R, 1. / 3) - a1 / 3;
        x1 = -std::pow(-R, 1. / 3) - a1 / 3;
      }
      x2 = 0;
      n = x0 == x1 ? 1 : 2;
      x1 = x0 == x1 ? 0 : x1;
    } else {
      d = std::sqrt(-d);
      double e = std::pow(d + std::fabs(R), 1. / 3);
      if (R > 0) e = -e;
      x0 = (e + Q / e) - a1 * (1. / 3);
      n = 1;
    }
  }
  r[0] = x0;
  r[1] = x1;
  r[2] = x2;
  return n;
}

// run7Point: up to 3 fundamental matrices (row-major 9 doubles each)
int run_7point(const float *m1, const float *m2, double F[27]) {
  double At[9][9];
  std::memset(At, 0, sizeof(At));
  for (int i = 0; i < 7; i++) {
    double x0 = m1[2 * i], y0 = m1[2 * i + 1];
    double x1 = m2[2 * i], y1 = m2[2 * i + 1];
    double *a = At[i];
    a[0] = x1 * x0; a[1] = x1 * y0; a[2] = x1;
    a[3] = y1 * x0; a[4] = y1 * y0; a[5] = y1;
    a[6] = x0; a[7] = y0; a[8] = 1;
  }
  if (!null_space_7x9(At)) return 0;
  double *f1 = At[7], *f2 = At[8];
  for (int i = 0; i < 9; i++) f1[i] -= f2[i];
  double c[4], r[3] = {0, 0, 0};
  double t0 = f2[4] * f2[8] - f2[5] * f2[7];
  double t1 = f2[3] * f2[8] - f2[5] * f2[6];
  double t2 = f2[3] * f2[7] - f2[4] * f2[6];
  c[3] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
  c[2] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 - f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
         f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) - f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
         f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) - f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
         f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
  t0 = f1[4] * f1[8] - f1[5] * f1[7];
  t1 = f1[3] * f1[8] - f1[5] * f1[6];
  t2 = f1[3] * f1[7] - f1[4] * f1[6];
  c[0] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;
  c[1] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 - f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
         f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) - f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
         f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) - f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
         f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
  int n = solve_cubic(c, r);
  if (n < 1 || n > 3) return n;
  for (int k = 0; k < n; k++) {
    double *fm = F + 9 * k;
    double lambda = r[k], mu = 1.;
    double s = f1[8] * r[k] + f2[8];
    if (std::fabs(s) > DBL_EPSILON) {
      mu = 1. / s;
      lambda *= mu;
      fm[8] = 1.;
    } else {
      fm[8] = 0.;
    }
    for (int i = 0; i < 8; i++) fm[i] = f1[i] * lambda + f2[i] * mu;
  }
  return n;
}

bool have_collinear(const float *m, int count) {
  int i = count - 1;
  for (int j = 0; j < i; j++) {
    double dx1 = m[2 * j] - m[2 * i];
    double dy1 = m[2 * j + 1] - m[2 * i + 1];
    for (int k = 0; k < j; k++) {
      double dx2 = m[2 * k] - m[2 * i];
      double dy2 = m[2 * k + 1] - m[2 * i + 1];
      if (std::fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (std::fabs(dx1) + std::fabs(dy1) + std::fabs(dx2) + std::fabs(dy2)))
        return true;
    }
  }
  return false;
}

bool get_subset(const float *m1, const float *m2, int count, float *ms1, float *ms2, CvRng &rng, int max_attempts) {
  const int model_points = 7;
  int idx[7];
  int i = 0, iters = 0;
  for (; iters < max_attempts; iters++) {
    for (i = 0; i < model_points && iters < max_attempts;) {
      int idx_i;
      for (idx_i = rng.uniform(0, count); std::find(idx, idx + i, idx_i) != idx + i; idx_i = rng.uniform(0, count)) {
      }
      idx[i] = idx_i;
      ms1[2 * i] = m1[2 * idx_i]; ms1[2 * i + 1] = m1[2 * idx_i + 1];
      ms2[2 * i] = m2[2 * idx_i]; ms2[2 * i + 1] = m2[2 * idx_i + 1];
      i++;
    }
    if (i == model_points && (have_collinear(ms1, i) || have_collinear(ms2, i))) continue;
    break;
  }
  return i == model_points && iters < max_attempts;
}

// FMEstimatorCallback::computeError on structure-of-arrays inputs (doubles prepared once per call); the loop has no
// cross-iteration dependence, so the compiler vectorises it.  Operation order per element is the library's.
struct Soa {
  std::vector<double> x1, y1, x2, y2;
};
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx2", "default")))
#endif
int count_inliers_chunk(const Soa &p, int i0, int i1, const double *F, float t, uint8_t *mask) {
  const double *X1 = p.x1.data(), *Y1 = p.y1.data(), *X2 = p.x2.data(), *Y2 = p.y2.data();
  const double f0 = F[0], f1 = F[1], f2 = F[2], f3 = F[3], f4 = F[4], f5 = F[5], f6 = F[6], f7 = F[7], f8 = F[8];
  int good = 0;
  for (int i = i0; i < i1; i++) {
    double x1 = X1[i], y1 = Y1[i], x2 = X2[i], y2 = Y2[i];
    double a = f0 * x1 + f1 * y1 + f2;
    double b = f3 * x1 + f4 * y1 + f5;
    double c = f6 * x1 + f7 * y1 + f8;
    double s2 = 1. / (a * a + b * b);
    double d2 = x2 * a + y2 * b + c;
    a = f0 * x2 + f3 * y2 + f6;
    b = f1 * x2 + f4 * y2 + f7;
    c = f2 * x2 + f5 * y2 + f8;
    double s1 = 1. / (a * a + b * b);
    double d1 = x1 * a + y1 * b + c;
    double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
    uint8_t f = (float)(e1 > e2 ? e1 : e2) <= t;
    mask[i] = f;
    good += f;
  }
  return good;
}

// findInliers with an exact early exit: a model is only ever used if its inlier count EXCEEDS `need`, so the scan
// stops (returning -1) as soon as the points still to come cannot lift it above that.
int count_inliers(const Soa &p, int count, const double *F, float t, uint8_t *mask, int need) {
  int good = 0;
  for (int i0 = 0; i0 < count; i0 += 64) {
    const int i1 = i0 + 64 < count ? i0 + 64 : count;
    good += count_inliers_chunk(p, i0, i1, F, t, mask);
    if (good + (count - i1) <= need) return -1;
  }
  return good;
}

void compute_error(const float *m1, const float *m2, int count, const double *F, float *err) {
  for (int i = 0; i < count; i++) {
    double x1 = m1[2 * i], y1 = m1[2 * i + 1], x2 = m2[2 * i], y2 = m2[2 * i + 1];
    double a = F[0] * x1 + F[1] * y1 + F[2];
    double b = F[3] * x1 + F[4] * y1 + F[5];
    double c = F[6] * x1 + F[7] * y1 + F[8];
    double s2 = 1. / (a * a + b * b);
    double d2 = x2 * a + y2 * b + c;
    a = F[0] * x2 + F[3] * y2 + F[6];
    b = F[1] * x2 + F[4] * y2 + F[7];
    c = F[2] * x2 + F[5] * y2 + F[8];
    double s1 = 1. / (a * a + b * b);
    double d1 = x1 * a + y1 * b + c;
    err[i] = (float)std::max(d1 * d1 * s1, d2 * d2 * s2);
  }
}

int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = std::max(p, 0.);
  p = std::min(p, 1.);
  ep = std::max(ep, 0.);
  ep = std::min(ep, 1.);
  double num = std::max(1. - p, DBL_MIN);
  double denom = 1. - std::pow(1. - ep, model_points);
  if (denom < DBL_MIN) return 0;
  num = std::log(num);
  denom = std::log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : cv_round(num / denom);
}

}  // namespace

// Returns the number of inliers (0: no model, mask all zero); mask[i] in {0,1}.
// *mask_valid = 0 reproduces cv::findFundamentalMat returning an empty mask (n < 7).
int ransac_fundamental(const float *m1, const float *m2, int count, double threshold, double confidence, uint8_t *mask,
                       int *mask_valid) {
  const int model_points = 7, max_iters = 1000;
  std::memset(mask, 0, count);
  *mask_valid = 0;
  if (count < 7) return 0;
  *mask_valid = 1;
  if (threshold <= 0) threshold = 3;
  if (confidence < DBL_EPSILON || confidence > 1 - DBL_EPSILON) confidence = 0.99;
  double F[27];
  if (count == 7) {
    int n = run_7point(m1, m2, F);
    if (n <= 0) return 0;
    std::memset(mask, 1, count);
    return count;
  }
  // A repeated frame: every point tracked onto itself (LK's first step is exactly zero on identical images).  Every
  // 7-point system is then rank 6 — its null space is the antisymmetric matrices, for which x' F x = 0 holds for EVERY
  // point — and the library either returns such an F with all points as inliers or trips an internal assertion
  // (OpenCV 4.13: cv::Exception out of findFundamentalMat, which the reference does not catch).  The useful one of the
  // two outcomes is returned: all points are inliers.  (With a few moved points among static ones the sampled systems
  // that contain a moved point have full rank and the loop below finds the static set as inliers by itself.)
  {
    bool all_static = true;
    for (int k = 0; k < 2 * count && all_static; k++) all_static = m1[k] == m2[k];
    if (all_static) {
      std::memset(mask, 1, count);
      return count;
    }
  }
  std::vector<float> err(count);
  std::vector<uint8_t> cur(count), best(count, 0);
  float ms1[14], ms2[14];
  CvRng rng((uint64_t)-1);
  if (count >= 15) {
    // ---- RANSACPointSetRegistrator::run
    int niters = max_iters, max_good = 0;
    const float t = (float)(threshold * threshold);
    Soa soa;
    soa.x1.resize(count); soa.y1.resize(count); soa.x2.resize(count); soa.y2.resize(count);
    for (int k = 0; k < count; k++) {
      soa.x1[k] = m1[2 * k]; soa.y1[k] = m1[2 * k + 1];
      soa.x2[k] = m2[2 * k]; soa.y2[k] = m2[2 * k + 1];
    }
    for (int iter = 0; iter < niters; iter++) {
      if (!get_subset(m1, m2, count, ms1, ms2, rng, 10000)) {
        if (iter == 0) return 0;
        break;
      }
      int nmodels = run_7point(ms1, ms2, F);
      if (nmodels <= 0) continue;
      for (int i = 0; i < nmodels; i++) {
        int good = count_inliers(soa, count, F + 9 * i, t, cur.data(), std::max(max_good, model_points - 1));
        if (good > std::max(max_good, model_points - 1)) {
          std::swap(cur, best);
          max_good = good;
          niters = ransac_update_num_iters(confidence, (double)(count - good) / count, model_points, niters);
        }
      }
    }
    if (max_good > 0) std::memcpy(mask, best.data(), count);
    return max_good;
  }
  // ---- LMeDSPointSetRegistrator::run (8 <= count < 15).  Mask-exact against the library for count == 14.  For
  // count <= 13 the median is the (count / 2 + 1)-th smallest error with count / 2 <= 6, i.e. the error of one of the 7
  // SAMPLE points of the model, which is zero up to rounding (1e-33): the library's choice of the "best" model is then
  // decided by the rounding noise of its own 7-point solver and cannot be reproduced by any other build of the same
  // algorithm (tests/test_host_cpu.py documents it); this is a faithful LMedS, not a bit-copy, in that regime.
  {
    const double outlier_ratio = 0.45;
    int niters = ransac_update_num_iters(confidence, outlier_ratio, model_points, max_iters);
    niters = std::max(niters, 3);
    double min_median = DBL_MAX;
    bool found = false;
    double bestF[9];
    std::vector<float> sorted(count);
    for (int iter = 0; iter < niters; iter++) {
      if (!get_subset(m1, m2, count, ms1, ms2, rng, 1000)) {
        if (iter == 0) return 0;
        break;
      }
      int nmodels = run_7point(ms1, ms2, F);
      if (nmodels <= 0) continue;
      for (int i = 0; i < nmodels; i++) {
        compute_error(m1, m2, count, F + 9 * i, err.data());
        std::copy(err.begin(), err.end(), sorted.begin());
        std::nth_element(sorted.begin(), sorted.begin() + count / 2, sorted.end());
        double median = sorted[count / 2];
        if (median < min_median) {
          min_median = median;
          std::memcpy(bestF, F + 9 * i, sizeof(bestF));
          found = true;
        }
      }
    }
    if (!found || min_median >= DBL_MAX) return 0;
    double sigma = 2.5 * 1.4826 * (1 + 5. / (count - model_points)) * std::sqrt(min_median);
    sigma = std::max(sigma, 0.001);
    compute_error(m1, m2, count, bestF, err.data());
    const float t = (float)(sigma * sigma);
    int good = 0;
    for (int k = 0; k < count; k++) {
      uint8_t f = err[k] <= t;
      mask[k] = f;
      good += f;
    }
    return good;
  }
}

}  // namespace plviwo
