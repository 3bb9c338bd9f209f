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

#include "ransac_core.h"

namespace plviwo {

namespace {

using namespace ransac_core;

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
