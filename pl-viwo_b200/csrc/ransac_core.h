// Core of the fundamental-matrix RANSAC gate (cv::findFundamentalMat(FM_RANSAC) restated), shared by
// the host implementation (ransac.cpp) and the single-CTA device gate of the stream group (kernels_glue.cu): OpenCV's
// multiply-with-carry RNG, the 7-point solver with OpenCV's null-space basis, cv::solveCubic, the collinearity test of the
// sample, the symmetric epipolar error and the adaptive iteration count.  Both builds compile this without floating-point
// contraction (-ffp-contract=off / --fmad=false), so host and device evaluate the same IEEE operations; they can differ only
// through the last-ulp behaviour of acos / cos / pow / log, which decides nothing but exact ties.
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define PLVIWO_HD __host__ __device__
#else
#define PLVIWO_HD
#endif

namespace plviwo {
namespace ransac_core {

struct CvRng {  // cv::RNG: state = (state & 0xffffffff) * 4164903690 + (state >> 32)
  uint64_t state;
  PLVIWO_HD explicit CvRng(uint64_t s) : state(s ? s : 0xffffffffu) {}
  PLVIWO_HD unsigned next() {
    state = (uint64_t)(unsigned)state * 4164903690u + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  PLVIWO_HD int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

PLVIWO_HD inline int cv_round(double v) { return (int)nearbyint(v); }

// Null-space basis of the 7 x 9 system exactly as cv::SVDecomp(A, W, U, Vt, MODIFY_A | FULL_UV) delivers it in rows
// 7 and 8 of Vt.  OpenCV's JacobiSVD (no LAPACK in the build) orthogonalises the 7 rows, then produces the two
// missing right singular vectors by starting from fixed pseudo-random +-1/9 vectors (cv::RNG(0x12345678)) and
// projecting out every earlier row twice (with an L1 renormalisation after each projection).  That completion is a
// projection onto the orthogonal complement of the row space, so it does not depend on WHICH orthonormal basis of
// the row space is used: a re-orthogonalised Gram-Schmidt basis (1 us) gives the same two vectors as the Jacobi
// sweeps (10+ us) to rounding error, and the root order of the cubic — which breaks inlier-count ties — with it.
// Returns false when the 7 rows are numerically rank deficient (the caller then treats the sample as degenerate).
PLVIWO_HD inline bool null_space_7x9(double At[9][9]) {
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
    double s = 1 / sqrt(norm);
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
            asum += fabs(t);
          }
          asum = asum > eps * 100 ? 1 / asum : 0;
          for (int k = 0; k < m; k++) At[i][k] *= asum;
        }
      }
      sd = 0;
      for (int k = 0; k < m; k++) sd += At[i][k] * At[i][k];
      sd = sqrt(sd);
    }
    double s = sd > minval ? 1 / sd : 0.;
    for (int k = 0; k < m; k++) At[i][k] *= s;
  }
  return true;
}

// cv::solveCubic for a0 x^3 + a1 x^2 + a2 x + a3 (double coefficients)
PLVIWO_HD inline int solve_cubic(const double c[4], double r[3]) {
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
        d = sqrt(d);
        double q1 = (-a2 + d) * 0.5;
        double q2 = (a2 + d) * -0.5;
        if (fabs(q1) > fabs(q2)) {
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
      double theta = acos(R / sqrt(Qcubed));
      double sqrtQ = sqrt(Q);
      double t0 = -2 * sqrtQ;
      double t1 = theta * (1. / 3);
      double t2 = a1 * (1. / 3);
      x0 = t0 * cos(t1) - t2;
      x1 = t0 * cos(t1 + (2. * kPi / 3)) - t2;
      x2 = t0 * cos(t1 + (4. * kPi / 3)) - t2;
      n = 3;
    } else if (d == 0) {
      if (R >= 0) {
        x0 = -2 * pow(R, 1. / 3) - a1 / 3;
        x1 = pow(R, 1. / 3) - a1 / 3;
      } else {
        x0 = 2 * pow(-R, 1. / 3) - a1 / 3;
        x1 = -pow(-R, 1. / 3) - a1 / 3;
      }
      x2 = 0;
      n = x0 == x1 ? 1 : 2;
      x1 = x0 == x1 ? 0 : x1;
    } else {
      d = sqrt(-d);
      double e = pow(d + fabs(R), 1. / 3);
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
PLVIWO_HD inline int run_7point(const float *m1, const float *m2, double F[27]) {
  double At[9][9];
  for (int i = 0; i < 9; i++)
    for (int k = 0; k < 9; k++) At[i][k] = 0;
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
    if (fabs(s) > DBL_EPSILON) {
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

PLVIWO_HD inline bool have_collinear(const float *m, int count) {
  int i = count - 1;
  for (int j = 0; j < i; j++) {
    double dx1 = m[2 * j] - m[2 * i];
    double dy1 = m[2 * j + 1] - m[2 * i + 1];
    for (int k = 0; k < j; k++) {
      double dx2 = m[2 * k] - m[2 * i];
      double dy2 = m[2 * k + 1] - m[2 * i + 1];
      if (fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2)))
        return true;
    }
  }
  return false;
}

PLVIWO_HD inline bool get_subset(const float *m1, const float *m2, int count, float *ms1, float *ms2, CvRng &rng, int max_attempts) {
  const int model_points = 7;
  int idx[7];
  int i = 0, iters = 0;
  for (; iters < max_attempts; iters++) {
    for (i = 0; i < model_points && iters < max_attempts;) {
      int idx_i;
      for (;;) {
        idx_i = rng.uniform(0, count);
        bool dup = false;
        for (int q = 0; q < i; q++) dup = dup || idx[q] == idx_i;
        if (!dup) break;
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

// FMEstimatorCallback::computeError for one correspondence: max of the two squared point-to-epipolar-line distances
PLVIWO_HD inline float epipolar_error(const double *F, double x1, double y1, double x2, double y2) {
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
  double e1 = d1 * d1 * s1, e2 = d2 * d2 * s2;
  return (float)(e1 > e2 ? e1 : e2);
}

PLVIWO_HD inline int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = fmax(p, 0.);
  p = fmin(p, 1.);
  ep = fmax(ep, 0.);
  ep = fmin(ep, 1.);
  double num = fmax(1. - p, DBL_MIN);
  double denom = 1. - pow(1. - ep, model_points);
  if (denom < DBL_MIN) return 0;
  num = log(num);
  denom = log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : cv_round(num / denom);
}


}  // namespace ransac_core
}  // namespace plviwo
