// ORACLE — test infrastructure only.  Nothing under oracle/ is linked, imported or executed by the product
// path (pl-viwo_b200/csrc); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may use it, as the checker / CPU baseline.
//
// Two pieces of the reference's CPU path that have no Python binding in this image:
//
//  (1) libstdc++ std::sort with the reference's comparator (Grider_FAST.h:57, called at Grider_GRID.h:128).
//      The sort is UNSTABLE and FAST responses are small integers, so which keypoints survive the
//      top-num_features_grid cut is decided inside tie groups by introsort's exact permutation.
//
//  (2) cv::ximgproc::FastLineDetector::detect (called at PL-VIWO/src/update/cam/TrackLSD.cpp:200-205).
//      opencv_contrib is absent from /root/reference and from this image (cv2.ximgproc missing), so this
//      is a restatement of the published algorithm (opencv_contrib 4.x modules/ximgproc/src/
//      fast_line_detector.cpp).  PARITY UNPINNED: there is no executable FLD here to
//      check it against; cv::Canny and cv::fitLine — its building blocks — are pinned against cv2 in
//      tests/test_oracle_pins.py.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace {

struct Kp {
  float x, y, size, angle, response;
  int octave, class_id;  // same footprint as cv::KeyPoint; class_id carries the original index
};
// Grider_FAST.h:57 — by-value comparator, strict '>' on response
bool compare_response(Kp first, Kp second) { return first.response > second.response; }

struct Pt2i { int x, y; };
struct Seg { float x1, y1, x2, y2; };

inline int cv_round(double v) { return (int)std::nearbyint(v); }  // round-half-even like cvRound

// cv::fitLine(points, DIST_L2, 0, 0.01, 0.01) on 2-D points == fitLine2D_wods (closed form)
void fit_line_l2(const std::vector<Pt2i> &pts, float line[4]) {
  double x = 0, y = 0, x2 = 0, y2 = 0, xy = 0;
  for (const Pt2i &p : pts) {
    float px = (float)p.x, py = (float)p.y;
    x += px; y += py;
    x2 += px * px; y2 += py * py; xy += px * py;
  }
  double w = (float)pts.size();
  x /= w; y /= w; x2 /= w; y2 /= w; xy /= w;
  double dx2 = x2 - x * x, dy2 = y2 - y * y, dxy = xy - x * y;
  float t = (float)std::atan2(2 * dxy, dx2 - dy2) / 2;
  line[0] = (float)std::cos(t);
  line[1] = (float)std::sin(t);
  line[2] = (float)x;
  line[3] = (float)y;
}

inline void cross3(const double a[3], const double b[3], double c[3]) {
  double c0 = a[1] * b[2] - a[2] * b[1];
  double c1 = a[2] * b[0] - a[0] * b[2];
  double c2 = a[0] * b[1] - a[1] * b[0];
  c[0] = c0; c[1] = c1; c[2] = c2;
}

// distPointLine: normalises l IN PLACE, then returns l . p
inline double dist_point_line(const double p[3], double l[3]) {
  double x = l[0], y = l[1];
  double w = std::sqrt(x * x + y * y);
  l[0] = x / w; l[1] = y / w; l[2] = l[2] / w;
  return l[0] * p[0] + l[1] * p[1] + l[2] * p[2];
}

struct Fld {
  int W, H, length_threshold;
  float distance_threshold;

  void incident_point(const double l[3], float &px, float &py) const {
    double a[3] = {(double)px, (double)py, 1.0};
    double b[3] = {l[0], l[1], 0.0};
    double lk[3], xk[3];
    cross3(a, b, lk);
    cross3(lk, l, xk);
    double alpha = 1.0 / xk[2];
    double X = xk[0] * alpha, Y = xk[1] * alpha;
    float fx = (float)X, fy = (float)Y;
    px = fx < 0.0f ? 0.0f : (fx >= (W - 1.0f) ? (W - 1.0f) : fx);
    py = fy < 0.0f ? 0.0f : (fy >= (H - 1.0f) ? (H - 1.0f) : fy);
  }
  void incident_point_i(const double l[3], Pt2i &p) const {
    float fx = (float)p.x, fy = (float)p.y;
    incident_point(l, fx, fy);
    p.x = cv_round(fx);  // cv::Point_<int>(Point2f) == saturate_cast<int> == cvRound
    p.y = cv_round(fy);
  }

  bool get_point_chain(const uint8_t *img, Pt2i pt, Pt2i &chained, int &direction, int step) const {
    static const int indices[8][2] = {{1, 1}, {1, 0}, {1, -1}, {0, -1}, {-1, -1}, {-1, 0}, {-1, 1}, {0, 1}};
    float min_dir_diff = 7.0f;
    Pt2i consistent_pt = {0, 0};
    int consistent_direction = 0;
    for (int i = 0; i < 8; i++) {
      int ci = pt.x + indices[i][1];
      int ri = pt.y + indices[i][0];
      if (ri < 0 || ri == H || ci < 0 || ci == W) continue;
      if (img[ri * W + ci] == 0) continue;
      if (step == 0) {
        chained.x = ci; chained.y = ri;
        direction = i > 4 ? i - 8 : i;
        return true;
      }
      int curr_dir = i > 4 ? i - 8 : i;
      int dir_diff = std::abs(curr_dir - direction);
      dir_diff = dir_diff > 4 ? 8 - dir_diff : dir_diff;
      if (dir_diff <= min_dir_diff) {
        min_dir_diff = (float)dir_diff;
        consistent_pt.x = ci; consistent_pt.y = ri;
        consistent_direction = curr_dir;
      }
    }
    if (min_dir_diff < 2) {
      chained = consistent_pt;
      direction = (direction * step + consistent_direction) / (step + 1);
      return true;
    }
    return false;
  }

  void line_from_fit(const float line[4], double l[3]) const {
    double a[3] = {line[2], line[3], 1.0};
    double b[3] = {(double)(line[2] + line[0]), (double)(line[3] + line[1]), 1.0};
    cross3(a, b, l);
  }

  void extract_segments(const std::vector<Pt2i> &points, std::vector<Seg> &segments) const {
    int total = (int)points.size();
    std::vector<Pt2i> l_points;
    int i, j;
    for (i = 0; i + length_threshold < total; i++) {
      Pt2i ps = points[i];
      Pt2i pe = points[i + length_threshold];
      double a[3] = {(double)ps.x, (double)ps.y, 1};
      double b[3] = {(double)pe.x, (double)pe.y, 1};
      double l[3];
      cross3(a, b, l);
      bool is_line = true;
      l_points.clear();
      l_points.push_back(ps);
      for (j = 1; j < length_threshold; j++) {
        Pt2i pt = points[i + j];
        double p[3] = {(double)pt.x, (double)pt.y, 1.0};
        double dist = dist_point_line(p, l);
        if (std::fabs(dist) > distance_threshold) { is_line = false; break; }
        l_points.push_back(pt);
      }
      if (!is_line) continue;
      l_points.push_back(pe);

      float line[4];
      fit_line_l2(l_points, line);
      line_from_fit(line, l);
      incident_point_i(l, ps);

      for (j = length_threshold + 1; i + j < total; j++) {
        Pt2i pt = points[i + j];
        double p[3] = {(double)pt.x, (double)pt.y, 1.0};
        double dist = dist_point_line(p, l);
        if (std::fabs(dist) > distance_threshold) {
          fit_line_l2(l_points, line);
          line_from_fit(line, l);
          dist = dist_point_line(p, l);
          if (std::fabs(dist) > distance_threshold) { j--; break; }
        }
        pe = pt;
        l_points.push_back(pt);
      }
      fit_line_l2(l_points, line);
      line_from_fit(line, l);
      float e1x = (float)ps.x, e1y = (float)ps.y, e2x = (float)pe.x, e2y = (float)pe.y;
      incident_point(l, e1x, e1y);
      incident_point(l, e2x, e2y);
      segments.push_back({e1x, e1y, e2x, e2y});
      i = i + j;
    }
  }

  // additionalOperationsOnSegment: orient the segment so the brighter side is consistent
  void orient(const uint8_t *src, int stride, Seg &seg) const {
    if (seg.x1 == 0.0f && seg.x2 == 0.0f && seg.y1 == 0.0f && seg.y2 == 0.0f) return;
    double ang = std::atan2((double)(seg.y2 - seg.y1), (double)(seg.x2 - seg.x1));
    double dx = (double)seg.x2 - (double)seg.x1, dy = (double)seg.y2 - (double)seg.y1;
    const int num_points = 10;
    float px[num_points], py[num_points];
    px[0] = seg.x1; py[0] = seg.y1;
    px[num_points - 1] = seg.x2; py[num_points - 1] = seg.y2;
    for (int i = 1; i < num_points - 1; i++) {
      px[i] = px[0] + ((float)dx / float(num_points - 1) * (float)i);
      py[i] = py[0] + ((float)dy / float(num_points - 1) * (float)i);
    }
    const double kPi = 3.1415926535897932384626433832795;
    double gap = 1.0;
    double c = gap * std::cos(90.0 * kPi / 180.0 + ang), s = gap * std::sin(90.0 * kPi / 180.0 + ang);
    int iR = 0, iL = 0;
    for (int i = 0; i < num_points; i++) {
      int rx = cv_round(px[i] + c), ry = cv_round(py[i] + s);
      int lx = cv_round(px[i] - c), ly = cv_round(py[i] - s);
      rx = std::min(std::max(rx, 0), W - 1); ry = std::min(std::max(ry, 0), H - 1);
      lx = std::min(std::max(lx, 0), W - 1); ly = std::min(std::max(ly, 0), H - 1);
      iR += src[ry * stride + rx];
      iL += src[ly * stride + lx];
    }
    if (iR > iL) { std::swap(seg.x1, seg.x2); std::swap(seg.y1, seg.y2); }
  }

  int detect(const uint8_t *src, int stride, uint8_t *canny, float *out, int cap) const {
    // only the two corner blocks are cleared (colRange(0,6).rowRange(0,6); last 5 rows x last 5 cols)
    for (int r = 0; r < std::min(6, H); r++)
      for (int c = 0; c < std::min(6, W); c++) canny[r * W + c] = 0;
    for (int r = std::max(H - 5, 0); r < H; r++)
      for (int c = std::max(W - 5, 0); c < W; c++) canny[r * W + c] = 0;
    std::vector<Pt2i> points;
    std::vector<Seg> segments;
    int n_out = 0;
    for (int r = 0; r < H; r++) {
      for (int c = 0; c < W; c++) {
        if (canny[r * W + c] == 0) continue;
        Pt2i pt = {c, r};
        points.push_back(pt);
        canny[r * W + c] = 0;
        int direction = 0, step = 0;
        while (get_point_chain(canny, pt, pt, direction, step)) {
          points.push_back(pt);
          step++;
          canny[pt.y * W + pt.x] = 0;
        }
        if (points.size() < (unsigned)length_threshold + 1) { points.clear(); continue; }
        extract_segments(points, segments);
        for (Seg seg : segments) {
          float length = std::sqrt((seg.x1 - seg.x2) * (seg.x1 - seg.x2) + (seg.y1 - seg.y2) * (seg.y1 - seg.y2));
          if (length < length_threshold) continue;
          if ((seg.x1 <= 5.0f && seg.x2 <= 5.0f) || (seg.y1 <= 5.0f && seg.y2 <= 5.0f) ||
              (seg.x1 >= W - 5.0f && seg.x2 >= W - 5.0f) || (seg.y1 >= H - 5.0f && seg.y2 >= H - 5.0f))
            continue;
          orient(src, stride, seg);
          if (n_out < cap) {
            out[4 * n_out + 0] = seg.x1; out[4 * n_out + 1] = seg.y1;
            out[4 * n_out + 2] = seg.x2; out[4 * n_out + 3] = seg.y2;
          }
          n_out++;
        }
        points.clear();
        segments.clear();
      }
    }
    return n_out;
  }
};

}  // namespace

extern "C" {

// perm_out[k] = original index of the keypoint that std::sort leaves at position k
void oracle_sort_perm(const float *response, int n, int *perm_out) {
  std::vector<Kp> v(n);
  for (int i = 0; i < n; i++) v[i] = Kp{0.f, 0.f, 7.f, -1.f, response[i], 0, i};
  std::sort(v.begin(), v.end(), compare_response);
  for (int i = 0; i < n; i++) perm_out[i] = v[i].class_id;
}

// src: the 8-bit image FLD was given (for segment orientation); canny: W*H edge map (0 / non-zero), consumed.
// Returns the number of segments found (may exceed cap; only cap are written), each (x1, y1, x2, y2).
int oracle_fld_detect(const uint8_t *src, int width, int height, int stride, uint8_t *canny, int length_threshold,
                      float distance_threshold, float *out_lines, int cap) {
  Fld f{width, height, length_threshold, distance_threshold};
  return f.detect(src, stride, canny, out_lines, cap);
}

void oracle_fit_line(const int *xy, int n, float *line4) {
  std::vector<Pt2i> pts(n);
  for (int i = 0; i < n; i++) pts[i] = {xy[2 * i], xy[2 * i + 1]};
  fit_line_l2(pts, line4);
}

}  // extern "C"
