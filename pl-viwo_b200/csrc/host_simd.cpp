// Host-side inner loops of the line tracker that are worth vectorising (compiled by g++ only).
#include <cstddef>
#include <cstdint>

namespace plviwo {

// AssignPointToLines (TrackLSD.cpp:770-781), candidate test for ONE line against all points, structure-of-arrays:
//   pass[j] = inside the (mis-indexed) bounding box of :775  AND  not farther than 6 px from the supporting line.
// The second test is a conservative pre-filter (a point more than 6 px from the infinite line is more than 5 px from
// the segment whatever branch PointLineDistance takes); survivors go through the exact function.  Comparing the
// float coordinate with a double that was converted from a float is the same as comparing the two floats.
#if defined(__x86_64__) && defined(__GNUC__) && !defined(__clang__)
__attribute__((target_clones("avx2", "default")))
#endif
int line_candidates(const float *px, const float *py, int n, float min_lx, float max_lx, float min_ly, float max_ly, float pa,
                    float pb, float pc, float plen2, uint8_t *pass) {
  int any = 0;
  for (int j = 0; j < n; j++) {
    const float x = px[j], y = py[j];
    const float t = pa * x + pb * y + pc;
    const int ok = (!(x < min_lx) & !(x > max_lx) & !(y < min_ly) & !(y > max_ly)) & !(t * t > plen2);
    pass[j] = (uint8_t)ok;
    any |= ok;
  }
  return any;
}

// cv::pyrDown of an 8-bit image to (dw, dh) on the host: separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8.
// Only used for the tracking MASK when cfg.downsample is set (UpdaterCamera.cpp:92-93); masks never leave the host.
static inline int reflect101_host(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i < 0 ? 0 : (i >= n ? n - 1 : i);
}
void pyr_down_host(const uint8_t *src, int sw, int sh, int stride, uint8_t *dst, int dw, int dh) {
  // horizontal pass of the 5 source rows a destination row needs, kept in a ring of 5 int rows
  int *rows = new int[(size_t)5 * dw];
  int have[5] = {-1, -1, -1, -1, -1};   // source row currently held by each ring slot
  for (int y = 0; y < dh; y++) {
    const int *r[5];
    for (int k = 0; k < 5; k++) {
      const int sy = reflect101_host(2 * y + k - 2, sh);
      const int slot = sy % 5;
      int *row = rows + (size_t)slot * dw;
      if (have[slot] != sy) {
        const uint8_t *s = src + (size_t)sy * stride;
        for (int x = 0; x < dw; x++) {
          const int c = 2 * x;
          if (c >= 2 && c + 2 < sw)
            row[x] = s[c - 2] + 4 * s[c - 1] + 6 * s[c] + 4 * s[c + 1] + s[c + 2];
          else
            row[x] = s[reflect101_host(c - 2, sw)] + 4 * s[reflect101_host(c - 1, sw)] + 6 * s[reflect101_host(c, sw)] +
                     4 * s[reflect101_host(c + 1, sw)] + s[reflect101_host(c + 2, sw)];
        }
        have[slot] = sy;
      }
      r[k] = row;
    }
    uint8_t *d = dst + (size_t)y * dw;
    for (int x = 0; x < dw; x++) d[x] = (uint8_t)((r[0][x] + 4 * r[1][x] + 6 * r[2][x] + 4 * r[3][x] + r[4][x] + 128) >> 8);
  }
  delete[] rows;
}

}  // namespace plviwo
