// Host-side inner loops of the line tracker that are worth vectorising (compiled by g++ only).
#include <cstddef>
#include <cstdint>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace plviwo {

// AssignPointToLines (TrackLSD.cpp:770-781), candidate test for ONE line against all points, structure-of-arrays:
//   bit j = inside the (mis-indexed) bounding box of :775  AND  not farther than 6 px from the supporting line.
// The second test is a conservative pre-filter (a point more than 6 px from the infinite line is more than 5 px from
// the segment whatever branch PointLineDistance takes); survivors go through the exact function.  Comparing the
// float coordinate with a double that was converted from a float is the same as comparing the two floats.
// Result: one bit per point (bit j & 7 of bits[j >> 3]); returns non-zero if any point passed.  The caller pads px / py
// with 16 readable floats and bits with 2 writable bytes beyond n points.
static int line_candidates_scalar(const float *px, const float *py, int n, float min_lx, float max_lx, float min_ly, float max_ly,
                                  float pa, float pb, float pc, float plen2, uint8_t *bits) {
  int any = 0;
  for (int j0 = 0; j0 < n; j0 += 8) {
    unsigned m = 0;
    for (int k = 0; k < 8 && j0 + k < n; k++) {
      const float x = px[j0 + k], y = py[j0 + k];
      const float t = pa * x + pb * y + pc;
      const int ok = (!(x < min_lx) & !(x > max_lx) & !(y < min_ly) & !(y > max_ly)) & !(t * t > plen2);
      m |= (unsigned)ok << k;
    }
    bits[j0 >> 3] = (uint8_t)m;
    any |= (int)m;
  }
  return any;
}

#if defined(__x86_64__) && defined(__GNUC__)
// 8 points per iteration; the bounding-box test comes first and the distance test only runs for groups with a point inside
// the box (most groups have none).  Multiplications and additions stay separate instructions (no FMA), as in the scalar form.
__attribute__((target("avx2"))) static int line_candidates_avx2(const float *px, const float *py, int n, float min_lx,
                                                               float max_lx, float min_ly, float max_ly, float pa, float pb,
                                                               float pc, float plen2, uint8_t *bits) {
  const __m256 vminx = _mm256_set1_ps(min_lx), vmaxx = _mm256_set1_ps(max_lx), vminy = _mm256_set1_ps(min_ly),
               vmaxy = _mm256_set1_ps(max_ly), va = _mm256_set1_ps(pa), vb = _mm256_set1_ps(pb), vc = _mm256_set1_ps(pc),
               vl = _mm256_set1_ps(plen2);
  int any = 0;
  const int n8 = (n + 7) & ~7;   // the arrays are padded; padding points are masked out below
  for (int j = 0; j < n8; j += 8) {
    const __m256 x = _mm256_loadu_ps(px + j), y = _mm256_loadu_ps(py + j);
    // reject = x < min | x > max | y < min | y > max | t * t > plen2   (ordered comparisons: false on NaN, as the scalar form)
    __m256 rej = _mm256_or_ps(_mm256_cmp_ps(x, vminx, _CMP_LT_OQ), _mm256_cmp_ps(x, vmaxx, _CMP_GT_OQ));
    rej = _mm256_or_ps(rej, _mm256_or_ps(_mm256_cmp_ps(y, vminy, _CMP_LT_OQ), _mm256_cmp_ps(y, vmaxy, _CMP_GT_OQ)));
    unsigned m = ~(unsigned)_mm256_movemask_ps(rej) & 0xffu;
    if (m) {
      const __m256 t = _mm256_add_ps(_mm256_add_ps(_mm256_mul_ps(va, x), _mm256_mul_ps(vb, y)), vc);
      m &= ~(unsigned)_mm256_movemask_ps(_mm256_cmp_ps(_mm256_mul_ps(t, t), vl, _CMP_GT_OQ));
      if (j + 8 > n) m &= (1u << (n - j)) - 1u;
    }
    bits[j >> 3] = (uint8_t)m;
    any |= (int)m;
  }
  return any;
}

// the same, 16 points per iteration
__attribute__((target("avx512f"))) static int line_candidates_avx512(const float *px, const float *py, int n, float min_lx,
                                                                    float max_lx, float min_ly, float max_ly, float pa, float pb,
                                                                    float pc, float plen2, uint8_t *bits) {
  const __m512 vminx = _mm512_set1_ps(min_lx), vmaxx = _mm512_set1_ps(max_lx), vminy = _mm512_set1_ps(min_ly),
               vmaxy = _mm512_set1_ps(max_ly), va = _mm512_set1_ps(pa), vb = _mm512_set1_ps(pb), vc = _mm512_set1_ps(pc),
               vl = _mm512_set1_ps(plen2);
  int any = 0;
  const int n16 = (n + 15) & ~15;   // padded arrays (16 points of slack)
  for (int j = 0; j < n16; j += 16) {
    const __m512 x = _mm512_loadu_ps(px + j), y = _mm512_loadu_ps(py + j);
    const __mmask16 rej = _mm512_cmp_ps_mask(x, vminx, _CMP_LT_OQ) | _mm512_cmp_ps_mask(x, vmaxx, _CMP_GT_OQ) |
                          _mm512_cmp_ps_mask(y, vminy, _CMP_LT_OQ) | _mm512_cmp_ps_mask(y, vmaxy, _CMP_GT_OQ);
    unsigned m = ~(unsigned)rej & 0xffffu;
    if (m) {
      const __m512 t = _mm512_add_ps(_mm512_add_ps(_mm512_mul_ps(va, x), _mm512_mul_ps(vb, y)), vc);
      m &= ~(unsigned)_mm512_cmp_ps_mask(_mm512_mul_ps(t, t), vl, _CMP_GT_OQ);
      if (j + 16 > n) m &= n - j >= 16 ? 0xffffu : ((1u << (n - j)) - 1u);
    }
    bits[j >> 3] = (uint8_t)m;
    bits[(j >> 3) + 1] = (uint8_t)(m >> 8);
    any |= (int)m;
  }
  return any;
}
#endif

int line_candidates(const float *px, const float *py, int n, float min_lx, float max_lx, float min_ly, float max_ly, float pa,
                    float pb, float pc, float plen2, uint8_t *bits) {
#if defined(__x86_64__) && defined(__GNUC__)
  static const int level = __builtin_cpu_supports("avx512f") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
  if (level == 2) return line_candidates_avx512(px, py, n, min_lx, max_lx, min_ly, max_ly, pa, pb, pc, plen2, bits);
  if (level == 1) return line_candidates_avx2(px, py, n, min_lx, max_lx, min_ly, max_ly, pa, pb, pc, plen2, bits);
#endif
  return line_candidates_scalar(px, py, n, min_lx, max_lx, min_ly, max_ly, pa, pb, pc, plen2, bits);
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
