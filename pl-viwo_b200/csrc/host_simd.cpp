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

}  // namespace plviwo
