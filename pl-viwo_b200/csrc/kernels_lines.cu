// Line-segment extraction kernels (sm_100a): Canny edge map of the half-resolution equalised frame and the
// FastLineDetector chain walk + incremental segment fit.
//
// Reference ops replaced (both inside cv::ximgproc::FastLineDetector::detect, called at TrackLSD.cpp:200-205 with
// length 20, distance 1.414213562, Canny 50/50 aperture 3, no merge — TrackLSD.h:269-273):
//   cv::Canny(src, 50, 50, 3, L1)    SURVEY.md Appendix A8: low == high, so every gradient-direction local
//                                    maximum above the threshold is an edge and no hysteresis walk is needed
//   lineDetection / getPointChain / extractSegments / additionalOperationsOnSegment    Appendix B
//
// Structure.  The chain walk is order dependent (visited pixels are consumed, seeds are taken in raster order), but
// a walk never leaves the 8-connected component of its seed, so different components cannot influence each other:
// the sequential algorithm is exactly "for every connected component, run the raster-order walk on that component
// alone; then list all chains by the raster index of their seed".  So
//   1. connected components of the edge map by union-find (label = smallest raster index = the component's first seed),
//   2. one warp per component with >= length_threshold + 1 pixels: it copies the component's pixels into a private
//      bit map in shared memory, its lanes scan for the next seed together and lane 0 walks the chain,
//   3. chains are ranked by seed index, segments are fitted by one thread per chain (running double-precision sums,
//      identical to refitting from scratch) and an ordered compaction restores the reference's output order.
// The first version walked the whole frame with a single thread: 14.6 ms per 640x280 frame.
#include "fe_kernels.h"

#include <algorithm>
#include <cmath>

namespace plviwo {

// ------------------------------------------------------------------------------------------------ Canny
constexpr int kCnW = 64, kCnH = 16, kCnThreads = 256;

__global__ void __launch_bounds__(kCnThreads)
    k_canny(const uint8_t *__restrict__ img, int w, int h, int pitch, int low, unsigned *__restrict__ edges,
            int words_per_row) {
  __shared__ uint8_t pix[kCnH + 4][kCnW + 4];
  __shared__ short sdx[kCnH + 2][kCnW + 2];
  __shared__ short sdy[kCnH + 2][kCnW + 2];
  __shared__ unsigned short mag[kCnH + 2][kCnW + 2];
  const int tx0 = blockIdx.x * kCnW, ty0 = blockIdx.y * kCnH;
  const int tid = threadIdx.x;
  // pixels with a halo of 2, BORDER_REPLICATE
  for (int i = tid; i < (kCnH + 4) * (kCnW + 4); i += kCnThreads) {
    int r = i / (kCnW + 4), c = i - r * (kCnW + 4);
    int gx = min(max(tx0 - 2 + c, 0), w - 1), gy = min(max(ty0 - 2 + r, 0), h - 1);
    pix[r][c] = img[(size_t)gy * pitch + gx];
  }
  __syncthreads();
  // Sobel + L1 magnitude with a halo of 1; zero outside the image
  for (int i = tid; i < (kCnH + 2) * (kCnW + 2); i += kCnThreads) {
    int r = i / (kCnW + 2), c = i - r * (kCnW + 2);
    int gx = tx0 - 1 + c, gy = ty0 - 1 + r;
    int dx = 0, dy = 0, m = 0;
    if (gx >= 0 && gx < w && gy >= 0 && gy < h) {
      const uint8_t *p = &pix[r + 1][c + 1];
      const int s = kCnW + 4;
      dx = (p[-s + 1] + 2 * p[1] + p[s + 1]) - (p[-s - 1] + 2 * p[-1] + p[s - 1]);
      dy = (p[s - 1] + 2 * p[s] + p[s + 1]) - (p[-s - 1] + 2 * p[-s] + p[-s + 1]);
      m = abs(dx) + abs(dy);
    }
    sdx[r][c] = (short)dx;
    sdy[r][c] = (short)dy;
    mag[r][c] = (unsigned short)m;
  }
  __syncthreads();
  // non-maximum suppression; warp wv handles rows 2wv, 2wv+1; 32 pixels per ballot
  const int wv = tid >> 5, lane = tid & 31;
#pragma unroll
  for (int q = 0; q < 4; q++) {
    int r = 2 * wv + (q >> 1), c = (q & 1) * 32 + lane;
    int gx = tx0 + c, gy = ty0 + r;
    bool edge = false;
    if (gx < w && gy < h) {
      int m = mag[r + 1][c + 1];
      if (m > low) {
        int dx = sdx[r + 1][c + 1], dy = sdy[r + 1][c + 1];
        int x = abs(dx), y = abs(dy) << 15;
        int tg22x = x * 13573;
        if (y < tg22x) {
          edge = m > mag[r + 1][c] && m >= mag[r + 1][c + 2];
        } else {
          int tg67x = tg22x + (x << 16);
          if (y > tg67x) {
            edge = m > mag[r][c + 1] && m >= mag[r + 2][c + 1];
          } else {
            int s = (dx ^ dy) < 0 ? -1 : 1;
            edge = m > mag[r][c + 1 - s] && m > mag[r + 2][c + 1 + s];
          }
        }
      }
      // FastLineDetector clears the two corner blocks of the edge map before walking
      if (gy < 6 && gx < 6) edge = false;
      if (gy >= h - 5 && gx >= w - 5) edge = false;
    }
    unsigned bits = __ballot_sync(0xffffffffu, edge);
    if (lane == 0 && gy < h && (tx0 + (q & 1) * 32) < w) edges[(size_t)gy * words_per_row + ((tx0 + (q & 1) * 32) >> 5)] = bits;
  }
}

void launch_canny(const DevImage &half, float th_low, float th_high, FldBuffers &fb, cudaStream_t s) {
  (void)th_high;  // low == high is enforced at create time (no hysteresis pass is implemented)
  dim3 grid((half.w + kCnW - 1) / kCnW, (half.h + kCnH - 1) / kCnH);
  k_canny<<<grid, kCnThreads, 0, s>>>(half.p, half.w, half.h, half.pitch, (int)floorf(th_low), fb.edges, fb.words_per_row);
}

__global__ void k_unpack_edges(const unsigned *__restrict__ edges, int words_per_row, int w, int h,
                               uint8_t *__restrict__ out) {
  int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x < w && y < h) out[(size_t)y * w + x] = ((edges[(size_t)y * words_per_row + (x >> 5)] >> (x & 31)) & 1u) ? 255 : 0;
}
void launch_unpack_edges(const FldBuffers &fb, int w, int h, uint8_t *d_out, cudaStream_t s) {
  dim3 grid((w + 255) / 256, h);
  k_unpack_edges<<<grid, 256, 0, s>>>(fb.edges, fb.words_per_row, w, h, d_out);
}

// --------------------------------------------------------------------------- connected components (8-conn)
__device__ __forceinline__ int ccl_find(const int *parent, int i) {
  int p = __ldcg(parent + i);
  while (p != i) {
    i = p;
    p = __ldcg(parent + i);
  }
  return i;
}
__device__ __forceinline__ void ccl_union(int *parent, int a, int b) {
  while (true) {
    a = ccl_find(parent, a);
    b = ccl_find(parent, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }   // link the larger root under the smaller index
    int old = atomicMin(parent + a, b);
    if (old == a) return;
    a = old;
  }
}

__global__ void k_ccl_init(const unsigned *__restrict__ edges, int words_per_row, int w, int h, int *__restrict__ label,
                           int *__restrict__ cnt, int *__restrict__ bbox /* maxy, minx, maxx planes */,
                           int *__restrict__ counters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 8) counters[i] = 0;
  if (i >= w * h) return;
  const int y = i / w, x = i - y * w;
  const bool e = (edges[(size_t)y * words_per_row + (x >> 5)] >> (x & 31)) & 1u;
  label[i] = e ? i : -1;
  cnt[i] = 0;
  bbox[i] = 0;                 // max y
  bbox[w * h + i] = 0x7fffffff;  // min x
  bbox[2 * w * h + i] = -1;    // max x
}

__global__ void k_ccl_merge(int w, int h, int *__restrict__ label) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  if (label[i] < 0) return;
  const int y = i / w, x = i - y * w;
  // backward half of the 8-neighbourhood: W, NW, N, NE
  if (x > 0 && label[i - 1] >= 0) ccl_union(label, i, i - 1);
  if (y > 0) {
    if (x > 0 && label[i - w - 1] >= 0) ccl_union(label, i, i - w - 1);
    if (label[i - w] >= 0) ccl_union(label, i, i - w);
    if (x < w - 1 && label[i - w + 1] >= 0) ccl_union(label, i, i - w + 1);
  }
}

// flatten + per-component pixel count and bounding box (warp-aggregated atomics keyed by the root)
__global__ void k_ccl_flatten(int w, int h, int *__restrict__ label, int *__restrict__ cnt, int *__restrict__ bbox) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  int root = -1;
  if (i < w * h && label[i] >= 0) {
    root = ccl_find(label, i);
    label[i] = root;
  }
  const unsigned active = __ballot_sync(0xffffffffu, root >= 0);
  if (root < 0) return;
  const int y = i / w, x = i - y * w;
  const unsigned peers = __match_any_sync(active, root);
  const int n = __popc(peers);
  const int ymax = __reduce_max_sync(peers, y);
  const int xmin = __reduce_min_sync(peers, x);
  const int xmax = __reduce_max_sync(peers, x);
  if ((threadIdx.x & 31) == __ffs(peers) - 1) {
    atomicAdd(cnt + root, n);
    atomicMax(bbox + root, ymax);
    atomicMin(bbox + w * h + root, xmin);
    atomicMax(bbox + 2 * w * h + root, xmax);
  }
}

// components large enough to hold a chain of length_threshold + 1 pixels
__global__ void k_ccl_roots(int w, int h, const int *__restrict__ label, const int *__restrict__ cnt, int min_pixels,
                            int *__restrict__ comp_root, int *__restrict__ counters, int max_comps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= w * h) return;
  if (label[i] == i && cnt[i] >= min_pixels) {
    int k = atomicAdd(counters + 0, 1);
    if (k < max_comps) comp_root[k] = i;
  }
}

// ------------------------------------------------------------------------------------------- chain walk
// getPointChain for step > 0: among the set neighbours take the one whose direction is closest to the running
// direction (circular difference; ties go to the LATER neighbour index), accept only a difference < 2.
__device__ __forceinline__ int choose_neighbour(unsigned mask, int direction) {
  const int i0 = direction < 0 ? direction + 8 : direction;      // neighbour index with difference 0
  if ((mask >> i0) & 1u) return i0;
  const int ia = (i0 + 1) & 7, ib = (i0 + 7) & 7;               // the two neighbours with difference 1
  const bool sa = (mask >> ia) & 1u, sb = (mask >> ib) & 1u;
  if (sa && sb) return max(ia, ib);
  if (sa) return ia;
  if (sb) return ib;
  return 8;
}

// padded private bit map: row stride ws words, 1 zero row above and below, local x stored at bit x + 1
__device__ __forceinline__ unsigned row3(const unsigned *bm, int ws, int py, int x) {
  const unsigned *r = bm + py * ws;
  const int wi = x >> 5, sh = x & 31;
  return __funnelshift_r(r[wi], r[wi + 1], sh) & 7u;
}

// counters: [0] components, [1] next component to take, [2] chain-point cursor, [3] chains, [4] segments
constexpr int kWalkThreads = 128;
constexpr int kWalkCtas = 48;   // per frame; components are pulled from an atomic queue

__global__ void __launch_bounds__(kWalkThreads)
    k_fld_walk_cc(const int *__restrict__ label, const int *__restrict__ cnt, const int *__restrict__ bbox, int w, int h,
                  const int *__restrict__ comp_root, int *__restrict__ counters, int max_comps, int length_threshold,
                  int2 *__restrict__ chain_pts, int *__restrict__ chain_seed, int *__restrict__ chain_off,
                  int *__restrict__ chain_len, int max_chains) {
  extern __shared__ unsigned bm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int ncomp = min(counters[0], max_comps);
  __shared__ int s_ci;
  while (true) {
    // dynamic work queue: a few long-lived CTAs pull components, so the walk never floods the SMs' shared memory
    // while the latency-critical tracking kernels of the same or other camera streams want to start
    __syncthreads();   // the previous component's walk is finished
    if (tid == 0) s_ci = atomicAdd(counters + 1, 1);
    __syncthreads();
    const int ci = s_ci;
    if (ci >= ncomp) break;
    const int root = comp_root[ci];
    const int y0 = root / w, y1 = bbox[root];
    const int x0 = bbox[w * h + root], x1 = bbox[2 * w * h + root];
    const int bw = x1 - x0 + 1, bh = y1 - y0 + 1;
    const int ws = ((bw + 2 + 31) >> 5) + 1;
    for (int i = tid; i < (bh + 2) * ws; i += kWalkThreads) bm[i] = 0;
    __syncthreads();
    // private bit map of this component: one coalesced label load + ballot per 32 pixels, 4 loads in flight per warp
    const int groups = (bw + 31) >> 5;
    const int items = bh * groups;
    for (int it0 = warp * 4; it0 < items; it0 += (kWalkThreads / 32) * 4) {
      int lab[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int it = it0 + u;
        lab[u] = -2;
        if (it < items) {
          const int ly = it / groups, g = it - ly * groups;
          const int lx = (g << 5) + lane;
          if (lx < bw) lab[u] = label[(size_t)(y0 + ly) * w + x0 + lx];
        }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int it = it0 + u;
        const unsigned word = __ballot_sync(0xffffffffu, lab[u] == root);
        if (lane == 0 && word && it < items) {
          const int ly = it / groups, g = it - ly * groups;
          unsigned *dst = bm + (ly + 1) * ws + g;
          atomicOr(dst, word << 1);
          if (word >> 31) atomicOr(dst + 1, 1u);
        }
      }
    }
    __syncthreads();
    if (warp != 0) continue;   // warp 0 scans and walks; the others wait at the barrier above
    int base = 0;
    if (lane == 0) base = atomicAdd(counters + 2, cnt[root]);   // this component's slice of the chain-point pool
    base = __shfl_sync(0xffffffffu, base, 0);
    int npts = base;
    // raster scan over the private map
    int y = 0, xw = 0;
    while (y < bh) {
      const int g = xw + lane;
      unsigned v = 0;
      if (g < groups) {
        const unsigned *r = bm + (y + 1) * ws;
        v = __funnelshift_r(r[g], r[g + 1], 1);  // local pixels 32g .. 32g+31
      }
      const unsigned any = __ballot_sync(0xffffffffu, v != 0);
      if (!any) {
        xw += 32;
        if (xw >= groups) { xw = 0; y++; }
        continue;
      }
      const int first = __ffs(any) - 1;
      const unsigned vv = __shfl_sync(0xffffffffu, v, first);
      const int bit = __ffs(vv) - 1;
      const int sx = ((xw + first) << 5) + bit, sy = y;
      xw += first;
      if (lane == 0) {
        // ---- walk one chain (the getPointChain loop of lineDetection); coordinates local to the bounding box
        const int start = npts;
        int cx = sx, cy = sy;
        chain_pts[npts++] = make_int2(x0 + cx, y0 + cy);
        bm[(cy + 1) * ws + ((cx + 1) >> 5)] &= ~(1u << ((cx + 1) & 31));
        int direction = 0, step = 0;
        while (true) {
          const unsigned t = row3(bm, ws, cy, cx), m = row3(bm, ws, cy + 1, cx), b = row3(bm, ws, cy + 2, cx);
          const unsigned mask = ((b >> 2) & 1u) | (((b >> 1) & 1u) << 1) | ((b & 1u) << 2) | ((m & 1u) << 3) |
                                ((t & 1u) << 4) | (((t >> 1) & 1u) << 5) | (((t >> 2) & 1u) << 6) | (((m >> 2) & 1u) << 7);
          if (!mask) break;
          int i;
          if (step == 0) {
            i = __ffs(mask) - 1;
            direction = i > 4 ? i - 8 : i;
          } else {
            i = choose_neighbour(mask, direction);
            if (i == 8) break;
            const int cd = i > 4 ? i - 8 : i;
            direction = (direction * step + cd) / (step + 1);
          }
          const int dr = (i <= 2) ? 1 : ((i == 3 || i == 7) ? 0 : -1);
          const int dc = (i == 0 || i == 6 || i == 7) ? 1 : ((i == 1 || i == 5) ? 0 : -1);
          cx += dc;
          cy += dr;
          chain_pts[npts++] = make_int2(x0 + cx, y0 + cy);
          step++;
          bm[(cy + 1) * ws + ((cx + 1) >> 5)] &= ~(1u << ((cx + 1) & 31));
        }
        if (npts - start < length_threshold + 1) {
          npts = start;  // chain too short: dropped (its pixels stay consumed)
        } else {
          const int c = atomicAdd(counters + 3, 1);
          if (c < max_chains) {
            chain_seed[c] = (y0 + sy) * w + (x0 + sx);
            chain_off[c] = start;
            chain_len[c] = npts - start;
          }
        }
      }
      npts = __shfl_sync(0xffffffffu, npts, 0);
      __syncwarp();
    }
  }
}

// rank of every chain by the raster index of its seed (seeds are distinct pixels): order[rank] = chain
__global__ void k_fld_order(const int *__restrict__ chain_seed, const int *__restrict__ counters, int max_chains,
                            int *__restrict__ order) {
  const int n = min(counters[3], max_chains);
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n) return;
  const int mine = chain_seed[c];
  int rank = 0;
  for (int k = 0; k < n; k++) rank += chain_seed[k] < mine ? 1 : 0;
  order[rank] = c;
}

// --------------------------------------------------------------------------------- segments from chains
struct FitSums {
  double x, y, x2, y2, xy;
  int n;
};
__device__ __forceinline__ void fs_add(FitSums &s, int2 p) {
  float px = (float)p.x, py = (float)p.y;
  s.x += px; s.y += py;
  s.x2 += px * px; s.y2 += py * py; s.xy += px * py;
  s.n++;
}
// cv::fitLine(.., DIST_L2, 0, 0.01, 0.01) == fitLine2D_wods, then the homogeneous line through (x0,y0),(x0+vx,y0+vy)
__device__ __forceinline__ void fs_line(const FitSums &s, double l[3]) {
  double w = (float)s.n;
  double x = s.x / w, y = s.y / w, x2 = s.x2 / w, y2 = s.y2 / w, xy = s.xy / w;
  double dx2 = x2 - x * x, dy2 = y2 - y * y, dxy = xy - x * y;
  float t = (float)atan2(2 * dxy, dx2 - dy2) / 2;
  float l0 = cosf(t), l1 = sinf(t), l2 = (float)x, l3 = (float)y;
  double a[3] = {l2, l3, 1.0};
  double b[3] = {(double)(l2 + l0), (double)(l3 + l1), 1.0};
  l[0] = a[1] * b[2] - a[2] * b[1];
  l[1] = a[2] * b[0] - a[0] * b[2];
  l[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ double dist_point_line(double px, double py, double l[3]) {
  double x = l[0], y = l[1];
  double w = sqrt(x * x + y * y);
  l[0] = x / w; l[1] = y / w; l[2] = l[2] / w;
  return l[0] * px + l[1] * py + l[2];
}
__device__ __forceinline__ void incident_point(const double l[3], float &px, float &py, int W, int H) {
  double a[3] = {(double)px, (double)py, 1.0};
  double b[3] = {l[0], l[1], 0.0};
  double lk[3] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]};
  double xk[3] = {lk[1] * l[2] - lk[2] * l[1], lk[2] * l[0] - lk[0] * l[2], lk[0] * l[1] - lk[1] * l[0]};
  double alpha = 1.0 / xk[2];
  float fx = (float)(xk[0] * alpha), fy = (float)(xk[1] * alpha);
  px = fx < 0.0f ? 0.0f : (fx >= (W - 1.0f) ? (W - 1.0f) : fx);
  py = fy < 0.0f ? 0.0f : (fy >= (H - 1.0f) ? (H - 1.0f) : fy);
}

// one thread per chain, in seed order; chain c writes its segments to slots chain_off[c] / 21 + j (collision free:
// every segment consumes at least 21 chain points and the chains' point ranges are disjoint)
__global__ void k_fld_segments(const uint8_t *__restrict__ img, int W, int H, int pitch, int T, float dist_thr,
                               const int2 *__restrict__ chain_pts, const int *__restrict__ chain_off,
                               const int *__restrict__ chain_len, const int *__restrict__ order,
                               const int *__restrict__ counters, int max_chains, float4 *__restrict__ segs,
                               int *__restrict__ seg_cnt, int *__restrict__ seg_base) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= min(counters[3], max_chains)) return;
  const int c = order[r];
  const int2 *points = chain_pts + chain_off[c];
  const int total = chain_len[c];
  const int slot0 = chain_off[c] / kSegsPerChainDiv;
  int nseg = 0;
  int i, j;
  for (i = 0; i + T < total; i++) {
    int2 ps = points[i], pe = points[i + T];
    double l[3] = {(double)ps.y - (double)pe.y, (double)pe.x - (double)ps.x,
                   (double)ps.x * (double)pe.y - (double)ps.y * (double)pe.x};
    bool is_line = true;
    FitSums fs = {0, 0, 0, 0, 0, 0};
    fs_add(fs, ps);
    for (j = 1; j < T; j++) {
      int2 pt = points[i + j];
      double dist = dist_point_line((double)pt.x, (double)pt.y, l);
      if (fabs(dist) > dist_thr) { is_line = false; break; }
      fs_add(fs, pt);
    }
    if (!is_line) continue;
    fs_add(fs, pe);
    fs_line(fs, l);
    {  // incidentPoint(l, ps) on the INTEGER point: result rounded (cv::Point2i(Point2f) == cvRound)
      float fx = (float)ps.x, fy = (float)ps.y;
      incident_point(l, fx, fy, W, H);
      ps.x = __float2int_rn(fx);
      ps.y = __float2int_rn(fy);
    }
    for (j = T + 1; i + j < total; j++) {
      int2 pt = points[i + j];
      double dist = dist_point_line((double)pt.x, (double)pt.y, l);
      if (fabs(dist) > dist_thr) {
        fs_line(fs, l);
        dist = dist_point_line((double)pt.x, (double)pt.y, l);
        if (fabs(dist) > dist_thr) { j--; break; }
      }
      pe = pt;
      fs_add(fs, pt);
    }
    fs_line(fs, l);
    float e1x = (float)ps.x, e1y = (float)ps.y, e2x = (float)pe.x, e2y = (float)pe.y;
    incident_point(l, e1x, e1y, W, H);
    incident_point(l, e2x, e2y, W, H);
    i = i + j;
    // ---- per-segment filters of lineDetection
    float len = sqrtf((e1x - e2x) * (e1x - e2x) + (e1y - e2y) * (e1y - e2y));
    if (len < (float)T) continue;
    if ((e1x <= 5.0f && e2x <= 5.0f) || (e1y <= 5.0f && e2y <= 5.0f) || (e1x >= W - 5.0f && e2x >= W - 5.0f) ||
        (e1y >= H - 5.0f && e2y >= H - 5.0f))
      continue;
    // ---- additionalOperationsOnSegment: orient by the brighter side
    if (!(e1x == 0.0f && e2x == 0.0f && e1y == 0.0f && e2y == 0.0f)) {
      double ang = atan2((double)(e2y - e1y), (double)(e2x - e1x));
      double dx = (double)e2x - (double)e1x, dy = (double)e2y - (double)e1y;
      const double kPi = 3.1415926535897932384626433832795;
      double cs = cos(90.0 * kPi / 180.0 + ang), sn = sin(90.0 * kPi / 180.0 + ang);
      int iR = 0, iL = 0;
      for (int k = 0; k < 10; k++) {
        float qx, qy;
        if (k == 0) { qx = e1x; qy = e1y; }
        else if (k == 9) { qx = e2x; qy = e2y; }
        else {
          qx = e1x + ((float)dx / 9.0f * (float)k);
          qy = e1y + ((float)dy / 9.0f * (float)k);
        }
        int rx = __double2int_rn(qx + cs), ry = __double2int_rn(qy + sn);
        int lx = __double2int_rn(qx - cs), ly = __double2int_rn(qy - sn);
        rx = min(max(rx, 0), W - 1); ry = min(max(ry, 0), H - 1);
        lx = min(max(lx, 0), W - 1); ly = min(max(ly, 0), H - 1);
        iR += img[(size_t)ry * pitch + rx];
        iL += img[(size_t)ly * pitch + lx];
      }
      if (iR > iL) {
        float tx = e1x, ty = e1y;
        e1x = e2x; e1y = e2y; e2x = tx; e2y = ty;
      }
    }
    segs[slot0 + nseg] = make_float4(e1x, e1y, e2x, e2y);
    nseg++;
  }
  seg_cnt[r] = nseg;      // indexed by RANK so the compaction below walks chains in seed order
  seg_base[r] = slot0;
}

// ordered compaction of the per-chain segment slots (one block)
__global__ void __launch_bounds__(256)
    k_fld_compact(const float4 *__restrict__ segs, const int *__restrict__ seg_cnt, const int *__restrict__ seg_base,
                  int *__restrict__ counters, int max_chains, float4 *__restrict__ out, int out_cap) {
  __shared__ int warp_tot[8];
  __shared__ int s_running;
  const int n = min(counters[3], max_chains);
  const int tid = threadIdx.x;
  if (tid == 0) s_running = 0;
  __syncthreads();
  for (int base = 0; base < n; base += 256) {
    int c = base + tid;
    int cnt = c < n ? seg_cnt[c] : 0;
    int v = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, v, d);
      if ((tid & 31) >= d) v += t;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = v;
    __syncthreads();
    int off = s_running;
    for (int k = 0; k < (tid >> 5); k++) off += warp_tot[k];
    off += v - cnt;
    for (int k = 0; k < cnt; k++)
      if (off + k < out_cap) out[off + k] = segs[seg_base[c] + k];
    __syncthreads();
    if (tid == 255) s_running = off + cnt;
    __syncthreads();
  }
  if (tid == 0) counters[4] = s_running;
}

int FldBuffers::alloc(int w, int h, int length_threshold, int out_capacity) {
  const size_t n = (size_t)w * h;
  words_per_row = (w + 31) / 32;
  max_chains = (int)(n / (length_threshold + 1)) + 1;
  out_cap = out_capacity;
  bool ok = true;
  auto A = [&](auto **p, size_t bytes) { ok = ok && cudaMalloc((void **)p, bytes) == cudaSuccess; };
  A(&edges, (size_t)words_per_row * h * sizeof(unsigned));
  A(&label, n * sizeof(int));
  A(&cnt, n * sizeof(int));
  A(&bbox, 3 * n * sizeof(int));
  A(&comp_root, (size_t)max_chains * sizeof(int));
  A(&counters, 8 * sizeof(int));
  A(&chain_pts, n * sizeof(int2));
  A(&chain_seed, (size_t)max_chains * sizeof(int));
  A(&chain_off, (size_t)max_chains * sizeof(int));
  A(&chain_len, (size_t)max_chains * sizeof(int));
  A(&order, (size_t)max_chains * sizeof(int));
  A(&segs, (n / kSegsPerChainDiv + 2) * sizeof(float4));
  A(&seg_cnt, (size_t)2 * max_chains * sizeof(int));
  A(&out, (size_t)out_cap * sizeof(float4));
  if (ok) ok = cudaMemset(counters, 0, 8 * sizeof(int)) == cudaSuccess;
  return ok ? 0 : 1;
}

void FldBuffers::release() {
  cudaFree(edges); cudaFree(label); cudaFree(cnt); cudaFree(bbox); cudaFree(comp_root); cudaFree(counters);
  cudaFree(chain_pts); cudaFree(chain_seed); cudaFree(chain_off); cudaFree(chain_len); cudaFree(order);
  cudaFree(segs); cudaFree(seg_cnt); cudaFree(out);
  *this = FldBuffers();
}

void launch_fld(const DevImage &half, int length_threshold, float distance_threshold, FldBuffers &fb, cudaStream_t s) {
  const int w = half.w, h = half.h, n = w * h;
  const int tpb = 256, nb = (n + tpb - 1) / tpb;
  k_ccl_init<<<nb, tpb, 0, s>>>(fb.edges, fb.words_per_row, w, h, fb.label, fb.cnt, fb.bbox, fb.counters);
  k_ccl_merge<<<nb, tpb, 0, s>>>(w, h, fb.label);
  k_ccl_flatten<<<nb, tpb, 0, s>>>(w, h, fb.label, fb.cnt, fb.bbox);
  k_ccl_roots<<<nb, tpb, 0, s>>>(w, h, fb.label, fb.cnt, length_threshold + 1, fb.comp_root, fb.counters, fb.max_chains);
  const int ws = ((w + 2 + 31) >> 5) + 1;
  size_t smem = (size_t)(h + 2) * ws * sizeof(unsigned);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(k_fld_walk_cc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  k_fld_walk_cc<<<kWalkCtas, kWalkThreads, smem, s>>>(fb.label, fb.cnt, fb.bbox, w, h, fb.comp_root, fb.counters, fb.max_chains,
                                               length_threshold, fb.chain_pts, fb.chain_seed, fb.chain_off, fb.chain_len,
                                               fb.max_chains);
  int blocks = (fb.max_chains + 127) / 128;
  k_fld_order<<<blocks, 128, 0, s>>>(fb.chain_seed, fb.counters, fb.max_chains, fb.order);
  k_fld_segments<<<(fb.max_chains + 63) / 64, 64, 0, s>>>(half.p, w, h, half.pitch, length_threshold, distance_threshold,
                                                          fb.chain_pts, fb.chain_off, fb.chain_len, fb.order, fb.counters,
                                                          fb.max_chains, fb.segs, fb.seg_cnt, fb.seg_cnt + fb.max_chains);
  k_fld_compact<<<1, 256, 0, s>>>(fb.segs, fb.seg_cnt, fb.seg_cnt + fb.max_chains, fb.counters, fb.max_chains, fb.out,
                                  fb.out_cap);
}

}  // namespace plviwo
