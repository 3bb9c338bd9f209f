// Line-segment extraction kernels (sm_100a): Canny edge map of the half-resolution equalised frame and the
// FastLineDetector chain walk + incremental segment fit.
//
// Reference ops replaced (both inside cv::ximgproc::FastLineDetector::detect, called at TrackLSD.cpp:200-205 with
// length 20, distance 1.414213562, Canny 50/50 aperture 3, no merge — TrackLSD.h:269-273):
//   cv::Canny(src, 50, 50, 3, L1)    SURVEY.md Appendix A8: low == high, so every gradient-direction local
//                                    maximum above the threshold is an edge and no hysteresis walk is needed
//   lineDetection / getPointChain / extractSegments / additionalOperationsOnSegment    Appendix B
//
// Structure.  The chain walk is order dependent (visited pixels are consumed, seeds are taken in raster order),
// so it stays sequential per frame: ONE warp owns the whole bit-packed edge map in shared memory (640x280 bits =
// 22 KB), its lanes scan for the next seed together and lane 0 walks the chain with a 2 KB decision table.
// Frames/streams run concurrently on different SMs.  Everything after the walk is parallel over chains (one
// thread per chain fits segments with running double-precision sums, identical to refitting from scratch) and
// an ordered compaction restores the reference's output order.
#include "fe_kernels.h"

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

// ------------------------------------------------------------------------------------------- chain walk
// Decision table of getPointChain for step > 0: [direction + 3][8-neighbour mask] -> neighbour index, 8 = stop.
__constant__ uint8_t c_chain_lut[8 * 256];
static bool g_chain_lut_ready[64] = {false};

static void build_chain_lut(uint8_t *lut) {
  for (int d = -3; d <= 4; d++) {
    for (int mask = 0; mask < 256; mask++) {
      float min_dir_diff = 7.0f;
      int chosen = 8;
      for (int i = 0; i < 8; i++) {
        if (!((mask >> i) & 1)) continue;
        int curr_dir = i > 4 ? i - 8 : i;
        int dir_diff = std::abs(curr_dir - d);
        dir_diff = dir_diff > 4 ? 8 - dir_diff : dir_diff;
        if (dir_diff <= min_dir_diff) {
          min_dir_diff = (float)dir_diff;
          chosen = i;
        }
      }
      lut[(d + 3) * 256 + mask] = (uint8_t)(min_dir_diff < 2 ? chosen : 8);
    }
  }
}

// padded bitmap in shared memory: row stride ws words, 1 zero row above and below, x shifted by +1
__device__ __forceinline__ unsigned row3(const unsigned *bm, int ws, int py, int x) {
  // bits (x-1, x, x+1) of padded row py, as the low 3 bits (x in image coordinates, stored at bit x + 1)
  const unsigned *r = bm + py * ws;
  int wi = x >> 5, sh = x & 31;  // bit position of (x - 1) in padded coords is x
  return __funnelshift_r(r[wi], r[wi + 1], sh) & 7u;
}

__global__ void __launch_bounds__(32)
    k_fld_walk(const unsigned *__restrict__ edges, int words_per_row, int w, int h, int length_threshold,
               int2 *__restrict__ chain_pts, int *__restrict__ chain_off, int *__restrict__ n_chains, int max_chains) {
  extern __shared__ unsigned bm[];
  const int lane = threadIdx.x;
  const int ws = ((w + 2 + 31) >> 5) + 1;  // +1 so the funnel shift may read one word past the row
  const int prow = h + 2;
  for (int i = lane; i < prow * ws; i += 32) bm[i] = 0;
  __syncwarp();
  // copy with a 1-bit shift (padded x = x + 1)
  for (int i = lane; i < h * words_per_row; i += 32) {
    int y = i / words_per_row, wi = i - y * words_per_row;
    unsigned v = edges[i];
    if (wi == words_per_row - 1 && (w & 31)) v &= (1u << (w & 31)) - 1u;
    unsigned *dst = bm + (y + 1) * ws + wi;
    // lanes of one row touch neighbouring words: use shared atomics for the carry-over bit
    atomicOr(dst, v << 1);
    if (v >> 31) atomicOr(dst + 1, 1u);
  }
  __syncwarp();

  int nchains = 0, npts = 0;
  // raster scan over image words (padded row y + 1, padded bit x + 1)
  int y = 0, xw = 0;            // current row and 32-pixel group
  unsigned done_mask = 0;       // pixels of the current group already passed (bits below the last seed, inclusive)
  while (y < h) {
    // each lane inspects one 32-pixel group of the current row, starting at xw
    const int groups = (w + 31) >> 5;
    int g = xw + lane;
    unsigned v = 0;
    if (g < groups) {
      const unsigned *r = bm + (y + 1) * ws;
      v = __funnelshift_r(r[g], r[g + 1], 1);  // pixels 32g .. 32g+31
      if (g == xw) v &= ~done_mask;
    }
    unsigned any = __ballot_sync(0xffffffffu, v != 0);
    if (!any) {
      xw += 32;
      done_mask = 0;
      if (xw >= groups) { xw = 0; y++; }
      continue;
    }
    int first = __ffs(any) - 1;
    unsigned vv = __shfl_sync(0xffffffffu, v, first);
    int bit = __ffs(vv) - 1;
    int sx = ((xw + first) << 5) + bit, sy = y;
    if (first != 0) done_mask = 0;
    xw += first;
    done_mask |= (bit == 31) ? 0xffffffffu : ((2u << bit) - 1u);

    if (lane == 0) {
      // ---- walk one chain (getPointChain loop of lineDetection)
      int start = npts;
      int cx = sx, cy = sy;
      chain_pts[npts++] = make_int2(cx, cy);
      bm[(cy + 1) * ws + ((cx + 1) >> 5)] &= ~(1u << ((cx + 1) & 31));
      int direction = 0, step = 0;
      while (true) {
        unsigned t = row3(bm, ws, cy, cx), m = row3(bm, ws, cy + 1, cx), b = row3(bm, ws, cy + 2, cx);
        unsigned mask = ((b >> 2) & 1u) | (((b >> 1) & 1u) << 1) | ((b & 1u) << 2) | ((m & 1u) << 3) | ((t & 1u) << 4) |
                        (((t >> 1) & 1u) << 5) | (((t >> 2) & 1u) << 6) | (((m >> 2) & 1u) << 7);
        if (!mask) break;
        int i;
        if (step == 0) {
          i = __ffs(mask) - 1;
          direction = i > 4 ? i - 8 : i;
        } else {
          i = c_chain_lut[(direction + 3) * 256 + mask];
          if (i == 8) break;
          int cd = i > 4 ? i - 8 : i;
          direction = (direction * step + cd) / (step + 1);
        }
        const int dr = (i <= 2) ? 1 : ((i == 3 || i == 7) ? 0 : -1);
        const int dc = (i == 0 || i == 6 || i == 7) ? 1 : ((i == 1 || i == 5) ? 0 : -1);
        cx += dc;
        cy += dr;
        chain_pts[npts++] = make_int2(cx, cy);
        step++;
        bm[(cy + 1) * ws + ((cx + 1) >> 5)] &= ~(1u << ((cx + 1) & 31));
      }
      if (npts - start < length_threshold + 1 || nchains >= max_chains) {
        npts = start;  // chain too short: dropped (its pixels stay consumed)
      } else {
        chain_off[nchains++] = start;
      }
    }
    __syncwarp();
  }
  if (lane == 0) {
    chain_off[nchains] = npts;
    n_chains[0] = nchains;
  }
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

__global__ void k_fld_segments(const uint8_t *__restrict__ img, int W, int H, int pitch, int T, float dist_thr,
                               const int2 *__restrict__ chain_pts, const int *__restrict__ chain_off,
                               const int *__restrict__ n_chains, float4 *__restrict__ segs, int *__restrict__ seg_cnt,
                               int *__restrict__ seg_base) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains[0]) return;
  const int2 *points = chain_pts + chain_off[c];
  const int total = chain_off[c + 1] - chain_off[c];
  // segment slots of chain c start at chain_off[c] / 21 + c (every chain of n points yields <= n/21 + 1 segments)
  const int slot0 = chain_off[c] / kSegsPerChainDiv + c;
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
  seg_cnt[c] = nseg;
  seg_base[c] = slot0;
}

// ordered compaction of the per-chain segment slots (one block)
__global__ void __launch_bounds__(256)
    k_fld_compact(const float4 *__restrict__ segs, const int *__restrict__ seg_cnt, const int *__restrict__ seg_base,
                  int *__restrict__ n_chains, float4 *__restrict__ out, int out_cap) {
  __shared__ int warp_tot[8];
  __shared__ int s_running;
  const int n = n_chains[0];
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
  if (tid == 0) n_chains[1] = s_running;
}

void launch_fld(const DevImage &half, int length_threshold, float distance_threshold, FldBuffers &fb, cudaStream_t s) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (!g_chain_lut_ready[dev & 63]) {
    uint8_t lut[8 * 256];
    build_chain_lut(lut);
    cudaMemcpyToSymbol(c_chain_lut, lut, sizeof(lut));
    g_chain_lut_ready[dev & 63] = true;
  }
  const int ws = ((half.w + 2 + 31) >> 5) + 1;
  size_t smem = (size_t)(half.h + 2) * ws * sizeof(unsigned);
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaFuncSetAttribute(k_fld_walk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = smem;
  }
  k_fld_walk<<<1, 32, smem, s>>>(fb.edges, fb.words_per_row, half.w, half.h, length_threshold, fb.chain_pts, fb.chain_off,
                                 fb.n_chains, fb.max_chains);
  int blocks = (fb.max_chains + 63) / 64;
  k_fld_segments<<<blocks, 64, 0, s>>>(half.p, half.w, half.h, half.pitch, length_threshold, distance_threshold,
                                       fb.chain_pts, fb.chain_off, fb.n_chains, fb.segs, fb.seg_cnt, fb.seg_cnt + fb.max_chains);
  k_fld_compact<<<1, 256, 0, s>>>(fb.segs, fb.seg_cnt, fb.seg_cnt + fb.max_chains, fb.n_chains, fb.out, fb.out_cap);
}

}  // namespace plviwo
