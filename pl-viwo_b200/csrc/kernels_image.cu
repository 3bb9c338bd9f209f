// Image-domain kernels of the front end (sm_100a): histogram, equalise + pyramid level 1 + half-resolution
// image in one pass, and the remaining pyramid levels in one launch.
//
// Reference ops replaced (all are OpenCV calls made by the reference; arithmetic per SURVEY.md Appendix A):
//   cv::equalizeHist                    TrackKLT.cpp:59, TrackLSD.cpp:83   (A1)  -> k_hist + k_eq_pyr1 (LUT)
//   cv::buildOpticalFlowPyramid         TrackKLT.cpp:71                    (A2)  -> k_eq_pyr1 (level 1) + k_pyr_rest
//   cv::resize(.., 0.5, 0.5, LINEAR)    TrackLSD.cpp:204                   (A7)  -> k_eq_pyr1 (half)
// The reference equalises the same frame twice (point and line tracker) and materialises Scharr derivative
// planes with a 15 px border; here the frame is read once, the equalised image is written once and shared,
// and derivatives are formed on the fly inside the LK kernel.
#include "fe_kernels.h"

namespace plviwo {

__device__ __forceinline__ int reflect101(int i, int n) {
  // BORDER_REFLECT_101, valid for -n < i < 2n - 1
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

// ------------------------------------------------------------------------------------------------ histogram
// Each thread walks 16-byte chunks (uint4 = 16 pixels); per-warp shared sub-histograms keep shared-memory
// atomics off a single copy, one global atomic per non-empty bin per CTA at the end.
constexpr int kHistThreads = 256;
constexpr int kHistWarps = kHistThreads / 32;

__global__ void __launch_bounds__(kHistThreads) k_hist(const uint8_t *__restrict__ src, int w, int h, int pitch,
                                                       unsigned *__restrict__ hist) {
  __shared__ unsigned sh[kHistWarps][256];
  for (int i = threadIdx.x; i < kHistWarps * 256; i += kHistThreads) (&sh[0][0])[i] = 0;
  __syncthreads();
  unsigned *my = sh[threadIdx.x >> 5];
  const int chunks_per_row = (w + 15) >> 4;
  const int total = chunks_per_row * h;
  for (int c = blockIdx.x * kHistThreads + threadIdx.x; c < total; c += gridDim.x * kHistThreads) {
    int y = c / chunks_per_row;
    int x = (c - y * chunks_per_row) << 4;
    const uint8_t *row = src + (size_t)y * pitch + x;
    if (x + 16 <= w) {
      uint4 v = *reinterpret_cast<const uint4 *>(row);
      unsigned wds[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int k = 0; k < 4; k++) {
        atomicAdd(&my[wds[k] & 0xff], 1u);
        atomicAdd(&my[(wds[k] >> 8) & 0xff], 1u);
        atomicAdd(&my[(wds[k] >> 16) & 0xff], 1u);
        atomicAdd(&my[wds[k] >> 24], 1u);
      }
    } else {
      for (int k = 0; x + k < w; k++) atomicAdd(&my[row[k]], 1u);
    }
  }
  __syncthreads();
  for (int b = threadIdx.x; b < 256; b += kHistThreads) {
    unsigned s = 0;
#pragma unroll
    for (int k = 0; k < kHistWarps; k++) s += sh[k][b];
    if (s) atomicAdd(&hist[b], s);
  }
}

void launch_hist(const DevImage &src, unsigned *d_hist, cudaStream_t s) {
  int chunks = ((src.w + 15) >> 4) * src.h;
  int grid = (chunks + kHistThreads * 2 - 1) / (kHistThreads * 2);  // ~2 chunks per thread
  if (grid < 1) grid = 1;
  if (grid > 148 * 4) grid = 148 * 4;
  k_hist<<<grid, kHistThreads, 0, s>>>(src.p, src.w, src.h, src.pitch, d_hist);
}

// ------------------------------------------------------------------------- equalise + level 0/1 + half-res
// One CTA produces a 64 x 16 tile of level 1, i.e. consumes a (128 + 4) x (32 + 4) window of the raw frame
// (5-tap [1 4 6 4 1] pyrDown halo of 2), staged in shared memory AFTER the LUT so level 0, level 1 and the
// 2x2-mean half-resolution image all come from one read of the frame.
constexpr int kT1W = 64, kT1H = 16;            // level-1 tile
constexpr int kT0W = 2 * kT1W, kT0H = 2 * kT1H;  // level-0 interior of the tile
constexpr int kTileCols = kT0W + 8;            // staged columns: [2*tx0 - 4, 2*tx0 + 132), word aligned
constexpr int kTileRows = kT0H + 4;            // staged rows:    [2*ty0 - 2, 2*ty0 + 34)
constexpr int kEqThreads = 256;

__global__ void __launch_bounds__(kEqThreads)
    k_eq_pyr1(const uint8_t *__restrict__ src, int w, int h, int spitch, unsigned *__restrict__ hist,
              unsigned *__restrict__ counter, int equalize, uint8_t *__restrict__ l0, int l0pitch,
              uint8_t *__restrict__ l1, int w1, int h1, int l1pitch, uint8_t *__restrict__ half, int wh, int hh,
              int hpitch) {
  __shared__ __align__(16) uint8_t tile[kTileRows][kTileCols];
  __shared__ unsigned short hbuf[kTileRows][kT1W];
  __shared__ uint8_t lut[256];
  __shared__ unsigned warp_tot[kEqThreads / 32];
  __shared__ int s_i0;
  __shared__ unsigned s_h0;
  __shared__ unsigned s_last;

  const int tid = threadIdx.x;
  // ---- LUT (cv::equalizeHist, Appendix A1): every CTA rebuilds it from the 1 KB histogram
  if (equalize) {
    if (tid == 0) s_i0 = 256;
    __syncthreads();
    unsigned hv = hist[tid];
    if (hv) atomicMin(&s_i0, tid);
    // inclusive scan over 256 bins
    unsigned v = hv;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      unsigned t = __shfl_up_sync(0xffffffffu, v, d);
      if ((tid & 31) >= d) v += t;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = v;
    __syncthreads();
    unsigned base = 0;
    for (int k = 0; k < (tid >> 5); k++) base += warp_tot[k];
    unsigned cum = v + base;  // sum hist[0..tid]
    int i0 = s_i0;
    unsigned total = (unsigned)w * (unsigned)h;
    if (tid == i0) s_h0 = hv;
    __syncthreads();
    unsigned hist0 = s_h0;
    uint8_t out;
    if (hist0 == total) {
      out = (uint8_t)i0;  // single-valued image: dst.setTo(i0)
    } else if (tid <= i0) {
      out = 0;
    } else {
      float scale = __fdiv_rn(255.f, (float)(total - hist0));
      int sum = (int)(cum - hist0);  // bins i0+1..tid (bins below i0 are empty)
      int r = __float2int_rn(__fmul_rn((float)sum, scale));
      out = (uint8_t)min(max(r, 0), 255);
    }
    lut[tid] = out;
  } else {
    lut[tid] = (uint8_t)tid;
  }
  __syncthreads();

  // ---- stage the raw window through the LUT
  const int tx0 = blockIdx.x * kT1W, ty0 = blockIdx.y * kT1H;
  const int gx_base = 2 * tx0 - 4, gy_base = 2 * ty0 - 2;
  constexpr int kWordsPerRow = kTileCols / 4;
  for (int i = tid; i < kTileRows * kWordsPerRow; i += kEqThreads) {
    int r = i / kWordsPerRow, wi = i - r * kWordsPerRow;
    int gy = gy_base + r, gx = gx_base + 4 * wi;
    unsigned out = 0;
    if (gy > -h && gy < 2 * h - 1) {
      int ry = reflect101(gy, h);
      const uint8_t *row = src + (size_t)ry * spitch;
      if (gx >= 0 && gx + 4 <= w) {
        unsigned v = *reinterpret_cast<const unsigned *>(row + gx);
        out = (unsigned)lut[v & 0xff] | ((unsigned)lut[(v >> 8) & 0xff] << 8) | ((unsigned)lut[(v >> 16) & 0xff] << 16) |
              ((unsigned)lut[v >> 24] << 24);
      } else {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          int x = gx + k;
          if (x > -w && x < 2 * w - 1) out |= (unsigned)lut[row[reflect101(x, w)]] << (8 * k);
        }
      }
    }
    *reinterpret_cast<unsigned *>(&tile[r][4 * wi]) = out;
  }
  __syncthreads();

  // ---- level 0 (equalised frame): interior words of the tile
  for (int i = tid; i < kT0H * (kT0W / 4); i += kEqThreads) {
    int r = i / (kT0W / 4), wi = i - r * (kT0W / 4);
    int gy = 2 * ty0 + r, gx = 2 * tx0 + 4 * wi;
    if (gy < h && gx < w) {
      unsigned v = *reinterpret_cast<const unsigned *>(&tile[r + 2][4 + 4 * wi]);
      uint8_t *dst = l0 + (size_t)gy * l0pitch + gx;
      if (gx + 4 <= w) {
        *reinterpret_cast<unsigned *>(dst) = v;
      } else {
        for (int k = 0; gx + k < w; k++) dst[k] = (uint8_t)(v >> (8 * k));
      }
    }
  }

  // ---- horizontal [1 4 6 4 1] on all staged rows
  for (int i = tid; i < kTileRows * kT1W; i += kEqThreads) {
    int r = i / kT1W, c = i - r * kT1W;
    const uint8_t *p = &tile[r][2 * c + 2];
    hbuf[r][c] = (unsigned short)(p[0] + 4 * p[1] + 6 * p[2] + 4 * p[3] + p[4]);
  }
  __syncthreads();

  // ---- vertical pass -> level 1, and the 2x2 mean -> half-resolution image; 4 outputs per thread, one word store
  {
    int r = tid / (kT1W / 4), c4 = (tid - r * (kT1W / 4)) * 4;
    int oy = ty0 + r, ox = tx0 + c4;
    if (oy < h1 && ox < w1) {
      unsigned packed = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        int c = c4 + k;
        int s = hbuf[2 * r][c] + 4 * hbuf[2 * r + 1][c] + 6 * hbuf[2 * r + 2][c] + 4 * hbuf[2 * r + 3][c] + hbuf[2 * r + 4][c];
        packed |= (unsigned)((s + 128) >> 8) << (8 * k);
      }
      uint8_t *dst = l1 + (size_t)oy * l1pitch + ox;
      if (ox + 4 <= w1) {
        *reinterpret_cast<unsigned *>(dst) = packed;
      } else {
        for (int k = 0; ox + k < w1; k++) dst[k] = (uint8_t)(packed >> (8 * k));
      }
    }
    if (half != nullptr && oy < hh && ox < wh) {
      unsigned packed = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const uint8_t *p0 = &tile[2 * r + 2][2 * (c4 + k) + 4];
        const uint8_t *p1 = &tile[2 * r + 3][2 * (c4 + k) + 4];
        packed |= (unsigned)((p0[0] + p0[1] + p1[0] + p1[1] + 2) >> 2) << (8 * k);
      }
      uint8_t *dst = half + (size_t)oy * hpitch + ox;
      if (ox + 4 <= wh) {
        *reinterpret_cast<unsigned *>(dst) = packed;
      } else {
        for (int k = 0; ox + k < wh; k++) dst[k] = (uint8_t)(packed >> (8 * k));
      }
    }
  }

  // ---- the last CTA to finish clears the histogram for the next frame
  if (equalize) {
    __syncthreads();
    if (tid == 0) {
      __threadfence();
      s_last = (atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1) ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {
      hist[tid] = 0;
      if (tid == 0) *counter = 0;
    }
  }
}

void launch_eq_pyr1(const DevImage &src, unsigned *d_hist, unsigned *d_counter, int equalize, const DevImage &l0,
                    const DevImage &l1, const DevImage &half, cudaStream_t s) {
  // tiles are laid over the level-1 footprint of the frame even when level 1 itself is not wanted (l1.p == null)
  const int w1 = (src.w + 1) / 2, h1 = (src.h + 1) / 2;
  dim3 grid((w1 + kT1W - 1) / kT1W, (h1 + kT1H - 1) / kT1H);
  k_eq_pyr1<<<grid, kEqThreads, 0, s>>>(src.p, src.w, src.h, src.pitch, d_hist, d_counter, equalize, l0.p, l0.pitch, l1.p,
                                        l1.p ? l1.w : 0, l1.p ? l1.h : 0, l1.pitch, half.p, half.w, half.h, half.pitch);
}

// ---------------------------------------------------------------------------------------- remaining levels
__device__ __forceinline__ int pyr_down_px(const uint8_t *__restrict__ src, int sw, int sh, int spitch, int x, int y,
                                           bool cg) {
  int acc = 0;
  const int wgt[5] = {1, 4, 6, 4, 1};
#pragma unroll
  for (int j = 0; j < 5; j++) {
    int ry = reflect101(2 * y + j - 2, sh);
    const uint8_t *row = src + (size_t)ry * spitch;
    int hsum = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
      int rx = reflect101(2 * x + i - 2, sw);
      int v = cg ? (int)__ldcg(row + rx) : (int)row[rx];
      hsum += wgt[i] * v;
    }
    acc += wgt[j] * hsum;
  }
  return (acc + 128) >> 8;
}

struct PyrRestArgs {
  uint8_t *p[kMaxLevels];
  int w[kMaxLevels], h[kMaxLevels], pitch[kMaxLevels];
  int n;
};

constexpr int kRestThreads = 256;
constexpr int kRestTileW = 32, kRestTileH = 8;

__global__ void __launch_bounds__(kRestThreads) k_pyr_rest(PyrRestArgs a, unsigned *__restrict__ counter) {
  __shared__ unsigned s_last;
  {  // level 2 from level 1: one output per thread
    int x = blockIdx.x * kRestTileW + (threadIdx.x & (kRestTileW - 1));
    int y = blockIdx.y * kRestTileH + (threadIdx.x / kRestTileW);
    if (x < a.w[2] && y < a.h[2]) a.p[2][(size_t)y * a.pitch[2] + x] = (uint8_t)pyr_down_px(a.p[1], a.w[1], a.h[1], a.pitch[1], x, y, false);
  }
  if (a.n <= 3) return;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int l = 3; l < a.n; l++) {
    int n = a.w[l] * a.h[l];
    for (int i = threadIdx.x; i < n; i += kRestThreads) {
      int y = i / a.w[l], x = i - y * a.w[l];
      a.p[l][(size_t)y * a.pitch[l] + x] = (uint8_t)pyr_down_px(a.p[l - 1], a.w[l - 1], a.h[l - 1], a.pitch[l - 1], x, y, true);
    }
    __threadfence();
    __syncthreads();
  }
  if (threadIdx.x == 0) *counter = 0;
}

void launch_pyr_rest(const Pyramid &pyr, unsigned *d_counter, cudaStream_t s) {
  if (pyr.n <= 2) return;
  PyrRestArgs a;
  a.n = pyr.n;
  for (int l = 0; l < pyr.n; l++) {
    a.p[l] = pyr.lvl[l].p;
    a.w[l] = pyr.lvl[l].w;
    a.h[l] = pyr.lvl[l].h;
    a.pitch[l] = pyr.lvl[l].pitch;
  }
  dim3 grid((a.w[2] + kRestTileW - 1) / kRestTileW, (a.h[2] + kRestTileH - 1) / kRestTileH);
  k_pyr_rest<<<grid, kRestThreads, 0, s>>>(a, d_counter);
}

}  // namespace plviwo
