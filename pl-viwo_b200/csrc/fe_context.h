// Host side of the front end: one FeContext per camera stream.  Mirrors the reference's tracker objects
// (ov_core::TrackKLT, viw::TrackLSD) with the same method names and the same per-frame state machine, but every
// pixel-touching step is a kernel launch on the context's own CUDA streams (fe_kernels.h).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "../../include/plviwo_fe.h"
#include "fe_kernels.h"

namespace plviwo {

struct Pt {
  float x, y;
};

int ransac_fundamental(const float *m1, const float *m2, int count, double threshold, double confidence, uint8_t *mask,
                       int *mask_valid);

// Everything that belongs to one submitted frame and can be produced without tracker state.
struct FrameSlot {
  DevImage raw;          // staged input (device)
  Pyramid pyr;           // equalised pyramid, level 0 = equalised frame
  DevImage half;         // half-resolution equalised frame (line detector input)
  FldBuffers fld;
  uint8_t *h_raw = nullptr;      // pinned staging for the host image
  float4 *h_segs = nullptr;      // pinned: detected segments (half-res)
  int *h_fld_counts = nullptr;   // pinned: [n_chains, n_segments]
  std::vector<uint8_t> mask;     // host copy of the caller's mask (empty = all zero)
  double timestamp = 0;
  double vp[6] = {0, 0, 0, 0, 0, 0};
  bool has_vp = false;
  bool busy = false;
  cudaEvent_t ev_pyr = nullptr, ev_lines = nullptr;
  cudaEvent_t ev_t[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool timed = false;
};

class FeContext {
 public:
  FeContext(const FeConfig &cfg, int device);
  ~FeContext();
  int init();  // allocates; returns FeStatus

  // ---- reference-shaped API ----
  int set_calib(const double K[4], const double D[4]);
  int submit(double t, const uint8_t *image, int stride, bool on_device, const uint8_t *mask, int mask_stride,
             const double vp[6]);
  int collect(FeFrameInfo *info);
  int feed(double t, const uint8_t *image, int w, int h, int stride, bool on_device, const uint8_t *mask,
           int mask_stride, const double vp[6], FeFrameInfo *info);

  // TrackBase::get_last_obs / get_last_ids
  const std::vector<Pt> &get_last_obs() const { return pts_last_; }
  const std::vector<uint64_t> &get_last_ids() const { return ids_last_; }
  void set_num_features(int n) { cfg_.num_features = n; }
  void change_feat_id(uint64_t id_old, uint64_t id_new);

  int get_state(void *buf, size_t cap, size_t *n_bytes);
  int set_state(const void *buf, size_t n_bytes);
  int tap(int what, void *buf, size_t cap, size_t *n_bytes);

  const FeConfig &cfg() const { return cfg_; }
  std::string last_error;
  std::vector<FePointRow> point_rows;
  std::vector<FeLineRow> line_rows;
  std::vector<FeLinePoint> line_points;
  std::vector<float> sample_uv;
  std::vector<uint8_t> sample_status;
  bool timing = false;
  FeStageTimes times{};

 private:
  int fail(cudaError_t e, const char *what);
  int alloc_image(DevImage &im, int w, int h);
  int enqueue_frame_independent(FrameSlot &s);
  // TrackKLT
  int klt_feed(FrameSlot &cur, FeFrameInfo *info);
  int perform_detection(const FrameSlot &img, std::vector<Pt> &pts, std::vector<uint64_t> &ids, FeFrameInfo *info);
  int perform_matching(const FrameSlot &f0, const FrameSlot &f1, std::vector<Pt> &pts0, std::vector<Pt> &pts1,
                       std::vector<uint8_t> &mask_out, bool &mask_empty, FeFrameInfo *info);
  // TrackLSD
  int lsd_feed(FrameSlot &cur, FeFrameInfo *info);
  void undistort_host(float u, float v, float &un, float &vn) const;

  FeConfig cfg_;
  int device_;
  int W_, H_;
  cudaStream_t s_img_ = nullptr, s_pt_ = nullptr, s_line_ = nullptr;
  std::vector<FrameSlot> slots_;
  std::vector<int> queue_;      // submitted, not yet collected (slot indices, FIFO)
  int last_slot_ = -1;          // slot holding the previous frame's pyramid (img_pyramid_last)
  unsigned *d_hist_ = nullptr, *d_counters_ = nullptr;

  // ---- point tracker state (TrackBase.h:173-192)
  std::vector<Pt> pts_last_;
  std::vector<uint64_t> ids_last_;
  uint64_t currid_ = 1;
  // ---- line tracker state (TrackLSD.h:248-279)
  std::vector<float4> lines_last_;
  std::vector<uint64_t> line_ids_last_;
  std::vector<std::map<int, double>> pol_last_;
  uint64_t line_currid_ = 1;

  // ---- detection scratch
  FastCell *d_cells_ = nullptr, *h_cells_ = nullptr;
  unsigned *d_fast_total_ = nullptr, *d_kps_ = nullptr, *h_kps_ = nullptr;
  int *d_band_off_ = nullptr, *d_band_cnt_ = nullptr, *h_band_ = nullptr;  // h_band_: [total, off..., cnt...]
  int max_cells_ = 0, max_bands_ = 0, kps_cap_ = 0;
  std::vector<uint64_t> occ_bits_;
  // ---- tracking scratch
  int max_pts_ = 0;
  float2 *d_pts0_ = nullptr, *d_pts1_ = nullptr, *d_p0n_ = nullptr, *d_p1n_ = nullptr;
  uint8_t *d_status_ = nullptr;
  float2 *h_pts0_ = nullptr, *h_pts1_ = nullptr, *h_p0n_ = nullptr, *h_p1n_ = nullptr;
  uint8_t *h_status_ = nullptr;
  cudaEvent_t ev_pt_[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // ---- taps
  std::vector<int32_t> tap_fast_;
  std::vector<float> tap_lk_, tap_subpix_, tap_fld_;
  int cur_slot_ = -1;
};

}  // namespace plviwo
