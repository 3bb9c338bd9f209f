#pragma once
#include <atomic>
#include <map>
#include <memory>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <opencv2/core.hpp>
#include "cam/CamBase.h"
#include "feat/FeatureDatabase.h"
#include "utils/sensor_data.h"
namespace ov_core {
class TrackBase {
 public:
  enum HistogramMethod { NONE, HISTOGRAM, CLAHE };
  TrackBase(std::unordered_map<size_t, std::shared_ptr<CamBase>> cameras, int numfeats, int numaruco, bool stereo, HistogramMethod histmethod)
      : camera_calib(cameras), database(new FeatureDatabase()), num_features(numfeats), use_stereo(stereo), histogram_method(histmethod) {
    currid = 4 * (size_t)numaruco + 1;
  }
  virtual ~TrackBase() {}
  virtual void feed_new_camera(const CameraData &message) = 0;
  std::shared_ptr<FeatureDatabase> get_feature_database() { return database; }
  void change_feat_id(size_t id_old, size_t id_new) { database->change_feat_id(id_old, id_new); }   // non-virtual (TrackBase.h)
 protected:
  std::unordered_map<size_t, std::shared_ptr<CamBase>> camera_calib;
  std::shared_ptr<FeatureDatabase> database;
  int num_features;
  bool use_stereo;
  HistogramMethod histogram_method;
  std::mutex mtx_last_vars;
  std::map<size_t, cv::Mat> img_last, img_mask_last;
  std::unordered_map<size_t, std::vector<cv::KeyPoint>> pts_last;
  std::unordered_map<size_t, std::vector<size_t>> ids_last;
  std::atomic<size_t> currid;
};
}  // namespace ov_core
