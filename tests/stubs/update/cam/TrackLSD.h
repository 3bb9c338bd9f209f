#pragma once
#include <map>
#include <memory>
#include <unordered_map>
#include <vector>
#include "cam/CamBase.h"
#include "track/TrackBase.h"
#include "update/cam/linefeat/LineFeatureDatabase.h"
namespace viw {
class TrackLSD {
 public:
  TrackLSD(std::unordered_map<size_t, std::shared_ptr<ov_core::CamBase>> cameras, bool stereo, ov_core::TrackBase::HistogramMethod histmethod,
           std::map<int, std::shared_ptr<ov_core::TrackBase>> _trackFEATS)
      : camera_calib(cameras), database(new LineFeatureDatabase), use_stereo(stereo), histogram_method(histmethod), trackFEATS(_trackFEATS) {}
  virtual ~TrackLSD() {}
  void feed_new_camera(const ov_core::CameraData &message, std::vector<Eigen::Vector2d> &vanishing_points);
  std::shared_ptr<LineFeatureDatabase> get_feature_database() { return database; }
 protected:
  std::unordered_map<size_t, std::shared_ptr<ov_core::CamBase>> camera_calib;
  std::shared_ptr<LineFeatureDatabase> database;
  bool use_stereo;
  ov_core::TrackBase::HistogramMethod histogram_method;
  std::map<int, std::shared_ptr<ov_core::TrackBase>> trackFEATS;
};
}  // namespace viw
