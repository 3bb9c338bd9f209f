#pragma once
#include <opencv2/core.hpp>
#include <vector>
namespace ov_core {
struct CameraData {
  double timestamp;
  std::vector<int> sensor_ids;
  std::vector<cv::Mat> images;
  std::vector<cv::Mat> masks;
};
}  // namespace ov_core
