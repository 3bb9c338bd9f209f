#pragma once
#include <Eigen/Eigen>
#include <map>
#include <vector>
namespace viw {
class LineFeatureDatabase {
 public:
  void update_feature(size_t id, double timestamp, size_t cam_id, Eigen::Vector4f line, Eigen::Vector4f line_n,
                      std::map<int, double> points_line, std::vector<Eigen::Vector2f> points, int D) {
    (void)id; (void)timestamp; (void)cam_id; (void)line; (void)line_n; (void)points_line; (void)points; (void)D;
  }
};
}  // namespace viw
