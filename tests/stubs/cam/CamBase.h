#pragma once
#include <Eigen/Eigen>
namespace ov_core {
class CamBase {
 public:
  virtual ~CamBase() {}
  Eigen::MatrixXd get_value() { return camera_values; }
 protected:
  Eigen::MatrixXd camera_values;
};
}  // namespace ov_core
