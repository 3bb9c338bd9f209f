#pragma once
#include <cstddef>
namespace ov_core {
class FeatureDatabase {
 public:
  void update_feature(size_t id, double timestamp, size_t cam_id, float u, float v, float u_n, float v_n) {
    (void)id; (void)timestamp; (void)cam_id; (void)u; (void)v; (void)u_n; (void)v_n;
  }
  void change_feat_id(size_t id_old, size_t id_new) { (void)id_old; (void)id_new; }
};
}  // namespace ov_core
