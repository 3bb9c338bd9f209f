// Instantiates both adaptor classes so that every member function is compiled (tests/test_adaptor_syntax.py).
#include "plviwo_ov_adaptor.hpp"

int adaptor_check() {
  std::unordered_map<size_t, std::shared_ptr<ov_core::CamBase>> cams;
  auto kl = std::make_shared<plviwo::TrackB200>(cams, 200, 0, false, ov_core::TrackBase::HISTOGRAM, 20, 5, 5, 10);
  std::map<int, std::shared_ptr<ov_core::TrackBase>> feats{{0, kl}};
  plviwo::TrackLSDB200 ls(cams, false, ov_core::TrackBase::HISTOGRAM, feats);
  ov_core::CameraData msg;
  std::vector<Eigen::Vector2d> vps(3);
  kl->feed_new_camera(msg);
  ls.feed_new_camera(msg, vps);
  return 0;
}
