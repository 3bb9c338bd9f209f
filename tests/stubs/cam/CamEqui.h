#pragma once
#include "cam/CamBase.h"
namespace ov_core {
class CamEqui : public CamBase {
 public:
  CamEqui(int, int) {}
};
}  // namespace ov_core
