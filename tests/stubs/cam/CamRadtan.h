#pragma once
#include "cam/CamBase.h"
namespace ov_core {
class CamRadtan : public CamBase {
 public:
  CamRadtan(int, int) {}
};
}  // namespace ov_core
