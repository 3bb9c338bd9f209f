#pragma once
#include <cstddef>
namespace cv {
struct Point2f { float x = 0, y = 0; };
struct KeyPoint { Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1; };
struct Mat {
  int rows = 0, cols = 0;
  size_t step = 0;
  unsigned char *data = nullptr;
  bool empty() const { return data == nullptr; }
};
}  // namespace cv
