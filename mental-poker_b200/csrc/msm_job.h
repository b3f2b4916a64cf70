// One MSM job: scalars [scalar_off, scalar_off + len) against points [point_off, point_off + len).
#pragma once
#include <stdint.h>

namespace mp {
struct MsmJob {
  uint32_t scalar_off, point_off, len;
};
}  // namespace mp
