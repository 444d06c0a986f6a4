// focr_decode.cuh -- host interface of the focr line-decode kernels (focr_decode.cu).
#pragma once
#include "common.cuh"

namespace focr {
}  // namespace focr
