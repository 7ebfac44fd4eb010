// images.cuh -- device-resident batch of 8-bit gray images (what cv::imread(..., IMREAD_GRAYSCALE)
// hands to EZCalibrator::processImage, /root/reference/src/ezcalibrator.cpp:44), shared by all detectors.
#pragma once
#include "ezc_common.cuh"

struct ezc_images {
  ezc_context* ctx = nullptr;
  const uint8_t* d = nullptr;   // device pointer
  ezc::DevBuf own;              // backing store when uploaded from the host
  int n = 0, width = 0, height = 0;
  int64_t pitch = 0;            // bytes between rows
  int64_t image_stride = 0;     // bytes between images
};

namespace ezc {
// cornerSubPix on device-resident data (used by the detectors and by the C entry point).
// corners: device float2 [n] in/out; image_index: device int32 [n].
ezc_status subpix_device(ezc_context* ctx, const ezc_images* imgs, float2* corners, const int32_t* image_index, int n,
                         int win_w, int win_h, int max_iter, double eps);
}  // namespace ezc
