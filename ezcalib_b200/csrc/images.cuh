// images.cuh -- device-resident batch of 8-bit gray images (what cv::imread(..., IMREAD_GRAYSCALE)
// hands to EZCalibrator::processImage, /root/reference/src/ezcalibrator.cpp:44), shared by all detectors.
#pragma once
#include <algorithm>
#include <vector>

#include "ezc_common.cuh"

struct ezc_images {
  ezc_context* ctx = nullptr;
  const uint8_t* d = nullptr;   // device pointer
  ezc::DevBuf own;              // backing store when uploaded from the host
  int n = 0, width = 0, height = 0;
  int64_t pitch = 0;            // bytes between rows
  int64_t image_stride = 0;     // bytes between images
  // asynchronous upload (ezc_images_upload_async): copies run on ctx->copy_stream; ready[c] is recorded after images
  // [c*chunk, (c+1)*chunk) have landed.  Consumers order themselves with ezc::images_wait.
  std::vector<cudaEvent_t> ready;
  int chunk = 0;
  ~ezc_images() {
    if (!ready.empty() && ctx && ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    for (cudaEvent_t e : ready) cudaEventDestroy(e);
  }
};

namespace ezc {
// Make ctx->stream wait until images [0, last] of the batch are resident (no-op for synchronous uploads / wrapped memory).
inline cudaError_t images_wait(ezc_context* ctx, const ezc_images* imgs, int last) {
  if (imgs->ready.empty() || last < 0) return cudaSuccess;
  const int c = std::min<int>((int)imgs->ready.size() - 1, last / imgs->chunk);
  return cudaStreamWaitEvent(ctx->stream, imgs->ready[c], 0);
}
// cornerSubPix on device-resident data (used by the detectors and by the C entry point).
// corners: device float2 [n] in/out; image_index: device int32 [n].
ezc_status subpix_device(ezc_context* ctx, const ezc_images* imgs, float2* corners, const int32_t* image_index, int n,
                         int win_w, int win_h, int max_iter, double eps);
}  // namespace ezc
