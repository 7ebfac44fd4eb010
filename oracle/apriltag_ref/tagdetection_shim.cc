// Restatement of the three TagDetection members that extractTags needs
// (/root/reference/thirdparty/ethz_apriltag2/src/TagDetection.cc:21-74).  The rest of that file
// (getRelativeTransform / draw, :76-159) is off the hot path and needs the real OpenCV (solvePnP, line,
// putText), so TagDetection.cc is the ONE reference source that is not compiled unmodified.
// TEST INFRASTRUCTURE ONLY.
#include "apriltags/TagDetection.h"
#include "apriltags/MathUtil.h"

namespace AprilTags {

TagDetection::TagDetection()
  : good(false), obsCode(), code(), id(), hammingDistance(), rotation(), p(),
    cxy(), observedPerimeter(), homography(), hxy() {
  homography.setZero();
}

TagDetection::TagDetection(int _id)
  : good(false), obsCode(), code(), id(_id), hammingDistance(), rotation(), p(),
    cxy(), observedPerimeter(), homography(), hxy() {
  homography.setZero();
}

std::pair<float,float> TagDetection::interpolate(float x, float y) const {
  float z = homography(2,0)*x + homography(2,1)*y + homography(2,2);
  if ( z == 0 )
    return std::pair<float,float>(0,0);
  float newx = (homography(0,0)*x + homography(0,1)*y + homography(0,2))/z + hxy.first;
  float newy = (homography(1,0)*x + homography(1,1)*y + homography(1,2))/z + hxy.second;
  return std::pair<float,float>(newx,newy);
}

bool TagDetection::overlapsTooMuch(const TagDetection &other) const {
  float radius =
    ( MathUtil::distance2D(p[0], p[1]) +
      MathUtil::distance2D(p[1], p[2]) +
      MathUtil::distance2D(p[2], p[3]) +
      MathUtil::distance2D(p[3], p[0]) +
      MathUtil::distance2D(other.p[0], other.p[1]) +
      MathUtil::distance2D(other.p[1], other.p[2]) +
      MathUtil::distance2D(other.p[2], other.p[3]) +
      MathUtil::distance2D(other.p[3], other.p[0]) ) / 16.0f;
  float dist = MathUtil::distance2D(cxy, other.cxy);
  return ( dist < radius );
}

} // namespace
