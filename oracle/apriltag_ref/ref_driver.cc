// C entry points around the reference's own AprilTag detector, compiled from the sources where they lie
// under /root/reference/thirdparty/ethz_apriltag2 (see Makefile).  TEST INFRASTRUCTURE ONLY: the oracle
// for the AprilGrid half of the detection path, and the CPU baseline ("reference" kind) of bench.py.
//
//   ref_extract_tags  = AprilTags::TagDetector(tagCodes36h11).extractTags(image)   (TagDetector.cc:40-586)
//   ref_stages        = the intermediate products of extractTags' steps 1-7, obtained by calling the
//                       reference's own classes (FloatImage, Gaussian, Edge, UnionFindSimple,
//                       GLineSegment2D, Gridder, Quad) in the order TagDetector.cc:43-399 does; only the
//                       glue loops are restated here, for stage-by-stage parity tests.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <vector>

#include "apriltags/Edge.h"
#include "apriltags/FloatImage.h"
#include "apriltags/GLine2D.h"
#include "apriltags/GLineSegment2D.h"
#include "apriltags/Gaussian.h"
#include "apriltags/Gridder.h"
#include "apriltags/MathUtil.h"
#include "apriltags/Quad.h"
#include "apriltags/Segment.h"
#include "apriltags/TagDetector.h"
#include "apriltags/Tag36h11.h"
#include "apriltags/UnionFindSimple.h"
#include "apriltags/XYWeight.h"

using namespace AprilTags;

extern "C" {

// Returns the number of detections (may exceed max_det; only the first max_det are written).
int ref_extract_tags(const uint8_t* img, int W, int H, int max_det, int32_t* id, int32_t* hamming, int32_t* rotation,
                     float* p /*[n][4][2]*/, float* cxy /*[n][2]*/, float* perimeter, uint64_t* obs_code) {
  static TagDetector detector(tagCodes36h11);
  cv::Mat image(H, W, const_cast<unsigned char*>(img));
  std::vector<TagDetection> dets = detector.extractTags(image);
  const int n = static_cast<int>(dets.size());
  for (int i = 0; i < n && i < max_det; ++i) {
    const TagDetection& d = dets[i];
    id[i] = d.id; hamming[i] = d.hammingDistance; rotation[i] = d.rotation;
    for (int k = 0; k < 4; ++k) { p[(i * 4 + k) * 2] = d.p[k].first; p[(i * 4 + k) * 2 + 1] = d.p[k].second; }
    cxy[i * 2] = d.cxy.first; cxy[i * 2 + 1] = d.cxy.second;
    perimeter[i] = d.observedPerimeter;
    obs_code[i] = static_cast<uint64_t>(d.obsCode);
  }
  return n;
}

// Stage products.  All output pointers may be null.  Capacities are given by the caller.
//   seg_out [W*H]   blurred image (fimSeg);  theta_out/mag_out [W*H]
//   edges_out [max_edges][3] = (cost, pixelIdxA, pixelIdxB) AFTER the stable sort; returns n_edges in counts[0]
//   rep_out [W*H]   union-find representative of every pixel after mergeEdges; size_out [W*H] its set size
//   segs_out [max_segs][6] = x0,y0,x1,y1,theta,length in creation order; counts[1] = n_segments
//   quads_out [max_quads][9] = 4 corners + observedPerimeter in creation order; counts[2] = n_quads
void ref_stages(const uint8_t* img, int width, int height, float* seg_out, float* theta_out, float* mag_out,
                int32_t* edges_out, int max_edges, int32_t* rep_out, int32_t* size_out, float* segs_out, int max_segs,
                float* quads_out, int max_quads, int32_t* counts) {
  FloatImage fimOrig(width, height);
  {
    int i = 0;
    for (int y = 0; y < height; y++)
      for (int x = 0; x < width; x++) { fimOrig.set(x, y, img[i] / 255.); i++; }
  }
  std::pair<int, int> opticalCenter(width / 2, height / 2);
  float segSigma = 0.8f;
  FloatImage fimSeg;
  {
    int filtsz = ((int)std::max(3.0f, 3 * segSigma)) | 1;
    std::vector<float> filt = Gaussian::makeGaussianFilter(segSigma, filtsz);
    fimSeg = fimOrig;
    fimSeg.filterFactoredCentered(filt, filt);
  }
  FloatImage fimTheta(width, height), fimMag(width, height);
  for (int y = 1; y < height - 1; y++)
    for (int x = 1; x < width - 1; x++) {
      float Ix = fimSeg.get(x + 1, y) - fimSeg.get(x - 1, y);
      float Iy = fimSeg.get(x, y + 1) - fimSeg.get(x, y - 1);
      float mag = Ix * Ix + Iy * Iy;
      float theta = atan2(Iy, Ix);
      fimTheta.set(x, y, theta);
      fimMag.set(x, y, mag);
    }
  if (seg_out) for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) seg_out[y * width + x] = fimSeg.get(x, y);
  if (theta_out) for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) theta_out[y * width + x] = fimTheta.get(x, y);
  if (mag_out) for (int y = 0; y < height; y++) for (int x = 0; x < width; x++) mag_out[y * width + x] = fimMag.get(x, y);

  UnionFindSimple uf(width * height);
  std::vector<Edge> edges(static_cast<size_t>(width) * height * 4);
  size_t nEdges = 0;
  {
    std::vector<float> storage(static_cast<size_t>(width) * height * 4);
    float* tmin = &storage[static_cast<size_t>(width) * height * 0];
    float* tmax = &storage[static_cast<size_t>(width) * height * 1];
    float* mmin = &storage[static_cast<size_t>(width) * height * 2];
    float* mmax = &storage[static_cast<size_t>(width) * height * 3];
    for (int y = 0; y + 1 < height; y++)
      for (int x = 0; x + 1 < width; x++) {
        float mag0 = fimMag.get(x, y);
        if (mag0 < Edge::minMag) continue;
        mmax[y * width + x] = mag0; mmin[y * width + x] = mag0;
        float theta0 = fimTheta.get(x, y);
        tmin[y * width + x] = theta0; tmax[y * width + x] = theta0;
        Edge::calcEdges(theta0, x, y, fimTheta, fimMag, edges, nEdges);
      }
    edges.resize(nEdges);
    std::stable_sort(edges.begin(), edges.end());
    if (edges_out)
      for (size_t i = 0; i < nEdges && i < static_cast<size_t>(max_edges); ++i) {
        edges_out[i * 3] = edges[i].cost; edges_out[i * 3 + 1] = edges[i].pixelIdxA; edges_out[i * 3 + 2] = edges[i].pixelIdxB;
      }
    Edge::mergeEdges(edges, uf, tmin, tmax, mmin, mmax);
  }
  counts[0] = static_cast<int32_t>(nEdges);
  if (rep_out || size_out)
    for (int i = 0; i < width * height; ++i) {
      if (rep_out) rep_out[i] = uf.getRepresentative(i);
      if (size_out) size_out[i] = uf.getSetSize(i);
    }

  std::map<int, std::vector<XYWeight> > clusters;
  for (int y = 0; y + 1 < height; y++)
    for (int x = 0; x + 1 < width; x++) {
      if (uf.getSetSize(y * width + x) < Segment::minimumSegmentSize) continue;
      int rep = (int)uf.getRepresentative(y * width + x);
      clusters[rep].push_back(XYWeight(x, y, fimMag.get(x, y)));
    }
  std::vector<Segment> segments;
  for (std::map<int, std::vector<XYWeight> >::const_iterator it = clusters.begin(); it != clusters.end(); it++) {
    std::vector<XYWeight> points = it->second;
    GLineSegment2D gseg = GLineSegment2D::lsqFitXYW(points);
    float length = MathUtil::distance2D(gseg.getP0(), gseg.getP1());
    if (length < Segment::minimumLineLength) continue;
    Segment seg;
    float dy = gseg.getP1().second - gseg.getP0().second;
    float dx = gseg.getP1().first - gseg.getP0().first;
    float tmpTheta = std::atan2(dy, dx);
    seg.setTheta(tmpTheta);
    seg.setLength(length);
    float flip = 0, noflip = 0;
    for (unsigned int i = 0; i < points.size(); i++) {
      XYWeight xyw = points[i];
      float theta = fimTheta.get((int)xyw.x, (int)xyw.y);
      float mag = fimMag.get((int)xyw.x, (int)xyw.y);
      float err = MathUtil::mod2pi(theta - seg.getTheta());
      if (err < 0) noflip += mag; else flip += mag;
    }
    if (flip > noflip) { float temp = seg.getTheta() + (float)M_PI; seg.setTheta(temp); }
    float dot = dx * std::cos(seg.getTheta()) + dy * std::sin(seg.getTheta());
    if (dot > 0) {
      seg.setX0(gseg.getP1().first); seg.setY0(gseg.getP1().second);
      seg.setX1(gseg.getP0().first); seg.setY1(gseg.getP0().second);
    } else {
      seg.setX0(gseg.getP0().first); seg.setY0(gseg.getP0().second);
      seg.setX1(gseg.getP1().first); seg.setY1(gseg.getP1().second);
    }
    segments.push_back(seg);
  }
  counts[1] = static_cast<int32_t>(segments.size());
  if (segs_out)
    for (size_t i = 0; i < segments.size() && i < static_cast<size_t>(max_segs); ++i) {
      float* o = segs_out + i * 6;
      o[0] = segments[i].getX0(); o[1] = segments[i].getY0(); o[2] = segments[i].getX1(); o[3] = segments[i].getY1();
      o[4] = segments[i].getTheta(); o[5] = segments[i].getLength();
    }

  Gridder<Segment> gridder(0, 0, width, height, 10);
  for (unsigned int i = 0; i < segments.size(); i++) gridder.add(segments[i].getX0(), segments[i].getY0(), &segments[i]);
  for (unsigned i = 0; i < segments.size(); i++) {
    Segment& parentseg = segments[i];
    GLine2D parentLine(std::pair<float, float>(parentseg.getX0(), parentseg.getY0()),
                       std::pair<float, float>(parentseg.getX1(), parentseg.getY1()));
    Gridder<Segment>::iterator iter = gridder.find(parentseg.getX1(), parentseg.getY1(), 0.5f * parentseg.getLength());
    while (iter.hasNext()) {
      Segment& child = iter.next();
      if (MathUtil::mod2pi(child.getTheta() - parentseg.getTheta()) > 0) continue;
      GLine2D childLine(std::pair<float, float>(child.getX0(), child.getY0()),
                        std::pair<float, float>(child.getX1(), child.getY1()));
      std::pair<float, float> p = parentLine.intersectionWith(childLine);
      if (p.first == -1) continue;
      float parentDist = MathUtil::distance2D(p, std::pair<float, float>(parentseg.getX1(), parentseg.getY1()));
      float childDist = MathUtil::distance2D(p, std::pair<float, float>(child.getX0(), child.getY0()));
      if (std::max(parentDist, childDist) > parentseg.getLength()) continue;
      parentseg.children.push_back(&child);
    }
  }
  std::vector<Quad> quads;
  std::vector<Segment*> tmp(5);
  for (unsigned int i = 0; i < segments.size(); i++) {
    tmp[0] = &segments[i];
    Quad::search(fimOrig, tmp, segments[i], 0, quads, opticalCenter);
  }
  counts[2] = static_cast<int32_t>(quads.size());
  if (quads_out)
    for (size_t i = 0; i < quads.size() && i < static_cast<size_t>(max_quads); ++i) {
      float* o = quads_out + i * 9;
      for (int k = 0; k < 4; ++k) { o[2 * k] = quads[i].quadPoints[k].first; o[2 * k + 1] = quads[i].quadPoints[k].second; }
      o[8] = quads[i].observedPerimeter;
    }
}

int ref_num_codes36h11(void) { return static_cast<int>(tagCodes36h11.codes.size()); }
uint64_t ref_code36h11(int i) { return tagCodes36h11.codes[i]; }

}  // extern "C"
