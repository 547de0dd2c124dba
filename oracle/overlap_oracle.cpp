/*
 * oracle/overlap_oracle.cpp -- TEST INFRASTRUCTURE ONLY (CPU oracle of K1; never on the product path).
 *
 * Restatement of the keyframe-overlap masks of Frontend::doWeNeedANewKeyframe (reference okvis_frontend/src/Frontend.cpp:
 * 1058-1167) and ViSlamBackend::overlapFraction (okvis_ceres/src/ViSlamBackend.cpp:2341-2426): per camera image two
 * (rows/10) x (cols/10) CV_8UC1 masks, "detections" and "matches", into which cv::circle(mask, keypoint.pt*0.1,
 * int(radius), 255, cv::FILLED) is drawn for every keypoint / every matched keypoint, radius = min(rows, cols) * kptrad;
 * then countNonZero(matches & detections) and countNonZero(matches | detections).
 *
 * cv::circle with thickness FILLED, LINE_8, shift 0 runs OpenCV's Circle() (modules/imgproc/src/drawing.cpp; OpenCV is
 * not vendored in /root/reference -- algorithm restated from the published source and PINNED against cv2 4.13.0 masks
 * in tests/golden/circle_cv2_4_13.npz): the midpoint iteration (dx, dy) with horizontal spans
 * [cx-dx, cx+dx] on rows cy-+dy and [cx-dy, cx+dy] on rows cy-+dx, clipped to the image. The centre is
 * Point(cvRound(float(pt.x * 0.1)), cvRound(float(pt.y * 0.1))): Point2f * double is evaluated in double and stored as
 * float, the Point2f -> Point conversion rounds half to even.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace {
void hline(uint8_t* row, int x0, int x1) { for (int x = x0; x <= x1; x++) row[x] = 255; }

void circle_filled(uint8_t* img, int width, int height, int step, int cx, int cy, int radius)
{
  int err = 0, dx = radius, dy = 0, plus = 1, minus = (radius << 1) - 1;
  const int inside = cx >= radius && cx < width - radius && cy >= radius && cy < height - radius;
  while (dx >= dy) {
    int mask;
    int y11 = cy - dy, y12 = cy + dy, y21 = cy - dx, y22 = cy + dx;
    int x11 = cx - dx, x12 = cx + dx, x21 = cx - dy, x22 = cx + dy;
    if (inside) {
      hline(img + y11 * step, x11, x12); hline(img + y12 * step, x11, x12);
      hline(img + y21 * step, x21, x22); hline(img + y22 * step, x21, x22);
    } else if (x11 < width && x12 >= 0 && y21 < height && y22 >= 0) {
      x11 = std::max(x11, 0); x12 = std::min(x12, width - 1);
      if ((unsigned)y11 < (unsigned)height) hline(img + y11 * step, x11, x12);
      if ((unsigned)y12 < (unsigned)height) hline(img + y12 * step, x11, x12);
      if (x21 < width && x22 >= 0) {
        x21 = std::max(x21, 0); x22 = std::min(x22, width - 1);
        if ((unsigned)y21 < (unsigned)height) hline(img + y21 * step, x21, x22);
        if ((unsigned)y22 < (unsigned)height) hline(img + y22 * step, x21, x22);
      }
    }
    dy++;
    err += plus;
    plus += 2;
    mask = (err <= 0) - 1;
    err -= minus & mask;
    dx += mask;
    minus -= mask & 2;
  }
}
}  // namespace

extern "C" void okvo_circle_filled(uint8_t* img, int width, int height, int cx, int cy, int radius)
{
  circle_filled(img, width, height, width, cx, cy, radius);
}

/* one camera image: image_rows x image_cols pixels, n keypoints (pt.x, pt.y floats), matched[k] != 0 -> also drawn into
 * the matches mask. Optionally returns the two masks (rows/10 x cols/10 bytes each). */
extern "C" void okvo_overlap_counts(int image_rows, int image_cols, int n, const float* xy, const uint8_t* matched, double kptrad,
                                    int32_t* intersection, int32_t* uni, uint8_t* det_out, uint8_t* mat_out)
{
  const int rows = image_rows / 10, cols = image_cols / 10;
  std::vector<uint8_t> det((size_t)rows * cols, 0), mat((size_t)rows * cols, 0);
  const double radius = double(std::min(rows, cols)) * kptrad;
  for (int k = 0; k < n; k++) {
    const float px = (float)((double)xy[2 * k] * 0.1), py = (float)((double)xy[2 * k + 1] * 0.1);
    const int cx = (int)lrintf(px), cy = (int)lrintf(py);
    circle_filled(det.data(), cols, rows, cols, cx, cy, int(radius));
    if (matched[k]) circle_filled(mat.data(), cols, rows, cols, cx, cy, int(radius));
  }
  int ic = 0, uc = 0;
  for (size_t i = 0; i < det.size(); i++) { ic += (mat[i] & det[i]) != 0; uc += (mat[i] | det[i]) != 0; }
  *intersection = ic; *uni = uc;
  if (det_out) memcpy(det_out, det.data(), det.size());
  if (mat_out) memcpy(mat_out, mat.data(), mat.size());
}
