/* oracle/oracle.h -- TEST INFRASTRUCTURE ONLY. C interface of the CPU oracle (see brisk_oracle.c / match_oracle.c).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this. */
#ifndef OKVO_ORACLE_H
#define OKVO_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* 28-byte POD with the field order of cv::KeyPoint (pt.x, pt.y, size, angle, response, octave, class_id) */
typedef struct {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} okvo_keypoint_t;

typedef struct okvo_brisk okvo_brisk_t;

void okvo_resize_area(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh, int dstride);
int okvo_oast916_bstar(const uint8_t* p, int stride);
int okvo_agast58_bstar(const uint8_t* p, int stride);
void okvo_dense_b0(const uint8_t* img, int w, int h, uint8_t* out);
void okvo_integral(const uint8_t* img, int W, int H, int stride, int32_t* integral);

okvo_brisk_t* okvo_brisk_create(int threshold, int octaves, float pattern_scale);
void okvo_brisk_destroy(okvo_brisk_t* b);
int okvo_brisk_descriptor_bytes(const okvo_brisk_t* b);
int okvo_brisk_num_pairs(const okvo_brisk_t* b, int* n_short, int* n_long);
const float* okvo_brisk_pattern(const okvo_brisk_t* b);
void okvo_brisk_size_list(const okvo_brisk_t* b, unsigned* out);
void okvo_brisk_pairs(const okvo_brisk_t* b, unsigned* short_ij, int* long_ijw);
int okvo_brisk_kscale(float size);

int okvo_brisk_detect_raw(okvo_brisk_t* b, const uint8_t* img, int W, int H, int stride, okvo_keypoint_t* kp, int cap);
int okvo_brisk_num_layers(const okvo_brisk_t* b);
int okvo_brisk_layer(const okvo_brisk_t* b, int i, int* w, int* h, float* scale, float* offset, const uint8_t** img,
                     const uint8_t** scores);
int okvo_brisk_compute(const okvo_brisk_t* b, const uint8_t* image, int W, int H, okvo_keypoint_t* kp, int n, uint8_t* desc);
int okvo_cap_strongest(okvo_keypoint_t* kp, int n, int max_kp);
int okvo_brisk_detect_and_compute(okvo_brisk_t* b, const uint8_t* img, int W, int H, int stride, int max_kp,
                                  okvo_keypoint_t* kp, int cap, uint8_t* desc);


/* Harris + uniformity detector and 48-byte BRISK2 extractor (PARITY UNPINNED vs smartroboticslab/brisk) */
okvo_brisk_t* okvo_brisk_create_dmax(int threshold, int octaves, float pattern_scale, double dmax_factor);
void okvo_harris_scores(const uint8_t* img, int W, int H, int stride, int32_t* score);
int okvo_harris_maxima(const int32_t* score, int W, int H, int threshold, int* xy, int cap);
float okvo_harris_lut(float radius, int dx, int dy);
int okvo_harris_detect(const uint8_t* img, int W, int H, int stride, float radius, int threshold, int max_kp, okvo_keypoint_t* kp, int cap);
int okvo_brisk2_basic_scale(void);
int okvo_brisk2_warp(const float e[3], const float J[6], const float d[3], float fu, float M[4]);
int okvo_brisk2_compute(const okvo_brisk_t* b, const uint8_t* image, int W, int H, okvo_keypoint_t* kp, int n, uint8_t* desc,
                        const float* rays, const float* jac, float fu, const float* dir);

#ifdef __cplusplus
}
#endif
#endif
