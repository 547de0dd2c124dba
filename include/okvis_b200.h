/*
 * okvis_b200.h -- C ABI of libokvis_b200.so: the B200-native OKVIS2 vision front-end hot path
 * (detect -> describe -> match). Plain pointers and sizes only; no CUDA/torch/OpenCV/Eigen types.
 *
 * Every entry point names the reference interface it replaces (paths relative to the okvis2 tree @464180aa).
 * All functions return 0 (OKB_OK) or a negative okb_status; okb_last_error() gives the text. Nothing throws
 * across this boundary; the C++ adapter turns a non-zero status into OKVIS_THROW(okvis::Frontend::Exception)
 * (reference okvis_frontend/include/okvis/Frontend.hpp:60). There is NO CPU fallback: without a CUDA device
 * okb_create fails with OKB_ERR_NO_DEVICE.
 *
 * Memory: all pointers are HOST pointers owned by the caller and only used during the call, except in the
 * *_device variants (device pointers, used by the benchmark to time the resident-input path).
 */
#ifndef OKVIS_B200_H
#define OKVIS_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  OKB_OK = 0,
  OKB_ERR_NO_DEVICE = -1,   /* no CUDA device / driver: the product path refuses to run */
  OKB_ERR_CUDA = -2,        /* a CUDA runtime call or kernel failed */
  OKB_ERR_ARGUMENT = -3,    /* bad argument (null pointer, size out of range, unsupported D) */
  OKB_ERR_CAPACITY = -4,    /* a fixed-capacity device buffer overflowed (raise the config capacity) */
  OKB_ERR_UNSUPPORTED = -5, /* feature not built (octaves > 0 with descriptor_bytes = 48, okb_fetch_layer on such a camera) */
  OKB_ERR_NCCL = -6
} okb_status;

/* 28-byte POD with the exact field order of cv::KeyPoint (pt.x, pt.y, size, angle, response, octave, class_id),
 * so the adapter can memcpy into std::vector<cv::KeyPoint> (okvis::Frame::keypoints_,
 * okvis_cv/include/okvis/Frame.hpp:247-265). */
typedef struct {
  float x, y, size, angle, response;
  int32_t octave, class_id;
} okb_keypoint_t;

/* Per-camera detector/extractor configuration == the constructor arguments of the detector/extractor pair made
 * in Frontend::initialiseBriskFeatureDetectors (okvis_frontend/src/Frontend.cpp:2398-2417) from
 * FrontendParameters (okvis_common/include/okvis/Parameters.hpp:123-133). */
typedef struct {
  int32_t width, height;    /* image size (u8, single channel) */
  int32_t threshold;        /* D = 64: AGAST corner threshold (1..254); D = 48: absolute Harris corner threshold (>= 1) */
  int32_t octaves;          /* 0 = single scale (the shipped okvis setting), n>0 = 2n scale-space layers */
  int32_t max_keypoints;    /* FrontendParameters::max_num_keypoints: keep the N strongest; 0 = no cap */
  int32_t descriptor_bytes; /* 64 = AGAST + BRISK-512 (north_star, pinned to OpenCV 4.13). 48 = the pair OKVIS2 itself constructs
                             * (Frontend.cpp:2406-2412): Harris + uniformity-enforcement detector and 48-byte BRISK2 extractor; then
                             * `threshold` is the absolute Harris threshold (Parameters.hpp:125), `uniformity_radius` the radius in
                             * pixels (Parameters.hpp:124), octaves must be 0. PARITY UNPINNED vs smartroboticslab/brisk@1ef8b42a. */
  int32_t max_batch;        /* frames per batched call (>=1) */
  float pattern_scale;      /* BRISK pattern scale (1.0) */
  float uniformity_radius;  /* D = 48 only: FrontendParameters::detection_threshold ("uniformity radius in pixels") */
} okb_camera_config_t;

typedef struct okb_context okb_context_t;

/* ---- life cycle (replaces Frontend::Frontend + the six setters, Frontend.cpp:2398-2417; called where
 *      ThreadedSlam::init pushes the yaml parameters, okvis_multisensor_processing/src/ThreadedSlam.cpp:89-94) */
int okb_create(int device, int n_cams, const okb_camera_config_t* cfgs, okb_context_t** out);
void okb_destroy(okb_context_t* ctx);
const char* okb_last_error(void);
const char* okb_version(void);
/* 1 when the cos() the matchers' gates use (cos(2.6 sigma), cos(6 sigma) of triangulateFast, okvis_frontend/src/
 * stereo_triangulation.cpp:82-127) was verified at okb_create to return exactly what this machine's libm returns on 65 536
 * arguments. It is one function for the host-buffer and the device-resident matcher forms (a restatement of glibc's
 * algorithm, csrc/okb_gatecos.h), so those agree by construction either way. */
int okb_gate_cos_exact(const okb_context_t* ctx);
/* number of kernels of this library launched on this context so far (bench.py's gpu_launches) */
int64_t okb_launch_count(const okb_context_t* ctx);
/* the CUDA stream (cudaStream_t as void*) camera `cam` works on; okb_sync waits for all of them */
void* okb_stream(okb_context_t* ctx, int cam);
int okb_sync(okb_context_t* ctx);
/* How the host-buffer entry points (okb_detect_describe[_batch], okb_match_map3d_batch, okb_match_stereo_batch) wait for
 * the device: 0 (default) = spin, lowest latency; 1 = sleep on a blocking event, for hosts that run more waiting threads than
 * cores (many ranks x sequences per box). The reference has no counterpart: its calls are synchronous CPU code. */
int okb_set_blocking_sync(okb_context_t* ctx, int on);

/* ---- detect + describe: replaces okvis::Frame::detect + okvis::Frame::describe, i.e. the calls
 *      detector_->detect(image_, keypoints_) and extractor_->compute(image_, keypoints_, descriptors_)
 *      (okvis_cv/include/okvis/implementation/Frame.hpp:140-154,160-175) issued by
 *      Frontend::detectAndDescribe (Frontend.cpp:221-269). Like compute(), it may drop border keypoints; the
 *      keypoints returned are exactly those that own a descriptor row. Thread-safe for different `cam`
 *      (one stream + workspace per camera; reference: per-camera mutex, Frontend.cpp:226).
 *      kp_out: cap records; desc_out: cap x descriptor_bytes, row-major, continuous (Frame.hpp:287-289);
 *      *n_out: number written. */
int okb_detect_describe(okb_context_t* ctx, int cam, const uint8_t* image, size_t stride_bytes,
                        okb_keypoint_t* kp_out, uint8_t* desc_out, int cap, int* n_out);
/* n_frames <= max_batch images of the same camera in one submission (batch replay, BASELINE config 5).
 * images: n_frames x height x stride_bytes; kp_out: n_frames x cap; desc_out: n_frames x cap x D; n_out: n_frames */
int okb_detect_describe_batch(okb_context_t* ctx, int cam, int n_frames, const uint8_t* images, size_t stride_bytes,
                              okb_keypoint_t* kp_out, uint8_t* desc_out, int cap, int* n_out);
/* Same, but `d_images` (n_frames x height x width, pitch = width) is a DEVICE pointer and results stay on the
 * device inside the context (read them with okb_fetch_features). Asynchronous on okb_stream(ctx, cam). */
int okb_detect_describe_batch_device(okb_context_t* ctx, int cam, int n_frames, const uint8_t* d_images);
int okb_fetch_features(okb_context_t* ctx, int cam, int frame, okb_keypoint_t* kp_out, uint8_t* desc_out, int cap,
                       int* n_out);
/* device pointers of the last result of camera `cam` (valid until the next detect call on it):
 * keypoints [max_batch][capacity], descriptors [max_batch][capacity][D], counts [max_batch] */
int okb_device_features(okb_context_t* ctx, int cam, const okb_keypoint_t** d_kp, const uint8_t** d_desc,
                        const int32_t** d_count, int* capacity);

/* device pointers of the back-projections (D4) of the last detect call of camera `cam` (camera model set):
 * rays [max_batch][capacity][3] doubles (x, y, 1), valid [max_batch][capacity]; valid until the next detect call on it */
int okb_device_back_projections(okb_context_t* ctx, int cam, const double** d_rays, const uint8_t** d_valid);

/* Descriptor rows on the device. okb_device_features hands out rows of the camera's own width (descriptor_bytes: 64 or 48). The
 * device-resident matcher forms built on the tensor-core Hamming scans (okb_match_stereo_device*, okb_match_motion_stereo_device*) and the
 * feature block below read rows in 64-BYTE SLOTS; a 48-byte row sits in the first 48 bytes of its slot with a zero tail (the Hamming
 * distances are the same). A D = 48 camera keeps that second layout itself, so the forms that take a camera index work on it unchanged;
 * blocks the caller passes by pointer (okb_older_view_t::d_desc, the *_ptr forms) use the slot layout. okb_process_multiframe takes
 * cameras of either width (pool rows and returned rows in the camera's own width, older views in slots).
 *
 * Fixed-capacity feature block of a batch, the unit the camera-sharded multi-GPU mode all-gathers:
 *   [counts: n_frames x int32, padded to 256 B][keypoints: n_frames x capacity x 28 B][descriptors: n_frames x capacity x 64 B]
 * okb_export_features packs the last result of camera `cam` into the caller's DEVICE buffer (asynchronous on the camera
 * stream; capacity = okb_device_features). The block layout is what okb_match_stereo_device_ptr consumes. */
size_t okb_feature_block_bytes(int n_frames, int capacity);
int okb_export_features(okb_context_t* ctx, int cam, int n_frames, void* d_block);


/* ---- multi-GPU: the cameras of an NCameraSystem sharded over the GPUs. detect / describe / M1 / M3 are
 *      independent per camera (the reference runs one thread per camera, ThreadedSlam.cpp:432-448); only Frontend::matchStereo
 *      couples cameras, pairwise and only where NCameraSystem::hasOverlap (Frontend.cpp:1990-2000). The exchange is ONE NCCL
 *      all-gather of the fixed-capacity feature blocks (okb_export_features) per batch. NCCL is loaded at run time (dlopen),
 *      errors are OKB_ERR_NCCL.
 *      okb_comm_init_all : one process drives all GPUs (the reference is a single process): ncclCommInitAll.
 *      okb_comm_init_rank: one process per GPU; rank 0 makes a 128-byte id with okb_comm_unique_id and shares it out of band. */
typedef struct okb_comm okb_comm_t;
int okb_comm_unique_id(void* id128);
int okb_comm_init_all(int n_devices, const int* devices, okb_comm_t** out);
int okb_comm_init_rank(int world, int rank, const void* id128, int device, okb_comm_t** out);
void okb_comm_destroy(okb_comm_t* comm);
int okb_comm_world(const okb_comm_t* comm);
int okb_comm_local_ranks(const okb_comm_t* comm);   /* n_devices of okb_comm_init_all, 1 for okb_comm_init_rank */
/* All-gather of bytes_per_rank bytes per rank: d_send[i] (device i) -> d_recv[i] (world x bytes_per_rank, device i) for the
 * n_local local ranks of the communicator (one call covers them all: ncclGroupStart / End). It is enqueued on the stream of
 * camera cams[i] of context ctxs[i], i.e. right behind that camera's okb_export_features, so it overlaps the matchers of the other
 * cameras; okb_comm_wait makes `stream` (e.g. the stereo matcher's) wait for the gathered data of local rank i. Asynchronous. */
int okb_allgather_features(okb_comm_t* comm, int n_local, okb_context_t* const* ctxs, const int* cams, const void* const* d_send,
                           void* const* d_recv, size_t bytes_per_rank);
int okb_comm_wait(okb_comm_t* comm, int local_rank, void* stream);

/* inspection hooks used by the parity tests: layer geometry, layer images and the dense AGAST score maps
 * (b0 = largest threshold at which the pixel is still a 9-16 corner, 0 in the 3-pixel margin) of the last call */
int okb_num_layers(okb_context_t* ctx, int cam);
int okb_layer_info(okb_context_t* ctx, int cam, int layer, int* width, int* height, float* scale, float* offset);
int okb_fetch_layer(okb_context_t* ctx, int cam, int frame, int layer, uint8_t* image_out, uint8_t* score_out);
/* SM cycle stamps (clock64) of the phases of the two single-CTA kernels (tie resolution, selection) of the last call:
 * out16[0..4] resolve phases, [5] rounds, [6] ties, [8..12] finalize phases, [13] keypoints before the cap, [14] candidates */
int okb_debug_stamps(okb_context_t* ctx, int cam, int frame, long long* out16);
/* algorithmic bytes of the pyramid+score pass for one image of camera `cam` (read base + write
 * reduced layers + write score maps, from the actual layer sizes) */
int64_t okb_pyramid_score_bytes(okb_context_t* ctx, int cam);
/* accumulated device time (ms, CUDA events on the camera stream) and launches of the pyramid+score kernels since
 * the last okb_reset_timers; enabled by okb_enable_timers(ctx, 1) */
int okb_enable_timers(okb_context_t* ctx, int on);
int okb_reset_timers(okb_context_t* ctx);
int okb_get_timers(okb_context_t* ctx, int cam, double* pyramid_score_ms, int64_t* pyramid_score_launches,
                   double* total_ms);
/* accumulated device time of the fused score + non-max kernel alone (the dominant kernel of the pass) */
int okb_get_score_kernel_ms(okb_context_t* ctx, int cam, double* score_ms);

/* ---- D4: back-projection. Replaces okvis::Frame::computeBackProjections (okvis_cv/include/okvis/implementation/Frame.hpp:
 *      178-193) -> PinholeCamera<D>::backProject (cameras/implementation/PinholeCamera.hpp:574-592) -> Distortion::undistort.
 *      model: 0 = no distortion, 1 = radial-tangential (k1,k2,p1,p2; 5 Gauss-Newton iterations, exact fp64),
 *             2 = equidistant (k1..k4; 20 iterations; uses atan, equal to libm within the last bits). */
typedef struct {
  int32_t model, reserved;
  double fu, fv, cu, cv;
  double k[4];
} okb_camera_model_t;
int okb_set_camera_model(okb_context_t* ctx, int cam, const okb_camera_model_t* model);
/* rays_out: n x 3 doubles (x, y, 1); valid_out: n success flags (Frame::backProjectionsValid_). Host buffers. */
int okb_back_project(okb_context_t* ctx, int cam, int n, const okb_keypoint_t* kp, double* rays_out, uint8_t* valid_out);


/* ---- D5: camera-awareness maps. Replaces PinholeCamera<D>::initialiseCameraAwarenessMaps (okvis_cv/include/okvis/cameras/
 *      implementation/PinholeCamera.hpp:179-208, called once per camera at configuration load, ViParametersReader.cpp:105): for
 *      every pixel the normalised back-projection (rays: height x width x 3 floats, CV_32FC3; zero where backProject fails) and the
 *      2 x 3 Jacobian of the projection at that ray (jac: height x width x 6 floats, CV_32FC(6), row-major; zero where the
 *      projection is not Successful -- the reference leaves those entries uninitialised). Uses the model of okb_set_camera_model
 *      and the camera's width / height; the maps also stay on the device for the camera-aware extractor. Either output may be
 *      NULL. Synchronous. */
int okb_camera_awareness_maps(okb_context_t* ctx, int cam, float* rays_out, float* jac_out);

/* D1: the extraction direction Frontend::detectAndDescribe hands to the extractor before every frame (okvis_frontend/src/Frontend.cpp:
 * 245-251): gravity in the camera frame, T_WC.inverse().C() * (0, 0, -1), as three floats. The library stores it per camera and hands
 * it back. The D = 48 extractor (descriptor_bytes = 48) aligns its pattern with it when the camera-awareness maps are on the device
 * (okb_camera_awareness_maps after okb_set_camera_model; without the maps it falls back to the gradient orientation of plain BRISK);
 * the BRISK-512 extractor is rotation-invariant by its own orientation estimate and does not consume it. C_WC: row-major 3x3 rotation
 * of T_WC. */
int okb_set_extraction_direction(okb_context_t* ctx, int cam, const double C_WC[9]);
int okb_get_extraction_direction(okb_context_t* ctx, int cam, float dir_out[3]);

/* ---- NCameraSystem::computeOverlaps (okvis_cv/src/NCameraSystem.cpp:48-118): for every ordered camera pair (seenBy, cam) every
 *      pixel of `cam` is back-projected, rotated into `seenBy` (C_rel[(seenBy * n + cam) * 9 ..]: the rotation
 *      (T_SC[seenBy]->inverse() * *T_SC[cam]).C(), row-major), projected, and verified by a back-projection of the image point
 *      (|cos - 1| < 1e-10). overlaps_out: n x n bytes = NCameraSystem::hasOverlap(seenBy, cam), the test Frontend::matchStereo
 *      uses to pick its camera pairs (Frontend.cpp:1990-2000). mats_out: NULL, or n x n host pointers (NULL entries allowed) that
 *      receive overlapMats_[seenBy][cam] (height[cam] x width[cam] bytes). Synchronous. */
int okb_compute_overlaps(okb_context_t* ctx, int n_cams, const okb_camera_model_t* models, const int32_t* widths, const int32_t* heights,
                         const double* C_rel, uint8_t* overlaps_out, uint8_t* const* mats_out);

/* When a camera model is set, okb_detect_describe* also back-projects the keypoints it returns (same kernel as
 * okb_back_project) and keeps the rays in pinned host memory: this call only copies them out (frame = index inside the
 * last batch of camera `cam`). It is how the adapter fills Frame::backProjections_ without a second device round trip. */
int okb_last_back_projections(okb_context_t* ctx, int cam, int frame, int cap, double* rays_out, uint8_t* valid_out, int* n_out);

/* ---- matchers. Descriptors are n x D u8, D in {48, 64}. "First in the reference's iteration order wins ties"
 *      (strict <) is honoured bit-exactly; geometric gates are evaluated on the device in fp64 without FMA
 *      contraction. Index outputs are -1 / distance outputs are match_threshold when nothing matched. ---- */

/* M1: replaces Frontend::matchToMapByThread (Frontend.cpp:1515-1590) for one camera.
 * Candidates are the pooled landmark descriptors (descriptorPool, Frontend.cpp:1221-1223,1348) in ascending
 * LandmarkId order, cand_lm[c] = landmark slot of descriptor c (non-decreasing), lm_proj = LandmarkToMatch::projection
 * (Frontend.cpp:1259), lm_is3d = LandmarkToMatch::is3d. kp_use[k]=0 skips keypoint k (Frontend.cpp:1546-1550). */
int okb_match_map3d(okb_context_t* ctx, int D, int n_kp, const uint8_t* kp_desc, const double* kp_xy,
                    const uint8_t* kp_use, int n_cand, const uint8_t* cand_desc, const int32_t* cand_lm, int n_lm,
                    const double* lm_proj, const uint8_t* lm_is3d, double reprojection_threshold,
                    uint32_t match_threshold, uint32_t* out_dist, int32_t* out_lm);

/* M2: replaces Frontend::matchToMapByThreadUnitialised (Frontend.cpp:1594-1720). kp_e_W: world-frame unit rays
 * T_WC1.C()*e1_C.normalized() (Frontend.cpp:1620-1627); cand_e_W / cand_r_W: LandmarkToMatch::e_W / r_W per pooled
 * descriptor (Frontend.cpp:1333-1334); r_WC1 = T_WC1.r(); sigma = 1/focalLength (Frontend.cpp:1636).
 * kp_prev_lm (may be NULL): landmark slot already assigned to k (loop-closure mode early break, :1706-1709);
 * out_ctr counts those early breaks. out_hp_W: n_kp x 4, written only by non-parallel accepted matches (:1713-1715). */
int okb_match_map_uninit(okb_context_t* ctx, int D, int n_kp, const uint8_t* kp_desc, const double* kp_e_W,
                         const uint8_t* kp_use, const int32_t* kp_prev_lm, int n_cand, const uint8_t* cand_desc,
                         const int32_t* cand_lm, const double* cand_e_W, const double* cand_r_W, int n_lm,
                         const uint8_t* lm_is3d, const double r_WC1[3], double sigma, uint32_t match_threshold,
                         uint32_t* out_dist, int32_t* out_lm, double* out_hp_W, int32_t* out_ctr);
/* M2, device-resident batched form: the queries are the features camera `cam` detected last (okb_detect_describe_batch_device): their
 * descriptors and back-projections stay in HBM, e1_W = T_WC1.C() * e1_C.normalized() is formed on the device per frame (T_WC1: host,
 * n_frames x 12 = C row-major 9 + r 3; Frontend.cpp:1606-1627), keypoints without a valid back-projection are skipped. The pool
 * (device pointers, argument meaning as above) is shared by the frames of the batch. d_kp_use (may be NULL): n_frames x capacity mask
 * of the keypoints to match (`use[k]`, :1630-1633); d_kp_prev_lm (may be NULL): n_frames x capacity; outputs n_frames x capacity
 * (x 4 for hp_W), d_out_ctr (may be NULL): n_frames counters. Asynchronous on the camera's stream. */
int okb_match_map_uninit_device(okb_context_t* ctx, int cam, int n_frames, int n_cand, const uint8_t* d_cand_desc, const int32_t* d_cand_lm,
                                const double* d_cand_e_W, const double* d_cand_r_W, int n_lm, const uint8_t* d_lm_is3d, const double* T_WC1,
                                double sigma, uint32_t match_threshold, const uint8_t* d_kp_use, const int32_t* d_kp_prev_lm,
                                uint32_t* d_out_dist, int32_t* d_out_lm, double* d_out_hp_W, int32_t* d_out_ctr);

/* M3: replaces the worker lambda of Frontend::matchMotionStereo (Frontend.cpp:1809-1907) for one (older frame,
 * camera) pair, up to and including the choice of k1_max/hps_W/initialisable; `quality` (acos, :1888) and the
 * final 4 px re-projection check (:1897-1904) stay with the caller, who owns the camera model.
 * use0[k0]: keypoint k0 is eligible and has a valid back-projection (:1813-1842); e0_W = (T_WC0.C()*e0_C).normalized();
 * size_over_f0[k0] = size0/f0 (sigma = that * 0.125, :1838). Frame-1 arrays are the compacted unmatched set k1s
 * (:1789-1801): valid1 = getBackProjection succeeded, e1_W likewise normalised. T_CW0/T_CW1: row-major 3x4 [R|t] of
 * T_WC.inverse(). out_k1 indexes the compacted set. */
int okb_match_motion_stereo(okb_context_t* ctx, int D, int n0, const uint8_t* desc0, const uint8_t* use0,
                            const double* e0_W, const double* size_over_f0, int n1, const uint8_t* desc1,
                            const uint8_t* valid1, const double* e1_W, const double r_WC0[3], const double r_WC1[3],
                            const double T_CW0[12], const double T_CW1[12], uint32_t match_threshold,
                            int32_t* out_k1, uint32_t* out_dist, double* out_hp_W, uint8_t* out_initialisable);

/* M4: replaces the k0/k1 double loop of Frontend::matchStereo (Frontend.cpp:2016-2074) for one overlapping camera
 * pair (im0 < im1). valid0/valid1 = getBackProjection flags; size_over_f = size/f per keypoint
 * (sigma = max(size0/f0, size1/f1)*0.125, :2035). The serial insertion logic (:2076-2141) stays with the caller. */
int okb_match_stereo(okb_context_t* ctx, int D, int n0, const uint8_t* desc0, const uint8_t* valid0,
                     const double* e0_W, const double* size_over_f0, int n1, const uint8_t* desc1,
                     const uint8_t* valid1, const double* e1_W, const double* size_over_f1, const double r_WC0[3],
                     const double r_WC1[3], const double T_CW0[12], const double T_CW1[12], uint32_t match_threshold,
                     int32_t* out_k1, uint32_t* out_dist, double* out_hp_W, uint8_t* out_initialisable);

/* M5: replaces the descriptor matching loop of Frontend::verifyRecognisedPlace (Frontend.cpp:329-355) for one camera:
 * per old-frame landmark (descriptors lm_offsets[i]..lm_offsets[i+1]) the best keypoint over (descriptor, k) order. */
int okb_match_place(okb_context_t* ctx, int D, int n_lm, const int32_t* lm_offsets, const uint8_t* lm_desc, int n_kp,
                    const uint8_t* kp_desc, uint32_t match_threshold, int32_t* out_k, uint32_t* out_dist);

/* H0: brisk::Hamming::PopcntofXORed over all pairs (call sites Frontend.cpp:341,1580,1661,1846,2024): the plain
 * n_a x n_b distance matrix, for tests and the DBoW2 adapter (okvis_frontend/src/FBrisk.cpp:66). */
int okb_hamming_matrix(okb_context_t* ctx, int D, int n_a, const uint8_t* a, int n_b, const uint8_t* b,
                       uint16_t* out_dist);

/* Device-resident, batched M1 (benchmark "value" leg: inputs already in HBM): matches the features that the last
 * okb_detect_describe_batch_device call left on the device for frames 0..n_frames-1 of camera `cam` against a
 * device-resident landmark pool. d_lm_proj holds one projection table per frame (n_frames x n_lm x 2 doubles: the
 * camera moves between frames). d_out_* hold n_frames x capacity entries (capacity from okb_device_features).
 * D = bytes per pool row; it must equal the camera's descriptor_bytes (e.g. a D = 48 pool from okb_prepare_landmarks
 * against a 64-byte camera is OKB_ERR_ARGUMENT, not a stride mismatch).
 * All d_* pointers are device pointers. Asynchronous on okb_stream(ctx, cam). */
int okb_match_map3d_device(okb_context_t* ctx, int cam, int D, int n_frames, int n_cand, const uint8_t* d_cand_desc,
                           const int32_t* d_cand_lm, int n_lm, const double* d_lm_proj, const uint8_t* d_lm_is3d,
                           double reprojection_threshold, uint32_t match_threshold, uint32_t* d_out_dist,
                           int32_t* d_out_lm);

/* Host-buffer, batched M1 (replay use; benchmark "e2e" leg): same queries as okb_match_map3d_device -- the features
 * the last okb_detect_describe[_batch] call of camera `cam` left on the device -- matched against a landmark pool given in
 * HOST memory (Frontend.cpp:1515-1590; pool layout Frontend.cpp:1221-1223,1259,1267). lm_proj: n_frames x n_lm x 2
 * doubles (one projection table per frame). out_*: n_frames x cap entries in host memory; rows k >= n_out[b] of frame b
 * are unspecified. Page-locked caller buffers are read / written by the copy engines directly, pageable ones go through
 * the library's pinned staging. Synchronous (returns when the outputs are in place). */
int okb_match_map3d_batch(okb_context_t* ctx, int cam, int D, int n_frames, int n_cand, const uint8_t* cand_desc,
                          const int32_t* cand_lm, int n_lm, const double* lm_proj, const uint8_t* lm_is3d,
                          double reprojection_threshold, uint32_t match_threshold, int cap, uint32_t* out_dist,
                          int32_t* out_lm);

/* Host-buffer, batched M4 (Frontend.cpp:2016-2074): okb_match_stereo_device with the outputs delivered to HOST memory
 * (n_frames x cap entries each; out_hp_W n_frames x cap x 4 doubles). Queries = camera cam0, candidates = camera cam1,
 * both as left on the device by the last okb_detect_describe[_batch]; camera models from okb_set_camera_model.
 * Synchronous. */
int okb_match_stereo_batch(okb_context_t* ctx, int cam0, int cam1, int n_frames, const double C_WC0[9], const double r_WC0[3],
                           const double C_WC1[9], const double r_WC1[3], uint32_t match_threshold, int cap, int32_t* out_k1,
                           uint32_t* out_dist, double* out_hp_W, uint8_t* out_initialisable);

/* Device-resident, batched M4 for the benchmark "value" leg and the camera-sharded multi-GPU mode: stereo-matches the
 * features of camera cam0 (queries) against camera cam1 left on the device by okb_detect_describe_batch_device, frame by
 * frame. Back-projection (D4, camera models from okb_set_camera_model), e_W = (C_WC * e_C).normalized(), size/f and the
 * cos tables are computed on the device (by the same function as the host-buffer form: csrc/okb_gatecos.h, equal to libm). C_WC: row-major 3x3 rotation, r_WC: position.
 * Outputs: n_frames x capacity(cam0) entries, device pointers. */
int okb_match_stereo_device(okb_context_t* ctx, int cam0, int cam1, int n_frames, const double C_WC0[9], const double r_WC0[3],
                            const double C_WC1[9], const double r_WC1[3], uint32_t match_threshold, int32_t* d_out_k1,
                            uint32_t* d_out_dist, double* d_out_hp_W, uint8_t* d_out_initialisable);
/* Same on explicit device feature blocks (e.g. a peer camera's block received by NCCL all-gather): keypoints
 * [n_frames][cap], descriptors [n_frames][cap][64], counts [n_frames]. `stream` = cudaStream_t (NULL: the context's
 * match stream); the caller orders it after the producers of the inputs. The call uses one scratch area per context:
 * calls of one context must be issued on streams that serialise them (different contexts are independent). */
int okb_match_stereo_device_ptr(okb_context_t* ctx, int n_frames, int cap0, const okb_keypoint_t* d_kp0, const uint8_t* d_desc0,
                                const int32_t* d_count0, const okb_camera_model_t* model0, const double C_WC0[9],
                                const double r_WC0[3], int cap1, const okb_keypoint_t* d_kp1, const uint8_t* d_desc1,
                                const int32_t* d_count1, const okb_camera_model_t* model1, const double C_WC1[9],
                                const double r_WC1[3], uint32_t match_threshold, void* stream, int32_t* d_out_k1,
                                uint32_t* d_out_dist, double* d_out_hp_W, uint8_t* d_out_initialisable);


/* M3 as a device-resident SEQUENCE over the older keyframes: replaces the loop of Frontend::matchMotionStereo over
 * matchFrameIds for one camera (okvis_frontend/src/Frontend.cpp:1775-1958): per older keyframe the worker loop (:1809-1895), the
 * 4 px re-projection check (:1897-1904; PinholeCamera::projectHomogeneous of T_WC1.inverse() * hp_W with the camera model of
 * okb_set_camera_model) and the serial insertion (:1915-1954: ascending k0, the first k0 that claims a current keypoint wins),
 * whose result -- the current keypoints that now carry a landmark -- shrinks the candidate set k1s of the NEXT older keyframe
 * (:1789-1801). All of it stays on the device, batched over the n_frames current frames of the last detect call of `cam`.
 * older[b * n_older + v]: view v of current frame b (device blocks; the estimator-state tests :1813-1821,1840,1923-1932 are the
 * caller's d_use flags). T_WC1 / T_CW1: n_frames x 12 doubles, pose of the current camera and its inverse as
 * okvis::kinematics::Transformation gives them (C row-major, then r). d_matched1: n_frames x capacity bytes, IN: 1 = the current
 * keypoint already has a landmark (e.g. from M1), OUT: after the insertions of all views. Outputs: n_frames x n_older x cap0
 * entries (d_out_hp_W x 4 doubles); flags bit 0 = matching (matchInfos[k0].matching), bit 1 = initialisable, bit 2 = inserted.
 * `quality` (acos, :1888) only feeds estimator.setLandmark and stays with the caller. Asynchronous on okb_stream(ctx, cam). */
typedef struct {
  const uint8_t* d_desc;   /* n x 64 descriptors (device) */
  const double* d_rays;    /* n x 3 back-projections (x, y, 1), Frame::getBackProjection */
  const uint8_t* d_valid;  /* n: back-projection succeeded */
  const float* d_size;     /* n: cv::KeyPoint::size */
  const uint8_t* d_use;    /* n eligibility flags, or NULL = all eligible */
  int32_t n, reserved;
  double T_WC[12], T_CW[12];   /* pose of this camera in the older frame and its inverse */
} okb_older_view_t;
int okb_match_motion_stereo_device(okb_context_t* ctx, int cam, int n_frames, const double* T_WC1, const double* T_CW1, int n_older,
                                   const okb_older_view_t* older, int cap0, uint32_t match_threshold, uint8_t* d_matched1,
                                   int32_t* d_out_k1, uint32_t* d_out_dist, double* d_out_hp_W, uint8_t* d_out_flags);
/* d_matched[b][k] = (d_lm[b][k] >= 0) for the keypoints of frame b of the last detect call of `cam`, 0 beyond its count: the
 * `landmarkId != 0` test (Frontend.cpp:1792-1795) on the output of okb_match_map3d_device. Asynchronous on okb_stream(ctx, cam). */
int okb_matched_mask_device(okb_context_t* ctx, int cam, int n_frames, const int32_t* d_lm, uint8_t* d_matched);
/* Host-buffer form (replay use; benchmark "e2e" leg): the older views stay device blocks (the keyframe feature store), poses and
 * the matched mask come from / go to HOST memory (matched1: n_frames x cap bytes, in/out), and only the MATCHING entries return:
 * per (frame, view) n_match and, in ascending k0 (the order of the insertion loop, Frontend.cpp:1915), up to cap_m entries of
 * m_k0, m_k1, m_flags, m_hp_W (x 4). More than cap_m matches of a view -> OKB_ERR_CAPACITY. Synchronous. */
int okb_match_motion_stereo_batch(okb_context_t* ctx, int cam, int n_frames, const double* T_WC1, const double* T_CW1, int n_older,
                                  const okb_older_view_t* older, int cap0, uint32_t match_threshold, int cap, uint8_t* matched1,
                                  int cap_m, int32_t* n_match, int32_t* m_k0, int32_t* m_k1, uint8_t* m_flags, double* m_hp_W);
/* Same on explicit device feature blocks of the current frames (keypoints [n_frames][cap1], descriptors [n_frames][cap1][64],
 * counts [n_frames]); `stream` = cudaStream_t or NULL. */
int okb_match_motion_stereo_device_ptr(okb_context_t* ctx, int n_frames, int cap1, const okb_keypoint_t* d_kp1, const uint8_t* d_desc1,
                                       const int32_t* d_count1, const okb_camera_model_t* model, int width, int height,
                                       const double* T_WC1, const double* T_CW1, int n_older, const okb_older_view_t* older, int cap0,
                                       uint32_t match_threshold, void* stream, uint8_t* d_matched1, int32_t* d_out_k1,
                                       uint32_t* d_out_dist, double* d_out_hp_W, uint8_t* d_out_flags);


/* ---- live use: ONE call per multiframe. Replaces the per-frame sequence of ThreadedSlam::processFrame
 *      (okvis_multisensor_processing/src/ThreadedSlam.cpp:429-463,512-533): Frontend::detectAndDescribe for every camera
 *      (Frontend.cpp:221-269), then inside dataAssociationAndInitialization the matchers M1 (matchToMapByThread), M3
 *      (matchMotionStereo) per camera and M4 (matchStereo) per overlapping pair. Everything is enqueued at once (one stream per
 *      camera, the pairs behind them), the results come back in one synchronisation, and from the second call with the same
 *      shapes on the launch sequence is replayed as a captured CUDA graph. HOST buffers in and out; the M1 landmark pool is
 *      re-uploaded only when pool_changed is set, the M3 views are device blocks (the keyframe feature store).
 *      One okb_multiframe_cam_t per camera of the context, in camera order. Synchronous. */
typedef struct {
  /* in */
  const uint8_t* image; size_t stride_bytes;
  int32_t n_cand, n_lm, pool_changed, reserved;           /* M1 pool; n_cand = 0: no M1. pool_changed: (re)upload cand_desc / cand_lm / lm_is3d */
  const uint8_t* cand_desc; const int32_t* cand_lm; const uint8_t* lm_is3d;
  const double* lm_proj;                                   /* n_lm x 2, every frame */
  const double* T_WC1; const double* T_CW1;                /* 12 doubles each (C row-major, r): pose of the camera and its inverse */
  int32_t n_older, cap0; const okb_older_view_t* older;    /* M3: n_older views with at most cap0 keypoints; n_older = 0: no M3 */
  /* out */
  int32_t cap, n;                                          /* caller capacity of the rows below; n = keypoints of this frame */
  okb_keypoint_t* kp; uint8_t* desc;                       /* cap rows (desc: x 64) */
  double* rays; uint8_t* rays_valid;                       /* may be NULL; back-projections (camera model set) */
  uint32_t* m1_dist; int32_t* m1_lm;                       /* cap entries */
  int32_t cap_m, reserved2;                                /* M3: list capacity per view */
  int32_t* m3_n;                                           /* n_older counts */
  int32_t* m3_k0; int32_t* m3_k1; uint8_t* m3_flags; double* m3_hp_W;   /* n_older x cap_m entries (hp x 4), ascending k0 */
} okb_multiframe_cam_t;
typedef struct {
  int32_t cam0, cam1;                                      /* im0 < im1 with NCameraSystem::hasOverlap */
  double C_WC0[9], r_WC0[3], C_WC1[9], r_WC1[3];
  int32_t* k1; uint32_t* dist; double* hp_W; uint8_t* initialisable;   /* capacity of camera cam0's rows (cams[cam0].cap) */
} okb_multiframe_stereo_t;
int okb_process_multiframe(okb_context_t* ctx, int n_cams, okb_multiframe_cam_t* cams, int n_pairs, okb_multiframe_stereo_t* pairs,
                           double reprojection_threshold, uint32_t match_threshold);
/* 0 = always submit the launches directly (default 1: CUDA-graph replay); counters of both kinds of submissions */
int okb_stream_use_graph(okb_context_t* ctx, int on);
int okb_stream_stats(okb_context_t* ctx, long long* graph_launches, long long* direct_calls);
/* host seconds okb_process_multiframe has spent so far in: staging the inputs, submitting, waiting for the device, copying the results out */
int okb_stream_timing(okb_context_t* ctx, double* seconds4, int reset);
/* the per-older-keyframe step of the M3 sequence exists in two forms with identical results: one launch per view (one CTA per frame;
 * the default) and separate gate / finish / check / commit kernels. mode: -1 default, 0 separate kernels, 1 one launch per view. Process-wide; a testing / tuning hook. */
void okb_m3_set_fused(int mode);
/* the Hamming scans of the device-resident M3 / M4 matchers exist in three forms with identical results: mode 2 (default) tcgen05
 * integer MMA with TMEM accumulators, 1 legacy integer MMA (mma.sync), 0 POPC. Process-wide; a testing / tuning hook. */
void okb_scan_set_mma(int mode);

/* ---- P1: landmark-candidate preparation. Replaces the serial host loop of Frontend::matchToMap that
 *      builds landmarksToMatch / descriptorPool for one camera (okvis_frontend/src/Frontend.cpp:1196-1360; pose lookups
 *      ViGraph.cpp:632-645): projection of every landmark into the current view (PinholeCamera::projectHomogeneous,
 *      cameras/implementation/PinholeCamera.hpp:257-292,493-502), field-of-view gate +- reprThreshold, is3d decision,
 *      viewpoint (0.6 rad) and scale (50 %) pruning of the observations walked newest first, "best 3" descriptor pool with
 *      the loop's own bookkeeping (row `o` written, cropped to `o` rows; oracle/prepare_oracle.cpp states what survives).
 *      Output = what M1/M2 take: cand_desc (pool rows), cand_lm (row -> output landmark), lm_proj, lm_is3d (+ e_W, r_W).
 *
 *      The descriptors / back-projections of the frames in the window live in a device-resident store:
 *      okb_store_configure(n_slots, n_cams, D); okb_store_frame(slot, cam, ...) per (multiframe, camera), the slot being
 *      the caller's index for the multiframe id (KeypointIdentifier::frameId). */
int okb_store_configure(okb_context_t* ctx, int n_slots, int n_cams, int D);
/* desc: n x D bytes (Frame::keypointDescriptor), rays: n x 3 doubles (Frame::getBackProjection); host pointers */
int okb_store_frame(okb_context_t* ctx, int slot, int cam, int n, const uint8_t* desc, const double* rays);
/* same, device to device from frame `batch_index` of the last okb_detect_describe[_batch] call of camera `cam`
 * (needs okb_set_camera_model: the rays are the device D4 output) */
int okb_store_frame_from_last(okb_context_t* ctx, int slot, int cam, int batch_index);

typedef struct {
  double T_WC1[12];            /* current camera pose T_WS1 * T_SC: C_WC row-major (9) then r_WC (3)  (Frontend.cpp:1214-1215) */
  double T_CW1[12];            /* its inverse, same layout                                              (Frontend.cpp:1216) */
  okb_camera_model_t model;    /* geometryAs<CAMERA_GEOMETRY>(im) */
  int32_t width, height;
  double repr_threshold;       /* 20.0 with IMU, 150.0 without (Frontend.cpp:1180) */
  int32_t exclusive, reserved; /* 1 = loopClosureLandmarksToUseExclusively given: the caller passes only those landmarks
                                  and the viewpoint / scale gates are off (Frontend.cpp:1300,1306) */
} okb_prepare_view_t;

/* Landmarks in ascending LandmarkId order (MapPoints is an ordered map): hp_W n_lm x 4, quality n_lm, observations of
 * landmark i = obs[obs_begin[i] .. obs_begin[i+1]) in std::set<KeypointIdentifier> order (frame slot, camera, keypoint
 * index; 3 x int32 each) -- the kernel walks them in reverse like the reference. T_WC_old: n_slots x n_cams x 12 doubles
 * (T_WS_old * T_SC_old, layout as T_WC1). Outputs (host, caller-sized: n_lm landmarks, 2 * n_lm pool rows):
 *   out_lm[j]         input index of output landmark j (ascending)          out_proj[j]  projection (2 doubles)
 *   out_is3d[j], out_p_W[j] (3 doubles)                                    out_desc_begin[j .. j+1)  its pool rows
 *   out_pool rows x D, out_cand_lm[row] = j, out_e_W / out_r_W rows x 3 doubles, out_kid rows x 3 int32.
 * Any out_* may be NULL. The same arrays stay on the device for okb_prepared_device. Synchronous. */
int okb_prepare_landmarks(okb_context_t* ctx, const okb_prepare_view_t* view, int n_lm, const double* hp_W, const double* quality,
                          const int32_t* obs_begin, int n_obs, const int32_t* obs, const double* T_WC_old,
                          int32_t* n_out, int32_t* n_rows, int32_t* out_lm, double* out_proj, uint8_t* out_is3d, double* out_p_W,
                          int32_t* out_desc_begin, uint8_t* out_pool, int32_t* out_cand_lm, double* out_e_W, double* out_r_W,
                          int32_t* out_kid);
/* device pointers of the last okb_prepare_landmarks result, in the argument layout of okb_match_map3d_device (valid
 * until the next okb_prepare_landmarks / okb_store_configure) */
int okb_prepared_device(okb_context_t* ctx, const uint8_t** d_cand_desc, const int32_t** d_cand_lm, const double** d_lm_proj,
                        const uint8_t** d_lm_is3d, int32_t* n_cand, int32_t* n_lm);

/* ---- K1: keyframe-overlap masks. Replaces the mask painting and counting inside
 *      Frontend::doWeNeedANewKeyframe (okvis_frontend/src/Frontend.cpp:1068-1101 for the current frame, :1117-1151 for
 *      every keyframe of the window) and ViSlamBackend::overlapFraction (okvis_ceres/src/ViSlamBackend.cpp:2341-2426):
 *      per view (one camera image of one multiframe) the "detections" mask gets cv::circle(mask, pt*0.1,
 *      int(min(rows/10, cols/10) * kptrad), 255, FILLED) for every keypoint and the "matches" mask for every keypoint with
 *      matched[k] != 0 (the caller evaluates `lmId != 0 [&& lmIds.count(lmId)]`); the outputs are
 *      countNonZero(matches & detections) and countNonZero(matches | detections) per view. The caller sums them over
 *      the cameras of a multiframe and divides, as the reference does (kptrad = 0.09, Frontend.cpp:104). */
typedef struct { int32_t image_rows, image_cols, first_keypoint, n_keypoints; } okb_overlap_view_t;
/* xy: n_keypoints x 2 floats (cv::KeyPoint::pt of all views, concatenated), matched: n_keypoints bytes. Synchronous. */
int okb_overlap_counts(okb_context_t* ctx, int n_views, const okb_overlap_view_t* views, int n_keypoints, const float* xy,
                       const uint8_t* matched, double kptrad, int32_t* out_intersection, int32_t* out_union);

/* ---- B1: DBoW2 vocabulary descent with the FBrisk distance. Replaces
 *      TemplatedVocabulary<FBrisk::TDescriptor, FBrisk>::transform(feature, word_id, weight, nid, levelsup) as reached from
 *      the loop-closure queries of the front-end (okvis_frontend/src/Frontend.cpp:661-672,756-760,896-899) with
 *      FBrisk::distance = brisk::Hamming::PopcntofXORed (okvis_frontend/src/FBrisk.cpp:64-67). external/DBoW2 is an empty
 *      submodule in the reference tree: the descent is restated from the published algorithm (oracle/bow_oracle.py).
 *      Vocabulary as stored in resources/small_voc.yml.gz: nodes IN FILE ORDER (nodeId, parentId, weight, descriptor of D
 *      bytes; ids 1..n_nodes, 0 = root), words (wordId, nodeId). */
int okb_bow_load(okb_context_t* ctx, int D, int k, int L, int n_nodes, const int32_t* node_id, const int32_t* parent_id,
                 const double* weight, const uint8_t* desc, int n_words, const int32_t* word_id, const int32_t* word_node);
/* desc: n x D bytes. out_word[i] = word id of the leaf reached, out_weight[i] its weight, out_node[i] = id of the node on
 * the path at level L - levelsup (0 when that is the root), as DBoW2's direct index wants it. out_weight / out_node may be
 * NULL. Synchronous. */
int okb_bow_transform(okb_context_t* ctx, int n, const uint8_t* desc, int levelsup, int32_t* out_word, double* out_weight,
                      int32_t* out_node);

#ifdef __cplusplus
}
#endif
#endif /* OKVIS_B200_H */
