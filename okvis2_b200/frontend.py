"""Host-side mirror of the reference's front-end interface for the detect -> describe -> match path.

Same names, argument meaning and error behaviour as okvis::Frame / okvis::MultiFrame
(reference okvis_cv/include/okvis/Frame.hpp:247-265, MultiFrame.hpp:53-332) and okvis::Frontend
(reference okvis_frontend/include/okvis/Frontend.hpp:87-115,201-238; src/Frontend.cpp:221-269, 1515-2074).
All arithmetic of the path runs inside libokvis_b200.so (CUDA, sm_100a) through the C ABI; nothing here computes
features or distances on the host.
"""
import threading

import numpy as np

from . import lib as _l
from .lib import KP_DTYPE, CameraConfig, CameraModel, OkbError, check, ptr

import ctypes as C


class Frame:
    """okvis::Frame: image + keypoints + contiguous N x D descriptor matrix + landmark ids."""

    def __init__(self):
        self.image = None
        self.keypoints = np.zeros(0, KP_DTYPE)          # std::vector<cv::KeyPoint>
        self.descriptors = np.zeros((0, 64), np.uint8)  # cv::Mat N x D CV_8UC1, continuous
        self.landmarkIds = np.zeros(0, np.uint64)       # zero-filled after describe (Frame.hpp:170)
        self.backProjections = np.zeros((0, 3), np.float64)
        self.backProjectionsValid = np.zeros(0, np.uint8)
        self.extractionDirection = None                 # gravity in the camera frame as handed to the extractor (Frontend.cpp:245-251)

    def numKeypoints(self):
        return len(self.keypoints)

    def keypointDescriptor(self, k):
        return self.descriptors[k]

    def resetKeypoints(self, keypoints):
        self.keypoints = np.ascontiguousarray(keypoints, KP_DTYPE)
        self.landmarkIds = np.zeros(len(self.keypoints), np.uint64)

    def resetDescriptors(self, descriptors):
        self.descriptors = np.ascontiguousarray(descriptors, np.uint8)


class MultiFrame:
    """okvis::MultiFrame: one Frame per camera of the NCameraSystem (MultiFrame.hpp:327-331)."""

    def __init__(self, numCameras, timestamp=0.0, id=0):
        self.frames = [Frame() for _ in range(numCameras)]
        self.timestamp = timestamp
        self.id = id

    def numFrames(self):
        return len(self.frames)

    def setImage(self, cameraIdx, image):
        image = np.asarray(image)
        if image.dtype != np.uint8 or image.ndim != 2:
            raise OkbError(_l.OKB_ERR_ARGUMENT, "setImage: expected a single-channel u8 image")
        self.frames[cameraIdx].image = image

    def numKeypoints(self, cameraIdx=None):
        if cameraIdx is None:
            return sum(f.numKeypoints() for f in self.frames)
        return self.frames[cameraIdx].numKeypoints()

    def keypointDescriptor(self, cameraIdx, k):
        return self.frames[cameraIdx].descriptors[k]

    def landmarkId(self, cameraIdx, k):
        return int(self.frames[cameraIdx].landmarkIds[k])

    def setLandmarkId(self, cameraIdx, k, lmId):
        self.frames[cameraIdx].landmarkIds[k] = lmId


class Frontend:
    """okvis::Frontend restricted to the hot path. One CUDA context (okb_context_t) per instance."""

    def __init__(self, numCameras, width=752, height=480, device=0, max_batch=1, descriptor_bytes=64):
        self.numCameras = numCameras
        self._geom = [(width, height)] * numCameras if np.isscalar(width) else list(zip(width, height))
        # defaults of okvis::Frontend::Frontend (Frontend.cpp:133-147), threshold re-interpreted as the AGAST threshold
        self.briskDetectionOctaves_ = 0
        self.briskDetectionThreshold_ = 40.0
        self.briskDetectionAbsoluteThreshold_ = 200.0
        self.briskDetectionMaximumKeypoints_ = 450
        self.briskDescriptionRotationInvariance_ = True
        self.briskDescriptionScaleInvariance_ = False
        self.briskMatchingThreshold_ = 60.0
        self._device, self._max_batch, self._D = device, max_batch, descriptor_bytes
        self._ctx = None
        self._models = {}
        self._locks = [threading.Lock() for _ in range(numCameras)]  # featureDetectorMutexes_ (Frontend.cpp:226)
        self.initialiseBriskFeatureDetectors()

    # ---- setters (Frontend.hpp:201-238): each one re-creates the detectors/extractors, as in the reference
    def setBriskDetectionOctaves(self, octaves):
        self.briskDetectionOctaves_ = int(octaves); self.initialiseBriskFeatureDetectors()

    def setBriskDetectionThreshold(self, threshold):
        self.briskDetectionThreshold_ = float(threshold); self.initialiseBriskFeatureDetectors()

    def setBriskDetectionAbsoluteThreshold(self, threshold):
        self.briskDetectionAbsoluteThreshold_ = float(threshold); self.initialiseBriskFeatureDetectors()

    def setBriskDetectionMaximumKeypoints(self, maxKeypoints):
        self.briskDetectionMaximumKeypoints_ = int(maxKeypoints); self.initialiseBriskFeatureDetectors()

    def setBriskDescriptionRotationInvariance(self, invariance):
        if not invariance:
            raise OkbError(_l.OKB_ERR_UNSUPPORTED, "only rotation-invariant BRISK is implemented")
        self.briskDescriptionRotationInvariance_ = True

    def setBriskDescriptionScaleInvariance(self, invariance):
        self.briskDescriptionScaleInvariance_ = bool(invariance)

    def setBriskMatchingThreshold(self, threshold):
        self.briskMatchingThreshold_ = float(threshold)

    def configure(self, threshold=None, octaves=None, max_keypoints=None, matching_threshold=None, absolute_threshold=None):
        """Set several parameters with a single re-initialisation."""
        if threshold is not None: self.briskDetectionThreshold_ = float(threshold)
        if absolute_threshold is not None: self.briskDetectionAbsoluteThreshold_ = float(absolute_threshold)
        if octaves is not None: self.briskDetectionOctaves_ = int(octaves)
        if max_keypoints is not None: self.briskDetectionMaximumKeypoints_ = int(max_keypoints)
        if matching_threshold is not None: self.briskMatchingThreshold_ = float(matching_threshold)
        self.initialiseBriskFeatureDetectors()

    # ---- camera models (D4)
    MODELS = {"none": 0, "radialtangential": 1, "equidistant": 2}

    def setCameraModel(self, cameraIndex, distortion_type, focal_length, principal_point, distortion_coefficients):
        """Same fields as the `cameras:` entries of the okvis yaml files (config/euroc.yaml:9-12)."""
        m = CameraModel()
        m.model = self.MODELS[distortion_type]
        m.fu, m.fv = focal_length
        m.cu, m.cv = principal_point
        for i in range(4):
            m.k[i] = distortion_coefficients[i] if i < len(distortion_coefficients) else 0.0
        self._models[cameraIndex] = m
        check(_l.lib().okb_set_camera_model(self._ctx, cameraIndex, C.byref(m)))
        self._aware = set(getattr(self, "_aware", ())) - {cameraIndex}   # maps of the previous model are stale

    def computeBackProjections(self, frameOut, cameraIndex):
        """MultiFrame::computeBackProjections(im) (Frame.hpp:178-193): fills backProjections / backProjectionsValid."""
        fr = frameOut.frames[cameraIndex]
        n = len(fr.keypoints)
        rays = np.zeros((n, 3), np.float64); valid = np.zeros(n, np.uint8)
        kp = np.ascontiguousarray(fr.keypoints, KP_DTYPE)
        check(_l.lib().okb_back_project(self._ctx, cameraIndex, n, ptr(kp), ptr(rays), ptr(valid)))
        fr.backProjections, fr.backProjectionsValid = rays, valid
        return int(valid.sum())


    # ---- D5 / rig overlaps
    def cameraAwarenessMaps(self, cameraIndex):
        """PinholeCamera::initialiseCameraAwarenessMaps (PinholeCamera.hpp:179-208) for the camera model set with setCameraModel:
        (rays H x W x 3 float32, imageJacobians H x W x 6 float32). The maps also stay on the device."""
        w, h = self._geom[cameraIndex]
        rays = np.zeros((h, w, 3), np.float32); jac = np.zeros((h, w, 6), np.float32)
        check(_l.lib().okb_camera_awareness_maps(self._ctx, cameraIndex, ptr(rays), ptr(jac)))
        self._aware = set(getattr(self, "_aware", ())) | {cameraIndex}
        return rays, jac

    def computeOverlaps(self, models, intrinsics, widths, heights, C_rel, masks=False):
        """NCameraSystem::computeOverlaps (NCameraSystem.cpp:48-118). models: 0 none / 1 radtan / 2 equidistant; intrinsics: per
        camera fu fv cu cv k0..k3; C_rel[s][c] = (T_SC[s]^-1 * T_SC[c]).C(). Returns hasOverlap (n x n bool) [and the masks]."""
        n = len(models)
        ms = (CameraModel * n)()
        for i in range(n):
            ms[i].model = int(models[i]); ms[i].fu, ms[i].fv, ms[i].cu, ms[i].cv = (float(x) for x in intrinsics[i][:4])
            for k in range(4):
                ms[i].k[k] = float(intrinsics[i][4 + k]) if 4 + k < len(intrinsics[i]) else 0.0
        W = np.ascontiguousarray(widths, np.int32); H = np.ascontiguousarray(heights, np.int32)
        Cr = np.ascontiguousarray(np.asarray(C_rel, np.float64).reshape(n, n, 9))
        out = np.zeros((n, n), np.uint8)
        mats = ptrs = None
        if masks:
            mats = [[np.zeros((int(H[c]), int(W[c])), np.uint8) for c in range(n)] for _ in range(n)]
            ptrs = (C.c_void_p * (n * n))(*[mats[s][c].ctypes.data for s in range(n) for c in range(n)])
        check(_l.lib().okb_compute_overlaps(self._ctx, n, ms, ptr(W), ptr(H), ptr(Cr), ptr(out), ptrs))
        return (out.astype(bool), mats) if masks else out.astype(bool)

    # ---- P1: landmark-candidate preparation (the loop of Frontend::matchToMap before the matching threads start)
    def configureFeatureStore(self, n_slots, D=64):
        """Device-resident store of the descriptors / back-projections of the multiframes in the window."""
        check(_l.lib().okb_store_configure(self._ctx, int(n_slots), self.numCameras, int(D)))
        self._store_D = int(D)

    def storeFrame(self, slot, cameraIndex, descriptors, backProjections):
        d = np.ascontiguousarray(descriptors, np.uint8); r = np.ascontiguousarray(backProjections, np.float64)
        check(_l.lib().okb_store_frame(self._ctx, int(slot), int(cameraIndex), len(d), ptr(d), ptr(r)))

    def storeLastFrame(self, slot, cameraIndex, batch_index=0):
        check(_l.lib().okb_store_frame_from_last(self._ctx, int(slot), int(cameraIndex), int(batch_index)))

    def prepareLandmarksToMatch(self, cameraIndex, T_WC1, T_CW1, width, height, hp_W, quality, obs_begin, obs, T_WC_old,
                                reprThreshold=20.0, exclusive=False):
        """landmarksToMatch / descriptorPool of one camera (Frontend.cpp:1196-1360). Poses are (C 3x3, r 3) pairs packed
        as 12 doubles; returns a dict with the packed pool (cand_desc, cand_lm, lm_proj, lm_is3d, ...)."""
        v = _l.PrepareView()
        v.T_WC1[:] = list(np.asarray(T_WC1, np.float64).ravel()); v.T_CW1[:] = list(np.asarray(T_CW1, np.float64).ravel())
        v.model = self._models[cameraIndex]; v.width, v.height = int(width), int(height)
        v.repr_threshold = float(reprThreshold); v.exclusive = 1 if exclusive else 0
        hp_W = np.ascontiguousarray(hp_W, np.float64); quality = np.ascontiguousarray(quality, np.float64)
        obs_begin = np.ascontiguousarray(obs_begin, np.int32); obs = np.ascontiguousarray(obs, np.int32).reshape(-1, 3)
        T_WC_old = np.ascontiguousarray(T_WC_old, np.float64)
        n_lm = len(quality); D = self._store_D
        out = dict(lm=np.zeros(n_lm, np.int32), lm_proj=np.zeros((n_lm, 2)), lm_is3d=np.zeros(n_lm, np.uint8), p_W=np.zeros((n_lm, 3)),
                   desc_begin=np.zeros(n_lm + 1, np.int32), cand_desc=np.zeros((2 * n_lm, D), np.uint8),
                   cand_lm=np.zeros(2 * n_lm, np.int32), e_W=np.zeros((2 * n_lm, 3)), r_W=np.zeros((2 * n_lm, 3)),
                   kid=np.zeros((2 * n_lm, 3), np.int32))
        n_out = C.c_int32(); n_rows = C.c_int32()
        check(_l.lib().okb_prepare_landmarks(self._ctx, C.byref(v), n_lm, ptr(hp_W), ptr(quality), ptr(obs_begin), len(obs), ptr(obs),
                                             ptr(T_WC_old), C.byref(n_out), C.byref(n_rows), ptr(out["lm"]), ptr(out["lm_proj"]),
                                             ptr(out["lm_is3d"]), ptr(out["p_W"]), ptr(out["desc_begin"]), ptr(out["cand_desc"]),
                                             ptr(out["cand_lm"]), ptr(out["e_W"]), ptr(out["r_W"]), ptr(out["kid"])))
        nl, nr = n_out.value, n_rows.value
        for k in ("lm", "lm_proj", "lm_is3d", "p_W"):
            out[k] = out[k][:nl]
        out["desc_begin"] = out["desc_begin"][:nl + 1]
        for k in ("cand_desc", "cand_lm", "e_W", "r_W", "kid"):
            out[k] = out[k][:nr]
        return out

    def matchToMap(self, cameraIndex, T_WC1, T_CW1, width, height, hp_W, quality, obs_begin, obs, T_WC_old,
                   reprThreshold=20.0, exclusive=False):
        """The device part of Frontend::matchToMap for one camera (Frontend.cpp:1196-1408): prepare the landmark pool
        (P1), then match the keypoints of the LAST detectAndDescribe of this camera against it (M1) without the pool leaving
        the device. Returns (distances, landmark index per keypoint row into the caller's hp_W order or -1, the prepared pool);
        the arrays have one entry per keypoint ROW of the device block (capacity), the first numKeypoints are the frame's."""
        pool = self.prepareLandmarksToMatch(cameraIndex, T_WC1, T_CW1, width, height, hp_W, quality, obs_begin, obs, T_WC_old,
                                            reprThreshold, exclusive)
        L = _l.lib()
        p = [C.c_void_p() for _ in range(4)]; nc = C.c_int32(); nl = C.c_int32()
        check(L.okb_prepared_device(self._ctx, *[C.byref(x) for x in p], C.byref(nc), C.byref(nl)))
        d_kp = C.c_void_p(); d_desc = C.c_void_p(); d_cnt = C.c_void_p(); cap = C.c_int(0)
        check(L.okb_device_features(self._ctx, cameraIndex, C.byref(d_kp), C.byref(d_desc), C.byref(d_cnt), C.byref(cap)))
        import torch   # device buffers for the M1 outputs (plumbing only)
        dist = torch.zeros(cap.value, dtype=torch.int32, device=f"cuda:{self._device}")
        lm = torch.zeros(cap.value, dtype=torch.int32, device=f"cuda:{self._device}")
        check(L.okb_match_map3d_device(self._ctx, cameraIndex, self._store_D, 1, nc.value, p[0], p[1], nl.value, p[2], p[3], float(reprThreshold),
                                       int(self.briskMatchingThreshold_), dist.data_ptr(), lm.data_ptr()))
        check(L.okb_sync(self._ctx))
        dist = dist.cpu().numpy().view(np.uint32); lm = lm.cpu().numpy()   # rows >= the frame's keypoint count: "no match"
        idx = np.full(len(lm), -1, np.int32)
        if len(pool["lm"]):
            hit = lm >= 0
            idx[hit] = pool["lm"][lm[hit]]
        return dist, idx, pool

    # ---- K1: keyframe-overlap masks (Frontend::doWeNeedANewKeyframe, ViSlamBackend::overlapFraction)
    keyframeInsertionOverlapThreshold_ = np.float32(0.55)   # Frontend.cpp:145
    kptrad = 0.09                                           # Frontend.cpp:104, ViSlamBackend.hpp:684

    def _overlap_counts(self, views, kptrad=None):
        """views: (image_rows, image_cols, xy (n, 2) float32, matched (n,) bool). One device call for all of them."""
        vs = (_l.OverlapView * len(views))()
        xy, m, first = [], [], 0
        for i, (r, c, p, f) in enumerate(views):
            vs[i] = _l.OverlapView(int(r), int(c), first, len(p)); first += len(p)
            xy.append(np.asarray(p, np.float32).reshape(-1, 2)); m.append(np.asarray(f, np.uint8))
        xy = np.ascontiguousarray(np.concatenate(xy)) if views else np.zeros((0, 2), np.float32)
        m = np.ascontiguousarray(np.concatenate(m)) if views else np.zeros(0, np.uint8)
        inter = np.zeros(len(views), np.int32); uni = np.zeros(len(views), np.int32)
        check(_l.lib().okb_overlap_counts(self._ctx, len(views), vs, len(m), ptr(xy), ptr(m), self.kptrad if kptrad is None else kptrad,
                                          ptr(inter), ptr(uni)))
        return inter, uni

    @staticmethod
    def _views_of(mf, matched_of):
        out = []
        for fr in mf.frames:
            xy = np.stack([fr.keypoints["x"], fr.keypoints["y"]], 1) if len(fr.keypoints) else np.zeros((0, 2), np.float32)
            out.append((fr.image.shape[0], fr.image.shape[1], xy, matched_of(fr)))
        return out

    def doWeNeedANewKeyframe(self, numFramesInEstimator, currentFrame, otherFrames, isInitialized=True):
        """Frontend::doWeNeedANewKeyframe (Frontend.cpp:1058-1167). otherFrames = the multiframes of estimator.keyFrames()
        + loopClosureFrames() + keyframes in the IMU window (the caller's bookkeeping, :1105-1115)."""
        if numFramesInEstimator < 4:
            return True
        if not isInitialized:
            return False
        lmIds = set()
        for fr in currentFrame.frames:
            lmIds.update(int(x) for x in fr.landmarkIds if x != 0)
        ids = np.array(sorted(lmIds), np.uint64)
        views = self._views_of(currentFrame, lambda fr: fr.landmarkIds != 0)
        for mf in otherFrames:
            views += self._views_of(mf, lambda fr: (fr.landmarkIds != 0) & np.isin(fr.landmarkIds, ids))
        inter, uni = self._overlap_counts(views)
        nc = currentFrame.numFrames()
        with np.errstate(divide="ignore", invalid="ignore"):
            overlap = np.float64(inter[:nc].sum()) / np.float64(uni[:nc].sum())
            overlapOthers = np.float64(0.0)
            pos = nc
            for mf in otherFrames:
                n = mf.numFrames()
                x = np.float64(inter[pos:pos + n].sum()) / np.float64(uni[pos:pos + n].sum()); pos += n
                overlapOthers = x if overlapOthers < x else overlapOthers   # std::max(overlapOthers, x)
        overlap = overlap if overlap < overlapOthers else overlapOthers     # std::min(overlapOthers, overlap)
        if currentFrame.numKeypoints() < 7 * nc:
            return False
        return not (np.float32(overlap) > self.keyframeInsertionOverlapThreshold_)

    def overlapFraction(self, frameA, frameB, kptradius=0.09):
        """ViSlamBackend::overlapFraction (ViSlamBackend.cpp:2341-2426)."""
        lm = [set(int(x) for fr in f.frames for x in fr.landmarkIds if x != 0) for f in (frameA, frameB)]
        matches = np.array(sorted(lm[0] & lm[1]), np.uint64)
        if len(matches) == 0:
            return 0.0
        views = []
        for f in (frameA, frameB):
            views += self._views_of(f, lambda fr: np.isin(fr.landmarkIds, matches))
        inter, uni = self._overlap_counts(views, kptradius)
        n = frameA.numFrames()
        with np.errstate(divide="ignore", invalid="ignore"):
            o = [np.float64(inter[i * n:(i + 1) * n].sum()) / np.float64(uni[i * n:(i + 1) * n].sum()) for i in range(2)]
        return float(o[1] if o[1] < o[0] else o[0])   # std::min(overlap[0], overlap[1])

    # ---- B1: DBoW2 vocabulary descent with the FBrisk distance
    def loadVocabulary(self, k, L, node_id, parent_id, weight, descriptors, word_id, word_node):
        """The node / word lists of a DBoW2 vocabulary file, nodes in file order (resources/small_voc.yml.gz)."""
        a = lambda x, t: np.ascontiguousarray(x, t)
        d = a(descriptors, np.uint8)
        ni, pi, w, wi, wn = a(node_id, np.int32), a(parent_id, np.int32), a(weight, np.float64), a(word_id, np.int32), a(word_node, np.int32)
        check(_l.lib().okb_bow_load(self._ctx, d.shape[1], int(k), int(L), len(ni), ptr(ni), ptr(pi), ptr(w), ptr(d), len(wi), ptr(wi), ptr(wn)))

    def bowTransform(self, descriptors, levelsup=0):
        """TemplatedVocabulary::transform for every row: (word ids, weights, node ids `levelsup` levels above the leaves)."""
        d = np.ascontiguousarray(descriptors, np.uint8)
        word = np.zeros(len(d), np.int32); weight = np.zeros(len(d), np.float64); node = np.zeros(len(d), np.int32)
        check(_l.lib().okb_bow_transform(self._ctx, len(d), ptr(d), int(levelsup), ptr(word), ptr(weight), ptr(node)))
        return word, weight, node

    def bowTransformImage(self, descriptors, levelsup=0):
        """TemplatedVocabulary::transform(features, bowVector, featureVector, levelsup) for TF_IDF weighting + L1 scoring: the
        per-feature descents run on the device (okb_bow_transform), the accumulation into the sparse vectors -- DBoW2's own
        bookkeeping -- on the host in feature order. Returns (sorted [(word, weight)], {node: [feature indices]})."""
        word, weight, node = self.bowTransform(descriptors, levelsup)
        v, fv = {}, {}
        for i in range(len(word)):
            if weight[i] > 0:
                w = int(word[i])
                v[w] = v.get(w, 0.0) + float(weight[i])
                fv.setdefault(int(node[i]), []).append(i)
        items = sorted(v.items())
        norm = 0.0
        for _, w in items:
            norm += abs(w)
        if norm > 0.0:
            items = [(k, w / norm) for k, w in items]
        return items, fv

    @staticmethod
    def bowScoreL1(v1, v2):
        """DBoW2::L1Scoring::score of two normalised BowVectors."""
        score, i, j = 0.0, 0, 0
        while i < len(v1) and j < len(v2):
            if v1[i][0] == v2[j][0]:
                score += abs(v1[i][1] - v2[j][1]) - abs(v1[i][1]) - abs(v2[j][1]); i += 1; j += 1
            elif v1[i][0] < v2[j][0]:
                i += 1
            else:
                j += 1
        return -score / 2.0

    def initialiseBriskFeatureDetectors(self):
        """Frontend::initialiseBriskFeatureDetectors (Frontend.cpp:2398-2417): (re)create the per-camera objects."""
        self.close()
        cfgs = (CameraConfig * self.numCameras)()
        for i, (w, h) in enumerate(self._geom):
            if self._D == 48:   # the reference's own pair (Frontend.cpp:2406-2409): (uniformity radius, octaves, absolute threshold, max keypoints)
                cfgs[i] = CameraConfig(w, h, int(self.briskDetectionAbsoluteThreshold_), self.briskDetectionOctaves_,
                                       self.briskDetectionMaximumKeypoints_, 48, self._max_batch, 1.0, float(self.briskDetectionThreshold_))
            else:               # AGAST + BRISK-512: briskDetectionThreshold_ re-interpreted as the AGAST threshold
                cfgs[i] = CameraConfig(w, h, int(self.briskDetectionThreshold_), self.briskDetectionOctaves_,
                                       self.briskDetectionMaximumKeypoints_, self._D, self._max_batch, 1.0, 0.0)
        ctx = C.c_void_p()
        check(_l.lib().okb_create(self._device, self.numCameras, cfgs, C.byref(ctx)))
        self._ctx = ctx
        for i, m in getattr(self, "_models", {}).items():   # the camera models survive a re-initialisation
            check(_l.lib().okb_set_camera_model(self._ctx, i, C.byref(m)))
        for i in getattr(self, "_aware", ()):               # and so do the camera-awareness maps the D = 48 extractor reads
            if i in self._models:
                check(_l.lib().okb_camera_awareness_maps(self._ctx, i, None, None))

    def close(self):
        if getattr(self, "_ctx", None):
            _l.lib().okb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def ctx(self):
        return self._ctx

    # ---- detect + describe
    def _capacity(self, cam):
        n = self.briskDetectionMaximumKeypoints_
        return n if n > 0 else 16384

    def detectAndDescribe(self, cameraIndex, frameOut, T_WC=None, keypoints=None):
        """Frontend::detectAndDescribe (Frontend.cpp:221-269). Thread-safe per camera; returns True."""
        if keypoints is not None:  # Frontend.cpp:229: "external keypoints currently not supported"
            raise OkbError(_l.OKB_ERR_UNSUPPORTED, "external keypoints currently not supported")
        with self._locks[cameraIndex]:
            fr = frameOut.frames[cameraIndex]
            if T_WC is not None:
                # ExtractionDirection == gravity direction in camera frame (Frontend.cpp:245-251); T_WC: 3x3 / 3x4 / 4x4 / 12 values
                Cm = np.ascontiguousarray(np.asarray(T_WC, np.float64).reshape(-1, 4 if np.size(T_WC) in (12, 16) else 3)[:3, :3])
                check(_l.lib().okb_set_extraction_direction(self._ctx, cameraIndex, ptr(Cm)))
                d = np.zeros(3, np.float32)
                check(_l.lib().okb_get_extraction_direction(self._ctx, cameraIndex, ptr(d)))
                fr.extractionDirection = d
            img = np.ascontiguousarray(fr.image)
            w, h = self._geom[cameraIndex]
            if img.shape != (h, w):
                raise OkbError(_l.OKB_ERR_ARGUMENT, f"image is {img.shape}, camera {cameraIndex} is {(h, w)}")
            cap = self._capacity(cameraIndex)
            kp = np.zeros(cap, KP_DTYPE)
            desc = np.zeros((cap, self._D), np.uint8)
            n = C.c_int(0)
            check(_l.lib().okb_detect_describe(self._ctx, cameraIndex, img.ctypes.data, img.strides[0], kp.ctypes.data,
                                               desc.ctypes.data, cap, C.byref(n)))
            fr.keypoints = kp[:n.value].copy()
            fr.descriptors = np.ascontiguousarray(desc[:n.value])
            fr.landmarkIds = np.zeros(n.value, np.uint64)  # Frame::describe zero-fills them (Frame.hpp:170)
        return True

    def detectAndDescribeBatch(self, cameraIndex, images):
        """Batch replay: `images` is n x H x W u8. Returns a list of (keypoints, descriptors)."""
        images = np.ascontiguousarray(images, np.uint8)
        n_frames = images.shape[0]
        cap = self._capacity(cameraIndex)
        kp = np.zeros((n_frames, cap), KP_DTYPE)
        desc = np.zeros((n_frames, cap, self._D), np.uint8)
        n = np.zeros(n_frames, np.int32)
        with self._locks[cameraIndex]:
            check(_l.lib().okb_detect_describe_batch(self._ctx, cameraIndex, n_frames, images.ctypes.data,
                                                     images.strides[1], kp.ctypes.data, desc.ctypes.data, cap,
                                                     n.ctypes.data))
        return [(kp[b, :n[b]].copy(), desc[b, :n[b]].copy()) for b in range(n_frames)]

    def layers(self, cameraIndex, frame=0):
        """Pyramid layer images and thresholded score maps of the last detect call (test hook)."""
        L = _l.lib()
        out = []
        for i in range(L.okb_num_layers(self._ctx, cameraIndex)):
            w, h, s, o = C.c_int(), C.c_int(), C.c_float(), C.c_float()
            check(L.okb_layer_info(self._ctx, cameraIndex, i, C.byref(w), C.byref(h), C.byref(s), C.byref(o)))
            img = np.zeros((h.value, w.value), np.uint8)
            sc = np.zeros((h.value, w.value), np.uint8)
            check(L.okb_fetch_layer(self._ctx, cameraIndex, frame, i, img.ctypes.data, sc.ctypes.data))
            out.append((img, sc, s.value, o.value))
        return out

    # ---- matchers (thin, array-in / array-out forms of the five loops)
    def matchToMapByThread(self, kp_desc, kp_xy, kp_use, cand_desc, cand_lm, lm_proj, lm_is3d, use_imu=True):
        """Frontend::matchToMapByThread (Frontend.cpp:1515-1590) for all keypoints of one camera."""
        kp_desc = np.ascontiguousarray(kp_desc, np.uint8); cand_desc = np.ascontiguousarray(cand_desc, np.uint8)
        n, D = kp_desc.shape[0], kp_desc.shape[1] if kp_desc.ndim == 2 else self._D
        kp_xy = np.ascontiguousarray(kp_xy, np.float64); cand_lm = np.ascontiguousarray(cand_lm, np.int32)
        lm_proj = np.ascontiguousarray(lm_proj, np.float64); lm_is3d = np.ascontiguousarray(lm_is3d, np.uint8)
        kp_use = None if kp_use is None else np.ascontiguousarray(kp_use, np.uint8)
        dist = np.zeros(n, np.uint32); lm = np.zeros(n, np.int32)
        thr = 20.0 if use_imu else 150.0  # Frontend.cpp:1530
        check(_l.lib().okb_match_map3d(self._ctx, D, n, ptr(kp_desc), ptr(kp_xy), ptr(kp_use), len(cand_desc),
                                       ptr(cand_desc), ptr(cand_lm), len(lm_is3d), ptr(lm_proj), ptr(lm_is3d), thr,
                                       int(self.briskMatchingThreshold_), ptr(dist), ptr(lm)))
        return dist, lm

    def matchToMapByThreadUnitialised(self, kp_desc, kp_e_W, kp_use, cand_desc, cand_lm, cand_e_W, cand_r_W, lm_is3d,
                                      r_WC1, focalLength, kp_prev_lm=None):
        """Frontend::matchToMapByThreadUnitialised (Frontend.cpp:1594-1720)."""
        kp_desc = np.ascontiguousarray(kp_desc, np.uint8); cand_desc = np.ascontiguousarray(cand_desc, np.uint8)
        n, D = kp_desc.shape
        kp_e_W = np.ascontiguousarray(kp_e_W, np.float64); cand_e_W = np.ascontiguousarray(cand_e_W, np.float64)
        cand_r_W = np.ascontiguousarray(cand_r_W, np.float64); cand_lm = np.ascontiguousarray(cand_lm, np.int32)
        lm_is3d = np.ascontiguousarray(lm_is3d, np.uint8); r = np.ascontiguousarray(r_WC1, np.float64)
        kp_use = None if kp_use is None else np.ascontiguousarray(kp_use, np.uint8)
        kp_prev_lm = None if kp_prev_lm is None else np.ascontiguousarray(kp_prev_lm, np.int32)
        dist = np.zeros(n, np.uint32); lm = np.zeros(n, np.int32); hp = np.zeros((n, 4), np.float64); ctr = C.c_int32(0)
        check(_l.lib().okb_match_map_uninit(self._ctx, D, n, ptr(kp_desc), ptr(kp_e_W), ptr(kp_use), ptr(kp_prev_lm),
                                            len(cand_desc), ptr(cand_desc), ptr(cand_lm), ptr(cand_e_W), ptr(cand_r_W),
                                            len(lm_is3d), ptr(lm_is3d), ptr(r), 1.0 / focalLength,
                                            int(self.briskMatchingThreshold_), ptr(dist), ptr(lm), ptr(hp), C.byref(ctr)))
        return dist, lm, hp, ctr.value

    def _stereo(self, fn, desc0, use0, e0, sof0, desc1, valid1, e1, sof1, r0, r1, T0, T1):
        desc0 = np.ascontiguousarray(desc0, np.uint8); desc1 = np.ascontiguousarray(desc1, np.uint8)
        n0, D = desc0.shape
        a = lambda x, t: None if x is None else np.ascontiguousarray(x, t)
        use0, valid1 = a(use0, np.uint8), a(valid1, np.uint8)
        e0, e1, sof0, sof1 = a(e0, np.float64), a(e1, np.float64), a(sof0, np.float64), a(sof1, np.float64)
        r0, r1, T0, T1 = a(r0, np.float64), a(r1, np.float64), a(T0, np.float64), a(T1, np.float64)
        k1 = np.zeros(n0, np.int32); dist = np.zeros(n0, np.uint32); hp = np.zeros((n0, 4), np.float64)
        init = np.zeros(n0, np.uint8)
        thr = int(self.briskMatchingThreshold_)
        if fn == "motion":
            check(_l.lib().okb_match_motion_stereo(self._ctx, D, n0, ptr(desc0), ptr(use0), ptr(e0), ptr(sof0), len(desc1),
                                                   ptr(desc1), ptr(valid1), ptr(e1), ptr(r0), ptr(r1), ptr(T0), ptr(T1), thr,
                                                   ptr(k1), ptr(dist), ptr(hp), ptr(init)))
        else:
            check(_l.lib().okb_match_stereo(self._ctx, D, n0, ptr(desc0), ptr(use0), ptr(e0), ptr(sof0), len(desc1),
                                            ptr(desc1), ptr(valid1), ptr(e1), ptr(sof1), ptr(r0), ptr(r1), ptr(T0), ptr(T1),
                                            thr, ptr(k1), ptr(dist), ptr(hp), ptr(init)))
        return k1, dist, hp, init

    def matchMotionStereo(self, desc0, use0, e0_W, size_over_f0, desc1, valid1, e1_W, r_WC0, r_WC1, T_CW0, T_CW1):
        """Worker loop of Frontend::matchMotionStereo (Frontend.cpp:1809-1907) for one (older frame, camera)."""
        return self._stereo("motion", desc0, use0, e0_W, size_over_f0, desc1, valid1, e1_W, None, r_WC0, r_WC1, T_CW0, T_CW1)

    def matchStereo(self, desc0, valid0, e0_W, size_over_f0, desc1, valid1, e1_W, size_over_f1, r_WC0, r_WC1, T_CW0, T_CW1):
        """k0/k1 loops of Frontend::matchStereo (Frontend.cpp:2016-2074) for one overlapping camera pair."""
        return self._stereo("stereo", desc0, valid0, e0_W, size_over_f0, desc1, valid1, e1_W, size_over_f1, r_WC0, r_WC1,
                            T_CW0, T_CW1)

    def matchToMapBatch(self, cameraIndex, n_frames, cand_desc, cand_lm, lm_proj, lm_is3d, use_imu=True, out=None):
        """Frontend::matchToMapByThread for every frame of the last detectAndDescribeBatch of `cameraIndex` (the queries
        stay on the device); lm_proj is n_frames x n_lm x 2. `out` = (dist, lm) arrays to fill (e.g. page-locked)."""
        cap = self._capacity(cameraIndex)
        cand_desc = np.ascontiguousarray(cand_desc, np.uint8); cand_lm = np.ascontiguousarray(cand_lm, np.int32)
        lm_proj = np.ascontiguousarray(lm_proj, np.float64); lm_is3d = np.ascontiguousarray(lm_is3d, np.uint8)
        assert lm_proj.shape == (n_frames, len(lm_is3d), 2)
        dist, lm = out if out is not None else (np.zeros((n_frames, cap), np.uint32), np.zeros((n_frames, cap), np.int32))
        with self._locks[cameraIndex]:
            check(_l.lib().okb_match_map3d_batch(self._ctx, cameraIndex, cand_desc.shape[1], n_frames, len(cand_desc), ptr(cand_desc), ptr(cand_lm),
                                                 len(lm_is3d), ptr(lm_proj), ptr(lm_is3d), 20.0 if use_imu else 150.0,
                                                 int(self.briskMatchingThreshold_), dist.shape[1], ptr(dist), ptr(lm)))
        return dist, lm

    def matchStereoBatch(self, cam0, cam1, n_frames, C_WC0, r_WC0, C_WC1, r_WC1, out=None):
        """Frontend::matchStereo k0/k1 loops for every frame of the last detectAndDescribeBatch of the two cameras."""
        cap = self._capacity(cam0)
        a = lambda x: np.ascontiguousarray(x, np.float64)
        C0, r0, C1, r1 = a(C_WC0), a(r_WC0), a(C_WC1), a(r_WC1)
        if out is None:
            out = (np.zeros((n_frames, cap), np.int32), np.zeros((n_frames, cap), np.uint32),
                   np.zeros((n_frames, cap, 4), np.float64), np.zeros((n_frames, cap), np.uint8))
        k1, dist, hp, init = out
        check(_l.lib().okb_match_stereo_batch(self._ctx, cam0, cam1, n_frames, ptr(C0), ptr(r0), ptr(C1), ptr(r1),
                                              int(self.briskMatchingThreshold_), k1.shape[1], ptr(k1), ptr(dist), ptr(hp), ptr(init)))
        return k1, dist, hp, init

    def verifyRecognisedPlaceMatch(self, lm_offsets, lm_desc, kp_desc):
        """Descriptor matching loop of Frontend::verifyRecognisedPlace (Frontend.cpp:329-355)."""
        lm_offsets = np.ascontiguousarray(lm_offsets, np.int32); lm_desc = np.ascontiguousarray(lm_desc, np.uint8)
        kp_desc = np.ascontiguousarray(kp_desc, np.uint8)
        n_lm = len(lm_offsets) - 1
        D = kp_desc.shape[1]
        k = np.zeros(n_lm, np.int32); dist = np.zeros(n_lm, np.uint32)
        check(_l.lib().okb_match_place(self._ctx, D, n_lm, ptr(lm_offsets), ptr(lm_desc), len(kp_desc), ptr(kp_desc),
                                       int(self.briskMatchingThreshold_), ptr(k), ptr(dist)))
        return k, dist

    def hammingMatrix(self, a, b):
        """brisk::Hamming::PopcntofXORed over all pairs."""
        a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
        out = np.zeros((len(a), len(b)), np.uint16)
        check(_l.lib().okb_hamming_matrix(self._ctx, a.shape[1], len(a), ptr(a), len(b), ptr(b), ptr(out)))
        return out
