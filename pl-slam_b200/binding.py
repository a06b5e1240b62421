"""ctypes binding of libplslam_b200.so — Python mirror of the reference's operator classes.

Class and method names follow the reference (ORBextractor, LINEextractor, ORBmatcher, LSDmatcher, Optimizer);
see include/plslam_b200.h for the C ABI each method calls.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PLSLAM_B200_LIB") or os.path.join(_HERE, "libplslam_b200.so")   # env override: A/B builds in tools/

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])


class PLError(RuntimeError):
    pass


class PLOrbConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("nfeatures", C.c_int), ("scale_factor", C.c_float),
                ("nlevels", C.c_int), ("ini_th_fast", C.c_int), ("min_th_fast", C.c_int), ("max_batch", C.c_int),
                ("cell_slot_cap", C.c_int)]


_lib = None
vp = C.c_void_p


def lib():
    """Load the CUDA library; there is deliberately no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise PLError(f"{LIB_PATH} is missing: run __graft_entry__.build() (no CPU fallback exists)")
        L = C.CDLL(LIB_PATH)
        L.pl_last_error.restype = C.c_char_p
        L.pl_launch_count.restype = C.c_ulonglong
        L.pl_orb_create.argtypes = [C.POINTER(PLOrbConfig), C.POINTER(vp)]
        L.pl_orb_destroy.argtypes = [vp]
        L.pl_orb_capacity.argtypes = [vp]
        L.pl_orb_tables.argtypes = [vp] * 8
        L.pl_orb_extract.argtypes = [vp, vp, C.c_int, vp, vp, vp]
        L.pl_orb_extract_batch.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_int, vp, vp, vp]
        L.pl_orb_extract_batch_dev.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_int, vp, vp, vp, vp]
        L.pl_orb_get_level.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int]
        L.pl_orb_debug_candidates.argtypes = [vp, C.c_int, C.c_int, vp, C.c_int]
        _lib = L
    return _lib


def check(rc):
    if rc < 0:
        raise PLError(f"plslam_b200 error {rc}: {lib().pl_last_error().decode()}")
    return rc


def _p(a):
    return a.ctypes.data_as(vp) if a is not None else None


def launch_count():
    return int(lib().pl_launch_count())


class ORBextractor:
    """Mirror of ORB_SLAM2::ORBextractor (reference include/ORBextractor.h:45-111).

    `__call__(image)` == operator()(image, mask, keypoints, descriptors); the mask is ignored as in the
    reference.  `extract_batch` runs B frames per launch sequence.
    """

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, width=640, height=480, max_batch=1,
                 cell_slot_cap=0):
        self.cfg = PLOrbConfig(width, height, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_batch,
                               cell_slot_cap)
        self._h = vp()
        check(lib().pl_orb_create(C.byref(self.cfg), C.byref(self._h)))
        self.capacity = check(lib().pl_orb_capacity(self._h))
        n = nlevels
        self._scale, self._inv, self._s2, self._is2 = (np.zeros(n, np.float32) for _ in range(4))
        self.mnFeaturesPerLevel = np.zeros(n, np.int32)
        self.level_w, self.level_h = np.zeros(n, np.int32), np.zeros(n, np.int32)
        check(lib().pl_orb_tables(self._h, _p(self._scale), _p(self._inv), _p(self._s2), _p(self._is2),
                                  _p(self.mnFeaturesPerLevel), _p(self.level_w), _p(self.level_h)))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().pl_orb_destroy(self._h)
            self._h = vp()

    def GetLevels(self): return self.cfg.nlevels
    def GetScaleFactor(self): return self.cfg.scale_factor
    def GetScaleFactors(self): return self._scale
    def GetInverseScaleFactors(self): return self._inv
    def GetScaleSigmaSquares(self): return self._s2
    def GetInverseScaleSigmaSquares(self): return self._is2

    def __call__(self, image, mask=None):
        image = np.ascontiguousarray(image, np.uint8)
        if image.size == 0:
            return np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        assert image.shape == (self.cfg.height, self.cfg.width)
        kps = np.zeros(self.capacity, KP_DTYPE)
        desc = np.zeros((self.capacity, 32), np.uint8)
        n = C.c_int(0)
        check(lib().pl_orb_extract(self._h, _p(image), image.strides[0], _p(kps), _p(desc), C.byref(n)))
        return kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images):
        images = np.ascontiguousarray(images, np.uint8)
        B = images.shape[0]
        kps = np.zeros((B, self.capacity), KP_DTYPE)
        desc = np.zeros((B, self.capacity, 32), np.uint8)
        n = np.zeros(B, np.int32)
        check(lib().pl_orb_extract_batch(self._h, _p(images), images.strides[1], images.strides[0], B, _p(kps),
                                         _p(desc), _p(n)))
        return kps, desc, n

    def extract_batch_dev(self, img_ptr, stride, frame_stride, B, kps_ptr, desc_ptr, n_ptr, stream=None):
        """Device-pointer variant (asynchronous)."""
        check(lib().pl_orb_extract_batch_dev(self._h, img_ptr, stride, frame_stride, B, kps_ptr, desc_ptr, n_ptr,
                                             stream))

    def mvImagePyramid(self, level, frame=0, with_border=False):
        w, h = int(self.level_w[level]), int(self.level_h[level])
        if with_border:
            w, h = w + 38, h + 38
        out = np.zeros((h, w), np.uint8)
        check(lib().pl_orb_get_level(self._h, frame, level, _p(out), int(with_border)))
        return out

    def debug_candidates(self, level, frame=0):
        n = check(lib().pl_orb_debug_candidates(self._h, frame, level, None, 0))
        out = np.zeros(max(n, 1), KP_DTYPE)
        check(lib().pl_orb_debug_candidates(self._h, frame, level, _p(out), n))
        return out[:n]


# ---------------------------------------------------------------------------------------------- matching
def _u8(a):
    return np.ascontiguousarray(a, np.uint8)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def _i32(a):
    return np.ascontiguousarray(a, np.int32)


def frame_assign_grid(keys_un, bounds):
    """Frame::AssignFeaturesToGrid (reference src/Frame.cc:278-294) -> CSR (cell_start[3073], cell_items[n])."""
    keys = np.ascontiguousarray(keys_un); b = _f32(bounds)
    start = np.zeros(64 * 48 + 1, np.int32); items = np.zeros(max(len(keys), 1), np.int32)
    check(lib().pl_frame_assign_grid(_p(keys), len(keys), _p(b), _p(start), _p(items)))
    return start, items[:start[-1]]


class ORBmatcher:
    """Mirror of ORB_SLAM2::ORBmatcher (reference include/ORBmatcher.h:37-102) on flat frame arrays."""
    TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30

    def __init__(self, nnratio=0.6, checkOri=True):
        self.mfNNratio, self.mbCheckOrientation = float(nnratio), bool(checkOri)

    @staticmethod
    def DescriptorDistance(a, b):
        a, b = _u8(a).reshape(-1, 32), _u8(b).reshape(-1, 32)
        out = np.zeros(len(a), np.int32)
        check(lib().pl_descriptor_distance_batch(_p(a), _p(b), len(a), _p(out)))
        return out if len(out) > 1 else int(out[0])

    def SearchForInitialization(self, keys1, desc1, keys2, desc2, bounds, vbPrevMatched, windowSize=10):
        k1, k2 = np.ascontiguousarray(keys1), np.ascontiguousarray(keys2)
        d1, d2 = _u8(desc1), _u8(desc2)
        pm = _f32(vbPrevMatched).copy(); b = _f32(bounds)
        m = np.zeros(max(len(k1), 1), np.int32)
        f = lib().pl_orb_search_for_initialization
        f.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, C.c_int, C.c_float, C.c_int]
        nm = check(f(_p(k1), _p(d1), len(k1), _p(k2), _p(d2), len(k2), _p(b), _p(pm), _p(m), int(windowSize),
                     self.mfNNratio, int(self.mbCheckOrientation)))
        return nm, m[:len(k1)], pm

    def SearchByProjectionLast(self, keys_cur, desc_cur, bounds, Tcw, K, scale_factors, last_valid, last_pos,
                               last_desc, last_octave, last_angle, th, preassigned=None):
        """SearchByProjection(CurrentFrame, LastFrame, th, bMono=True)."""
        kc, dc = np.ascontiguousarray(keys_cur), _u8(desc_cur)
        arr = [_f32(bounds), _f32(Tcw), _f32(K), _f32(scale_factors), _u8(last_valid), _f32(last_pos), _u8(last_desc),
               _i32(last_octave), _f32(last_angle)]
        pre = None if preassigned is None else _u8(preassigned)
        m = np.zeros(max(len(kc), 1), np.int32)
        f = lib().pl_orb_search_by_projection_last
        f.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_float, C.c_int, vp, vp]
        nm = check(f(_p(kc), _p(dc), len(kc), _p(arr[0]), _p(arr[1]), _p(arr[2]), _p(arr[3]), len(arr[3]), len(arr[4]),
                     _p(arr[4]), _p(arr[5]), _p(arr[6]), _p(arr[7]), _p(arr[8]), float(th),
                     int(self.mbCheckOrientation), _p(pre), _p(m)))
        return nm, m[:len(kc)]

    def SearchByProjectionPoints(self, keys, desc, bounds, scale_factors, in_view, proj, level, view_cos, mp_desc, th=3,
                                 preassigned=None):
        """SearchByProjection(F, vpMapPoints, th)."""
        k, d = np.ascontiguousarray(keys), _u8(desc)
        arr = [_f32(bounds), _f32(scale_factors), _u8(in_view), _f32(proj), _i32(level), _f32(view_cos), _u8(mp_desc)]
        pre = None if preassigned is None else _u8(preassigned)
        m = np.zeros(max(len(k), 1), np.int32)
        f = lib().pl_orb_search_by_projection_points
        f.argtypes = [vp, vp, C.c_int, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp]
        nm = check(f(_p(k), _p(d), len(k), _p(arr[0]), _p(arr[1]), len(arr[1]), len(arr[2]), _p(arr[2]), _p(arr[3]),
                     _p(arr[4]), _p(arr[5]), _p(arr[6]), float(th), self.mfNNratio, _p(pre), _p(m)))
        return nm, m[:len(k)]


class LSDmatcher:
    """Mirror of ORB_SLAM2::LSDmatcher (reference include/LSDmatcher.h:22-76) on flat arrays."""
    TH_HIGH, TH_LOW = 80, 50

    def __init__(self, nnratio=0.7, checkOri=True):
        self.mfNNratio, self.mbCheckOrientation = float(nnratio), bool(checkOri)

    DescriptorDistance = ORBmatcher.DescriptorDistance

    @staticmethod
    def knnMatch(d1, d2):
        """cv::BFMatcher(NORM_HAMMING).knnMatch(d1, d2, k=2) as used by FrameBFMatch."""
        d1, d2 = _u8(d1), _u8(d2)
        idx = np.zeros((max(len(d1), 1), 2), np.int32); dist = np.zeros((max(len(d1), 1), 2), np.int32)
        check(lib().pl_match_bf_knn2(_p(d1), len(d1), _p(d2), len(d2), _p(idx), _p(dist)))
        return idx[:len(d1)], dist[:len(d1)]

    def FrameBFMatch(self, ldesc1, ldesc2, TH=50.0):
        d1, d2 = _u8(ldesc1), _u8(ldesc2)
        m = np.zeros(max(len(d1), 1), np.int32)
        f = lib().pl_lsd_frame_bf_match
        f.argtypes = [vp, C.c_int, vp, C.c_int, C.c_float, C.c_float, vp]
        check(f(_p(d1), len(d1), _p(d2), len(d2), float(TH), self.mfNNratio, _p(m)))
        return m[:len(d1)]

    def SearchDouble(self, ldesc1, ldesc2):
        d1, d2 = _u8(ldesc1), _u8(ldesc2)
        m = np.zeros(max(len(d1), 1), np.int32)
        f = lib().pl_lsd_search_double
        f.argtypes = [vp, C.c_int, vp, C.c_int, C.c_float, vp]
        nm = check(f(_p(d1), len(d1), _p(d2), len(d2), self.mfNNratio, _p(m)))
        return nm, m[:len(d1)]


# ---------------------------------------------------------------------------------------------- pose-only LM
class Optimizer:
    """Mirror of ORB_SLAM2::Optimizer's pose-only entry points (reference include/Optimizer.h:56-65) on flat
    arrays: each call takes what the reference reads from the Frame and returns (n_inliers, Tcw, mvbOutlier,
    mvbLineOutlier, lm_iterations)."""

    @staticmethod
    def _run(mode, Tcw, K, pt_obs, pt_inv_sigma2, pt_Xw, line_func, line_Xw):
        Tcw, K = _f32(Tcw), _f32(K)
        po = _f32(pt_obs).reshape(-1, 2); pw = _f32(pt_inv_sigma2); px = _f32(pt_Xw).reshape(-1, 3)
        lf = np.ascontiguousarray(line_func, np.float64).reshape(-1, 3)
        lx = np.ascontiguousarray(line_Xw, np.float64).reshape(-1, 6)
        Tout = np.zeros((4, 4), np.float32)
        pout = np.zeros(max(len(po), 1), np.uint8); lout = np.zeros(max(len(lf), 1), np.uint8)
        its = C.c_int(0)
        f = lib().pl_pose_optimization
        f.argtypes = [C.c_int, vp, vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, vp, vp, vp, vp]
        n = check(f(mode, _p(Tcw), _p(K), len(po), _p(po), _p(pw), _p(px), len(lf), _p(lf), _p(lx), _p(Tout), _p(pout),
                    _p(lout), C.byref(its)))
        return n, Tout, pout[:len(po)].astype(bool), lout[:len(lf)].astype(bool), its.value

    @staticmethod
    def PoseOptimization(Tcw, K, pt_obs, pt_inv_sigma2, pt_Xw, line_func, line_Xw):
        return Optimizer._run(0, Tcw, K, pt_obs, pt_inv_sigma2, pt_Xw, line_func, line_Xw)

    @staticmethod
    def PoseOptimizationWithPoints(Tcw, K, pt_obs, pt_inv_sigma2, pt_Xw):
        return Optimizer._run(1, Tcw, K, pt_obs, pt_inv_sigma2, pt_Xw, np.zeros((0, 3)), np.zeros((0, 6)))

    @staticmethod
    def PoseOptimizationWithLines(Tcw, K, line_func, line_Xw):
        return Optimizer._run(2, Tcw, K, np.zeros((0, 2)), np.zeros(0), np.zeros((0, 3)), line_func, line_Xw)


# ---------------------------------------------------------------------------------------------- line features
KEYLINE_DTYPE = np.dtype([("angle", "<f4"), ("class_id", "<i4"), ("octave", "<i4"), ("ptx", "<f4"), ("pty", "<f4"),
                          ("response", "<f4"), ("size", "<f4"), ("startPointX", "<f4"), ("startPointY", "<f4"),
                          ("endPointX", "<f4"), ("endPointY", "<f4"), ("sPointInOctaveX", "<f4"),
                          ("sPointInOctaveY", "<f4"), ("ePointInOctaveX", "<f4"), ("ePointInOctaveY", "<f4"),
                          ("lineLength", "<f4"), ("numOfPixels", "<i4")])


class PLLineConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("nfeatures", C.c_int), ("min_line_length", C.c_double),
                ("max_batch", C.c_int), ("segment_cap", C.c_int), ("lsd_used_in_global", C.c_int)]


class LINEextractor:
    """Mirror of ORB_SLAM2::LINEextractor (reference include/LineExtractor.h:20-62).

    ctor (numOctaves, scale, nLSDFeature, min_line_length) as in the reference; `scale` reaches the detector as
    (int)scale == 1 and numOctaves is 1 in every shipped config, which is what is implemented.
    `__call__(image, mask)` == operator()(image, mask, keylines, descriptors, lineVec2d).
    """

    def __init__(self, numOctaves=1, scale=1.2, nLSDFeature=200, min_line_length=0.0, width=640, height=480, max_batch=1,
                 segment_cap=0):
        if numOctaves != 1 or int(scale) != 1:
            raise PLError("only numOctaves == 1 and int(scale) == 1 are supported (all reference configs)")
        self.cfg = PLLineConfig(width, height, nLSDFeature, float(min_line_length), max_batch, segment_cap, 0)
        self._h = vp()
        L = lib()
        L.pl_line_create.argtypes = [C.POINTER(PLLineConfig), C.POINTER(vp)]
        L.pl_line_destroy.argtypes = [vp]
        L.pl_line_capacity.argtypes = [vp]
        L.pl_line_extract.argtypes = [vp, vp, C.c_int, vp, vp, vp, vp, vp]
        L.pl_line_extract_batch.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_int, vp, vp, vp, vp, vp]
        L.pl_line_extract_batch_dev.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_int, vp, vp, vp, vp, vp, vp]
        L.pl_line_debug_segments.argtypes = [vp, C.c_int, vp, C.c_int]
        L.pl_line_debug_scaled.argtypes = [vp, C.c_int, vp, vp, vp]
        L.pl_line_debug_sobel.argtypes = [vp, C.c_int, vp, vp]
        L.pl_line_debug_order.argtypes = [vp, C.c_int, vp, C.c_int]
        check(L.pl_line_create(C.byref(self.cfg), C.byref(self._h)))
        self.capacity = check(L.pl_line_capacity(self._h))

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().pl_line_destroy(self._h)
            self._h = vp()

    def __call__(self, image, mask=None):
        image = np.ascontiguousarray(image, np.uint8)
        if image.size == 0:
            return np.zeros(0, KEYLINE_DTYPE), np.zeros((0, 32), np.uint8), np.zeros((0, 3))
        if mask is not None and (mask.shape != image.shape or mask.dtype != np.uint8):
            raise PLError("Mask error while detecting lines: please check its dimensions and that data type is CV_8UC1")
        kl = np.zeros(self.capacity, KEYLINE_DTYPE); desc = np.zeros((self.capacity, 32), np.uint8)
        lf = np.zeros((self.capacity, 3), np.float64); n = C.c_int(0)
        m = None if mask is None else np.ascontiguousarray(mask)
        check(lib().pl_line_extract(self._h, _p(image), image.strides[0], _p(m), _p(kl), _p(desc), _p(lf), C.byref(n)))
        return kl[:n.value].copy(), desc[:n.value].copy(), lf[:n.value].copy()

    def extract_batch(self, images, mask=None):
        images = np.ascontiguousarray(images, np.uint8)
        B = images.shape[0]
        kl = np.zeros((B, self.capacity), KEYLINE_DTYPE); desc = np.zeros((B, self.capacity, 32), np.uint8)
        lf = np.zeros((B, self.capacity, 3), np.float64); n = np.zeros(B, np.int32)
        m = None if mask is None else np.ascontiguousarray(mask)
        check(lib().pl_line_extract_batch(self._h, _p(images), images.strides[1], images.strides[0], B, _p(m), _p(kl),
                                          _p(desc), _p(lf), _p(n)))
        return kl, desc, lf, n

    def extract_batch_dev(self, img_ptr, stride, frame_stride, B, mask_ptr, kl_ptr, desc_ptr, lf_ptr, n_ptr, stream=None):
        check(lib().pl_line_extract_batch_dev(self._h, img_ptr, stride, frame_stride, B, mask_ptr, kl_ptr, desc_ptr,
                                              lf_ptr, n_ptr, stream))

    # parity taps
    def debug_segments(self, frame=0):
        n = check(lib().pl_line_debug_segments(self._h, frame, None, 0))
        out = np.zeros((max(n, 1), 4), np.float32)
        check(lib().pl_line_debug_segments(self._h, frame, _p(out), n))
        return out[:n]

    def debug_scaled(self, frame=0):
        sw, sh = C.c_int(), C.c_int()
        check(lib().pl_line_debug_scaled(self._h, frame, None, C.byref(sw), C.byref(sh)))
        out = np.zeros((sh.value, sw.value), np.uint8)
        check(lib().pl_line_debug_scaled(self._h, frame, _p(out), C.byref(sw), C.byref(sh)))
        return out

    def debug_sobel(self, frame=0):
        dx = np.zeros((self.cfg.height, self.cfg.width), np.int16); dy = np.zeros_like(dx)
        check(lib().pl_line_debug_sobel(self._h, frame, _p(dx), _p(dy)))
        return dx, dy

    def debug_order(self, frame=0):
        n = check(lib().pl_line_debug_order(self._h, frame, None, 0))
        out = np.zeros(max(n, 1), np.uint32)
        check(lib().pl_line_debug_order(self._h, frame, _p(out), n))
        return out[:n]


# ---------------------------------------------------------------------------------------------- front-end pipeline
class PLFrontendConfig(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("max_batch", C.c_int), ("orb_nfeatures", C.c_int),
                ("orb_scale_factor", C.c_float), ("orb_nlevels", C.c_int), ("orb_ini_th", C.c_int), ("orb_min_th", C.c_int),
                ("line_nfeatures", C.c_int), ("line_min_length", C.c_double), ("lm_cap_points", C.c_int),
                ("lm_cap_lines", C.c_int)]


class Frontend:
    """Batch front-end: ORB + LSD/LBD extraction, frame-to-frame matching, 2 x PoseOptimization per frame."""

    def __init__(self, width=640, height=480, max_batch=8, orb=(1000, 1.2, 8, 20, 7), lines=(200, 0.0), lm_caps=(512, 128)):
        self.cfg = PLFrontendConfig(width, height, max_batch, orb[0], orb[1], orb[2], orb[3], orb[4], lines[0], lines[1],
                                    lm_caps[0], lm_caps[1])
        self._h = vp()
        L = lib()
        L.pl_frontend_create.argtypes = [C.POINTER(PLFrontendConfig), C.POINTER(vp)]
        L.pl_frontend_destroy.argtypes = [vp]
        L.pl_frontend_capacities.argtypes = [vp, vp, vp]
        L.pl_frontend_set_pose_problems.argtypes = [vp, C.c_int] + [vp] * 9
        L.pl_frontend_run_dev.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_int, vp]
        L.pl_frontend_run.argtypes = [vp, vp, C.c_int, C.c_size_t, C.c_int] + [vp] * 13
        L.pl_frontend_io_bytes.argtypes = [vp, vp, vp]
        L.pl_frontend_fetch.argtypes = [vp, C.c_int] + [vp] * 12
        L.pl_frontend_set_timing.argtypes = [vp, C.c_int]
        L.pl_frontend_grow_ms.argtypes = [vp, vp]
        L.pl_frontend_grow_bytes_per_frame.argtypes = [vp]
        L.pl_frontend_grow_bytes_per_frame.restype = C.c_longlong
        L.pl_frontend_copy_poses_dev.argtypes = [vp, C.c_int, vp, vp]
        check(L.pl_frontend_create(C.byref(self.cfg), C.byref(self._h)))
        ck, cl = C.c_int(), C.c_int()
        check(L.pl_frontend_capacities(self._h, C.byref(ck), C.byref(cl)))
        self.capK, self.capL = ck.value, cl.value
        self.cap_points, self.cap_lines = lm_caps

    def __del__(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().pl_frontend_destroy(self._h)
            self._h = vp()

    def set_pose_problems(self, problems):
        """problems: list of dicts from synth.synth_pose_problem (one per frame of the batch)."""
        B, cp, cl = len(problems), self.cap_points, self.cap_lines
        T0 = np.zeros((B, 16), np.float32); K = np.zeros((B, 4), np.float32)
        npt = np.zeros(B, np.int32); nln = np.zeros(B, np.int32)
        obs = np.zeros((B, cp, 2), np.float32); w = np.zeros((B, cp), np.float32); X = np.zeros((B, cp, 3), np.float32)
        lf = np.zeros((B, cl, 3), np.float64); lX = np.zeros((B, cl, 6), np.float64)
        for b, p in enumerate(problems):
            n, m = len(p["pt_obs"]), len(p["line_func"])
            assert n <= cp and m <= cl
            T0[b] = p["Tcw0"].ravel(); K[b] = p["K"]; npt[b] = n; nln[b] = m
            obs[b, :n] = p["pt_obs"]; w[b, :n] = p["pt_inv_sigma2"]; X[b, :n] = p["pt_Xw"]
            lf[b, :m] = p["line_func"]; lX[b, :m] = p["line_Xw"]
        self._problems = (T0, K, npt, obs, w, X, nln, lf, lX)
        check(lib().pl_frontend_set_pose_problems(self._h, B, _p(T0), _p(K), _p(npt), _p(obs), _p(w), _p(X), _p(nln), _p(lf), _p(lX)))

    def set_wrap(self, on=True):
        """Frame 0 is matched against the last frame of the SAME batch (closed loop) instead of the previous step's last frame."""
        lib().pl_frontend_set_wrap.argtypes = [vp, C.c_int]
        check(lib().pl_frontend_set_wrap(self._h, int(on)))

    def set_tracking(self, on=True):
        """Add the steady-state projection searches (Tracking.cc:1345-1357,1799,1855) to the step; needs pose problems (Tcw0, K)."""
        lib().pl_frontend_set_tracking.argtypes = [vp, C.c_int]
        check(lib().pl_frontend_set_tracking(self._h, int(on)))

    def fetch_tracking(self, B, which=0):
        """dict(pt_match [B][capK], n_pt, line_match [B][capL], n_line, map_pos [B][capK][3], pt_in_view, line_in_view)."""
        o = dict(pt_match=np.zeros((B, self.capK), np.int32), n_pt=np.zeros(B, np.int32), line_match=np.zeros((B, self.capL), np.int32),
                 n_line=np.zeros(B, np.int32), map_pos=np.zeros((B, self.capK, 3), np.float32), pt_in_view=np.zeros((B, self.capK), np.uint8),
                 line_in_view=np.zeros((B, self.capL), np.uint8))
        lib().pl_frontend_fetch_tracking.argtypes = [vp, C.c_int, C.c_int] + [vp] * 7
        check(lib().pl_frontend_fetch_tracking(self._h, B, which, _p(o["pt_match"]), _p(o["n_pt"]), _p(o["line_match"]), _p(o["n_line"]),
                                               _p(o["map_pos"]), _p(o["pt_in_view"]), _p(o["line_in_view"])))
        return o

    def dump(self, B, path):
        """pl_frontend_dump: the last step's results in the replay container trajectory.load_frontend reads."""
        lib().pl_frontend_dump.argtypes = [vp, C.c_int, C.c_char_p]
        check(lib().pl_frontend_dump(self._h, B, str(path).encode()))

    def pack_pose_problems(self, problems, pinned=True):
        """Pack the batch's pose problems into (pinned) host arrays for upload_pose_problems()."""
        import torch
        B, cp, cl = len(problems), self.cap_points, self.cap_lines
        shapes = [((B, 16), np.float32), ((B, 4), np.float32), ((B,), np.int32), ((B, cp, 2), np.float32), ((B, cp), np.float32),
                  ((B, cp, 3), np.float32), ((B,), np.int32), ((B, cl, 3), np.float64), ((B, cl, 6), np.float64)]
        arrs, keep = [], []
        for shp, dt in shapes:
            nbytes = int(np.prod(shp)) * np.dtype(dt).itemsize
            t = torch.zeros(max(nbytes, 1), dtype=torch.uint8, pin_memory=pinned)
            keep.append(t)
            arrs.append(t.numpy()[:nbytes].view(dt).reshape(shp))
        T0, K, npt, obs, w, X, nln, lf, lX = arrs
        for b, p in enumerate(problems):
            n, m = len(p["pt_obs"]), len(p["line_func"])
            assert n <= cp and m <= cl
            T0[b] = p["Tcw0"].ravel(); K[b] = p["K"]; npt[b] = n; nln[b] = m
            obs[b, :n] = p["pt_obs"]; w[b, :n] = p["pt_inv_sigma2"]; X[b, :n] = p["pt_Xw"]
            lf[b, :m] = p["line_func"]; lX[b, :m] = p["line_Xw"]
        self._packed = (arrs, keep, B)
        return self._packed

    def upload_pose_problems(self, stream=None):
        """Enqueue the upload of the packed problems (no synchronisation); returns the bytes enqueued."""
        arrs, _, B = self._packed
        f = lib().pl_frontend_set_pose_problems_async
        f.argtypes = [vp, C.c_int] + [vp] * 9 + [vp]
        f.restype = C.c_longlong
        return check(f(self._h, B, *[_p(a) for a in arrs], stream))

    def set_camera(self, K, distCoef):
        """mK / mDistCoef of the sequence: with k1 != 0 the step undistorts frames (for lines) and keypoints (for matching)."""
        K = _f32(K); D = _f32(distCoef)
        assert K.shape == (4,) and D.shape == (5,)
        check(lib().pl_frontend_set_camera(self._h, _p(K), _p(D)))

    def fetch_keys_un(self, B):
        out = np.zeros((B, self.capK), KP_DTYPE)
        check(lib().pl_frontend_fetch_keys_un(self._h, C.c_int(B), _p(out)))
        return out

    def alloc_outputs(self, B, pinned=False):
        shapes = dict(kps=((B, self.capK), KP_DTYPE), desc=((B, self.capK, 32), np.uint8), n=((B,), np.int32),
                      keylines=((B, self.capL), KEYLINE_DTYPE), ldesc=((B, self.capL, 32), np.uint8),
                      linefunc=((B, self.capL, 3), np.float64), nl=((B,), np.int32), pt_matches=((B, self.capK), np.int32),
                      n_pt_matches=((B,), np.int32), line_matches=((B, self.capL), np.int32), n_line_matches=((B,), np.int32),
                      poses=((2, B, 16), np.float32), inliers=((2, B), np.int32))
        out = {}
        for k, (shp, dt) in shapes.items():
            if pinned:
                import torch
                nbytes = int(np.prod(shp)) * np.dtype(dt).itemsize
                t = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=True)
                out["_pin_" + k] = t
                out[k] = t.numpy()[:nbytes].view(dt).reshape(shp)
            else:
                out[k] = np.zeros(shp, dt)
        return out

    ORDER = ["kps", "desc", "n", "keylines", "ldesc", "linefunc", "nl", "pt_matches", "n_pt_matches", "line_matches",
             "n_line_matches", "poses", "inliers"]

    def run(self, images, out=None):
        """End-to-end on host buffers (images: uint8 [B][H][W])."""
        B = images.shape[0]
        out = out or self.alloc_outputs(B)
        check(lib().pl_frontend_run(self._h, _p(images), images.strides[1], images.strides[0], B,
                                    *[_p(out[k]) for k in self.ORDER]))
        return out

    def submit(self, images, out):
        """Streaming form of run(): enqueue one step (pinned host buffers); results are valid after wait()."""
        B = images.shape[0]
        check(lib().pl_frontend_submit(self._h, _p(images), C.c_int(images.strides[1]), C.c_size_t(images.strides[0]), C.c_int(B),
                                       *[_p(out[k]) for k in self.ORDER]))

    def wait(self, keep_in_flight=0):
        check(lib().pl_frontend_wait(self._h, C.c_int(keep_in_flight)))

    def run_dev(self, img_ptr, stride, frame_stride, B, stream=None):
        check(lib().pl_frontend_run_dev(self._h, img_ptr, stride, frame_stride, B, stream))

    def fetch(self, B):
        out = self.alloc_outputs(B)
        order = [k for k in self.ORDER if k != "linefunc"]
        check(lib().pl_frontend_fetch(self._h, B, *[_p(out[k]) for k in order]))
        return out

    def io_bytes(self):
        a, b = C.c_longlong(), C.c_longlong()
        check(lib().pl_frontend_io_bytes(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def set_timing(self, on=True):
        check(lib().pl_frontend_set_timing(self._h, int(on)))

    def grow_ms(self):
        ms = C.c_float()
        check(lib().pl_frontend_grow_ms(self._h, C.byref(ms)))
        return ms.value

    def grow_bytes_per_frame(self):
        return int(lib().pl_frontend_grow_bytes_per_frame(self._h))

    def copy_poses_dev(self, B, dst_ptr, stream=None):
        check(lib().pl_frontend_copy_poses_dev(self._h, B, dst_ptr, stream))


# ---------------------------------------------------------------------------------------------- local BA
class PLBAProblem(C.Structure):
    _fields_ = [("n_kf", C.c_int), ("kf_Tcw", vp), ("kf_fixed", vp), ("kf_K", vp), ("K_end", C.c_float * 4),
                ("n_pt", C.c_int), ("pt_Xw", vp), ("n_ln", C.c_int), ("ln_Xw", vp),
                ("n_pe", C.c_int), ("pe_kf", vp), ("pe_pt", vp), ("pe_obs", vp), ("pe_inv_sigma2", vp),
                ("n_le", C.c_int), ("le_kf", vp), ("le_ln", vp), ("le_func", vp)]


def LocalBundleAdjustmentWithLine(p, stop_flag_dev=None):
    """Optimizer::LocalBundleAdjustmentWithLine on a flattened local window (dict as made by synth.synth_ba_problem).
    Returns dict(kf_Tcw, pt_Xw, ln_Xw, pe_erase, le_erase, le_erase_kf, its)."""
    a = {k: np.ascontiguousarray(v) for k, v in p.items() if isinstance(v, np.ndarray)}
    n_kf, n_pt, n_ln, n_pe, n_le = len(a["kf_fixed"]), len(a["pt_Xw"]), len(a["ln_Xw"]), len(a["pe_kf"]), len(a["le_kf"])
    P = PLBAProblem(n_kf, _p(a["kf_Tcw"]), _p(a["kf_fixed"]), _p(a["kf_K"]), (C.c_float * 4)(*[float(v) for v in a["K_end"]]),
                    n_pt, _p(a["pt_Xw"]), n_ln, _p(a["ln_Xw"]), n_pe, _p(a["pe_kf"]), _p(a["pe_pt"]), _p(a["pe_obs"]),
                    _p(a["pe_inv_sigma2"]), n_le, _p(a["le_kf"]), _p(a["le_ln"]), _p(a["le_func"]))
    out = dict(kf_Tcw=np.zeros((n_kf, 16), np.float32), pt_Xw=np.zeros((max(n_pt, 1), 3), np.float32),
               ln_Xw=np.zeros((max(n_ln, 1), 6), np.float64), pe_erase=np.zeros(max(n_pe, 1), np.uint8),
               le_erase=np.zeros(max(n_le, 1), np.uint8), le_erase_kf=np.zeros(max(n_le, 1), np.int32))
    its = C.c_int(0)
    f = lib().pl_local_ba
    f.argtypes = [C.POINTER(PLBAProblem), vp, vp, vp, vp, vp, vp, vp, vp]
    check(f(C.byref(P), stop_flag_dev, _p(out["kf_Tcw"]), _p(out["pt_Xw"]), _p(out["ln_Xw"]), _p(out["pe_erase"]),
            _p(out["le_erase"]), _p(out["le_erase_kf"]), C.byref(its)))
    out["its"] = its.value
    for k, n in (("pt_Xw", n_pt), ("ln_Xw", n_ln), ("pe_erase", n_pe), ("le_erase", n_le), ("le_erase_kf", n_le)):
        out[k] = out[k][:n]
    return out


Optimizer.LocalBundleAdjustmentWithLine = staticmethod(LocalBundleAdjustmentWithLine)


def GlobalBundleAdjustemnt(p, nIterations=5, bRobust=True, stop_flag=None):
    """Optimizer::GlobalBundleAdjustemnt / BundleAdjustment with lines (Optimizer.cc:41-58,275-638; the reference's spelling) on
    a flattened map (dict as made by synth.synth_ba_problem; K_end is not read).  stop_flag: None or an int32 numpy scalar
    array the caller may set while the call runs.  Returns dict(kf_Tcw, pt_Xw, ln_Xw, its, solve_ms)."""
    a = {k: np.ascontiguousarray(v) for k, v in p.items() if isinstance(v, np.ndarray)}
    n_kf, n_pt, n_ln, n_pe, n_le = len(a["kf_fixed"]), len(a["pt_Xw"]), len(a["ln_Xw"]), len(a["pe_kf"]), len(a["le_kf"])
    P = PLBAProblem(n_kf, _p(a["kf_Tcw"]), _p(a["kf_fixed"]), _p(a["kf_K"]), (C.c_float * 4)(0, 0, 0, 0),
                    n_pt, _p(a["pt_Xw"]), n_ln, _p(a["ln_Xw"]), n_pe, _p(a["pe_kf"]), _p(a["pe_pt"]), _p(a["pe_obs"]),
                    _p(a["pe_inv_sigma2"]), n_le, _p(a["le_kf"]), _p(a["le_ln"]), _p(a["le_func"]))
    out = dict(kf_Tcw=np.zeros((n_kf, 16), np.float32), pt_Xw=np.zeros((max(n_pt, 1), 3), np.float32),
               ln_Xw=np.zeros((max(n_ln, 1), 6), np.float64))
    its = C.c_int(0); ms = C.c_float(0)
    f = lib().pl_global_ba
    f.argtypes = [C.POINTER(PLBAProblem), C.c_int, C.c_int, vp, vp, vp, vp, vp, vp]
    sf = None if stop_flag is None else _p(stop_flag)
    check(f(C.byref(P), int(nIterations), int(bool(bRobust)), sf, _p(out["kf_Tcw"]), _p(out["pt_Xw"]), _p(out["ln_Xw"]),
            C.byref(its), C.byref(ms)))
    out["its"] = its.value; out["solve_ms"] = ms.value
    out["pt_Xw"] = out["pt_Xw"][:n_pt]; out["ln_Xw"] = out["ln_Xw"][:n_ln]
    return out


Optimizer.GlobalBundleAdjustemnt = staticmethod(GlobalBundleAdjustemnt)


# ---------------------------------------------------------------------------------------------- line matching by projection
def frame_assign_grid_lines(keylines_un, bounds):
    """Frame::AssignFeaturesToGridForLine (reference src/Frame.cc:296-320) -> CSR (cell_start[3073], cell_items)."""
    kl = np.ascontiguousarray(keylines_un); b = _f32(bounds)
    cap = max(len(kl), 1) * 120
    start = np.zeros(64 * 48 + 1, np.int32); items = np.zeros(cap, np.int32)
    f = lib().pl_frame_assign_grid_lines
    f.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int]
    n = check(f(_p(kl), len(kl), _p(b), _p(start), _p(items), cap))
    return start, items[:n]


def _lsd_search_last(self, keylines_cur, linefunc_cur, desc_cur, bounds, last_valid, last_proj, last_desc, last_length, th, preassigned=None):
    """LSDmatcher::SearchByProjection(CurrentFrame, LastFrame, th)."""
    kl = np.ascontiguousarray(keylines_cur)
    a = [np.ascontiguousarray(linefunc_cur, np.float64), _u8(desc_cur), _f32(bounds), _u8(last_valid), _f32(last_proj), _u8(last_desc), _f32(last_length)]
    pre = None if preassigned is None else _u8(preassigned)
    m = np.zeros(max(len(kl), 1), np.int32)
    f = lib().pl_lsd_search_by_projection_last
    f.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, C.c_float, vp, vp]
    nm = check(f(_p(kl), _p(a[0]), _p(a[1]), len(kl), _p(a[2]), len(a[3]), _p(a[3]), _p(a[4]), _p(a[5]), _p(a[6]), float(th), _p(pre), _p(m)))
    return nm, m[:len(kl)]


def _lsd_search_lines(self, keylines, linefunc, desc, bounds, in_view, proj, view_cos, ml_desc, th=3, preassigned=None):
    """LSDmatcher::SearchByProjection(F, vpMapLines, th)."""
    kl = np.ascontiguousarray(keylines)
    a = [np.ascontiguousarray(linefunc, np.float64), _u8(desc), _f32(bounds), _u8(in_view), _f32(proj), _f32(view_cos), _u8(ml_desc)]
    pre = None if preassigned is None else _u8(preassigned)
    m = np.zeros(max(len(kl), 1), np.int32)
    f = lib().pl_lsd_search_by_projection_lines
    f.argtypes = [vp, vp, vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, C.c_float, C.c_float, vp, vp]
    nm = check(f(_p(kl), _p(a[0]), _p(a[1]), len(kl), _p(a[2]), len(a[3]), _p(a[3]), _p(a[4]), _p(a[5]), _p(a[6]), float(th),
                 self.mfNNratio, _p(pre), _p(m)))
    return nm, m[:len(kl)]


LSDmatcher.SearchByProjectionLast = _lsd_search_last
LSDmatcher.SearchByProjectionLines = _lsd_search_lines


# ----------------------------------------------------------------- Frame glue (reference src/Frame.cc)
class Undistorter:
    """The undistortion the mono Frame constructor does for every frame (reference src/Frame.cc:220-222:
    initUndistortRectifyMap + remap) and Frame::UndistortKeyPoints (:915-945); the map is built once per camera."""

    def __init__(self, K, distCoef, width, height):
        self.K = _f32(K); self.D = _f32(distCoef); self.w = int(width); self.h = int(height)
        assert self.K.shape == (4,) and self.D.shape == (5,)
        self.handle = vp()
        check(lib().pl_undistort_create(_p(self.K), _p(self.D), C.c_int(self.w), C.c_int(self.h), C.byref(self.handle)))

    def __del__(self):
        if getattr(self, "handle", None):
            lib().pl_undistort_destroy(self.handle); self.handle = None

    def remap(self, image):
        img = _u8(image)
        if img.shape != (self.h, self.w):
            raise PLError(f"image {img.shape} does not match the undistorter ({self.h}, {self.w})")
        out = np.empty_like(img)
        check(lib().pl_undistort_remap(self.handle, _p(img), C.c_int(self.w), _p(out), C.c_int(self.w)))
        return out

    def remap_batch_dev(self, src_ptr, sstride, sframe, B, dst_ptr, dstride, dframe, stream=None):
        check(lib().pl_undistort_remap_batch_dev(self.handle, vp(src_ptr), C.c_int(sstride), C.c_size_t(sframe), C.c_int(B),
                                                 vp(dst_ptr), C.c_int(dstride), C.c_size_t(dframe), vp(stream or 0)))

    def UndistortKeyPoints(self, keys):
        keys = np.ascontiguousarray(keys, KP_DTYPE); out = np.empty_like(keys)
        check(lib().pl_undistort_keypoints(self.handle, _p(keys), C.c_int(len(keys)), _p(out)))
        return out

    def undistort_keypoints_dev(self, kps_ptr, n_ptr, cap, B, out_ptr, stream=None):
        check(lib().pl_undistort_keypoints_dev(self.handle, vp(kps_ptr), vp(n_ptr), C.c_int(cap), C.c_int(B), vp(out_ptr),
                                               vp(stream or 0)))

    def ComputeImageBounds(self):
        return ComputeImageBounds(self.K, self.D, self.w, self.h)


def ComputeImageBounds(K, distCoef, width, height):
    """Frame::ComputeImageBounds (reference src/Frame.cc:947-985) -> [mnMinX, mnMinY, mnMaxX, mnMaxY]."""
    K = _f32(K); D = _f32(distCoef); b = np.empty(4, np.float32)
    check(lib().pl_frame_image_bounds(_p(K), _p(D), C.c_int(width), C.c_int(height), _p(b)))
    return b


def isInFrustum(Tcw, Ow, K, bounds, log_scale_factor, n_scale_levels, viewingCosLimit, pos, normal, min_dist, max_dist):
    """Frame::isInFrustum(MapPoint*, viewingCosLimit) over n map points (reference src/Frame.cc:560-620)."""
    n = len(pos)
    Tcw = _f32(Tcw); Ow = _f32(Ow); K = _f32(K); bounds = _f32(bounds)
    pos = _f32(pos); normal = _f32(normal); min_dist = _f32(min_dist); max_dist = _f32(max_dist)
    inview = np.zeros(n, np.uint8); proj = np.zeros((n, 2), np.float32); level = np.zeros(n, np.int32); vc = np.zeros(n, np.float32)
    check(lib().pl_frame_is_in_frustum_points(_p(Tcw), _p(Ow), _p(K), _p(bounds), C.c_float(log_scale_factor),
                                              C.c_int(n_scale_levels), C.c_float(viewingCosLimit), C.c_int(n), _p(pos), _p(normal),
                                              _p(min_dist), _p(max_dist), _p(inview), _p(proj), _p(level), _p(vc)))
    return inview, proj, level, vc


def isInFrustumLines(Tcw, Ow, K, bounds, log_scale_factor, viewingCosLimit, pos, normal, min_dist, max_dist):
    """Frame::isInFrustum(MapLine*, viewingCosLimit) over n map lines (reference src/Frame.cc:622-702)."""
    n = len(pos)
    Tcw = _f32(Tcw); Ow = _f32(Ow); K = _f32(K); bounds = _f32(bounds)
    pos = np.ascontiguousarray(pos, np.float64); normal = np.ascontiguousarray(normal, np.float64)
    min_dist = _f32(min_dist); max_dist = _f32(max_dist)
    inview = np.zeros(n, np.uint8); proj = np.zeros((n, 4), np.float32); level = np.zeros(n, np.int32); vc = np.zeros(n, np.float32)
    check(lib().pl_frame_is_in_frustum_lines(_p(Tcw), _p(Ow), _p(K), _p(bounds), C.c_float(log_scale_factor),
                                             C.c_float(viewingCosLimit), C.c_int(n), _p(pos), _p(normal), _p(min_dist),
                                             _p(max_dist), _p(inview), _p(proj), _p(level), _p(vc)))
    return inview, proj, level, vc


# ----------------------------------------------------------------- LocalMapping matchers (reference src/ORBmatcher.cc:720-1065)
def _fv_csr(fv):
    """DBoW2::FeatureVector (dict node -> feature indices, std::map order = ascending node id) as CSR arrays."""
    nodes = np.array(sorted(fv), np.uint32)
    start = np.zeros(len(nodes) + 1, np.int32); items = []
    for i, nd in enumerate(nodes):
        items += list(fv[int(nd)]); start[i + 1] = len(items)
    return nodes, start, np.array(items, np.int32).reshape(-1)


def _search_for_triangulation(self, keys1_un, desc1, has_mp1, keys2_un, desc2, has_mp2, fv1, fv2, F12, Cw1, R2w, t2w, K2,
                              scale_factors2, level_sigma2_2):
    """ORBmatcher::SearchForTriangulation(pKF1, pKF2, F12, vMatchedPairs, false) -> (nmatches, matches12[n1])."""
    k1 = np.ascontiguousarray(keys1_un, KP_DTYPE); k2 = np.ascontiguousarray(keys2_un, KP_DTYPE)
    d1 = _u8(desc1); d2 = _u8(desc2); m1 = _u8(has_mp1); m2 = _u8(has_mp2)
    n1a, s1, i1 = _fv_csr(fv1); n2a, s2, i2 = _fv_csr(fv2)
    F = _f32(F12); Cw = _f32(Cw1); R = _f32(R2w); t = _f32(t2w); K = _f32(K2); sf = _f32(scale_factors2); sg = _f32(level_sigma2_2)
    out = np.full(len(k1), -1, np.int32)
    nm = check(lib().pl_orb_search_for_triangulation(
        _p(k1), _p(d1), _p(m1), C.c_int(len(k1)), _p(k2), _p(d2), _p(m2), C.c_int(len(k2)), _p(n1a), _p(s1), _p(i1), C.c_int(len(n1a)),
        _p(n2a), _p(s2), _p(i2), C.c_int(len(n2a)), _p(F), _p(Cw), _p(R), _p(t), _p(K), _p(sf), _p(sg), C.c_int(len(sf)),
        C.c_int(int(self.mbCheckOrientation)), _p(out)))
    return nm, out


def _fuse_search(self, keys_un, desc, bounds, Tcw, Ow, K, scale_factors, inv_level_sigma2, log_scale_factor, skip, pos, normal,
                 min_dist, max_dist, mp_desc, th=3.0):
    """Search half of ORBmatcher::Fuse(pKF, vpMapPoints, th) -> (best_idx[n_mp], best_dist[n_mp])."""
    keys = np.ascontiguousarray(keys_un, KP_DTYPE); desc = _u8(desc)
    b = _f32(bounds); T = _f32(Tcw); O = _f32(Ow); Kc = _f32(K); sf = _f32(scale_factors); iv = _f32(inv_level_sigma2)
    n_mp = len(pos)
    sk = None if skip is None else _u8(skip)
    pos = _f32(pos); normal = _f32(normal); mn = _f32(min_dist); mx = _f32(max_dist); md = _u8(mp_desc)
    bi = np.zeros(n_mp, np.int32); bd = np.zeros(n_mp, np.int32)
    check(lib().pl_orb_fuse_search(_p(keys), _p(desc), C.c_int(len(keys)), _p(b), _p(T), _p(O), _p(Kc), _p(sf), _p(iv),
                                   C.c_int(len(sf)), C.c_float(log_scale_factor), C.c_int(n_mp), _p(sk), _p(pos), _p(normal), _p(mn),
                                   _p(mx), _p(md), C.c_float(th), _p(bi), _p(bd)))
    return bi, bd


ORBmatcher.SearchForTriangulation = _search_for_triangulation
ORBmatcher.FuseSearch = _fuse_search


def _lsd_search_for_triangulation(self, ldesc1, has_ml1, ldesc2, has_ml2, isDouble=True, th=None):
    """LSDmatcher::SearchForTriangulation(pKF1, pKF2, vMatchedPairs, isDouble) (reference src/LSDmatcher.cpp:727-776)
    -> (nmatches, vMatchedPairs[NL1]); th defaults to TH_HIGH (the 4-argument overload), th = TH_LOW + isDouble is the
    pair<> overload (:672-725)."""
    th = float(self.TH_HIGH if th is None else th)
    d1 = _u8(ldesc1).reshape(-1, 32); d2 = _u8(ldesc2).reshape(-1, 32); m1 = _u8(has_ml1); m2 = _u8(has_ml2)
    out = np.full(len(d1), -1, np.int32)
    nm = check(lib().pl_lsd_search_for_triangulation(_p(d1), _p(m1), C.c_int(len(d1)), _p(d2), _p(m2), C.c_int(len(d2)),
                                                     C.c_float(th), C.c_float(self.mfNNratio), C.c_int(int(isDouble)), _p(out)))
    return nm, out


LSDmatcher.SearchForTriangulation = _lsd_search_for_triangulation


def _search_by_bow(self, keysKF_un, descKF, has_mp_kf, keysF, descF, fvKF, fvF):
    """ORBmatcher::SearchByBoW(pKF, F, vpMapPointMatches) (reference src/ORBmatcher.cc:187-327) -> (nmatches, matchesF[F.N]):
    matchesF[j] = index of the keyframe feature whose MapPoint frame feature j receives."""
    kK = np.ascontiguousarray(keysKF_un, KP_DTYPE); kF = np.ascontiguousarray(keysF, KP_DTYPE)
    dK = _u8(descKF); dF = _u8(descF); mp = _u8(has_mp_kf)
    n1a, s1, i1 = _fv_csr(fvKF); n2a, s2, i2 = _fv_csr(fvF)
    out = np.full(len(kF), -1, np.int32)
    nm = check(lib().pl_orb_search_by_bow(_p(kK), _p(dK), _p(mp), C.c_int(len(kK)), _p(kF), _p(dF), C.c_int(len(kF)), _p(n1a), _p(s1),
                                          _p(i1), C.c_int(len(n1a)), _p(n2a), _p(s2), _p(i2), C.c_int(len(n2a)),
                                          C.c_float(self.mfNNratio), C.c_int(int(self.mbCheckOrientation)), _p(out)))
    return nm, out


ORBmatcher.SearchByBoW = _search_by_bow


def _search_by_projection_keyframe(self, keys_cur_un, desc_cur, bounds, Tcw, Ow, K, scale_factors, log_scale_factor, kf_valid, pos,
                                   mp_desc, min_dist, max_dist, kf_angle, th, ORBdist, preassigned=None):
    """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (reference src/ORBmatcher.cc:1587-1716)
    -> (nmatches, cur_match[F.N]) with cur_match[i2] = index into the keyframe's map-point list."""
    kc = np.ascontiguousarray(keys_cur_un, KP_DTYPE); dc = _u8(desc_cur)
    b = _f32(bounds); T = _f32(Tcw); O = _f32(Ow); Kc = _f32(K); sf = _f32(scale_factors)
    v = _u8(kf_valid); pos = _f32(pos); md = _u8(mp_desc); mn = _f32(min_dist); mx = _f32(max_dist); ang = _f32(kf_angle)
    pre = None if preassigned is None else _u8(preassigned)
    out = np.full(len(kc), -1, np.int32)
    nm = check(lib().pl_orb_search_by_projection_keyframe(
        _p(kc), _p(dc), C.c_int(len(kc)), _p(b), _p(T), _p(O), _p(Kc), _p(sf), C.c_int(len(sf)), C.c_float(log_scale_factor),
        C.c_int(len(v)), _p(v), _p(pos), _p(md), _p(mn), _p(mx), _p(ang), C.c_float(th), C.c_int(int(ORBdist)),
        C.c_int(int(self.mbCheckOrientation)), _p(pre), _p(out)))
    return nm, out


ORBmatcher.SearchByProjectionKeyFrame = _search_by_projection_keyframe


def _search_by_bow_keyframes(self, keys1_un, desc1, has_mp1, keys2_un, desc2, has_mp2, fv1, fv2):
    """ORBmatcher::SearchByBoW(pKF1, pKF2, vpMatches12) (reference src/ORBmatcher.cc:574-709) -> (nmatches, matches12[n1])."""
    k1 = np.ascontiguousarray(keys1_un, KP_DTYPE); k2 = np.ascontiguousarray(keys2_un, KP_DTYPE)
    d1 = _u8(desc1); d2 = _u8(desc2); m1 = _u8(has_mp1); m2 = _u8(has_mp2)
    n1a, s1, i1 = _fv_csr(fv1); n2a, s2, i2 = _fv_csr(fv2)
    out = np.full(len(k1), -1, np.int32)
    nm = check(lib().pl_orb_search_by_bow_keyframes(_p(k1), _p(d1), _p(m1), C.c_int(len(k1)), _p(k2), _p(d2), _p(m2), C.c_int(len(k2)),
                                                    _p(n1a), _p(s1), _p(i1), C.c_int(len(n1a)), _p(n2a), _p(s2), _p(i2), C.c_int(len(n2a)),
                                                    C.c_float(self.mfNNratio), C.c_int(int(self.mbCheckOrientation)), _p(out)))
    return nm, out


ORBmatcher.SearchByBoWKeyFrames = _search_by_bow_keyframes


# ---------------------------------------------------------------------------------------------- LocalMapping: map-point descriptor, line fusion
def ComputeDistinctiveDescriptors(desc, offsets, return_desc=False):
    """MapPoint::ComputeDistinctiveDescriptors (reference src/MapPoint.cc:249-314) for a batch of map points: descriptors of
    point m = rows offsets[m]..offsets[m+1] of desc.  Returns best index inside each point's list (-1 for an empty list)."""
    desc = _u8(desc).reshape(-1, 32); off = np.ascontiguousarray(offsets, np.int32)
    n = len(off) - 1
    best = np.zeros(max(n, 1), np.int32); out = np.zeros((max(n, 1), 32), np.uint8) if return_desc else None
    f = lib().pl_mappoint_distinctive_descriptors
    f.argtypes = [vp, vp, C.c_int, vp, vp]
    check(f(_p(desc), _p(off), n, _p(best), _p(out)))
    return (best[:n], out[:n]) if return_desc else best[:n]


def _lsd_fuse_search(self, keylines, kf_point_desc, bounds, Tcw, Ow, K, scale_line, log_scale_factor_line, skip, pos, normal,
                     min_dist, max_dist, ml_desc, th=3.0):
    """Search half of LSDmatcher::Fuse(pKF, vpMapLines, th) (reference src/LSDmatcher.cpp:860-1011) -> (best_idx, best_dist,
    stop_at); quirks in include/plslam_b200.h (pl_lsd_fuse_search)."""
    kl = np.ascontiguousarray(keylines); pd = _u8(kf_point_desc).reshape(-1, 32)
    b = _f32(bounds); T = _f32(Tcw); O = _f32(Ow); Kc = _f32(K)
    n = len(pos)
    sk = _u8(skip); P = np.ascontiguousarray(pos, np.float64); Nn = np.ascontiguousarray(normal, np.float64)
    mn = _f32(min_dist); mx = _f32(max_dist); md = _u8(ml_desc).reshape(-1, 32)
    bi = np.zeros(max(n, 1), np.int32); bd = np.zeros(max(n, 1), np.int32); stop = C.c_int(n)
    f = lib().pl_lsd_fuse_search
    f.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp, C.c_float, C.c_float, C.c_int, vp, vp, vp, vp, vp, vp, C.c_float, vp, vp, vp]
    check(f(_p(kl), len(kl), _p(pd), len(pd), _p(b), _p(T), _p(O), _p(Kc), scale_line, log_scale_factor_line, n, _p(sk), _p(P), _p(Nn),
            _p(mn), _p(mx), _p(md), th, _p(bi), _p(bd), C.byref(stop)))
    return bi[:n], bd[:n], stop.value


LSDmatcher.FuseSearch = _lsd_fuse_search
