"""ctypes binding of libplslam_b200.so — Python mirror of the reference's operator classes.

Class and method names follow the reference (ORBextractor, LINEextractor, ORBmatcher, LSDmatcher, Optimizer);
see include/plslam_b200.h for the C ABI each method calls.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libplslam_b200.so")

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
