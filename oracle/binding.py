"""ctypes binding of oracle/liboracle.so (test infrastructure only)."""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
assert KP_DTYPE.itemsize == 28


def build(force=False):
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".cpp")]
    if force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.oracle_orb_create.restype = C.c_void_p
        _lib.oracle_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        _lib.oracle_orb_destroy.argtypes = [C.c_void_p]
        _lib.oracle_orb_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_int]
        _lib.oracle_orb_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 6
        _lib.oracle_orb_level_dims.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _lib.oracle_orb_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _lib.oracle_orb_blurred.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib.oracle_orb_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _lib.oracle_orb_selected.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        _lib.oracle_fast_atan2.restype = C.c_float
        _lib.oracle_fast_atan2.argtypes = [C.c_float, C.c_float]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OrbOracle:
    """Restatement of ORB_SLAM2::ORBextractor (reference src/ORBextractor.cc)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7):
        self.nlevels = nlevels
        self.nfeatures = nfeatures
        self.h = lib().oracle_orb_create(nfeatures, scale_factor, nlevels, ini_th, min_th)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_orb_destroy(self.h)
            self.h = None

    def tables(self):
        n = self.nlevels
        sc, isc, s2, is2 = (np.zeros(n, np.float32) for _ in range(4))
        per = np.zeros(n, np.int32)
        umax = np.zeros(16, np.int32)
        lib().oracle_orb_tables(self.h, _p(sc), _p(isc), _p(s2), _p(is2), _p(per), _p(umax))
        return dict(scale=sc, inv_scale=isc, sigma2=s2, inv_sigma2=is2, per_level=per, umax=umax)

    def extract(self, img):
        img = np.ascontiguousarray(img, np.uint8)
        cap = self.nfeatures * 2 + 64
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = lib().oracle_orb_extract(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps),
                                     _p(desc), cap)
        assert n >= 0
        return kps[:n].copy(), desc[:n].copy()

    def level_dims(self, l):
        w, h = C.c_int(), C.c_int()
        lib().oracle_orb_level_dims(self.h, l, C.byref(w), C.byref(h))
        return w.value, h.value

    def level(self, l, with_border=False):
        w, h = self.level_dims(l)
        if with_border:
            w, h = w + 38, h + 38
        out = np.zeros((h, w), np.uint8)
        lib().oracle_orb_level(self.h, l, _p(out), int(with_border))
        return out

    def blurred(self, l):
        w, h = self.level_dims(l)
        out = np.zeros((h, w), np.uint8)
        rc = lib().oracle_orb_blurred(self.h, l, _p(out))
        return out if rc == 0 else None

    def candidates(self, l):
        n = lib().oracle_orb_candidates(self.h, l, None, 0)
        out = np.zeros(max(n, 1), KP_DTYPE)
        lib().oracle_orb_candidates(self.h, l, _p(out), n)
        return out[:n]

    def selected(self, l):
        n = lib().oracle_orb_selected(self.h, l, None, 0)
        out = np.zeros(max(n, 1), KP_DTYPE)
        lib().oracle_orb_selected(self.h, l, _p(out), n)
        return out[:n]


def resize_linear_u8(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().oracle_resize_linear_u8(_p(src), src.shape[1], src.shape[0], _p(dst), dw, dh)
    return dst


def blur_u8(src, ksize):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros_like(src)
    lib().oracle_blur_u8(_p(src), src.shape[1], src.shape[0], _p(dst), ksize)
    return dst


def fast_atan2(y, x):
    return lib().oracle_fast_atan2(float(y), float(x))


def fast_detect(img, threshold):
    img = np.ascontiguousarray(img, np.uint8)
    cap = img.size
    out = np.zeros(cap, KP_DTYPE)
    n = lib().oracle_fast_detect(_p(img), img.shape[1], img.shape[0], threshold, _p(out), cap)
    return out[:n].copy()


def fast_score_map(img):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros(img.shape, np.int32)
    lib().oracle_fast_score_map(_p(img), img.shape[1], img.shape[0], _p(out))
    return out


def distribute(kps, minX, maxX, minY, maxY, N):
    kps = np.ascontiguousarray(kps)
    out = np.zeros(max(len(kps), 1), KP_DTYPE)
    n = lib().oracle_distribute(_p(kps), len(kps), minX, maxX, minY, maxY, N, _p(out))
    return out[:n].copy()
