"""Wire / disk formats around the path: the monocular trajectory writers of the reference
(System::SaveKeyFrameTrajectoryTUM, src/System.cc:396-431; System::SaveKeyFrameTrajectoryMonoKitti, :433-464) and a flat
binary dump of per-frame front-end results for offline replay (the same arrays pl_frontend_run returns).

Host-side formatting only; poses are the 4x4 float32 Tcw that Tracking / pl_pose_optimization produce."""
import struct
import numpy as np

MAGIC = b"PLSB200\x01"


def to_quaternion(R):
    """Converter::toQuaternion (src/Converter.cc): float rotation -> Eigen::Quaterniond(R) -> (x, y, z, w) as float32.
    Eigen's rotation-to-quaternion branches (trace > 0, else largest diagonal) in fp64."""
    m = np.asarray(R, np.float32).astype(np.float64)
    q = np.zeros(4)          # x y z w
    t = m[0, 0] + m[1, 1] + m[2, 2]
    if t > 0:
        t = np.sqrt(t + 1.0)
        q[3] = 0.5 * t
        t = 0.5 / t
        q[0] = (m[2, 1] - m[1, 2]) * t; q[1] = (m[0, 2] - m[2, 0]) * t; q[2] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]:
            i = 1
        if m[2, 2] > m[i, i]:
            i = 2
        j = (i + 1) % 3; k = (j + 1) % 3
        t = np.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[i] = 0.5 * t
        t = 0.5 / t
        q[3] = (m[k, j] - m[j, k]) * t; q[j] = (m[j, i] + m[i, j]) * t; q[k] = (m[k, i] + m[i, k]) * t
    return q.astype(np.float32)


def _rt(Tcw):
    """KeyFrame::GetRotation().t() and GetCameraCenter() (Ow = -Rcw^T tcw, fp32 gemm in cv's accumulation order)."""
    T = np.asarray(Tcw, np.float32).reshape(4, 4)
    Rwc = T[:3, :3].T.copy()
    f = np.float32
    Ow = np.array([-(f(f(f(Rwc[i, 0] * T[0, 3]) + f(Rwc[i, 1] * T[1, 3])) + f(Rwc[i, 2] * T[2, 3]))) for i in range(3)], np.float32)
    return Rwc, Ow


def format_keyframe_trajectory_tum(timestamps, poses_Tcw, bad=None):
    """Lines of System::SaveKeyFrameTrajectoryTUM: `f << fixed << setprecision(6) << stamp << setprecision(7) << " " << t...q`."""
    out = []
    for i, (ts, T) in enumerate(zip(timestamps, poses_Tcw)):
        if bad is not None and bad[i]:
            continue
        R, t = _rt(T)
        q = to_quaternion(R)
        out.append("%.6f" % float(ts) + "".join(" %.7f" % float(v) for v in (*t, *q)) + "\n")
    return "".join(out)


def format_keyframe_trajectory_mono_kitti(poses_Tcw, bad=None):
    """Lines of System::SaveKeyFrameTrajectoryMonoKitti: the 3x4 [Rwc | Ow] row-major with setprecision(9)."""
    out = []
    for i, T in enumerate(poses_Tcw):
        if bad is not None and bad[i]:
            continue
        R, t = _rt(T)
        vals = [R[0, 0], R[0, 1], R[0, 2], t[0], R[1, 0], R[1, 1], R[1, 2], t[1], R[2, 0], R[2, 1], R[2, 2], t[2]]
        out.append(" ".join("%.9f" % float(v) for v in vals) + "\n")
    return "".join(out)


def SaveKeyFrameTrajectoryTUM(filename, timestamps, poses_Tcw, bad=None):
    with open(filename, "w") as f:
        f.write(format_keyframe_trajectory_tum(timestamps, poses_Tcw, bad))


def SaveKeyFrameTrajectoryMonoKitti(filename, poses_Tcw, bad=None):
    with open(filename, "w") as f:
        f.write(format_keyframe_trajectory_mono_kitti(poses_Tcw, bad))


# ------------------------------------------------------------------------------------------------ binary replay dump
_FIELDS = ("kps", "desc", "n", "keylines", "ldesc", "linefunc", "nl", "pt_matches", "n_pt_matches", "line_matches",
           "n_line_matches", "poses", "inliers")


def dump_frontend(path, out, B):
    """Flat little-endian dump of one pl_frontend_run result (the dict returned by Frontend.run / alloc_outputs):
    magic, B, then per field: name length, name, dtype string length, dtype string, ndim, shape, raw bytes."""
    with open(path, "wb") as f:
        f.write(MAGIC); f.write(struct.pack("<i", B))
        for k in _FIELDS:
            a = np.ascontiguousarray(out[k])
            name = k.encode(); dt = repr(a.dtype.descr if a.dtype.fields else a.dtype.str).encode()
            f.write(struct.pack("<i", len(name))); f.write(name)
            f.write(struct.pack("<i", len(dt))); f.write(dt)
            f.write(struct.pack("<i", a.ndim)); f.write(struct.pack("<%dq" % a.ndim, *a.shape))
            f.write(a.tobytes())


def load_frontend(path):
    import ast
    with open(path, "rb") as f:
        if f.read(len(MAGIC)) != MAGIC:
            raise ValueError("not a plslam_b200 front-end dump")
        (B,) = struct.unpack("<i", f.read(4))
        out = {}
        for _ in _FIELDS:
            (ln,) = struct.unpack("<i", f.read(4)); name = f.read(ln).decode()
            (ld,) = struct.unpack("<i", f.read(4)); dt = np.dtype(ast.literal_eval(f.read(ld).decode()))
            (nd,) = struct.unpack("<i", f.read(4)); shape = struct.unpack("<%dq" % nd, f.read(8 * nd))
            out[name] = np.frombuffer(f.read(int(np.prod(shape)) * dt.itemsize), dt).reshape(shape).copy()
    return B, out


# ------------------------------------------------------------------------------------------------ the C ABI forms (wire.cu)
def _lib():
    from . import binding
    return binding


def c_format_keyframe_trajectory_tum(timestamps, poses_Tcw, bad=None):
    """pl_trajectory_format_tum: the same text, pose records computed on the GPU."""
    import ctypes as C
    b = _lib(); L = b.lib()
    ts = np.ascontiguousarray(timestamps, np.float64); P = np.ascontiguousarray(poses_Tcw, np.float32).reshape(-1, 16)
    bd = None if bad is None else np.ascontiguousarray(bad, np.uint8)
    L.pl_trajectory_format_tum.restype = C.c_longlong
    L.pl_trajectory_format_tum.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    need = L.pl_trajectory_format_tum(b._p(ts), b._p(P), b._p(bd), len(P), None, 0)
    if need < 0:
        b.check(int(need))
    buf = C.create_string_buffer(need + 1)
    L.pl_trajectory_format_tum(b._p(ts), b._p(P), b._p(bd), len(P), buf, need + 1)
    return buf.raw[:need].decode()


def c_format_keyframe_trajectory_mono_kitti(poses_Tcw, bad=None):
    import ctypes as C
    b = _lib(); L = b.lib()
    P = np.ascontiguousarray(poses_Tcw, np.float32).reshape(-1, 16)
    bd = None if bad is None else np.ascontiguousarray(bad, np.uint8)
    L.pl_trajectory_format_mono_kitti.restype = C.c_longlong
    L.pl_trajectory_format_mono_kitti.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t]
    need = L.pl_trajectory_format_mono_kitti(b._p(P), b._p(bd), len(P), None, 0)
    if need < 0:
        b.check(int(need))
    buf = C.create_string_buffer(need + 1)
    L.pl_trajectory_format_mono_kitti(b._p(P), b._p(bd), len(P), buf, need + 1)
    return buf.raw[:need].decode()


def c_save(filename, poses_Tcw, timestamps=None, bad=None):
    """pl_save_keyframe_trajectory_tum (timestamps given) / _mono_kitti."""
    import ctypes as C
    b = _lib(); L = b.lib()
    P = np.ascontiguousarray(poses_Tcw, np.float32).reshape(-1, 16)
    bd = None if bad is None else np.ascontiguousarray(bad, np.uint8)
    if timestamps is not None:
        ts = np.ascontiguousarray(timestamps, np.float64)
        L.pl_save_keyframe_trajectory_tum.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        b.check(L.pl_save_keyframe_trajectory_tum(str(filename).encode(), b._p(ts), b._p(P), b._p(bd), len(P)))
    else:
        L.pl_save_keyframe_trajectory_mono_kitti.argtypes = [C.c_char_p, C.c_void_p, C.c_void_p, C.c_int]
        b.check(L.pl_save_keyframe_trajectory_mono_kitti(str(filename).encode(), b._p(P), b._p(bd), len(P)))
