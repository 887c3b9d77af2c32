"""ctypes binding of libfastlivo_b200.so (include/fastlivo_b200.h).

No CPU fallback: if the library is missing it is built with nvcc; if there is no CUDA
device ``Handle()`` raises ``FlbError`` (FLB_ERR_NO_DEVICE).  Nothing here imports the
oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfastlivo_b200.so")

FLB_OK = 0
ERRORS = {-1: "FLB_ERR_INVALID", -2: "FLB_ERR_CUDA", -3: "FLB_ERR_NO_DEVICE", -4: "FLB_ERR_STATE",
          -5: "FLB_ERR_NUMERIC", -6: "FLB_ERR_COMM", -7: "FLB_ERR_TIMEOUT"}


class FlbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("device", C.c_int), ("cell_size", C.c_double), ("knn_max_d2", C.c_double),
                ("plane_threshold", C.c_double), ("persistent", C.c_int), ("reserved", C.c_int * 7)]


class ImuParams(C.Structure):
    """flb_imu_params (include/fastlivo_b200.h)."""
    _fields_ = [("cov_gyr", C.c_double * 3), ("cov_acc", C.c_double * 3), ("cov_bias_gyr", C.c_double * 3),
                ("cov_bias_acc", C.c_double * 3), ("G_m_s2", C.c_double), ("mean_acc_norm", C.c_double),
                ("R_LI", C.c_double * 9), ("t_LI", C.c_double * 3)]


class ImuCarry(C.Structure):
    """flb_imu_carry: last_lidar_end_time_, acc_s_last, angvel_last of ImuProcess."""
    _fields_ = [("last_lidar_end_time", C.c_double), ("acc_s_last", C.c_double * 3), ("angvel_last", C.c_double * 3)]


class State18(C.Structure):
    _fields_ = [("rot", C.c_double * 9), ("pos", C.c_double * 3), ("vel", C.c_double * 3),
                ("bg", C.c_double * 3), ("ba", C.c_double * 3), ("grav", C.c_double * 3),
                ("cov", C.c_double * 324)]

    @classmethod
    def make(cls, R, p, vel=None, bg=None, ba=None, grav=None, cov=None):
        s = cls()
        s.rot[:] = np.asarray(R, np.float64).ravel()
        s.pos[:] = np.asarray(p, np.float64)
        s.vel[:] = np.zeros(3) if vel is None else vel
        s.bg[:] = np.zeros(3) if bg is None else bg
        s.ba[:] = np.zeros(3) if ba is None else ba
        s.grav[:] = np.zeros(3) if grav is None else grav
        s.cov[:] = (np.eye(18) if cov is None else np.asarray(cov, np.float64)).ravel()
        return s

    @classmethod
    def from_frame(cls, frame, prop=True):
        R = frame["R_prop"] if prop else frame["R_true"]
        p = frame["p_prop"] if prop else frame["p_true"]
        return cls.make(R, p, frame["vel"], frame["bg"], frame["ba"], frame["grav"], frame["cov"])

    def copy(self):
        o = State18()
        C.memmove(C.byref(o), C.byref(self), C.sizeof(State18))
        return o

    @property
    def R(self):
        return np.array(self.rot[:]).reshape(3, 3)

    @property
    def p(self):
        return np.array(self.pos[:])

    @property
    def P(self):
        return np.array(self.cov[:]).reshape(18, 18)

    def vector(self):
        return np.concatenate([self.rot[:], self.pos[:], self.vel[:], self.bg[:], self.ba[:], self.grav[:]])


class LioParams(C.Structure):
    _fields_ = [("R_LI", C.c_double * 9), ("t_LI", C.c_double * 3), ("laser_point_cov", C.c_double),
                ("max_iteration", C.c_int), ("conv_rot_deg", C.c_double), ("conv_pos_cm", C.c_double)]


class LioReport(C.Structure):
    _fields_ = [("passes", C.c_int), ("knn_passes", C.c_int), ("n_eff_last", C.c_int),
                ("res_mean_last", C.c_double), ("rows_total", C.c_int64), ("converged_last", C.c_int),
                ("status", C.c_int)]


class NormalEq(C.Structure):
    _fields_ = [("width", C.c_int), ("n_eff", C.c_int), ("sum_abs_res", C.c_double),
                ("HTH", C.c_double * 144), ("HTh", C.c_double * 12)]


class StateIkfom(C.Structure):
    """flb_state_ikfom == state_ikfom (reference include/use-ikfom.hpp:12-21); quaternions (x, y, z, w)."""
    _fields_ = [("pos", C.c_double * 3), ("rot", C.c_double * 4), ("offset_R_L_I", C.c_double * 4),
                ("offset_T_L_I", C.c_double * 3), ("vel", C.c_double * 3), ("bg", C.c_double * 3),
                ("ba", C.c_double * 3), ("grav", C.c_double * 3), ("P", C.c_double * 529)]

    def vector(self):
        return np.concatenate([self.pos[:], self.rot[:], self.offset_R_L_I[:], self.offset_T_L_I[:], self.vel[:],
                               self.bg[:], self.ba[:], self.grav[:]])

    @property
    def cov(self):
        return np.array(self.P[:]).reshape(23, 23)


class IkfomParams(C.Structure):
    _fields_ = [("laser_point_cov", C.c_double), ("max_iteration", C.c_int), ("limit", C.c_double * 23)]


class IkfomReport(C.Structure):
    _fields_ = [("passes", C.c_int), ("knn_passes", C.c_int), ("n_eff_last", C.c_int), ("converged_last", C.c_int),
                ("res_mean_last", C.c_double), ("rows_total", C.c_int64), ("status", C.c_int)]


class Camera(C.Structure):
    _fields_ = [("width", C.c_int), ("height", C.c_int), ("fx", C.c_double), ("fy", C.c_double),
                ("cx", C.c_double), ("cy", C.c_double), ("d", C.c_double * 5)]


class VioParams(C.Structure):
    _fields_ = [("Rcl", C.c_double * 9), ("Pcl", C.c_double * 3), ("R_LI", C.c_double * 9),
                ("t_LI", C.c_double * 3), ("img_point_cov", C.c_double), ("max_iteration", C.c_int),
                ("conv_rot_deg", C.c_float), ("conv_pos_cm", C.c_float), ("force_all_passes", C.c_int)]


class VioReport(C.Structure):
    _fields_ = [("passes", C.c_int * 3), ("last_error", C.c_float * 3), ("rows_total", C.c_int64),
                ("skipped_last", C.c_int), ("cov_updated", C.c_int), ("status", C.c_int)]


class FrameInputs(C.Structure):
    """flb_frame_inputs (include/fastlivo_b200.h)."""
    _fields_ = [("scan_xyz", C.c_void_p), ("n_scan", C.c_int), ("scan_stride", C.c_int),
                ("gray", C.c_void_p), ("width", C.c_int), ("height", C.c_int), ("stride_bytes", C.c_int),
                ("patch_pos", C.c_void_p), ("patch", C.c_void_p), ("search_level", C.c_void_p), ("Pn", C.c_int),
                ("x", C.POINTER(State18)), ("x_prop", C.POINTER(State18))]


class VmapParams(C.Structure):
    """flb_vmap_params (include/fastlivo_b200.h)."""
    _fields_ = [("grid_size", C.c_int), ("ncc_en", C.c_int), ("outlier_threshold", C.c_double), ("ncc_thre", C.c_double),
                ("Rcl", C.c_double * 9), ("Pcl", C.c_double * 3), ("R_LI", C.c_double * 9), ("t_LI", C.c_double * 3)]


class VioEq(C.Structure):
    _fields_ = [("HTH", C.c_double * 36), ("HTz", C.c_double * 6), ("error", C.c_float),
                ("n_meas", C.c_int64), ("skipped", C.c_int)]


# every symbol include/fastlivo_b200.h declares (checked by the CPU-only test tier)
SYMBOLS = ["flb_abi_version", "flb_create", "flb_destroy", "flb_last_error", "flb_set_stream", "flb_synchronize", "flb_host_alloc", "flb_host_free",
           "flb_map_upload", "flb_map_add_points", "flb_map_delete_boxes", "flb_map_size", "flb_map_download", "flb_scan_upload", "flb_knn", "flb_lio_pass", "flb_lio_export", "flb_lio_update", "flb_lio_update_ikfom",
           "flb_image_upload", "flb_patches_upload", "flb_camera_set", "flb_vio_pass", "flb_vio_export",
           "flb_vio_update", "flb_state_upload", "flb_state_download", "flb_lio_update_enqueue",
           "flb_vio_update_enqueue", "flb_state_reset_enqueue", "flb_state_set_prior_enqueue", "flb_profile_start", "flb_profile_stop",
           "flb_launch_count", "flb_trace_enable", "flb_trace_download", "flb_comm_unique_id", "flb_comm_init", "flb_comm_destroy", "flb_p2p_export", "flb_p2p_attach", "flb_p2p_detach",
           "flb_imu_undistort", "flb_visual_candidates", "flb_vio_errors",
           "flb_vio_update_level", "flb_state_download_enqueue", "flb_state_download_wait", "flb_frame_enqueue",
           "flb_vmap_reset", "flb_vmap_select", "flb_vmap_selected", "flb_vmap_grow", "flb_vmap_add_observations", "flb_vmap_counts",
           "flb_vmap_map_value", "flb_vmap_dump", "flb_colorize", "flb_voxel_grid",
           "flb_batch_begin", "flb_batch_set_frame", "flb_batch_state_reset_enqueue", "flb_batch_update_enqueue", "flb_batch_state_download",
           "flb_debug_set_packet_epoch", "flb_debug_block_stamps", "flb_debug_vio_stamps", "flb_debug_scan_order", "flb_debug_set_scan_sort"]


def lib_path() -> str:
    return _SO


def build(force: bool = False) -> str:
    """Compile libfastlivo_b200.so for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = [os.path.join(_HERE, "csrc", f) for f in os.listdir(os.path.join(_HERE, "csrc"))]
    srcs.append(os.path.join(_HERE, "..", "include", "fastlivo_b200.h"))
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(os.path.getmtime(s) for s in srcs)
    if force or stale:
        if not os.path.exists("/usr/local/cuda/bin/nvcc") and os.path.exists(_SO):
            return _SO  # a box without nvcc: use the prebuilt library that travelled with the repo
        subprocess.check_call(["bash", os.path.join(_HERE, "build.sh")])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp = C.c_void_p
        L.flb_abi_version.restype = C.c_int
        L.flb_last_error.restype = C.c_char_p
        L.flb_last_error.argtypes = [vp]
        L.flb_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
        L.flb_destroy.argtypes = [vp]
        L.flb_set_stream.argtypes = [vp, vp]
        L.flb_synchronize.argtypes = [vp]
        L.flb_host_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp)]
        L.flb_host_free.argtypes = [vp, vp]
        L.flb_map_upload.argtypes = [vp, vp, C.c_int, C.c_int]
        L.flb_scan_upload.argtypes = [vp, vp, C.c_int, C.c_int]
        L.flb_map_add_points.argtypes = [vp, vp, C.c_int, C.c_int, C.c_float]
        L.flb_map_delete_boxes.argtypes = [vp, vp, C.c_int]
        L.flb_map_size.argtypes = [vp]
        L.flb_map_download.argtypes = [vp, vp, C.c_int, C.POINTER(C.c_int)]
        L.flb_knn.argtypes = [vp, vp, C.c_int, vp, vp]
        L.flb_lio_pass.argtypes = [vp, C.POINTER(LioParams), vp, vp, C.c_int, C.c_int, C.POINTER(NormalEq)]
        L.flb_lio_export.argtypes = [vp] + [vp] * 9 + [C.POINTER(C.c_int)]
        L.flb_lio_update.argtypes = [vp, C.POINTER(LioParams), C.POINTER(State18), C.POINTER(State18), C.POINTER(LioReport)]
        L.flb_lio_update_ikfom.argtypes = [vp, C.POINTER(IkfomParams), C.POINTER(StateIkfom), C.POINTER(IkfomReport)]
        L.flb_imu_undistort.argtypes = [vp, C.POINTER(ImuParams), C.POINTER(ImuCarry), vp, C.c_int, C.c_double, C.c_double,
                                        vp, C.c_int, C.c_int, C.c_int, vp, vp, C.POINTER(C.c_int)]
        L.flb_visual_candidates.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp]
        L.flb_image_upload.argtypes = [vp, vp, C.c_int, C.c_int, C.c_int]
        L.flb_patches_upload.argtypes = [vp, vp, vp, vp, C.c_int]
        L.flb_camera_set.argtypes = [vp, C.POINTER(Camera)]
        L.flb_vio_pass.argtypes = [vp, C.POINTER(VioParams), vp, vp, C.c_int, C.POINTER(VioEq)]
        L.flb_vio_export.argtypes = [vp, vp, vp, vp]
        L.flb_vio_errors.argtypes = [vp, vp, C.c_int]
        L.flb_debug_scan_order.argtypes = [vp, vp, C.c_int]
        L.flb_debug_set_scan_sort.argtypes = [vp, C.c_int]
        L.flb_vio_update_level.argtypes = [vp, C.POINTER(VioParams), C.c_int, C.c_float, C.POINTER(State18), C.POINTER(State18),
                                           C.POINTER(C.c_float), vp, C.POINTER(VioReport)]
        L.flb_batch_begin.argtypes = [vp, C.c_int, C.c_int]
        L.flb_batch_set_frame.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int, C.POINTER(State18), C.POINTER(State18)]
        L.flb_batch_state_reset_enqueue.argtypes = [vp]
        L.flb_batch_update_enqueue.argtypes = [vp, C.POINTER(LioParams), C.POINTER(VioParams)]
        L.flb_batch_state_download.argtypes = [vp, C.c_int, C.POINTER(State18), C.POINTER(LioReport), C.POINTER(VioReport)]
        L.flb_voxel_grid.argtypes = [vp, vp, C.c_int, C.c_int, C.c_float, vp, C.c_int, C.POINTER(C.c_int)]
        L.flb_colorize.argtypes = [vp, vp, vp, vp, C.c_int, vp, C.c_int, C.c_int, vp, vp]
        L.flb_vmap_reset.argtypes = [vp, C.POINTER(VmapParams)]
        L.flb_vmap_select.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp]
        L.flb_vmap_selected.argtypes = [vp, C.c_int, C.POINTER(C.c_int)] + [vp] * 6
        L.flb_vmap_grow.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int]
        L.flb_vmap_add_observations.argtypes = [vp, vp, vp, C.c_int]
        L.flb_vmap_counts.argtypes = [vp] + [C.POINTER(C.c_int)] * 5
        L.flb_vmap_map_value.argtypes = [vp, vp, C.c_int]
        L.flb_vmap_dump.argtypes = [vp, C.c_int, C.c_int] + [vp] * 7
        L.flb_vio_update.argtypes = [vp, C.POINTER(VioParams), C.POINTER(State18), C.POINTER(State18), C.POINTER(VioReport)]
        L.flb_state_upload.argtypes = [vp, C.POINTER(State18), C.POINTER(State18)]
        L.flb_state_download.argtypes = [vp, C.POINTER(State18), C.POINTER(LioReport), C.POINTER(VioReport)]
        L.flb_lio_update_enqueue.argtypes = [vp, C.POINTER(LioParams)]
        L.flb_frame_enqueue.argtypes = [vp, C.POINTER(FrameInputs), C.POINTER(LioParams), C.POINTER(VioParams), C.c_int]
        L.flb_state_download_enqueue.argtypes = [vp, C.c_int]
        L.flb_state_download_wait.argtypes = [vp, C.c_int, C.POINTER(State18), C.POINTER(LioReport), C.POINTER(VioReport)]
        L.flb_vio_update_enqueue.argtypes = [vp, C.POINTER(VioParams)]
        L.flb_state_reset_enqueue.argtypes = [vp]
        L.flb_state_set_prior_enqueue.argtypes = [vp]
        L.flb_profile_start.argtypes = [vp]
        L.flb_profile_stop.argtypes = [vp, vp, vp]
        L.flb_trace_enable.argtypes = [vp, C.c_int]
        L.flb_trace_download.argtypes = [vp, C.c_int, vp, C.c_int, C.POINTER(C.c_int)]
        L.flb_launch_count.restype = C.c_int64
        L.flb_launch_count.argtypes = [vp]
        L.flb_comm_unique_id.argtypes = [vp]
        L.flb_comm_init.argtypes = [vp, vp, C.c_int, C.c_int]
        L.flb_comm_destroy.argtypes = [vp]
        L.flb_p2p_export.argtypes = [vp, vp]
        L.flb_p2p_attach.argtypes = [vp, C.c_int, C.c_int, vp]
        L.flb_p2p_detach.argtypes = [vp]
        _lib = L
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def lio_params(frame, max_iteration, early_stop=True) -> LioParams:
    p = LioParams()
    p.R_LI[:] = np.asarray(frame["R_LI"], np.float64).ravel()
    p.t_LI[:] = np.asarray(frame["t_LI"], np.float64)
    p.laser_point_cov = frame["cfg"].laser_point_cov
    p.max_iteration = max_iteration
    p.conv_rot_deg = 0.01 if early_stop else 0.0
    p.conv_pos_cm = 0.015 if early_stop else 0.0
    return p


def vio_params(frame, max_iteration, early_stop=True, force_all_passes=False) -> VioParams:
    p = VioParams()
    p.Rcl[:] = np.asarray(frame["Rcl"], np.float64).ravel()
    p.Pcl[:] = np.asarray(frame["Pcl"], np.float64)
    p.R_LI[:] = np.asarray(frame["R_LI"], np.float64).ravel()
    p.t_LI[:] = np.asarray(frame["t_LI"], np.float64)
    p.img_point_cov = frame["cfg"].img_point_cov
    p.max_iteration = max_iteration
    p.conv_rot_deg = 0.001 if early_stop else 0.0
    p.conv_pos_cm = 0.001 if early_stop else 0.0
    p.force_all_passes = int(force_all_passes)
    return p


class Handle:
    """Owns one flb_handle.  Method names follow the C ABI."""

    def __init__(self, device=0, cell_size=0.6, knn_max_d2=5.0, plane_threshold=0.1, persistent=1):
        self.L = lib()
        cfg = Config()
        cfg.device, cfg.cell_size, cfg.knn_max_d2, cfg.plane_threshold, cfg.persistent = \
            device, cell_size, knn_max_d2, plane_threshold, persistent
        self.h = C.c_void_p()
        rc = self.L.flb_create(C.byref(cfg), C.byref(self.h))
        if rc != FLB_OK:
            raise FlbError(rc, self.L.flb_last_error(None).decode())
        self.N = self.M = self.Pn = 0

    def close(self):
        if getattr(self, "h", None):
            self.L.flb_synchronize(self.h)
            for ptr in getattr(self, "_pinned", []):
                self.L.flb_host_free(self.h, ptr)
            self._pinned = []
            self.L.flb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != FLB_OK:
            raise FlbError(rc, self.L.flb_last_error(self.h).decode())

    # ---- uploads
    def map_upload(self, xyz):
        a = np.ascontiguousarray(xyz, np.float32)
        self._ck(self.L.flb_map_upload(self.h, _p(a), a.shape[0], a.shape[1]))
        self.M = a.shape[0]

    def map_add_points(self, world_xyz, downsample_size):
        a = np.ascontiguousarray(world_xyz, np.float32)
        self._ck(self.L.flb_map_add_points(self.h, _p(a), a.shape[0], a.shape[1], C.c_float(downsample_size)))
        self.M = self.L.flb_map_size(self.h)

    def map_delete_boxes(self, boxes):
        b = np.ascontiguousarray(boxes, np.float32).reshape(-1, 6)
        self._ck(self.L.flb_map_delete_boxes(self.h, _p(b), len(b)))
        self.M = self.L.flb_map_size(self.h)

    def map_download(self):
        n = self.L.flb_map_size(self.h)
        out = np.empty((n, 3), np.float32)
        m = C.c_int()
        self._ck(self.L.flb_map_download(self.h, _p(out), n, C.byref(m)))
        return out

    def voxel_grid(self, xyz, leaf):
        """flb_voxel_grid: pcl::VoxelGrid centroids (m, 3) float32."""
        a = np.ascontiguousarray(xyz, np.float32)
        out = np.zeros((len(a), 3), np.float32)
        n = C.c_int()
        self._ck(self.L.flb_voxel_grid(self.h, _p(a), len(a), a.shape[1] if a.ndim == 2 else 3, C.c_float(leaf), _p(out), len(a), C.byref(n)))
        return out[:n.value].copy()

    def scan_upload(self, body_xyz):
        a = np.ascontiguousarray(body_xyz, np.float32)
        self._ck(self.L.flb_scan_upload(self.h, _p(a), a.shape[0], a.shape[1] if a.ndim == 2 else 3))
        self.N = a.shape[0]

    def debug_set_scan_sort(self, mode):
        """0 automatic, 1 one-block ordering kernel, 2 device-wide sort (test aid)."""
        self._ck(self.L.flb_debug_set_scan_sort(self.h, int(mode)))

    def debug_scan_order(self):
        """Index in the caller's array of the point at each position of the uploaded (ordered) scan (test aid)."""
        out = np.zeros(max(self.N, 1), np.int32)
        self._ck(self.L.flb_debug_scan_order(self.h, _p(out), out.shape[0]))
        return out[:self.N]

    def image_upload(self, gray):
        a = np.ascontiguousarray(gray, np.uint8)
        self._ck(self.L.flb_image_upload(self.h, _p(a), a.shape[1], a.shape[0], a.shape[1]))

    def patches_upload(self, pos, patch, level):
        pos = np.ascontiguousarray(pos, np.float64).reshape(-1, 3)
        patch = np.ascontiguousarray(patch, np.float32).reshape(len(pos), 192)
        level = np.ascontiguousarray(level, np.int32)
        self._ck(self.L.flb_patches_upload(self.h, _p(pos), _p(patch), _p(level), len(pos)))
        self.Pn = len(pos)

    def camera_set(self, cam: dict):
        c = Camera()
        c.width, c.height = cam["width"], cam["height"]
        c.fx, c.fy, c.cx, c.cy = cam["fx"], cam["fy"], cam["cx"], cam["cy"]
        c.d[:] = cam["d"]
        self._ck(self.L.flb_camera_set(self.h, C.byref(c)))

    def load_frame(self, frame):
        """Upload everything a synthetic frame holds."""
        self.map_upload(frame["map_xyz"])
        self.scan_upload(frame["scan_body"])
        if len(frame["patch_pos"]):
            self.camera_set(frame["cam"])
            self.image_upload(frame["image"])
            self.patches_upload(frame["patch_pos"], frame["patch_ref"], frame["patch_level"])

    # ---- kNN
    def knn(self, q_world):
        q = np.ascontiguousarray(q_world, np.float32)
        idx = np.empty((len(q), 5), np.int32)
        d2 = np.empty((len(q), 5), np.float32)
        self._ck(self.L.flb_knn(self.h, _p(q), len(q), _p(idx), _p(d2)))
        return idx, d2

    # ---- LIO
    def lio_pass(self, prm: LioParams, R, p, rematch: bool, width=6, export=True):
        R = np.ascontiguousarray(R, np.float64)
        p = np.ascontiguousarray(p, np.float64)
        eq = NormalEq()
        self._ck(self.L.flb_lio_pass(self.h, C.byref(prm), _p(R), _p(p), int(rematch), width, C.byref(eq)))
        out = dict(n=eq.n_eff, total_residual=eq.sum_abs_res,
                   HTH=np.array(eq.HTH[:width * width]).reshape(width, width), HTh=np.array(eq.HTh[:width]))
        if export:
            N = self.N
            world = np.empty((N, 3), np.float32)
            nn_idx = np.empty((N, 5), np.int32)
            nn_d2 = np.empty((N, 5), np.float32)
            pabcd = np.empty((N, 4), np.float32)
            pd2 = np.empty(N, np.float32)
            sel = np.empty(N, np.uint8)
            rows = np.empty((N, width))
            meas = np.empty(N)
            sel_idx = np.empty(N, np.int32)
            n = C.c_int()
            self._ck(self.L.flb_lio_export(self.h, _p(world), _p(nn_idx), _p(nn_d2), _p(pabcd), _p(pd2), _p(sel), _p(rows),
                                           _p(meas), _p(sel_idx), C.byref(n)))
            out.update(world=world, nn_idx=nn_idx, nn_d2=nn_d2, pabcd=pabcd, pd2=pd2, rowmask=sel,
                       rows=rows[:n.value], meas=meas[:n.value], sel_idx=sel_idx[:n.value])
        return out

    def lio_update(self, prm: LioParams, x: State18, x_prop: State18) -> LioReport:
        rep = LioReport()
        self._ck(self.L.flb_lio_update(self.h, C.byref(prm), C.byref(x), C.byref(x_prop), C.byref(rep)))
        return rep

    def lio_update_ikfom(self, prm: IkfomParams, x: StateIkfom) -> IkfomReport:
        rep = IkfomReport()
        self._ck(self.L.flb_lio_update_ikfom(self.h, C.byref(prm), C.byref(x), C.byref(rep)))
        return rep

    # ---- VIO
    def vio_pass(self, prm: VioParams, R, p, level: int, export=True):
        R = np.ascontiguousarray(R, np.float64)
        p = np.ascontiguousarray(p, np.float64)
        eq = VioEq()
        self._ck(self.L.flb_vio_pass(self.h, C.byref(prm), _p(R), _p(p), level, C.byref(eq)))
        out = dict(error=np.float32(eq.error), n_meas=eq.n_meas, skipped=eq.skipped,
                   HTH6=np.array(eq.HTH[:]).reshape(6, 6), HTz6=np.array(eq.HTz[:]))
        if export:
            z = np.empty(self.Pn * 64)
            H = np.empty((self.Pn * 64, 6))
            err = np.empty(self.Pn, np.float32)
            self._ck(self.L.flb_vio_export(self.h, _p(z), _p(H), _p(err)))
            out.update(z=z, H_sub=H, errors=err)
        return out

    def vio_update_level(self, prm: VioParams, level: int, total_residual: float, x: State18, x_prop: State18):
        """flb_vio_update_level == LidarSelector::UpdateState.  Returns (last_error, G[18, 6], report)."""
        rep = VioReport()
        le = C.c_float()
        G = np.zeros((18, 6))
        self._ck(self.L.flb_vio_update_level(self.h, C.byref(prm), int(level), C.c_float(total_residual), C.byref(x), C.byref(x_prop),
                                             C.byref(le), _p(G), C.byref(rep)))
        return le.value, G, rep

    def vio_errors(self):
        """sub_sparse_map->errors as the last update's last pass left them (flb_vio_errors)."""
        err = np.zeros(self.Pn, np.float32)
        self._ck(self.L.flb_vio_errors(self.h, _p(err), self.Pn))
        return err

    def vio_update(self, prm: VioParams, x: State18, x_prop: State18) -> VioReport:
        rep = VioReport()
        self._ck(self.L.flb_vio_update(self.h, C.byref(prm), C.byref(x), C.byref(x_prop), C.byref(rep)))
        return rep

    # ---- visual-map growth: candidate scoring (row f4)
    def visual_candidates(self, Rcw, Pcw, world_xyz, grid_size, border, map_value):
        """flb_visual_candidates on the uploaded image / camera.  Returns (map_value_out, winner) per grid cell."""
        R = np.ascontiguousarray(Rcw, np.float64)
        P = np.ascontiguousarray(Pcw, np.float64)
        pts = np.ascontiguousarray(world_xyz, np.float32)
        mv = np.ascontiguousarray(map_value, np.float32).copy()
        win = np.zeros(len(mv), np.int32)
        self._ck(self.L.flb_visual_candidates(self.h, R.ctypes.data_as(C.c_void_p), P.ctypes.data_as(C.c_void_p),
                                              pts.ctypes.data_as(C.c_void_p), len(pts), pts.shape[1] if len(pts) else 3,
                                              int(grid_size), int(border), mv.ctypes.data_as(C.c_void_p),
                                              win.ctypes.data_as(C.c_void_p)))
        return mv, win

    # ---- batched frames
    def batch_begin(self, B, max_points):
        self._ck(self.L.flb_batch_begin(self.h, int(B), int(max_points)))

    def batch_set_frame(self, b, body_xyz, x: State18, x_prop: State18):
        a = np.ascontiguousarray(body_xyz, np.float32)
        self._ck(self.L.flb_batch_set_frame(self.h, int(b), _p(a), a.shape[0], a.shape[1], C.byref(x), C.byref(x_prop)))

    def batch_state_reset_enqueue(self):
        self._ck(self.L.flb_batch_state_reset_enqueue(self.h))

    def batch_update_enqueue(self, lprm, vprm=None):
        self._ck(self.L.flb_batch_update_enqueue(self.h, C.byref(lprm), C.byref(vprm) if vprm is not None else None))

    def batch_state_download(self, b):
        x, lr, vr = State18(), LioReport(), VioReport()
        self._ck(self.L.flb_batch_state_download(self.h, int(b), C.byref(x), C.byref(lr), C.byref(vr)))
        return x, lr, vr

    # ---- device-resident visual map (rows f2 / f4)
    def vmap_reset(self, seq_or_frame, grid_size=40, outlier_threshold=100.0, ncc_en=0, ncc_thre=0.0):
        """flb_vmap_reset; the extrinsics come from a synthetic frame / sequence dict (Rcl, Pcl, R_LI, t_LI)."""
        p = VmapParams()
        p.grid_size, p.ncc_en, p.outlier_threshold, p.ncc_thre = int(grid_size), int(ncc_en), float(outlier_threshold), float(ncc_thre)
        p.Rcl[:] = np.asarray(seq_or_frame["Rcl"], np.float64).ravel()
        p.Pcl[:] = np.asarray(seq_or_frame["Pcl"], np.float64)
        p.R_LI[:] = np.asarray(seq_or_frame["R_LI"], np.float64).ravel()
        p.t_LI[:] = np.asarray(seq_or_frame["t_LI"], np.float64)
        self._ck(self.L.flb_vmap_reset(self.h, C.byref(p)))

    @staticmethod
    def _pose(Rcw, Pcw):
        if Rcw is None:
            return None, None, None, None
        R = np.ascontiguousarray(Rcw, np.float64)
        P = np.ascontiguousarray(Pcw, np.float64)
        return R, P, R.ctypes.data_as(C.c_void_p), P.ctypes.data_as(C.c_void_p)

    def vmap_select(self, Rcw, Pcw, pg_down, blocking=True):
        """flb_vmap_select.  blocking: returns the number of selected patches; else enqueue-only (returns None)."""
        R, P, pr, pp = self._pose(Rcw, Pcw)
        pg = np.ascontiguousarray(pg_down, np.float32).reshape(-1, 3)
        n = C.c_int()
        self._ck(self.L.flb_vmap_select(self.h, pr, pp, _p(pg), len(pg), 3, C.byref(n) if blocking else None))
        if blocking:
            self.Pn = n.value
            return n.value
        return None

    def vmap_selected(self):
        n = C.c_int()
        self._ck(self.L.flb_vmap_selected(self.h, 0, C.byref(n), None, None, None, None, None, None))
        m = n.value
        o = dict(index=np.zeros(m, np.int32), point=np.zeros(m, np.int32), search_level=np.zeros(m, np.int32),
                 error=np.zeros(m, np.float32), pos=np.zeros((m, 3)), patch=np.zeros((m, 192), np.float32))
        if m:
            self._ck(self.L.flb_vmap_selected(self.h, m, C.byref(n), _p(o["index"]), _p(o["point"]), _p(o["search_level"]), _p(o["error"]),
                                              _p(o["pos"]), _p(o["patch"])))
        self.Pn = m
        return o

    def vmap_grow(self, Rcw, Pcw, pg, frame_id):
        R, P, pr, pp = self._pose(Rcw, Pcw)
        pg = np.ascontiguousarray(pg, np.float32).reshape(-1, 3)
        self._ck(self.L.flb_vmap_grow(self.h, pr, pp, _p(pg), len(pg), 3, int(frame_id)))

    def vmap_add_observations(self, Rcw, Pcw, frame_id):
        R, P, pr, pp = self._pose(Rcw, Pcw)
        self._ck(self.L.flb_vmap_add_observations(self.h, pr, pp, int(frame_id)))

    def colorize(self, Rcw, Pcw, bgr, world_xyz):
        """flb_colorize: (rgb (n, 3) uint8, valid (n,) bool)."""
        R, P, pr, pp = self._pose(Rcw, Pcw)
        img = np.ascontiguousarray(bgr, np.uint8)
        pts = np.ascontiguousarray(world_xyz, np.float32).reshape(-1, 3)
        rgb = np.zeros((len(pts), 3), np.uint8)
        val = np.zeros(len(pts), np.uint8)
        self._ck(self.L.flb_colorize(self.h, pr, pp, _p(img), img.shape[1] * 3, _p(pts), len(pts), 3, _p(rgb), _p(val)))
        return rgb, val.astype(bool)

    def vmap_counts(self):
        v = [C.c_int() for _ in range(5)]
        self._ck(self.L.flb_vmap_counts(self.h, *[C.byref(x) for x in v]))
        return dict(points=v[0].value, features=v[1].value, images=v[2].value, selected=v[3].value, last_added=v[4].value)

    def vmap_map_value(self, length):
        out = np.zeros(length, np.float32)
        self._ck(self.L.flb_vmap_map_value(self.h, _p(out), length))
        return out

    def vmap_dump(self):
        c = self.vmap_counts()
        n, m = c["points"], c["features"]
        d = dict(pos=np.zeros((n, 3)), value=np.zeros(n, np.float32), n_obs=np.zeros(n, np.int32), obs=np.zeros((n, 20), np.int32),
                 ft_geo=np.zeros((m, 17)), ft_score=np.zeros(m, np.float32), ft_level_id_img=np.zeros((m, 3), np.int32))
        self._ck(self.L.flb_vmap_dump(self.h, n, m, _p(d["pos"]), _p(d["value"]), _p(d["n_obs"]), _p(d["obs"]), _p(d["ft_geo"]),
                                      _p(d["ft_score"]), _p(d["ft_level_id_img"])))
        return d

    # ---- IMU propagation + undistortion (row f3)
    def imu_undistort(self, prm: ImuParams, carry: ImuCarry, v_imu, pcl_beg_time, pcl_end_time, pts, offset_index=3):
        """flb_imu_undistort on the device state.  v_imu (K,7) [t, gyr, acc]; pts (n, stride) float32 with x,y,z
        first and the time offset (ms) at column offset_index.  Returns (xyz (n,3) float32, IMUpose (n_poses,22));
        carry is updated in place, the device state is propagated in place."""
        v = np.ascontiguousarray(v_imu, np.float64)
        pts = np.ascontiguousarray(pts, np.float32)
        n = pts.shape[0]
        out = np.zeros((n, 3), np.float32)
        poses = np.zeros((len(v), 22), np.float64)
        npo = C.c_int()
        self._ck(self.L.flb_imu_undistort(self.h, C.byref(prm), C.byref(carry), v.ctypes.data_as(C.c_void_p), len(v),
                                          float(pcl_beg_time), float(pcl_end_time), pts.ctypes.data_as(C.c_void_p),
                                          pts.shape[1] if n else 4, offset_index, n, out.ctypes.data_as(C.c_void_p),
                                          poses.ctypes.data_as(C.c_void_p), C.byref(npo)))
        return out, poses[:npo.value]

    # ---- device-resident loop
    def state_upload(self, x: State18, x_prop: State18):
        self._ck(self.L.flb_state_upload(self.h, C.byref(x), C.byref(x_prop)))

    def state_download(self):
        x, lr, vr = State18(), LioReport(), VioReport()
        self._ck(self.L.flb_state_download(self.h, C.byref(x), C.byref(lr), C.byref(vr)))
        return x, lr, vr

    def frame_inputs(self, scan, x, x_prop, image=None, pos=None, patch=None, level=None):
        """A reusable flb_frame_inputs over caller-owned numpy arrays (keep them alive; pass page-locked arrays from
        pinned_like() to let the copy engine read them in place)."""
        fi = FrameInputs()
        fi._keep = (scan, image, pos, patch, level, x, x_prop)
        fi.scan_xyz, fi.n_scan, fi.scan_stride = scan.ctypes.data, scan.shape[0], scan.shape[1]
        if image is not None:
            fi.gray, fi.width, fi.height, fi.stride_bytes = image.ctypes.data, image.shape[1], image.shape[0], image.shape[1]
        if pos is not None:
            fi.patch_pos, fi.patch, fi.search_level, fi.Pn = pos.ctypes.data, patch.ctypes.data, level.ctypes.data, len(pos)
            self.Pn = len(pos)
        fi.x, fi.x_prop = C.pointer(x), C.pointer(x_prop)
        self.N = scan.shape[0]
        return fi

    def frame_enqueue(self, fi, lprm, vprm=None, slot=-1):
        self._ck(self.L.flb_frame_enqueue(self.h, C.byref(fi), C.byref(lprm), C.byref(vprm) if vprm is not None else None, int(slot)))

    def state_download_enqueue(self, slot):
        self._ck(self.L.flb_state_download_enqueue(self.h, int(slot)))

    def state_download_wait(self, slot):
        x, lr, vr = State18(), LioReport(), VioReport()
        self._ck(self.L.flb_state_download_wait(self.h, int(slot), C.byref(x), C.byref(lr), C.byref(vr)))
        return x, lr, vr

    def lio_update_enqueue(self, prm):
        self._ck(self.L.flb_lio_update_enqueue(self.h, C.byref(prm)))

    def vio_update_enqueue(self, prm):
        self._ck(self.L.flb_vio_update_enqueue(self.h, C.byref(prm)))

    def state_reset_enqueue(self):
        self._ck(self.L.flb_state_reset_enqueue(self.h))

    def state_set_prior_enqueue(self):
        self._ck(self.L.flb_state_set_prior_enqueue(self.h))

    def pinned_like(self, a):
        """A page-locked copy of numpy array `a` (flb_host_alloc); freed with the handle."""
        a = np.ascontiguousarray(a)
        ptr = C.c_void_p()
        self._ck(self.L.flb_host_alloc(self.h, max(a.nbytes, 1), C.byref(ptr)))
        self._pinned = getattr(self, "_pinned", [])
        self._pinned.append(ptr)
        buf = (C.c_ubyte * max(a.nbytes, 1)).from_address(ptr.value)
        out = np.frombuffer(buf, dtype=a.dtype, count=a.size).reshape(a.shape)
        out[...] = a
        return out

    def set_stream(self, stream_ptr):
        self._ck(self.L.flb_set_stream(self.h, C.c_void_p(stream_ptr)))

    def synchronize(self):
        self._ck(self.L.flb_synchronize(self.h))

    def launch_count(self) -> int:
        return int(self.L.flb_launch_count(self.h))

    def profile_start(self):
        self._ck(self.L.flb_profile_start(self.h))

    def profile_stop(self):
        ms = np.zeros(4)
        n = np.zeros(4, np.int64)
        self._ck(self.L.flb_profile_stop(self.h, _p(ms), _p(n)))
        return ms, n

    def trace_enable(self, on=True):
        self._ck(self.L.flb_trace_enable(self.h, int(on)))

    def trace_download(self, which):
        us = np.zeros(128)
        n = C.c_int()
        self._ck(self.L.flb_trace_download(self.h, which, _p(us), 128, C.byref(n)))
        return us[:n.value]

    # ---- multi-GPU
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        rc = lib().flb_comm_unique_id(buf)
        if rc != FLB_OK:
            raise FlbError(rc, lib().flb_last_error(None).decode())
        return buf.raw

    def p2p_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._ck(self.L.flb_p2p_export(self.h, buf))
        return buf.raw

    def p2p_attach(self, rank: int, world: int, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        buf = C.create_string_buffer(blob, len(blob))
        self._ck(self.L.flb_p2p_attach(self.h, rank, world, buf))

    def p2p_detach(self):
        self._ck(self.L.flb_p2p_detach(self.h))

    def comm_init(self, unique_id: bytes, rank: int, world: int):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.L.flb_comm_init(self.h, buf, rank, world))
