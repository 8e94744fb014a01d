"""Host-side mirror of the reference's sensor interface for the depth observation model.

Names, argument meaning and error behaviour follow what dbot_ros builds in
ref:source/dbot_ros/tracker/particle_tracker_node.cpp:84-203 (and the copy in
ref:source/dbot_ros/tracker/object_tracker_service_node.cpp:136-175):

    object_model   = ObjectModel(loader, center_object_frame)            (:92-97)
    camera_data    = CameraData(provider)                                 (:102-121)
    params_obsrv   = RbSensorBuilder.Parameters  (use_gpu, sample_count, occlusion.*, kinect.*,
                     delta_time, shader keys)                             (:164-199)
    sensor_builder = RbSensorBuilder(object_model, camera_data, params)  (:201-203)
    sensor         = sensor_builder.build()          -> RbSensor
    sensor.reset() / sensor.set_observation(image) / sensor.loglikes(deltas, indices, update)

All arithmetic happens in libdbot_b200.so (C-ABI, include/dbot_b200.h); this module only does
what dbot's KinectImageModelGPU does on the host before its GPU call: compose each particle's
delta pose with the integrated (default) pose and hand R|t per object to the device.
"""
from __future__ import annotations

import atexit
import ctypes as C
import weakref
from dataclasses import dataclass, field

import numpy as np

from . import abi

# dbot KinectPixelModel constructor defaults that particle_tracker.yaml does not expose
HALF_LIFE_DEPTH = 1.0
MAX_DEPTH = 6.0


@dataclass
class OcclusionParameters:
    """ref:config/particle_tracker.yaml:41-43"""

    p_occluded_visible: float = 0.1
    p_occluded_occluded: float = 0.7
    initial_occlusion_prob: float = 0.1


@dataclass
class KinectParameters:
    """ref:config/particle_tracker.yaml:47-49"""

    tail_weight: float = 0.01
    model_sigma: float = 0.003
    sigma_factor: float = 0.0014247


@dataclass
class Parameters:
    """dbot::RbSensorBuilder<State>::Parameters as filled at
    ref:source/dbot_ros/tracker/particle_tracker_node.cpp:164-199."""

    use_gpu: bool = True
    sample_count: int = 2000
    delta_time: float = 1.0 / 30.0
    occlusion: OcclusionParameters = field(default_factory=OcclusionParameters)
    kinect: KinectParameters = field(default_factory=KinectParameters)
    # OpenGL-only keys of the reference's GPU evaluator: accepted, ignored (there is no OpenGL)
    use_custom_shaders: bool = False
    vertex_shader_file: str = ""
    fragment_shader_file: str = ""
    geometry_shader_file: str = "none"
    # B200 extension (optional YAML key gpu/cull_backfaces, default true): skip triangles facing away
    # from the camera for meshes proven closed; false = draw every triangle as dbot's renderer does
    cull_backfaces: bool = True

    @classmethod
    def from_dict(cls, tree):
        """`tree` is the content of particle_tracker.yaml (with or without the top-level
        `particle_filter` key).  Key lookup mirrors the node: use_gpu picks gpu/ or cpu/
        sample_count (ref:...particle_tracker_node.cpp:165-174); delta_time is fixed to 1/30 (:189)."""
        pf = tree.get("particle_filter", tree)
        p = cls()
        p.use_gpu = bool(pf["use_gpu"])
        p.sample_count = int(pf["gpu" if p.use_gpu else "cpu"]["sample_count"])
        occ = pf["observation"]["occlusion"]
        kin = pf["observation"]["kinect"]
        p.occlusion = OcclusionParameters(
            float(occ["p_occluded_visible"]), float(occ["p_occluded_occluded"]),
            float(occ["initial_occlusion_prob"]))
        p.kinect = KinectParameters(
            float(kin["tail_weight"]), float(kin["model_sigma"]), float(kin["sigma_factor"]))
        p.delta_time = 1.0 / 30.0
        gpu = pf.get("gpu", {})
        p.use_custom_shaders = bool(gpu.get("use_custom_shaders", False))
        p.vertex_shader_file = str(gpu.get("vertex_shader_file", ""))
        p.fragment_shader_file = str(gpu.get("fragment_shader_file", ""))
        p.geometry_shader_file = str(gpu.get("geometry_shader_file", "none"))
        p.cull_backfaces = bool(gpu.get("cull_backfaces", True))
        return p

    @classmethod
    def from_yaml(cls, path):
        import yaml

        with open(path) as f:
            return cls.from_dict(yaml.safe_load(f))


class ObjectModel:
    """dbot::ObjectModel: per-part vertex and triangle lists; with center_object_frame the
    vertices of every part are shifted so that their mean is the origin
    (ref:config/particle_tracker.yaml:27-30, ref:...particle_tracker_node.cpp:92-97)."""

    def __init__(self, vertices, triangles, center_object_frame=False):
        if len(vertices) != len(triangles) or len(vertices) == 0:
            raise ValueError("ObjectModel: need one vertex list and one triangle list per part")
        self.centers = []
        self.vertices = []
        self.triangles = []
        for v, t in zip(vertices, triangles):
            v = np.asarray(v, dtype=np.float64).reshape(-1, 3)
            t = np.ascontiguousarray(t, dtype=np.int32).reshape(-1, 3)
            c = v.mean(axis=0) if (center_object_frame and len(v)) else np.zeros(3)
            self.centers.append(c)
            self.vertices.append(np.ascontiguousarray(v - c, dtype=np.float32))
            self.triangles.append(t)

    @property
    def count_parts(self):
        return len(self.vertices)


class CameraData:
    """dbot::CameraData: camera matrix and resolution AFTER downsampling.  `from_native` applies
    what RosCameraDataProvider does: rows 0-1 of K divided by the factor
    (ref:source/dbot_ros/util/ros_camera_data_provider.cpp:72) and resolution = native / factor
    (ref:config/camera.yaml:5-9)."""

    def __init__(self, camera_matrix, rows, cols, downsampling_factor=1):
        self.camera_matrix = np.asarray(camera_matrix, dtype=np.float64).reshape(3, 3)
        self.rows, self.cols = int(rows), int(cols)
        self.downsampling_factor = int(downsampling_factor)

    @classmethod
    def from_native(cls, camera_matrix, width, height, downsampling_factor):
        K = np.array(camera_matrix, dtype=np.float64).reshape(3, 3)
        K[:2, :] /= float(downsampling_factor)
        return cls(K, height // downsampling_factor, width // downsampling_factor,
                   downsampling_factor)

    def depth_image_vector(self, image):
        """ri::to_eigen_vector: row-major vector with stride-N decimation
        (ref:source/dbot_ros/util/ros_interface.h:152-168)."""
        img = np.asarray(image, dtype=np.float32)
        ds = self.downsampling_factor
        return np.ascontiguousarray(img[: self.rows * ds : ds, : self.cols * ds : ds]).reshape(-1)

    @property
    def pixels(self):
        return self.rows * self.cols


def rotation_matrices(euler_vectors):
    """dbot EulerVector::rotation_matrix for an array [..., 3] of rotation vectors (float64)."""
    v = np.asarray(euler_vectors, dtype=np.float64)
    angle = np.linalg.norm(v, axis=-1)
    safe = np.where(angle > 0.0, angle, 1.0)
    axis = v / safe[..., None]
    x, y, z = axis[..., 0], axis[..., 1], axis[..., 2]
    c, s = np.cos(angle), np.sin(angle)
    C1 = 1.0 - c
    R = np.empty(v.shape[:-1] + (3, 3))
    R[..., 0, 0] = c + x * x * C1
    R[..., 0, 1] = x * y * C1 - z * s
    R[..., 0, 2] = x * z * C1 + y * s
    R[..., 1, 0] = y * x * C1 + z * s
    R[..., 1, 1] = c + y * y * C1
    R[..., 1, 2] = y * z * C1 - x * s
    R[..., 2, 0] = z * x * C1 - y * s
    R[..., 2, 1] = z * y * C1 + x * s
    R[..., 2, 2] = c + z * z * C1
    ident = np.broadcast_to(np.eye(3), R.shape)
    return np.where((angle > 0.0)[..., None, None], R, ident)


def compose_poses(deltas, default_poses):
    """pose = delta (x) default for every particle and object: t = t_delta + t_default,
    R = R_delta R_default, in double, then rounded to the float32 [n][n_objects][12] (R row-major, t)
    the C-ABI takes.  deltas: [n, n_objects*12] FreeFloatingRigidBodiesState or [n, n_objects*6]."""
    default_poses = np.asarray(default_poses, dtype=np.float64).reshape(-1, 6)
    n_obj = default_poses.shape[0]
    d = np.asarray(deltas, dtype=np.float64)
    n = d.shape[0]
    d = d.reshape(n, n_obj, -1)
    if d.shape[2] not in (6, 12):
        raise ValueError("loglikes: states must have 6 or 12 numbers per object")
    R = rotation_matrices(d[:, :, 3:6]) @ rotation_matrices(default_poses[:, 3:6])[None]
    t = d[:, :, 0:3] + default_poses[None, :, 0:3]
    out = np.empty((n, n_obj, 12), dtype=np.float32)
    out[:, :, :9] = R.reshape(n, n_obj, 9)
    out[:, :, 9:] = t
    return out


class RbSensor:
    """dbot::RbSensor<State> (model/rb_sensor.h): the interface the particle filter drives."""

    def reset(self):
        raise NotImplementedError

    def set_observation(self, image):
        raise NotImplementedError

    def loglikes(self, deltas, indices, update=False):
        raise NotImplementedError


# Handles still open when the interpreter exits are closed from a Python atexit hook, i.e. BEFORE the CUDA
# runtime's own exit handler runs: a __del__ that reaches cudaFree / cudaStreamDestroy after the runtime has
# torn itself down (module-level sensors are finalised that late) crashes the process on its way out.
_LIVE_SENSORS = weakref.WeakSet()
_LIVE_FILTERS = weakref.WeakSet()     # device-resident filters hold their sensor's stream: closed first
_LIVE_WRAPPERS = weakref.WeakSet()    # sharded wrappers hold torch buffers tied to a sensor's stream: closed before it


def _close_all_at_exit():
    for group in (_LIVE_WRAPPERS, _LIVE_FILTERS, _LIVE_SENSORS):
        for obj in list(group):
            try:
                obj.close()
            except Exception:
                pass


atexit.register(_close_all_at_exit)


class B200KinectImageModel(RbSensor):
    """The sensor RbSensorBuilder.build() returns when use_gpu is true: stands where dbot's
    KinectImageModelGPU stood, backed by libdbot_b200.so."""

    def __init__(self, camera_matrix, rows, cols, max_sample_count, object_model, params,
                 device_id=0, occlusion_mode=0, device_ids=None, gather_mode=0):
        self._lib = abi.load_library()
        self._h = C.c_void_p()
        self.rows, self.cols = int(rows), int(cols)
        self.n_objects = object_model.count_parts
        self.max_sample_count = int(max_sample_count)
        self.params = params
        cfg, self._keep = abi.make_config(
            rows, cols, camera_matrix, max_sample_count, object_model.vertices,
            object_model.triangles,
            tail_weight=params.kinect.tail_weight, model_sigma=params.kinect.model_sigma,
            sigma_factor=params.kinect.sigma_factor, half_life_depth=HALF_LIFE_DEPTH,
            max_depth=MAX_DEPTH, p_occluded_visible=params.occlusion.p_occluded_visible,
            p_occluded_occluded=params.occlusion.p_occluded_occluded,
            initial_occlusion_prob=params.occlusion.initial_occlusion_prob,
            delta_time=params.delta_time, device_id=device_id, occlusion_mode=occlusion_mode,
            cull_backfaces=0 if getattr(params, "cull_backfaces", True) else 1,
            device_ids=device_ids, gather_mode=gather_mode)
        abi.check(self._lib, self._lib.dbot_b200_create(C.byref(cfg), C.byref(self._h)))
        _LIVE_SENSORS.add(self)
        # RbSensor::default_poses_ -- one PoseVector (position, euler vector) per object
        self._default_poses = np.zeros((self.n_objects, 6))

    # -- RbSensor ---------------------------------------------------------------------------
    def reset(self):
        abi.check(self._lib, self._lib.dbot_b200_reset(self._h))

    def set_observation(self, image):
        img = np.ascontiguousarray(image, dtype=np.float32).reshape(-1)
        if img.size != self.rows * self.cols:
            raise ValueError("set_observation: expected %d pixels, got %d" % (self.rows * self.cols, img.size))
        abi.check(self._lib, self._lib.dbot_b200_set_observation(self._h, abi.ptr(img)))

    def loglikes(self, deltas, indices, update=False):
        """deltas: [n, n_objects*12] FreeFloatingRigidBodiesState (position, euler vector, linear
        and angular velocity per body) or [n, n_objects*6] pose vectors; indices: int32 [n],
        occlusion slot of each particle's parent, overwritten with 0..n-1 when update."""
        poses = self.compose(deltas)
        return self.loglikes_poses(poses, indices, update)

    # -- pieces ------------------------------------------------------------------------------
    @property
    def integrated_poses(self):
        return self._default_poses.copy()

    @integrated_poses.setter
    def integrated_poses(self, poses):
        self._default_poses = np.asarray(poses, dtype=np.float64).reshape(self.n_objects, 6).copy()

    def compose(self, deltas):
        return compose_poses(deltas, self._default_poses)

    def loglikes_poses(self, poses, indices, update=False):
        poses = np.ascontiguousarray(poses, dtype=np.float32).reshape(-1, self.n_objects, 12)
        n = poses.shape[0]
        if not (isinstance(indices, np.ndarray) and indices.dtype == np.int32 and indices.flags.c_contiguous):
            raise TypeError("loglikes: indices must be a contiguous int32 numpy array (it is written)")
        if indices.size != n:
            raise ValueError("loglikes: %d states but %d indices" % (n, indices.size))
        out = np.empty(n, dtype=np.float64)
        abi.check(self._lib, self._lib.dbot_b200_loglikes(
            self._h, abi.ptr(poses), abi.ptr(indices), n, 1 if update else 0, abi.ptr(out)))
        return out

    def render(self, pose):
        pose = np.ascontiguousarray(pose, dtype=np.float32).reshape(self.n_objects, 12)
        out = np.empty(self.rows * self.cols, dtype=np.float32)
        abi.check(self._lib, self._lib.dbot_b200_render(self._h, abi.ptr(pose), abi.ptr(out)))
        return out.reshape(self.rows, self.cols)

    # -- device-resident variants (torch tensors or raw device pointers) ----------------------
    def set_observation_device(self, d_depth_ptr):
        abi.check(self._lib, self._lib.dbot_b200_set_observation_device(self._h, C.c_void_p(d_depth_ptr)))

    def loglikes_device(self, d_poses_ptr, d_indices_ptr, n, update, d_out_ptr):
        abi.check(self._lib, self._lib.dbot_b200_loglikes_device(
            self._h, C.c_void_p(d_poses_ptr), C.c_void_p(d_indices_ptr), int(n), 1 if update else 0,
            C.c_void_p(d_out_ptr)))

    def synchronize(self):
        abi.check(self._lib, self._lib.dbot_b200_synchronize(self._h))

    # -- occlusion-slot migration (particle-sharded filter, cross-GPU parents) -------------------
    @property
    def slot_floats(self):
        return int(self._lib.dbot_b200_slot_floats(self._h))

    def export_slot_device(self, slot, d_dst_ptr):
        abi.check(self._lib, self._lib.dbot_b200_export_slot_device(self._h, int(slot), C.c_void_p(d_dst_ptr)))

    def import_slot_device(self, d_src_ptr):
        out = C.c_int32(-1)
        abi.check(self._lib, self._lib.dbot_b200_import_slot_device(self._h, C.c_void_p(d_src_ptr), C.byref(out)))
        return out.value

    # -- gather of log-weights over NVLink peer memory, one process per GPU (csrc/peer_gather.cuh) ---
    def peer_gather_export(self, world, rank, width_max):
        """allocate this rank's gathered vectors; returns the 64-byte CUDA IPC handle the other ranks need"""
        handle = C.create_string_buffer(abi.IPC_HANDLE_BYTES)
        abi.check(self._lib, self._lib.dbot_b200_peer_gather_export(
            self._h, int(world), int(rank), int(width_max), C.cast(handle, C.c_void_p)))
        return handle.raw

    def peer_gather_connect(self, handles):
        """handles: the world exported handles in rank order"""
        blob = b"".join(handles)
        buf = C.create_string_buffer(blob, len(blob))
        abi.check(self._lib, self._lib.dbot_b200_peer_gather_connect(self._h, C.cast(buf, C.c_void_p)))

    def loglikes_gather(self, poses_local, indices_local, update, width, out_all):
        """this rank's slice in, every rank's log-weights out (out_all: float64[world * width], written)"""
        n = int(indices_local.shape[0])
        abi.check(self._lib, self._lib.dbot_b200_loglikes_gather(
            self._h, abi.ptr(poses_local) if n else None, abi.ptr(indices_local) if n else None, n,
            1 if update else 0, int(width), abi.ptr(out_all)))

    def loglikes_gather_device(self, d_poses_ptr, d_indices_ptr, n_local, update, d_out_ptr, width):
        abi.check(self._lib, self._lib.dbot_b200_loglikes_gather_device(
            self._h, C.c_void_p(d_poses_ptr), C.c_void_p(d_indices_ptr), int(n_local), 1 if update else 0,
            C.c_void_p(d_out_ptr), int(width)))

    def set_staging_threads(self, n):
        """host threads that help to stage pageable depth frames (default 2; 0 = the caller's thread alone)"""
        abi.check(self._lib, self._lib.dbot_b200_set_option(self._h, abi.OPT_STAGING_THREADS, int(n)))

    @property
    def peer_gather_result(self):
        """device pointer of the gathered vector the last gathering call filled"""
        return self._lib.dbot_b200_peer_gather_result(self._h)

    @property
    def stream(self):
        return self._lib.dbot_b200_stream(self._h)

    # -- state read-back and measurement -------------------------------------------------------
    @property
    def n_slots(self):
        return self._lib.dbot_b200_n_slots(self._h)

    @property
    def device_count(self):
        """GPUs behind this sensor (particles are sharded over them inside the library)"""
        return self._lib.dbot_b200_device_count(self._h)

    @property
    def migrated_slots(self):
        return int(self._lib.dbot_b200_migrated_slots(self._h))

    def update_empty(self):
        abi.check(self._lib, self._lib.dbot_b200_update_empty(self._h))

    def cull_sign(self, obj):
        """+1 / -1: back faces of this object are skipped (mesh proven closed); 0: all triangles drawn"""
        return self._lib.dbot_b200_cull_sign(self._h, int(obj))

    def slot_rows(self, slot):
        """(first, last) image row in which the slot's stored image may differ from the background; first > last
        when it is the background everywhere"""
        a, b = C.c_int32(), C.c_int32()
        abi.check(self._lib, self._lib.dbot_b200_slot_rows(self._h, int(slot), C.byref(a), C.byref(b)))
        return a.value, b.value

    def slot_rect(self, slot):
        """(first row, last row, first column, last column) of the rectangle in which the slot's stored image may
        differ from the background (columns in steps of 8 pixels); only this part of a parent is copied"""
        r = (C.c_int32 * 4)()
        abi.check(self._lib, self._lib.dbot_b200_slot_rect(self._h, int(slot), r))
        return tuple(int(v) for v in r)

    @property
    def free_slots(self):
        return self._lib.dbot_b200_free_slot_count(self._h)

    @property
    def observation_time(self):
        return self._lib.dbot_b200_observation_time(self._h)

    def occlusion(self, slot):
        out = np.empty(self.rows * self.cols, dtype=np.float32)
        abi.check(self._lib, self._lib.dbot_b200_get_occlusion(self._h, int(slot), abi.ptr(out)))
        return out.reshape(self.rows, self.cols)

    def timer_begin(self):
        abi.check(self._lib, self._lib.dbot_b200_timer_begin(self._h))

    def timer_end(self):
        ms = C.c_float()
        abi.check(self._lib, self._lib.dbot_b200_timer_end(self._h, C.byref(ms)))
        return ms.value

    def profile_enable(self, on=True):
        abi.check(self._lib, self._lib.dbot_b200_profile_enable(self._h, 1 if on else 0))

    def profile_read(self):
        ms = np.zeros(len(abi.KERNEL_NAMES), dtype=np.float32)
        cnt = np.zeros(len(abi.KERNEL_NAMES), dtype=np.int64)
        abi.check(self._lib, self._lib.dbot_b200_profile_read(self._h, abi.ptr(ms), abi.ptr(cnt)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(abi.KERNEL_NAMES)}

    @property
    def launch_count(self):
        return self._lib.dbot_b200_launch_count(self._h)

    @property
    def device_bytes(self):
        return self._lib.dbot_b200_device_bytes(self._h)

    def close(self):
        if self._h:
            self._lib.dbot_b200_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RbSensorBuilder:
    """dbot::RbSensorBuilder<State>(object_model, camera_data, params)
    (ref:source/dbot_ros/tracker/particle_tracker_node.cpp:201-203).  build() with use_gpu
    returns the B200 sensor; with use_gpu false it raises, as dbot's builder does for the GPU
    branch when it was compiled without DBOT_BUILD_GPU -- this package has no CPU model."""

    Parameters = Parameters

    def __init__(self, object_model, camera_data, params, device_id=0, device_ids=None, gather_mode=0):
        """device_ids: several CUDA ordinals -> one sensor whose particles are sharded over those GPUs inside
        the library (one process, BASELINE config 4); gather_mode 0 = ncclAllGather, 1 = fused peer stores"""
        self.object_model = object_model
        self.camera_data = camera_data
        self.params = params
        self.device_id = device_id
        self.device_ids = device_ids
        self.gather_mode = gather_mode

    def build(self):
        if not self.params.use_gpu:
            raise RuntimeError(
                "RbSensorBuilder: use_gpu is false and dbot_ros_b200 ships no CPU sensor "
                "(the CPU model stays dbot's KinectImageModel)")
        return B200KinectImageModel(
            self.camera_data.camera_matrix, self.camera_data.rows, self.camera_data.cols,
            self.params.sample_count, self.object_model, self.params, device_id=self.device_id,
            device_ids=self.device_ids, gather_mode=self.gather_mode)
