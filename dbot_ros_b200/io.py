"""Input adapters either side of the hot path: Wavefront .obj meshes as
dbot's SimpleWavefrontObjectModelLoader reads them for ObjectModel
(ref:source/dbot_ros/tracker/particle_tracker_node.cpp:84-97; mesh names come from
ref:config/object.yaml:3-6), and 32FC1 depth images as ri::to_eigen_vector hands them to the
tracker (ref:source/dbot_ros/util/ros_interface.h:152-168 -- CameraData.depth_image_vector)."""
import os

import numpy as np

from .rb_sensor import ObjectModel


def load_obj(path):
    """-> (vertices float64 [nv,3], triangles int32 [nt,3]).  Reads `v` and `f` records; faces may
    be `i`, `i/t`, `i//n` or `i/t/n`, 1-based or negative (relative) indices; polygons are fanned
    into triangles.  Everything else (vt, vn, groups, materials) is ignored."""
    verts, tris = [], []
    with open(path) as f:
        for ln, line in enumerate(f, 1):
            tok = line.split("#", 1)[0].split()
            if not tok:
                continue
            if tok[0] == "v":
                if len(tok) < 4:
                    raise ValueError("%s:%d: vertex needs 3 coordinates" % (path, ln))
                verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif tok[0] == "f":
                idx = []
                for t in tok[1:]:
                    k = int(t.split("/")[0])
                    k = k - 1 if k > 0 else len(verts) + k
                    if k < 0 or k >= len(verts):
                        raise ValueError("%s:%d: face index %s out of range" % (path, ln, t))
                    idx.append(k)
                if len(idx) < 3:
                    raise ValueError("%s:%d: face needs 3 vertices" % (path, ln))
                for j in range(1, len(idx) - 1):
                    tris.append([idx[0], idx[j], idx[j + 1]])
    return np.asarray(verts, dtype=np.float64).reshape(-1, 3), np.asarray(tris, dtype=np.int32).reshape(-1, 3)


def save_obj(path, vertices, triangles):
    with open(path, "w") as f:
        for v in np.asarray(vertices):
            f.write("v %.9g %.9g %.9g\n" % tuple(v))
        for t in np.asarray(triangles):
            f.write("f %d %d %d\n" % tuple(int(i) + 1 for i in t))


def object_model_from_files(directory, meshes, center_object_frame=True):
    """ObjectModel from the `object: {directory, meshes}` entries of object.yaml
    (ref:config/object.yaml:3-6), one part per mesh file."""
    V, T = zip(*[load_obj(os.path.join(directory, m)) for m in meshes])
    return ObjectModel(list(V), list(T), center_object_frame=center_object_frame)


# ---- the reference's offline dataset layout -------------------------------------
# TrackingDataset (ref:source/dbot_ros/util/tracking_dataset.cpp:92-101) is a directory holding
# `measurements.bag` (depth images + CameraInfo; needs ROS, which this image does not have) and
# `ground_truth.txt`.  The text file is plain: StoreTextFile (ref:...tracking_dataset.cpp:325-357) writes one
# line per frame that has a ground truth, "<image stamp> <state values...>", the state being the tracker's
# State vector (12 doubles per object: position, rotation vector, linear and angular velocity --
# FreeFloatingRigidBodiesState; 6 per object are accepted too).  LoadTextFile (ref:...:229-277) attaches a line
# to the frames whose stamp is within admissible_delta_time_ (ref:...tracking_dataset.h:134) of it.
GROUND_TRUTH_FILENAME = "ground_truth.txt"      # ref:source/dbot_ros/util/tracking_dataset.cpp:97
ADMISSIBLE_DELTA_TIME = 0.02                    # ref:source/dbot_ros/util/tracking_dataset.cpp:98


def load_ground_truth(path):
    """-> (stamps float64 [F], states float64 [F, k]); `path` is the file or the dataset directory"""
    if os.path.isdir(path):
        path = os.path.join(path, GROUND_TRUTH_FILENAME)
    stamps, states = [], []
    with open(path) as f:
        for ln, line in enumerate(f, 1):
            tok = line.split()
            if not tok:
                continue
            vals = [float(t) for t in tok]
            if len(vals) < 7 or (len(vals) - 1) % 6:
                raise ValueError("%s:%d: expected a stamp and 6 or 12 numbers per object, got %d values"
                                 % (path, ln, len(vals)))
            if states and len(vals) - 1 != len(states[0]):
                raise ValueError("%s:%d: state size changes from %d to %d" % (path, ln, len(states[0]), len(vals) - 1))
            stamps.append(vals[0])
            states.append(vals[1:])
    return np.asarray(stamps, dtype=np.float64), np.asarray(states, dtype=np.float64).reshape(len(stamps), -1)


def save_ground_truth(path, stamps, states):
    """the writer side of the same layout (TrackingDataset::StoreTextFile)"""
    if os.path.isdir(path):
        path = os.path.join(path, GROUND_TRUTH_FILENAME)
    states = np.asarray(states, dtype=np.float64).reshape(len(stamps), -1)
    with open(path, "w") as f:
        for t, s in zip(stamps, states):
            f.write("%.9f %s\n" % (float(t), " ".join("%.17g" % v for v in s)))


def ground_truth_poses(states, n_objects):
    """[F, k] state vectors -> [F, n_objects, 6] poses (position, rotation vector), whether the file holds
    12-number states or 6-number poses per object"""
    states = np.asarray(states, dtype=np.float64)
    per = states.shape[1] // n_objects
    if per not in (6, 12) or per * n_objects != states.shape[1]:
        raise ValueError("ground truth holds %d values per frame, not 6 or 12 per object for %d objects"
                         % (states.shape[1], n_objects))
    return states.reshape(len(states), n_objects, per)[:, :, :6].copy()


def match_ground_truth(frame_stamps, gt_stamps, admissible_delta_time=ADMISSIBLE_DELTA_TIME):
    """index of the ground-truth line for every frame, -1 where none lies within the admissible time difference
    (the matching rule of TrackingDataset::LoadTextFile)"""
    frame_stamps = np.asarray(frame_stamps, dtype=np.float64)
    gt_stamps = np.asarray(gt_stamps, dtype=np.float64)
    out = np.full(len(frame_stamps), -1, dtype=np.int64)
    if len(gt_stamps) == 0:
        return out
    order = np.argsort(gt_stamps)
    pos = np.searchsorted(gt_stamps[order], frame_stamps)
    for i, (t, p) in enumerate(zip(frame_stamps, pos)):
        best, best_d = -1, admissible_delta_time
        for q in (p - 1, p):
            if 0 <= q < len(order) and abs(gt_stamps[order[q]] - t) <= best_d:
                best, best_d = int(order[q]), abs(gt_stamps[order[q]] - t)
        out[i] = best
    return out


def load_dataset(directory):
    """A replayable dataset directory: `ground_truth.txt` in the reference's layout plus the depth frames the
    bag would hold, unpacked to `measurements.npz` (depth [F, rows, cols] float32 metres, NaN = no data;
    stamps [F]; K [3, 3] -- the CameraInfo matrix at native resolution).  -> dict"""
    z = np.load(os.path.join(directory, "measurements.npz"))
    stamps_gt, states = load_ground_truth(directory)
    return {"depth": z["depth"], "stamps": z["stamps"], "K": z["K"], "gt_stamps": stamps_gt, "gt_states": states,
            "gt_index": match_ground_truth(z["stamps"], stamps_gt)}


def save_dataset(directory, depth, stamps, K, states):
    os.makedirs(directory, exist_ok=True)
    np.savez_compressed(os.path.join(directory, "measurements.npz"), depth=np.asarray(depth, dtype=np.float32),
                        stamps=np.asarray(stamps, dtype=np.float64), K=np.asarray(K, dtype=np.float64))
    save_ground_truth(directory, stamps, states)
