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
