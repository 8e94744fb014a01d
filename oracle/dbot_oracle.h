/*
 * oracle/dbot_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the depth observation model that dbot_ros selects through
 * dbot::RbSensorBuilder<State> (ref:source/dbot_ros/tracker/particle_tracker_node.cpp:164-203)
 * and drives through tracker_->track(image) (ref:source/dbot_ros/object_tracker_ros.hpp:49).
 *
 * The arithmetic itself lives in the un-vendored, un-pinned dependency
 * `bayesian-object-tracking/dbot` (ref:package.xml:28, ref:CMakeLists.txt:44,
 * ref:README.md:26), which is NOT under /root/reference.  This file restates its
 * published algorithm (dbot model/kinect_image_model.h, model/kinect_pixel_model.h,
 * model/occlusion_model.h, rigid_body_renderer.cpp) as written down in
 * SURVEY.md Appendix A.  The reference tree holds no tests, golden vectors or
 * fixtures for this path, so:
 *
 *     *** PARITY UNPINNED ***  -- this oracle is pinned only by closed-form
 *     known-answer tests (tests/test_oracle_known_answers.py) and by an
 *     independent numpy re-derivation (tests/np_model.py); it cannot be checked
 *     against outputs of the real dbot library in this environment.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.
 */
#ifndef DBOT_ORACLE_H
#define DBOT_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_sensor oracle_sensor;

typedef struct
{
    int32_t rows, cols;  /* after downsampling (ref:source/dbot_ros/util/ros_interface.h:155-156) */
    double K[9];         /* row-major; rows 0-1 already / downsampling
                            (ref:source/dbot_ros/util/ros_camera_data_provider.cpp:72) */
    int32_t n_objects;
    const float* const* vertices;    /* [n_objects] -> float[3*n_vertices[o]] */
    const int32_t* n_vertices;
    const int32_t* const* triangles; /* [n_objects] -> int32[3*n_triangles[o]] */
    const int32_t* n_triangles;
    /* ref:config/particle_tracker.yaml:47-49 */
    double tail_weight, model_sigma, sigma_factor;
    /* dbot KinectPixelModel constructor defaults (not exposed in YAML) */
    double half_life_depth, max_depth;
    /* ref:config/particle_tracker.yaml:41-43 */
    double p_occluded_visible, p_occluded_occluded, initial_occlusion_prob;
    /* ref:source/dbot_ros/tracker/particle_tracker_node.cpp:189 */
    double delta_time;
    /* 0: lazy, per-pixel time stamps (dbot CPU KinectImageModel);
       1: eager, whole image propagated on every update (dbot GPU evaluator). */
    int32_t eager;
    /* 1: round p_vis*(1-o), p_occ*o, p_inf and o to float before the log, as
       the dbot CPU model's `float` locals do (soft detail). */
    int32_t float_intermediates;
    /* 0: every triangle is rasterised, as upstream (dbot's renderer does not cull).
       1: triangles facing away from the camera are skipped for every object whose mesh is
          proven closed and consistently oriented (oracle_cull_sign != 0) -- the rule the CUDA
          library applies by default.  For such a mesh the nearest surface along any ray from
          outside is camera-facing, so the rendered depth is the same except where rounding
          decides between the two triangles that meet at a silhouette edge
          (tests/test_oracle_known_answers.py counts those pixels). */
    int32_t cull_backfaces;
} oracle_config;

int oracle_create(const oracle_config* cfg, oracle_sensor** out);
int oracle_destroy(oracle_sensor* s);
int oracle_reset(oracle_sensor* s);
int oracle_set_observation(oracle_sensor* s, const float* depth_m);
/* poses: [n][n_objects][12] = R row-major (9) + t (3), float32, camera frame */
int oracle_loglikes(oracle_sensor* s, const float* poses, int32_t* indices,
                    int32_t n, int32_t update, double* out_loglikes);
/* one pose set -> full-frame depth, +INFINITY where nothing is hit */
int oracle_render(oracle_sensor* s, const float* pose, float* out_depth);
/* state read-back for parity tests */
int oracle_n_slots(const oracle_sensor* s);
double oracle_observation_time(const oracle_sensor* s);
/* occlusion probability of slot `slot` as of the current observation time
   (i.e. propagated to t_obs), one float per pixel */
int oracle_get_occlusion(const oracle_sensor* s, int32_t slot, float* out_occ);
/* raw stored state (occ + time stamp) */
int oracle_get_occlusion_raw(const oracle_sensor* s, int32_t slot,
                             float* out_occ, double* out_time);
void oracle_set_num_threads(int n);
int oracle_get_max_threads(void);

/* scalar pieces, exported for known-answer tests */
double oracle_p_visible(const oracle_sensor* s, double y, double d);
double oracle_p_occluded(const oracle_sensor* s, double y, double d);
double oracle_p_occluded_inf(const oracle_sensor* s, double y);
double oracle_propagate(const oracle_sensor* s, double o, double dt);

/* sign of the snapped screen-space area (x right, y down) of the camera-facing triangles of
   object o: +1 / -1 when its mesh is closed (every directed edge of the position-welded mesh
   is matched by exactly one opposite edge) with every connected component oriented the same
   way, 0 otherwise.  Independent of cull_backfaces. */
int oracle_cull_sign(const oracle_sensor* s, int32_t o);

/* test support: append a slot exported from another instance (see dbot_oracle.c) */
int oracle_import_slot(oracle_sensor* s, const float* occ, const double* occ_time);

#ifdef __cplusplus
}
#endif
#endif
