/*
 * include/dbot_b200.h -- C-ABI of libdbot_b200.so, the B200 (sm_100a) depth
 * observation model for the dbot / dbot_ros particle tracker.
 *
 * Drop-in boundary.  In the reference the sensor is
 * obtained at
 *     ref:source/dbot_ros/tracker/particle_tracker_node.cpp:164-203
 *     ref:source/dbot_ros/tracker/object_tracker_service_node.cpp:136-175
 * (`dbot::RbSensorBuilder<State>(object_model, camera_data, params)`), where
 * `params.use_gpu` (ref:config/particle_tracker.yaml:3) selects dbot's
 * OpenGL+CUDA `KinectImageModelGPU`.  This library replaces that evaluator:
 * each entry point below names the `dbot::RbSensor<State>` member (dbot
 * model/rb_sensor.h, un-vendored) it stands in for.  INTEGRATION.md shows the
 * C++ adapter a dbot maintainer would add.
 *
 * Conventions: plain C types only; caller owns every host buffer; the library
 * owns all device memory and its CUDA stream; every call returns 0 on
 * success, <0 for argument errors (-12: entry point not available on a multi-GPU handle, -13: reset() required
 * after an update that failed on the device, -14: NCCL), >0 for a CUDA error code; text through
 * dbot_b200_last_error() (thread-local).  One caller at a time per handle
 * (the tracker serialises track() -- ref:source/dbot_ros/object_tracker_ros.hpp:111-125).
 * There is no CPU fallback: without a CUDA device create() fails.
 */
#ifndef DBOT_B200_H
#define DBOT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DBOT_B200_ABI_VERSION 2
#define DBOT_B200_MAX_OBJECTS 8
#define DBOT_B200_MAX_DEVICES 8

typedef struct dbot_b200_sensor dbot_b200_sensor; /* opaque */

/* Mirrors dbot::RbSensorBuilder<State>::Parameters + CameraData + ObjectModel, i.e.
 * everything ref:source/dbot_ros/tracker/particle_tracker_node.cpp:84-203 hands to the builder. */
typedef struct
{
    int32_t rows, cols;   /* resolution after downsampling
                             (ref:source/dbot_ros/util/ros_interface.h:155-156) */
    double K[9];          /* camera matrix, row-major, rows 0-1 already divided by the
                             downsampling factor (ref:source/dbot_ros/util/ros_camera_data_provider.cpp:72);
                             K[3]=K[6]=K[7]=0 and K[8]=1 required */
    int32_t max_particles; /* = sample_count (ref:config/particle_tracker.yaml:7,11) */
    int32_t n_objects;     /* <= DBOT_B200_MAX_OBJECTS */
    const float* const* vertices;    /* [n_objects] -> float[3*n_vertices[o]], object frame
                                        (centred when center_object_frame,
                                        ref:source/dbot_ros/tracker/particle_tracker_node.cpp:92-97) */
    const int32_t* n_vertices;
    const int32_t* const* triangles; /* [n_objects] -> int32[3*n_triangles[o]] */
    const int32_t* n_triangles;
    double tail_weight, model_sigma, sigma_factor; /* ref:config/particle_tracker.yaml:47-49 */
    double half_life_depth, max_depth;             /* dbot KinectPixelModel defaults: 1.0, 6.0 */
    double p_occluded_visible, p_occluded_occluded,
        initial_occlusion_prob;                    /* ref:config/particle_tracker.yaml:41-43 */
    double delta_time;   /* ref:source/dbot_ros/tracker/particle_tracker_node.cpp:189 (1/30) */
    int32_t device_id;   /* CUDA device ordinal of a single-device handle (n_devices <= 1 and no device_ids) */
    int32_t occlusion_mode; /* 0 = dense: one full-frame fp32 occlusion image per particle with the
                               semantics of dbot's GPU evaluator (whole image propagated on every
                               update; the propagation is kept as two numbers, and an update moves only the
                               image rows an ancestor has ever written).  Other values are reserved. */
    int32_t cull_backfaces; /* 0 = triangles facing away from the camera are skipped for every object whose
                               mesh is proven closed and consistently oriented (dbot_b200_cull_sign != 0;
                               the nearest surface of such a mesh is always camera-facing, so the rendered
                               depth is unchanged); 1 = never: every triangle is drawn, as dbot's
                               RigidBodyRenderer does.  Open or inconsistently oriented meshes are always
                               drawn completely. */
    int32_t n_devices;      /* 0 or 1: one GPU (device_id, or device_ids[0] when given).  2..8: the particles of
                               every loglikes call -- and their occlusion slots -- are sharded over that many GPUs
                               of this process; frame, mesh and camera are replicated; one all-gather of the
                               log-weights per call over NVLink (BASELINE config 4).
                               Results are bit-identical to a single-GPU handle. */
    const int32_t* device_ids; /* [n_devices] CUDA ordinals, NULL = 0 .. n_devices-1 (an ordinal may repeat only
                                  with gather_mode 1: shards then share that GPU) */
    int32_t gather_mode;    /* n_devices > 1: 0 = ncclAllGather (libnccl.so.2 bound at run time), 1 = fused: the
                               particle kernel stores each log-weight into every GPU's gathered vector through
                               NVLink peer memory, no collective launch */
} dbot_b200_config;

/* RbSensorBuilder<State>::build() with use_gpu == true */
int dbot_b200_create(const dbot_b200_config* cfg, dbot_b200_sensor** out);
int dbot_b200_destroy(dbot_b200_sensor* s);

/* RbSensor<State>::reset(): one occlusion slot at initial_occlusion_prob, observation time 0 */
int dbot_b200_reset(dbot_b200_sensor* s);

/* RbSensor<State>::set_observation(image): rows*cols float32 metres, row-major
 * (row*cols+col, ref:source/dbot_ros/util/ros_interface.h:152-168), NaN = no data.
 * Advances the observation time by delta_time. */
int dbot_b200_set_observation(dbot_b200_sensor* s, const float* depth_m);

/* RbSensor<State>::loglikes(deltas, indices, update).
 *   poses        [n][n_objects][12] float32: R row-major (9) then t (3), absolute pose in the
 *                camera frame (delta (x) default pose already composed by the adapter)
 *   indices      [n] in: occlusion slot of each particle's parent; out: identity when update
 *   out_loglikes [n] double
 * update == 0 leaves all state untouched. */
int dbot_b200_loglikes(dbot_b200_sensor* s, const float* poses, int32_t* indices, int32_t n,
                       int32_t update, double* out_loglikes);

/* dbot RigidBodyRenderer::Render for one pose set: out_depth[rows*cols], +INFINITY where no
 * triangle is hit.  Used for visualisation and by the parity tests. */
int dbot_b200_render(dbot_b200_sensor* s, const float* pose, float* out_depth);

/* ---- device-resident variants: all pointers are device pointers on cfg.device_id; work is
 * enqueued on the handle's stream and NOT synchronised (use dbot_b200_synchronize). ---- */
int dbot_b200_set_observation_device(dbot_b200_sensor* s, const float* d_depth_m);
int dbot_b200_loglikes_device(dbot_b200_sensor* s, const float* d_poses, int32_t* d_indices,
                              int32_t n, int32_t update, double* d_out_loglikes);
void* dbot_b200_stream(dbot_b200_sensor* s); /* cudaStream_t */
int dbot_b200_synchronize(dbot_b200_sensor* s);

/* ---- occlusion-slot migration between handles (particle-sharded filter: a child whose parent lives
 * on another GPU).  Both handles must be at the same observation / update
 * time (they are when every rank runs the same call sequence).  d_dst / d_src: device buffers of
 * dbot_b200_slot_floats() floats; work is enqueued on the handle's stream. ---- */
int64_t dbot_b200_slot_floats(dbot_b200_sensor* s);
int dbot_b200_export_slot_device(dbot_b200_sensor* s, int32_t slot, float* d_dst);
/* appends a logical slot (index returned in *out_slot) that the next loglikes calls may name as a
 * parent; like every slot it is dropped by the next update unless a child adopts it */
int dbot_b200_import_slot_device(dbot_b200_sensor* s, const float* d_src, int32_t* out_slot);

/* ---- device-resident filter step around the sensor: dbot's
 * RaoBlackwellCoordinateParticleFilter::filter + ParticleTracker::on_track, i.e. what
 * tracker->initialize(poses) / tracker->track(image) run per frame
 * (ref:source/dbot_ros/tracker/particle_tracker_node.cpp:252, ref:source/dbot_ros/object_tracker_ros.hpp:49),
 * with every O(N) step as a kernel on the sensor's stream: states, poses, parent indices and weights
 * stay in HBM; per frame the depth image goes in and n_objects * 6 doubles come out. ---- */
typedef struct dbot_b200_filter dbot_b200_filter; /* opaque */
typedef struct
{
    double linear_sigma[3], angular_sigma[3]; /* ref:config/particle_tracker.yaml:55-60 */
    double velocity_factor;                   /* ref:config/particle_tracker.yaml:63 */
    double max_kl_divergence;                 /* ref:config/particle_tracker.yaml:36 */
    uint64_t seed;                            /* counter-based noise: same seed, same track */
    int32_t n_particles;                      /* evaluation_count / n_objects, <= max_particles */
    double moving_average_update_rate;        /* ref:config/particle_tracker.yaml:25, in (0, 1]; 1 = no smoothing */
    double centers[DBOT_B200_MAX_OBJECTS][3]; /* center_object_frame (ref:config/particle_tracker.yaml:30,
                                                 ref:...particle_tracker_node.cpp:210-214): the offset by which
                                                 every part's vertices were shifted (ObjectModel centers); poses
                                                 cross this API in the MODEL frame and are converted on the way
                                                 in and out.  Zeros when the meshes were not centred. */
} dbot_b200_filter_config;
int dbot_b200_filter_create(dbot_b200_sensor* s, const dbot_b200_filter_config* cfg, dbot_b200_filter** out);
int dbot_b200_filter_destroy(dbot_b200_filter* f);
/* poses: [n_objects][6] position + rotation vector of the model frame in the camera frame; resets the sensor */
int dbot_b200_filter_initialize(dbot_b200_filter* f, const double* poses);
/* out_states: [n_objects][12] the state the reference publishes after this frame
 * (ref:source/dbot_ros/object_tracker_ros.hpp:49-61): position, rotation vector, linear and angular velocity,
 * moving average applied, model frame; out_resamplings (optional): resampling steps since initialize */
int dbot_b200_filter_track(dbot_b200_filter* f, const float* depth_m, double* out_states,
                           int32_t* out_resamplings);

/* particle-sharded evaluation with one single-device handle per process (dbot_ros_b200/sharded.py): the update
 * of a call in which this shard received no particle -- every slot is given back and the handle's update clock
 * advances as on the other shards */
int dbot_b200_update_empty(dbot_b200_sensor* s);
/* ---- the gather of log-weights for one process per GPU, fused into the particle kernel over NVLink peer
 * memory (csrc/peer_gather.cuh; SURVEY.md section 8e asks for ONE all-gather of log-weights per call -- here the
 * kernel's epilogue stores every log-weight straight into every rank's gathered vector, shared between the
 * processes as CUDA IPC memory, and one warp exchanges arrival counters; no collective is launched).
 *   export   allocates this rank's gathered vectors for slices of up to width_max particles and returns the
 *            DBOT_B200_IPC_HANDLE_BYTES-byte handle the other ranks need;
 *   connect  takes all world handles (rank order, this rank's own entry is ignored) and maps the peers;
 *   loglikes_gather         RbSensor::loglikes on this rank's slice (host arrays, parents in this rank's own
 *            slot numbering); out_all[g * width + i] = log-weight of particle i of rank g's slice, for all
 *            ranks.  Every rank must make the same sequence of gathering calls (n_local may be 0);
 *   loglikes_gather_device  the same with device pointers and no wait; the gathered vector is at
 *            peer_gather_result() once the stream has run (valid until the next gathering call but one).
 * Errors: -15 a peer's memory cannot be mapped, -16 a peer did not arrive within about 10 s. ---- */
#define DBOT_B200_IPC_HANDLE_BYTES 64
int dbot_b200_peer_gather_export(dbot_b200_sensor* s, int32_t world, int32_t rank, int32_t width_max,
                                 void* out_handle);
int dbot_b200_peer_gather_connect(dbot_b200_sensor* s, const void* handles);
int dbot_b200_loglikes_gather(dbot_b200_sensor* s, const float* poses_local, const int32_t* indices_local,
                              int32_t n_local, int32_t update, int32_t width, double* out_all);
int dbot_b200_loglikes_gather_device(dbot_b200_sensor* s, const float* d_poses, int32_t* d_indices,
                                     int32_t n_local, int32_t update, double* d_out, int32_t width);
void* dbot_b200_peer_gather_result(dbot_b200_sensor* s);
/* host-side check of a sharded call's global parent vector: 0 = every child's parent lives on the child's own
 * rank, 1 = not (out_first: the first such child), -7 = a parent is not a slot.  Ranks own contiguous balanced
 * slices of the n particles and of the n_slots slots (the first n % world ranks hold one more). */
int dbot_b200_shard_scan(const int32_t* indices, int32_t n, int32_t n_slots, int32_t world, int32_t* out_first);
/* number of GPUs behind the handle; occlusion slots that have moved between them since create */
int dbot_b200_device_count(dbot_b200_sensor* s);
int64_t dbot_b200_migrated_slots(dbot_b200_sensor* s);

/* ---- state read-back (tests, debugging) ---- */
int dbot_b200_n_slots(dbot_b200_sensor* s);
/* back-face culling decision for one object: +1 / -1 = the sign of the snapped screen-space area of the
 * triangles that are drawn, 0 = all triangles are drawn */
int dbot_b200_cull_sign(dbot_b200_sensor* s, int32_t object);
double dbot_b200_observation_time(dbot_b200_sensor* s);
/* first and last image row in which a slot's stored image may differ from the background value (first > last:
 * nowhere).  An update copies only these rows of a parent to a child that does not adopt its slot. */
int dbot_b200_slot_rows(dbot_b200_sensor* s, int32_t slot, int32_t* out_first, int32_t* out_last);
/* the same with columns: out_rect[4] = { first row, last row, first column, last column } (columns in steps of 8
 * pixels, the copy's 256-bit unit); only this rectangle of a parent is copied */
int dbot_b200_slot_rect(dbot_b200_sensor* s, int32_t slot, int32_t* out_rect);
/* physical occlusion slots not in use (pool of 2*max_particles; n_slots + free == pool size) */
int dbot_b200_free_slot_count(dbot_b200_sensor* s);
/* occlusion probability image of one slot as of the current observation time, rows*cols floats */
int dbot_b200_get_occlusion(dbot_b200_sensor* s, int32_t slot, float* out_occ);

/* ---- tuning ---- */
/* host threads beside the caller's that help to stage a pageable depth frame into page-locked memory (default 2;
 * 0: the caller's thread stages alone).  Helpers take a third off a single tracker's frame time; ranks of a
 * process-per-GPU group run in lock step, where every extra thread is a source of stragglers: sharded.py sets 0. */
#define DBOT_B200_OPT_STAGING_THREADS 1
int dbot_b200_set_option(dbot_b200_sensor* s, int32_t option, int64_t value);

/* ---- measurement hooks (bench.py) ---- */
/* CUDA events on the handle's stream: begin records, end records + synchronises and returns ms */
int dbot_b200_timer_begin(dbot_b200_sensor* s);
int dbot_b200_timer_end(dbot_b200_sensor* s, float* out_ms);
/* per-kernel accounting: when enabled every kernel launch is bracketed by events */
#define DBOT_B200_KERNEL_OBSERVATION 0 /* observation table precompute           */
#define DBOT_B200_KERNEL_LOGLIKES 1    /* raster + likelihood, update == 0       */
#define DBOT_B200_KERNEL_UPDATE 2      /* raster + likelihood + occlusion stream */
#define DBOT_B200_KERNEL_RENDER 3
#define DBOT_B200_KERNEL_PLAN 4        /* update planning (slot adoption / copies)  */
#define DBOT_B200_KERNEL_COUNT 5
int dbot_b200_profile_enable(dbot_b200_sensor* s, int32_t on);
/* sums since the last call; arrays of DBOT_B200_KERNEL_COUNT entries */
int dbot_b200_profile_read(dbot_b200_sensor* s, float* out_ms, int64_t* out_launches);
/* total kernels launched by this handle since create */
int64_t dbot_b200_launch_count(dbot_b200_sensor* s);
/* bytes of device memory held by the handle */
int64_t dbot_b200_device_bytes(dbot_b200_sensor* s);

int dbot_b200_abi_version(void);
const char* dbot_b200_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
