/*
 * include/dbot_b200_adapter.hpp -- the C++ side of the drop-in: a dbot::RbSensor<State> backed by
 * libdbot_b200.so.  A dbot maintainer adds this header next to dbot's gpu/kinect_image_model_gpu.h and
 * returns it from RbSensorBuilder<State>::create_gpu_based_model(); dbot_ros then
 * needs no change at all: ref:source/dbot_ros/tracker/particle_tracker_node.cpp:164-203 keeps filling
 * RbSensorBuilder<State>::Parameters and use_gpu:true (ref:config/particle_tracker.yaml:3) now selects
 * the B200 evaluator.
 *
 * It needs dbot's own headers (dbot/model/rb_sensor.h, dbot/pose/*) and Eigen, none of which exist in
 * this repository's build image; tests/test_adapter_compiles.py compiles it against minimal stand-ins
 * of those headers (tests/stubs/) and, on a GPU, runs set_observation / loglikes through it and compares the
 * log-weights with the oracle (tests/stubs/adapter_run.cpp).
 *
 * What it does is what dbot's KinectImageModelGPU does on the host before its GPU call: compose every
 * particle's delta pose with the integrated (default) pose in double, round R|t to float32, and pass
 * [n][n_objects][12] to the evaluator.  No arithmetic of the observation model happens here.
 */
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include <dbot/model/rb_sensor.h>

#include "dbot_b200.h"

namespace dbot
{
template <typename State>
class B200KinectImageModel : public RbSensor<State>
{
public:
    typedef typename RbSensor<State>::StateArray StateArray;
    typedef typename RbSensor<State>::RealArray RealArray;
    typedef typename RbSensor<State>::IntArray IntArray;
    typedef typename RbSensor<State>::Observation Observation;

    /* camera_matrix: 3x3 (row, col) accessor, rows 0-1 already divided by the downsampling factor
     * (ref:source/dbot_ros/util/ros_camera_data_provider.cpp:72); vertices[o][v] / triangles[o][t]:
     * per-part lists exactly as dbot::ObjectModel hands them to KinectImageModelGPU. */
    template <typename Matrix3, typename VertexLists, typename TriangleLists>
    B200KinectImageModel(const Matrix3& camera_matrix, int n_rows, int n_cols, int max_sample_count,
                         const VertexLists& vertices, const TriangleLists& triangles,
                         double tail_weight, double model_sigma, double sigma_factor,
                         double p_occluded_visible, double p_occluded_occluded,
                         double initial_occlusion_prob, double delta_time, int device_id = 0,
                         double half_life_depth = 1.0, double max_depth = 6.0,
                         const std::vector<int>& device_ids = std::vector<int>(), int gather_mode = 0,
                         bool cull_backfaces = true)
        : n_objects_((int)vertices.size()), handle_(nullptr)
    {
        /* device_ids: two or more CUDA ordinals shard the particles over those GPUs inside the library (one
         * process, one sensor object -- what RbSensorBuilder::build() must return); gather_mode 0 = NCCL
         * all-gather of the log-weights, 1 = stores over NVLink peer memory from the kernel's epilogue.
         * cull_backfaces = false draws every triangle, as dbot's RigidBodyRenderer does. */
        dbot_b200_config cfg = dbot_b200_config();
        cfg.rows = n_rows;
        cfg.cols = n_cols;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 3; ++c) cfg.K[3 * r + c] = camera_matrix(r, c);
        cfg.max_particles = max_sample_count;
        cfg.n_objects = n_objects_;
        std::vector<std::vector<float> > v(n_objects_);
        std::vector<std::vector<int32_t> > t(n_objects_);
        std::vector<const float*> vp(n_objects_);
        std::vector<const int32_t*> tp(n_objects_);
        std::vector<int32_t> nv(n_objects_), nt(n_objects_);
        for (int o = 0; o < n_objects_; ++o)
        {
            for (size_t i = 0; i < vertices[o].size(); ++i)
                for (int k = 0; k < 3; ++k) v[o].push_back((float)vertices[o][i][k]);
            for (size_t i = 0; i < triangles[o].size(); ++i)
                for (int k = 0; k < 3; ++k) t[o].push_back((int32_t)triangles[o][i][k]);
            vp[o] = v[o].data();
            tp[o] = t[o].data();
            nv[o] = (int32_t)vertices[o].size();
            nt[o] = (int32_t)triangles[o].size();
        }
        cfg.vertices = vp.data();
        cfg.n_vertices = nv.data();
        cfg.triangles = tp.data();
        cfg.n_triangles = nt.data();
        cfg.tail_weight = tail_weight;
        cfg.model_sigma = model_sigma;
        cfg.sigma_factor = sigma_factor;
        cfg.half_life_depth = half_life_depth;
        cfg.max_depth = max_depth;
        cfg.p_occluded_visible = p_occluded_visible;
        cfg.p_occluded_occluded = p_occluded_occluded;
        cfg.initial_occlusion_prob = initial_occlusion_prob;
        cfg.delta_time = delta_time;
        cfg.device_id = device_id;
        cfg.occlusion_mode = 0;
        cfg.cull_backfaces = cull_backfaces ? 0 : 1;
        std::vector<int32_t> ids(device_ids.begin(), device_ids.end());
        cfg.n_devices = (int32_t)ids.size();
        cfg.device_ids = ids.empty() ? nullptr : ids.data();
        cfg.gather_mode = gather_mode;
        check(dbot_b200_create(&cfg, &handle_), "create");
        this->default_poses_.resize(n_objects_);
    }

    virtual ~B200KinectImageModel()
    {
        if (handle_) dbot_b200_destroy(handle_);
    }

    void reset() { check(dbot_b200_reset(handle_), "reset"); }

    void set_observation(const Observation& image)
    {
        /* Observation is a vector of fl::Real (double) holding float32 pixel values
         * (ref:source/dbot_ros/util/ros_interface.h:152-168) */
        frame_.resize((size_t)image.size());
        for (size_t i = 0; i < frame_.size(); ++i) frame_[i] = (float)image[i];
        check(dbot_b200_set_observation(handle_, frame_.data()), "set_observation");
    }

    RealArray loglikes(const StateArray& deltas, IntArray& indices, const bool& update = false)
    {
        const int n = (int)deltas.size();
        poses_.resize((size_t)n * n_objects_ * 12);
        parents_.resize(n);
        out_.resize(n);
        for (int i = 0; i < n; ++i)
        {
            parents_[i] = (int32_t)indices[i];
            for (int o = 0; o < n_objects_; ++o)
            {
                /* pose = delta (x) default: R = R_delta R_default, t = t_delta + t_default, in double */
                const auto Rd = deltas[i].component(o).orientation().rotation_matrix();
                const auto R0 = this->default_poses_[o].orientation().rotation_matrix();
                const auto td = deltas[i].component(o).position();
                const auto t0 = this->default_poses_[o].position();
                float* p = &poses_[((size_t)i * n_objects_ + o) * 12];
                for (int r = 0; r < 3; ++r)
                    for (int c = 0; c < 3; ++c)
                    {
                        double acc = 0.0;
                        for (int k = 0; k < 3; ++k) acc += Rd(r, k) * R0(k, c);
                        p[3 * r + c] = (float)acc;
                    }
                for (int k = 0; k < 3; ++k) p[9 + k] = (float)(td[k] + t0[k]);
            }
        }
        check(dbot_b200_loglikes(handle_, poses_.data(), parents_.data(), n, update ? 1 : 0, out_.data()),
              "loglikes");
        RealArray result;
        result.resize(n);
        for (int i = 0; i < n; ++i) result[i] = out_[i];
        if (update)
            for (int i = 0; i < n; ++i) indices[i] = parents_[i]; /* identity after the update */
        return result;
    }

private:
    void check(int rc, const char* what)
    {
        /* the C-ABI never throws; dbot's GPU model reports "too many poses" etc. as exceptions */
        if (rc != 0)
            throw std::runtime_error(std::string("dbot_b200_") + what + ": " + dbot_b200_last_error());
    }

    int n_objects_;
    dbot_b200_sensor* handle_;
    std::vector<float> frame_, poses_;
    std::vector<int32_t> parents_;
    std::vector<double> out_;
};
}  // namespace dbot
