// TEST STAND-IN (not dbot's header): the smallest set of types that lets
// include/dbot_b200_adapter.hpp compile in an image without dbot / Eigen / fl.  Shapes follow dbot's
// model/rb_sensor.h and pose/* as described in SURVEY.md section 2c.
#pragma once
#include <cstddef>
#include <vector>

namespace fl { typedef double Real; }

namespace Eigen {
template <typename T> struct Array1 {
    std::vector<T> v;
    size_t size() const { return v.size(); }
    void resize(size_t n) { v.resize(n); }
    T& operator[](size_t i) { return v[i]; }
    const T& operator[](size_t i) const { return v[i]; }
};
struct Matrix3 { double m[9]; double operator()(int r, int c) const { return m[3 * r + c]; } };
struct Vector3 { double m[3]; double operator[](int i) const { return m[i]; } };
}  // namespace Eigen

namespace dbot {
struct EulerVector { Eigen::Matrix3 R; Eigen::Matrix3 rotation_matrix() const { return R; } };
struct PoseVector {
    Eigen::Vector3 p; EulerVector e;
    Eigen::Vector3 position() const { return p; }
    EulerVector orientation() const { return e; }
};
template <int N> struct FreeFloatingRigidBodiesState {
    std::vector<PoseVector> bodies;
    const PoseVector& component(int i) const { return bodies[i]; }
};
template <typename State> class RbSensor {
public:
    typedef Eigen::Array1<State> StateArray;
    typedef Eigen::Array1<fl::Real> RealArray;
    typedef Eigen::Array1<int> IntArray;
    typedef Eigen::Array1<fl::Real> Observation;
    typedef std::vector<PoseVector> PoseArray;
    virtual ~RbSensor() {}
    virtual RealArray loglikes(const StateArray& deltas, IntArray& indices, const bool& update = false) = 0;
    virtual void set_observation(const Observation& image) = 0;
    virtual void reset() = 0;
    virtual PoseArray& integrated_poses() { return default_poses_; }
protected:
    PoseArray default_poses_;
};
}  // namespace dbot
