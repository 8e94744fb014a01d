// TEST: instantiate the adapter against the stand-in dbot types and link it with libdbot_b200.so.
#include <array>
#include <cstdio>
#include <vector>
#include "dbot_b200_adapter.hpp"

int main()
{
    typedef dbot::FreeFloatingRigidBodiesState<-1> State;
    Eigen::Matrix3 K = {{72.5, 0, 40, 0, 72.5, 30, 0, 0, 1}};
    std::vector<std::vector<std::array<double, 3> > > verts(1);
    std::vector<std::vector<std::array<int, 3> > > tris(1);
    verts[0] = {{{0, 0, 0}}, {{0.1, 0, 0}}, {{0, 0.1, 0}}};
    tris[0] = {{{0, 1, 2}}};
    try
    {
        dbot::B200KinectImageModel<State> sensor(K, 60, 80, 8, verts, tris, 0.01, 0.003, 0.0014247, 0.1, 0.7,
                                                 0.1, 1.0 / 30.0);
        dbot::RbSensor<State>& base = sensor;
        base.reset();
        std::printf("created\n");
    }
    catch (const std::exception& e)
    {
        std::printf("exception: %s\n", e.what()); /* expected on a machine without a CUDA device */
    }
    return 0;
}
