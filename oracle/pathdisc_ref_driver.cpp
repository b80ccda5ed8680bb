// TEST INFRASTRUCTURE.  C entry point around the reference's UNMODIFIED PathDiscretizer
// (/root/reference/src/nmpc_nav_control/PathDiscretizer.cpp, compiled in place by `make -C oracle ref`; nothing is
// copied) with the stand-in TPath of oracle/stubs.  Output library: oracle/_ref/libpathdisc_ref.so.
#include <list>
#include <vector>
#include "nmpc_nav_control/PathDiscretizer.h"

// segments [n_seg][16] (kind, vel, th0, th1, cx[6], cy[6]); poses_out [num_poses][3]; returns the number of poses
extern "C" int pathdisc_ref(const double* segments, int n_seg, double nearest_sample_u, double sample_period, int num_poses,
                            int is_holonomic, double* poses_out)
{
    nmpc_nav_control::TPathList lst;
    for (int s = 0; s < n_seg; s++) {
        parametric_trajectories_common::TPath p;
        const double* r = segments + 16 * s;
        p.kind = r[0]; p.vel = r[1]; p.th0 = r[2]; p.th1 = r[3];
        for (int i = 0; i < 6; i++) { p.cx[i] = r[4 + i]; p.cy[i] = r[10 + i]; }
        lst.push_back(p);
    }
    nmpc_nav_control::PathDiscretizer d(sample_period, num_poses, is_holonomic != 0);
    std::vector<nmpc_nav_control::PathDiscretizer::Pose> out;
    d.getNextNPoses(lst, nearest_sample_u, out);
    for (size_t i = 0; i < out.size(); i++) { poses_out[3 * i] = out[i].x; poses_out[3 * i + 1] = out[i].y; poses_out[3 * i + 2] = out[i].theta; }
    return (int)out.size();
}
