// PathDiscretizer::getNextNPoses (src/nmpc_nav_control/PathDiscretizer.cpp:14-63) for one robot,
// over segments of sixteen doubles (nmpc_path_segment in include/nmpc_b200.h: kind, vel, th0, th1, cx[6], cy[6]).
//
// The walk is inherently sequential per robot (each step length depends on the tangent at the previous sample, each
// emitted pose on the chord length accumulated since the last one), so the batch dimension is the parallel one: a
// thread per robot, the poses written structure-of-arrays ([pose][x|y|theta][robot]) so that consecutive robots store
// consecutive doubles - the layout nmpc_ctrl_tick_device reads.  NMPC_HD so that tests/host_emul runs the same code.
//
// Deviations from the reference, both where it has undefined behaviour: the segment index taken from
// nearest_sample_u (:23) is clamped to the path; the walk is bounded by PATH_MAX_STEPS steps (a vanishing step length
// never terminates upstream), after which the list is padded with the path's end pose like a finished path.
#pragma once
#include "platform.cuh"

namespace nmpc {

struct PathDisc {
    static constexpr int SEG = 16;
    static constexpr int PATH_MAX_STEPS = 1 << 18;
    static constexpr double PI = 3.14159265358979323846;

    NMPC_HD static double poly(const double* c, double u) { return ((((c[5] * u + c[4]) * u + c[3]) * u + c[2]) * u + c[1]) * u + c[0]; }
    NMPC_HD static double dpoly(const double* c, double u)
    {
        return (((5.0 * c[5] * u + 4.0 * c[4]) * u + 3.0 * c[3]) * u + 2.0 * c[2]) * u + c[1];
    }
    // getPoseSample / getVelSample's clamping of (segment, parameter), :67-76
    NMPC_HD static const double* locate(const double* segs, int nseg, double su, double& u)
    {
        const double fl = floor(su);
        int num = (int)fl;
        u = su - fl;
        if (!(fl < (double)nseg)) { num = nseg - 1; u = 1.0; }      // also a NaN parameter: upstream indexes out of range
        else if (fl < 0.0) { num = 0; u = 0.0; }
        return segs + (size_t)num * SEG;
    }
    // position and tangent at path parameter su; returns the segment and its local parameter for heading()
    NMPC_HD static const double* sample(const double* segs, int nseg, double su, double& u, double& x, double& y, double& dx, double& dy)
    {
        const double* s = locate(segs, nseg, su, u);
        if (s[0] == 0.0) {
            x = poly(s + 4, u); y = poly(s + 10, u); dx = dpoly(s + 4, u); dy = dpoly(s + 10, u);
        } else {
            double sn, cs;
            nmpc_sincos(s[6] + s[7] * u, &sn, &cs);
            x = s[4] + s[5] * cs; y = s[10] + s[5] * sn; dx = -s[5] * s[7] * sn; dy = s[5] * s[7] * cs;
        }
        return s;
    }
    // the heading getPoseSample returns (:78-84); only evaluated for the poses that are emitted - upstream computes it
    // at every sub-step and throws nine of ten away
    NMPC_HD static double heading(const double* s, double u, double dx, double dy, bool hol)
    {
        if (hol) return s[2] + (s[3] - s[2]) * u;                   // GetThetaHolomonic
        double th = atan2(dy, dx);
        if (!(s[1] >= 0.0)) th += PI;
        return th;
    }
    NMPC_HD static double seg_speed(const double* segs, int nseg, double fl)
    {
        int i = fl < 0.0 ? 0 : (fl < (double)nseg ? (int)fl : nseg - 1);
        return fabs(segs[(size_t)i * SEG + 1]);
    }

    // out [num_poses][3] with stride ld between consecutive doubles of one robot
    NMPC_HD static void next_poses(const double* segs, int nseg, double u0, double period, int num_poses, bool hol, double* out,
                                   size_t ld)
    {
        const double per_cycle = period >= 1.0 ? 20.0 : 10.0;      // :9-11
        const double thr = 1e-2, N = (double)nseg;
        double goal = seg_speed(segs, nseg, floor(u0)) * period;
        double rel = goal / per_cycle;
        double u = u0, ox, oy, dx, dy, ul;
        sample(segs, nseg, u0, ul, ox, oy, dx, dy);
        double step = rel / sqrt(dx * dx + dy * dy);
        double curr = 0.0;
        int n = 0;
        for (int itc = 0; u < N && itc < PATH_MAX_STEPS; itc++) {  // :34-56
            u += step;
            u = (N < u) ? N : u;                                    // std::min(u, N): a NaN stays and ends the walk
            double nx, ny;
            const double* sg = sample(segs, nseg, u, ul, nx, ny, dx, dy);
            const double ex = nx - ox, ey = ny - oy;
            curr += sqrt(ex * ex + ey * ey);
            if ((goal - curr) <= thr * goal) {
                out[(size_t)(3 * n) * ld] = nx; out[(size_t)(3 * n + 1) * ld] = ny;
                out[(size_t)(3 * n + 2) * ld] = heading(sg, ul, dx, dy, hol);
                n++;
                const double fl = floor(u);
                goal = seg_speed(segs, nseg, fl < N - 1.0 ? fl : N - 1.0) * period;
                rel = goal / per_cycle;
                curr = 0.0;
            }
            if (n == num_poses) break;
            step = rel / sqrt(dx * dx + dy * dy);
            ox = nx; oy = ny;
        }
        if (n < num_poses) {                                        // :58-63
            double lx, ly;
            const double* sg = sample(segs, nseg, N, ul, lx, ly, dx, dy);
            const double lt = heading(sg, ul, dx, dy, hol);
            for (; n < num_poses; n++) {
                out[(size_t)(3 * n) * ld] = lx; out[(size_t)(3 * n + 1) * ld] = ly; out[(size_t)(3 * n + 2) * ld] = lt;
            }
        }
    }
};

}  // namespace nmpc
