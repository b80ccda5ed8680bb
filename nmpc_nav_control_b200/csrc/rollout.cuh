// the two pieces a closed-loop rollout needs besides the path discretiser (path_disc.cuh) and the
// controller tick (ctrl_glue.cuh + the RTI step), per robot, NMPC_HD so that tests/host_emul runs the same code.
//
//   plant_step : x+ = phi_RK4(x, u_0 + noise) over one control period with the OCP's own model (the nominal plant of
//                BASELINE config 4; scripts/test_scripts/acados_sim_diff.py:136-160 does the same with an Euler step and
//                Gaussian noise on the accelerations), then the measurements NMPCNavControl*::run reads: pose, body
//                twist (inverse kinematics of the actuator states) and, for tric, the steering angle.
//   nearest_u  : the path parameter nearest to the robot, searched forward from the previous one.  The reference uses
//                parametric_trajectories_common::TPathProcessMinDist (NMPCNavControlROS.cpp:596-600), a private class
//                that is not in the tree; this is a stand-in with its own definition (window search + ternary
//                refinement), not a restatement.
#pragma once
#include "platform.cuh"
#include "rti_core.cuh"
#include "ctrl_glue.cuh"
#include "path_disc.cuh"

namespace nmpc {

template <class M>
struct Rollout {
    using S = Rti<M>;
    static constexpr int NV = M::NV, NX = S::NX;

    // x [NX] stride ld (in/out); u0 [NV] stride ldu; noise [NV] stride ld or null; pose, vel [3] stride ld; steer or null
    NMPC_HD static void plant_step(double* x, size_t ld, const double* u0, size_t ldu, const double* noise, const double* p, double dt,
                                   double* pose, double* vel, double* steer)
    {
        double xs[NX], us[NV], xn[NX], Ep[3][S::NC], lti[4 * NV];
        for (int i = 0; i < NX; i++) xs[i] = x[i * ld];
        for (int i = 0; i < NV; i++) us[i] = u0[i * ldu] + (noise ? noise[i * ld] : 0.0);
        S::rk4_sens(xs, us, p, dt, xn, Ep, lti);
        for (int i = 0; i < NX; i++) x[i * ld] = xn[i];
        pose[0] = xn[0]; pose[ld] = xn[1]; pose[2 * ld] = xn[2];
        double c[3];
        if constexpr (M::ID == 2) {
            // tric: run() reads v and the steering angle; the yaw rate is reported for completeness
            double g[3], Jth[3], Jv[3][NV];
            M::pose_rates(xn[2], &xn[3], p, g, Jth, Jv);
            c[0] = xn[3]; c[1] = 0.0; c[2] = g[2];
            if (steer) steer[0] = xn[4];
        } else {
            CtrlGlue<M>::inverse_kinematics(&xn[3], p, c);
            if constexpr (M::ID == 0) { c[2] = c[1]; c[1] = 0.0; }      // diff returns (v, w, 0): the twist is (v, 0, w)
        }
        vel[0] = c[0]; vel[ld] = c[1]; vel[2 * ld] = c[2];
    }
};

struct PathNearest {
    NMPC_HD static double dist2(const double* segs, int nseg, double su, double px, double py)
    {
        double u, x, y, dx, dy;
        PathDisc::sample(segs, nseg, su, u, x, y, dx, dy);
        return (x - px) * (x - px) + (y - py) * (y - py);
    }
    // search window [u_prev - back, u_prev + ahead] clipped to the path, NS coarse samples, then ternary refinement
    NMPC_HD static double nearest_u(const double* segs, int nseg, double u_prev, double px, double py, double back, double ahead)
    {
        constexpr int NS = 24, NT = 30;
        const double N = (double)nseg;
        double lo = u_prev - back, hi = u_prev + ahead;
        if (lo < 0.0) lo = 0.0;
        if (hi > N) hi = N;
        if (!(lo < hi)) return hi < 0.0 ? 0.0 : hi;
        const double h = (hi - lo) / NS;
        int best = 0;
        double dbest = dist2(segs, nseg, lo, px, py);
        for (int i = 1; i <= NS; i++) {
            const double d = dist2(segs, nseg, lo + h * i, px, py);
            if (d < dbest) { dbest = d; best = i; }
        }
        double a = lo + h * (best > 0 ? best - 1 : 0), b = lo + h * (best < NS ? best + 1 : NS);
        for (int it = 0; it < NT; it++) {
            const double m1 = a + (b - a) / 3.0, m2 = b - (b - a) / 3.0;
            if (dist2(segs, nseg, m1, px, py) <= dist2(segs, nseg, m2, px, py)) b = m2; else a = m1;
        }
        return 0.5 * (a + b);
    }
};

}  // namespace nmpc
