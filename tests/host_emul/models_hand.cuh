// TEST INFRASTRUCTURE: the hand-written model functions of round 1, kept as the comparison target of the emitted
// csrc/models_gen.cuh (tests/test_emit_cpu.py).  Robot models as device functions (K1 input).
//
// All three reference models share one structure:
//   x = [pose(3) = (x, y, theta) | actual(NV) | ref(NV)],  u = d(ref)/dt (NV)
//   pose_dot   = g(theta, actual; p)                       (model specific, nonlinear)
//   actual_dot = (ref - actual) / tau_c                    (first-order lag per channel)
//   ref_dot    = u
// so a model is described by NV, NP, tau_c(p) and g with its Jacobian w.r.t. (theta, actual).
//
//   DiffModel  : scripts/diff/diff_amr_model.py:42-60    (x=[x,y,th,vl,vr,vl_ref,vr_ref], p=[dist_b,tau_v])
//   Omni4Model : scripts/omni4/omni4_amr_model.py:52-73  (x=[x,y,th,v1..4,v1..4_ref],    p=[l1_plus_l2,tau_v])
//   TricModel  : scripts/tric/tric_amr_model.py:43-59    (x=[x,y,th,v,alpha,v_ref,alpha_ref], p=[dist_d,tau_v,tau_a])
//                tric_amr_model.py:45 defines cos_alpha = sin(alpha); reproduced (TRIC_FAITHFUL_COS_BUG).
#pragma once
#include "../../nmpc_nav_control_b200/csrc/platform.cuh"

#ifndef TRIC_FAITHFUL_COS_BUG
#define TRIC_FAITHFUL_COS_BUG 1
#endif

namespace nmpc_hand {
using nmpc::nmpc_sincos;

struct DiffModel {
    static constexpr int NV = 2, NP = 2, ID = 0;
    static constexpr bool THETA_ROW_LTI = true;
    NMPC_HD static double inv_tau(int, const double* p) { return 1.0 / p[1]; }
    // g[3], Jth[3] = dg/dtheta, Jv[3][NV] = dg/dactual
    NMPC_HD static void pose_rates(double th, const double* a, const double* p, double* g, double* Jth, double (*Jv)[NV]) {
        double s, c; nmpc_sincos(th, &s, &c);
        const double v = (a[1] + a[0]) / 2.0;
        const double ib = 1.0 / p[0];
        g[0] = v * c; g[1] = v * s; g[2] = (a[1] - a[0]) * ib;
        Jth[0] = -v * s; Jth[1] = v * c; Jth[2] = 0.0;
        Jv[0][0] = 0.5 * c; Jv[0][1] = 0.5 * c;
        Jv[1][0] = 0.5 * s; Jv[1][1] = 0.5 * s;
        Jv[2][0] = -ib;     Jv[2][1] = ib;
    }
};

struct Omni4Model {
    static constexpr int NV = 4, NP = 2, ID = 1;
    static constexpr bool THETA_ROW_LTI = true;
    NMPC_HD static double inv_tau(int, const double* p) { return 1.0 / p[1]; }
    NMPC_HD static void pose_rates(double th, const double* a, const double* p, double* g, double* Jth, double (*Jv)[NV]) {
        double s, c; nmpc_sincos(th, &s, &c);
        const double v  = ( a[0] - a[1] + a[2] - a[3]) / 4.0;
        const double vn = (-a[0] - a[1] + a[2] + a[3]) / 4.0;
        const double iw = 1.0 / (2.0 * p[0]);
        g[0] = v * c - vn * s; g[1] = v * s + vn * c; g[2] = (-a[0] - a[1] - a[2] - a[3]) * iw;
        Jth[0] = -v * s - vn * c; Jth[1] = v * c - vn * s; Jth[2] = 0.0;
        const double sv[4]  = { 1.0, -1.0, 1.0, -1.0};
        const double svn[4] = {-1.0, -1.0, 1.0,  1.0};
#pragma unroll
        for (int i = 0; i < 4; i++) {
            Jv[0][i] = (sv[i] * c - svn[i] * s) / 4.0;
            Jv[1][i] = (sv[i] * s + svn[i] * c) / 4.0;
            Jv[2][i] = -iw;
        }
    }
};

struct TricModel {
    static constexpr int NV = 2, NP = 3, ID = 2;
    static constexpr bool THETA_ROW_LTI = false;
    NMPC_HD static double inv_tau(int c, const double* p) { return 1.0 / p[1 + c]; }
    NMPC_HD static void pose_rates(double th, const double* a, const double* p, double* g, double* Jth, double (*Jv)[NV]) {
        double s, c; nmpc_sincos(th, &s, &c);
        double sa, ca; nmpc_sincos(a[1], &sa, &ca);
#if TRIC_FAITHFUL_COS_BUG
        const double f = sa, df = ca;          // "cos_alpha" = sin(alpha)
#else
        const double f = ca, df = -sa;
#endif
        const double id = 1.0 / p[0];
        g[0] = a[0] * c * f; g[1] = a[0] * s * f; g[2] = a[0] * id * sa;
        Jth[0] = -a[0] * s * f; Jth[1] = a[0] * c * f; Jth[2] = 0.0;
        Jv[0][0] = c * f;   Jv[0][1] = a[0] * c * df;
        Jv[1][0] = s * f;   Jv[1][1] = a[0] * s * df;
        Jv[2][0] = id * sa; Jv[2][1] = a[0] * id * ca;
    }
};

}  // namespace nmpc_hand
