/*
 * ORACLE (test infrastructure, NOT product code) — robot model ODEs and Jacobians.
 *
 * CPU restatement of the three continuous-time models of JorgeDFR/nmpc_nav_control.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load anything under oracle/.  PARITY UNPINNED: the reference's arithmetic
 * lives in acados/HPIPM/BLASFEO/CasADi (un-vendored, version unpinned, package.xml:29),
 * none of which exists in this environment, and the reference holds no golden vectors.
 *
 * Follows:
 *   diff  : scripts/diff/diff_amr_model.py:42-60   (second vl_dot/vr_dot assignment wins, :53-54)
 *   omni4 : scripts/omni4/omni4_amr_model.py:52-73
 *   tric  : scripts/tric/tric_amr_model.py:43-59   (cos_alpha = sin(alpha) at :45 is reproduced)
 *
 * Each model gives f(x,u,p) and the dense Jacobian J = [df/dx | df/du] (nx x (nx+nu), row-major).
 */
#ifndef ORC_MODELS_H
#define ORC_MODELS_H
#include <math.h>
#include <string.h>

/* ---------------- diff2amr: nx=7 nu=2 np=2 ---------------- */
/* x=[x,y,theta,vl,vr,vl_ref,vr_ref] u=[dvl_ref,dvr_ref] p=[dist_b,tau_v] */
static inline void orc_diff_f(const double *x, const double *u, const double *p, double *f)
{
    const double c = cos(x[2]), s = sin(x[2]);
    const double v = (x[4] + x[3]) / 2.0;
    const double w = (x[4] - x[3]) / p[0];
    const double it = 1.0 / p[1];
    f[0] = v * c;
    f[1] = v * s;
    f[2] = w;
    f[3] = -it * x[3] + it * x[5];
    f[4] = -it * x[4] + it * x[6];
    f[5] = u[0];
    f[6] = u[1];
}
static inline void orc_diff_jac(const double *x, const double *u, const double *p, double *J)
{
    (void)u;
    enum { NX = 7, NZ = 9 };
    const double c = cos(x[2]), s = sin(x[2]);
    const double v = (x[4] + x[3]) / 2.0;
    const double it = 1.0 / p[1];
    memset(J, 0, sizeof(double) * NX * NZ);
    J[0 * NZ + 2] = -v * s; J[0 * NZ + 3] = 0.5 * c; J[0 * NZ + 4] = 0.5 * c;
    J[1 * NZ + 2] =  v * c; J[1 * NZ + 3] = 0.5 * s; J[1 * NZ + 4] = 0.5 * s;
    J[2 * NZ + 3] = -1.0 / p[0]; J[2 * NZ + 4] = 1.0 / p[0];
    J[3 * NZ + 3] = -it; J[3 * NZ + 5] = it;
    J[4 * NZ + 4] = -it; J[4 * NZ + 6] = it;
    J[5 * NZ + 7] = 1.0;
    J[6 * NZ + 8] = 1.0;
}

/* ---------------- omni4amr: nx=11 nu=4 np=2 ---------------- */
/* x=[x,y,theta,v1..v4,v1_ref..v4_ref] u=[dv1_ref..dv4_ref] p=[l1_plus_l2,tau_v] */
static inline void orc_omni4_f(const double *x, const double *u, const double *p, double *f)
{
    const double c = cos(x[2]), s = sin(x[2]);
    const double v  = ( x[3] - x[4] + x[5] - x[6]) / 4.0;
    const double vn = (-x[3] - x[4] + x[5] + x[6]) / 4.0;
    const double w  = (-x[3] - x[4] - x[5] - x[6]) / (2.0 * p[0]);
    const double it = 1.0 / p[1];
    f[0] = v * c - vn * s;
    f[1] = v * s + vn * c;
    f[2] = w;
    for (int i = 0; i < 4; i++) f[3 + i] = -it * x[3 + i] + it * x[7 + i];
    for (int i = 0; i < 4; i++) f[7 + i] = u[i];
}
static inline void orc_omni4_jac(const double *x, const double *u, const double *p, double *J)
{
    (void)u;
    enum { NX = 11, NZ = 15 };
    const double c = cos(x[2]), s = sin(x[2]);
    const double v  = ( x[3] - x[4] + x[5] - x[6]) / 4.0;
    const double vn = (-x[3] - x[4] + x[5] + x[6]) / 4.0;
    const double it = 1.0 / p[1];
    static const double sv[4]  = { 1.0, -1.0, 1.0, -1.0};   /* dv /dvi * 4 */
    static const double svn[4] = {-1.0, -1.0, 1.0,  1.0};   /* dvn/dvi * 4 */
    memset(J, 0, sizeof(double) * NX * NZ);
    J[0 * NZ + 2] = -v * s - vn * c;
    J[1 * NZ + 2] =  v * c - vn * s;
    for (int i = 0; i < 4; i++) {
        J[0 * NZ + 3 + i] = (sv[i] * c - svn[i] * s) / 4.0;
        J[1 * NZ + 3 + i] = (sv[i] * s + svn[i] * c) / 4.0;
        J[2 * NZ + 3 + i] = -1.0 / (2.0 * p[0]);
        J[(3 + i) * NZ + 3 + i] = -it;
        J[(3 + i) * NZ + 7 + i] = it;
        J[(7 + i) * NZ + 11 + i] = 1.0;
    }
}

/* ---------------- tric3amr: nx=7 nu=2 np=3 ---------------- */
/* x=[x,y,theta,v,alpha,v_ref,alpha_ref] u=[dv_ref,dalpha_ref] p=[dist_d,tau_v,tau_a] */
static inline void orc_tric_f(const double *x, const double *u, const double *p, double *f)
{
    const double c = cos(x[2]), s = sin(x[2]);
    const double sa = sin(x[4]);
    const double ca_bug = sin(x[4]);           /* tric_amr_model.py:45  cos_alpha = ca.sin(alpha) */
    const double itv = 1.0 / p[1], ita = 1.0 / p[2];
    f[0] = x[3] * c * ca_bug;
    f[1] = x[3] * s * ca_bug;
    f[2] = x[3] / p[0] * sa;
    f[3] = -itv * x[3] + itv * x[5];
    f[4] = -ita * x[4] + ita * x[6];
    f[5] = u[0];
    f[6] = u[1];
}
static inline void orc_tric_jac(const double *x, const double *u, const double *p, double *J)
{
    (void)u;
    enum { NX = 7, NZ = 9 };
    const double c = cos(x[2]), s = sin(x[2]);
    const double sa = sin(x[4]), dsa = cos(x[4]);   /* d/dalpha of sin(alpha) (both factors) */
    const double itv = 1.0 / p[1], ita = 1.0 / p[2];
    memset(J, 0, sizeof(double) * NX * NZ);
    J[0 * NZ + 2] = -x[3] * s * sa; J[0 * NZ + 3] = c * sa; J[0 * NZ + 4] = x[3] * c * dsa;
    J[1 * NZ + 2] =  x[3] * c * sa; J[1 * NZ + 3] = s * sa; J[1 * NZ + 4] = x[3] * s * dsa;
    J[2 * NZ + 3] = sa / p[0];      J[2 * NZ + 4] = x[3] / p[0] * dsa;
    J[3 * NZ + 3] = -itv; J[3 * NZ + 5] = itv;
    J[4 * NZ + 4] = -ita; J[4 * NZ + 6] = ita;
    J[5 * NZ + 7] = 1.0;
    J[6 * NZ + 8] = 1.0;
}

#endif
