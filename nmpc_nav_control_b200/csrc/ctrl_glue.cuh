// the controller's per-tick glue around one RTI step, per instance.
//
// What NMPCNavControl{Diff,Omni4,Tric}::run does before and after `{m}_acados_solve`
// (src/nmpc_nav_control/NMPCNavControlDiff.cpp:82-175, NMPCNavControlOmni4.cpp:91-177,
// NMPCNavControlTric.cpp:88-181), restated for a batch: one thread = one robot, every array
// structure-of-arrays with the instance index fastest (the pointers passed in are already offset to
// the instance, `ld` / `ldv` are the strides between components).
//
//   pre : x0 = [pose | measured actuator states (direct kinematics of the body twist) | reference
//         states carried from the previous tick]; yref rows = reference poses with the heading
//         unwrapped along the chain starting at the robot's heading, the list padded with its last
//         pose up to N+1 rows; (diff) terminal weights W_e[0..2] = 100 x W[0..2] when the last two
//         reference rows coincide, else 1 x.
//   post: reference states advanced by u_0 * dt, inverse kinematics -> velocity command; an
//         instance whose solve did not return 0 keeps its carried state and its command (the
//         reference throws before touching either, NMPCNavControl.cpp:15-24).
//
// NMPC_HD functions so that tests/host_emul runs the same code on the CPU.
#pragma once
#include "platform.cuh"
#include "models.cuh"

namespace nmpc {

template <class M>
struct CtrlGlue {
    static constexpr int NV = M::NV, NX = 3 + 2 * NV, NREF = NSTAGE + 1;
    static constexpr double PI = 3.14159265358979323846;

    // Diff.cpp:183-187, Omni4.cpp:185-192, Tric.cpp:96-97 (v and the measured steering angle)
    NMPC_HD static void direct_kinematics(double v, double vn, double w, double steer, const double* p, double* a)
    {
        if constexpr (M::ID == 0) {
            a[0] = v - 0.5 * p[0] * w;
            a[1] = v + 0.5 * p[0] * w;
        } else if constexpr (M::ID == 1) {
            a[0] =  v - vn - 0.5 * p[0] * w;
            a[1] = -v - vn - 0.5 * p[0] * w;
            a[2] =  v + vn - 0.5 * p[0] * w;
            a[3] = -v + vn - 0.5 * p[0] * w;
        } else {
            a[0] = v; a[1] = steer;
        }
    }
    // Diff.cpp:189-193 -> (v, w, 0); Omni4.cpp:194-200 -> (v, vn, w); Tric.cpp:165-166 -> (v, alpha, 0)
    NMPC_HD static void inverse_kinematics(const double* r, const double* p, double* cmd)
    {
        if constexpr (M::ID == 0) {
            cmd[0] = (r[1] + r[0]) / 2.0;
            cmd[1] = (r[1] - r[0]) / p[0];
            cmd[2] = 0.0;
        } else if constexpr (M::ID == 1) {
            cmd[0] = ( r[0] - r[1] + r[2] - r[3]) / 4.0;
            cmd[1] = (-r[0] - r[1] + r[2] + r[3]) / 4.0;
            cmd[2] = (-r[0] - r[1] - r[2] - r[3]) / (2.0 * p[0]);
        } else {
            cmd[0] = r[0]; cmd[1] = r[1]; cmd[2] = 0.0;
        }
    }
    // NMPCNavControl.cpp:26-32
    NMPC_HD static double unwrap(double cur, double prev)
    {
        const double d = cur - prev;
        if (d > PI) cur -= 2.0 * PI;
        else if (d < -PI) cur += 2.0 * PI;
        return cur;
    }

    // pose, vel (v, vn, w), refs [nref_max][3], x0bar [NX], yref [N+1][3], We [NX]: stride ld; vref [NV]: stride ldv.
    // steer may be null (diff, omni4); We null = the shared table is used (omni4, tric: Omni4.cpp:132-139).
    NMPC_HD static void pre(const double* pose, const double* vel, const double* steer, const double* refs, int nref,
                            const double* vref, size_t ldv, const double* p, const double* W0, const double* We_tab,
                            double* x0bar, double* yref, double* We, size_t ld)
    {
        const double th0 = pose[2 * ld];
        x0bar[0] = pose[0]; x0bar[ld] = pose[ld]; x0bar[2 * ld] = th0;
        double a[NV];
        direct_kinematics(vel[0], vel[ld], vel[2 * ld], steer ? steer[0] : 0.0, p, a);
        for (int i = 0; i < NV; i++) {
            x0bar[(3 + i) * ld] = a[i];
            x0bar[(3 + NV + i) * ld] = vref[i * ldv];
        }
        if (nref < 1) nref = 1;                  // the reference reads yref[-1] for an empty list; one pose is the least meaningful input
        if (nref > NREF) nref = NREF;
        double rx = 0.0, ry = 0.0, rt = th0, px = 0.0, py = 0.0, pt = 0.0;
        for (int i = 0; i < NREF; i++) {
            px = rx; py = ry; pt = rt;
            if (i < nref) {
                rx = refs[(size_t)(3 * i) * ld]; ry = refs[(size_t)(3 * i + 1) * ld];
                rt = unwrap(refs[(size_t)(3 * i + 2) * ld], rt);
            }
            yref[(size_t)(3 * i) * ld] = rx; yref[(size_t)(3 * i + 1) * ld] = ry; yref[(size_t)(3 * i + 2) * ld] = rt;
        }
        if (We) {                                // Diff.cpp:126-139: rows N and N-1 equal (bitwise, as the reference compares)
            const double f = (rx == px && ry == py && rt == pt) ? 100.0 : 1.0;
            for (int i = 0; i < NX; i++) We[i * ld] = i < 3 ? f * W0[i] : We_tab[i];
        }
    }

    // u0 [NV] stride ldu (stage 0 of the iterate after the step); cmd [3] stride ld
    NMPC_HD static void post(int status, const double* x0bar, size_t ld, const double* u0, size_t ldu, double dt, const double* p,
                             double* vref, size_t ldv, double* cmd)
    {
        if (status != 0) return;
        double r[NV], c[3];
        for (int i = 0; i < NV; i++) r[i] = x0bar[(3 + NV + i) * ld] + u0[i * ldu] * dt;    // Diff.cpp:155-157
        inverse_kinematics(r, p, c);
        cmd[0] = c[0]; cmd[ld] = c[1]; cmd[2 * ld] = c[2];
        for (int i = 0; i < NV; i++) vref[i * ldv] = r[i];                                   // Diff.cpp:171-172
    }
};

}  // namespace nmpc
