// Minimal C++ caller of the batched C ABI (include/nmpc_b200.h), the shape a maintainer of the reference would write
// to drive a fleet of diff robots from one process: the constructor part of NMPCNavControlDiff
// (src/nmpc_nav_control/NMPCNavControlDiff.cpp:6-74) becomes nmpc_create + nmpc_set_*, run() (:82-175) becomes
// nmpc_ctrl_tick_host, reset_mpc() (:177-181) becomes nmpc_reset.
//
//   g++ -std=c++14 -O2 -I include examples/fleet_tick.cpp -L nmpc_nav_control_b200 -lnmpc_b200 \
//       -Wl,-rpath,$PWD/nmpc_nav_control_b200 -o fleet_tick && ./fleet_tick 1024
//
// Exit code 0 = ticks done, 2 = the library reported an error (e.g. no CUDA device: there is no CPU fallback).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "nmpc_b200.h"

#define CHECK(call)                                                               \
    do {                                                                          \
        int rc_ = (call);                                                         \
        if (rc_ != 0) { std::fprintf(stderr, "%s -> %d: %s\n", #call, rc_, nmpc_last_error()); return 2; } \
    } while (0)

int main(int argc, char** argv)
{
    const int B = argc > 1 ? std::atoi(argv[1]) : 256;
    const double dt = 0.025;                              // 1 / control_freq (NMPCNavControlROS.cpp:82)
    nmpc_dims_t d;
    CHECK(nmpc_dims(NMPC_MODEL_DIFF, &d));
    nmpc_solver* s = nullptr;
    CHECK(nmpc_create(NMPC_MODEL_DIFF, B, /*device*/ 0, &s));

    // constructor part: W from W_diag on every stage, W_e from W_diag[0..nx) (Diff.cpp:24-41, 68-73)
    const double W_diag[9] = {10, 10, 5, 0, 0, 0, 0, 1, 1};
    std::vector<double> W((size_t)d.n * d.ny), We(d.nx);
    for (int k = 0; k < d.n; k++) for (int i = 0; i < d.ny; i++) W[(size_t)k * d.ny + i] = W_diag[i];
    for (int i = 0; i < d.nx; i++) We[i] = W_diag[i];
    CHECK(nmpc_set_weights(s, W.data(), We.data()));
    CHECK(nmpc_reset(s));                                 // reset_mpc() on a new path (NMPCNavControlROS.cpp:309-326)
    CHECK(nmpc_ctrl_reset(s, nullptr));

    // every robot follows a circle of its own radius; N+1 reference poses per tick, instance-major host arrays
    const int NREF = NMPC_N + 1;
    std::vector<double> pose((size_t)B * 3), vel((size_t)B * 3, 0.0), refs((size_t)B * NREF * 3), cmd((size_t)B * 3, 0.0);
    std::vector<int> status(B), qp_iter(B);
    for (int tick = 0; tick < 5; tick++) {
        for (int i = 0; i < B; i++) {
            const double R = 1.0 + 0.001 * i, v = 0.4, a0 = v * dt * tick / R;
            pose[3 * i] = R * std::cos(a0); pose[3 * i + 1] = R * std::sin(a0); pose[3 * i + 2] = std::remainder(a0 + M_PI / 2, 2 * M_PI);
            vel[3 * i] = tick ? cmd[3 * i] : 0.0; vel[3 * i + 2] = tick ? cmd[3 * i + 1] : 0.0;      // (v, vn, w) <- last command (v, w, 0)
            for (int k = 0; k < NREF; k++) {
                const double a = a0 + v * dt * k / R;
                double* r = &refs[((size_t)i * NREF + k) * 3];
                r[0] = R * std::cos(a); r[1] = R * std::sin(a); r[2] = std::remainder(a + M_PI / 2, 2 * M_PI);   // wrapped heading
            }
        }
        CHECK(nmpc_ctrl_tick_host(s, B, pose.data(), vel.data(), /*steer*/ nullptr, refs.data(), /*nref*/ nullptr, NREF, dt,
                                  cmd.data(), status.data(), qp_iter.data()));
        int bad = 0; double it = 0;
        for (int i = 0; i < B; i++) { bad += status[i] != 0; it += qp_iter[i]; }
        std::printf("tick %d: robot 0 cmd v=%.6f w=%.6f, mean qp_iter %.2f, failed %d\n", tick, cmd[0], cmd[1], it / B, bad);
    }
    CHECK(nmpc_destroy(s));
    return 0;
}
