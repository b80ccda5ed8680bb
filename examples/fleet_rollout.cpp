// C++ caller of the closed-loop rollout engine (include/nmpc_b200.h: nmpc_rollout_device), BASELINE config 4 generalised:
// B tricycle robots follow a path for T ticks, every tick = nearest path parameter -> N+1 reference poses -> controller
// tick with SQP to convergence -> nominal plant step -> warm-start shift, all enqueued by ONE library call
// (mirrors NMPCNavControlROS::processFollowPath, src/nmpc_nav_control/NMPCNavControlROS.cpp:648-720, closed through the
// plant like scripts/test_scripts/acados_sim_diff.py:119-163).
//
//   g++ -std=c++14 -O2 -I include -I /usr/local/cuda/include examples/fleet_rollout.cpp -L nmpc_nav_control_b200 -lnmpc_b200 \
//       -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/nmpc_nav_control_b200 -o fleet_rollout && ./fleet_rollout 256 200
//
// Exit code 0 = rollout done, 2 = the library (or CUDA) reported an error, e.g. no device: there is no CPU fallback.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime_api.h>
#include "nmpc_b200.h"

#define CHECK(call)                                                               \
    do {                                                                          \
        int rc_ = (call);                                                         \
        if (rc_ != 0) { std::fprintf(stderr, "%s -> %d: %s\n", #call, rc_, nmpc_last_error()); return 2; } \
    } while (0)
#define CU(call)                                                                  \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) { std::fprintf(stderr, "%s: %s\n", #call, cudaGetErrorString(e_)); return 2; } \
    } while (0)

int main(int argc, char** argv)
{
    const int B = argc > 1 ? std::atoi(argv[1]) : 256, T = argc > 2 ? std::atoi(argv[2]) : 200;
    nmpc_dims_t d;
    CHECK(nmpc_dims(NMPC_MODEL_TRIC, &d));
    nmpc_solver* s = nullptr;
    CHECK(nmpc_create(NMPC_MODEL_TRIC, B, /*device*/ 0, &s));       // fails without a CUDA device
    // constructor part of NMPCNavControlTric (Tric.cpp:24-41, 68-73): W from W_diag, W_e from W_diag[0..nx)
    const double W_diag[9] = {10, 10, 5, 0, 0, 0, 0, 1, 1};
    std::vector<double> W((size_t)d.n * d.ny), We(d.nx);
    for (int k = 0; k < d.n; k++) for (int i = 0; i < d.ny; i++) W[(size_t)k * d.ny + i] = W_diag[i];
    for (int i = 0; i < d.nx; i++) We[i] = W_diag[i];
    CHECK(nmpc_set_weights(s, W.data(), We.data()));
    CHECK(nmpc_reset(s));
    CHECK(nmpc_ctrl_reset(s, nullptr));

    // one path: a 2 m line followed by a quarter circle of radius 1 m, 0.4 m/s
    nmpc_path_segment seg[2] = {};
    seg[0].kind = 0; seg[0].vel = 0.4; seg[0].cx[0] = 0.0; seg[0].cx[1] = 2.0; seg[0].cy[0] = 0.0;
    seg[1].kind = 1; seg[1].vel = 0.4; seg[1].cx[0] = 2.0; seg[1].cx[1] = 1.0; seg[1].cx[2] = -M_PI / 2; seg[1].cx[3] = M_PI / 2; seg[1].cy[0] = 1.0;
    const int offs[2] = {0, 2};
    double *d_seg, *d_u, *d_x, *d_pose, *d_vel, *d_steer, *d_traj, *d_cmds;
    int *d_off, *d_fail;
    CU(cudaMalloc((void**)&d_seg, sizeof(seg))); CU(cudaMalloc((void**)&d_off, sizeof(offs)));
    CU(cudaMemcpy(d_seg, seg, sizeof(seg), cudaMemcpyHostToDevice)); CU(cudaMemcpy(d_off, offs, sizeof(offs), cudaMemcpyHostToDevice));
    CU(cudaMalloc((void**)&d_u, sizeof(double) * B)); CU(cudaMalloc((void**)&d_x, sizeof(double) * d.nx * B));
    CU(cudaMalloc((void**)&d_pose, sizeof(double) * 3 * B)); CU(cudaMalloc((void**)&d_vel, sizeof(double) * 3 * B)); CU(cudaMalloc((void**)&d_steer, sizeof(double) * B));
    CU(cudaMalloc((void**)&d_traj, sizeof(double) * (size_t)(T + 1) * 3 * B)); CU(cudaMalloc((void**)&d_cmds, sizeof(double) * (size_t)T * 3 * B));
    CU(cudaMalloc((void**)&d_fail, sizeof(int) * T));
    // robots start at rest beside the path start, a few centimetres apart
    std::vector<double> pose((size_t)3 * B), x((size_t)d.nx * B, 0.0);
    for (int i = 0; i < B; i++) { pose[i] = 0.0; pose[B + i] = 0.05 * std::sin(0.37 * i); pose[2 * B + i] = 0.1 * std::cos(0.11 * i); }
    for (int j = 0; j < 3; j++) for (int i = 0; i < B; i++) x[(size_t)j * B + i] = pose[(size_t)j * B + i];
    CU(cudaMemcpy(d_pose, pose.data(), sizeof(double) * 3 * B, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_x, x.data(), sizeof(double) * d.nx * B, cudaMemcpyHostToDevice));
    CU(cudaMemset(d_vel, 0, sizeof(double) * 3 * B)); CU(cudaMemset(d_steer, 0, sizeof(double) * B)); CU(cudaMemset(d_u, 0, sizeof(double) * B));

    nmpc_rollout_opts o;
    o.dt = 0.025; o.back = 0.05; o.ahead = 0.5; o.is_holonomic = 0;
    o.sqp_max_iter = 6; o.sqp_tol = 1e-8; o.shift = 1;              // BASELINE config 4: SQP to convergence + warm-start shift
    CHECK(nmpc_rollout_device(s, B, T, &o, d_seg, d_off, 1, /*path_id*/ nullptr, d_u, d_x, d_pose, d_vel, d_steer, /*noise*/ nullptr,
                              d_traj, d_cmds, d_fail, /*stream*/ nullptr));
    CU(cudaDeviceSynchronize());
    std::vector<double> traj((size_t)(T + 1) * 3 * B);
    std::vector<int> fail(T);
    CU(cudaMemcpy(traj.data(), d_traj, traj.size() * sizeof(double), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(fail.data(), d_fail, sizeof(int) * T, cudaMemcpyDeviceToHost));
    int nfail = 0;
    for (int t = 0; t < T; t++) nfail += fail[t];
    std::printf("%d robots x %d ticks: robot 0 moved from (%.3f, %.3f) to (%.3f, %.3f), failed solves %d, launches %d\n", B, T,
                traj[0], traj[B], traj[(size_t)T * 3 * B], traj[(size_t)T * 3 * B + B], nfail, nmpc_last_launches(s));
    nmpc_destroy(s);
    return nfail ? 1 : 0;
}
