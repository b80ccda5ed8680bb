// Conformance driver (test infrastructure, our code): drives the reference's UNMODIFIED solver
// wrappers  NMPCNavControl{Diff,Omni4,Tric}  (compiled from /root/reference/src where they lie, never
// copied) against this repository's acados-compatible headers and libraries, the way
// NMPCNavControlROS::executeNMPC does (NMPCNavControlROS.cpp:700-720): construct, reset_mpc(), then
// run(pose, vel, refs, cmd, cpu_time) once per tick.
//
//   usage: conformance_driver <diff|omni4|tric> <input.txt>
//   input: line 1: n_ticks n_ref
//          per tick: pose(3) vel(3) steer(1), then n_ref lines "x y theta"
//   output (stdout), per tick: "tick i ok=<0|1> cpu_ms=<..> cmd=<a> <b> <c>"   (%.17g)
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>
#include <list>

#include "nmpc_nav_control/NMPCNavControlDiff.h"
#include "nmpc_nav_control/NMPCNavControlOmni4.h"
#include "nmpc_nav_control/NMPCNavControlTric.h"

using namespace nmpc_nav_control;

// The wrappers never initialise their yref staging array beyond columns 0..2 (SURVEY.md Appendix C
// #3); placing the object in zeroed storage makes the unspecified entries 0, the value the
// restatement uses.
template <class T, class... A>
static T* make_zeroed(A&&... a)
{
    void* mem = std::calloc(1, sizeof(T));
    return new (mem) T(std::forward<A>(a)...);
}

int main(int argc, char** argv)
{
    if (argc < 3) { std::fprintf(stderr, "usage: %s model input\n", argv[0]); return 2; }
    const std::string model = argv[1];
    FILE* f = std::fopen(argv[2], "r");
    if (!f) { std::perror("input"); return 2; }
    int n_ticks = 0, n_ref = 0;
    if (std::fscanf(f, "%d %d", &n_ticks, &n_ref) != 2) return 2;
    const double dt = 1.0 / 40.0;
    const double deg = M_PI / 180.0;
    NMPCNavControl* ctl = nullptr;
    try {
        // numeric set-up = config/nmpc_nav_control_acados_models.yaml (codegen defaults), W_diag = [Q, R]
        if (model == "diff")
            ctl = make_zeroed<NMPCNavControlDiff>(dt, 0.270, 0.1, 1.0, 2.0, std::vector<double>{10, 10, 5, 0, 0, 0, 0, 1, 1});
        else if (model == "omni4")
            ctl = make_zeroed<NMPCNavControlOmni4>(dt, 0.535, 0.1, 1.0, 1.0,
                                                   std::vector<double>{10, 10, 10, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1});
        else if (model == "tric")
            ctl = make_zeroed<NMPCNavControlTric>(dt, 0.270, 0.1, 0.5, 1.0, 1.0, -30.0 * deg, 30.0 * deg, 120.0 * deg,
                                                  std::vector<double>{10, 10, 5, 0, 0, 0, 0, 1, 1});
        else { std::fprintf(stderr, "unknown model\n"); return 2; }
    } catch (const std::exception& e) {
        std::printf("create_failed %s\n", e.what());
        return 3;
    }
    std::printf("horizon %g dt %.17g\n", ctl->getHorizon(), ctl->getDeltaTime());
    ctl->reset_mpc();
    for (int t = 0; t < n_ticks; t++) {
        NMPCNavControl::Pose pose; NMPCNavControl::Vel vel; double steer = 0.0;
        if (std::fscanf(f, "%lf %lf %lf %lf %lf %lf %lf", &pose.x, &pose.y, &pose.theta, &vel.v, &vel.vn, &vel.w, &steer) != 7) return 2;
        std::list<NMPCNavControl::Pose> refs;
        for (int i = 0; i < n_ref; i++) {
            NMPCNavControl::Pose p;
            if (std::fscanf(f, "%lf %lf %lf", &p.x, &p.y, &p.theta) != 3) return 2;
            refs.push_back(p);
        }
        double cpu_ms = 0.0, a = 0.0, b = 0.0, c = 0.0;
        bool ok = false;
        try {
            if (model == "diff") {
                NMPCNavControlDiff::CmdVelDiff cmd;
                ok = ctl->run(pose, vel, refs, cmd, cpu_ms); a = cmd.v; b = cmd.w;
            } else if (model == "omni4") {
                NMPCNavControlOmni4::CmdVelOmni4 cmd;
                ok = ctl->run(pose, vel, refs, cmd, cpu_ms); a = cmd.v; b = cmd.vn; c = cmd.w;
            } else {
                static_cast<NMPCNavControlTric*>(ctl)->setSteeringWheelAngle(steer);
                NMPCNavControlTric::CmdVelTric cmd;
                ok = ctl->run(pose, vel, refs, cmd, cpu_ms); a = cmd.v; b = cmd.alpha;
            }
        } catch (const std::exception& e) {
            std::printf("tick %d exception %s\n", t, e.what());
            return 4;
        }
        std::printf("tick %d ok=%d cpu_ms=%.6f cmd=%.17g %.17g %.17g\n", t, ok ? 1 : 0, cpu_ms, a, b, c);
    }
    ctl->~NMPCNavControl();
    std::free(ctl);
    std::fclose(f);
    return 0;
}
