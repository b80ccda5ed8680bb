// TEST INFRASTRUCTURE ONLY.  Lane-by-lane CPU emulation of the CUDA per-lane logic in
// nmpc_nav_control_b200/csrc/rti_core.cuh, compiled by g++ with -DNMPC_HOST_EMUL.  It exists so
// the kernels' arithmetic can be checked against the oracle in the GPU-less build container;
// it is never linked into the product library and the Python package cannot load it.
#include <vector>
#include <cstring>
#include <cstdlib>
#include "../../nmpc_nav_control_b200/csrc/rti_core.cuh"
#include "../../nmpc_nav_control_b200/csrc/rti_coop.cuh"
#include "../../nmpc_nav_control_b200/csrc/rti_solo.cuh"

using namespace nmpc;

template <class M>
static int run(int B, const double* W, const double* We, const double* lbx, const double* ubx,
               const double* lbu, const double* ubu, const double* p, double dt, const IpmOpts* o,
               const double* x0bar, const double* yref, int nyref, const double* We_inst,
               double* x, double* u, int* status, int* iters, double* stats, bool solo = false)
{
    using S = Rti<M>;
    using R = typename S::R;
    constexpr int NX = S::NX, NU = S::NU, NV = S::NV;
    std::vector<double> lti(NSTAGE * 4 * NV), thr(NSTAGE * S::NC);
    for (int k = 0; k < NSTAGE; k++) {
        double x0[NX] = {0}, u0[NU] = {0}, xn[NX], Ep[3][S::NC];
        S::rk4_sens(x0, u0, p + k * S::NP, dt, xn, Ep, &lti[k * 4 * NV]);
        for (int c = 0; c < S::NC; c++) thr[k * S::NC + c] = Ep[2][c];
    }
    Tables tb{W, We, lbx, ubx, lbu, ubu, p, lti.data(), dt, thr.data(), nullptr};
    std::vector<double> tile(R::tile_doubles);
    for (int i = 0; i < B; i++) {
        std::fill(tile.begin(), tile.end(), 0.0);
        const int lane = i % LANES;
        double* base = tile.data() + lane;
        double* xi = x + (size_t)i * (NSTAGE + 1) * NX;
        double* ui = u + (size_t)i * NSTAGE * NU;
        const double* yi = yref + (size_t)i * (NSTAGE + 1) * nyref;
        const double* wei = We_inst ? We_inst + (size_t)i * NX : We;
        for (int k = 0; k <= NSTAGE; k++)
            S::linearize_stage(k, xi + k * NX, ui + (k < NSTAGE ? k : 0) * NU, xi + (k < NSTAGE ? k + 1 : k) * NX,
                               yi + k * nyref, nyref, x0bar + (size_t)i * NX, tb, wei,
                               base + R::OFF_LIN + (size_t)k * R::NF_LIN * LANES, base + R::OFF_IT + (size_t)k * R::NF_IT * LANES);
        typename S::LaneStats st;
        if (solo) {
            // the block-per-instance mapping (rti_solo.cuh) run by one "thread": its phases execute item by item
            std::vector<double> sm(Solo<M>::SM_DOUBLES, 0.0), gs(Solo<M>::GSCR_DOUBLES + 1, 0.0);
            Solo<M>::run(sm.data(), typename Solo<M>::TileIO{base}, tb, wei, 1, *o, &st, gs.data());
        } else
            S::qp_ipm_lane(base, tb, wei, *o, st);
        status[i] = st.status; iters[i] = st.iter;
        if (stats) { for (int q = 0; q < 4; q++) stats[i * 8 + q] = st.res[q]; stats[i * 8 + 4] = st.mu; stats[i * 8 + 5] = st.lin_res; stats[i * 8 + 6] = st.cond_fallbacks; }
        if (st.status == 0 || st.status == 1)
            for (int k = 0; k <= NSTAGE; k++)
                S::step_stage(k, base + R::OFF_IT + (size_t)k * R::NF_IT * LANES, x0bar + (size_t)i * NX, xi + k * NX, ui + (k < NSTAGE ? k : 0) * NU);
    }
    return 0;
}

extern "C" int emul_rti(int model, int B, const double* W, const double* We, const double* lbx, const double* ubx,
                        const double* lbu, const double* ubu, const double* p, double dt, const IpmOpts* o,
                        const double* x0bar, const double* yref, int nyref, const double* We_inst,
                        double* x, double* u, int* status, int* iters, double* stats)
{
    switch (model) {
        case 0: return run<DiffModel>(B, W, We, lbx, ubx, lbu, ubu, p, dt, o, x0bar, yref, nyref, We_inst, x, u, status, iters, stats);
        case 1: return run<Omni4Model>(B, W, We, lbx, ubx, lbu, ubu, p, dt, o, x0bar, yref, nyref, We_inst, x, u, status, iters, stats);
        case 2: return run<TricModel>(B, W, We, lbx, ubx, lbu, ubu, p, dt, o, x0bar, yref, nyref, We_inst, x, u, status, iters, stats);
    }
    return -1;
}

extern "C" int emul_rti_solo(int model, int B, const double* W, const double* We, const double* lbx, const double* ubx,
                             const double* lbu, const double* ubu, const double* p, double dt, const IpmOpts* o,
                             const double* x0bar, const double* yref, int nyref, const double* We_inst,
                             double* x, double* u, int* status, int* iters, double* stats)
{
    switch (model) {
        case 0: return run<DiffModel>(B, W, We, lbx, ubx, lbu, ubu, p, dt, o, x0bar, yref, nyref, We_inst, x, u, status, iters, stats, true);
        case 1: return run<Omni4Model>(B, W, We, lbx, ubx, lbu, ubu, p, dt, o, x0bar, yref, nyref, We_inst, x, u, status, iters, stats, true);
        case 2: return run<TricModel>(B, W, We, lbx, ubx, lbu, ubu, p, dt, o, x0bar, yref, nyref, We_inst, x, u, status, iters, stats, true);
    }
    return -1;
}

// ---- the persistent lane-cooperative K3 (rti_coop.cuh): one emulated warp of 32 lanes, phases run lane by lane ----
template <class M, class GP>
static int run_group(int B, const double* W, const double* We, const double* lbx, const double* ubx,
                     const double* lbu, const double* ubu, const double* p, double dt, const IpmOpts* o,
                     const double* x0bar, const double* yref, int nyref, const double* We_inst,
                     double* x, double* u, int* status, int* iters, double* stats)
{
    using S = Rti<M>;
    using GR = typename GP::R;
    constexpr int NX = S::NX, NU = S::NU, NV = S::NV;
    std::vector<double> lti(NSTAGE * 4 * NV), thr(NSTAGE * S::NC);
    for (int k = 0; k < NSTAGE; k++) {
        double x0[NX] = {0}, u0[NU] = {0}, xn[NX], Ep[3][S::NC];
        S::rk4_sens(x0, u0, p + k * S::NP, dt, xn, Ep, &lti[k * 4 * NV]);
        for (int c = 0; c < S::NC; c++) thr[k * S::NC + c] = Ep[2][c];
    }
    std::vector<double> stg((NSTAGE + 1) * GP::TROW, 0.0);
    for (int k = 0; k < NSTAGE; k++) {
        for (int i = 0; i < 4 * NV; i++) stg[k * GP::TROW + i] = lti[k * 4 * NV + i];
        stg[k * GP::TROW + GP::LT_ONE] = 1.0;
        for (int i = 0; i < S::NY; i++) stg[k * GP::TROW + GP::T_W + i] = W[k * S::NY + i];
    }
    Tables tb{W, We, lbx, ubx, lbu, ubu, p, lti.data(), dt, thr.data(), stg.data()};
    std::vector<double> ws((size_t)B * GR::inst_doubles, 0.0);
    std::vector<double> WeT;                  // kernel layout of the per-instance terminal weights: [nx][B]
    if (We_inst) { WeT.resize((size_t)NX * B); for (int i = 0; i < B; i++) for (int j = 0; j < NX; j++) WeT[(size_t)j * B + i] = We_inst[(size_t)i * NX + j]; }
    for (int i = 0; i < B; i++) {
        double* xi = x + (size_t)i * (NSTAGE + 1) * NX;
        double* ui = u + (size_t)i * NSTAGE * NU;
        const double* yi = yref + (size_t)i * (NSTAGE + 1) * nyref;
        const double* wei = We_inst ? We_inst + (size_t)i * NX : We;
        for (int k = 0; k <= NSTAGE; k++) {
            double* rec = GP::rec_of(ws.data(), i, k);
            S::template linearize_stage<GR, 1>(k, xi + k * NX, ui + (k < NSTAGE ? k : 0) * NU, xi + (k < NSTAGE ? k + 1 : k) * NX,
                                               yi + k * nyref, nyref, x0bar + (size_t)i * NX, tb, wei, rec, rec);
            S::template coldstart_stage<GR, 1>(k, *o, rec, rec);
        }
    }
    std::vector<int> qs(B, -1), qi(B, 0);
    std::vector<double> st((size_t)8 * B, 0.0);
    GrpOut out{qs.data(), qi.data(), st.data(), B};
    int next = 0;
    const int nwarps = 2;                     // the second warp finds the queue empty unless B is large
    for (int wp = 0; wp < nwarps; wp++) {
        std::vector<typename GP::Lane> lanes(32);
        std::vector<double> sm(GP::WARP_D, 0.0);
        for (int l = 0; l < 32; l++) GP::init_lane(lanes[l], l, 0);
        GP::run_warp(lanes.data(), sm.data(), ws.data(), 0, B, &next, tb, We_inst ? WeT.data() : nullptr, B, *o, out, GrpResume{nullptr, nullptr, nullptr});
    }
    for (int i = 0; i < B; i++) {
        status[i] = qs[i]; iters[i] = qi[i];
        if (stats) for (int q = 0; q < 7; q++) stats[i * 8 + q] = st[(size_t)q * B + i];
        double* xi = x + (size_t)i * (NSTAGE + 1) * NX;
        double* ui = u + (size_t)i * NSTAGE * NU;
        if (qs[i] == 0 || qs[i] == 1)
            for (int k = 0; k <= NSTAGE; k++)
                S::template step_stage<GR, 1>(k, GP::rec_of(ws.data(), i, k), x0bar + (size_t)i * NX, xi + k * NX, ui + (k < NSTAGE ? k : 0) * NU);
    }
    return 0;
}

extern "C" int emul_rti_coop(int model, int G, int B, const double* W, const double* We, const double* lbx, const double* ubx,
                             const double* lbu, const double* ubu, const double* p, double dt, const IpmOpts* o,
                             const double* x0bar, const double* yref, int nyref, const double* We_inst,
                             double* x, double* u, int* status, int* iters, double* stats)
{
#define RC(MODEL, GG) return run_group<MODEL, Coop<MODEL, GG>>(B, W, We, lbx, ubx, lbu, ubu, p, dt, o, x0bar, yref, nyref, We_inst, x, u, status, iters, stats)
    (void)G;
    if (model == 0) RC(DiffModel, 8);
    if (model == 1) RC(Omni4Model, 16);
    if (model == 2) RC(TricModel, 8);
#undef RC
    return -1;
}

// ---- hybrid schedule: K iterations of the per-lane sweeps, then the unfinished instances are handed to the
// lane-cooperative kernel in the middle of an iteration (what the hybrid schedule does on the device) ----
template <class M, class GP>
static int run_hybrid(int K, int B, const double* W, const double* We, const double* lbx, const double* ubx,
                      const double* lbu, const double* ubu, const double* p, double dt, const IpmOpts* o,
                      const double* x0bar, const double* yref, int nyref, const double* We_inst,
                      double* x, double* u, int* status, int* iters, double* stats)
{
    using S = Rti<M>;
    using R = typename S::R;
    using GR = typename GP::R;
    constexpr int NX = S::NX, NU = S::NU, NV = S::NV;
    std::vector<double> lti(NSTAGE * 4 * NV), thr(NSTAGE * S::NC);
    for (int k = 0; k < NSTAGE; k++) {
        double x0[NX] = {0}, u0[NU] = {0}, xn[NX], Ep[3][S::NC];
        S::rk4_sens(x0, u0, p + k * S::NP, dt, xn, Ep, &lti[k * 4 * NV]);
        for (int c = 0; c < S::NC; c++) thr[k * S::NC + c] = Ep[2][c];
    }
    std::vector<double> stg((NSTAGE + 1) * GP::TROW, 0.0);
    for (int k = 0; k < NSTAGE; k++) {
        for (int i = 0; i < 4 * NV; i++) stg[k * GP::TROW + i] = lti[k * 4 * NV + i];
        stg[k * GP::TROW + GP::LT_ONE] = 1.0;
        for (int i = 0; i < S::NY; i++) stg[k * GP::TROW + GP::T_W + i] = W[k * S::NY + i];
    }
    Tables tb{W, We, lbx, ubx, lbu, ubu, p, lti.data(), dt, thr.data(), stg.data()};
    std::vector<std::vector<double>> tiles(B, std::vector<double>(R::tile_doubles, 0.0));
    std::vector<typename S::LaneCtl> ctl(B);
    std::vector<int> qs(B, -1), qi(B, 0), list;
    std::vector<double> st((size_t)8 * B, 0.0);
    std::vector<double> WeT;
    if (We_inst) { WeT.resize((size_t)NX * B); for (int i = 0; i < B; i++) for (int j = 0; j < NX; j++) WeT[(size_t)j * B + i] = We_inst[(size_t)i * NX + j]; }
    std::vector<typename S::LaneCtl> ctl_g;
    for (int i = 0; i < B; i++) {
        double* base = tiles[i].data() + (i % LANES);
        double* xi = x + (size_t)i * (NSTAGE + 1) * NX;
        double* ui = u + (size_t)i * NSTAGE * NU;
        const double* yi = yref + (size_t)i * (NSTAGE + 1) * nyref;
        const double* wei = We_inst ? We_inst + (size_t)i * NX : We;
        for (int k = 0; k <= NSTAGE; k++)
            S::linearize_stage(k, xi + k * NX, ui + (k < NSTAGE ? k : 0) * NU, xi + (k < NSTAGE ? k + 1 : k) * NX,
                               yi + k * nyref, nyref, x0bar + (size_t)i * NX, tb, wei,
                               base + R::OFF_LIN + (size_t)k * R::NF_LIN * LANES, base + R::OFF_IT + (size_t)k * R::NF_IT * LANES);
        typename S::LaneCtl& c = ctl[i];
        c.init(true);
        double scratch[S::CarryB::SC_N * S::PSTRIDE];
        S::template run_phase<S::SW_B_FIRST>(base, tb, wei, *o, false, c, scratch);
        for (int it = 0; it < K && !c.done; it++) {
            S::template run_phase<S::SW_FDF>(base, tb, wei, *o, true, c, scratch);      // true: defer a centering repeat (done = 2)
            S::template run_phase<S::SW_B>(base, tb, wei, *o, false, c, scratch);
        }
        if (c.done == 2) c.done = 0;
        if (c.done) {
            qs[i] = c.status; qi[i] = c.iter;
            for (int q = 0; q < 4; q++) st[(size_t)q * B + i] = c.nrm[q];
            st[(size_t)4 * B + i] = c.mu; st[(size_t)5 * B + i] = c.lin_res; st[(size_t)6 * B + i] = c.nfb;
        } else { list.push_back(i); ctl_g.push_back(c); }
    }
    const int nres = (int)list.size();
    std::vector<double> ws((size_t)std::max(nres, 1) * GR::inst_doubles, 0.0);
    for (int q = 0; q < nres; q++)
        for (int k = 0; k <= NSTAGE; k++)
            GP::tile_to_record(tiles[list[q]].data() + (list[q] % LANES), k, GP::rec_of(ws.data(), q, k), thr.data() + (k < NSTAGE ? k : 0) * S::NC);
    GrpOut out{qs.data(), qi.data(), st.data(), B};
    int next = 0;
    {
        std::vector<typename GP::Lane> lanes(32);
        std::vector<double> sm(GP::WARP_D, 0.0);
        for (int l = 0; l < 32; l++) GP::init_lane(lanes[l], l, 0);
        GP::run_warp(lanes.data(), sm.data(), ws.data(), 0, 0, &next, tb, We_inst ? WeT.data() : nullptr, B, *o, out,
                     GrpResume{&nres, list.data(), ctl_g.data()});
    }
    std::vector<int> map(B, -1);
    for (int q = 0; q < nres; q++) map[list[q]] = q;
    for (int i = 0; i < B; i++) {
        status[i] = qs[i]; iters[i] = qi[i];
        if (stats) for (int q = 0; q < 7; q++) stats[i * 8 + q] = st[(size_t)q * B + i];
        double* xi = x + (size_t)i * (NSTAGE + 1) * NX;
        double* ui = u + (size_t)i * NSTAGE * NU;
        if (qs[i] == 0 || qs[i] == 1)
            for (int k = 0; k <= NSTAGE; k++) {
                if (map[i] >= 0) S::template step_stage<GR, 1>(k, GP::rec_of(ws.data(), map[i], k), x0bar + (size_t)i * NX, xi + k * NX, ui + (k < NSTAGE ? k : 0) * NU);
                else S::step_stage(k, tiles[i].data() + (i % LANES) + R::OFF_IT + (size_t)k * R::NF_IT * LANES, x0bar + (size_t)i * NX, xi + k * NX, ui + (k < NSTAGE ? k : 0) * NU);
            }
    }
    return nres;
}

extern "C" int emul_rti_hybrid(int model, int K, int B, const double* W, const double* We, const double* lbx, const double* ubx,
                               const double* lbu, const double* ubu, const double* p, double dt, const IpmOpts* o,
                               const double* x0bar, const double* yref, int nyref, const double* We_inst,
                               double* x, double* u, int* status, int* iters, double* stats)
{
#define RH(MODEL, GG) return run_hybrid<MODEL, Coop<MODEL, GG>>(K, B, W, We, lbx, ubx, lbu, ubu, p, dt, o, x0bar, yref, nyref, We_inst, x, u, status, iters, stats)
    if (model == 0) RH(DiffModel, 8);
    if (model == 1) RH(Omni4Model, 16);
    if (model == 2) RH(TricModel, 8);
#undef RH
    return -1;
}

// ---- controller glue (ctrl_glue.cuh), one instance after another; arrays SoA with leading dimension B ----
#include "../../nmpc_nav_control_b200/csrc/ctrl_glue.cuh"

template <class M>
static void ctrl_pre_all(int B, const double* pose, const double* vel, const double* steer, const double* refs, const int* nref,
                         int nref_max, const double* vref, const double* p, const double* W0, const double* We_tab,
                         double* x0bar, double* yref, double* We)
{
    for (int i = 0; i < B; i++) {
        int n = nref ? nref[i] : nref_max;
        if (n > nref_max) n = nref_max;
        CtrlGlue<M>::pre(pose + i, vel + i, steer ? steer + i : nullptr, refs + i, n, vref + i, (size_t)B, p, W0, We_tab,
                         x0bar + i, yref + i, We ? We + i : nullptr, (size_t)B);
    }
}
template <class M>
static void ctrl_post_all(int B, const int* status, const double* x0bar, const double* u0, double dt, const double* p,
                          double* vref, double* cmd)
{
    for (int i = 0; i < B; i++)
        CtrlGlue<M>::post(status[i], x0bar + i, (size_t)B, u0 + i, (size_t)B, dt, p, vref + i, (size_t)B, cmd + i);
}

extern "C" int emul_ctrl_pre(int model, int B, const double* pose, const double* vel, const double* steer, const double* refs,
                             const int* nref, int nref_max, const double* vref, const double* p, const double* W0,
                             const double* We_tab, double* x0bar, double* yref, double* We)
{
    switch (model) {
        case 0: ctrl_pre_all<DiffModel>(B, pose, vel, steer, refs, nref, nref_max, vref, p, W0, We_tab, x0bar, yref, We); break;
        case 1: ctrl_pre_all<Omni4Model>(B, pose, vel, steer, refs, nref, nref_max, vref, p, W0, We_tab, x0bar, yref, We); break;
        default: ctrl_pre_all<TricModel>(B, pose, vel, steer, refs, nref, nref_max, vref, p, W0, We_tab, x0bar, yref, We); break;
    }
    return 0;
}
extern "C" int emul_ctrl_post(int model, int B, const int* status, const double* x0bar, const double* u0, double dt,
                              const double* p, double* vref, double* cmd)
{
    switch (model) {
        case 0: ctrl_post_all<DiffModel>(B, status, x0bar, u0, dt, p, vref, cmd); break;
        case 1: ctrl_post_all<Omni4Model>(B, status, x0bar, u0, dt, p, vref, cmd); break;
        default: ctrl_post_all<TricModel>(B, status, x0bar, u0, dt, p, vref, cmd); break;
    }
    return 0;
}

// ---- path discretisation (path_disc.cuh), one robot after another; poses SoA [num_poses][3][B] ----
#include "../../nmpc_nav_control_b200/csrc/path_disc.cuh"
extern "C" int emul_path_discretize(int B, const double* segs, const int* path_off, int n_paths, const int* path_id, const double* u0,
                                    double period, int num_poses, int holonomic, double* out)
{
    for (int i = 0; i < B; i++) {
        int p = path_id ? path_id[i] : 0;
        p = p < 0 ? 0 : (p >= n_paths ? n_paths - 1 : p);
        PathDisc::next_poses(segs + (size_t)path_off[p] * PathDisc::SEG, path_off[p + 1] - path_off[p], u0[i], period, num_poses,
                             holonomic != 0, out + i, (size_t)B);
    }
    return 0;
}

// ---- closed-loop pieces (rollout.cuh) ----
#include "../../nmpc_nav_control_b200/csrc/rollout.cuh"
extern "C" int emul_plant_step(int model, int B, double* xp, const double* u0, const double* noise, const double* p, double dt,
                               double* pose, double* vel, double* steer)
{
    for (int i = 0; i < B; i++) {
        const double* nz = noise ? noise + i : nullptr;
        switch (model) {
            case 0: Rollout<DiffModel>::plant_step(xp + i, B, u0 + i, B, nz, p, dt, pose + i, vel + i, steer + i); break;
            case 1: Rollout<Omni4Model>::plant_step(xp + i, B, u0 + i, B, nz, p, dt, pose + i, vel + i, steer + i); break;
            default: Rollout<TricModel>::plant_step(xp + i, B, u0 + i, B, nz, p, dt, pose + i, vel + i, steer + i); break;
        }
    }
    return 0;
}
extern "C" double emul_nearest_u(const double* segs, int nseg, double u_prev, double px, double py, double back, double ahead)
{
    return PathNearest::nearest_u(segs, nseg, u_prev, px, py, back, ahead);
}
