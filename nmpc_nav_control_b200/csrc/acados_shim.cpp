// libacados.so — the `ocp_nlp_*` setters / getters the reference's solver wrapper calls
// (include/acados_c/ocp_nlp_interface.h), plus the capsule core shared by the three
// libacados_ocp_solver_<m>.so libraries.  One capsule = one OCP instance = a batch-of-1 solver of
// libnmpc_b200.so; setters stage values on the host, `solve` pushes what changed and runs one
// SQP-RTI iteration on the GPU (nmpc_rti_solve_host), getters read the host copy of the iterate.
//
// Semantics follow what the reference relies on:
//   * after create: tables = code-generation defaults, iterate x_k = (0,0,pi,0..), u = 0
//   * the iterate persists between solves and is never shifted; reset zeroes it
//   * stage-0 "lbx"/"ubx" carry the measured state (nbx_0 = nx, equality)
//   * "W" arrives as a dense column-major matrix; only its diagonal may be non-zero
//   * status: 0 ok, 1 NaN, 2 max iter (not produced by RTI), 3 min step, 4 QP failure
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "acados_shim.h"
#include "../../include/nmpc_b200.h"

struct nmpc_acados_core {
    int model;
    nmpc_dims_t d;
    nmpc_solver* s;
    ocp_nlp_config cfg;
    ocp_nlp_dims dims;
    ocp_nlp_in in;
    ocp_nlp_out out;
    ocp_nlp_solver solver;
    // host staging (what the setters write)
    std::vector<double> W, We, lbx, ubx, lbu, ubu, p;      // tables, layout of nmpc_set_*
    std::vector<double> lbx0, ubx0, yref;                   // stage-0 state bounds; yref [N+1][ny]
    bool w_dirty, b_dirty, p_dirty;
    std::vector<unsigned char> w_bad;                       // per stage: the last W set for it was not diagonal
    // host copy of the iterate after the last solve / reset / create; after a solve only x_0, x_1 and u_0 are current
    // (what run() reads, Diff.cpp:145-152), the rest is fetched from the device on the first request for another stage
    std::vector<double> x, u;
    bool iter_stale;
    double time_tot;
    int qp_iter, status;
};

static void fetch_iterate(nmpc_acados_core* c)
{
    nmpc_get_iterate_host(c->s, 1, c->x.data(), c->u.data());
    c->iter_stale = false;
}

extern "C" nmpc_acados_core* nmpc_acados_core_create(int model)
{
    nmpc_dims_t d;
    if (nmpc_dims(model, &d) != 0) return nullptr;
    nmpc_acados_core* c = new (std::nothrow) nmpc_acados_core();
    if (!c) return nullptr;
    c->model = model; c->d = d; c->s = nullptr;
    int device = 0;
    if (const char* e = std::getenv("NMPC_ACADOS_DEVICE")) device = std::atoi(e);
    if (nmpc_create(model, 1, device, &c->s) != 0) { delete c; return nullptr; }   // no device -> create fails, no CPU fallback
    const int n = d.n;
    c->W.resize((size_t)n * d.ny); c->We.resize(d.nx);
    c->lbx.resize((size_t)n * d.nbx); c->ubx.resize((size_t)n * d.nbx);
    c->lbu.resize((size_t)n * d.nbu); c->ubu.resize((size_t)n * d.nbu);
    c->p.resize((size_t)n * d.np);
    nmpc_get_tables(c->s, c->W.data(), c->We.data(), c->lbx.data(), c->ubx.data(), c->lbu.data(), c->ubu.data(), c->p.data());
    c->x.assign((size_t)(n + 1) * d.nx, 0.0); c->u.assign((size_t)n * d.nu, 0.0);
    fetch_iterate(c);
    c->lbx0.assign(c->x.begin(), c->x.begin() + d.nx);      // constraints.x0 default (generate_c_code.py:58-60)
    c->ubx0 = c->lbx0;
    c->yref.assign((size_t)(n + 1) * d.ny, 0.0);
    c->w_dirty = c->b_dirty = c->p_dirty = false;
    c->w_bad.assign((size_t)n + 1, 0);
    c->time_tot = 0.0; c->qp_iter = 0; c->status = 0; c->iter_stale = false;
    c->cfg.core = c; c->in.core = c; c->out.core = c; c->out.inf_norm_res = 0.0; c->solver.core = c;
    c->dims.core = c; c->dims.N = n; c->dims.nx = d.nx; c->dims.nu = d.nu; c->dims.ny = d.ny; c->dims.nyn = d.nyn;
    c->dims.np = d.np; c->dims.nbx = d.nbx; c->dims.nbu = d.nbu;
    return c;
}

extern "C" void nmpc_acados_core_views(nmpc_acados_core* c, ocp_nlp_config** cfg, ocp_nlp_dims** dims, ocp_nlp_in** in,
                                       ocp_nlp_out** out, ocp_nlp_solver** solver)
{
    *cfg = &c->cfg; *dims = &c->dims; *in = &c->in; *out = &c->out; *solver = &c->solver;
}

extern "C" void nmpc_acados_core_free(nmpc_acados_core* c)
{
    if (!c) return;
    nmpc_destroy(c->s);
    delete c;
}

extern "C" int nmpc_acados_core_update_params(nmpc_acados_core* c, int stage, const double* p, int np)
{
    if (!p || np != c->d.np || stage < 0 || stage > c->d.n) return 1;
    if (stage == c->d.n) return 0;        // the terminal cost does not depend on p
    std::memcpy(&c->p[(size_t)stage * np], p, sizeof(double) * np);
    c->p_dirty = true;
    return 0;
}

extern "C" int nmpc_acados_core_reset(nmpc_acados_core* c, int)
{
    if (nmpc_reset(c->s) != 0) return 1;
    std::fill(c->x.begin(), c->x.end(), 0.0);
    std::fill(c->u.begin(), c->u.end(), 0.0);
    c->iter_stale = false;
    return 0;
}

extern "C" int nmpc_acados_core_solve(nmpc_acados_core* c)
{
    const auto t0 = std::chrono::steady_clock::now();
    const nmpc_dims_t& d = c->d;
    for (unsigned char b : c->w_bad)
        if (b) return NMPC_QP_FAILURE;                          // a stage holds a non-diagonal W
    for (int j = 0; j < d.nx; j++)
        if (c->lbx0[j] != c->ubx0[j]) return NMPC_QP_FAILURE;   // stage 0 must be the x0 equality
    // the terminal weight changes every tick in the diff wrapper (Diff.cpp:126-139): it travels as the per-instance
    // W_e argument of the solve, so only a change of a path-stage W touches the device tables
    if (c->w_dirty && nmpc_set_weights(c->s, c->W.data(), nullptr) != 0) return NMPC_QP_FAILURE;
    if (c->b_dirty && nmpc_set_bounds(c->s, c->lbx.data(), c->ubx.data(), c->lbu.data(), c->ubu.data()) != 0) return NMPC_QP_FAILURE;
    if (c->p_dirty && nmpc_set_params(c->s, c->p.data()) != 0) return NMPC_QP_FAILURE;
    c->w_dirty = c->b_dirty = c->p_dirty = false;
    double u0[4], x1[11];
    int status = 0, qp_iter = 0;
    if (nmpc_rti_solve_host(c->s, 1, c->lbx0.data(), c->yref.data(), d.ny, c->We.data(), u0, x1, &status, &qp_iter) != 0)
        return NMPC_QP_FAILURE;
    if (status == 0) {
        std::memcpy(c->x.data(), c->lbx0.data(), sizeof(double) * d.nx);       // K4: x_0 <- the measured state
        std::memcpy(c->x.data() + d.nx, x1, sizeof(double) * d.nx);
        std::memcpy(c->u.data(), u0, sizeof(double) * d.nu);
    }
    c->iter_stale = true;
    double stats[8];
    c->out.inf_norm_res = 0.0;
    if (nmpc_last_stats_host(c->s, 1, stats) == 0)
        c->out.inf_norm_res = std::fmax(std::fmax(stats[0], stats[1]), std::fmax(stats[2], stats[3]));
    c->status = status; c->qp_iter = qp_iter;
    c->time_tot = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return status;
}

static bool is(const char* a, const char* b) { return std::strcmp(a, b) == 0; }

extern "C" int ocp_nlp_constraints_model_set(ocp_nlp_config*, ocp_nlp_dims*, ocp_nlp_in* in, ocp_nlp_out*, int stage,
                                             const char* field, void* value)
{
    if (!in || !in->core || !field || !value) return 1;
    nmpc_acados_core* c = in->core;
    const nmpc_dims_t& d = c->d;
    const double* v = (const double*)value;
    if (is(field, "lbx") || is(field, "ubx")) {
        const bool lo = field[0] == 'l';
        if (stage == 0) { std::memcpy((lo ? c->lbx0 : c->ubx0).data(), v, sizeof(double) * d.nx); return 0; }
        if (stage < 1 || stage > d.n) return 1;
        std::memcpy(&(lo ? c->lbx : c->ubx)[(size_t)(stage - 1) * d.nbx], v, sizeof(double) * d.nbx);
        c->b_dirty = true;
        return 0;
    }
    if (is(field, "lbu") || is(field, "ubu")) {
        if (stage < 0 || stage >= d.n) return 1;
        std::memcpy(&(field[0] == 'l' ? c->lbu : c->ubu)[(size_t)stage * d.nbu], v, sizeof(double) * d.nbu);
        c->b_dirty = true;
        return 0;
    }
    return 1;
}

extern "C" int ocp_nlp_cost_model_set(ocp_nlp_config*, ocp_nlp_dims*, ocp_nlp_in* in, int stage, const char* field, void* value)
{
    if (!in || !in->core || !field || !value) return 1;
    nmpc_acados_core* c = in->core;
    const nmpc_dims_t& d = c->d;
    const double* v = (const double*)value;
    if (stage < 0 || stage > d.n) return 1;
    const int ny = stage < d.n ? d.ny : d.nyn;
    if (is(field, "W")) {
        double* dst = stage < d.n ? &c->W[(size_t)stage * d.ny] : c->We.data();
        bool bad = false;
        for (int j = 0; j < ny && !bad; j++)
            for (int i = 0; i < ny; i++)
                if (i != j && v[i + (size_t)ny * j] != 0.0) { bad = true; break; }
        c->w_bad[stage] = bad ? 1 : 0;                           // a later diagonal W for the stage clears it
        if (bad) return 1;                                       // nothing of a rejected W is stored
        for (int j = 0; j < ny; j++) dst[j] = v[j + (size_t)ny * j];
        if (stage < d.n) c->w_dirty = true;
        return 0;
    }
    if (is(field, "yref") || is(field, "y_ref")) {
        double* dst = &c->yref[(size_t)stage * d.ny];
        std::memcpy(dst, v, sizeof(double) * ny);
        return 0;
    }
    return 1;
}

extern "C" void ocp_nlp_out_get(ocp_nlp_config*, ocp_nlp_dims*, ocp_nlp_out* out, int stage, const char* field, void* value)
{
    if (!out || !out->core || !field || !value) return;
    nmpc_acados_core* c = out->core;
    const nmpc_dims_t& d = c->d;
    if (is(field, "x") && stage >= 0 && stage <= d.n) {
        if (c->iter_stale && (stage > 1 || c->status != 0)) fetch_iterate(c);
        std::memcpy(value, &c->x[(size_t)stage * d.nx], sizeof(double) * d.nx);
    } else if (is(field, "u") && stage >= 0 && stage < d.n) {
        if (c->iter_stale && (stage > 0 || c->status != 0)) fetch_iterate(c);
        std::memcpy(value, &c->u[(size_t)stage * d.nu], sizeof(double) * d.nu);
    }
}

extern "C" void ocp_nlp_get(ocp_nlp_solver* solver, const char* field, void* return_value_)
{
    if (!solver || !solver->core || !field || !return_value_) return;
    nmpc_acados_core* c = solver->core;
    if (is(field, "time_tot")) *(double*)return_value_ = c->time_tot;
    else if (is(field, "sqp_iter")) *(int*)return_value_ = 1;
    else if (is(field, "qp_iter")) *(int*)return_value_ = c->qp_iter;
    else if (is(field, "status")) *(int*)return_value_ = c->status;
}
