// libacados_ocp_solver_tric3amr.so — the generated-solver entry points of the reference for the tric3amr
// model (declared in include/acados_solver_tric3amr.h), implemented on the B200-native batched solver
// with batch = 1 through the shared capsule code in acados_shim.cpp.
#include "acados_solver_tric3amr.h"
#include "acados_shim.h"

extern "C" {
tric3amr_solver_capsule* tric3amr_acados_create_capsule(void)
{
    tric3amr_solver_capsule* c = new (std::nothrow) tric3amr_solver_capsule();
    if (c) { c->nlp_config = nullptr; c->nlp_dims = nullptr; c->nlp_in = nullptr; c->nlp_out = nullptr; c->nlp_solver = nullptr; c->core = nullptr; }
    return c;
}
int tric3amr_acados_free_capsule(tric3amr_solver_capsule* c) { delete c; return 0; }
int tric3amr_acados_create(tric3amr_solver_capsule* c)
{
    if (!c) return 1;
    nmpc_acados_core* core = nmpc_acados_core_create(2);
    if (!core) return 1;
    c->core = core;
    nmpc_acados_core_views(core, &c->nlp_config, &c->nlp_dims, &c->nlp_in, &c->nlp_out, &c->nlp_solver);
    return 0;
}
int tric3amr_acados_update_params(tric3amr_solver_capsule* c, int stage, double* p, int np)
{
    return (c && c->core) ? nmpc_acados_core_update_params(c->core, stage, p, np) : 1;
}
int tric3amr_acados_solve(tric3amr_solver_capsule* c) { return (c && c->core) ? nmpc_acados_core_solve(c->core) : 1; }
int tric3amr_acados_reset(tric3amr_solver_capsule* c, int reset_qp_solver_mem)
{
    return (c && c->core) ? nmpc_acados_core_reset(c->core, reset_qp_solver_mem) : 1;
}
int tric3amr_acados_free(tric3amr_solver_capsule* c)
{
    if (c && c->core) { nmpc_acados_core_free(c->core); c->core = nullptr; }
    return 0;
}
}
