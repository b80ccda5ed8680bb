// Internal interface between libacados.so (acados_shim.cpp) and the three
// libacados_ocp_solver_<m>.so entry-point libraries.  Host C++ only; no CUDA types.
#pragma once
#include <new>
#include "acados_c/ocp_nlp_interface.h"

extern "C" {
nmpc_acados_core* nmpc_acados_core_create(int model);
void nmpc_acados_core_views(nmpc_acados_core* c, ocp_nlp_config** cfg, ocp_nlp_dims** dims, ocp_nlp_in** in,
                            ocp_nlp_out** out, ocp_nlp_solver** solver);
int nmpc_acados_core_update_params(nmpc_acados_core* c, int stage, const double* p, int np);
int nmpc_acados_core_solve(nmpc_acados_core* c);
int nmpc_acados_core_reset(nmpc_acados_core* c, int reset_qp_solver_mem);
void nmpc_acados_core_free(nmpc_acados_core* c);
}
