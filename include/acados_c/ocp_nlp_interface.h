/*
 * acados_c/ocp_nlp_interface.h — the slice of acados' C interface that the reference's solver
 * wrapper binds (src/nmpc_nav_control/NMPCNavControl{Diff,Omni4,Tric}.cpp), re-exported by the
 * B200-native solver.  Source compatible, not binary compatible with acados: the wrapper is
 * recompiled against these headers and linked with libacados.so + libacados_ocp_solver_<m>.so
 * from this repository (CMakeLists.txt:108-115 of the reference names the same libraries).
 *
 * Call sites replaced (reference file:line):
 *   ocp_nlp_constraints_model_set  NMPCNavControlDiff.cpp:49-65 (set-up), :96-101 (x0 per tick)
 *   ocp_nlp_cost_model_set         NMPCNavControlDiff.cpp:68-73 (W, W_e), :121-124 (yref), :138-139 (W_e per tick)
 *   ocp_nlp_out_get                NMPCNavControlDiff.cpp:151-152 (u_0), :168-169 (x_1)
 *   ocp_nlp_get                    NMPCNavControlDiff.cpp:148 ("time_tot", seconds)
 *   nlp_out->inf_norm_res          NMPCNavControlDiff.cpp:146
 * All pointers are caller-owned host buffers and are copied synchronously.  Nothing throws.
 */
#ifndef NMPC_B200_ACADOS_C_OCP_NLP_INTERFACE_H
#define NMPC_B200_ACADOS_C_OCP_NLP_INTERFACE_H

#ifdef __cplusplus
extern "C" {
#endif

struct nmpc_acados_core;   /* one OCP instance behind a capsule (csrc/acados_shim.cpp) */

typedef struct ocp_nlp_config { struct nmpc_acados_core* core; } ocp_nlp_config;
typedef struct ocp_nlp_dims { struct nmpc_acados_core* core; int N, nx, nu, ny, nyn, np, nbx, nbu; } ocp_nlp_dims;
typedef struct ocp_nlp_in { struct nmpc_acados_core* core; } ocp_nlp_in;
typedef struct ocp_nlp_out {
    struct nmpc_acados_core* core;
    double inf_norm_res;   /* inf-norm of the final QP residuals of the last solve */
} ocp_nlp_out;
typedef struct ocp_nlp_solver { struct nmpc_acados_core* core; } ocp_nlp_solver;

/* fields: "lbx" "ubx" (stage 0: nx values = the measured state, lbx must equal ubx at solve time;
 * stages 1..N: nbx values), "lbu" "ubu" (stages 0..N-1: nbu values).  Returns 0, or 1 on a bad
 * stage / unknown field. */
int ocp_nlp_constraints_model_set(ocp_nlp_config* config, ocp_nlp_dims* dims, ocp_nlp_in* in, ocp_nlp_out* out,
                                  int stage, const char* field, void* value);

/* fields: "W" (stage < N: ny x ny column-major, stage N: nyn x nyn; must be diagonal, which is all
 * the reference ever sets — off-diagonal entries make the next solve return status 4),
 * "yref" / "y_ref" (stage < N: ny values, stage N: nyn values).  Returns 0 or 1. */
int ocp_nlp_cost_model_set(ocp_nlp_config* config, ocp_nlp_dims* dims, ocp_nlp_in* in, int stage, const char* field, void* value);

/* fields: "x" (stage 0..N, nx values), "u" (stage 0..N-1, nu values) of the current iterate */
void ocp_nlp_out_get(ocp_nlp_config* config, ocp_nlp_dims* dims, ocp_nlp_out* out, int stage, const char* field, void* value);

/* fields: "time_tot" (double, seconds, host wall clock of the last solve incl. copies), "sqp_iter" (int, 1),
 * "qp_iter" (int), "status" (int) */
void ocp_nlp_get(ocp_nlp_solver* solver, const char* field, void* return_value_);

#ifdef __cplusplus
}
#endif
#endif
