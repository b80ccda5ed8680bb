/* acados/ocp_nlp/ocp_nlp_cost_ls.h — include-path shim.  The reference's NMPCNavControl.h:10-17 includes this header but uses
 * no symbol from it; it exists so the wrapper compiles unchanged against the
 * B200-native solver. */
#ifndef NMPC_B200_SHIM_ACADOS_OCP_NLP_OCP_NLP_COST_LS_H
#define NMPC_B200_SHIM_ACADOS_OCP_NLP_OCP_NLP_COST_LS_H
#endif
