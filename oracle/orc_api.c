/*
 * ORACLE (test infrastructure, NOT product code) — instantiates the RTI restatement for the
 * three models of JorgeDFR/nmpc_nav_control.  PARITY UNPINNED (see orc_models.h).
 * Sizes: scripts/diff/diff_amr_model.py:15-26, scripts/omni4/omni4_amr_model.py:19-34,
 * scripts/tric/tric_amr_model.py:15-27; bounded indices: scripts/<m>/generate_c_code.py:45-55.
 */
#include <stdlib.h>
#include <stddef.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "orc_common.h"
#include "orc_models.h"

#define ORC_CAT_(a, b) a##b
#define ORC_CAT(a, b) ORC_CAT_(a, b)

/* diff2amr */
#define ORC_NX 7
#define ORC_NU 2
#define ORC_NP 2
#define ORC_NBX 2
#define ORC_NBU 2
#define ORC_IDXBX {5, 6}
#define ORC_F orc_diff_f
#define ORC_JAC orc_diff_jac
#define ORC_FN(n) ORC_CAT(orc_diff_, n)
#include "orc_rti_core.inc"
#undef ORC_NX
#undef ORC_NU
#undef ORC_NP
#undef ORC_NBX
#undef ORC_NBU
#undef ORC_IDXBX
#undef ORC_F
#undef ORC_JAC
#undef ORC_FN

/* omni4amr */
#define ORC_NX 11
#define ORC_NU 4
#define ORC_NP 2
#define ORC_NBX 4
#define ORC_NBU 4
#define ORC_IDXBX {7, 8, 9, 10}
#define ORC_F orc_omni4_f
#define ORC_JAC orc_omni4_jac
#define ORC_FN(n) ORC_CAT(orc_omni4_, n)
#include "orc_rti_core.inc"
#undef ORC_NX
#undef ORC_NU
#undef ORC_NP
#undef ORC_NBX
#undef ORC_NBU
#undef ORC_IDXBX
#undef ORC_F
#undef ORC_JAC
#undef ORC_FN

/* tric3amr */
#define ORC_NX 7
#define ORC_NU 2
#define ORC_NP 3
#define ORC_NBX 2
#define ORC_NBU 2
#define ORC_IDXBX {5, 6}
#define ORC_F orc_tric_f
#define ORC_JAC orc_tric_jac
#define ORC_FN(n) ORC_CAT(orc_tric_, n)
#include "orc_rti_core.inc"

void orc_default_opts(orc_ipm_opts *o) { orc_ipm_opts_default(o); }
int orc_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
