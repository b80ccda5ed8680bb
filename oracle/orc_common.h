/*
 * ORACLE (test infrastructure, NOT product code) — shared option / statistics structs.
 *
 * PARITY UNPINNED (see orc_models.h).  The constants below restate what SURVEY.md
 * Appendix B.1/B.4 records for acados' PARTIAL_CONDENSING_HPIPM defaults (BALANCE mode
 * with acados' overrides); they cannot be verified here and therefore live in ONE struct.
 */
#ifndef ORC_COMMON_H
#define ORC_COMMON_H

#ifndef ORC_N
#define ORC_N 80           /* scripts/{diff,omni4,tric}/common.py: N = ceil(tf_ini*freq) = 80; -DORC_N=... for another horizon */
#endif
#define ORC_HIST 64

typedef struct {
    double mu0;            /* 1.0   : acados override of HPIPM's mode default           */
    double alpha_min;      /* 1e-8                                                      */
    double res_g_max;      /* 1e-6                                                      */
    double res_b_max;      /* 1e-8                                                      */
    double res_d_max;      /* 1e-8                                                      */
    double res_m_max;      /* 1e-8                                                      */
    double reg_prim;       /* 1e-15 : added to the Hessian diagonal before factorising  */
    double lam_min;        /* 1e-16                                                     */
    double t_min;          /* 1e-16                                                     */
    double tau_min;        /* 1e-16                                                     */
    double thr0;           /* 0.1   : cold-start slack threshold                        */
    int iter_max;          /* 50                                                        */
    int pred_corr;         /* 1                                                         */
    int cond_pred_corr;    /* 1                                                         */
    int itref_corr_max;    /* 2                                                         */
    int lq_fact;           /* 1 : only the >1e-5 accuracy test is evaluated and counted */
} orc_ipm_opts;

static inline void orc_ipm_opts_default(orc_ipm_opts *o)
{
    o->mu0 = 1.0; o->alpha_min = 1e-8;
    o->res_g_max = 1e-6; o->res_b_max = 1e-8; o->res_d_max = 1e-8; o->res_m_max = 1e-8;
    o->reg_prim = 1e-15; o->lam_min = 1e-16; o->t_min = 1e-16; o->tau_min = 1e-16;
    o->thr0 = 0.1; o->iter_max = 50; o->pred_corr = 1; o->cond_pred_corr = 1;
    o->itref_corr_max = 2; o->lq_fact = 1;
}

typedef struct {
    int status;            /* acados: 0 ok, 1 NaN, 2 maxiter, 3 minstep, 4 QP failure   */
    int qp_status;         /* hpipm : 0 ok, 1 maxiter, 2 minstep, 3 NaN                  */
    int qp_iter;
    int itref_solves;      /* refinement solves actually performed                      */
    int lq_flags;          /* iterations whose predictor lin-residual exceeded 1e-5      */
    int cond_fallbacks;    /* iterations that fell back to the pure centering direction */
    double res[4];         /* final inf-norms res_g,res_b,res_d,res_m                    */
    double mu;
    double lin_res_max[4]; /* max over iterations of the corrector's lin-system residuals*/
    double alpha_hist[ORC_HIST];
    double mu_hist[ORC_HIST];
} orc_stats;

#endif
