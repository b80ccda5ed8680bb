/*
 * nmpc_b200.h — batched C ABI of the B200-native SQP-RTI solver (libnmpc_b200.so).
 *
 * This is the NEW batched entry point the north star asks for; the reference has no analogue
 * (its caller solves one OCP per `{m}_acados_solve`, src/nmpc_nav_control/NMPCNavControlDiff.cpp:142).
 * The closest upstream shape is acados' `{m}_acados_batch_solve(capsule**, int*, int)`.
 * The single-instance acados-compatible ABI the reference's C++ controller binds is declared
 * in include/acados_solver_{diff2amr,omni4amr,tric3amr}.h and include/acados_c/ocp_nlp_interface.h
 * and is implemented on top of the functions below with batch = 1.
 *
 * Conventions: plain C, no exceptions, int return (0 = ok, <0 = NMPC_E_*).  "host" pointers
 * are caller-owned and copied synchronously; "device" pointers are caller-owned CUDA device
 * memory on the solver's device.  Device batches are structure-of-arrays with the instance
 * index fastest: a[stage][component][instance].
 */
#ifndef NMPC_B200_H
#define NMPC_B200_H
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#include "nmpc_horizon.h"               /* NMPC_N, NMPC_DT: scripts/<m>/common.py:5-9, N = ceil(tf_ini * freq); emitted, see emit.py.
                                         * nmpc_dims().n is the horizon of the loaded library */

enum { NMPC_MODEL_DIFF = 0, NMPC_MODEL_OMNI4 = 1, NMPC_MODEL_TRIC = 2 };

/* per-instance status = acados status codes */
enum { NMPC_SUCCESS = 0, NMPC_NAN_DETECTED = 1, NMPC_MAXITER = 2, NMPC_MINSTEP = 3, NMPC_QP_FAILURE = 4 };

enum {
    NMPC_E_ARG = -1,        /* bad argument / unsupported option                          */
    NMPC_E_CUDA = -2,       /* CUDA runtime error (message via nmpc_last_error)           */
    NMPC_E_NODEVICE = -3,   /* no CUDA device: there is no CPU fallback                   */
    NMPC_E_CAPACITY = -4    /* batch larger than the capacity given at creation           */
};

/* interior-point options; defaults restate acados' PARTIAL_CONDENSING_HPIPM settings
 *  */
typedef struct {
    double mu0, alpha_min, res_g_max, res_b_max, res_d_max, res_m_max;
    double reg_prim, lam_min, t_min, tau_min, thr0;
    int iter_max, cond_pred_corr;
} nmpc_ipm_opts;

typedef struct nmpc_solver nmpc_solver;

typedef struct { int nx, nu, np, ny, nyn, nbx, nbu, n; } nmpc_dims_t;

int nmpc_dims(int model, nmpc_dims_t* out);
void nmpc_default_opts(nmpc_ipm_opts* o);
const char* nmpc_last_error(void);

/* One solver = one model on one device with room for `max_batch` instances.  Tables start at
 * the code-generation defaults of config/nmpc_nav_control_acados_models.yaml (what
 * `{m}_acados_create` yields, NMPCNavControlDiff.cpp:10-12); the iterate starts at x_k = default
 * x0 (0,0,pi,0,..), u = 0. */
int nmpc_create(int model, int max_batch, int device, nmpc_solver** out);
int nmpc_destroy(nmpc_solver* s);

/* stage-wise tables shared by all instances (host pointers).  Any pointer may be NULL = keep.
 *   W_diag [N][ny]  diagonal of W in y order [x;u]   (ocp_nlp_cost_model_set(i,"W"), Diff.cpp:68-71)
 *   We_diag [nx]    diagonal of W_e                   (ocp_nlp_cost_model_set(N,"W"), Diff.cpp:72-73)
 *   lbx/ubx [N][nbx] row k = stage k+1                (constraints_model_set(i,"lbx"/"ubx"), Diff.cpp:49-56)
 *   lbu/ubu [N][nbu] stages 0..N-1                    (constraints_model_set(i,"lbu"/"ubu"), Diff.cpp:58-65)
 *   p [N][np]                                         ({m}_acados_update_params, Diff.cpp:44-46)  */
int nmpc_set_weights(nmpc_solver* s, const double* W_diag, const double* We_diag);
int nmpc_set_bounds(nmpc_solver* s, const double* lbx, const double* ubx, const double* lbu, const double* ubu);
int nmpc_set_params(nmpc_solver* s, const double* p);
/* read the tables back (same layouts; any pointer may be NULL = skip) */
int nmpc_get_tables(const nmpc_solver* s, double* W_diag, double* We_diag, double* lbx, double* ubx, double* lbu, double* ubu,
                    double* p);
int nmpc_set_opts(nmpc_solver* s, const nmpc_ipm_opts* o);
int nmpc_get_opts(const nmpc_solver* s, nmpc_ipm_opts* o);

/* the persisted iterate (x [N+1][nx][cap], u [N][nu][cap], device, SoA with leading dim cap =
 * max_batch).  reset: zero it ({m}_acados_reset, Diff.cpp:177-181). */
int nmpc_iterate_device(nmpc_solver* s, double** d_x, double** d_u, int* leading_dim);
int nmpc_reset(nmpc_solver* s);
/* same, enqueued on `stream` (cudaStream_t, NULL = the default stream) without host sync */
int nmpc_reset_async(nmpc_solver* s, void* stream);
/* instance-major host copies: x [B][N+1][nx], u [B][N][nu].  They run on the solver's own stream and wait only for it:
 * synchronise the stream of any device call (solve, shift, rollout) that touched the iterate before reading it back. */
int nmpc_set_iterate_host(nmpc_solver* s, int B, const double* x, const double* u);
int nmpc_get_iterate_host(nmpc_solver* s, int B, double* x, double* u);

/* ---- the batched entry point: one SQP-RTI iteration of B independent OCPs ------------------
 * All pointers are DEVICE pointers, SoA with leading dimension B:
 *   d_x0bar [nx][B]            measured initial state          (stage-0 lbx = ubx, Diff.cpp:96-101)
 *   d_yref  [N+1][nyref][B]    nyref = 3 (pose, rest 0) or ny  (cost_model_set(i,"yref"), Diff.cpp:121-124);
 *                              the terminal row uses its first nx entries
 *   d_We    [nx][B] or NULL    per-instance terminal weights   (the diff wrapper's W_e switch, Diff.cpp:127-139)
 *   d_x, d_u                   iterate in/out; NULL = the solver's own persisted iterate
 *   d_status [B] (int)         acados status per instance
 *   d_qp_iter [B] (int)        interior-point iterations per instance
 *   d_stats [8][B] or NULL     res_g,res_b,res_d,res_m, mu, lin_res, cond_fallbacks, qp_status
 * Asynchronous on `stream` (cudaStream_t; NULL = the default stream); no host synchronisation inside.
 * Which kernels run is chosen from the batch size (same results to rounding, same iteration counts):
 *   B <= 4 x SMs (diff, tric)   one thread block per instance, state in shared memory - the latency path (batch 1: 0.7 ms)
 *   B <  24,576                 the persistent lane-cooperative kernel
 *   larger                      hybrid: lockstep per-lane sweeps, then the lane-cooperative kernel for the stragglers
 * Environment overrides read at nmpc_create, for experiments: NMPC_K3 = sweep | group | hybrid, NMPC_SOLO_MAX = <largest batch
 * of the block-per-instance kernel, 0 = never>, NMPC_HYB_MIN, NMPC_HYB_FRAC, NMPC_HYB_KMAX, NMPC_CHUNK. */
int nmpc_rti_solve_device(nmpc_solver* s, int B, const double* d_x0bar, const double* d_yref, int nyref,
                          const double* d_We, double* d_x, double* d_u, int ldxu,
                          int* d_status, int* d_qp_iter, double* d_stats, void* stream);

/* ---- BASELINE config 4 / north-star kernel (4): warm-start shift and SQP to convergence ---------
 * The reference never shifts and does one RTI per tick (NMPCNavControlROS.cpp:309,316,326 ->
 * NMPCNavControlDiff.cpp:142); the shapes mirrored here are its simulation scripts: the explicit warm start from the
 * previous solution (scripts/test_scripts/casadi_sim_diff.py:104-106) and the closed loop of
 * scripts/test_scripts/acados_sim_diff.py:119-163.
 * nmpc_shift_device: x_k <- x_{k+1} (k < N), x_N kept; u_k <- u_{k+1} (k < N-1), u_{N-1} kept.
 *   d_x, d_u NULL = the persisted iterate; d_mask [B] (int) or NULL: only instances with a non-zero entry shift.
 * nmpc_sqp_solve_device: per instance, repeats the RTI step (same arguments as nmpc_rti_solve_device) until the inf-norm
 *   of its step (dx, du over all stages) is <= tol, its status is non-zero, or max_iter steps were taken.  A device-side
 *   mask keeps the instances that stopped out of the following passes (their queue is rebuilt on the device before every
 *   pass); all max_iter passes are enqueued, nothing synchronises with the host.
 *   d_status [B]: status of the instance's last step; d_sqp_iter [B] or NULL: RTI steps taken; d_qp_iter [B] or NULL:
 *   interior-point iterations summed over them. */
int nmpc_shift_device(nmpc_solver* s, int B, double* d_x, double* d_u, int ldxu, const int* d_mask, void* stream);
int nmpc_sqp_solve_device(nmpc_solver* s, int B, const double* d_x0bar, const double* d_yref, int nyref, const double* d_We,
                          double* d_x, double* d_u, int ldxu, int max_iter, double tol, int* d_status, int* d_sqp_iter,
                          int* d_qp_iter, void* stream);

/* ---- host-buffer call (what a controller process would bind): instance-major host arrays ---
 *   x0bar [B][nx], yref [B][N+1][nyref], We [B][nx] or NULL   -> copied H2D and transposed on device
 *   u0 [B][nu], x1 [B][nx], status [B], qp_iter [B]            <- copied D2H (what run() reads, Diff.cpp:151-169)
 * Uses and updates the solver's persisted iterate.  Synchronous.  Calls with at most 64 instances (the ROS drop-in's one
 * robot per call, NMPCNavControlDiff.cpp:96-169) go through one pinned staging buffer with ONE copy each way; the results are
 * the same bits as the general path's. */
int nmpc_rti_solve_host(nmpc_solver* s, int B, const double* x0bar, const double* yref, int nyref, const double* We,
                        double* u0, double* x1, int* status, int* qp_iter);

/* ---- the controller's per-tick glue, batched and on the device ---------------
 * One call = NMPCNavControl{Diff,Omni4,Tric}::run for B robots (src/nmpc_nav_control/NMPCNavControlDiff.cpp:82-175,
 * NMPCNavControlOmni4.cpp:91-177, NMPCNavControlTric.cpp:88-181): initial state from pose + direct kinematics of the
 * measured twist + the reference states carried from the previous tick, heading references unwrapped along the chain
 * that starts at the robot's heading, reference list padded with its last pose to N+1 rows, (diff) terminal weights
 * W_e[0..2] = 100 x W[0..2] when the last two reference rows coincide, one RTI step on the solver's persisted iterate,
 * reference states advanced by u_0 * dt, inverse kinematics.  The kinematic constant is p[0] of stage 0 (dist_b /
 * l1_plus_l2), W and W_e are the solver's tables (the wrappers initialise W_e from Q: set it with nmpc_set_weights).
 *   d_pose [3][B]           x, y, theta
 *   d_vel  [3][B]           v, vn, w (measured body twist; vn is used by omni4 only)
 *   d_steer [B] or NULL     tric: measured steering-wheel angle
 *   d_refs [nref_max][3][B] reference poses x, y, theta (theta wrapped; rows beyond N+1 are ignored)
 *   d_nref [B] or NULL      number of valid reference poses per robot, clamped to 1..nref_max (NULL = nref_max for all)
 *   dt                      controller period (1 / control_freq, NMPCNavControlROS.cpp:82)
 *   d_cmd  [3][B]           out: diff (v, w, 0), omni4 (v, vn, w), tric (v, alpha, 0); a robot whose solve did not
 *                           return 0 keeps its previous command and carried state (run() throws before writing them)
 *   d_status, d_qp_iter     [B] or NULL
 * Asynchronous on `stream`.  nmpc_ctrl_reset zeroes the carried reference states (the constructors' state,
 * Diff.cpp:14); nmpc_reset is reset_mpc().  nmpc_ctrl_state_device exposes them: [nv][max_batch]. */
int nmpc_ctrl_tick_device(nmpc_solver* s, int B, const double* d_pose, const double* d_vel, const double* d_steer,
                          const double* d_refs, const int* d_nref, int nref_max, double dt, double* d_cmd,
                          int* d_status, int* d_qp_iter, void* stream);
/* the same tick with SQP to convergence in place of the single RTI step (sqp_max_iter, sqp_tol as nmpc_sqp_solve_device) */
int nmpc_ctrl_tick_sqp_device(nmpc_solver* s, int B, const double* d_pose, const double* d_vel, const double* d_steer,
                              const double* d_refs, const int* d_nref, int nref_max, double dt, int sqp_max_iter, double sqp_tol,
                              double* d_cmd, int* d_status, int* d_qp_iter, void* stream);
int nmpc_ctrl_reset(nmpc_solver* s, void* stream);
int nmpc_ctrl_state_device(nmpc_solver* s, double** d_vref, int* leading_dim);
/* host-buffer form: pose [B][3], vel [B][3], steer [B] or NULL, refs [B][nref_max][3] (nref_max <= N+1), nref [B] or
 * NULL, cmd [B][3] in/out, status [B], qp_iter [B] or NULL.  Synchronous. */
int nmpc_ctrl_tick_host(nmpc_solver* s, int B, const double* pose, const double* vel, const double* steer,
                        const double* refs, const int* nref, int nref_max, double dt, double* cmd, int* status,
                        int* qp_iter);

/* ---- batched path discretisation ----------------------------------------------
 * PathDiscretizer::getNextNPoses (src/nmpc_nav_control/PathDiscretizer.cpp:14-63) for B robots: from the robot's
 * nearest path parameter, walk the path in steps of a tenth of the distance covered in one sample period at the
 * segment's speed, emit a pose each time the accumulated chord length reaches that distance (within 1 %), pad with the
 * path's end pose.  The reference's curves are `parametric_trajectories_common::TPath`, a private dependency; here a
 * path is a run of segments, sixteen doubles each: */
typedef struct {
    double kind;        /* 0: x(u) = sum cx[i] u^i, y(u) = sum cy[i] u^i (line, cubic Bezier, up to degree 5)            */
                        /* 1: arc x(u) = cx[0] + cx[1] cos(cx[2] + cx[3] u), y(u) = cy[0] + cx[1] sin(cx[2] + cx[3] u)    */
    double vel;         /* signed speed on the segment (TPath::GetVelocity); negative = heading flipped by pi            */
    double th0, th1;    /* holonomic heading at u = 0 and u = 1 (TPath::GetThetaHolomonic, linear in between)            */
    double cx[6], cy[6];
} nmpc_path_segment;
/*   d_segments [n_seg][16]     all paths' segments, path q = segments d_path_offsets[q] .. d_path_offsets[q+1]-1
 *   d_path_offsets [n_paths+1] (int)
 *   d_path_id [B] (int) or NULL (= path 0 for every robot), clamped to 0..n_paths-1
 *   d_nearest_u [B]            segment index + parameter in [0,1) (active_path_u_, NMPCNavControlROS.cpp:668)
 *   sample_period, num_poses, is_holonomic   the constructor arguments (PathDiscretizer.cpp:5-12; the node passes
 *                              dt, N+1, false: NMPCNavControlROS.cpp:666)
 *   d_poses [num_poses][3][B]  out: x, y, theta - the d_refs layout of nmpc_ctrl_tick_device
 * Stateless; asynchronous on `stream` of device `device`. */
int nmpc_path_discretize_device(int device, int B, const double* d_segments, const int* d_path_offsets, int n_paths,
                                const int* d_path_id, const double* d_nearest_u, double sample_period, int num_poses,
                                int is_holonomic, double* d_poses, void* stream);

/* ---- the remaining pieces of a closed loop on the device --------------------------
 * nmpc_plant_step_device: the nominal plant of a rollout, x+ = phi_RK4(x, u_0 + noise) over `dt` with the solver's own
 * model and parameters (stage 0), u_0 = stage-0 controls of the solver's persisted iterate (what the tick just
 * computed); scripts/test_scripts/acados_sim_diff.py:136-160 is the reference's version (Euler, Gaussian noise on the
 * accelerations).  Writes the measurements the next tick reads.
 *   d_noise [nu][B] or NULL    added to u_0 (the caller draws it; seeded)
 *   d_xplant [nx][B]           plant state, in/out
 *   d_pose [3][B], d_vel [3][B] (v, vn, w: inverse kinematics of the actuator states), d_steer [B] or NULL (tric)
 * nmpc_path_nearest_device: path parameter nearest to each robot's position, searched in
 * [u - back, u + ahead] (clipped to the path) by 24 coarse samples and a ternary refinement.  The reference uses the
 * private parametric_trajectories_common::TPathProcessMinDist (NMPCNavControlROS.cpp:596-600); this is a stand-in
 * with its own definition.   d_u [B]: in = previous parameter, out = nearest. */
int nmpc_plant_step_device(nmpc_solver* s, int B, double dt, const double* d_noise, double* d_xplant, double* d_pose,
                           double* d_vel, double* d_steer, void* stream);
int nmpc_path_nearest_device(int device, int B, const double* d_segments, const int* d_path_offsets, int n_paths,
                             const int* d_path_id, const double* d_pose, double back, double ahead, double* d_u, void* stream);

/* nmpc_rollout_device: the whole closed loop for B robots, `ticks` ticks enqueued by ONE call with nothing between them on
 * the host.  Per tick, in the order of NMPCNavControlROS::processFollowPath (NMPCNavControlROS.cpp:648-720): nearest
 * path parameter -> N+1 reference poses -> controller tick (one RTI step, or SQP to convergence) -> nominal plant step
 * (scripts/test_scripts/acados_sim_diff.py:136-160) -> optional warm-start shift of the iterate.
 *   d_u [B]                  path parameter per robot, in/out
 *   d_xplant [nx][B], d_pose [3][B], d_vel [3][B], d_steer [B] (tric; else NULL)    plant state and measurements, in/out
 *   d_noise [ticks][nu][B] or NULL      added to u_0 in the plant (drawn and seeded by the caller)
 *   d_traj [ticks+1][3][B] or NULL      out: pose before the first tick and after every tick
 *   d_cmds [ticks][3][B] or NULL        out: velocity commands
 *   d_nfail [ticks] (int) or NULL       out: robots whose solve returned a non-zero status, per tick
 * The controller state (carried reference states, iterate) is the solver's: nmpc_ctrl_reset / nmpc_reset start a run. */
typedef struct {
    double dt;              /* controller period = sample period of the reference poses                        */
    double back, ahead;     /* nearest-point search window (path-parameter units), see nmpc_path_nearest_device */
    double sqp_tol;         /* SQP: step inf-norm at which an instance stops iterating                          */
    int is_holonomic;       /* PathDiscretizer flag (the node passes false)                                     */
    int sqp_max_iter;       /* 1 = one RTI step per tick (the reference); > 1 = SQP to convergence              */
    int shift;              /* 1 = shift the iterate one stage after every tick (the reference: 0)              */
} nmpc_rollout_opts;
int nmpc_rollout_device(nmpc_solver* s, int B, int ticks, const nmpc_rollout_opts* o, const double* d_segments,
                        const int* d_path_offsets, int n_paths, const int* d_path_id, double* d_u, double* d_xplant,
                        double* d_pose, double* d_vel, double* d_steer, const double* d_noise, double* d_traj, double* d_cmds,
                        int* d_nfail, void* stream);

/* statistics of the last nmpc_rti_solve_host call, stats [8][B] (rows as d_stats above), host pointer */
int nmpc_last_stats_host(nmpc_solver* s, int B, double* stats);

/* timing of the last nmpc_rti_solve_* call on this solver, measured with CUDA events on its
 * stream: ms[0] = K1+K2 linearise, ms[1] = K3 QP, ms[2] = K4 step, ms[3] = total incl. copies */
int nmpc_last_timing(nmpc_solver* s, double* ms4);
/* number of kernel launches issued by the last solve call */
int nmpc_last_launches(const nmpc_solver* s);

/* tiny device benchmark used by bench.py to measure the fp64 FMA peak (roofline denominator);
 * returns achieved TFLOP/s */
double nmpc_dfma_peak_tflops(int device, int iters);

#ifdef __cplusplus
}
#endif
#endif
