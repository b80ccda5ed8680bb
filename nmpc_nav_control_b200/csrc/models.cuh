// Robot models as device functions (K1 input).
//
// All three reference models share one structure:
//   x = [pose(3) = (x, y, theta) | actual(NV) | ref(NV)],  u = d(ref)/dt (NV)
//   pose_dot   = g(theta, actual; p)                       (model specific, nonlinear)
//   actual_dot = (ref - actual) / tau_c                    (first-order lag per channel)
//   ref_dot    = u
// so a model is described by NV, NP, tau_c(p) and g with its Jacobian w.r.t. (theta, actual).
//
// The functions themselves are GENERATED: nmpc_nav_control_b200/models_def.py states g symbolically, following
// scripts/diff/diff_amr_model.py:42-60, scripts/omni4/omni4_amr_model.py:52-73 and scripts/tric/tric_amr_model.py:43-59
// term by term, and `python -m nmpc_nav_control_b200.emit --models` differentiates it with sympy and writes
// csrc/models_gen.cuh - the counterpart of the reference regenerating its solver's C code from the CasADi model
// (scripts/<m>/generate_c_code.py:12-77).  tric_amr_model.py:45 defines cos_alpha = sin(alpha); reproduced unless
// TRIC_FAITHFUL_COS_BUG is 0.
#pragma once
#include "platform.cuh"

#ifndef TRIC_FAITHFUL_COS_BUG
#define TRIC_FAITHFUL_COS_BUG 1
#endif

#include "models_gen.cuh"
