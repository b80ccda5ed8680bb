"""Declarative descriptions of the three robot models, the input of the offline emitter (emit.py).

The reference describes a model once, symbolically, in scripts/<m>/<m>_amr_model.py (CasADi) and regenerates the solver's
C code from it (scripts/<m>/generate_c_code.py).  CasADi is not available here; the same role is played by sympy: a
model is the pose-rate vector g(theta, actual; p) and the lag time constants, written below as sympy expressions that
follow the reference's model files term by term, and `emit.emit_models` differentiates them and writes the CUDA device
functions (csrc/models_gen.cuh).  All three models share the cascade of SURVEY.md Appendix A.5:

    x = [pose (x, y, theta) | actual (nv) | ref (nv)],  u = d(ref)/dt
    pose_dot = g(theta, actual; p),  actual_dot = (ref - actual) / tau_c,  ref_dot = u
"""
from __future__ import annotations

from dataclasses import dataclass

import sympy as sp


@dataclass(frozen=True)
class ModelDef:
    name: str            # C++ struct name
    model_id: int
    nv: int              # actuator channels
    np_: int             # parameters
    g: tuple             # pose rates (x_dot, y_dot, theta_dot) as sympy expressions in th, a[i], p[i]
    tau: tuple           # parameter index of every channel's time constant
    source: str          # where the reference defines it


th = sp.Symbol("th", real=True)


def _syms(nv, np_):
    return sp.symbols(f"a0:{nv}", real=True), sp.symbols(f"p0:{np_}", real=True)


def diff_model() -> ModelDef:
    """scripts/diff/diff_amr_model.py:42-60: v = (vr + vl) / 2, w = (vr - vl) / dist_b (the lag definition of vl_dot /
    vr_dot at :53-54 is the live one, the first assignment at :51-52 is dead code)"""
    (vl, vr), (dist_b, tau_v) = _syms(2, 2)
    v = (vr + vl) / 2
    w = (vr - vl) / dist_b
    return ModelDef("DiffModel", 0, 2, 2, (v * sp.cos(th), v * sp.sin(th), w), (1, 1), "scripts/diff/diff_amr_model.py:42-60")


def omni4_model() -> ModelDef:
    """scripts/omni4/omni4_amr_model.py:52-73"""
    (v1, v2, v3, v4), (l12, tau_v) = _syms(4, 2)
    v = (v1 - v2 + v3 - v4) / 4
    vn = (-v1 - v2 + v3 + v4) / 4
    w = (-v1 - v2 - v3 - v4) / (2 * l12)
    return ModelDef("Omni4Model", 1, 4, 2, (v * sp.cos(th) - vn * sp.sin(th), v * sp.sin(th) + vn * sp.cos(th), w), (1, 1, 1, 1),
                    "scripts/omni4/omni4_amr_model.py:52-73")


def tric_model(faithful_cos_bug: bool = True) -> ModelDef:
    """scripts/tric/tric_amr_model.py:43-59; :45 defines `cos_alpha = ca.sin(alpha)` - reproduced by default (the acados
    build of the reference has it), `faithful_cos_bug=False` gives the intended kinematics"""
    (v, alpha), (dist_d, tau_v, tau_a) = _syms(2, 3)
    cos_alpha = sp.sin(alpha) if faithful_cos_bug else sp.cos(alpha)
    return ModelDef("TricModel", 2, 2, 3, (v * sp.cos(th) * cos_alpha, v * sp.sin(th) * cos_alpha, v / dist_d * sp.sin(alpha)), (1, 2),
                    "scripts/tric/tric_amr_model.py:43-59")


def all_models(faithful_cos_bug: bool = True):
    return [diff_model(), omni4_model(), tric_model(faithful_cos_bug)]
