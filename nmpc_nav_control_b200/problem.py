"""OCP definitions of the three robot models (host-side data, no arithmetic on the hot path).

Mirrors, for the SQP-RTI path only, what the reference fixes at code-generation time:

* sizes / state ordering  - scripts/diff/diff_amr_model.py:15-26,
                            scripts/omni4/omni4_amr_model.py:19-34,
                            scripts/tric/tric_amr_model.py:15-27
* horizon                 - scripts/<m>/common.py:5-9  (N = ceil(tf_ini*freq) = 80, dt = 0.025)
* weights, bounds, params - config/nmpc_nav_control_acados_models.yaml:2-75 through
                            scripts/<m>/generate_c_code.py:30-60
* default x0              - scripts/<m>/generate_c_code.py:58-60  (0,0,pi,0,...)

`ModelSpec.codegen_defaults()` gives the stage-wise tables the acados-generated solver would
hold right after `{m}_acados_create` (what BASELINE config 1 names as "N and weights from
config/nmpc_nav_control_acados_models.yaml").
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np



def _read_horizon():
    """N and dt of the build this package drives: the emitted include/nmpc_horizon.h (emit.py; the reference fixes them at
    code-generation time, scripts/<m>/common.py:5-9).  NMPC_HORIZON_H selects the header of an alternate build."""
    import os
    import re
    path = os.environ.get("NMPC_HORIZON_H") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                             "include", "nmpc_horizon.h")
    txt = open(path).read()
    n = int(re.search(r"#define\s+NMPC_N\s+(\d+)", txt).group(1))
    dt = float(re.search(r"#define\s+NMPC_DT\s+([0-9.eE+-]+)", txt).group(1))
    return n, dt


N_HORIZON, DT = _read_horizon()   # the reference's yaml: freq 40 -> dt = 0.025, N = ceil(2.0 / 0.025) = 80, TF = N dt = 2.0


@dataclass(frozen=True)
class ModelSpec:
    name: str              # short name used by this package ("diff", "omni4", "tric")
    acados_name: str       # symbol prefix of the generated solver the reference links
    model_id: int          # index used by the C ABI and the synthetic-input RNG
    nx: int
    nu: int
    np_: int
    idxbx: tuple           # bounded state indices (the *reference* velocity states)
    p: tuple               # default model parameters
    Q: tuple
    R: tuple
    QN: tuple
    lbx: tuple
    ubx: tuple
    lbu: tuple
    ubu: tuple
    n: int = N_HORIZON
    dt: float = DT
    x0_default: tuple = field(default=())

    @property
    def ny(self) -> int:
        return self.nx + self.nu

    @property
    def nbx(self) -> int:
        return len(self.idxbx)

    @property
    def nbu(self) -> int:
        return self.nu

    @property
    def nv(self) -> int:
        """number of actuated channels (= nu); state = [pose(3), actual(nv), ref(nv)]"""
        return self.nu

    def codegen_defaults(self) -> dict:
        """Stage-wise problem tables as the generated solver holds them after create()."""
        n = self.n
        W = np.tile(np.array(self.Q + self.R, dtype=np.float64), (n, 1))
        return dict(
            dt=self.dt,
            W=W,                                                   # [N][ny] diagonal, order [x;u]
            We=np.array(self.QN, dtype=np.float64),                # [nx]
            lbx=np.tile(np.array(self.lbx, dtype=np.float64), (n, 1)),   # row k -> stage k+1
            ubx=np.tile(np.array(self.ubx, dtype=np.float64), (n, 1)),
            lbu=np.tile(np.array(self.lbu, dtype=np.float64), (n, 1)),
            ubu=np.tile(np.array(self.ubu, dtype=np.float64), (n, 1)),
            p=np.tile(np.array(self.p, dtype=np.float64), (n, 1)),
        )


_DEG = math.pi / 180.0

DIFF = ModelSpec(
    name="diff", acados_name="diff2amr", model_id=0, nx=7, nu=2, np_=2, idxbx=(5, 6),
    p=(0.270, 0.1),
    Q=(10.0, 10.0, 5.0, 0.0, 0.0, 0.0, 0.0), R=(1.0, 1.0),
    QN=(1000.0, 1000.0, 500.0, 0.0, 0.0, 0.0, 0.0),
    lbx=(-1.0, -1.0), ubx=(1.0, 1.0), lbu=(-2.0, -2.0), ubu=(2.0, 2.0),
    x0_default=(0.0, 0.0, math.pi, 0.0, 0.0, 0.0, 0.0),
)

OMNI4 = ModelSpec(
    name="omni4", acados_name="omni4amr", model_id=1, nx=11, nu=4, np_=2, idxbx=(7, 8, 9, 10),
    p=(0.535, 0.1),
    Q=(10.0, 10.0, 10.0) + (0.0,) * 8, R=(1.0,) * 4,
    QN=(10.0, 10.0, 10.0) + (0.0,) * 8,
    lbx=(-1.0,) * 4, ubx=(1.0,) * 4, lbu=(-1.0,) * 4, ubu=(1.0,) * 4,
    x0_default=(0.0, 0.0, math.pi) + (0.0,) * 8,
)

TRIC = ModelSpec(
    name="tric", acados_name="tric3amr", model_id=2, nx=7, nu=2, np_=3, idxbx=(5, 6),
    p=(0.270, 0.1, 0.5),
    Q=(10.0, 10.0, 5.0, 0.0, 0.0, 0.0, 0.0), R=(1.0, 1.0),
    QN=(1000.0, 1000.0, 500.0, 0.0, 0.0, 0.0, 0.0),
    # scripts/tric/common.py:17-19 converts deg -> rad
    lbx=(-1.0, -30.0 * _DEG), ubx=(1.0, 30.0 * _DEG),
    lbu=(-1.0, -120.0 * _DEG), ubu=(1.0, 120.0 * _DEG),
    x0_default=(0.0, 0.0, math.pi, 0.0, 0.0, 0.0, 0.0),
)

MODELS = {"diff": DIFF, "omni4": OMNI4, "tric": TRIC}
MODELS_BY_ACADOS_NAME = {m.acados_name: m for m in MODELS.values()}


def get_model(name: str) -> ModelSpec:
    if name in MODELS:
        return MODELS[name]
    if name in MODELS_BY_ACADOS_NAME:
        return MODELS_BY_ACADOS_NAME[name]
    raise KeyError(f"unknown model {name!r}; expected one of {sorted(MODELS)}")
