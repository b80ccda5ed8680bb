"""Offline emitter: YAML -> the model constants the kernels are built with.

Mirrors `scripts/generate_acados_libs.py` (the reference regenerates its acados solver libraries from
config/nmpc_nav_control_acados_models.yaml through scripts/<m>/common.py `load_parameters` and
scripts/<m>/generate_c_code.py:30-60): the same YAML schema (`omni4_params`, `diff_params`, `tric_params`) is read,
the same derived quantities are formed (dt = 1/freq, N = ceil(tf_ini/dt), degrees -> radians for tric, bounds
-/+ v_max on the reference-velocity states and -/+ a_max on the controls), and the result is written as
`csrc/model_defaults.inc`, the table `nmpc_create` starts from ({m}_acados_create's defaults); `--build` then
recompiles the libraries, as the reference's script recompiles the generated C code.

    python -m nmpc_nav_control_b200.emit config.yaml [--out csrc/model_defaults.inc] [--build]

The emitter owns everything the reference fixes at code-generation time:

* the horizon - dt = 1 / freq, N = ceil(tf_ini / dt) (scripts/<m>/common.py:5-9) are written to include/nmpc_horizon.h
  (NMPC_N, NMPC_DT), which the public header, the acados-compatible headers and the kernels (record / tile layouts,
  sweep loops) all take them from; the three models must agree on it, as one library serves all three;
* the model functions - `--models` writes csrc/models_gen.cuh from the symbolic descriptions of models_def.py
  (pose rates and their Jacobian by sympy differentiation + common-subexpression elimination), the role CasADi plays
  for the reference (scripts/<m>/<m>_amr_model.py -> generate_c_code.py);
* the default tables - csrc/model_defaults.inc.

`--out-dir DIR` leaves the tree alone and writes DIR/include/nmpc_horizon.h, DIR/model_defaults.inc (and, with `--build`,
DIR/libnmpc_b200.so compiled against them): an alternate build, selected at run time with NMPC_B200_LIB / NMPC_HORIZON_H.
Weights, bounds and parameters can also be changed at run time without rebuilding (nmpc_set_weights / _bounds /
_params); `specs_from_yaml` returns the `ModelSpec`s for that.
"""
from __future__ import annotations

import argparse
import math
import os
from dataclasses import replace

from .problem import MODELS, ModelSpec

PKG = os.path.dirname(os.path.abspath(__file__))
DEFAULT_OUT = os.path.join(PKG, "csrc", "model_defaults.inc")
_KEYS = {"diff": "diff_params", "omni4": "omni4_params", "tric": "tric_params"}


def horizon_of(prm: dict):
    """scripts/<m>/common.py:5-9: dt = 1 / freq, N = ceil(tf_ini / dt)"""
    dt = 1.0 / float(prm["freq"])
    n = int(math.ceil(float(prm["tf_ini"]) / dt))
    if n < 2 or n > 1000:
        raise ValueError(f"tf_ini={prm['tf_ini']}, freq={prm['freq']} give N={n}: outside 2..1000")
    return n, dt


def spec_from_params(name: str, prm: dict) -> ModelSpec:
    """one model's YAML block -> ModelSpec (what generate_c_code.py bakes into the generated solver)"""
    base = MODELS[name]
    n_h, dt_h = horizon_of(prm)
    nv = base.nv
    q, r, qn = tuple(map(float, prm["Q_diag"])), tuple(map(float, prm["R_diag"])), tuple(map(float, prm["QN_diag"]))
    if len(q) != base.nx or len(r) != base.nu or len(qn) != base.nx:
        raise ValueError(f"{name}: Q_diag / R_diag / QN_diag must have {base.nx} / {base.nu} / {base.nx} entries")
    vmax, amax = float(prm["v_max"]), float(prm["a_max"])
    if name == "diff":                                     # scripts/diff/generate_c_code.py:45-55
        p = (float(prm["dist_b"]), float(prm["tau_v"]))
        lbx, ubx, lbu, ubu = (-vmax,) * nv, (vmax,) * nv, (-amax,) * nv, (amax,) * nv
    elif name == "omni4":                                  # scripts/omni4/generate_c_code.py:45-55
        p = (float(prm["l1_plus_l2"]), float(prm["tau_v"]))
        lbx, ubx, lbu, ubu = (-vmax,) * nv, (vmax,) * nv, (-amax,) * nv, (amax,) * nv
    else:                                                  # scripts/tric/common.py:11-19, generate_c_code.py:45-56
        p = (float(prm["dist_d"]), float(prm["tau_v"]), float(prm["tau_a"]))
        amin, amx = prm["alpha_min"] * math.pi / 180.0, prm["alpha_max"] * math.pi / 180.0
        dmax = prm["dalpha_max"] * math.pi / 180.0
        lbx, ubx, lbu, ubu = (-vmax, amin), (vmax, amx), (-amax, -dmax), (amax, dmax)
    return replace(base, p=p, Q=q, R=r, QN=qn, lbx=lbx, ubx=ubx, lbu=lbu, ubu=ubu, n=n_h, dt=dt_h)


def specs_from_yaml(path: str) -> dict:
    import yaml
    with open(path) as f:
        cfg = yaml.safe_load(f)
    out = {}
    for name, key in _KEYS.items():
        if key in cfg:                                     # generate_acados_libs.py:31-52: absent blocks are skipped
            out[name] = spec_from_params(name, cfg[key])
    if not out:
        raise ValueError(f"{path}: none of {sorted(_KEYS.values())} found")
    hz = {(sp_.n, sp_.dt) for sp_ in out.values()}
    if len(hz) != 1:
        raise ValueError(f"{path}: the models disagree on the horizon {sorted(hz)}; one library (one NMPC_N) serves all three")
    return out


def emit_horizon(n: int, dt: float, source: str = "") -> str:
    """include/nmpc_horizon.h"""
    return (f"/* GENERATED by `python -m nmpc_nav_control_b200.emit` - do not edit.  The horizon of this build, from\n"
            f" * {source or 'config/nmpc_nav_control_acados_models.yaml'} through scripts/<m>/common.py:5-9: dt = 1 / freq, N = ceil(tf_ini / dt).\n"
            f" * Included by include/nmpc_b200.h, include/acados_solver_<m>.h and the kernels (csrc/platform.cuh). */\n"
            f"#ifndef NMPC_HORIZON_H\n#define NMPC_HORIZON_H\n#ifndef NMPC_N\n#define NMPC_N {int(n)}\n#endif\n"
            f"#ifndef NMPC_DT\n#define NMPC_DT {float(dt)!r}\n#endif\n#endif\n")


# ---- model functions from the symbolic descriptions (models_def.py) ------------------------------------------------
def _model_struct(md, variants) -> str:
    """one C++ struct; `variants` = [(preprocessor condition or None, ModelDef)] sharing the interface"""
    import sympy as sp
    from .models_def import th
    nv, np_ = md.nv, md.np_
    a = sp.symbols(f"a0:{nv}", real=True); p = sp.symbols(f"p0:{np_}", real=True)

    def body(d):
        g = list(d.g)
        outs, names = [], []
        for i in range(3):
            outs.append(g[i]); names.append(f"g[{i}]")
        for i in range(3):
            outs.append(sp.diff(g[i], th)); names.append(f"Jth[{i}]")
        for i in range(3):
            for c in range(nv):
                outs.append(sp.diff(g[i], a[c])); names.append(f"Jv[{i}][{c}]")
        # one sincos per distinct angle
        angles = []
        for e in outs:
            for f in e.atoms(sp.sin, sp.cos):
                if f.args[0] not in angles:
                    angles.append(f.args[0])
        angles.sort(key=lambda z: sp.default_sort_key(z))
        pre, sub = [], {}
        for k, ang in enumerate(angles):
            sname, cname = sp.Symbol(f"s{k}", real=True), sp.Symbol(f"c{k}", real=True)
            pre.append(f"double s{k}, c{k}; nmpc_sincos({sp.ccode(_cargs(ang, nv, np_))}, &s{k}, &c{k});")
            sub[sp.sin(ang)] = sname; sub[sp.cos(ang)] = cname
        outs = [e.subs(sub) for e in outs]
        rep, red = sp.cse(outs, symbols=sp.numbered_symbols("t"), optimizations="basic")
        lines = list(pre)
        for sym, ex in rep:
            lines.append(f"const double {sym} = {sp.ccode(_cargs(ex, nv, np_))};")
        for nm, ex in zip(names, red):
            lines.append(f"{nm} = {sp.ccode(_cargs(ex, nv, np_))};")
        lti = (sp.diff(d.g[2], th) == 0) and all(not sp.diff(d.g[2], a[c]).free_symbols & ({th} | set(a)) for c in range(nv))
        return lines, lti

    bodies = [(cond, *body(d)) for cond, d in variants]
    lti = all(b[2] for b in bodies)
    taus = ", ".join(str(t) for t in md.tau)
    out = [f"// {md.source}", f"struct {md.name} {{",
           f"    static constexpr int NV = {nv}, NP = {np_}, ID = {md.model_id};",
           f"    static constexpr bool THETA_ROW_LTI = {'true' if lti else 'false'};   // theta_dot is linear in the actuator states with constant coefficients",
           f"    NMPC_HD static double inv_tau(int c, const double* p) {{ const int ti[{nv}] = {{{taus}}}; return 1.0 / p[ti[c]]; }}",
           "    // g[3] = pose rates, Jth[3] = dg/dtheta, Jv[3][NV] = dg/dactual",
           "    NMPC_HD static void pose_rates(double th, const double* a, const double* p, double* g, double* Jth, double (*Jv)[NV]) {"]
    for k, (cond, lines, _) in enumerate(bodies):
        if cond is not None:
            out.append(("#if " if k == 0 else "#else  // ") + cond if k == 0 else "#else")
        out += ["        " + ln for ln in lines]
    if any(c is not None for c, _, _ in bodies):
        out.append("#endif")
    out += ["    }", "};", ""]
    return "\n".join(out)


def _cargs(expr, nv, np_):
    """symbols a3 / p1 -> array elements a[3] / p[1] for the C printer"""
    import sympy as sp
    sub = {}
    for i in range(nv):
        sub[sp.Symbol(f"a{i}", real=True)] = sp.Symbol(f"a[{i}]", real=True)
    for i in range(np_):
        sub[sp.Symbol(f"p{i}", real=True)] = sp.Symbol(f"p[{i}]", real=True)
    return expr.subs(sub) if hasattr(expr, "subs") else expr


def emit_models() -> str:
    """csrc/models_gen.cuh from models_def.py"""
    from . import models_def as md
    parts = ["// GENERATED by `python -m nmpc_nav_control_b200.emit --models` from nmpc_nav_control_b200/models_def.py - do not edit.",
             "// Pose rates of the three robot models and their Jacobians (sympy differentiation + common-subexpression elimination).",
             "#pragma once", "", "namespace nmpc {", ""]
    parts.append(_model_struct(md.diff_model(), [(None, md.diff_model())]))
    parts.append(_model_struct(md.omni4_model(), [(None, md.omni4_model())]))
    parts.append(_model_struct(md.tric_model(True), [("TRIC_FAITHFUL_COS_BUG", md.tric_model(True)), ("", md.tric_model(False))]))
    parts.append("}  // namespace nmpc")
    return "\n".join(parts) + "\n"


def _arr(v, n):
    v = list(v) + [0.0] * (n - len(v))
    return "{" + ", ".join(repr(float(x)) for x in v) + "}"


def emit_inc(specs: dict, source: str = "") -> str:
    """the three ModelInfo initialisers of csrc/rti_kernels.cu (models missing from `specs` keep the package defaults)"""
    lines = ["// GENERATED by `python -m nmpc_nav_control_b200.emit` - do not edit.  ModelInfo{nx, nu, np, nv, p[3], Q[11], R[4], QN[11],",
             "// lbx[4], ubx[4], lbu[4], ubu[4]} for diff2amr, omni4amr, tric3amr, from " + (source or "the package defaults") + "."]
    for name in ("diff", "omni4", "tric"):
        s = specs.get(name, MODELS[name])
        lines.append(f"    // {s.acados_name}")
        lines.append(f"    {{{s.nx}, {s.nu}, {s.np_}, {s.nv}, {_arr(s.p, 3)}, {_arr(s.Q, 11)}, {_arr(s.R, 4)}, {_arr(s.QN, 11)},")
        lines.append(f"     {_arr(s.lbx, 4)}, {_arr(s.ubx, 4)}, {_arr(s.lbu, 4)}, {_arr(s.ubu, 4)}}},")
    return "\n".join(lines) + "\n"


def _write(path: str, text: str) -> bool:
    old = open(path).read() if os.path.exists(path) else None
    if old != text:
        os.makedirs(os.path.dirname(path), exist_ok=True)
        with open(path, "w") as f:
            f.write(text)
    return old != text


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(description=__doc__.split("\n")[0])
    ap.add_argument("yaml", nargs="?", help="parameter file with the schema of config/nmpc_nav_control_acados_models.yaml")
    ap.add_argument("--out", default=None, help="model_defaults.inc to write (default: csrc/model_defaults.inc, or OUT_DIR/model_defaults.inc)")
    ap.add_argument("--out-dir", default=None, help="alternate build directory: the tree is left untouched")
    ap.add_argument("--models", action="store_true", help="regenerate csrc/models_gen.cuh from models_def.py")
    ap.add_argument("--build", action="store_true", help="recompile the CUDA library (and, in the tree, the acados-compatible libraries)")
    a = ap.parse_args(argv)
    root = os.path.dirname(PKG)
    if a.models:
        changed = _write(os.path.join(PKG, "csrc", "models_gen.cuh"), emit_models())
        print(f"csrc/models_gen.cuh: {'written' if changed else 'unchanged'}")
    if a.yaml:
        specs = specs_from_yaml(a.yaml)
        any_spec = next(iter(specs.values()))
        src = "config/" + os.path.basename(a.yaml)
        inc = a.out or (os.path.join(a.out_dir, "model_defaults.inc") if a.out_dir else DEFAULT_OUT)
        hdr = os.path.join(a.out_dir, "include", "nmpc_horizon.h") if a.out_dir else os.path.join(root, "include", "nmpc_horizon.h")
        c1 = _write(inc, emit_inc(specs, src))
        c2 = _write(hdr, emit_horizon(any_spec.n, any_spec.dt, src))
        print(f"{inc}: {'written' if c1 else 'unchanged'}; {hdr}: {'written' if c2 else 'unchanged'} "
              f"(N = {any_spec.n}, dt = {any_spec.dt!r}; {', '.join(sorted(specs))} from {a.yaml})")
    if a.build:
        from . import build
        if a.out_dir:
            out = build.build_alternate(a.out_dir)
            print(f"{out} built (select it with NMPC_B200_LIB={out} NMPC_HORIZON_H={os.path.join(a.out_dir, 'include', 'nmpc_horizon.h')})")
        else:
            build.build_all(force=True)
            print("libraries rebuilt")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
