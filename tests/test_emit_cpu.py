"""the offline emitter (nmpc_nav_control_b200/emit.py) against the reference's own parameter file when
the reference tree is mounted, and against the package defaults restated as a YAML otherwise."""
import os

import numpy as np
import pytest
import yaml

from nmpc_nav_control_b200 import emit
from nmpc_nav_control_b200.problem import MODELS

REF_YAML = "/root/reference/config/nmpc_nav_control_acados_models.yaml"
INC = os.path.join(os.path.dirname(emit.__file__), "csrc", "model_defaults.inc")


def _yaml_of_defaults():
    """the package defaults written in the reference's schema (degrees for tric)"""
    deg = 180.0 / np.pi
    d, o, t = MODELS["diff"], MODELS["omni4"], MODELS["tric"]
    com = lambda s: dict(tf_ini=2.0, freq=40, tau_v=s.p[1], v_max=s.ubx[0], a_max=s.ubu[0], Q_diag=list(s.Q), R_diag=list(s.R),
                         QN_diag=list(s.QN))
    return dict(diff_params=dict(com(d), dist_b=d.p[0]), omni4_params=dict(com(o), l1_plus_l2=o.p[0]),
                tric_params=dict(com(t), dist_d=t.p[0], tau_a=t.p[2], alpha_min=round(t.lbx[1] * deg, 9),
                                 alpha_max=round(t.ubx[1] * deg, 9), dalpha_max=round(t.ubu[1] * deg, 9)))


def _same(a, b):
    for f in ("p", "Q", "R", "QN", "lbx", "ubx", "lbu", "ubu"):
        assert np.allclose(getattr(a, f), getattr(b, f), rtol=1e-15, atol=0), f
    assert (a.nx, a.nu, a.np_, a.n, a.dt) == (b.nx, b.nu, b.np_, b.n, b.dt)


@pytest.mark.skipif(not os.path.exists(REF_YAML), reason="reference tree not mounted")
def test_reference_yaml_gives_the_package_defaults_and_the_committed_table():
    specs = emit.specs_from_yaml(REF_YAML)
    assert sorted(specs) == ["diff", "omni4", "tric"]
    for n, s in specs.items():
        _same(s, MODELS[n])
    assert emit.emit_inc(specs, "config/" + os.path.basename(REF_YAML)) == open(INC).read()


def test_defaults_round_trip_and_partial_files(tmp_path):
    y = tmp_path / "m.yaml"
    y.write_text(yaml.safe_dump(_yaml_of_defaults()))
    specs = emit.specs_from_yaml(str(y))
    for n, s in specs.items():
        _same(s, MODELS[n])
    # the committed table holds the same numbers
    body = lambda t: [ln for ln in t.splitlines() if not ln.startswith("//")]
    assert body(emit.emit_inc(specs)) == body(open(INC).read())
    # a file with one block: the other models keep their defaults (generate_acados_libs.py:31-52 skips absent blocks)
    cfg = _yaml_of_defaults()
    one = dict(diff_params=dict(cfg["diff_params"], dist_b=0.4, v_max=0.8, a_max=1.5, Q_diag=[20, 20, 8, 0, 0, 0, 0]))
    y.write_text(yaml.safe_dump(one))
    sp = emit.specs_from_yaml(str(y))
    assert list(sp) == ["diff"] and sp["diff"].p[0] == 0.4 and sp["diff"].ubx == (0.8, 0.8) and sp["diff"].lbu == (-1.5, -1.5)
    inc = emit.emit_inc(sp)
    assert "{0.4, 0.1, 0.0}" in inc and "{20.0, 20.0, 8.0," in inc and "{0.535, 0.1, 0.0}" in inc
    tb = sp["diff"].codegen_defaults()
    assert tb["W"].shape == (80, 9) and tb["W"][0, 0] == 20.0 and tb["ubx"][0, 0] == 0.8


def test_horizon_comes_from_the_yaml_and_bad_sizes_are_rejected(tmp_path):
    """scripts/<m>/common.py:5-9: dt = 1 / freq, N = ceil(tf_ini / dt); one library serves all three models, so they must agree"""
    cfg = _yaml_of_defaults()
    y = tmp_path / "m.yaml"
    y.write_text(yaml.safe_dump(dict(diff_params=dict(cfg["diff_params"], tf_ini=1.0))))
    sp = emit.specs_from_yaml(str(y))["diff"]
    assert (sp.n, sp.dt) == (40, 0.025) and sp.codegen_defaults()["W"].shape == (40, 9)
    y.write_text(yaml.safe_dump(dict(diff_params=dict(cfg["diff_params"], tf_ini=1.01, freq=50))))
    sp = emit.specs_from_yaml(str(y))["diff"]
    assert (sp.n, sp.dt) == (51, 0.02)                           # ceil(1.01 * 50)
    assert "#define NMPC_N 51" in emit.emit_horizon(sp.n, sp.dt) and "#define NMPC_DT 0.02\n" in emit.emit_horizon(sp.n, sp.dt)
    mixed = dict(diff_params=dict(cfg["diff_params"], tf_ini=1.0), tric_params=cfg["tric_params"])
    y.write_text(yaml.safe_dump(mixed))
    with pytest.raises(ValueError, match="disagree on the horizon"):
        emit.specs_from_yaml(str(y))
    y.write_text(yaml.safe_dump(dict(diff_params=dict(cfg["diff_params"], Q_diag=[1, 2, 3]))))
    with pytest.raises(ValueError, match="entries"):
        emit.specs_from_yaml(str(y))
    y.write_text(yaml.safe_dump(dict(other=1)))
    with pytest.raises(ValueError, match="none of"):
        emit.specs_from_yaml(str(y))
    # the committed header is the emitter's output for the reference's horizon
    root = os.path.dirname(os.path.dirname(emit.__file__))
    body = lambda t: [ln for ln in t.splitlines() if ln.startswith("#")]
    assert body(emit.emit_horizon(80, 0.025)) == body(open(os.path.join(root, "include", "nmpc_horizon.h")).read())


def test_generated_model_functions_are_current_and_equal_the_hand_written_ones(tmp_path):
    """csrc/models_gen.cuh is what emit.emit_models() gives for models_def.py, and the emitted pose_rates / Jacobians equal
    the hand-written round-1 functions (tests/host_emul/models_hand.cuh) to a few ulp on random inputs, both tric variants"""
    import subprocess
    root = os.path.dirname(os.path.dirname(emit.__file__))
    gen = os.path.join(root, "nmpc_nav_control_b200", "csrc", "models_gen.cuh")
    assert emit.emit_models() == open(gen).read()
    src = tmp_path / "cmp.cpp"
    src.write_text(r"""
#include <cstdio>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include "nmpc_nav_control_b200/csrc/models.cuh"
#include "tests/host_emul/models_hand.cuh"
static double ulps(double a, double b) {
    if (a == b) return 0.0;
    // in units of the last place of the larger operand, or of 1 (the scale of the inputs) where the value is a small
    // difference of O(1) terms: a different association of the same sum is allowed its cancellation error
    const double m = std::fmax(1.0, std::fmax(std::fabs(a), std::fabs(b)));
    return std::fabs(a - b) / (m * 2.220446049250313e-16);
}
template <class A, class B> static double cmp(unsigned seed) {
    srand(seed);
    double worst = 0.0;
    for (int t = 0; t < 2000; t++) {
        double th = 6.0 * (rand() / (double)RAND_MAX - 0.5), a[4], p[3];
        for (int i = 0; i < 4; i++) a[i] = 2.0 * (rand() / (double)RAND_MAX - 0.5);
        for (int i = 0; i < 3; i++) p[i] = 0.1 + rand() / (double)RAND_MAX;
        double g1[3], g2[3], j1[3], j2[3], v1[3][A::NV], v2[3][A::NV];
        A::pose_rates(th, a, p, g1, j1, v1); B::pose_rates(th, a, p, g2, j2, v2);
        for (int i = 0; i < 3; i++) {
            worst = std::fmax(worst, std::fmax(ulps(g1[i], g2[i]), ulps(j1[i], j2[i])));
            for (int c = 0; c < A::NV; c++) worst = std::fmax(worst, ulps(v1[i][c], v2[i][c]));
        }
        for (int c = 0; c < A::NV; c++) if (A::inv_tau(c, p) != B::inv_tau(c, p)) worst = 1e9;
    }
    static_assert(A::NV == B::NV && A::NP == B::NP && A::ID == B::ID && A::THETA_ROW_LTI == B::THETA_ROW_LTI, "interface");
    return worst;
}
int main() {
    const double d = cmp<nmpc::DiffModel, nmpc_hand::DiffModel>(1), o = cmp<nmpc::Omni4Model, nmpc_hand::Omni4Model>(2),
                 t = cmp<nmpc::TricModel, nmpc_hand::TricModel>(3);
    std::printf("%g %g %g\n", d, o, t);
    return (d <= 4.0 && o <= 4.0 && t <= 4.0) ? 0 : 1;
}
""")
    for bug in (1, 0):
        exe = tmp_path / f"cmp{bug}"
        r = subprocess.run(["g++", "-std=c++17", "-O1", "-ffp-contract=off", "-DNMPC_HOST_EMUL", f"-DTRIC_FAITHFUL_COS_BUG={bug}", "-I", root,
                            str(src), "-o", str(exe)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-3000:]
        r = subprocess.run([str(exe)], capture_output=True, text=True)
        assert r.returncode == 0, ("worst ulp differences (diff, omni4, tric):", r.stdout)


def test_alternate_horizon_build_passes_parity_at_n40(tmp_path):
    """a YAML with tf_ini = 1.0 (N = 40): the emitter writes the horizon header and the table into an alternate build
    directory; the kernels' logic compiled for that horizon (host emulation) agrees with the oracle compiled for it -
    the lane-cooperative K3 and the lockstep sweeps - in a fresh interpreter that reads the alternate header"""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(emit.__file__))
    cfg = _yaml_of_defaults()
    for k in cfg:
        cfg[k]["tf_ini"] = 1.0
    y = tmp_path / "n40.yaml"
    y.write_text(yaml.safe_dump(cfg))
    out = tmp_path / "alt"
    assert emit.main([str(y), "--out-dir", str(out)]) == 0
    hdr = out / "include" / "nmpc_horizon.h"
    assert "#define NMPC_N 40" in hdr.read_text() and (out / "model_defaults.inc").exists()
    code = r"""
import sys
sys.path.insert(0, %r); sys.path.insert(0, %r)
import numpy as np
from nmpc_nav_control_b200.problem import MODELS, N_HORIZON
assert N_HORIZON == 40 and MODELS["diff"].n == 40
import emul, helpers
from oracle import orc
assert orc.ORC_N == 40
for name, B in (("diff", 24), ("tric", 16), ("omni4", 8)):
    spec, x0, yref, _ = helpers.instances(name, 50, B)
    assert yref.shape[1] == 41
    ref = helpers.oracle_solve(orc, name, x0, yref)
    for kw in (dict(coop=True), dict()):
        out = emul.emul_rti(name, x0, yref, **kw)
        assert (out["qp_iter"] == ref["qp_iter"]).all(), (name, kw)
        assert helpers.parity_report(out["x"], ref["x"])[0] == 0 and helpers.parity_report(out["u"], ref["u"])[0] == 0, (name, kw)
print("n40 parity ok")
""" % (root, os.path.join(root, "tests"))
    env = dict(os.environ, NMPC_HORIZON_H=str(hdr))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=900)
    assert r.returncode == 0 and "n40 parity ok" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])


def test_cli_writes_the_table(tmp_path):
    y = tmp_path / "m.yaml"; out = tmp_path / "t.inc"
    y.write_text(yaml.safe_dump(_yaml_of_defaults()))
    assert emit.main([str(y), "--out", str(out)]) == 0
    assert out.read_text().count("\n    {") == 3 and "tric3amr" in out.read_text()
    assert emit.main([str(y), "--out", str(out)]) == 0          # second run: unchanged
