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


def test_other_horizons_and_bad_sizes_are_rejected(tmp_path):
    cfg = _yaml_of_defaults()
    y = tmp_path / "m.yaml"
    for key, val in (("tf_ini", 3.0), ("freq", 50)):
        bad = dict(diff_params=dict(cfg["diff_params"], **{key: val}))
        y.write_text(yaml.safe_dump(bad))
        with pytest.raises(ValueError, match="compiled for N=80"):
            emit.specs_from_yaml(str(y))
    y.write_text(yaml.safe_dump(dict(diff_params=dict(cfg["diff_params"], Q_diag=[1, 2, 3]))))
    with pytest.raises(ValueError, match="entries"):
        emit.specs_from_yaml(str(y))
    y.write_text(yaml.safe_dump(dict(other=1)))
    with pytest.raises(ValueError, match="none of"):
        emit.specs_from_yaml(str(y))


def test_cli_writes_the_table(tmp_path):
    y = tmp_path / "m.yaml"; out = tmp_path / "t.inc"
    y.write_text(yaml.safe_dump(_yaml_of_defaults()))
    assert emit.main([str(y), "--out", str(out)]) == 0
    assert out.read_text().count("\n    {") == 3 and "tric3amr" in out.read_text()
    assert emit.main([str(y), "--out", str(out)]) == 0          # second run: unchanged
