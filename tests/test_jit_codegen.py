"""Run-time code generation (rectipy_b200/jit.py) on the CPU box: parsing, variable roles, generated source, NVRTC -> sm_100a cubin.
Loading and running the image needs a device (tests/test_gpu_jit.py)."""
import numpy as np
import pytest

from rectipy_b200 import templates, jit, _cabi as abi
from test_gpu_jit import YAML


@pytest.fixture
def user_templates(tmp_path, monkeypatch):
    (tmp_path / "mymodels").mkdir()
    (tmp_path / "mymodels" / "custom.yaml").write_text(YAML)
    monkeypatch.chdir(tmp_path)


def test_unmatched_equations_become_a_run_time_compiled_spec(user_templates):
    spec = templates.resolve_template("mymodels.custom.adex")
    assert spec.model == abi.RP_JIT and spec.jit_program is None
    assert [k for k, _ in spec.state_vars] == ["adex_op/v", "adex_op/w", "adex_op/s"]          # the reference's y order: equation order
    assert set(spec.input_vars) == {"adex_op/I_ext", "adex_op/spike", "adex_op/s_in"}
    assert abi.RP_P_K not in [slot for slot, _ in spec.params.values()]                       # the fold slot stays free
    bound = jit.bind_spec(spec, "s", "s_in", "I_ext", "spike", "v")
    prog = bound.jit_program
    assert prog.spiking and prog.nsv == 3 and prog.src_plane == bound.planes["adex_op/s"] == 2 and bound.planes["adex_op/v"] == 0
    assert bound.params[jit.ONE_KEY] == (abi.RP_P_K, 1.0)
    assert prog.image[:4] == b"\x7fELF"
    for sym in (b"rp_jit_fwd_step", b"rp_jit_adj_step", b"rp_jit_init_src"):
        assert sym in prog.image
    assert "expf(" in prog.source and "double" not in prog.source.split("struct JitInitSrcArgs")[0].split("jit_field")[1]
    again = jit.bind_spec(spec, "s", "s_in", "I_ext", "spike", "v")
    assert again.jit_program is prog                                                         # cached by source hash


def test_reset_variable_moves_to_plane_zero_and_sources_may_be_expressions(user_templates, tmp_path):
    (tmp_path / "mymodels" / "other.yaml").write_text(
        "o_op:\n  base: OperatorTemplate\n  equations:\n    - \"a' = -a/tau_a + spike\"\n    - \"v' = -v + I_ext + J*a_in\"\n"
        "  variables:\n    a: output(0.0)\n    v: variable(0.0)\n    tau_a: 2.0\n    J: 0.5\n    I_ext: input(0.0)\n    spike: input(0.0)\n"
        "    a_in: input(0.0)\no:\n  base: NodeTemplate\n  operators:\n    - o_op\n")
    spec = templates.resolve_template("mymodels.other.o")
    b = jit.bind_spec(spec, "a", "a_in", "I_ext", "spike", "v")
    assert b.planes == {"o_op/v": 0, "o_op/a": 1} and b.plane_order == [1, 0]
    fhn = jit.bind_spec(templates.resolve_template("mymodels.custom.fhn"), "r", "r_in", "I_ext", None, None)
    assert fhn.jit_program.src_plane == -1 and not fhn.jit_program.spiking and fhn.source_var == "fhn_op/r"


def test_generator_refuses_what_the_kernels_cannot_express(user_templates, tmp_path):
    (tmp_path / "mymodels" / "bad.yaml").write_text(
        "m_op:\n  base: OperatorTemplate\n  equations:\n    - \"v' = -v + mean(v) + I_ext\"\n  variables:\n    v: output(0.0)\n    I_ext: input(0.0)\n"
        "m:\n  base: NodeTemplate\n  operators:\n    - m_op\n"
        "q_op:\n  base: OperatorTemplate\n  equations:\n    - \"v' = -v + r_in*r_in + I_ext\"\n  variables:\n    v: output(0.0)\n    I_ext: input(0.0)\n    r_in: input(0.0)\n"
        "q:\n  base: NodeTemplate\n  operators:\n    - q_op\n")
    with pytest.raises(NotImplementedError):
        templates.resolve_template("mymodels.bad.m")
    with pytest.raises(NotImplementedError):         # d f / d u depends on u: the adjoint product would need the drive of the step
        jit.bind_spec(templates.resolve_template("mymodels.bad.q"), "v", "r_in", "I_ext", None, None)
    with pytest.raises(KeyError):
        jit.bind_spec(templates.resolve_template("mymodels.custom.fhn"), "nope", "r_in", "I_ext", None, None)
