"""The oracle port (oracle/rectipy_oracle.py) must reproduce the fixtures minted from the unmodified reference."""
import numpy as np
import pytest
import torch

from golden_util import Case, RUN_CASES, oracle_run_case, rel_err, orc, GOLDEN
import os


@pytest.mark.parametrize("name", RUN_CASES)
@pytest.mark.parametrize("dtype_name", ["float64", "float32"])
def test_oracle_matches_reference_fixture(name, dtype_name):
    case = Case(name)
    res = oracle_run_case(case, dtype_name)
    # same torch ops in the same order as the reference => agreement far below the parity tolerance
    tol = 1e-12 if dtype_name == "float64" else 2e-6
    assert np.array_equal(res["steps"], case.ref(dtype_name, "steps"))
    for key, val in res.items():
        if key == "steps":
            continue
        ref = case.ref(dtype_name, key)
        assert val.shape == ref.shape, key
        assert rel_err(val, ref) <= tol, (key, rel_err(val, ref))


def test_spiking_fixtures_do_spike():
    for name, thr in (("qif_bptt", 100.0), ("qif_sfa_fwd", 100.0), ("lif_bptt", 10.0)):
        case = Case(name)
        s = case.ref("float64", "out")
        assert np.abs(s).max() > 0.0, name


def test_edges_fixture():
    z = np.load(os.path.join(GOLDEN, "edges.npz"))
    w = torch.tensor(z["lin_w"])
    out = np.stack([orc.linear_forward(w, torch.tensor(x)).numpy() for x in z["xs"]])
    assert rel_err(out, z["lin_out"]) < 1e-14
    # Linear == torch.nn.Linear without bias (rectipy_tests/test_edges.py:80)
    lin = torch.nn.Linear(w.shape[1], w.shape[0], bias=False, dtype=torch.float64)
    with torch.no_grad():
        lin.weight.copy_(w)
    assert rel_err(lin(torch.tensor(z["xs"])).detach().numpy(), z["lin_out"]) < 1e-12
    rls = orc.OracleRLS(w.shape[1], w.shape[0], beta=float(z["rls_beta"]), alpha=float(z["rls_alpha"]))
    for i, (x, y) in enumerate(zip(z["xs"], z["ys"])):
        xt, yt = torch.tensor(x), torch.tensor(y)
        rls.update(xt, yt, rls.forward(xt))
        assert rel_err(rls.weights.numpy(), z["rls_w"][i]) < 1e-12
        assert rel_err(rls.P.numpy(), z["rls_p"][i]) < 1e-12
        assert abs(float(rls.loss) - z["rls_loss"][i]) < 1e-10


def test_ridge_fixture():
    z = np.load(os.path.join(GOLDEN, "ridge.npz"))
    node = orc.make_node("li_tanh", z["W"].shape[0], z["W"], float(z["dt"]), params=dict(tau=1.0))
    net = orc.OracleNet(node, w_in=torch.tensor(z["w_in"]))
    r = net.run(torch.tensor(z["inputs"]), sampling_steps=int(z["S"]), enable_grad=False)
    X = torch.stack(r["out"])
    assert rel_err(X.numpy(), z["X"]) < 1e-12
    w, y = orc.ridge_fit(X, torch.tensor(z["targets"]), float(z["alpha"]))
    assert rel_err(w.numpy(), z["w_out"]) < 1e-6
    assert rel_err(y.numpy(), z["y"]) < 1e-6
