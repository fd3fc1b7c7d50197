"""CPU: the C-ABI library builds/loads and exports every symbol include/rectipy_b200.h declares (no compute calls)."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from rectipy_b200 import _cabi
    if not os.path.exists(_cabi.library_path()):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "rectipy_b200", "csrc")])
    return _cabi.load()


def test_header_symbols_are_exported(lib):
    from rectipy_b200 import _cabi
    header = open(os.path.join(ROOT, "include", "rectipy_b200.h")).read()
    declared = set(re.findall(r"\b(rp_[a-z_0-9]+)\s*\(", header))
    declared -= {"rp_desc", "rp_plan"}
    assert declared == set(_cabi.EXPORTS), declared ^ set(_cabi.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_host_only_entry_points(lib):
    from rectipy_b200 import _cabi as abi
    assert lib.rp_abi_version() == abi.RP_ABI_VERSION
    assert [lib.rp_num_state_vars(m) for m in range(6)] == [1, 1, 2, 3, 2, 3]
    assert [lib.rp_num_history_planes(m) for m in range(6)] == [1, 1, 2, 3, 2, 4]
    assert lib.rp_num_state_vars(99) == -1
    # record counting must match Network.run's windowing (network.py:590-597)
    for T, S, cut in [(100, 1, 0), (100, 2, 0), (600, 5, 7), (10, 3, 9), (10, 3, 10), (5, 100, 0), (0, 1, 0), (40, 4, 39)]:
        expect = len([t for t in range(T) if t >= cut and t % S == 0])
        assert lib.rp_num_records(T, S, cut) == expect, (T, S, cut)


def test_struct_sizes_match_header(lib, tmp_path):
    """ctypes mirrors of rp_desc / rp_fwd_args / rp_bwd_args must have the C compiler's size and layout."""
    from rectipy_b200 import _cabi as abi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "rectipy_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(rp_desc), sizeof(rp_fwd_args), sizeof(rp_bwd_args),'
                   ' offsetof(rp_fwd_args, history), offsetof(rp_bwd_args, g_x), offsetof(rp_desc, param_per_neuron));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    sizes = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert sizes == [ctypes.sizeof(abi.rp_desc), ctypes.sizeof(abi.rp_fwd_args), ctypes.sizeof(abi.rp_bwd_args),
                     abi.rp_fwd_args.history.offset, abi.rp_bwd_args.g_x.offset, abi.rp_desc.param_per_neuron.offset]


def test_plan_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rectipy_b200 import _cabi as abi
    d = abi.rp_desc()
    d.model, d.n, d.batch, d.dt, d.in_mode, d.out_mode = abi.RP_QIF, 8, 1, 1e-3, abi.RP_IN_DENSE, abi.RP_OUT_DENSE
    d.out_var = abi.RP_VAR_S
    h = ctypes.c_void_p()
    assert lib.rp_plan_create(ctypes.byref(d), ctypes.byref(h)) != 0
    assert b"no CUDA device" in lib.rp_last_error()
