"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/cbx.h declares,
and the ctypes struct mirrors agree with the compiled structs.  No compute calls (no GPU here)."""
import ctypes as C
import os
import re

import pytest

from marlon_b200 import _abi, _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()
    return _lib.load()


def test_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, "include", "cbx.h")).read()
    declared = set(re.findall(r"\b(cbx_[a-z_0-9]+)\s*\(", header))
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name


def test_struct_mirrors(lib):
    assert lib.cbx_abi_version() == _abi.ABI_VERSION
    assert lib.cbx_abi_sizeof(0) == C.sizeof(_abi.Config)
    assert lib.cbx_abi_sizeof(1) == C.sizeof(_abi.Views)
    assert lib.cbx_abi_sizeof(2) == C.sizeof(_abi.Tape)
    c = _abi.Config()
    assert lib.cbx_config_default(C.byref(c)) == 0
    assert (c.maximum_node_count, c.maximum_total_credentials, c.winning_reward) == (100, 1000, 5000.0)
    assert list(c.kind_of_index) == [_abi.KIND_CONNECT, _abi.KIND_LOCAL, _abi.KIND_REMOTE]
    assert c.def_sla_worsening_penalty_scale == 200.0 and c.att_max_timesteps == 2000


def test_scenario_blob_validation(lib):
    from marlon_b200 import scenario, scenarios

    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    blob = comp.tobytes()
    h = C.c_void_p()
    assert lib.cbx_scenario_create(blob, len(blob), C.byref(h)) == 0
    assert lib.cbx_scenario_destroy(h) == 0
    bad = b"\0" * len(blob)
    assert lib.cbx_scenario_create(bad, len(bad), C.byref(h)) == -1
    assert b"magic" in lib.cbx_last_error()


def test_no_cpu_fallback(lib):
    """device < 0 (or no CUDA device) is an error, never a silent CPU path."""
    import torch

    from marlon_b200 import config, scenario, scenarios

    comp = scenario.compile_scenario(scenarios.toyctf_environment())
    blob = comp.tobytes()
    s, b = C.c_void_p(), C.c_void_p()
    assert lib.cbx_scenario_create(blob, len(blob), C.byref(s)) == 0
    cfg = config.make_config(maximum_node_count=12, maximum_total_credentials=10)
    assert lib.cbx_batch_create(s, 4, C.byref(cfg), -1, C.byref(b)) == -4
    if not torch.cuda.is_available():
        assert lib.cbx_batch_create(s, 4, C.byref(cfg), 0, C.byref(b)) == -4
        from marlon_b200.batch import Batch

        with pytest.raises(RuntimeError):
            Batch(comp, cfg, 4)
    lib.cbx_scenario_destroy(s)
