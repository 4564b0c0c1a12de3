"""CPU: the C-ABI library loads and exports every symbol include/azgomoku_b200.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def header_symbols():
    text = open(os.path.join(ROOT, "include", "azgomoku_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(azg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_header():
    import alphazero_gomoku_b200 as m
    lib = ctypes.CDLL(m.LIB_PATH)
    syms = header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
    assert lib.azg_abi_version() == 1


def test_binding_covers_header():
    from alphazero_gomoku_b200 import _lib
    assert sorted(_lib.PROTOTYPES) == header_symbols()


def test_struct_layout():
    from alphazero_gomoku_b200 import _lib
    assert ctypes.sizeof(_lib.azg_pos) == 96
    assert _lib.azg_pos.player.offset == 64 and _lib.azg_pos.plies.offset == 80


def test_no_cpu_fallback():
    """Without a GPU every compute entry point must fail loudly, never compute on the CPU."""
    import torch
    import pytest
    import alphazero_gomoku_b200 as m
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert m.lib.azg_device_count() == 0
    with pytest.raises(m.AzgError):
        m.SearchEngine(m.GOMOKU, 1)
