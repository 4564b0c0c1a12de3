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
    assert lib.azg_abi_version() == 3


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


def test_argument_errors_are_codes_with_text():
    """Error behaviour of the boundary: negative return code + azg_last_error() text, never an exception
    or a crash from C (checked on paths that return before touching the device)."""
    import alphazero_gomoku_b200 as m
    from alphazero_gomoku_b200 import _lib
    lib = m.lib
    h = ctypes.c_void_p()
    assert lib.azg_create(None, ctypes.byref(h)) < 0
    assert b"null" in lib.azg_last_error()
    cfg = _lib.azg_config()
    cfg.n_games, cfg.queue_len, cfg.node_capacity, cfg.rule = 4, 257, 1024, 0         # queue_len > AZG_MAX_QUEUE
    assert lib.azg_create(ctypes.byref(cfg), ctypes.byref(h)) < 0 and b"queue_len" in lib.azg_last_error()
    cfg.queue_len, cfg.rule = 32, 7
    assert lib.azg_create(ctypes.byref(cfg), ctypes.byref(h)) < 0
    assert lib.azg_net_create(0, 6, 100, 64, ctypes.byref(h)) < 0 and b"channels" in lib.azg_last_error()
    assert lib.azg_net_create(0, -1, 128, 64, ctypes.byref(h)) < 0
    for fn in (lib.azg_search_fill, lib.azg_search_read_counters):
        assert fn(None, None, None, None) < 0
    assert lib.azg_search_commit(None, None, None) < 0
    assert lib.azg_destroy(None) == 0 and lib.azg_net_destroy(None) == 0              # destroying nothing is fine
