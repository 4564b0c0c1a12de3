"""Recipe: build the UNMODIFIED reference into ``oracle/_ref/`` (TEST INFRASTRUCTURE ONLY).

The reference (shirongcan/AlphaZero-Gomoku) is pure Python, so "building" it means byte-compiling
the modules of the hot path from the sources where they lie under ``/root/reference`` into
source-less byte-code files (extension ``.refbin``: snapshot tools tend to drop ``*.pyc``):

    python -m oracle.build_ref            # writes oracle/_ref/{network,train}.refbin, games/*.refbin, mcts/*.refbin

No reference source is copied into the repository: ``oracle/_ref/`` holds compiler output only,
is listed in ``.gitignore`` (never in history) and not in ``.gpurunignore`` (it travels to the GPU
box like the built ``.so``).  ``bench.py --impl reference`` and ``bench.py``'s ``cpu_baseline`` leg
import those modules under their reference names (``activate()``) and run the reference's own ``MCTS``, ``PyTorchModel`` and
``play_game_and_collect`` (``cpu_baseline.kind`` = "reference"); when the directory is missing
they fall back to the oracle port (kind "port").  Nothing in the product package reads it.
"""
from __future__ import annotations

import importlib.machinery
import importlib.util
import os
import py_compile
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
# the modules of SURVEY section 8(a): rules, search, network, self-play driver
MODULES = ("games/__init__.py", "games/gomoku.py", "games/pente.py", "mcts/__init__.py", "mcts/new_mcts_alpha.py",
           "network.py", "train.py")


EXT = ".refbin"


def _out(m: str) -> str:
    return os.path.join(OUT, m[:-3] + EXT)


def available() -> bool:
    return all(os.path.exists(_out(m)) for m in MODULES)


def build(reference: str = "/root/reference", quiet: bool = False) -> bool:
    """Byte-compile MODULES from ``reference`` into oracle/_ref/.  Returns False (and changes nothing)
    when the reference tree is not present - e.g. on the GPU box, which only uses the prebuilt files."""
    if not os.path.isdir(reference):
        return available()
    for m in MODULES:
        src = os.path.join(reference, m)
        dst = _out(m)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        # dfile keeps reference-relative file names in tracebacks; unchecked-hash pycs never look for a source file
        py_compile.compile(src, cfile=dst, dfile=m, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    if not quiet:
        print(f"oracle/_ref: {len(MODULES)} reference modules byte-compiled from {reference} (Python {sys.version_info[0]}.{sys.version_info[1]})")
    return True


def activate() -> bool:
    """Make the byte-compiled reference importable under its own top-level names (``games.gomoku``,
    ``mcts.new_mcts_alpha``, ``network``, ``train`` - the reference imports its modules that way)."""
    if not available():
        return False
    for m in MODULES:
        name = m[:-3].replace("/", ".")
        is_pkg = name.endswith(".__init__")
        if is_pkg:
            name = name[:-9]
        if name in sys.modules:
            continue
        loader = importlib.machinery.SourcelessFileLoader(name, _out(m))
        spec = importlib.util.spec_from_loader(name, loader, is_package=is_pkg)
        mod = importlib.util.module_from_spec(spec)
        if is_pkg:
            mod.__path__ = [os.path.dirname(_out(m))]
        sys.modules[name] = mod
        loader.exec_module(mod)
    return True


if __name__ == "__main__":
    ok = build(os.environ.get("AZG_REFERENCE", "/root/reference"))
    sys.exit(0 if ok else 1)
