"""Swap the GPU ``VectorBase`` / ``EmbeddingIndex`` into an importable typeagent.

The reference has no plugin registry: ``VectorBase`` is a concrete class bound by name at
import time in six modules.  ``install()`` rebinds that name in each module that is importable (the reference's own
``EmbeddingIndex`` is a pure forwarder and then builds the GPU class), and returns what it patched;
``uninstall()`` restores the originals.  typeagent itself is optional: without it,
``install()`` patches nothing and says so.
"""

from __future__ import annotations

import importlib

# module -> names bound there (reference: knowpro/fuzzyindex.py:9,23; storage/memory/
# reltermsindex.py:10-14,266; storage/memory/convthreads.py:4,20; storage/sqlite/
# messageindex.py:12,33; storage/sqlite/reltermsindex.py:11,139; aitools/vectorbase.py:82)
_SITES = {
    "typeagent.aitools.vectorbase": ("VectorBase",),
    "typeagent.knowpro.fuzzyindex": ("VectorBase",),
    "typeagent.storage.memory.reltermsindex": ("VectorBase",),
    "typeagent.storage.memory.convthreads": ("VectorBase",),
    "typeagent.storage.sqlite.messageindex": ("VectorBase",),
    "typeagent.storage.sqlite.reltermsindex": ("VectorBase",),
}

_saved: dict[tuple[str, str], object] = {}


def install(**vectorbase_options) -> list[str]:
    """Rebind the names; ``vectorbase_options`` (device=, storage_dtype=, normalize=) become
    the defaults of every VectorBase typeagent constructs afterwards."""
    from . import vectorbase

    if vectorbase_options:
        base_cls = type(
            "VectorBase",
            (vectorbase.VectorBase,),
            {"__init__": lambda self, settings, **kw: vectorbase.VectorBase.__init__(
                self, settings, **{**vectorbase_options, **kw})},
        )
    else:
        base_cls = vectorbase.VectorBase
    replacements = {"VectorBase": base_cls}
    patched = []
    for mod_name, names in _SITES.items():
        try:
            mod = importlib.import_module(mod_name)
        except Exception:
            continue
        for name in names:
            if hasattr(mod, name):
                _saved.setdefault((mod_name, name), getattr(mod, name))
                setattr(mod, name, replacements[name])
                patched.append(f"{mod_name}.{name}")
    return patched


def uninstall() -> None:
    for (mod_name, name), original in list(_saved.items()):
        try:
            setattr(importlib.import_module(mod_name), name, original)
        except Exception:
            pass
        del _saved[(mod_name, name)]
