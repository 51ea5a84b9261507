"""``install()`` / ``uninstall()``: rebinding the name ``VectorBase`` in the modules that import it
(SURVEY.md §8b).  typeagent itself is not importable here (its LLM dependencies are absent), so the
six import sites are stood in for by stub modules with the reference's module names."""

from __future__ import annotations

import sys
import types

import numpy as np

import typeagent_py_b200 as tab
from oracle import vectorbase_oracle as O

SITES = [
    "typeagent.aitools.vectorbase",
    "typeagent.knowpro.fuzzyindex",
    "typeagent.storage.memory.reltermsindex",
    "typeagent.storage.memory.convthreads",
    "typeagent.storage.sqlite.messageindex",
    "typeagent.storage.sqlite.reltermsindex",
]


class ReferenceVectorBase:  # what the stub modules bind before install()
    pass


def _stub_typeagent(monkeypatch):
    names = set()
    for site in SITES:
        parts = site.split(".")
        for i in range(1, len(parts) + 1):
            names.add(".".join(parts[:i]))
    for name in sorted(names):
        mod = types.ModuleType(name)
        mod.__path__ = []
        if name in SITES:
            mod.VectorBase = ReferenceVectorBase
        monkeypatch.setitem(sys.modules, name, mod)


def test_install_rebinds_every_site_and_uninstall_restores(monkeypatch):
    _stub_typeagent(monkeypatch)
    patched = tab.install()
    assert sorted(patched) == sorted(f"{s}.VectorBase" for s in SITES)
    for site in SITES:
        assert sys.modules[site].VectorBase is tab.VectorBase
    tab.uninstall()
    for site in SITES:
        assert sys.modules[site].VectorBase is ReferenceVectorBase


def test_install_with_options_sets_constructor_defaults(monkeypatch):
    _stub_typeagent(monkeypatch)
    tab.install(storage_dtype="bfloat16", device=0)
    try:
        cls = sys.modules["typeagent.storage.memory.reltermsindex"].VectorBase
        assert cls is not tab.VectorBase and issubclass(cls, tab.VectorBase)
        base = cls(tab.TextEmbeddingIndexSettings(O.FakeEmbeddingModel()))
        assert base._storage_dtype == "bfloat16"
        base.add_embeddings(None, np.eye(4, dtype=np.float32))
        assert len(base) == 4
        explicit = cls(tab.TextEmbeddingIndexSettings(O.FakeEmbeddingModel()), storage_dtype="float32")
        assert explicit._storage_dtype == "float32"
    finally:
        tab.uninstall()


def test_install_without_typeagent_patches_nothing():
    import pytest

    if any(name == "typeagent" or name.startswith("typeagent.") for name in sys.modules):
        pytest.skip("a typeagent module (the oracle's reference loader) is already imported")
    assert tab.install() == []
    tab.uninstall()
