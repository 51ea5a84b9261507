"""Load the UNMODIFIED reference modules (``vectorbase.py``, ``fuzzyindex.py`` and the index
classes that call them) — from ``/root/reference`` in the build container, or from the
git-ignored copy ``oracle/_ref/`` that ``oracle/vendor_ref.py`` (run by
``__graft_entry__.build()``) makes so that the real files travel to the GPU box.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Used by
``tests/golden/make_golden.py`` to generate the committed fixtures, by the tests that pin
``oracle/vectorbase_oracle.py`` and ``install()`` against the real thing, and by
``bench.py --impl reference`` / the ``cpu_baseline`` leg as the timed CPU comparator
(``kind: "reference"``).  The product package never imports it.

The reference package cannot be imported whole: ``typeagent/__init__.py:6`` pulls in
``knowpro.factory`` (needs ``typechat``) and ``aitools/vectorbase.py:14`` imports
``.model_adapters`` (needs ``pydantic_ai``/``stamina``), none of which are installed and there
is no network.  So bare package objects are pre-registered in ``sys.modules`` (their
``__path__`` pointing into the reference tree), ``model_adapters`` / ``typechat`` / ``stamina``
are stubbed, and the reference files are then imported as they lie on disk, byte for byte
(SURVEY.md §8c).
"""

from __future__ import annotations

import importlib
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
MOUNTED_SRC = "/root/reference/src/typeagent"
VENDORED_SRC = os.path.join(_HERE, "_ref", "typeagent")

_PACKAGES = ("", "aitools", "knowpro", "storage", "storage/memory", "storage/sqlite")


def reference_root() -> str | None:
    """Directory of the reference's ``typeagent`` package: the mounted tree if present, else
    the vendored copy, else None."""
    for root in (MOUNTED_SRC, VENDORED_SRC):
        if os.path.isfile(os.path.join(root, "aitools", "vectorbase.py")):
            return root
    return None


def reference_available() -> bool:
    return reference_root() is not None


def reference_kind() -> str:
    root = reference_root()
    return "mounted" if root == MOUNTED_SRC else ("vendored" if root else "absent")


class NullEmbeddingModel:
    """Embedding model that can cache but never embeds (cf. the reference's own
    benchmark stub, tools/benchmark_vectorbase.py:28-46)."""

    model_name = "oracle-null"

    def __init__(self) -> None:
        self.cache: dict[str, object] = {}

    def add_embedding(self, key, embedding) -> None:
        self.cache[key] = embedding

    async def get_embedding(self, key):
        return self.cache[key]

    async def get_embeddings(self, keys):
        import numpy as np

        return np.array([self.cache[k] for k in keys], dtype=np.float32)

    get_embedding_nocache = get_embedding
    get_embeddings_nocache = get_embeddings


class _Anything:
    """Stands in for any class of a stubbed third-party module (subscriptable, callable)."""

    def __class_getitem__(cls, item):
        return cls

    def __init__(self, *a, **k):
        pass


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything


_loaded: tuple[types.ModuleType, types.ModuleType] | None = None


def _register_packages() -> str:
    root = reference_root()
    if root is None:
        raise RuntimeError("reference sources are neither mounted at /root/reference nor vendored "
                           "under oracle/_ref (run __graft_entry__.build() in the build container)")
    for sub in _PACKAGES:
        name = "typeagent" + ("." + sub.replace("/", ".") if sub else "")
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [os.path.join(root, sub) if sub else root]
            sys.modules[name] = pkg
    if "typeagent.aitools.model_adapters" not in sys.modules:
        stub = types.ModuleType("typeagent.aitools.model_adapters")

        def create_embedding_model(*_a, **_k):
            raise RuntimeError("the oracle never creates a network embedding model")

        stub.create_embedding_model = create_embedding_model
        sys.modules["typeagent.aitools.model_adapters"] = stub
    for third_party in ("typechat", "stamina"):  # LLM / retry plumbing, never executed here
        if third_party not in sys.modules:
            sys.modules[third_party] = _StubModule(third_party)
    return root


def load_reference() -> tuple[types.ModuleType, types.ModuleType]:
    """Return ``(vectorbase_module, fuzzyindex_module)`` of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    _register_packages()
    vb = importlib.import_module("typeagent.aitools.vectorbase")
    fz = importlib.import_module("typeagent.knowpro.fuzzyindex")
    _loaded = (vb, fz)
    return _loaded


def load_reference_module(name: str) -> types.ModuleType:
    """Any other reference module of the lookup path's callers, e.g.
    ``typeagent.storage.memory.reltermsindex`` — unmodified, imported in place."""
    load_reference()
    return importlib.import_module(name)


def make_reference_vectorbase(vectors=None, **settings_kw):
    """A reference ``VectorBase`` over ``vectors`` (float32 [N, D]) with a null model."""
    vb, _ = load_reference()
    settings = vb.TextEmbeddingIndexSettings(
        embedding_model=NullEmbeddingModel(), **settings_kw
    )
    base = vb.VectorBase(settings)
    if vectors is not None:
        base.add_embeddings(None, vectors)
    return base
