"""Load the UNMODIFIED reference ``vectorbase.py`` / ``fuzzyindex.py`` from /root/reference.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  Works only in the build
container, where ``/root/reference`` is mounted; the GPU box has no reference,
so nothing that runs there (``-m gpu`` tests, ``smoke()``, ``bench.py``) may call
this module.  It is used by ``tests/golden/make_golden.py`` to generate the
committed fixtures and by the ``not gpu`` tests (skipped when the reference is
absent) to pin ``oracle/vectorbase_oracle.py`` against the real thing.

The reference package cannot be imported whole: ``typeagent/__init__.py:6``
pulls in ``knowpro.factory`` (needs ``typechat``) and ``aitools/vectorbase.py:14``
imports ``.model_adapters`` (needs ``pydantic_ai``/``stamina``), none of which are
installed and there is no network.  So three stub packages are pre-registered in
``sys.modules`` and the two reference files are then imported as they lie on
disk, byte for byte.
"""

from __future__ import annotations

import importlib
import os
import sys
import types

REFERENCE_SRC = "/root/reference/src/typeagent"


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SRC, "aitools", "vectorbase.py"))


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


_loaded: tuple[types.ModuleType, types.ModuleType] | None = None


def load_reference() -> tuple[types.ModuleType, types.ModuleType]:
    """Return ``(vectorbase_module, fuzzyindex_module)`` of the reference."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not reference_available():
        raise RuntimeError("reference sources are not mounted at /root/reference")
    for name, sub in (
        ("typeagent", ""),
        ("typeagent.aitools", "/aitools"),
        ("typeagent.knowpro", "/knowpro"),
    ):
        if name not in sys.modules:
            pkg = types.ModuleType(name)
            pkg.__path__ = [REFERENCE_SRC + sub]
            sys.modules[name] = pkg
    if "typeagent.aitools.model_adapters" not in sys.modules:
        stub = types.ModuleType("typeagent.aitools.model_adapters")

        def create_embedding_model(*_a, **_k):
            raise RuntimeError("the oracle never creates a network embedding model")

        stub.create_embedding_model = create_embedding_model
        sys.modules["typeagent.aitools.model_adapters"] = stub
    vb = importlib.import_module("typeagent.aitools.vectorbase")
    fz = importlib.import_module("typeagent.knowpro.fuzzyindex")
    _loaded = (vb, fz)
    return _loaded


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
