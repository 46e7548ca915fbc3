"""Drop-in for the reference's ``nomad_audio`` package (``src/nomad_audio/__init__.py:1-2``).

``from nomad_audio import nomad`` yields a ready ``Nomad`` singleton exactly like the reference; it is
built on first access (PEP 562) instead of at import so that importing the package needs no GPU.
"""
from nomad_b200.nomad import Nomad  # noqa: F401

_singleton = None


def __getattr__(name):
    global _singleton
    if name == "nomad":
        if _singleton is None:
            _singleton = Nomad()
        return _singleton
    raise AttributeError(f"module 'nomad_audio' has no attribute {name!r}")
