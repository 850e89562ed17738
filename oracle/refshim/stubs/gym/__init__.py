"""Minimal stand-in for ``gym`` (absent from the image) so the reference Python imports. ORACLE ONLY."""
from . import envs  # noqa: F401

_registry = {}


class Env(object):
    metadata = {}


def make(env_id):
    module, cls = _registry[env_id].split(":")
    import importlib
    return getattr(importlib.import_module(module), cls)()
