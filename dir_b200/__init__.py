"""dir_b200 — B200-native implementation of DIR's eval-mode forward (one hot path)."""
from .capi import DirB200Error, Handle, load_library  # noqa: F401
from .module import DIR  # noqa: F401

__all__ = ["DIR", "Handle", "DirB200Error", "load_library"]
