"""Mirror of `timbre_trap.framework` (reference: timbre_trap/framework/__init__.py:1-4)."""

from .cqt import CQT

__all__ = ['CQT']
