"""Mirror of `timbre_trap.framework` (reference: timbre_trap/framework/__init__.py:1-4)."""

from .cqt import CQT
from .modules import *          # noqa: F401,F403
from .modules import __all__ as _modules_all

__all__ = ['CQT'] + list(_modules_all)
