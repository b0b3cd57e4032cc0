"""Mirror of `timbre_trap.framework` (reference: timbre_trap/framework/__init__.py:1-4)."""

from .objectives import *       # noqa: F401,F403
from .objectives import __all__ as _objectives_all
from .cqt import CQT
from .modules import *          # noqa: F401,F403
from .modules import __all__ as _modules_all
from .train import TrainStep, compute_step_losses
from .pipeline import HostPipeline
from .frontend import multi_pitch_to_activations, prepare_audio, signal_distortion_ratio

__all__ = list(_objectives_all) + ['CQT', 'compute_step_losses', 'TrainStep', 'HostPipeline', 'prepare_audio', 'multi_pitch_to_activations',
                                   'signal_distortion_ratio'] + list(_modules_all)
