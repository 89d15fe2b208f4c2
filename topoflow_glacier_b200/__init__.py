"""B200-native implementation of topoflow-glacier's energy-balance + melt update.

Public surface (mirrors reference ``src/topoflow_glacier/__init__.py:1-5``):
``BmiTopoflowGlacier``, ``configure_logging``, ``logger``, ``__version__`` -- plus ``MeltEngine``, the
N-cell device-resident core the BMI class is built on.
"""

from .logger import configure_logging, logger

__version__ = "0.1.0"
__all__ = ["__version__", "BmiTopoflowGlacier", "MeltEngine", "configure_logging", "logger"]


def __getattr__(name):  # torch + the CUDA library are only needed once a model is actually used
    if name == "BmiTopoflowGlacier":
        from .bmi import BmiTopoflowGlacier

        return BmiTopoflowGlacier
    if name == "MeltEngine":
        from .engine import MeltEngine

        return MeltEngine
    raise AttributeError(name)
