"""Import-path alias: ``from topoflow_glacier import BmiTopoflowGlacier`` keeps working.

The reference package is called ``topoflow_glacier`` (reference ``src/topoflow_glacier/__init__.py``); drivers
written against it import from that path.  Everything here re-exports ``topoflow_glacier_b200``.
"""

from topoflow_glacier_b200 import __version__, configure_logging, logger  # noqa: F401

__all__ = ["__version__", "BmiTopoflowGlacier", "configure_logging", "logger"]


def __getattr__(name):
    if name == "BmiTopoflowGlacier":
        from topoflow_glacier_b200.bmi import BmiTopoflowGlacier

        return BmiTopoflowGlacier
    raise AttributeError(name)
