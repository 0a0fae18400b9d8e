"""Same import surface as the reference's ``generators`` package (generators/generators.py,
generators/siren.py, generators/volumetric_rendering.py)."""
from . import siren, volumetric_rendering  # noqa: F401
from .generators import ImplicitGenerator3d  # noqa: F401
from .volumetric_rendering import *  # noqa: F401,F403
