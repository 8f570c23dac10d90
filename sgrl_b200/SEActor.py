"""Reference-named module: ``from sgrl_b200.SEActor import SEPolicy`` mirrors ``from SEActor import SEPolicy``."""
from .modules import SEPolicy  # noqa: F401
