"""Reference-named module: ``from sgrl_b200.SECritic import SECritic`` mirrors ``from SECritic import SECritic``."""
from .modules import SECritic  # noqa: F401
