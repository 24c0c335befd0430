"""Import-path alias: the reference keeps Strsym in substitution_models/strsym.py."""
from .dna import Strsym  # noqa: F401
