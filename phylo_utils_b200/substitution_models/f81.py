"""Import-path alias: the reference keeps F81 in substitution_models/f81.py."""
from .dna import F81  # noqa: F401
