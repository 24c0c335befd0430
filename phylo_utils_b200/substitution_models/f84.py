"""Import-path alias: the reference keeps F84 in substitution_models/f84.py."""
from .dna import F84  # noqa: F401
