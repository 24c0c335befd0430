"""Import-path alias: the reference keeps Unrest in substitution_models/unrest.py."""
from .dna import Unrest  # noqa: F401
