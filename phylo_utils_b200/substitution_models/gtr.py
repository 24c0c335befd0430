"""Import-path alias: the reference keeps GTR in substitution_models/gtr.py."""
from .dna import GTR  # noqa: F401
