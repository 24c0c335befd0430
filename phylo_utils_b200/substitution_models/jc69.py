"""Import-path alias: the reference keeps JC69 in substitution_models/jc69.py."""
from .dna import JC69  # noqa: F401
