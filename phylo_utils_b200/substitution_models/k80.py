"""Import-path alias: the reference keeps K80 in substitution_models/k80.py."""
from .dna import K80  # noqa: F401
