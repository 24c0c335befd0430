"""Import-path alias: the reference keeps HKY85 in substitution_models/hky85.py."""
from .dna import HKY85  # noqa: F401
