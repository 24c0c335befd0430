"""Import-path alias: the reference keeps TN93 in substitution_models/tn93.py."""
from .dna import TN93  # noqa: F401
