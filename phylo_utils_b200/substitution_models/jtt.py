"""Import-path alias: the reference keeps JTT in substitution_models/jtt.py."""
from .protein import JTT  # noqa: F401
