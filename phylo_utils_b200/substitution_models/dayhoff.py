"""Import-path alias: the reference keeps Dayhoff in substitution_models/dayhoff.py."""
from .protein import Dayhoff  # noqa: F401
