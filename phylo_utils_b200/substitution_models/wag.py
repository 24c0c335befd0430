"""Import-path alias: the reference keeps WAG in substitution_models/wag.py."""
from .protein import WAG  # noqa: F401
