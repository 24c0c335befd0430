"""Import-path alias: the reference keeps LG in substitution_models/lg.py."""
from .protein import LG  # noqa: F401
