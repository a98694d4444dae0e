"""Stub module (see the package docstring)."""
