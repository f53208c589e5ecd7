"""Utilities kept from the reference's sopht.utils that the hot path needs."""

from .field import VectorField
from .precision import get_real_t, get_test_tol

__all__ = ["VectorField", "get_real_t", "get_test_tol"]
