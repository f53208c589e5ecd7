"""Axis indices of vector fields (same contract as sopht/utils/field.py:8-39)."""


class VectorField:
    """Vector fields are stored (dim, [nz,] ny, nx) with x = 0, y = 1, z = 2."""

    @staticmethod
    def x_axis_idx() -> int:
        return 0

    @staticmethod
    def y_axis_idx() -> int:
        return 1

    @staticmethod
    def z_axis_idx() -> int:
        return 2
