"""Names of the two supported precisions -> numpy scalar types and test tolerances.

Public names and behaviour follow sopht/utils/precision.py:6-19 (`get_real_t`, `get_test_tol`); the table-driven
form also serves `_lib.dtype_code` style lookups elsewhere in the package.
"""

import numpy as np

_REAL_TYPES = {"single": np.float32, "double": np.float64}
_TOL_IN_EPS = 1e3  # test tolerance = this many machine epsilons of the chosen type


def get_real_t(precision: str = "single") -> type:
    try:
        return _REAL_TYPES[precision]
    except KeyError:
        raise ValueError("Precision argument must be single or double") from None


def get_test_tol(precision: str = "single") -> float:
    kind = get_real_t(precision)
    return kind(_TOL_IN_EPS) * np.finfo(kind).eps
