"""sopht_b200: B200 (sm_100a) implementation of SophT's Eulerian flow time step.

Hand-written CUDA kernels behind the reference's own factory API; see DESIGN.md.
"""

import importlib
import sys

__version__ = "0.2.0"

# the sub-packages that mirror the reference's layout (sopht/__init__.py, sopht/numeric/__init__.py, ...)
_MIRRORED = (
    "numeric",
    "numeric.eulerian_grid_ops",
    "numeric.immersed_boundary_ops",
    "simulator",
    "simulator.flow",
    "simulator.immersed_body",
    "utils",
    "utils.field",
    "utils.precision",
)


def install_as_sopht(force: bool = False) -> None:
    """Make ``import sopht...`` resolve to this package for the hot-path sub-packages, so that scripts and tests written
    against the reference (``import sopht.numeric.eulerian_grid_ops as spne``, ``import sopht.simulator as sps``,
    ``import sopht.utils as spu``) run on the B200 path without edits. Only the mirrored sub-packages are aliased
    (numeric, simulator, utils.field / utils.precision); plotting, HDF5 I/O and the other out-of-scope utilities of the
    reference are not provided. Refuses to shadow an importable real ``sopht`` unless ``force`` is set."""
    if "sopht" in sys.modules and not force:
        if getattr(sys.modules["sopht"], "__sopht_b200_alias__", False):
            return
        raise RuntimeError("a module named 'sopht' is already imported; pass force=True to replace it")
    if not force and importlib.util.find_spec("sopht") is not None:
        raise RuntimeError("the reference package 'sopht' is installed; pass force=True to shadow it")
    root = sys.modules[__name__]
    root.__sopht_b200_alias__ = True
    sys.modules["sopht"] = root
    for name in _MIRRORED:
        sys.modules["sopht." + name] = importlib.import_module(__name__ + "." + name)
