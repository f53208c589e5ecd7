"""CPU: the C-ABI shared library loads and exports every symbol include/sopht_b200.h declares
(no compute calls here — there is no GPU in the build container)."""

import ctypes
import os
import re

import pytest

from sopht_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "sopht_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sopht_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) > 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing


def test_python_binding_covers_the_header():
    assert sorted(_lib.exported_symbols()) == _declared_symbols()


def test_version_and_error_string():
    lib = _lib.load()
    assert lib.sopht_version() >= 1
    assert isinstance(lib.sopht_last_error(), bytes)


def test_bad_arguments_return_status_not_abort():
    lib = _lib.load()
    f = _lib.SophtField()
    # null data pointer -> SOPHT_ERR_SHAPE, invalid dtype -> SOPHT_ERR_DTYPE; neither touches the GPU
    assert lib.sopht_set_fixed_val(7, ctypes.byref(f), 0.0, None) == -1
    assert lib.sopht_diffusion_flux_3d(0, ctypes.byref(f), ctypes.byref(f), 1.0, 1, None) == -2
    assert b"diffusion" in lib.sopht_last_error()


def test_no_cpu_fallback():
    import numpy as np
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from sopht_b200.numeric.eulerian_grid_ops import gen_set_fixed_val_pyst_kernel_3d

    k = gen_set_fixed_val_pyst_kernel_3d(np.float32)
    with pytest.raises(_lib.SophtLibraryError):
        k(field=np.zeros((4, 4, 4), np.float32), fixed_val=1.0)


def test_install_as_sopht_aliases_the_reference_import_paths():
    """`import sopht...` of a script written against the reference resolves to this package (sopht/__init__.py,
    sopht/numeric/eulerian_grid_ops/__init__.py:3-133 export lists)."""
    import importlib
    import sys

    import sopht_b200

    saved = {k: v for k, v in sys.modules.items() if k == "sopht" or k.startswith("sopht.")}
    try:
        sopht_b200.install_as_sopht(force=True)
        spne = importlib.import_module("sopht.numeric.eulerian_grid_ops")
        spnib = importlib.import_module("sopht.numeric.immersed_boundary_ops")
        sps = importlib.import_module("sopht.simulator")
        spu = importlib.import_module("sopht.utils")
        assert callable(spne.gen_diffusion_timestep_euler_forward_pyst_kernel_3d)
        assert hasattr(spne, "UnboundedPoissonSolverPYFFTW3D") and hasattr(spne, "FFTPyFFTW2D")
        assert hasattr(spnib, "VirtualBoundaryForcing") and hasattr(spnib, "EulerianLagrangianGridCommunicator3D")
        assert hasattr(sps, "UnboundedNavierStokesFlowSimulator3D") and hasattr(sps, "CosseratRodFlowInteraction")
        assert spu.get_real_t("single").__name__ == "float32" and hasattr(spu, "VectorField")
        sopht_b200.install_as_sopht()  # idempotent
    finally:
        for k in [k for k in sys.modules if k == "sopht" or k.startswith("sopht.")]:
            del sys.modules[k]
        sys.modules.update(saved)
