"""Fused convolution Laplacian filter (csrc/filter3d.cu) against the oracle's pass-by-pass
restatement of laplacian_filter_3d.py:129-163 on grids that exercise partial x tiles, several line segments and lines
shorter than a segment, every order up to the rod case's 5, scalar / vector / strided-view inputs, and against the
pass-by-pass CUDA composition at the C3 grid size."""

import numpy as np
import pytest

TOL = {"float32": 1e-5, "float64": 1e-12}  # north_star: per-kernel relative L2 error


def _rel_l2(a, b):
    return float(np.linalg.norm((np.asarray(a, dtype=np.float64) - b).ravel()) / np.linalg.norm(np.ravel(b)))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float64"])
@pytest.mark.parametrize("grid", [(17, 19, 23), (70, 9, 33), (5, 131, 64), (3, 3, 3), (66, 65, 40)])
@pytest.mark.parametrize("order", [1, 2, 5, 8, 9])  # <= 8: register pipeline along y / z; 9: shared-memory segments
def test_fused_convolution_filter_matches_oracle(dtype, grid, order):
    import torch

    import sopht_b200.numeric.eulerian_grid_ops as spne
    from oracle import stencils as ost

    rng = np.random.default_rng(11)
    real_t = np.float32 if dtype == "float32" else np.float64
    f = rng.standard_normal(grid).astype(real_t)
    ref = f.astype(np.float64)
    ost.laplacian_filter_3d(ref, np.ones(grid), np.zeros(grid), order, "convolution")
    dev = torch.from_numpy(f).cuda()
    flux = torch.full(grid, 7.0, dtype=dev.dtype, device="cuda")  # stale scratch contents must not matter
    buf = torch.full(grid, -3.0, dtype=dev.dtype, device="cuda")
    kernel = spne.gen_laplacian_filter_kernel_3d(filter_order=order, filter_flux_buffer=flux, field_buffer=buf,
                                                 real_t=real_t, field_type="scalar", filter_type="convolution")
    kernel(scalar_field=dev)
    assert _rel_l2(dev.cpu().numpy(), ref) < TOL[dtype]
    # the ring never changes
    out = dev.cpu().numpy()
    for axis in range(3):
        np.testing.assert_array_equal(np.take(out, [0, -1], axis=axis), np.take(f, [0, -1], axis=axis))


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_fused_convolution_filter_vector_and_views(dtype):
    import torch

    import sopht_b200.numeric.eulerian_grid_ops as spne
    from oracle import stencils as ost

    rng = np.random.default_rng(12)
    real_t = np.float32 if dtype == "float32" else np.float64
    grid = (34, 21, 45)
    v = rng.standard_normal((3, *grid)).astype(real_t)
    ref = v.astype(np.float64)
    ost.laplacian_filter_3d_vector(ref, np.zeros(grid), np.zeros(grid), 3, "convolution")
    # the vector field is a strided view of a bigger allocation (padded rows), like a slab of a decomposed grid
    big = torch.zeros((3, grid[0], grid[1] + 2, grid[2] + 3), dtype=getattr(torch, dtype), device="cuda")
    view = big[:, :, 1:-1, 2:-1]
    view.copy_(torch.from_numpy(v))
    scratch = torch.zeros((2, *grid), dtype=view.dtype, device="cuda")
    kernel = spne.gen_laplacian_filter_kernel_3d(filter_order=3, filter_flux_buffer=scratch[0],
                                                 field_buffer=scratch[1], real_t=real_t, field_type="vector",
                                                 filter_type="convolution")
    kernel(vector_field=view)
    assert _rel_l2(view.cpu().numpy(), ref) < TOL[dtype]
    assert float(big[:, :, 0].abs().max()) == 0.0 and float(big[..., :2].abs().max()) == 0.0  # padding untouched
    # numpy in / out is staged like everywhere else
    v_np = v.copy()
    spne.gen_laplacian_filter_kernel_3d(filter_order=3, filter_flux_buffer=np.zeros(grid, dtype=real_t),
                                        field_buffer=np.zeros(grid, dtype=real_t), real_t=real_t,
                                        field_type="vector", filter_type="convolution")(vector_field=v_np)
    assert _rel_l2(v_np, ref) < TOL[dtype]


@pytest.mark.gpu
def test_fused_convolution_filter_at_rod_case_size():
    """256 x 128 x 128 (BASELINE config 3), order 5: fused kernels vs the pass-by-pass composition of the public
    primitive kernels, plus the constant-field fixed point of test_laplacian_filter_3d.py:102-131."""
    import torch

    from sopht_b200 import _lib
    import sopht_b200.numeric.eulerian_grid_ops as spne

    grid = (256, 128, 128)
    gen = torch.Generator(device="cuda").manual_seed(5)
    v = torch.randn((3, *grid), device="cuda", generator=gen)
    ref = v.clone()
    flux, buf = torch.zeros(grid, device="cuda"), torch.zeros(grid, device="cuda")
    dt = _lib.dtype_code(np.float32)
    _lib.call("sopht_set_fixed_val_at_boundaries", dt, flux, 1, [0.0], 0)
    for c in range(3):
        for axis in (0, 1, 2):
            _lib.call("sopht_elementwise_copy", dt, buf, ref[c])
            for _ in range(5):
                _lib.call("sopht_laplacian_filter_flux_3d", dt, flux, buf, axis)
                _lib.call("sopht_elementwise_copy", dt, buf, flux)
            _lib.call("sopht_elementwise_saxpby", dt, ref[c], ref[c], flux, 1.0, -1.0)
    kernel = spne.gen_laplacian_filter_kernel_3d(filter_order=5, filter_flux_buffer=flux, field_buffer=buf,
                                                 real_t=np.float32, field_type="vector", filter_type="convolution")
    kernel(vector_field=v)
    err = float(torch.linalg.vector_norm((v - ref).double()) / torch.linalg.vector_norm(ref.double()))
    assert err < 1e-6
    const = torch.full(grid, 3.0, device="cuda")
    spne.gen_laplacian_filter_kernel_3d(filter_order=5, filter_flux_buffer=flux, field_buffer=buf, real_t=np.float32,
                                        field_type="scalar", filter_type="convolution")(scalar_field=const)
    inner = const[1:-1, 1:-1, 1:-1]
    # away from the walls a constant has no flux; next to a wall the masked stencil sees the zeroed ring of the flux
    assert float((inner[5:-5, 5:-5, 5:-5] - 3.0).abs().max()) == 0.0


def _one_pass_filter_emulation(f, order):
    """numpy emulation of what csrc/filter3d.cu evaluates per direction: away from the line ends one symmetric
    (2 order + 1)-tap filter of the originals, within `order` cells of an end the pass-by-pass values of a
    2 order-cell window; lines on the ring of the other two axes untouched."""
    taps = np.zeros(2 * order + 1)
    taps[order] = 1.0
    for _ in range(order):
        padded = np.pad(taps, 1)
        taps = (0.25 * (-padded[2:] - padded[:-2] + 2 * padded[1:-1]))
    out = f.astype(np.float64).copy()
    for axis in (2, 1, 0):
        src = np.moveaxis(out, axis, -1).copy()
        n = src.shape[-1]
        flux = np.zeros_like(src)
        if n >= 4 * order:
            for j in range(-order, order + 1):
                flux[..., order : n - order] += taps[order + j] * src[..., order + j : n - order + j]
            for window, flip in ((src[..., : 2 * order], False), (src[..., ::-1][..., : 2 * order], True)):
                u = window.copy()
                for _ in range(order):
                    nxt = np.zeros_like(u)
                    nxt[..., 1:-1] = 0.25 * (-u[..., 2:] - u[..., :-2] + 2 * u[..., 1:-1])
                    nxt[..., -1] = 0.25 * (-u[..., -2] + 2 * u[..., -1])  # beyond the window: treated as 0
                    u = nxt
                if flip:
                    flux[..., n - order :] = u[..., :order][..., ::-1]
                else:
                    flux[..., :order] = u[..., :order]
        else:  # short lines: pass by pass over the whole line
            u = src.copy()
            for _ in range(order):
                nxt = np.zeros_like(u)
                nxt[..., 1:-1] = 0.25 * (-u[..., 2:] - u[..., :-2] + 2 * u[..., 1:-1])
                u = nxt
            flux = u
        # no flux on lines that lie on the ring of the other two axes
        flux[0], flux[-1] = 0.0, 0.0
        flux[:, 0], flux[:, -1] = 0.0, 0.0
        out = np.moveaxis(src - flux, -1, axis)
    return out


@pytest.mark.parametrize("grid", [(17, 19, 23), (9, 30, 12), (24, 8, 21)])
@pytest.mark.parametrize("order", [1, 2, 5])
def test_one_pass_filter_algorithm_matches_oracle(grid, order):
    """CPU: the single-filter-plus-end-windows formulation of the CUDA kernels equals the reference's pass-by-pass
    closure (oracle restatement of laplacian_filter_3d.py:129-163)."""
    from oracle import stencils as ost

    rng = np.random.default_rng(21)
    f = rng.standard_normal(grid)
    ref = f.copy()
    ost.laplacian_filter_3d(ref, np.zeros(grid), np.zeros(grid), order, "convolution")
    np.testing.assert_allclose(_one_pass_filter_emulation(f, order), ref, rtol=0, atol=1e-13)
