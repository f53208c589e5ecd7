"""CPU: the pruned FFT Poisson pipeline's kernels (csrc/poisson_pow2_phases.cuh) emulated thread by thread
on the host against a double-precision doubled-domain convolution (tests/host/fft_emul.cu). This pins the
index arithmetic of the CUDA kernels (digit-reversed spectrum order, Hermitian pre/post-processing, folded
Green's function, Nyquist plane) without a GPU. `build/fft_emul full` runs every transform length (minutes)."""

import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_fft_pipeline_emulation(tmp_path):
    exe = tmp_path / "fft_emul"
    subprocess.run(
        ["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Isopht_b200/csrc", "-Iinclude",
         "tests/host/fft_emul.cu", "-o", str(exe)], cwd=ROOT, check=True)
    out = subprocess.run([str(exe)], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ALL OK" in out.stdout
