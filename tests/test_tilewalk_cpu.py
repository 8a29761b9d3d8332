"""CPU test of host/device-shared logic: the tile GEMM's work enumeration (csrc/tile_walk.cuh) is compiled for the host with
nvcc and checked exhaustively over job shapes (triangular, rectangular, block-column-cyclic) and chunk sizes."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_tile_enumeration_covers_every_half_tile_once(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        pytest.skip("nvcc not available")
    exe = str(tmp_path / "tilewalk_host")
    src = os.path.join(ROOT, "tests", "cpp", "tilewalk_host.cu")
    subprocess.run([nvcc, "-O1", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src], check=True,
                   capture_output=True, text=True)
    out = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-2000:]
    assert "0 failures" in out.stdout
