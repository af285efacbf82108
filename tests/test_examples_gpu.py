"""The example scripts (written against `import xlb` / `import warp as wp` / `import jax.numpy as jnp` like the
reference's own examples) run end to end on the GPU."""

import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run(*args):
    proc = subprocess.run([sys.executable, *args], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert proc.returncode == 0, proc.stdout + proc.stderr
    return proc.stdout


@pytest.mark.parametrize("backend,precision", [("warp", "fp32/fp32"), ("jax", "fp32/fp16")])
def test_cavity_mlups_script(backend, precision):
    out = run("examples/cavity_mlups.py", "96", "50", backend, precision)
    assert "MLUPs:" in out and float(out.split("MLUPs:")[1].split()[0]) > 100.0


def test_sphere_kbc_script(tmp_path):
    env_prefix = str(tmp_path / "umag")
    os.environ["XLB_OUT_PREFIX"] = env_prefix
    out = run("examples/sphere_kbc.py", "128", "32", "32", "200")
    assert "MLUPS" in out and os.path.exists(env_prefix + "_0000200.pgm")


@pytest.mark.parametrize("collision", ["BGK", "KBC"])
def test_cavity_2d_script(tmp_path, collision):
    os.environ["XLB_OUT_PREFIX"] = str(tmp_path / "c2d")
    out = run("examples/cavity_2d.py", "128", "2000", collision)
    assert "max |u|" in out and os.path.exists(str(tmp_path / "c2d") + "_0002000.pgm")
