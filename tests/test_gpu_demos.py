"""Every demo runs to completion on the GPU (the reference's demos/test_demos.py)."""

import pathlib
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
DEMOS = pathlib.Path(__file__).resolve().parent.parent / "demos"


@pytest.mark.parametrize("script,args", [
    ("demo_Y_bifurcation.py", []), ("demo_double_Y_bifurcation.py", []), ("demo_tree.py", []),
    ("demo_arterial_tree.py", []), ("demo_perf.py", ["3", "6", "12"]),
])
def test_demo(script, args, tmp_path):
    out = subprocess.run([sys.executable, str(DEMOS / script), *args], capture_output=True, text=True,
                         timeout=600, cwd=tmp_path, env={**__import__("os").environ, "PYTHONPATH": str(DEMOS.parent)})
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    if script == "demo_Y_bifurcation.py":
        assert "0.77485177" in out.stdout
