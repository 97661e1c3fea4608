"""the engine plans a model from module class NAMES and a fixed set of attributes, so it accepts the
reference's own `models.yolo.Model` object as well as the mirror.  Checked where the reference exists (the authoring
container): the engine's view of the reference Model and of the mirror built from the same yaml is identical attribute by
attribute (n / s / m), and the reference state_dict loads strictly into the mirror.  Runs the check in a subprocess because
importing the reference package rearranges sys.path and the working directory (tests/golden/ref_import.py)."""
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.skipif(not Path("/root/reference/models/yolo.py").exists(), reason="the reference is only mounted in the authoring container")
def test_engine_view_of_reference_model_equals_mirror():
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "golden" / "check_ref_model_compat.py")],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "COMPAT OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
