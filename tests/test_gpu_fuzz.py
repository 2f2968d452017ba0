"""Short randomised differential campaign (tools/fuzz.py): random codes, caps, erasure points, frame counts,
expurgation thresholds, engines and word widths; CUDA path vs the oracle, bit for bit."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_randomised_differential_campaign(built):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz.py"), "25", "20261018"],
                         capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-2000:]
    assert "0 failures" in res.stdout
