"""Experimental instantiations of the tile kernel (ie_set_option("encode_variant", 1|2)): byte-identical streams required.

The variants were written after this round's GPU budget was spent: their arithmetic is checked on the CPU
(tests/test_lean_variant_cpu.py) and their SASS was inspected, but they have not yet run on a B200.  They are NOT the default
path.  Until their first confirmed GPU run these tests are xfail(strict=False) -- a pass shows up as XPASS -- and each runs in
its own process so that nothing it does can affect the parity tests of the default path.  Remove the xfail once confirmed."""
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.mark.xfail(strict=False, reason="experimental kernel variant, not yet confirmed on hardware (not the default path)")
@pytest.mark.parametrize("variant", [1, 2])
def test_variant_streams_identical(variant):
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "_variant_worker.py"), str(variant)], capture_output=True, text=True,
                       timeout=600)
    print(r.stdout[-3000:], r.stderr[-3000:])
    assert r.returncode == 0
