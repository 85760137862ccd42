"""CPU checks of the tile kernel's packed arithmetic (ie_set_option("encode_variant", 1|2); 2 is the default kernel).

The variants change only per-lane register arithmetic of the tile kernel (imageencoder_b200/csrc/transform_fast.cuh, namespace
ie::lean); that code is host+device, so tests/host/lean_check.cu runs the very same C++ on the CPU against a transcription of
the default path: staged coefficients, guard-band mask, max bits_needed, segment flags and (variant 2) the packed transform's
outputs bit for bit."""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_lean_variant_arithmetic_matches_default_path(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    exe = tmp_path / "lean_check"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "--expt-relaxed-constexpr", "-Xcompiler",
           "-ffp-contract=off,-fno-fast-math", "-o", str(exe), str(ROOT / "tests" / "host" / "lean_check.cu")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe), "60000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "lean_check: ok" in r.stdout


def test_encode_variant_option_is_validated():
    import imageencoder_b200 as ie
    L = ie.lib()
    try:
        assert L.ie_set_option(b"encode_variant", 1) == 0
        assert L.ie_set_option(b"encode_variant", 2) == 0
        assert L.ie_set_option(b"encode_variant", 3) == 0
        assert L.ie_set_option(b"encode_variant", 4) == 0
        assert L.ie_set_option(b"encode_variant", 7) == 0
        assert L.ie_set_option(b"encode_variant", 8) == 0          # fused stream kernel (encode_fused.cu)
        assert L.ie_set_option(b"encode_variant", 9) == 0          # in-warp exact evaluation
        assert L.ie_set_option(b"encode_variant", 10) != 0
        assert L.ie_set_option(b"encode_variant", -1) != 0
    finally:
        assert L.ie_set_option(b"encode_variant", 2) == 0          # the default


def test_decode_variant_option_is_validated():
    import imageencoder_b200 as ie
    L = ie.lib()
    try:
        assert L.ie_set_option(b"decode_variant", 1) == 0
        assert L.ie_set_option(b"decode_variant", 2) != 0
        assert L.ie_set_option(b"decode_variant", -1) != 0
    finally:
        assert L.ie_set_option(b"decode_variant", 1) == 0          # the default


def test_me_variant_option_is_validated():
    import imageencoder_b200 as ie
    L = ie.lib()
    try:
        assert L.ie_set_option(b"me_variant", 1) == 0
        assert L.ie_set_option(b"me_variant", 0) == 0
        assert L.ie_set_option(b"me_variant", 3) != 0
        assert L.ie_set_option(b"me_variant", -1) != 0
    finally:
        assert L.ie_set_option(b"me_variant", 2) == 0              # the default


def test_copyout_variant_option_is_validated():
    import imageencoder_b200 as ie
    L = ie.lib()
    try:
        for v in (0, 1, 2, 3):
            assert L.ie_set_option(b"copyout_variant", v) == 0
        assert L.ie_set_option(b"copyout_variant", 4) != 0
        assert L.ie_set_option(b"copyout_variant", -1) != 0
    finally:
        assert L.ie_set_option(b"copyout_variant", 3) == 0         # the default
