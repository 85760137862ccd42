"""CPU tests of the boundary: the C-ABI library loads here (no GPU), exports every symbol include/*.h declares, and
fails loudly -- never falls back -- when asked to compute without a device.  Also the CLI's pre-GPU error behaviour."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "imageencoder_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return set(re.findall(r"\b(ie_[a-z0-9_]+)\s*\(", text))


def test_every_declared_symbol_is_exported_and_bound():
    import imageencoder_b200 as ie
    from imageencoder_b200 import _lib
    L = ie.lib()
    declared = _declared()
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/ but not exported"
    assert declared == set(_lib.SIGNATURES), "ctypes table and header disagree"
    nm = subprocess.run(["nm", "-D", "--defined-only", str(ie.lib_path())], capture_output=True, text=True).stdout
    exported = set(re.findall(r"\bT (ie_[a-z0-9_]+)", nm))
    assert declared <= exported


def test_no_oracle_or_cpu_fallback_in_product():
    """the product tree must not reference the oracle"""
    for p in list((ROOT / "imageencoder_b200").rglob("*.py")) + list((ROOT / "imageencoder_b200" / "csrc").rglob("*.c*")):
        txt = p.read_text()
        assert "import oracle" not in txt and "liboracle" not in txt and "oracle/" not in txt.replace("oracle/_ref", ""), p


def test_version_and_sizing():
    import imageencoder_b200 as ie
    L = ie.lib()
    assert b"sm_100a" in L.ie_version()
    assert L.ie_max_encoded_bytes(8192, 8192, 8, 1) >= (1024 * 1024 * (4 + 16 + 16 * 64)) // 8
    assert L.ie_max_encoded_bytes(8, 8, 3, 1) == 0


def test_fails_loudly_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import imageencoder_b200 as ie
    with pytest.raises(ie.IEError) as e:
        ie.encode_image(np.zeros(64, np.uint8), 8, 8, np.full(16, 2))
    assert e.value.code == -2
    with pytest.raises(ie.IEError):
        ie.decode_image(b"\x00" * 32, 4)


def test_cli_exit_codes(tmp_path):
    enc = ROOT / "bin" / "encoder"
    if not enc.exists():
        pytest.skip("CLIs not built")
    run = lambda *a: subprocess.run([str(enc), *a], capture_output=True, text=True, cwd=tmp_path).returncode
    assert run() == 1                                            # main.cpp:20-23
    assert run("missing.conf") == 2                              # main.cpp:27-31
    (tmp_path / "bad.conf").write_text("rawfile=a.raw\nencfile=a.enc\n")
    assert run("bad.conf") == 3                                  # main.cpp:47-52
    conf = "rawfile=a.raw\nencfile=a.enc\ndecfile=a_dec.raw\nwidth=8\nheight=8\nrle=1\nquantfile=q.txt\nlogfile=a.txt\n"
    (tmp_path / "a.conf").write_text(conf)
    assert run("a.conf") == 4                                    # quant matrix unreadable, main.cpp:80-82
    (tmp_path / "q.txt").write_text("1 2 3 4\n1 2 3 4\n1 2 3 4\n1 2 3 4\n")
    (tmp_path / "w.conf").write_text(conf.replace("width=8", "width=eight"))
    assert run("w.conf") == 5                                    # main.cpp:99-102
