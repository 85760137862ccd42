"""The drop-in CLIs (`encoder <conf>` / `decoder <conf>`, same .conf keys as the reference, main.cpp:19-185) and the
Python mirror of the reference classes, against the golden vectors made with the reference itself."""
import hashlib
import json
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, INPUTS

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]
sha = lambda b: hashlib.sha256(bytes(b)).hexdigest()


@pytest.mark.parametrize("name,W,H", [("ex0", 8, 8), ("ex6", 512, 256)])
def test_cli_image_roundtrip(gpu, tmp_path, name, W, H):
    gold = json.loads((GOLDEN / "golden.json").read_text())["images"]
    shutil.copy(INPUTS / f"{name}.raw", tmp_path)
    shutil.copy(INPUTS / "matrix.txt", tmp_path)
    (tmp_path / f"{name}.conf").write_text(f"rawfile={name}.raw\nencfile={name}.enc\ndecfile={name}_dec.raw\nwidth={W}\nheight={H}\n"
                                           f"rle=1\nquantfile=matrix.txt\nlogfile={name}.txt\n")
    for flag, kind in (("--huffman", "huff"), ("--no-huffman", "plain")):
        r = subprocess.run([str(ROOT / "bin" / "encoder"), flag, f"{name}.conf"], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        assert "Elapsed time:" in r.stdout
        g = gold[f"{name}|matrix.txt|rle1|{kind}"]
        assert sha((tmp_path / f"{name}.enc").read_bytes()) == g["enc_sha256"]
        r = subprocess.run([str(ROOT / "bin" / "decoder"), f"{name}.conf"], cwd=tmp_path, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr + r.stdout
        assert sha((tmp_path / f"{name}_dec.raw").read_bytes()) == g["dec_sha256"]


def test_cli_video_roundtrip(gpu, oracle_mod, tmp_path):
    from imageencoder_b200.synth import synth_video
    W, H, F = 64, 48, 7
    synth_video(W, H, F, 4000).tofile(tmp_path / "v.yuv")
    shutil.copy(INPUTS / "matrix.txt", tmp_path)
    (tmp_path / "v.conf").write_text(f"rawfile=v.yuv\nencfile=v.enc\ndecfile=v_dec.yuv\nwidth={W}\nheight={H}\nrle=1\n"
                                     "quantfile=matrix.txt\nlogfile=v.txt\ngop=4\nmerange=16\nmotioncompensation=1\n")
    gold = json.loads((GOLDEN / "golden.json").read_text())["video"]["synth4000|64x48x7|gop4|mer16|huff"]
    r = subprocess.run([str(ROOT / "bin" / "encoder"), "v.conf"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert sha((tmp_path / "v.enc").read_bytes()) == gold["enc_sha256"]
    r = subprocess.run([str(ROOT / "bin" / "decoder"), "v.conf"], cwd=tmp_path, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr + r.stdout
    assert sha((tmp_path / "v_dec.yuv").read_bytes()) == gold["dec_mc1_sha256"]


def test_python_classes_like_main_cpp(gpu, tmp_path):
    """main.cpp:105-150 flow with the mirrored classes"""
    gold = json.loads((GOLDEN / "golden.json").read_text())["images"]["ex3|matrix.txt|rle1|huff"]
    m = gpu.read_matrix(INPUTS / "matrix.txt")
    enc = gpu.ImageEncoder(INPUTS / "ex3.raw", tmp_path / "ex3.enc", 400, 400, True, m, huffman=True)
    assert enc.process()
    enc.saveResult()
    assert sha((tmp_path / "ex3.enc").read_bytes()) == gold["enc_sha256"]
    dec = gpu.ImageDecoder(tmp_path / "ex3.enc", tmp_path / "ex3_dec.raw")
    assert dec.process()
    dec.saveResult()
    assert sha((tmp_path / "ex3_dec.raw").read_bytes()) == gold["dec_sha256"]
