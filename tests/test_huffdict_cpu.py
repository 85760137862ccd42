"""CPU check of the host-side Huffman dictionary build (imageencoder_b200/csrc/huffman.cu: build_dictionary).

The reference's tie-breaking is libstdc++ container behaviour (unordered_map iteration order, priority_queue heap order, unstable
std::sort; Huffman.cpp:237-311).  The product keeps those containers and only pools their storage;
tests/host/huffdict_check.cu runs it against the plain-allocator transcription it replaced on random histograms: same codes,
same dictionary header bits.  (End to end the Huffman stage is compared with the oracle / the compiled reference by the GPU
tests.)"""
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def test_pooled_dictionary_build_matches_plain_allocator_transcription(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    import imageencoder_b200 as ie
    ie.lib()                                                   # the library must be built (harness links against it)
    libdir = ROOT / "imageencoder_b200"
    exe = tmp_path / "huffdict_check"
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-O2", "-fmad=false", "--expt-relaxed-constexpr",
           "-diag-suppress", "1650", "-o", str(exe), str(ROOT / "tests" / "host" / "huffdict_check.cu"), f"-L{libdir}",
           "-limageencoder_b200", "-Xlinker", f"-rpath={libdir}"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([str(exe), "20000"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "huffdict_check: ok" in r.stdout
