"""The drop-in CLI contract pinned against the reference's OWN binaries (oracle/_ref/ref_cli_encoder / ref_cli_decoder = its
main.cpp built as its makefile builds it): for a battery of settings files and quantisation-matrix files, `bin/encoder` and
`bin/decoder` must take the same decision as the reference -- the same exit code 1..5 (main.cpp:19-185: usage, unreadable
settings, incomplete settings, unreadable matrix, non-numeric setting), or "go on to process".  Past that point the reference
either succeeds (and segfaults at exit), aborts on nonsense dimensions or exit(-1)s on an unreadable input; the
product needs a GPU, so on this box it stops with its no-device code.  Only the decision is compared."""
import shutil
import subprocess
import tempfile
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
REF = ROOT / "oracle" / "_ref"
pytestmark = pytest.mark.skipif(not (REF / "ref_cli_encoder").exists() or not (ROOT / "bin" / "encoder").exists(),
                                reason="needs the compiled reference CLIs (oracle/build_ref.sh) and the built product CLIs")

BASE = dict(rawfile="a.raw", encfile="a.enc", decfile="a_dec.raw", width="8", height="8", rle="1", quantfile="q.txt", logfile="a.txt")
Q4 = "1 2 3 4\n1 2 3 4\n1 2 3 4\n1 2 3 4\n"


def conf(d, sep="=", extra=""):
    return "".join(f"{k}{sep}{v}\n" for k, v in d.items()) + extra


def drop(d, k):
    return {a: b for a, b in d.items() if a != k}


def encoder_cases():
    c = [("noargs", None, Q4, []), ("twoargs", conf(BASE), Q4, ["a.conf", "b.conf"]), ("missingconf", None, Q4, ["nope.conf"]),
         ("onlytwo", "rawfile=a.raw\nencfile=a.enc\n", Q4, None)]
    c += [(f"missing_{k}", conf(drop(BASE, k)), Q4, None) for k in BASE]
    c += [("same_raw_enc", conf(dict(BASE, encfile="a.raw")), Q4, None), ("noquantfile", conf(BASE), None, None)]
    for name, kv in (("width_word", dict(width="eight")), ("height_word", dict(height="x8")), ("width_trailing", dict(width="8px")),
                     ("width_hex", dict(width="0x8")), ("width_plus", dict(width="+8")), ("width_space", dict(width=" 8")),
                     ("rle_word", dict(rle="yes")), ("rle_2", dict(rle="2")), ("width_neg", dict(width="-8")),
                     ("width_big", dict(width="70000")), ("width_empty", dict(width=""))):
        c.append((name, conf(dict(BASE, **kv)), Q4, None))
    for name, q in (("q_short", "1 2 3 4\n" * 3), ("q_long_row", "1 2 3 4 5\n" + "1 2 3 4\n" * 3), ("q_word", "1 2 3 x\n" + "1 2 3 4\n" * 3),
                    ("q_neg", "1 2 3 -4\n" + "1 2 3 4\n" * 3), ("q_big", "1 2 3 70000\n" + "1 2 3 4\n" * 3), ("q_empty", ""),
                    ("q_extra_rows", Q4 + "1 2 3 4\n"), ("q_trailing_blank", Q4 + "\n"), ("q_tabs", Q4.replace(" ", "\t")),
                    ("q_multispace", Q4.replace(" ", "   ")), ("q_lead_trail_space", Q4.replace("\n", "  \n").replace("1 2", "  1 2")),
                    ("q_blank_lines", "1 2 3 4\n\n" + "1 2 3 4\n" * 3), ("q_crlf", Q4.replace("\n", "\r\n")), ("q_hex", Q4.replace("4", "0x4")),
                    ("q_trailing_junk", Q4.replace("4", "4;")), ("q_no_final_newline", Q4.rstrip("\n")), ("q_float", Q4.replace("2", "2.5"))):
        c.append((name, conf(BASE), q, None))
    c += [("conf_spaces", conf(BASE, sep=" = "), Q4, None), ("conf_comment", conf(BASE, extra="# comment\n"), Q4, None),
          ("conf_unknown_key", conf(BASE, extra="foo=bar\n"), Q4, None), ("conf_crlf", conf(BASE).replace("\n", "\r\n"), Q4, None),
          ("conf_dup_key", conf(BASE, extra="width=16\n"), Q4, None), ("conf_blank_lines", conf(BASE).replace("\n", "\n\n"), Q4, None),
          ("conf_noeq_line", conf(BASE, extra="garbage line\n"), Q4, None), ("conf_empty_key", conf(BASE, extra="=5\n"), Q4, None),
          ("conf_two_eq", conf(dict(BASE, logfile="a=b.txt")), Q4, None),
          ("video_partial_gop", conf(dict(BASE, gop="4")), Q4, None), ("video_partial_mer", conf(dict(BASE, merange="16")), Q4, None),
          ("video_word", conf(dict(BASE, gop="four", merange="16", motioncompensation="1")), Q4, None),
          ("video_mer_word", conf(dict(BASE, gop="4", merange="x", motioncompensation="1")), Q4, None),
          ("input_missing", conf(BASE), Q4, None)]
    return c


def decoder_cases():
    c = [("noargs", None, Q4, []), ("missingconf", None, Q4, ["nope.conf"])]
    c += [(f"missing_{k}", conf(drop(BASE, k)), Q4, None) for k in BASE]
    c += [("same_enc_dec", conf(dict(BASE, decfile="a.enc")), Q4, None),
          ("video_mc_word", conf(dict(BASE, gop="4", merange="16", motioncompensation="yes")), Q4, None),
          ("video_partial", conf(dict(BASE, motioncompensation="1")), Q4, None), ("input_missing", conf(BASE), Q4, None)]
    return c


def decision(rc):
    return rc if rc in (1, 2, 3, 4, 5, 255) else "process"       # 255 = exit(-1): unreadable input file (ImageBase.cpp:24-27)


def run_battery(ref_exe, our_exe, cases):
    diffs = []
    for name, ctext, q, args in cases:
        got = {}
        for which, exe in (("ref", ref_exe), ("ours", our_exe)):
            d = tempfile.mkdtemp()
            try:
                if ctext is not None:
                    Path(d, "a.conf").write_bytes(ctext.encode())
                if q is not None:
                    Path(d, "q.txt").write_bytes(q.encode())
                if name != "input_missing":
                    Path(d, "a.raw").write_bytes(b"\x80" * 64)
                    Path(d, "a.enc").write_bytes(b"\x00" * 8)    # a stream the decoders can at least open
                r = subprocess.run([str(exe), *(["a.conf"] if args is None else args)], cwd=d, capture_output=True, timeout=120)
                got[which] = decision(r.returncode)
            finally:
                shutil.rmtree(d, ignore_errors=True)
        if got["ref"] != got["ours"]:
            diffs.append((name, got))
    return diffs


def test_encoder_cli_decisions_match_the_reference_binary():
    diffs = run_battery(REF / "ref_cli_encoder", ROOT / "bin" / "encoder", encoder_cases())
    assert not diffs, diffs


def test_decoder_cli_decisions_match_the_reference_binary():
    diffs = run_battery(REF / "ref_cli_decoder", ROOT / "bin" / "decoder", decoder_cases())
    assert not diffs, diffs


def test_python_read_matrix_agrees_with_the_cli(tmp_path):
    """imageencoder_b200.read_matrix (the Python mirror of MatrixReader) accepts / rejects the same matrix files as bin/encoder --
    which the tests above tie to the reference's own binary -- and reads the same values"""
    import numpy as np

    import imageencoder_b200 as ie
    enc = ROOT / "bin" / "encoder"
    (tmp_path / "a.conf").write_text(conf(BASE))
    (tmp_path / "a.raw").write_bytes(b"\x80" * 64)
    for name, _c, q, _a in encoder_cases():
        if not name.startswith("q_") or q is None:
            continue
        (tmp_path / "q.txt").write_bytes(q.encode())
        r = subprocess.run([str(enc), "a.conf"], cwd=tmp_path, capture_output=True, text=True, timeout=120)
        cli_ok = r.returncode != 4
        try:
            m = ie.read_matrix(tmp_path / "q.txt")
            py_ok = True
        except ValueError:
            py_ok = False
        assert py_ok == cli_ok, (name, r.returncode)
        if cli_ok:
            # the CLI prints the matrix it read with setw(4) per entry (codec.cpp: MatrixReader::toString)
            printed = r.stdout.split("Quantization matrix:")[1].split("-------------------------")[1].strip("\n").split("\n")[: m.shape[0]]
            assert printed == ["".join(f"{int(v):4d}" for v in row) for row in np.asarray(m)], name
