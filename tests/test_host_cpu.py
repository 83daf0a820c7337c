"""Host-side logic of the drop-in that needs no GPU: the CSV row parser (readFragment rules) and the output
writer's line format, through the rk_hostcheck helper, against the oracle and the reference's golden bytes."""
import os
import subprocess

import numpy as np

from oracle import oracle as O
from repkiller_b200.frags import FRAG_DTYPE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOSTCHECK = os.path.join(ROOT, "repkiller_b200", "bin", "rk_hostcheck")
CLI = os.path.join(ROOT, "repkiller_b200", "bin", "repkiller")


def test_csv_parser_matches_oracle_on_fuzz(tmp_path, fuzz_cases):
    for c in fuzz_cases:
        inp = tmp_path / "in.csv"
        inp.write_text(c["csv"], newline="")
        out = tmp_path / "recs.bin"
        subprocess.check_call([HOSTCHECK, "parse", str(inp), str(out)])
        got = np.fromfile(out, dtype=FRAG_DTYPE)
        want, _, _, _ = O.load_csv(str(inp))
        assert got.tobytes() == want.tobytes(), f"fuzz seed {c['seed']}"


def test_parallel_parser_matches_oracle_for_any_thread_count(tmp_path, fuzz_cases):
    """FragmentsDatabase parses the rows with all host cores (ranges cut at row boundaries, concatenated in order): the
    records must be those of the row-by-row loop for every thread count, also when there are more threads than rows."""
    for c in fuzz_cases[::3]:
        inp = tmp_path / "in.csv"
        inp.write_text(c["csv"], newline="")
        want, _, _, _ = O.load_csv(str(inp))
        for threads in (1, 2, 3, 7, 16, 64):
            out = tmp_path / "recs.bin"
            subprocess.check_call([HOSTCHECK, "parse", str(inp), str(out), str(threads)])
            got = np.fromfile(out, dtype=FRAG_DTYPE)
            assert got.tobytes() == want.tobytes(), f"fuzz seed {c['seed']}, {threads} threads"


def test_float_fast_path_equals_strtof():
    """readFragment parses plain `digits[.digits]` similarity tokens without strtof; the result must be strtof's
    (correctly rounded float) on every token the fast path accepts."""
    p = subprocess.run([HOSTCHECK, "floatcheck", "3000000", "7"], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout[-1500:]
    assert " 0 mismatches" in p.stdout


def test_writer_line_format_matches_reference(tmp_path, fuzz_cases):
    """Every line the reference wrote for a singleton group (repval 0) must be reproduced byte for byte by the
    host writer from the same record (float formatting of similarity and identity, '-nan', strand bytes)."""
    checked = 0
    for c in fuzz_cases[:60]:
        inp = tmp_path / "in.csv"
        inp.write_text(c["csv"], newline="")
        out = tmp_path / "w.csv"
        subprocess.check_call([HOSTCHECK, "write", str(inp), str(out)])
        mine = out.read_bytes().split(b"\n")
        ref = c["ref_out"].encode("latin1").split(b"\n")
        assert mine[:16] == ref[:16]                      # header echoed verbatim
        def key(line):                                    # everything but the gid and repval columns
            f = line.split(b",")
            return tuple(f[:6] + f[7:13])
        mine_keys = {key(l) for l in mine[16:] if l}
        for l in ref[16:]:
            if l:
                assert key(l) in mine_keys, (c["seed"], l)
                checked += 1
    assert checked > 1000


def test_cli_argument_errors():
    p = subprocess.run([CLI], capture_output=True)
    assert p.returncode == 1
    assert b"Invalid number of arguments." in p.stderr
    assert p.stdout.startswith(b"Repkiller v0.9.b\nUsage: ./repkiller <input_file_path> <output_file_path> <length_ratio> <position_ratio>\n")
    p = subprocess.run([CLI, "/nonexistent.csv", "/tmp/o.csv", "0.05", "0.05"], capture_output=True)
    assert p.returncode != 0 and b"Could not open input file /nonexistent.csv." in p.stderr
