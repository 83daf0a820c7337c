"""The drop-in boundary: the reference's class and function names over the C ABI.
  * the reference's OWN generate_fragment_groups body (commonFunctions.cpp:41-80, compiled unchanged by oracle/Makefile
    `refbody` where /root/reference is present) against FragmentsDatabase::begin()/end() and SequenceOcupationList of
    this repo must reproduce the reference's golden bytes;
  * sort_groups is a pure function of (list, diag_func), like the reference's;
  * the bucket view visits the records the way the reference's bucket array does;
  * short CSV rows (readFragment's padding rule) through the CLI."""
import hashlib
import os
import subprocess

import numpy as np
import pytest

from oracle import oracle as O
from repkiller_b200 import frags, gen
from repkiller_b200.frags import FRAG_DTYPE, make_header

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "repkiller_b200", "bin", "repkiller")
HOSTCHECK = os.path.join(ROOT, "repkiller_b200", "bin", "rk_hostcheck")
REFBODY = os.path.join(ROOT, "oracle", "_ref", "ref_body_check")


def _need(path):
    if not os.path.exists(path):
        pytest.skip(f"{os.path.relpath(path, ROOT)} is built where the reference sources are present (python __graft_entry__.py)")


def test_reference_loop_body_on_the_facades_fuzz(tmp_path, fuzz_cases):
    _need(REFBODY)
    for c in fuzz_cases[::4]:
        inp = tmp_path / "in.csv"
        inp.write_text(c["csv"], newline="")
        outp = tmp_path / "out.csv"
        p = subprocess.run([REFBODY, str(inp), str(outp), repr(c["len_ratio"]), repr(c["pos_ratio"])], capture_output=True)
        assert p.returncode == 0, p.stderr
        assert outp.read_bytes() == c["ref_out"].encode("latin1"), f"fuzz seed {c['seed']}"


@pytest.mark.parametrize("name", ["dense", "c1"])
def test_reference_loop_body_on_the_facades_golden_md5(tmp_path, medium_cases, name):
    """100k fragments: ~200k single device queries, groups of thousands of members, std::sort tie order"""
    _need(REFBODY)
    c = medium_cases[name]
    w = gen.Workload(**c["workload"])
    inp = tmp_path / "in.csv"
    O.write_input_csv(str(inp), gen.generate(w), w.lx, w.ly)
    outp = tmp_path / "out.csv"
    p = subprocess.run([REFBODY, str(inp), str(outp), repr(w.len_ratio), repr(w.pos_ratio)], capture_output=True, timeout=900)
    assert p.returncode == 0, p.stderr
    assert hashlib.md5(outp.read_bytes()).hexdigest() == c["ref_md5"]


def test_sort_groups_is_a_pure_function(tmp_path, medium_cases):
    """a grouping with other ratios between generate_fragment_groups and sort_groups must not change the result"""
    c = medium_cases["c1"]
    w = gen.Workload(**c["workload"])
    inp = tmp_path / "c1.csv"
    O.write_input_csv(str(inp), gen.generate(w), w.lx, w.ly)
    outp = tmp_path / "out.csv"
    p = subprocess.run([HOSTCHECK, "steps_pure", str(inp), str(outp), "0.05", "0.05", "0.5", "0.5"], capture_output=True)
    assert p.returncode == 0, p.stderr
    assert hashlib.md5(outp.read_bytes()).hexdigest() == c["ref_md5"]


def test_sort_groups_with_any_table_equals_std_sort(tmp_path, medium_cases):
    w = gen.Workload(**medium_cases["dense"]["workload"])
    inp = tmp_path / "in.csv"
    O.write_input_csv(str(inp), gen.generate(w), w.lx, w.ly)
    p = subprocess.run([HOSTCHECK, "sort_any", str(inp), repr(w.len_ratio), repr(w.pos_ratio)], capture_output=True, text=True)
    assert p.returncode == 0, p.stdout + p.stderr
    n_big = int(p.stdout.split(" with more than 16 members")[0].split(",")[-1])
    assert n_big > 10 and " 0 differ" in p.stdout, p.stdout


def test_bucket_view_is_the_processing_order(tmp_path):
    w = gen.scaled(gen.WORKLOADS["c1"], 50_000)
    rec = gen.generate(w)
    rec["xStart"][:7] = (w.lx + 1) // 10 * 10   # some fragments of the never-visited last bucket (xStart/10 == vsize-1)
    inp = tmp_path / "in.csv"
    O.write_input_csv(str(inp), rec, w.lx, w.ly)
    outp = tmp_path / "visited.bin"
    p = subprocess.run([HOSTCHECK, "buckets", str(inp), str(outp)], capture_output=True)
    assert p.returncode == 0, p.stderr
    loaded, lx1, ly1, _ = O.load_csv(str(inp))
    g = O.group(loaded, lx1, ly1, 0.05, 0.05)
    assert g.n_kept < loaded.shape[0]
    got = np.fromfile(outp, dtype=FRAG_DTYPE)
    assert got.tobytes() == loaded[g.rank_fidx].tobytes()


def test_cli_short_rows_do_not_overflow_the_record_buffer(tmp_path):
    """readFragment accepts `Frag,5` (7 bytes, missing fields repeat the last one): far more records than bytes/28"""
    rows = "".join(f"Frag,{5 + (i % 90)}\n" for i in range(20000))
    text = make_header(2000, 2000, 20000) + rows
    inp = tmp_path / "in.csv"
    inp.write_text(text, newline="")
    outp = tmp_path / "out.csv"
    p = subprocess.run([CLI, str(inp), str(outp), "0.5", "0.5"], capture_output=True)
    assert p.returncode == 0, p.stderr
    rec, lx1, ly1, hdr = O.load_csv(str(inp))
    assert rec.shape[0] == 20000
    g = O.group(rec, lx1, ly1, 0.5, 0.5)
    want = tmp_path / "want.csv"
    O.write_output(str(want), hdr, rec, g)
    assert outp.read_bytes() == want.read_bytes()


def test_cli_values_beyond_32_bits_take_the_record_ingest(tmp_path):
    """the CLI packs records into the compact ingest; a 40-bit score (printed, never used for grouping) must switch it to
    the 109-byte ingest and come out digit for digit"""
    w = gen.scaled(gen.WORKLOADS["c1"], 5_000)
    rec = gen.generate(w)
    rec["score"][::7] = (1 << 40) + np.arange(rec[::7].shape[0], dtype=np.uint64)
    inp = tmp_path / "in.csv"
    O.write_input_csv(str(inp), rec, w.lx, w.ly)
    outp = tmp_path / "out.csv"
    p = subprocess.run([CLI, str(inp), str(outp), "0.05", "0.05"], capture_output=True)
    assert p.returncode == 0, p.stderr
    loaded, lx1, ly1, hdr = O.load_csv(str(inp))
    g = O.group(loaded, lx1, ly1, 0.05, 0.05)
    want = tmp_path / "want.csv"
    O.write_output(str(want), hdr, loaded, g)
    assert outp.read_bytes() == want.read_bytes()
    assert b"1099511627776" in outp.read_bytes()


def test_cli_over_several_ranks(tmp_path, medium_cases):
    """RK_DEVICES: the C++ CLI partitions the comparison over several contexts (rk_create_multi; the same device listed
    three times makes the ranks exchange through device copies) and writes the reference's bytes"""
    c = medium_cases["c1"]
    w = gen.Workload(**c["workload"])
    inp = tmp_path / "c1.csv"
    O.write_input_csv(str(inp), gen.generate(w), w.lx, w.ly)
    outp = tmp_path / "out.csv"
    p = subprocess.run([CLI, str(inp), str(outp), "0.05", "0.05"], capture_output=True, env=dict(os.environ, RK_DEVICES="0,0,0"))
    assert p.returncode == 0, p.stderr
    assert hashlib.md5(outp.read_bytes()).hexdigest() == c["ref_md5"]


def test_cli_reads_geckos_binary_container(tmp_path, medium_cases):
    """SURVEY §8f N4 (parity unpinned: the reference reads CSV only).  A .frags file — 16-byte header, 109-byte big-endian
    records (csrc/host/GeckoFrags.h) — loads as its CSV rendering does, so the CLI writes the lines the reference wrote for
    that CSV (md5 of the golden set, header excluded: a binary file carries no header text), on one GPU and over three ranks."""
    c = medium_cases["c1"]
    w = gen.Workload(**c["workload"])
    rec = gen.generate(w)
    csv_in, bin_in = tmp_path / "c1.csv", tmp_path / "c1.frags"
    O.write_input_csv(str(csv_in), rec, w.lx, w.ly)
    frags.write_gecko_binary(str(bin_in), rec, w.lx, w.ly)
    out_csv = tmp_path / "from_csv.out"
    p = subprocess.run([CLI, str(csv_in), str(out_csv), "0.05", "0.05"], capture_output=True)
    assert p.returncode == 0, p.stderr
    assert hashlib.md5(out_csv.read_bytes()).hexdigest() == c["ref_md5"]
    body = out_csv.read_bytes().split(b"\n", 16)[16]
    for env in ({}, {"RK_DEVICES": "0,0,0"}):
        out_bin = tmp_path / "from_bin.out"
        p = subprocess.run([CLI, str(bin_in), str(out_bin), "0.05", "0.05"], capture_output=True, env=dict(os.environ, **env))
        assert p.returncode == 0, p.stderr
        lines = out_bin.read_bytes().split(b"\n", 16)
        assert lines[16] == body, env
        assert lines[6] == b"SeqX length : %d" % w.lx and lines[7] == b"SeqY length : %d" % w.ly
        assert lines[12] == b"Total fragments : %d" % w.n
