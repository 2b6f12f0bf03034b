"""Hand-made inputs for individual quirks of the reference (SURVEY.md section 8, Q-list), run through the UNMODIFIED
reference and through this repo's host pipeline on the oracle.  Build container only."""
import os
import random
import re

import pytest

import refcmp
from afterqc_b200 import cli
from afterqc_b200.pipeline import seqFilter

pytestmark = pytest.mark.reference


def _rand(rng, n, alphabet="ACGT"):
    return "".join(rng.choice(alphabet) for _ in range(n))


def _write_pair(d, subs, recs1, recs2):
    for sub in subs:
        os.makedirs(os.path.join(d, sub), exist_ok=True)
        for name, recs in (("x_R1.fq", recs1), ("x_R2.fq", recs2)):
            if recs is None:
                continue
            with open(os.path.join(d, sub, name), "w") as f:
                for i, (s, q) in enumerate(recs):
                    f.write("@r%d\n%s\n+\n%s\n" % (i, s, q))


def _run_both(d, paired, extra, oracle_lib, mutate=None):
    from oracle import ref_loader
    mods = ref_loader.load()
    # reference: build the options exactly like after.main, optionally tweak them, then run seqFilter
    import sys
    argv = sys.argv
    try:
        sys.argv = ["after.py"] + refcmp.cli_args(d, "ref", paired, extra)
        ropt, _ = mods["after"].parseCommand()
    finally:
        sys.argv = argv
    ropt.version = "0.9.6"
    for k in ("trim_pair_same", "draw", "store_overlap"):
        setattr(ropt, k, mods["after"].parseBool(getattr(ropt, k)))
    ropt.trim_front2, ropt.trim_tail2 = ropt.trim_front, ropt.trim_tail
    ropt.barcode = False
    if mutate:
        mutate(ropt)
    mods["preprocesser"].seqFilter(ropt).run()
    opts, _ = cli.parseCommand(refcmp.cli_args(d, "new", paired, extra))
    cli.normalize_options(opts); opts.barcode = False
    if mutate:
        mutate(opts)
    sf = seqFilter(opts, backend_factory=lambda p: oracle_lib.Oracle(p))
    sf.run()
    a, b = refcmp.load_json(d, "ref"), refcmp.load_json(d, "new")
    diffs = [x for x in refcmp.json_diff(a, b) if not x[1].startswith("/command/")]
    assert not diffs, diffs[:5]
    for f in refcmp.output_files(paired, extra):
        assert open(os.path.join(d, "ref", f), "rb").read() == open(os.path.join(d, "new", f), "rb").read(), f
    return a, sf


def test_q2_q3_r2_ignored_by_lowquality_and_length_filters(tmp_path, oracle_lib):
    rng = random.Random(1)
    recs1, recs2 = [], []
    for i in range(60):
        recs1.append((_rand(rng, 100), "I" * 100))
        if i % 3 == 0:
            recs2.append((_rand(rng, 100), "#" * 100))          # all low quality: still good (only lowQual1 is tested)
        elif i % 3 == 1:
            recs2.append((_rand(rng, 12), "I" * 12))            # far below -s 35: still good (R2 length never checked)
        else:
            recs2.append((_rand(rng, 100), "I" * 100))
    d = str(tmp_path)
    _write_pair(d, ("ref", "new"), recs1, recs2)
    a, _ = _run_both(d, True, ["-f", "0", "-t", "0"], oracle_lib)
    assert a["afterqc_main_summary"]["good_reads"] == 60


def test_q4_trim_gate_is_keyed_on_r1_only(tmp_path, oracle_lib):
    rng = random.Random(2)
    recs1 = [(_rand(rng, 80), "I" * 80) for _ in range(40)]
    recs2 = [(_rand(rng, 80), "I" * 80) for _ in range(40)]
    d = str(tmp_path)
    _write_pair(d, ("ref", "new"), recs1, recs2)

    def r2_only_trim(o):            # R1 trim 0/0 closes the gate: the R2 values must be ignored (preprocesser.py:455)
        o.trim_front2, o.trim_tail2 = 7, 9
    a, _ = _run_both(d, True, ["-f", "0", "-t", "0"], oracle_lib, mutate=r2_only_trim)
    assert a["afterqc_main_summary"]["good_bases"] == 40 * 80
    out = open(os.path.join(d, "new", "good", "x_R2.good.fq")).read().split("\n")
    assert all(len(s) == 80 for s in out[1::4] if s)


def test_q7_overlap_histogram_is_taken_before_the_adapter_trim(tmp_path, oracle_lib):
    """overlap_histgram[overlap_len] is bumped with the FIRST scan (preprocesser.py:517), distance_histgram after the
    rescan (:536); the overlap histogram is only visible in the reference's HTML."""
    rng = random.Random(3)
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    recs1, recs2 = [], []
    for i in range(80):
        frag = _rand(rng, rng.choice([60, 90, 120, 200, 400]))        # < 150: adapter read-through (negative offset)
        rc = "".join(comp[c] for c in reversed(frag))
        recs1.append(((frag + _rand(rng, 150))[:150], "I" * 150))
        recs2.append(((rc + _rand(rng, 150))[:150], "I" * 150))
    d = str(tmp_path)
    _write_pair(d, ("ref", "new"), recs1, recs2)
    a, sf = _run_both(d, True, ["-f", "0", "-t", "0"], oracle_lib)
    assert a["afterqc_overlap"]["trimmed_adapter_reads"] > 10
    html = open(os.path.join(d, "ref", "QC", "x_R1.fq.html")).read()
    m = re.search(r"Plotly\.newPlot\('overlap_stat'", html)
    block = html[:m.start()].rsplit("var data=[", 1)[1]
    ys = [int(v) for v in re.search(r"y:\[([^\]]*)\]", block).group(1).split(",")]
    assert ys == sf.overlap_histgram


def test_q13_empty_line_ends_the_file(tmp_path, oracle_lib):
    rng = random.Random(4)
    d = str(tmp_path)
    for sub in ("ref", "new"):
        os.makedirs(os.path.join(d, sub))
        rng = random.Random(4)
        with open(os.path.join(d, sub, "x_R1.fq"), "w") as f:
            for i in range(50):
                s = _rand(rng, 70)
                f.write("@r%d\n%s\n+\n%s\n" % (i, s, "I" * 70))
                if i == 30:
                    f.write("\n")                                   # everything after this line is ignored
    a, _ = _run_both(d, False, ["-f", "0", "-t", "0"], oracle_lib)
    assert a["afterqc_main_summary"]["total_reads"] == 31
