"""Pins the oracle + host pipeline to the committed golden vectors made by the real reference
(oracle/make_golden.py).  Runs everywhere (no GPU, no /root/reference)."""
import json
import os

import numpy as np
import pytest

import cases
import golden_util


@pytest.mark.parametrize("name", golden_util.CASES)
def test_pipeline_on_oracle_matches_reference_golden(name, tmp_path, oracle_lib):
    problems = golden_util.run_case(name, tmp_path, lambda p: oracle_lib.Oracle(p))
    assert not problems, problems


@pytest.mark.parametrize("name", golden_util.BARCODE_CASES)
@pytest.mark.parametrize("batch_records", [1 << 18, 53])
def test_barcode_pipeline_on_oracle_matches_reference_golden(name, batch_records, tmp_path, oracle_lib):
    problems = golden_util.run_barcode_case(name, tmp_path, lambda p: oracle_lib.Oracle(p), batch_records)
    assert not problems, problems


def test_oracle_operators_match_reference_golden(oracle_lib):
    with open(os.path.join(golden_util.GOLD, "ops_adversarial.json")) as f:
        gold = json.load(f)
    batch = cases.adversarial_batch()
    assert batch.n == gold["n"]
    for pname, rows in gold["sets"].items():
        o = oracle_lib.Oracle(cases.make_params(pname))
        ops = o.ops_pairs(batch)
        got = np.stack([ops[k].astype(np.int64) for k in ("ov_offset", "ov_len", "ov_diff", "poly1", "poly2", "lowq1", "lowq2", "n1", "n2")], axis=1)
        want = np.array(rows, dtype=np.int64)
        bad = np.flatnonzero((got != want).any(axis=1))
        assert len(bad) == 0, (pname, bad[:5].tolist(), got[bad[:3]].tolist(), want[bad[:3]].tolist())
        o.close()


def test_known_answers_from_survey(oracle_lib):
    """KATs of SURVEY.md section 4: the reference's own self-check pairs (util.py:242-246)."""
    from afterqc_b200 import _abi
    from afterqc_b200.batch import PackedBatch
    o = oracle_lib.Oracle(_abi.Params.defaults())
    pairs = [
        ("CAGCGCCTACGGGCCCCTTTTTCTGCGCGACCGCGTGGCTGTGGGCGCGGATGCCTTTGAGCGCGGTGACTTCTCACTGCGTATCGAGCCGCTGGAGGTCTCCC",
         "ACCTCCAGCGGCTCGATACGCAGTGAGAAGTCACCGCGCTCAAAGGCATCCGCGCCCACAGCCACGCGGTCGCGCAGAAAAAGGGGCCCGTAGGCGCGGCTCCC", (-5, 99, 1)),
        ("CAGCGCCTACGGGCCCCTTTTTCTGCGCGACCGCGTGGCTGTGGGCGCGGATGCCTTTGAGCGCGGTGACTTCTCACTGCGTATCGAGC",
         "ACCTCCAGCGGCTCGATACGCAGTGAGAAGTCACCGCGCTCAAAGGCATCCGCGCCCACAGCCACGCGGTCGCGCAGAAAAAGGGGTCC", (10, 79, 1)),
    ]
    b = PackedBatch.from_reads([(p[0], "I" * len(p[0])) for p in pairs], [(p[1], "I" * len(p[1])) for p in pairs])
    ops = o.ops_pairs(b)
    for i, p in enumerate(pairs):
        assert (int(ops[i]["ov_offset"]), int(ops[i]["ov_len"]), int(ops[i]["ov_diff"])) == p[2]


def test_html_report_sections(tmp_path, oracle_lib):
    """QC/<R1>.html: one section per figure of the reference's report (preprocesser.py:700,771-772,785-819)."""
    import re
    import refcmp
    from afterqc_b200 import synth
    for cfg, paired, nfig in (("pe150", True, 3 + 4 * 5), ("se100", False, 1 + 2 * 5)):
        d = str(tmp_path / cfg)
        batch = synth.generate(cfg, 1300)
        refcmp.prepare_case(d, batch, subs=("new",))
        refcmp.run_ours(d, "new", paired, [], lambda p: oracle_lib.Oracle(p))
        html = open(os.path.join(d, "new", "QC", "x_R1.fq.html")).read()
        assert html.count("Plotly.newPlot(") == nfig
        titles = re.findall(r"<li class='menu-item'><a href='#[^']*'>\d+, ([^<]*)</a>", html)
        assert titles[0] == "AfterQC summary" and titles[1] == "Good reads and bad reads after filtering"
        assert ("Read1 kmer strand bias after filtering" in titles) == paired
        assert ("Kmer strand bias after filtering" in titles) == (not paired)
        assert "total reads:" in html and "auto trimming" in html
