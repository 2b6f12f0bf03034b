"""AQC_DEVICE_PARSE=1: the streaming driver with the FASTQ parser on the device (pipeline._filter_stream_device_parse) writes
the same files and the same JSON as the default driver (host parser threads).  Emulated engine; small read blocks so that
records straddle blocks, the mates hold different numbers of records per block and a step is re-parsed with a record cap."""
import filecmp
import os

import pytest

import refcmp
from afterqc_b200 import synth


def _run_both(tmp_path, monkeypatch, batch, extra, mutate=None, block="40000"):
    import emu
    d = str(tmp_path)
    refcmp.prepare_case(d, batch, subs=("host", "dev"))
    if mutate:
        for sub in ("host", "dev"):
            mutate(os.path.join(d, sub))
    paired = batch.paired
    monkeypatch.delenv("AQC_DEVICE_PARSE", raising=False)
    refcmp.run_ours(d, "host", paired, extra, lambda p: emu.EmuEngine(p))
    monkeypatch.setenv("AQC_DEVICE_PARSE", "1")
    monkeypatch.setenv("AQC_DEVICE_PARSE_BLOCK", block)
    calls = []
    from afterqc_b200 import pipeline
    real = pipeline.seqFilter._filter_stream_device_parse

    def spy(self, be, writers, block_bytes=None):
        calls.append(1)
        return real(self, be, writers, block_bytes)
    monkeypatch.setattr(pipeline.seqFilter, "_filter_stream_device_parse", spy)
    refcmp.run_ours(d, "dev", paired, extra, lambda p: emu.EmuEngine(p))
    assert calls, "the device-parse driver did not run"
    a, b = refcmp.load_json(d, "host"), refcmp.load_json(d, "dev")
    diffs = [x for x in refcmp.json_diff(a, b) if not x[1].startswith("/command/")]
    assert not diffs, diffs[:5]
    for f in refcmp.output_files(paired, extra):
        assert filecmp.cmp(os.path.join(d, "host", f), os.path.join(d, "dev", f), shallow=False), f


@pytest.mark.parametrize("extra", [[], ["-f", "2", "-t", "3", "--qc_sample", "700", "--store_overlap", "on"]])
def test_device_parse_driver_equals_host_driver(tmp_path, monkeypatch, oracle_lib, extra):
    _run_both(tmp_path, monkeypatch, synth.generate("pe150", 2600), extra)


def test_device_parse_single_end(tmp_path, monkeypatch, oracle_lib):
    _run_both(tmp_path, monkeypatch, synth.generate("se100", 1800), [])


def test_device_parse_shorter_mate_and_empty_line(tmp_path, monkeypatch, oracle_lib):
    def cut_r2(sub):                        # R2 ends 7 records early: the R1 record read just before that is still counted
        p = os.path.join(sub, "x_R2.fq")
        lines = open(p, "rb").read().split(b"\n")
        open(p, "wb").write(b"\n".join(lines[:4 * (1500 - 7)]) + b"\n")
    _run_both(tmp_path / "a", monkeypatch, synth.generate("pe150", 1500), [], mutate=cut_r2)

    def blank_r1(sub):                      # an empty line in R1 ends the file there (fastq.py:37-49), CRLF line ends elsewhere
        p = os.path.join(sub, "x_R1.fq")
        lines = open(p, "rb").read().split(b"\n")
        lines = [l + b"\r" for l in lines[:4 * 900]] + [b""] + lines[4 * 900:]
        open(p, "wb").write(b"\n".join(lines))
    _run_both(tmp_path / "b", monkeypatch, synth.generate("pe150", 1200), [], mutate=blank_r1)


def test_device_parse_one_block(tmp_path, monkeypatch, oracle_lib):
    _run_both(tmp_path, monkeypatch, synth.generate("pe150", 900), [], block="0")      # the default block: the whole file at once


def test_device_parse_driver_random_files(tmp_path, monkeypatch, oracle_lib):
    """random file pairs (ragged lengths, a shorter mate, an empty or blank line somewhere, CRLF) through both drivers with
    random small read blocks: same outcome -- the same files and JSON, or the same exception type"""
    import random
    import emu
    from afterqc_b200 import pipeline
    rng = random.Random(23)
    for case in range(14):
        n = rng.randint(1200, 2200)
        batch = synth.generate("pe150", n, len_jitter=rng.choice((0, 40, 90)), seed=100 + case)
        kind = rng.choice(("plain", "short_r2", "short_r1", "empty_r2", "blank_r1", "crlf"))

        def mutate(sub, kind=kind, n=n, cut=rng.randint(1, 300), at=rng.randint(1001, 1150), inner=rng.randrange(4)):
            for name, tag in (("x_R1.fq", "r1"), ("x_R2.fq", "r2")):
                p = os.path.join(sub, name)
                lines = open(p, "rb").read().split(b"\n")
                if kind == "short_" + tag:
                    lines = lines[:4 * (n - cut)] + [b""]
                elif kind == "empty_" + tag:
                    lines = lines[:4 * at + inner] + [b""] + lines[4 * at + inner:]
                elif kind == "blank_" + tag:
                    lines = lines[:4 * at + inner] + [b" \t"] + lines[4 * at + inner + 1:]
                elif kind == "crlf":
                    lines = [l + b"\r" if l else l for l in lines]
                open(p, "wb").write(b"\n".join(lines))
        d = str(tmp_path / ("c%d" % case))
        refcmp.prepare_case(d, batch, subs=("host", "dev"))
        for sub in ("host", "dev"):
            mutate(os.path.join(d, sub))
        extra = rng.choice(([], ["--qc_sample", "500"], ["-f", "1", "-t", "2"]))
        outcome = []
        for sub in ("host", "dev"):
            if sub == "dev":
                monkeypatch.setenv("AQC_DEVICE_PARSE", "1")
                monkeypatch.setenv("AQC_DEVICE_PARSE_BLOCK", str(rng.randint(3000, 90000)))
            else:
                monkeypatch.delenv("AQC_DEVICE_PARSE", raising=False)
            try:
                refcmp.run_ours(d, sub, True, extra, lambda p: emu.EmuEngine(p))
                outcome.append("ok")
            except Exception as e:                                   # noqa: BLE001 -- the type is what is compared
                outcome.append(type(e).__name__)
        assert outcome[0] == outcome[1], (case, kind, outcome)
        if outcome[0] != "ok":
            continue
        a, b = refcmp.load_json(d, "host"), refcmp.load_json(d, "dev")
        diffs = [x for x in refcmp.json_diff(a, b) if not x[1].startswith("/command/")]
        assert not diffs, (case, kind, diffs[:5])
        for f in refcmp.output_files(True, extra):
            assert filecmp.cmp(os.path.join(d, "host", f), os.path.join(d, "dev", f), shallow=False), (case, kind, f)


def test_device_parse_gz_inputs(tmp_path, monkeypatch, oracle_lib):
    """.gz inputs: the native reader's decoder threads feed text to the device parser (fastq_io.TextSource); outputs are .gz by
    extension (compared after decompression, quirk Q14)"""
    import gzip
    import json
    import emu
    from afterqc_b200 import cli
    from afterqc_b200.pipeline import seqFilter
    batch = synth.generate("pe150", 2400, len_jitter=30)
    outs = {}
    for sub in ("host", "dev"):
        d = str(tmp_path / sub)
        os.makedirs(d)
        synth.write_fastq(batch, os.path.join(d, "p_R1.fq"), os.path.join(d, "p_R2.fq"))
        for m in ("1", "2"):
            with open(os.path.join(d, "p_R%s.fq" % m), "rb") as f, gzip.open(os.path.join(d, "x_R%s.fq.gz" % m), "wb", compresslevel=1) as g:
                g.write(f.read())
        if sub == "dev":
            monkeypatch.setenv("AQC_DEVICE_PARSE", "1")
            monkeypatch.setenv("AQC_DEVICE_PARSE_BLOCK", "70000")
        else:
            monkeypatch.delenv("AQC_DEVICE_PARSE", raising=False)
        opts, _ = cli.parseCommand(["-1", os.path.join(d, "x_R1.fq.gz"), "-2", os.path.join(d, "x_R2.fq.gz"), "-g", os.path.join(d, "good")])
        cli.normalize_options(opts); opts.barcode = False
        seqFilter(opts, backend_factory=lambda p: emu.EmuEngine(p)).run()
        outs[sub] = d
    a = json.load(open(os.path.join(outs["host"], "QC", "x_R1.fq.gz.json")))
    b = json.load(open(os.path.join(outs["dev"], "QC", "x_R1.fq.gz.json")))
    diffs = [x for x in refcmp.json_diff(a, b) if not x[1].startswith("/command/")]
    assert not diffs, diffs[:5]
    for f in ("good/x_R1.good.fq.gz", "good/x_R2.good.fq.gz", "bad/x_R1.bad.fq.gz", "bad/x_R2.bad.fq.gz"):
        assert gzip.open(os.path.join(outs["host"], f)).read() == gzip.open(os.path.join(outs["dev"], f)).read(), f
