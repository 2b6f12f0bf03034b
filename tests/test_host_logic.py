"""Host-side logic that is not per-read: FASTQ parsing quirks, trimming slices, layout, CLI table."""
import gzip
import os
import re

import numpy as np

from afterqc_b200 import _abi, cli, fastq_io
from afterqc_b200.batch import PackedBatch
from afterqc_b200.pipeline import getMainName

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write(path, data):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "wb") as f:
        f.write(data)


def test_fastq_parse_basic_and_gz(tmp_path):
    data = b"@r1 x\nACGT\n+\nIIII\n@r2\nGG\n+r2\n#I\n"
    for name in ("a.fq", "a.fq.gz"):
        p = str(tmp_path / name)
        _write(p, data)
        rec = fastq_io.read_all(p)
        assert rec.n == 2
        assert rec.names.get(0) == b"@r1 x" and rec.seqs.get(1) == b"GG" and rec.plus.get(1) == b"+r2" and rec.quals.get(1) == b"#I"


def test_fastq_empty_line_is_eof_quirk_q13(tmp_path):
    # an empty (after rstrip) line ends the file, the record it belongs to is dropped (fastq.py:44-47)
    p = str(tmp_path / "b.fq")
    _write(p, b"@a\nAC\n+\nII\n@b\n \t\n+\nII\n@c\nAC\n+\nII\n")
    assert fastq_io.read_all(p).n == 1
    _write(p, b"@a\nAC\n+\nII\n\n@c\nAC\n+\nII\n")
    assert fastq_io.read_all(p).n == 1


def test_fastq_crlf_and_missing_final_newline_and_partial_record(tmp_path):
    p = str(tmp_path / "c.fq")
    _write(p, b"@a\r\nACG\r\n+\r\nIII\r\n@b\nTT\n+\nII")
    rec = fastq_io.read_all(p)
    assert rec.n == 2 and rec.seqs.get(0) == b"ACG" and rec.quals.get(1) == b"II"
    _write(p, b"@a\nACG\n+\nIII\n@b\nTT\n+\n")     # partial last record is dropped
    assert fastq_io.read_all(p).n == 1


def test_fastq_streaming_blocks_agree(tmp_path):
    p = str(tmp_path / "d.fq")
    rng = np.random.default_rng(1)
    recs = []
    for i in range(500):
        L = int(rng.integers(1, 90))
        s = bytes(rng.choice(list(b"ACGTN"), L).tolist())
        recs.append(b"@n%d\n%s\n+\n%s\n" % (i, s, b"I" * L))
    _write(p, b"".join(recs))
    whole = fastq_io.read_all(p)
    parts = list(fastq_io.iter_records(p, block_bytes=997))
    assert sum(c.n for c in parts) == whole.n == 500
    assert b"".join(c.seqs.data.tobytes() for c in parts) == whole.seqs.data.tobytes()


def test_quality_length_mismatch_is_rejected(tmp_path):
    p = str(tmp_path / "e.fq")
    _write(p, b"@a\nACGT\n+\nIII\n")
    try:
        fastq_io.read_all(p)
    except ValueError:
        return
    raise AssertionError("expected ValueError")


def test_main_name_and_flags():
    assert getMainName("/x/y/S1_R1_001.fastq.gz") == "S1_R1_001"
    assert getMainName("a.fq") == "a"
    assert cli.matchFlag("S_R1_001.fq", "R1") and not cli.matchFlag("SR1x.fq", "R1") and cli.matchFlag("xR1_.fq", "R1_")
    opts, _ = cli.parseCommand(["-1", "a.fq"])
    cli.normalize_options(opts)
    assert (opts.trim_front, opts.trim_tail, opts.qualified_quality_phred, opts.unqualified_base_limit, opts.poly_size_limit,
            opts.allow_mismatch_in_poly, opts.n_base_limit, opts.seq_len_req, opts.qc_sample, opts.qc_kmer, opts.compression) == \
        (-1, -1, 15, 60, 35, 2, 5, 35, 200000, 8, 2)
    assert opts.trim_pair_same is True and opts.store_overlap is False and opts.trim_front2 == -1


def test_packed_batch_roundtrip_and_slice():
    r1 = [("ACGT", "IIII"), ("", ""), ("NNACG", "#####")]
    r2 = [("TT", "II"), ("G", "I"), ("ACGTACGT", "IIIIIIII")]
    b = PackedBatch.from_reads(r1, r2, first_index=10)
    assert b.n == 3 and b.read(1, 2) == r1[2] and b.read(2, 0) == r2[0] and b.max_len() == 8
    s = b.slice(1, 3)
    assert s.n == 2 and s.first_index == 11 and s.read(2, 1) == r2[2] and int(s.off1[0]) == 0


def test_python_abi_mirror_matches_header():
    hdr = open(os.path.join(ROOT, "include", "afterqc_b200.h")).read()
    assert int(re.search(r"#define AQC_MAX_LEN (\d+)", hdr).group(1)) == _abi.MAX_LEN
    assert int(re.search(r"#define AQC_ABI_VERSION (\d+)", hdr).group(1)) == _abi.ABI_VERSION
    assert _abi.RESULT_DTYPE.itemsize == 32 and _abi.OPS_DTYPE.itemsize == 32
    import ctypes
    assert ctypes.sizeof(_abi.Params) == 24 * 4 and ctypes.sizeof(_abi.Batch) == 16 + 6 * 8
    names = re.findall(r"AQC_C_([A-Z0-9_]+)", hdr.split("enum {\n    AQC_C_TOTAL_READS")[1].split("AQC_C_ERR_MATRIX")[0])
    assert ["TOTAL_READS"] + names == list(_abi.CIDX)


def test_parallel_gzip_writer_roundtrip(tmp_path):
    import random
    p = str(tmp_path / "o.fq.gz")
    w = fastq_io.Writer(p, False, 2)
    w._f.BLOCK = 1000                       # force several members
    rng = random.Random(1)
    chunks = [bytes(rng.choice(b"ACGT\n@+I") for _ in range(rng.randint(0, 700))) for _ in range(60)]
    for c in chunks:
        w.write(c)
    w.close()
    assert gzip.open(p).read() == b"".join(chunks)
    q = str(tmp_path / "e.fq.gz")
    fastq_io.Writer(q, False, 2).close()
    assert gzip.open(q).read() == b""


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints exactly one JSON line with the contract's keys (CPU only)."""
    import json
    import subprocess
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0", "--cpu-sample", "4000"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0 and "workload" in d["config"]
