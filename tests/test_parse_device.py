"""aqc_fastq_parse_device (csrc/aqc_parse.cuh): FASTQ text -> packed columns in HBM, against the host parser aqc_fastq_parse
(csrc/aqc_fastq.cpp; itself pinned to fastq.Reader by tests/test_fastq_io.py) on the same bytes.  Runs on the SIMT emulator
here and on a B200 under -m gpu (tests/test_gpu_parse.py imports these cases)."""
import random

import numpy as np
import pytest

import cases
from afterqc_b200 import _abi, fastq_io


def fastq_text(batch, mate, eol=b"\n", last_newline=True, pad=b""):
    out = []
    for i in range(batch.n):
        s, q = batch.read(mate, i)
        s = s if isinstance(s, bytes) else s.encode()
        q = q if isinstance(q, bytes) else q.encode()
        out += [b"@r%d/%d some text" % (i, mate) + pad, s + pad, b"+", q + pad]
    t = eol.join(out)
    return t + (eol if last_newline else b"")


def host_parse(text, final=True):
    rec, consumed, eof = fastq_io._parse_block(bytes(text), final)
    return rec, consumed, eof


def check_same(engine, text, final=True, max_records=None):
    kw = {} if max_records is None else dict(max_records=max_records)
    p = engine.parse_fastq(text, final=final, **kw)
    if max_records is None:
        rec, consumed, eof = host_parse(text, final)
    else:                                    # the host wrapper has no cap: parse all, cut
        rec, consumed, eof = host_parse(text, final)
    d = p.fetch()
    n_host = 0 if rec is None else len(rec.seqs.off) - 1
    if max_records is not None:
        n_host = min(n_host, max_records)
    assert p.n == n_host
    if p.n == 0:
        return p
    so = rec.seqs.off[:p.n + 1]
    assert np.array_equal(d["off"].astype(np.int64), so - so[0])
    assert np.array_equal(d["seq"], rec.seqs.data[so[0]:so[p.n]])
    assert np.array_equal(d["qual"], rec.quals.data[so[0]:so[p.n]])
    t = np.frombuffer(bytes(text), dtype=np.uint8)
    for k, col in enumerate((rec.names, rec.seqs, rec.plus, rec.quals)):        # the line table reproduces all four columns
        for i in (0, p.n // 2, p.n - 1):
            s, l = int(d["line_start"][4 * i + k]), int(d["line_len"][4 * i + k])
            assert bytes(t[s:s + l]) == bytes(col.data[col.off[i]:col.off[i + 1]])
    if max_records is None:
        assert p.consumed == consumed and p.hit_eof == (eof or final)
    return p


def run_cases(make_engine):
    eng = make_engine(_abi.Params.defaults())
    small = cases.synthetic("pe150", 700, len_jitter=40)
    for kw in (dict(), dict(eol=b"\r\n"), dict(last_newline=False), dict(pad=b" \t"), dict(eol=b"\r\n", last_newline=False)):
        check_same(eng, fastq_text(small, 1, **kw))
    t = fastq_text(small, 2)
    check_same(eng, t, final=False)
    cut = t[:len(t) - 200]                                  # ends inside the sequence line of the last record
    p = check_same(eng, cut, final=False)
    assert p.consumed < len(cut) and not p.hit_eof
    check_same(eng, cut, final=True)                        # the partial record is dropped
    for text in (t[:len(t) - 57],):                         # ends inside the quality line: at the end of the file a short quality line
        with pytest.raises(ValueError):
            eng.parse_fastq(text, final=True)
        with pytest.raises(ValueError):
            host_parse(text, True)
        check_same(eng, text, final=False)
    check_same(eng, t, max_records=123)
    lines = t.split(b"\n")
    check_same(eng, b"\n".join(lines[:4 * 300] + [b""] + lines[4 * 300:]))      # an empty line ends the file
    check_same(eng, b"\n".join(lines[:4 * 10 + 2] + [b"   "] + lines[4 * 10 + 3:]))   # ... also one of blanks, inside a record
    assert eng.parse_fastq(b"").n == 0
    for text in (b"\n", b"@x\n\nAC", b"@x\nAC\n  \n"):      # an empty line among fewer than four lines still ends the file
        p, (rec, consumed, eof) = eng.parse_fastq(text, final=False), host_parse(text, False)
        assert p.n == 0 and rec is None and p.hit_eof and eof
    p = eng.parse_fastq(b"@x\nAC\n+", final=False)
    assert p.n == 0 and not p.hit_eof
    assert eng.parse_fastq(b"@x\nACGT\n", final=True).n == 0
    bad = list(lines)
    bad[4 * 55 + 3] = bad[4 * 55 + 3][:-1]
    with pytest.raises(ValueError, match="record 55"):
        eng.parse_fastq(b"\n".join(bad))
    with pytest.raises(ValueError, match="record 55"):
        host_parse(b"\n".join(bad))
    rng = random.Random(5)
    for _ in range(6):                                      # odd sizes around the 4 KB blocks and the 16-byte words
        n = rng.randint(1, 60)
        b = cases.synthetic("pe150", n, len_jitter=100)
        check_same(eng, fastq_text(b, 1, last_newline=rng.random() < 0.5))
    # many tiny records: the scan of the sequence lengths runs over more than 256 partial sums
    tiny = b"".join(b"@%d\n%s\n+\n%s\n" % (i, b"ACGTN"[i % 5:i % 5 + 1] * (1 + i % 3), b"IJK"[i % 3:i % 3 + 1] * (1 + i % 3)) for i in range(270000))
    p = check_same(eng, tiny)
    assert p.n == 270000
    # the parsed columns are a resident batch: same records as the host batch through the filter
    batch = cases.synthetic("pe150", 1500)
    p1 = eng.parse_fastq(fastq_text(batch, 1), slot=0)
    p2 = eng.parse_fastq(fastq_text(batch, 2), slot=1)
    from afterqc_b200.engine import ParsedDeviceBatch
    db = ParsedDeviceBatch(eng, p1, p2)
    eng.filter_pairs(db)
    a = eng.fetch_results(db)
    db.free()
    eng.reset()
    b = eng.filter_pairs(batch)
    assert np.array_equal(a, b)
    eng.close()


def test_parse_device_on_emulator():
    import emu
    run_cases(lambda p: emu.EmuEngine(p))


def test_parse_device_random_small_texts():
    """300 random little texts: records of random lengths, stray blanks / CR / tabs at line ends, blank or empty lines, a
    missing last newline, a bad record here and there; the device parser and the host parser must say the same"""
    import emu
    eng = emu.EmuEngine(_abi.Params.defaults())
    rng = random.Random(11)
    for case in range(300):
        lines = []
        for r in range(rng.randint(0, 12)):
            L = rng.randint(1, 40)
            qual_len = L if rng.random() > 0.04 else max(0, L + rng.choice((-1, 1)))
            rec = [b"@n%d" % r, bytes(rng.choice(b"ACGTN") for _ in range(L)), b"+", bytes(rng.randint(33, 73) for _ in range(qual_len))]
            for k in range(4):
                if rng.random() < 0.1:
                    rec[k] += rng.choice((b" ", b"\t", b"\r", b" \t\r"))
            if rng.random() < 0.03:
                rec[rng.randrange(4)] = rng.choice((b"", b" ", b"\t\r"))
            lines += rec
        if rng.random() < 0.2 and lines:
            lines = lines[:rng.randint(1, len(lines))]           # ends inside a record
        text = b"\n".join(lines) + (b"\n" if rng.random() < 0.6 and lines else b"")
        for final in (True, False):
            try:
                rec, consumed, eof = host_parse(text, final)
                err = None
            except ValueError as e:
                err = str(e)
            if err is not None:
                with pytest.raises(ValueError) as ei:
                    eng.parse_fastq(text, final=final)
                assert str(ei.value) == err, (case, text)
                continue
            p = eng.parse_fastq(text, final=final)
            n_host = 0 if rec is None else len(rec.seqs.off) - 1
            assert p.n == n_host and p.consumed == consumed and p.hit_eof == (eof or final), (case, final, text, p.n, n_host, p.consumed, consumed, p.hit_eof, eof)
            if p.n:
                d = p.fetch()
                assert np.array_equal(d["seq"], rec.seqs.data[:int(rec.seqs.off[p.n])]) and np.array_equal(d["qual"], rec.quals.data[:int(rec.seqs.off[p.n])])
                assert np.array_equal(d["off"].astype(np.int64), rec.seqs.off)
                t = np.frombuffer(text, dtype=np.uint8)
                for i in range(p.n):
                    s, l = int(d["line_start"][4 * i]), int(d["line_len"][4 * i])
                    assert bytes(t[s:s + l]) == bytes(rec.names.data[rec.names.off[i]:rec.names.off[i + 1]])
    eng.close()
