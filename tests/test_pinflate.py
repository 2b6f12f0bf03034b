"""The multi-threaded gzip decoder (csrc/aqc_pinflate.cpp) against zlib: FASTQ-like text (block search succeeds), binary
and highly compressible data (search finds nothing / soft stops), concatenated and bgzip-like members, decoy bit
patterns, corruption.  Any thread count must give the bytes zlib gives."""
import ctypes as C
import random
import zlib

import numpy as np
import pytest

from afterqc_b200 import _native


def gunzip_mt(data, threads, cap):
    L = _native.lib()
    src = np.frombuffer(data, dtype=np.uint8)
    out = np.empty(cap + 1, dtype=np.uint8)
    n = C.c_uint64(0)
    stats = (C.c_uint64 * 3)()
    err = C.create_string_buffer(256)
    rc = L.aqc_gunzip_buffer_mt(src.ctypes.data, len(data), out.ctypes.data, cap + 1, C.byref(n), threads, stats, err, 256)
    if rc:
        raise ValueError(err.value.decode() or "rc=%d" % rc)
    return out[:n.value].tobytes(), tuple(int(x) for x in stats)


def gz(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY):
    c = zlib.compressobj(level, zlib.DEFLATED, 31, 8, strategy)
    return c.compress(data) + c.flush()


def fastq_text(n_reads, seed, read_len=150):
    rng = np.random.default_rng(seed)
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, (n_reads, read_len))]
    quals = np.frombuffer(b"#5AFIIII", dtype=np.uint8)[rng.integers(0, 8, (n_reads, read_len))]
    out = []
    for i in range(n_reads):
        out.append(b"@SYN:1:FC:1:1101:%d:%d 1:N:0:A\n" % (i, i * 7)); out.append(bases[i].tobytes()); out.append(b"\n+\n")
        out.append(quals[i].tobytes()); out.append(b"\n")
    return b"".join(out)


FASTQ = fastq_text(60000, 3)          # ~20 MB of text, ~9 MB compressed: several search targets per round


@pytest.mark.parametrize("threads", [1, 2, 3, 8])
@pytest.mark.parametrize("level", [1, 6])
def test_fastq_stream_matches_zlib(threads, level):
    z = gz(FASTQ, level)
    out, (rounds, pieces, false_starts) = gunzip_mt(z, threads, len(FASTQ))
    assert out == FASTQ
    if threads > 1:
        assert pieces > rounds                 # the block search found real boundaries and their pieces chained exactly


def test_binary_and_compressible_data_fall_back_to_the_true_chain():
    rng = np.random.default_rng(1)
    blobs = {
        "random": rng.integers(0, 256, 6 << 20, dtype=np.uint8).tobytes(),                  # stored blocks: nothing to find
        "binary_lz": (rng.integers(0, 256, 50000, dtype=np.uint8).tobytes() * 200),          # dynamic blocks with binary literals
        "zeros": bytes(300 << 20),                                                          # 1000:1: soft stop at the output cap
        "text_runs": b"".join(bytes([65 + (i % 26)]) * (1 + i % 3000) for i in range(20000)),
    }
    for name, data in blobs.items():
        z = gz(data, 6)
        out, stats = gunzip_mt(z, 4, len(data))
        assert out == data, (name, stats)


def test_members_and_bgzip_like_files():
    parts = [FASTQ[i:i + 5_000_000] for i in range(0, len(FASTQ), 5_000_000)]
    z = b"".join(gz(p, 1 + (i % 6)) for i, p in enumerate(parts))
    out, _ = gunzip_mt(z + bytes(100), 4, len(FASTQ))
    assert out == FASTQ
    small = [FASTQ[i:i + 65000] for i in range(0, 4_000_000, 65000)]                        # bgzip-like: many small members
    z = b"".join(gz(p, 6) for p in small)
    out, _ = gunzip_mt(z, 4, len(b"".join(small)))
    assert out == b"".join(small)
    z = gz(b"", 6) + gz(FASTQ[:1000], 6) + gz(b"", 6)
    assert gunzip_mt(z, 4, 1000)[0] == FASTQ[:1000]


def test_decoy_headers_inside_the_payload_do_not_matter():
    """Text that itself contains the bytes of valid deflate blocks (stored inside a stored/literal context) offers the
    search false starts; the chain only accepts exact hand-overs, so the output cannot change."""
    inner = gz(FASTQ[:3_000_000], 6)[10:-8]                                                 # raw deflate blocks of text
    rng = random.Random(4)
    payload = b"".join([FASTQ[:4_000_000], inner, FASTQ[4_000_000:8_000_000], inner[rng.randrange(1, 99):], FASTQ[8_000_000:]])
    for level in (0, 1, 6):
        z = gz(payload, level)
        out, stats = gunzip_mt(z, 4, len(payload))
        assert out == payload, (level, stats)


def test_corruption_fails_loudly():
    z = bytearray(gz(FASTQ, 6))
    for cut in (len(z) // 3, len(z) - 5):
        with pytest.raises(ValueError):
            gunzip_mt(bytes(z[:cut]), 4, len(FASTQ))
    rng = random.Random(2)
    for _ in range(6):
        bad = bytearray(z)
        pos = rng.randrange(20, len(bad) - 8)
        bad[pos] ^= 1 << rng.randrange(8)
        try:
            out, _ = gunzip_mt(bytes(bad), 4, len(FASTQ) + 1000)
            assert out == FASTQ
        except ValueError:
            pass
    bad = bytearray(z); bad[-6] ^= 0x10
    with pytest.raises(ValueError, match="CRC|length"):
        gunzip_mt(bytes(bad), 4, len(FASTQ))


def test_reader_with_decoder_threads_matches_zlib(tmp_path, monkeypatch):
    """NativeStream on a .fq.gz larger than the parallel decoder's threshold: 4 decoder threads vs zlib."""
    from afterqc_b200 import fastq_io
    p = str(tmp_path / "big.fq.gz")
    with open(p, "wb") as f:
        f.write(gz(FASTQ + FASTQ, 1))                      # ~15 MB compressed, one member
    import os
    assert os.path.getsize(p) > (8 << 20)

    def digest():
        import hashlib
        h = hashlib.sha256()
        s = fastq_io.NativeStream(p, 20000)
        n = 0
        while True:
            k = s.available(20000)
            if not k:
                break
            v = s.take(k)
            for c in (v.names, v.seqs, v.plus, v.quals):
                h.update(c.data[int(c.off[0]):int(c.off[-1])].tobytes())
            n += v.n
            v.done()
        s.close()
        return n, h.hexdigest()
    monkeypatch.setenv("AQC_INFLATE_THREADS", "4")
    par = digest()
    monkeypatch.setenv("AQC_INFLATE", "zlib")
    assert digest() == par and par[0] == 120000


def _realistic_fastq(n, seed):
    """Illumina-like names, binned qualities in long runs, occasional N stretches and short reads"""
    rng = np.random.default_rng(seed)
    r = random.Random(seed)
    L = r.choice([76, 101, 151])
    out = []
    for i in range(n):
        name = b"@A00%d:%d:HXXXXDSXX:%d:%d:%d:%d 1:N:0:ACGT+TGCA" % (r.randrange(999), r.randrange(99), r.randrange(1, 5), 1101 + i // 5000,
                                                                     r.randrange(30000), r.randrange(30000))
        seq = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, L)].tobytes()
        if r.random() < 0.02:
            seq = seq[:r.randrange(1, L)] + b"NNNNN" + seq[:3]
        q = bytearray()
        while len(q) < len(seq):
            q += bytes([r.choice(b"FFFFFFFF:,#")]) * r.randrange(1, 60)
        out.append(name + b"\n" + seq + b"\n+\n" + bytes(q[:len(seq)]) + b"\n")
    return b"".join(out)


@pytest.mark.parametrize("case", range(6))
def test_random_zlib_parameters_on_realistic_fastq(case):
    r = random.Random(1000 + case)
    data = _realistic_fastq(r.choice([30000, 60000]), case)
    co = zlib.compressobj(r.randrange(1, 10), zlib.DEFLATED, 31, r.choice([8, 9, 5, 1]),
                          r.choice([zlib.Z_DEFAULT_STRATEGY] * 3 + [zlib.Z_FILTERED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FIXED]))
    step = r.choice([len(data), 1 << 20, 300000])
    parts = []
    for i in range(0, len(data), step):
        parts.append(co.compress(data[i:i + step]))
        if step < len(data) and r.random() < 0.5:
            parts.append(co.flush(r.choice([zlib.Z_SYNC_FLUSH, zlib.Z_FULL_FLUSH])))      # pigz-style flush points
    parts.append(co.flush())
    out, stats = gunzip_mt(b"".join(parts), r.choice([2, 3, 4, 7]), len(data))
    assert out == data, stats
