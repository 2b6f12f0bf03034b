"""The reader's own gzip/DEFLATE decoder (csrc/aqc_inflate.cpp) against zlib: every block type, compression level and
strategy, concatenated members, header flags, piece boundaries, corrupt and truncated streams (must fail, never crash)."""
import ctypes as C
import gzip
import io
import os
import random
import zlib

import numpy as np
import pytest

from afterqc_b200 import _native


def gunzip(data, cap=None):
    L = _native.lib()
    cap = cap if cap is not None else 64
    src = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, dtype=np.uint8)
    while True:
        out = np.empty(max(cap, 1), dtype=np.uint8)
        n = C.c_uint64(0)
        err = C.create_string_buffer(256)
        rc = L.aqc_gunzip_buffer(src.ctypes.data, len(data), out.ctypes.data, cap, C.byref(n), err, 256)
        if rc == 5:              # AQC_ERR_NOMEM: grow
            cap = cap * 4 + 1024
            continue
        if rc:
            raise ValueError(err.value.decode())
        return out[:n.value].tobytes()


def deflate_gz(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=31, memlevel=8):
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, memlevel, strategy)
    return c.compress(data) + c.flush()


def corpus():
    rng = random.Random(5)
    nprng = np.random.default_rng(5)
    fastq = b"".join(b"@SYN:1:FC:1:1101:%d:%d 1:N:0:A\n%s\n+\n%s\n" % (i, i * 7, bytes(rng.choice(b"ACGT") for _ in range(150)),
                                                                     bytes(rng.choice(b"#5AFIIII") for _ in range(150))) for i in range(3000))
    return {
        "empty": b"",
        "one": b"x",
        "fastq": fastq,
        "random": nprng.integers(0, 256, 300000, dtype=np.uint8).tobytes(),
        "zeros": bytes(500000),
        "runs": b"".join(bytes([rng.randrange(256)]) * rng.randrange(1, 700) for _ in range(2000)),
        "short_period": (b"abcdefg" * 50000) + (b"xy" * 70000) + (b"pqr" * 40000),
        "far_matches": (lambda blk: blk + nprng.integers(0, 256, 32768 - 300, dtype=np.uint8).tobytes() + blk + blk[:100] * 400)(
            nprng.integers(0, 4, 300, dtype=np.uint8).tobytes()),
        "text": (b"the quick brown fox jumps over the lazy dog. " * 20000)[:777777],
    }


CORPUS = corpus()


@pytest.mark.parametrize("name", sorted(CORPUS))
def test_levels_and_strategies_roundtrip(name):
    data = CORPUS[name]
    for level in (0, 1, 2, 6, 9):
        assert gunzip(deflate_gz(data, level)) == data, (name, level)
    for strat in (zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED):
        assert gunzip(deflate_gz(data, 6, strat)) == data, (name, strat)
    for memlevel in (1, 9):
        assert gunzip(deflate_gz(data, 6, memlevel=memlevel)) == data
    assert gunzip(deflate_gz(data, 6, wbits=25)) == data             # 512-byte window


def test_members_flags_and_padding():
    a, b, c = CORPUS["fastq"][:100000], CORPUS["text"][:50000], CORPUS["runs"][:70000]
    blob = deflate_gz(a, 1) + deflate_gz(b"", 6) + deflate_gz(b, 9) + deflate_gz(c, 0)
    assert gunzip(blob) == a + b + c
    assert gunzip(blob + bytes(1000)) == a + b + c                  # zero padding after the last member (python's gzip accepts it)
    bio = io.BytesIO()
    with gzip.GzipFile(filename="some_name.fq", mode="wb", fileobj=bio, mtime=123) as f:      # FNAME header field
        f.write(a)
    assert gunzip(bio.getvalue()) == a
    # hand-made header with FEXTRA + FNAME + FCOMMENT + FHCRC
    raw = deflate_gz(b, 6, wbits=-15)
    hdr = bytes([0x1f, 0x8b, 8, 2 | 4 | 8 | 16, 0, 0, 0, 0, 0, 255]) + (5).to_bytes(2, "little") + b"EXTRA" + b"name\0" + b"comment\0"
    hdr += (zlib.crc32(hdr) & 0xFFFF).to_bytes(2, "little")
    trailer = zlib.crc32(b).to_bytes(4, "little") + (len(b) & 0xFFFFFFFF).to_bytes(4, "little")
    assert gunzip(hdr + raw + trailer) == b
    # flush points (empty stored blocks) and many small blocks
    co = zlib.compressobj(6, zlib.DEFLATED, 31)
    parts = []
    for i in range(0, len(a), 777):
        parts.append(co.compress(a[i:i + 777])); parts.append(co.flush(zlib.Z_SYNC_FLUSH if i % 2 else zlib.Z_FULL_FLUSH))
    parts.append(co.flush())
    assert gunzip(b"".join(parts)) == a


def test_output_larger_than_the_piece_and_small_reads():
    data = CORPUS["fastq"] * 6 + CORPUS["random"] + CORPUS["short_period"]           # > 1 MiB pieces, window slides
    z = deflate_gz(data, 6)
    assert gunzip(z, cap=len(data)) == data                    # exact capacity: the stream must end exactly there
    assert gunzip(z, cap=7) == data                            # grown from tiny capacities (restarts)


def test_corrupt_streams_fail_loudly():
    data = CORPUS["fastq"][:200000]
    z = bytearray(deflate_gz(data, 6))
    for cut in (0, 5, 10, 11, 200, len(z) // 2, len(z) - 9, len(z) - 1):
        with pytest.raises(ValueError):
            gunzip(bytes(z[:cut]))
    bad = bytearray(z); bad[-5] ^= 1                           # CRC
    with pytest.raises(ValueError, match="CRC|length"):
        gunzip(bytes(bad))
    bad = bytearray(z); bad[-1] ^= 1                           # ISIZE
    with pytest.raises(ValueError, match="length"):
        gunzip(bytes(bad))
    with pytest.raises(ValueError, match="not a gzip"):
        gunzip(b"hello world, this is not gzip")
    with pytest.raises(ValueError, match="garbage"):
        gunzip(bytes(z) + b"garbage after the member")
    rng = random.Random(9)
    survived = 0
    for _ in range(400):                                        # random bit flips / byte smashes: error or identical output, never a crash
        bad = bytearray(z)
        for _k in range(rng.randrange(1, 4)):
            pos = rng.randrange(10, len(bad) - 8)
            bad[pos] = rng.randrange(256) if rng.random() < 0.5 else bad[pos] ^ (1 << rng.randrange(8))
        try:
            out = gunzip(bytes(bad))
            assert out == data                                  # only a no-op mutation can pass the CRC
            survived += 1
        except ValueError:
            pass
    assert survived < 40


def test_fuzz_small_streams_against_zlib():
    rng = random.Random(17)
    for i in range(300):
        n = rng.choice([0, 1, 2, 3, 7, 30, 258, 259, 1000, 5000])
        alpha = rng.choice([b"A", b"AC", b"ACGT", bytes(range(256))])
        data = bytes(rng.choice(alpha) for _ in range(n))
        z = deflate_gz(data, rng.choice([0, 1, 6, 9]), rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY]))
        assert gunzip(z) == data, i


def test_reader_uses_own_decoder_and_zlib_agree(tmp_path, monkeypatch):
    from afterqc_b200 import fastq_io
    p = str(tmp_path / "a.fq.gz")
    with open(p, "wb") as f:
        f.write(deflate_gz(CORPUS["fastq"][:400000], 6) + deflate_gz(CORPUS["fastq"][400000:], 1))

    def drain():
        s = fastq_io.NativeStream(p, 500)
        out = []
        while True:
            k = s.available(500)
            if not k:
                break
            v = s.take(k); out.append((v.names.data[int(v.names.off[0]):int(v.names.off[-1])].tobytes(), v.seqs.data[int(v.seqs.off[0]):int(v.seqs.off[-1])].tobytes())); v.done()
        s.close()
        return out
    own = drain()
    monkeypatch.setenv("AQC_INFLATE", "zlib")
    assert drain() == own and sum(len(x[1]) for x in own) == 3000 * 150
