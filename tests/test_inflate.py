"""The device DEFLATE decoder (sequali_b200/csrc/inflate_core.cuh) compiled for the host, against zlib:
stored, fixed-Huffman and dynamic-Huffman blocks, long matches, overlapping matches, long codes that
miss the primary tables, truncation and corruption.  The -m gpu half (tests/test_gpu_round2.py) runs the
same members through the kernel."""
import ctypes as C
import gzip
import zlib

import numpy as np
import pytest

from sequali_b200 import _lib, synth


def host_inflate(payload: bytes, cap: int):
    lib = _lib.load()
    out = np.zeros(max(cap, 1), np.uint8)
    n = C.c_uint32()
    src = np.frombuffer(payload, np.uint8)
    rc = lib.sq_selftest_inflate_host(src.ctypes.data, len(payload), out.ctypes.data, cap, C.byref(n))
    return rc, out[:n.value].tobytes()


def raw_deflate(data: bytes, level: int, strategy=zlib.Z_DEFAULT_STRATEGY) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return c.compress(data) + c.flush()


def samples():
    rng = np.random.default_rng(5)
    yield b""
    yield b"A"
    yield b"ACGT" * 5000                                      # overlapping matches (distance 4)
    yield bytes(rng.integers(0, 256, 30000, dtype=np.uint8))   # incompressible: stored blocks at any level
    yield synth.illumina_fastq(180, length=150, seed=3, n_tiles=4)
    yield bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), 65280))
    # a skewed alphabet: literal codes longer than the 10-bit primary table
    p = np.array([0.5 ** min(i + 1, 40) for i in range(200)])
    yield bytes(rng.choice(np.arange(200, dtype=np.uint8), 60000, p=p / p.sum()))
    yield bytes(1000) + b"x" * 70000                           # matches of the maximum length, > 64 KiB of text


@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_host_build_of_the_decoder_matches_zlib(level):
    for data in samples():
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY):
            payload = raw_deflate(data, level, strategy)
            rc, got = host_inflate(payload, len(data))
            assert rc == 0 and got == data, (level, strategy, len(data), rc)


def test_truncated_and_corrupt_streams_are_reported():
    data = synth.illumina_fastq(100, length=150, seed=4, n_tiles=2)
    payload = raw_deflate(data, 6)
    assert host_inflate(payload, len(data) - 1)[0] != 0           # output too small
    for cut in (1, 7, len(payload) // 2, len(payload) - 1):
        assert host_inflate(payload[:cut], len(data))[0] != 0, cut
    rng = np.random.default_rng(6)
    bad = 0
    for _ in range(200):  # random bit flips either break the stream or change the text
        b = bytearray(payload)
        b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        rc, got = host_inflate(bytes(b), len(data))
        bad += rc != 0 or got != data
    assert bad >= 195


def test_bgzf_scan_finds_every_member():
    text = synth.illumina_fastq(3000, length=150, seed=7, n_tiles=5)
    stream = synth.bgzf_compress(text, level=6, block_text=40000)
    assert gzip.decompress(stream) == text                         # a valid multi-member gzip file
    lib = _lib.load()
    blocks = (_lib.BgzfBlock * 100)()
    n, used, total = C.c_uint64(), C.c_uint64(), C.c_uint64()
    src = np.frombuffer(stream, np.uint8)
    assert lib.sq_bgzf_scan(src.ctypes.data, len(stream), blocks, 100, C.byref(n), C.byref(used), C.byref(total)) == 0
    assert used.value == len(stream) and total.value == len(text)
    assert n.value == -(-len(text) // 40000) + 1                   # + the empty end-of-file member
    got = b""
    for i in range(n.value):
        b = blocks[i]
        assert b.text_off == len(got)
        rc, part = host_inflate(stream[b.comp_off:b.comp_off + b.comp_len], b.text_len)
        assert rc == 0
        got += part
    assert got == text
    # a member cut short is left to the caller; plain gzip is refused
    assert lib.sq_bgzf_scan(src.ctypes.data, len(stream) - 5, blocks, 100, C.byref(n), C.byref(used), C.byref(total)) == 0
    assert used.value < len(stream) - 5
    plain = np.frombuffer(gzip.compress(text), np.uint8)
    assert lib.sq_bgzf_scan(plain.ctypes.data, len(plain), blocks, 100, C.byref(n), C.byref(used), C.byref(total)) == \
        _lib.SQ_E_FORMAT
