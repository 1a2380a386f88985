"""The ingest workers' whole-buffer DEFLATE decoder (rnaseqc_b200/csrc/host/fastinflate.cpp) against zlib.

It replaces zlib's inflate for BGZF blocks (the reference reads them through htslib's bgzf_read -> zlib, SeqLib/src/BamReader.cpp:233-256).
Contract: whenever it accepts a stream it has produced exactly zlib's bytes; whatever it refuses goes to zlib, so it may refuse
anything — but it must refuse every malformed or truncated stream, never write outside its output and never crash."""
import ctypes as C
import random
import zlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def inflate(libs):
    H = C.CDLL(libs["host"])
    H.rsqh_fast_inflate.restype = C.c_int
    H.rsqh_fast_inflate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]

    def run(comp: bytes, out_len: int):
        """-> bytes or None (refused); a guard zone behind the output proves that nothing is written past out_len"""
        src = np.frombuffer(comp, dtype=np.uint8).copy() if comp else np.zeros(0, np.uint8)
        dst = np.full(out_len + 64, 0xA5, dtype=np.uint8)
        ok = H.rsqh_fast_inflate(src.ctypes.data if len(src) else None, len(src), dst.ctypes.data, out_len)
        assert (dst[out_len:] == 0xA5).all(), "wrote past the end of the output"
        return dst[:out_len].tobytes() if ok else None
    return run


def raw_deflate(data: bytes, level: int, strategy=zlib.Z_DEFAULT_STRATEGY, memlevel=8) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, -15, memlevel, strategy)
    return c.compress(data) + c.flush()


def payloads():
    rng = random.Random(11)
    nrng = np.random.default_rng(12)
    out = {
        "empty": b"",
        "one_byte": b"x",
        "run": b"A" * 70000,
        "short_period": b"ACGT" * 9000 + b"AC" * 300 + b"ACGTTGC" * 2000,
        "text": (b"@SQ\tSN:chr1\tLN:248956422\n" * 400 + bytes(rng.choice(b"ACGTN") for _ in range(30000))),
        "random": nrng.integers(0, 256, 65280, dtype=np.uint8).tobytes(),
        "skewed": nrng.choice(np.arange(256, dtype=np.uint8), 65000, p=np.r_[[0.5, 0.2, 0.1], np.full(253, 0.2 / 253)]).tobytes(),
        "few_symbols": nrng.choice(np.frombuffer(b"IIIIFFF#", dtype=np.uint8), 60000).tobytes(),
    }
    # what a BGZF block of a BAM looks like: fixed fields, names with a counter, packed bases, flat qualities
    rec = bytearray()
    pos = 1000
    while len(rec) < 64000:
        pos += rng.randrange(40)
        name = b"read%09d\0" % rng.randrange(10**9)
        body = (np.int32([0, pos]).tobytes() + bytes([len(name), 60]) + b"\x12\x34" + np.uint16([1, 99]).tobytes() + np.int32([150, 0, pos + 200, 350]).tobytes()
                + name + np.uint32([150 << 4]).tobytes() + nrng.integers(0, 256, 75, dtype=np.uint8).tobytes() + b"I" * 150 + b"NMC\x01")
        rec += np.int32([len(body)]).tobytes() + body
    out["bam_like"] = bytes(rec[:65280])
    return out


@pytest.mark.parametrize("level", [1, 6, 9])
def test_accepted_streams_equal_zlib(inflate, level):
    for name, data in payloads().items():
        for strategy in (zlib.Z_DEFAULT_STRATEGY, zlib.Z_FIXED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE):
            comp = raw_deflate(data, level, strategy)
            got = inflate(comp, len(data))
            if strategy == zlib.Z_HUFFMAN_ONLY and got is None:
                continue                            # an alphabet without distance codes may legitimately be refused ...
            assert got == data, (name, level, strategy)
    # ... but stored blocks (level 0) and several deflate blocks in one stream are not
    big = payloads()["bam_like"] * 4 + payloads()["random"] + payloads()["run"]
    for lvl in (0, level):
        c = zlib.compressobj(lvl, zlib.DEFLATED, -15)
        comp = b"".join(c.compress(big[k:k + 50000]) + c.flush(zlib.Z_FULL_FLUSH) for k in range(0, len(big), 50000)) + c.flush()
        assert inflate(comp, len(big)) == big


def test_wrong_size_truncation_and_garbage_are_refused(inflate):
    data = payloads()["bam_like"]
    comp = raw_deflate(data, 6)
    assert inflate(comp, len(data)) == data
    assert inflate(comp, len(data) - 1) is None            # the stream holds more than the block header promised
    assert inflate(comp, len(data) + 1) is None            # ... or less
    assert inflate(comp + b"\0" * 9, len(data)) == data    # bytes behind the final block are ignored, as by zlib
    for cut in (0, 1, 2, 5, len(comp) // 2, len(comp) - 9, len(comp) - 1):
        assert inflate(comp[:cut], len(data)) is None, cut
    assert inflate(b"\x07", 10) is None                    # block type 3
    assert inflate(b"\x01\x05\x00\xfa\xff" + b"hello", 5) == b"hello"
    assert inflate(b"\x01\x05\x00\xfb\xff" + b"hello", 5) is None     # LEN / NLEN mismatch
    assert inflate(b"\x01\x05\x00\xfa\xff" + b"hell", 5) is None
    # a distance that reaches before the start of the block: fixed code, literal 'a', then match length 3 distance 2
    bits = "1" + "10" + format(0x30 + ord("a"), "08b") + "0000001" + "00001"
    bits += "0000000"                                      # end of block
    by = bytes(int("".join(reversed(bits[k:k + 8].ljust(8, "0"))), 2) for k in range(0, len(bits), 8))
    with pytest.raises(zlib.error):
        zlib.decompress(by, -15)
    assert inflate(by, 4) is None


def test_mutated_streams_never_disagree_with_zlib(inflate):
    """Byte-mutated and bit-flipped streams: either refused, or — where zlib also accepts and yields the promised size — the same bytes."""
    rng = random.Random(5)
    P = payloads()
    accepted = refused = 0
    for name in ("bam_like", "text", "skewed", "few_symbols", "short_period"):
        data = P[name]
        for level in (1, 6):
            comp = bytearray(raw_deflate(data, level))
            for _ in range(150):
                m = bytearray(comp)
                for _ in range(rng.choice((1, 1, 2, 5))):
                    k = rng.randrange(len(m))
                    m[k] = rng.randrange(256) if rng.random() < 0.5 else m[k] ^ (1 << rng.randrange(8))
                got = inflate(bytes(m), len(data))
                try:
                    d = zlib.decompressobj(-15)
                    want = d.decompress(bytes(m))
                    want_ok = d.eof and len(want) == len(data)
                except zlib.error:
                    want_ok = False
                if got is None:
                    refused += 1
                    continue
                accepted += 1
                assert want_ok and got == want, (name, level)
    assert accepted > 20 and refused > 200, (accepted, refused)


def test_random_lengths_and_alphabets(inflate):
    rng = np.random.default_rng(3)
    for trial in range(300):
        n = int(rng.integers(0, 70000)) if trial % 3 else int(rng.integers(0, 400))
        k = int(rng.integers(1, 257))
        alphabet = rng.integers(0, 256, k, dtype=np.uint8)
        p = rng.dirichlet(np.full(k, float(rng.choice([0.05, 0.3, 1.0, 10.0]))))
        data = rng.choice(alphabet, n, p=p).tobytes()
        if trial % 4 == 0 and n > 100:                      # plant long-range repeats
            b = bytearray(data)
            for _ in range(20):
                a, l, t = int(rng.integers(0, n - 50)), int(rng.integers(3, 300)), int(rng.integers(0, n - 50))
                b[t:t + l] = b[a:a + l][:max(0, n - t)]
            data = bytes(b[:n])
        comp = raw_deflate(data, int(rng.choice([1, 4, 6, 9])), memlevel=int(rng.choice([1, 8, 9])))
        assert inflate(comp, len(data)) == data, trial
