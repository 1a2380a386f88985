// fastinflate.h — whole-buffer raw-DEFLATE decoder for BGZF blocks (RFC 1951).
//
// BGZF hands the ingest workers what a streaming inflater is not built for: a complete compressed block and the exact size of
// its output, both in memory.  fast_inflate() uses that: a 64-bit bit buffer refilled with one unaligned load, an 11-bit
// literal/length table whose entries carry base value, extra-bit count and code length in one word, literals decoded two or
// three per refill, matches copied eight bytes at a time — no state machine, no window, no per-call allocation.
//
// It is STRICT: it returns false for anything but a well-formed stream that ends exactly at `out_len` bytes with its final
// block (over-subscribed or incomplete code sets, symbols 286/287 or distance codes 30/31, a distance before the start of the
// block, input or output running short).  Callers then hand the block to zlib, which either inflates it or produces the error
// the run reported before this decoder existed — results and error behaviour cannot change (tests/test_fastinflate.py:
// byte-identical against zlib on compressible, incompressible, stored, fixed-code and truncated / mutated streams).
#pragma once
#include <cstddef>
#include <cstdint>

namespace rsqh {

bool fast_inflate(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);

}  // namespace rsqh
