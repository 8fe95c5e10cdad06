// inflate_fast.hpp — raw-DEFLATE (RFC 1951) decoder used for BGZF members; see inflate_fast.cpp.
#pragma once
#include <cstddef>
#include <cstdint>

namespace mthh {

// Decodes the raw DEFLATE stream in[0, in_len) into exactly out_len bytes.  Returns false when the stream is malformed,
// does not produce exactly out_len bytes, or overruns its input; the caller then falls back to zlib.  Never writes outside
// out[0, out_len) and never reads outside in[0, in_len).
bool inflate_fast(const uint8_t* in, size_t in_len, uint8_t* out, size_t out_len);

// CRC-32 (gzip polynomial, as zlib's crc32(0, p, n)); PCLMULQDQ folding when the CPU has it.
uint32_t crc32_fast(const uint8_t* p, size_t n);

}  // namespace mthh
