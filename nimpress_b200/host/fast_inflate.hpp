// fast_inflate.hpp -- raw DEFLATE (RFC 1951) decoder for BGZF blocks.
//
// Host ingest is what bounds a run from disk: the reference inflates through
// htslib/zlib on one thread; this engine inflates blocks on a pool of threads, and each thread here
// runs a decoder written for this data -- 64-bit bit buffer refilled eight bytes at a time, one
// table lookup per symbol (10-bit primary table + subtables), word-wise match copies -- instead of
// zlib's general-purpose state machine.  The caller keeps zlib as the fallback: any block this decoder
// does not finish cleanly (return false) or whose CRC32 then differs is inflated again by zlib.
#pragma once
#include <cstddef>
#include <cstdint>

namespace nph {

struct FastInflateTables {                     // per worker thread, reused across blocks
    uint32_t lit[1024 + 288 * 32];
    uint32_t dist[256 + 32 * 128];
};

// in: in_len bytes of raw deflate data, READABLE for 16 bytes beyond (a BGZF block's CRC32 + ISIZE follow, the
// pool pads the rest).
// out: exactly out_len bytes are expected; the buffer must be WRITABLE for 16 bytes beyond.
// true: the stream ended with its final block after exactly out_len bytes.
bool fast_inflate(const uint8_t *in, size_t in_len, uint8_t *out, size_t out_len, FastInflateTables &t);

}  // namespace nph
