// inflate.cuh — raw DEFLATE (RFC 1951) on the GPU, one WARP per BGZF member.
//
// The reference reads BAM through htslib, which inflates one BGZF member after the other on one thread (bamutil.rs:4-11).
// BGZF members are independent raw-DEFLATE streams of at most 64 KiB of output, so a file is thousands of independent
// decode jobs: here each warp takes one member.  Huffman decoding is inherently serial, so all 32 lanes decode the SAME
// bit stream redundantly (identical registers, broadcast loads: no divergence, no shuffles) and the lanes only split the
// work that is parallel: filling the decode tables of a dynamic block and copying LZ77 matches.  Parallelism comes from
// the number of members in flight (148 SMs x up to 48 warps), not from within a stream.
//
// Decode tables per warp in shared memory: a 10-bit primary look-up for literal/length codes and an 8-bit one for
// distance codes (entry = symbol << 4 | code length); longer codes fall back to the canonical count/first-code walk
// (the method of zlib's contrib/puff).  Output goes straight to global memory; matches read it back through L2
// (ld.global.cg), ordered by __syncwarp().
#pragma once
#include "common.cuh"

namespace mth {

constexpr int INF_LL_BITS = 10, INF_D_BITS = 8;
constexpr int INF_MAXBITS = 15, INF_MAXL = 288, INF_MAXD = 30;

struct InflateTables {                      // per warp
    unsigned int ring[64];                  // compressed input staging (BitReader)
    uint16_t ll_fast[1 << INF_LL_BITS];     // (symbol << 4) | length, 0 = not a short code
    uint16_t d_fast[1 << INF_D_BITS];
    uint16_t ll_sym[INF_MAXL + 32], d_sym[INF_MAXD + 2];  // symbols ordered by (length, symbol): canonical order
    uint16_t ll_cnt[INF_MAXBITS + 1], d_cnt[INF_MAXBITS + 1];
    uint8_t len[INF_MAXL + INF_MAXD + 34];  // code lengths of the block being set up
};

enum : int { INF_OK = 0, INF_ERR_BTYPE = 1, INF_ERR_STORED = 2, INF_ERR_CODE = 3, INF_ERR_DIST = 4, INF_ERR_OVERRUN = 5,
             INF_ERR_INPUT = 6, INF_ERR_SIZE = 7, INF_ERR_TABLE = 8, INF_ERR_CRC = 9 };

// LSB-first bit reader; every lane of the warp holds the same state.  The compressed bytes come through a 256-byte ring in
// shared memory that the warp refills with one coalesced load (2 words per lane) every 64 words, so the serial decode loop
// waits on shared memory (~30 cycles), not on global memory (~500), for its input.
constexpr int INF_RING_WORDS = 64;
struct BitReader {
    const unsigned int* g;      // the stream as 32-bit words from the 4-byte aligned address at or below its first byte
    unsigned int* ring;         // shared memory, INF_RING_WORDS words, this warp's
    uint32_t n_words;           // words that contain stream bytes
    uint32_t n_bytes;           // stream length from the aligned base (skip + in_len)
    uint32_t rw;                // next word to move into the bit buffer
    uint32_t filled;            // words [0, filled) have been staged (multiple of INF_RING_WORDS)
    unsigned long long buf;
    int cnt;
    __device__ __forceinline__ void fill() {  // stage words [filled, filled + 64): every earlier word has been consumed
        const int lane = lane_id();
        __syncwarp();
#pragma unroll
        for (int k = 0; k < INF_RING_WORDS / 32; k++) {
            const uint32_t w = filled + lane + 32 * k;
            ring[lane + 32 * k] = w < n_words ? __ldg(g + w) : 0u;
        }
        filled += INF_RING_WORDS;
        __syncwarp();
    }
    __device__ __forceinline__ void init(const uint8_t* b, uint32_t len, unsigned int* ring_) {
        const uint32_t skip = (uint32_t)((uintptr_t)b & 3u);
        g = reinterpret_cast<const unsigned int*>(b - skip);
        ring = ring_;
        n_bytes = skip + len;
        n_words = (n_bytes + 3) >> 2;
        rw = 0; filled = 0; buf = 0; cnt = 0;
        // bytes of the first / last word outside the stream are never interpreted: `skip` bytes are dropped here and the decoder
        // stops at the end-of-block code (consumption past n_bytes is detected by overrun())
        refill();
        drop((int)skip * 8);
    }
    __device__ __forceinline__ void refill() {  // at least 32 valid bits afterwards (zeros past the end)
        while (cnt <= 32) {
            if (rw >= filled) fill();
            buf |= (unsigned long long)ring[rw & (INF_RING_WORDS - 1)] << cnt;
            rw++;
            cnt += 32;
        }
    }
    __device__ __forceinline__ uint32_t peek(int n) const { return (uint32_t)(buf & ((1ull << n) - 1ull)); }
    __device__ __forceinline__ void drop(int n) { buf >>= n; cnt -= n; }
    __device__ __forceinline__ uint32_t bits(int n) {  // n <= 16
        if (cnt < n) refill();
        uint32_t v = peek(n);
        drop(n);
        return v;
    }
    // byte position (from the aligned base) of the next unread bit, which must be byte aligned
    __device__ __forceinline__ uint32_t byte_pos() const { return rw * 4u - (uint32_t)(cnt >> 3); }
    __device__ __forceinline__ void seek(uint32_t pos) {  // continue at byte `pos`
        rw = pos >> 2;
        filled = rw & ~(uint32_t)(INF_RING_WORDS - 1);
        buf = 0; cnt = 0;
        fill();
        refill();
        drop((int)(pos & 3u) * 8);
    }
    __device__ __forceinline__ bool overrun() const { return (long long)rw * 32 - cnt > (long long)n_bytes * 8; }  // consumed bits past the end
};

// Canonical Huffman set-up from code lengths len[0..n): counts per length, symbols in canonical order (lane 0 — a few
// hundred steps), then the fast table in parallel (each lane fills the entries of its symbols).  Returns false for an
// over-subscribed code.
__device__ __forceinline__ bool build_table(const uint8_t* len, int n, uint16_t* cnt, uint16_t* sym, uint16_t* fast, int fast_bits) {
    const int lane = lane_id();
    __syncwarp();
    bool ok = true;
    if (lane == 0) {
        uint16_t offs[INF_MAXBITS + 2];
        for (int l = 0; l <= INF_MAXBITS; l++) cnt[l] = 0;
        for (int s = 0; s < n; s++) cnt[len[s]]++;
        int left = 1;
        for (int l = 1; l <= INF_MAXBITS; l++) {
            left <<= 1;
            left -= cnt[l];
            if (left < 0) ok = false;  // over-subscribed
        }
        offs[1] = 0;
        for (int l = 1; l < INF_MAXBITS; l++) offs[l + 1] = offs[l] + cnt[l];
        for (int s = 0; s < n; s++)
            if (len[s]) sym[offs[len[s]]++] = (uint16_t)s;
    }
    ok = __shfl_sync(FULL, ok ? 1 : 0, 0) != 0;
    for (int i = lane; i < (1 << fast_bits); i += 32) fast[i] = 0;
    __syncwarp();
    if (!ok) return false;
    // canonical codes: symbols of length l are numbered consecutively from first[l]; entry index = bit-reversed code
    int first = 0, index = 0;
    for (int l = 1; l <= fast_bits; l++) {
        const int c = cnt[l];
        for (int k = lane; k < c; k += 32) {
            const uint32_t code = (uint32_t)(first + k);
            const uint32_t rev = __brev(code) >> (32 - l);
            const uint16_t e = (uint16_t)((sym[index + k] << 4) | l);
            for (uint32_t x = rev; x < (1u << fast_bits); x += (1u << l)) fast[x] = e;
        }
        index += c;
        first = (first + c) << 1;
    }
    __syncwarp();
    return true;
}

// One symbol: fast table, else the canonical walk over the lengths above the fast width.
__device__ __forceinline__ int decode_sym(BitReader& br, const uint16_t* cnt, const uint16_t* sym, const uint16_t* fast, int fast_bits) {
    if (br.cnt < INF_MAXBITS) br.refill();
    const uint32_t e = fast[br.peek(fast_bits)];
    if (e) {
        br.drop((int)(e & 15u));
        return (int)(e >> 4);
    }
    int code = 0, first = 0, index = 0;
    unsigned long long b = br.buf;
    for (int l = 1; l <= INF_MAXBITS; l++) {
        code |= (int)(b & 1ull);
        b >>= 1;
        const int c = cnt[l];
        if (code - c < first) {
            br.drop(l);
            return sym[index + (code - first)];
        }
        index += c;
        first += c;
        first <<= 1;
        code <<= 1;
    }
    return -1;
}

// length / distance codes -> base value and number of extra bits (RFC 1951 3.2.5), computed instead of looked up
__device__ __forceinline__ void len_code(int ls, uint32_t* base, int* ext) {   // ls = symbol - 257, 0..28
    if (ls < 8) { *base = 3u + (uint32_t)ls; *ext = 0; return; }
    if (ls == 28) { *base = 258u; *ext = 0; return; }
    const int e = (ls - 4) >> 2;
    *ext = e;
    *base = 3u + ((4u + ((uint32_t)ls & 3u)) << e);
}
__device__ __forceinline__ void dist_code(int ds, uint32_t* base, int* ext) {  // 0..29
    if (ds < 4) { *base = 1u + (uint32_t)ds; *ext = 0; return; }
    const int e = (ds - 2) >> 1;
    *ext = e;
    *base = 1u + ((2u + ((uint32_t)ds & 1u)) << e);
}
__device__ const uint8_t INF_CLORDER[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};

// Inflates one raw-DEFLATE stream in[0, in_len) into out[0, out_cap); all 32 lanes of a warp call it with the same
// arguments.  Returns INF_OK and the number of bytes produced in *produced.
__device__ __forceinline__ int warp_inflate(const uint8_t* __restrict__ in, uint32_t in_len, uint8_t* out, uint32_t out_cap, InflateTables& T,
                                            uint32_t* produced) {
    const int lane = lane_id();
    BitReader br;
    br.init(in, in_len, T.ring);
    uint32_t pos = 0;
    int err = INF_OK;
    bool last = false;
    while (!last && err == INF_OK) {
        last = br.bits(1) != 0;
        const uint32_t type = br.bits(2);
        if (type == 0) {  // stored
            br.drop(br.cnt & 7);
            const uint32_t q0 = br.byte_pos();  // LEN NLEN data, byte aligned
            if (q0 + 4 > br.n_bytes) { err = INF_ERR_INPUT; break; }
            const uint8_t* q = reinterpret_cast<const uint8_t*>(br.g) + q0;
            const uint32_t len = (uint32_t)__ldg(q) | ((uint32_t)__ldg(q + 1) << 8), nlen = (uint32_t)__ldg(q + 2) | ((uint32_t)__ldg(q + 3) << 8);
            q += 4;
            if ((len ^ 0xFFFFu) != nlen || q0 + 4 + len > br.n_bytes) { err = INF_ERR_STORED; break; }
            if (pos + len > out_cap) { err = INF_ERR_SIZE; break; }
            for (uint32_t k = lane; k < len; k += 32) out[pos + k] = __ldg(q + k);
            __syncwarp();
            pos += len;
            br.seek(q0 + 4 + len);
            continue;
        }
        if (type == 3) { err = INF_ERR_BTYPE; break; }
        if (type == 1) {  // fixed code
            for (int s = lane; s < 288; s += 32) T.len[s] = s < 144 ? 8 : s < 256 ? 9 : s < 280 ? 7 : 8;
            for (int s = lane; s < 30; s += 32) T.len[288 + s] = 5;
            __syncwarp();
            if (!build_table(T.len, 288, T.ll_cnt, T.ll_sym, T.ll_fast, INF_LL_BITS) ||
                !build_table(T.len + 288, 30, T.d_cnt, T.d_sym, T.d_fast, INF_D_BITS)) { err = INF_ERR_TABLE; break; }
        } else {  // dynamic code
            const int nlen = (int)br.bits(5) + 257, ndist = (int)br.bits(5) + 1, ncode = (int)br.bits(4) + 4;
            if (nlen > 286 || ndist > 30) { err = INF_ERR_TABLE; break; }
            // code-length code: reuse the distance tables as scratch (7-bit codes, 19 symbols)
            for (int s = lane; s < 19; s += 32) T.len[s] = 0;
            __syncwarp();
            for (int k = 0; k < ncode; k++) {
                const uint32_t v = br.bits(3);
                if (lane == 0) T.len[INF_CLORDER[k]] = (uint8_t)v;
            }
            __syncwarp();
            if (!build_table(T.len, 19, T.d_cnt, T.d_sym, T.d_fast, 7)) { err = INF_ERR_TABLE; break; }
            // the literal/length and distance code lengths, run-length coded with the code-length code: decode into a
            // second area (the code-length lengths sit in T.len[0..19) while they are in use)
            uint8_t* L = T.len + 32;
            int idx = 0;
            while (idx < nlen + ndist) {
                const int s = decode_sym(br, T.d_cnt, T.d_sym, T.d_fast, 7);
                if (s < 0) { err = INF_ERR_CODE; break; }
                if (s < 16) {
                    if (lane == 0) L[idx] = (uint8_t)s;
                    idx++;
                } else {
                    int prev = 0, rep;
                    if (s == 16) {
                        if (idx == 0) { err = INF_ERR_CODE; break; }
                        __syncwarp();
                        prev = L[idx - 1];
                        rep = 3 + (int)br.bits(2);
                    } else if (s == 17) {
                        rep = 3 + (int)br.bits(3);
                    } else {
                        rep = 11 + (int)br.bits(7);
                    }
                    if (idx + rep > nlen + ndist) { err = INF_ERR_CODE; break; }
                    if (lane == 0)
                        for (int k = 0; k < rep; k++) L[idx + k] = (uint8_t)prev;
                    idx += rep;
                    __syncwarp();
                }
            }
            if (err != INF_OK) break;
            __syncwarp();
            if (L[256] == 0) { err = INF_ERR_CODE; break; }  // no end-of-block code
            if (!build_table(L, nlen, T.ll_cnt, T.ll_sym, T.ll_fast, INF_LL_BITS) ||
                !build_table(L + nlen, ndist, T.d_cnt, T.d_sym, T.d_fast, INF_D_BITS)) { err = INF_ERR_TABLE; break; }
        }
        // ---- the symbols of the block ----
        // (Deferring a match's stores behind the decoding of the next symbols, to hide the L2 latency of its loads, was measured
        // SLOWER — 52 ms against 35 ms for a 918 MB BAM: the kernel is bound by issued instructions, not by that latency.)
        while (true) {
            const int s = decode_sym(br, T.ll_cnt, T.ll_sym, T.ll_fast, INF_LL_BITS);
            if (s < 0) { err = INF_ERR_CODE; break; }
            if (s < 256) {
                if (pos >= out_cap) { err = INF_ERR_SIZE; break; }
                if (lane == 0) out[pos] = (uint8_t)s;
                pos++;
                continue;
            }
            if (s == 256) break;
            const int ls = s - 257;
            if (ls >= 29) { err = INF_ERR_CODE; break; }
            uint32_t lbase, dbase;
            int lext, dext;
            len_code(ls, &lbase, &lext);
            const uint32_t len = lbase + br.bits(lext);
            const int ds = decode_sym(br, T.d_cnt, T.d_sym, T.d_fast, INF_D_BITS);
            if (ds < 0 || ds >= 30) { err = INF_ERR_CODE; break; }
            dist_code(ds, &dbase, &dext);
            const uint32_t dist = dbase + br.bits(dext);
            if (dist > pos) { err = INF_ERR_DIST; break; }
            if (pos + len > out_cap) { err = INF_ERR_SIZE; break; }
            __syncwarp();  // the bytes written so far (literals by lane 0) are visible to every lane
            const uint8_t* src = out + pos - dist;
            for (uint32_t k = lane; k < len; k += 32) out[pos + k] = __ldcg(src + (dist >= len ? k : k % dist));
            __syncwarp();
            pos += len;
        }
        if (br.overrun()) err = INF_ERR_OVERRUN;
    }
    *produced = pos;
    return err;
}

// ---- CRC-32 of a member's output (the gzip trailer of every BGZF member carries it; htslib checks it) -------------------------
// Each lane takes one contiguous 1/32 of the bytes with the classic byte-wise table (in shared memory), then the 32 partial
// values are joined with crc(A || B) = crc(A) * x^(8 |B|) mod P  xor  crc(B) (zlib's crc32_combine, polynomial arithmetic on the
// bit-reflected representation).  ~1 % of the instructions the inflate itself needs.
constexpr uint32_t CRC_POLY = 0xedb88320u;

__device__ __forceinline__ uint32_t crc_multmodp(uint32_t a, uint32_t b) {  // a(x) * b(x) mod p(x), reflected
    uint32_t m = 1u << 31, p = 0;
    for (;;) {
        if (a & m) {
            p ^= b;
            if ((a & (m - 1)) == 0) break;
        }
        m >>= 1;
        b = (b & 1u) ? (b >> 1) ^ CRC_POLY : b >> 1;
    }
    return p;
}
__device__ __forceinline__ uint32_t crc_x2nmodp(uint32_t n, uint32_t k) {  // x^(n * 2^k) mod p(x)
    uint32_t p = 1u << 31, q = 1u << 30;  // q = x^(2^0)
    for (uint32_t i = 0; i < k; i++) q = crc_multmodp(q, q);
    while (n) {
        if (n & 1u) p = crc_multmodp(q, p);
        n >>= 1;
        if (n) q = crc_multmodp(q, q);
    }
    return p;
}
__device__ __forceinline__ void crc_table_init(uint32_t* tab256) {  // by the whole CTA
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        uint32_t c = i;
#pragma unroll
        for (int k = 0; k < 8; k++) c = (c & 1u) ? (c >> 1) ^ CRC_POLY : c >> 1;
        tab256[i] = c;
    }
}
// all 32 lanes of a warp; returns the CRC-32 of data[0, n) (same value on every lane)
__device__ __forceinline__ uint32_t warp_crc32(const uint8_t* data, uint32_t n, const uint32_t* __restrict__ tab256) {
    const int lane = lane_id();
    const uint32_t chunk = (n + 31u) >> 5;
    const uint32_t a = min(n, chunk * (uint32_t)lane), b = min(n, a + chunk);
    uint32_t c = 0xffffffffu;
    for (uint32_t k = a; k < b; k++) c = tab256[(c ^ __ldcg(data + k)) & 0xffu] ^ (c >> 8);
    c ^= 0xffffffffu;             // the CRC of this lane's bytes (0 for an empty range)
    uint32_t len = b - a;         // bytes it covers
    // join neighbours pairwise: after round r lane l (l % 2^(r+1) == 0) holds the CRC of 2^(r+1) chunks
#pragma unroll
    for (int r = 0; r < 5; r++) {
        const uint32_t oc = __shfl_down_sync(FULL, c, 1u << r), ol = __shfl_down_sync(FULL, len, 1u << r);
        if ((lane & ((2 << r) - 1)) == 0) {
            if (ol) c = crc_multmodp(crc_x2nmodp(ol, 3), c) ^ oc;
            len += ol;
        }
    }
    return __shfl_sync(FULL, c, 0);
}

}  // namespace mth
