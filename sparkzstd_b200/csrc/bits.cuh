// bits.cuh -- bit readers used by every entropy stage.
//
// RevBits replaces bitstream/reversebitstream.go:3-88 (Reversebitstream.Read /
// BitsStillInStream): bits are consumed from the MSB of the LAST byte towards byte 0, the
// first bit consumed is the MSB of the returned value, and bits "below" byte 0 read as zero
// while the remaining-bit count keeps going negative (callers detect over-read from it).
// Instead of div/mod per call, a 64-bit MSB-aligned window is kept in registers and topped up
// with aligned 32-bit loads walking backwards through global memory.
//
// FwdBits replaces bitstream/bitstream.go:8-90 (Bitstream.Read / UnwindBit): LSB-first
// forward reads; only the FSE table descriptions use it, so it is position based and
// "unwind one bit" is pos--.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SZB_HD __host__ __device__ __forceinline__
#else
#define SZB_HD inline
#endif

namespace szb {

struct RevBits {
    const uint8_t *base;  // first byte of the stream
    int32_t next;         // bytes [0, next) are not loaded yet
    int32_t avail;        // valid bits at the top of win (64 once the stream start was reached)
    int64_t remaining;    // real bits not consumed yet; negative after an over-read
    uint64_t win;         // next bits to consume, MSB first; zero below `avail`
};

SZB_HD uint32_t load_u32_aligned(const uint8_t *p) { return *reinterpret_cast<const uint32_t *>(p); }

// Top the window up so that at least 33 bits are valid (or the stream start is reached).
SZB_HD void rev_refill(RevBits &r) {
    if (r.avail <= 32) {
        if (r.next >= 4) {
            uint32_t w = load_u32_aligned(r.base + r.next - 4);
#if defined(__CUDA_ARCH__)
            // lanes of a warp walk unrelated streams in lock step: one lane's miss stalls all 32, so
            // every lane pulls the sectors it will need two refills-of-a-sector ahead into L1
            if (r.next >= 100) asm volatile("prefetch.global.L1 [%0];" ::"l"(r.base + r.next - 100));
#endif
            r.next -= 4;
            r.win |= (uint64_t)w << (32 - r.avail);
            r.avail += 32;
        } else {
            while (r.next > 0) {
                uint32_t b = r.base[--r.next];
                r.win |= (uint64_t)b << (56 - r.avail);
                r.avail += 8;
            }
            r.avail = 64;  // everything below is the zero fill of reversebitstream.go:23-27,67-75
        }
    }
}

// Positions the reader on a stream of len bytes.  Returns false on empty input.
SZB_HD bool rev_init(RevBits &r, const uint8_t *data, int32_t len) {
    r.base = data;
    r.next = len;
    r.avail = 0;
    r.win = 0;
    r.remaining = (int64_t)len * 8;
    if (len <= 0) {
        r.avail = 64;
        return false;
    }
    // byte loads until (base + next) is 4-byte aligned, then rev_refill can use aligned words
    while (r.next > 0 && ((reinterpret_cast<uintptr_t>(r.base + r.next)) & 3) != 0) {
        uint32_t b = r.base[--r.next];
        r.win |= (uint64_t)b << (56 - r.avail);
        r.avail += 8;
    }
    if (r.next == 0) r.avail = 64;
    rev_refill(r);
    return true;
}

// Reversebitstream.Read(n), n <= 32; the caller keeps avail >= n through rev_refill.
SZB_HD uint32_t rev_read(RevBits &r, uint32_t n) {
    uint32_t v = (uint32_t)((r.win >> 1) >> (63 - n));  // n == 0 -> 0 (reversebitstream.go:18-20)
    r.win <<= n;
    r.avail -= (int32_t)n;
    r.remaining -= n;
    return v;
}

SZB_HD uint32_t rev_peek(const RevBits &r, uint32_t n) { return (uint32_t)((r.win >> 1) >> (63 - n)); }

SZB_HD void rev_skip(RevBits &r, uint32_t n) {
    r.win <<= n;
    r.avail -= (int32_t)n;
    r.remaining -= n;
}

// The "skip padding" prologue shared by huffman.go:227-238, fse.go:313-324 and
// sequences.go:131-143: zero bits then the first 1 bit, at most 8 bits in total.
// Returns false for ErrBadPadding.
SZB_HD bool rev_skip_padding(RevBits &r) {
    uint32_t top = (uint32_t)(r.win >> 56);
    if (r.remaining <= 0 || top == 0) return false;
    uint32_t n = 1;
    while (!(top & 0x80)) {
        top <<= 1;
        n++;
    }
    rev_skip(r, n);
    return true;
}

struct FwdBits {
    const uint8_t *p;
    uint32_t nbytes;  // bytes available
    uint32_t pos;     // bit position
    bool eof;
};

SZB_HD void fwd_init(FwdBits &f, const uint8_t *p, uint32_t nbytes) {
    f.p = p;
    f.nbytes = nbytes;
    f.pos = 0;
    f.eof = false;
}

// Bitstream.Read(n), n <= 16
SZB_HD uint32_t fwd_read(FwdBits &f, uint32_t n) {
    if (n == 0) return 0;
    uint32_t byte = f.pos >> 3, sh = f.pos & 7;
    if (((f.pos + n + 7) >> 3) > f.nbytes) {
        f.eof = true;
        f.pos += n;
        return 0;
    }
    uint32_t v = f.p[byte];
    if (sh + n > 8) v |= (uint32_t)f.p[byte + 1] << 8;
    if (sh + n > 16) v |= (uint32_t)f.p[byte + 2] << 16;
    f.pos += n;
    return (v >> sh) & ((1u << n) - 1);
}

// Bitstream.UnwindBit (bitstream.go:20-37)
SZB_HD void fwd_unwind_bit(FwdBits &f) { f.pos--; }

// bytes consumed = ceil(bits / 8), fse.go:120-123
SZB_HD uint32_t fwd_bytes_used(const FwdBits &f) { return (f.pos + 7) >> 3; }

}  // namespace szb
