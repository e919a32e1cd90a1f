// fse.cuh -- FSE table description parse, decode-table construction and the predefined
// distributions.
//
// replaces: fse/fse.go:28-130 ReadTabledescriptionFromBitstream, :136-230 BuildDecodingTable,
// :235-249 BIT_highbit32 (-> clz), :253-301 InitState/PeekSymbol/NextState, :307-390
// DecodeInterleavedFSEStreams, and fse/predefined.go:3-78.
//
// Decode-table cell, packed in 32 bits so LL + ML + OF tables of the maximum accuracy logs
// (9, 9, 8) take 5 KB of shared memory per block in flight:
//   bits  0-15 Baseline        (fse.go:213)
//   bits 16-19 NumberOfBits    (fse.go:212)
//   bits 20-24 NumberOfAdditionalBits (LL/ML: predefined.go:18-20,47-50; OF: the code itself)
//   bits 25-31 the symbol (code); LL/ML base values are looked up from the code off the
//              critical state chain (predefined.go:5-11,36-40)
#pragma once
#include "bits.cuh"
#include "../../include/szb200.h"

namespace szb {

enum FseKind { KIND_LL = 0, KIND_OF = 1, KIND_ML = 2, KIND_HUFW = 3 };

constexpr uint32_t kMaxFseSymbols = 64;
constexpr uint32_t kMaxALLL = 9, kMaxALOF = 8, kMaxALML = 9, kMaxALHufW = 9;

#define SZB_LL_BASE_INIT                                                                                              \
    {0,  1,  2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 18, 20, 22, 24, 28, 32, 40, 48, 64, 128, 256, \
     512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 36, 37, 38, 39, 40, 41, 42, 43, 44, 45, 46, 47, 48, 49, 50,     \
     51, 52, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63}
#define SZB_LL_EXTRA_INIT                                                                                     \
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 3, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 14, \
     15, 16, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}
#define SZB_ML_BASE_INIT                                                                                             \
    {3,  4,  5,  6,  7,  8,  9,  10, 11, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22, 23, 24, 25, 26, 27, 28, 29,      \
     30, 31, 32, 33, 34, 35, 37, 39, 41, 43, 47, 51, 59, 67, 83, 99, 131, 259, 515, 1027, 2051, 4099, 8195, 16387,    \
     32771, 65539, 53, 54, 55, 56, 57, 58, 59, 60, 61, 62, 63}
#define SZB_ML_EXTRA_INIT                                                                                        \
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, \
     2, 3, 3, 4, 4, 5, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}

// Codes beyond the translation arrays stay untranslated with zero extra bits, as in
// fse.go:216-221 ("only do translation if necessary").
#if defined(__CUDACC__)
__device__ __constant__ uint32_t kLLBaseDev[64] = SZB_LL_BASE_INIT;
__device__ __constant__ uint8_t kLLExtraDev[64] = SZB_LL_EXTRA_INIT;
__device__ __constant__ uint32_t kMLBaseDev[64] = SZB_ML_BASE_INIT;
__device__ __constant__ uint8_t kMLExtraDev[64] = SZB_ML_EXTRA_INIT;
#endif
static const uint32_t kLLBaseHost[64] = SZB_LL_BASE_INIT;
static const uint8_t kLLExtraHost[64] = SZB_LL_EXTRA_INIT;
static const uint32_t kMLBaseHost[64] = SZB_ML_BASE_INIT;
static const uint8_t kMLExtraHost[64] = SZB_ML_EXTRA_INIT;

#if defined(__CUDA_ARCH__)
#define SZB_LUT(name) name##Dev
#else
#define SZB_LUT(name) name##Host
#endif

SZB_HD uint32_t ll_base(uint32_t c) { return SZB_LUT(kLLBase)[c & 63]; }
SZB_HD uint32_t ml_base(uint32_t c) { return SZB_LUT(kMLBase)[c & 63]; }
SZB_HD uint32_t extra_bits_for(int kind, uint32_t c) {
    if (kind == KIND_LL) return SZB_LUT(kLLExtra)[c & 63];
    if (kind == KIND_ML) return SZB_LUT(kMLExtra)[c & 63];
    if (kind == KIND_OF) return c;
    return 0;
}

// predefined.go:12-16, :42-45, :66-68
static const int8_t kLLDefaultNorm[36] = {4, 3, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 1, 1, 1, 2, 2,
                                          2, 2, 2, 2, 2, 2, 2, 3, 2, 1, 1, 1, 1, 1, -1, -1, -1, -1};
static const int8_t kMLDefaultNorm[53] = {1, 4, 3, 2, 2, 2, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1,
                                          1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1, -1, -1};
static const int8_t kOFDefaultNorm[29] = {1, 1, 1, 1, 1, 1, 2, 2, 2, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1, -1, -1, -1, -1, -1};

SZB_HD uint32_t fse_pack(uint32_t baseline, uint32_t nb, uint32_t extra, uint32_t code) {
    return (baseline & 0xFFFF) | (nb << 16) | (extra << 20) | (code << 25);
}
SZB_HD uint32_t fse_baseline(uint32_t e) { return e & 0xFFFF; }
SZB_HD uint32_t fse_nb(uint32_t e) { return (e >> 16) & 0xF; }
SZB_HD uint32_t fse_extra(uint32_t e) { return (e >> 20) & 0x1F; }
SZB_HD uint32_t fse_code(uint32_t e) { return e >> 25; }

// BIT_highbit32 (fse.go:235-249); the De Bruijn variant returns 0 for 0
SZB_HD uint32_t highbit32(uint32_t v) {
#if defined(__CUDA_ARCH__)
    return v ? 31 - __clz(v) : 0;
#else
    return v ? 31 - (uint32_t)__builtin_clz(v) : 0;
#endif
}

// ReadTabledescriptionFromBitstream (fse.go:28-130).  norm[s] receives the probability
// (-1 = "less than one"), i.e. Values[s]-1.  Serial: run by one lane.
SZB_HD int fse_read_description(const uint8_t *p, uint32_t avail, uint32_t max_al, int16_t *norm, uint32_t *nsym_out,
                                uint32_t *al_out, uint32_t *bytes_out) {
    FwdBits f;
    fwd_init(f, p, avail);
    uint32_t al = fwd_read(f, 4) + 5;  // fse.go:32-37
    if (f.eof) return SZB_ERR_UNEXPECTED_EOF;
    if (al > max_al) return SZB_ERR_UNSUPPORTED;
    int32_t remaining = 1 << al;
    uint32_t cur = 0;
    while (remaining > 0) {  // fse.go:45
        uint32_t nb = highbit32((uint32_t)remaining + 1) + 1;
        uint32_t v = fwd_read(f, nb);
        if (f.eof) return SZB_ERR_UNEXPECTED_EOF;
        uint32_t lower = (1u << (nb - 1)) - 1;                        // fse.go:63
        uint32_t thresh = (1u << nb) - 1 - ((uint32_t)remaining + 1);  // fse.go:64
        if ((v & lower) < thresh) {                                    // fse.go:66-77
            fwd_unwind_bit(f);
            v &= lower;
        } else if (v > lower) {  // fse.go:79-81
            v -= thresh;
        }
        if (cur >= kMaxFseSymbols) return SZB_ERR_UNSUPPORTED;
        int32_t prob = (int32_t)v - 1;
        norm[cur++] = (int16_t)prob;
        remaining -= (prob == -1) ? 1 : prob;  // fse.go:89-93
        if (prob == 0) {                        // fse.go:96-117
            uint32_t skip = 3;
            while (skip == 3) {
                skip = fwd_read(f, 2);
                if (f.eof) return SZB_ERR_UNEXPECTED_EOF;
                for (uint32_t i = 0; i < skip; i++) {
                    if (cur >= kMaxFseSymbols) return SZB_ERR_UNSUPPORTED;
                    norm[cur++] = 0;
                }
            }
        }
    }
    *nsym_out = cur;
    *al_out = al;
    *bytes_out = fwd_bytes_used(f);  // fse.go:120-123
    if (remaining != 0) return SZB_ERR_DIDNT_READ_ALL_PROBABILITIES;
    return SZB_OK;
}

// BuildDecodingTable (fse.go:136-230), serial form (one lane / host).  next is scratch for
// kMaxFseSymbols counters.
SZB_HD int fse_build_serial(const int16_t *norm, uint32_t nsym, uint32_t al, int kind, uint32_t *table, uint16_t *next) {
    const uint32_t size = 1u << al, mask = size - 1;
    const uint32_t step = (size >> 1) + (size >> 3) + 3;
    int32_t high = (int32_t)size - 1;
    for (uint32_t s = 0; s < nsym; s++) {  // fse.go:146-155
        if (norm[s] == -1) {
            if (high < 0) return SZB_ERR_PANIC;
            table[high--] = s;
            next[s] = 1;
        } else {
            next[s] = (uint16_t)norm[s];
        }
    }
    uint32_t pos = 0;
    for (uint32_t s = 0; s < nsym; s++) {  // fse.go:160-184
        for (int32_t i = 0; i < norm[s]; i++) {
            table[pos] = s;
            pos = (pos + step) & mask;
            while ((int32_t)pos > high) pos = (pos + step) & mask;
        }
    }
    if (pos != 0) return SZB_ERR_PANIC;  // fse.go:186-189
    for (uint32_t i = 0; i < size; i++) {  // fse.go:192-227
        uint32_t s = table[i];
        uint32_t n = next[s]++;
        uint32_t nb = al - highbit32(n);
        uint32_t baseline = (n << nb) - size;
        table[i] = fse_pack(baseline, nb, extra_bits_for(kind, s), s);
    }
    return SZB_OK;
}

#if defined(__CUDACC__)
// BuildDecodingTable, warp-cooperative form: same cells as fse_build_serial.
//  pass 1 (spread, fse.go:146-184): the stepping sequence p_j = j*step mod size visits every cell
//    once; the k-th visited cell that is not in the "-1" area belongs to the symbol whose
//    cumulative count covers k.  Lanes take 32 consecutive j, a ballot ranks the valid ones.
//  pass 2 (fse.go:192-227): cells in index order; match_any groups the lanes holding the same
//    symbol so each gets next[s] + (its rank inside the group).
// symk: scratch, size cells (uint8); next: scratch, kMaxFseSymbols counters.
__device__ __forceinline__ int fse_build_warp(const int16_t *norm, uint32_t nsym, uint32_t al, int kind, uint32_t *table,
                                              uint16_t *next, uint8_t *symk) {
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t size = 1u << al, mask = size - 1;
    const uint32_t step = (size >> 1) + (size >> 3) + 3;
    // lane 0: "-1" symbols to the top, counters, and symbol-of-k expansion
    int32_t high = (int32_t)size - 1;
    if (lane == 0) {
        for (uint32_t s = 0; s < nsym; s++) {
            if (norm[s] == -1) {
                table[high--] = s;
                next[s] = 1;
            } else {
                next[s] = (uint16_t)norm[s];
            }
        }
    }
    high = __shfl_sync(0xFFFFFFFFu, high, 0);
    // symk[k] = symbol of the k-th spread cell: each lane expands its own symbols
    {
        uint32_t start = 0;
        for (uint32_t s0 = 0; s0 < nsym; s0 += 32) {
            uint32_t s = s0 + lane;
            int32_t c = (s < nsym && norm[s] > 0) ? norm[s] : 0;
            uint32_t incl = (uint32_t)c;
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
                if ((int)lane >= d) incl += t;
            }
            uint32_t my = start + incl - (uint32_t)c;
            for (int32_t i = 0; i < c; i++) symk[my + i] = (uint8_t)s;
            start += __shfl_sync(0xFFFFFFFFu, incl, 31);
        }
    }
    __syncwarp();
    uint32_t k_base = 0;
    for (uint32_t j0 = 0; j0 < size; j0 += 32) {
        uint32_t j = j0 + lane;
        uint32_t p = (j * step) & mask;
        bool valid = (j < size) && ((int32_t)p <= high);
        uint32_t bal = __ballot_sync(0xFFFFFFFFu, valid);
        if (valid) table[p] = symk[k_base + __popc(bal & ((1u << lane) - 1))];
        k_base += __popc(bal);
    }
    __syncwarp();
    for (uint32_t i0 = 0; i0 < size; i0 += 32) {
        uint32_t i = i0 + lane;
        bool act = i < size;
        uint32_t s = act ? table[i] : (0x10000u + lane);
        uint32_t grp = __match_any_sync(0xFFFFFFFFu, s);
        if (act) {
            uint32_t n = (uint32_t)next[s] + __popc(grp & ((1u << lane) - 1));
            uint32_t nb = al - highbit32(n);
            table[i] = fse_pack((n << nb) - size, nb, extra_bits_for(kind, s), s);
        }
        __syncwarp();
        if (act && (grp >> lane) == 1) next[s] = (uint16_t)(next[s] + __popc(grp));  // highest lane of the group
        __syncwarp();
    }
    return SZB_OK;
}
#endif

}  // namespace szb
