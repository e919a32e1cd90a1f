// abi_layout.h -- the byte layout of the two descriptor tables (include/szb200.h), pinned at compile time and
// reported at run time.  Bindings that cannot include the header (Go structs in go/szb200, the ctypes mirror in
// sparkzstd_b200/_lib.py) compare their own sizes and offsets against szb_abi_layout() when they load.
#pragma once
#include <stddef.h>
#include <stdint.h>

#include "../../include/szb200.h"

#define SZB_FRAME_FIELDS(X)                                                                                          \
    X(src_off) X(src_len) X(window_size) X(content_size) X(dictionary_id) X(first_block) X(nblocks) X(checksum) \
    X(status) X(descriptor) X(single_segment) X(has_checksum) X(has_content_size) X(checksum_valid)
#define SZB_BLOCK_FIELDS(X)                                                                                         \
    X(src_off) X(lit_buf_off) X(seq_buf_off) X(block_size) X(frame) X(lit_regen) X(lit_comp) X(nseq) X(seq_off) \
    X(huf_origin) X(ll_origin) X(of_origin) X(ml_origin) X(type) X(last) X(lit_type) X(lit_streams)            \
    X(lit_hdr_bytes) X(seq_hdr_bytes) X(seq_modes) X(flags) X(hdr_status)

// what every binding assumes (go/szb200/szb200.go, sparkzstd_b200/_lib.py)
static_assert(sizeof(szb_frame_desc) == 64, "szb_frame_desc is 64 bytes");
static_assert(offsetof(szb_frame_desc, src_off) == 0 && offsetof(szb_frame_desc, src_len) == 8 &&
                  offsetof(szb_frame_desc, window_size) == 16 && offsetof(szb_frame_desc, content_size) == 24 &&
                  offsetof(szb_frame_desc, dictionary_id) == 32 && offsetof(szb_frame_desc, first_block) == 40 &&
                  offsetof(szb_frame_desc, nblocks) == 44 && offsetof(szb_frame_desc, checksum) == 48 &&
                  offsetof(szb_frame_desc, status) == 52 && offsetof(szb_frame_desc, descriptor) == 56 &&
                  offsetof(szb_frame_desc, single_segment) == 57 && offsetof(szb_frame_desc, has_checksum) == 58 &&
                  offsetof(szb_frame_desc, has_content_size) == 59 && offsetof(szb_frame_desc, checksum_valid) == 60,
              "szb_frame_desc field offsets");
static_assert(sizeof(szb_block_desc) == 80, "szb_block_desc is 80 bytes");
static_assert(offsetof(szb_block_desc, src_off) == 0 && offsetof(szb_block_desc, lit_buf_off) == 8 &&
                  offsetof(szb_block_desc, seq_buf_off) == 16 && offsetof(szb_block_desc, block_size) == 24 &&
                  offsetof(szb_block_desc, frame) == 28 && offsetof(szb_block_desc, lit_regen) == 32 &&
                  offsetof(szb_block_desc, lit_comp) == 36 && offsetof(szb_block_desc, nseq) == 40 &&
                  offsetof(szb_block_desc, seq_off) == 44 && offsetof(szb_block_desc, huf_origin) == 48 &&
                  offsetof(szb_block_desc, ll_origin) == 52 && offsetof(szb_block_desc, of_origin) == 56 &&
                  offsetof(szb_block_desc, ml_origin) == 60 && offsetof(szb_block_desc, type) == 64 &&
                  offsetof(szb_block_desc, last) == 65 && offsetof(szb_block_desc, lit_type) == 66 &&
                  offsetof(szb_block_desc, lit_streams) == 67 && offsetof(szb_block_desc, lit_hdr_bytes) == 68 &&
                  offsetof(szb_block_desc, seq_hdr_bytes) == 69 && offsetof(szb_block_desc, seq_modes) == 70 &&
                  offsetof(szb_block_desc, flags) == 71 && offsetof(szb_block_desc, hdr_status) == 72,
              "szb_block_desc field offsets");

static inline uint32_t szb_abi_layout_impl(uint32_t *out, uint32_t cap) {
    const uint32_t v[] = {
        (uint32_t)sizeof(szb_frame_desc),
#define X(f) (uint32_t)offsetof(szb_frame_desc, f),
        SZB_FRAME_FIELDS(X)
#undef X
        (uint32_t)sizeof(szb_block_desc),
#define X(f) (uint32_t)offsetof(szb_block_desc, f),
        SZB_BLOCK_FIELDS(X)
#undef X
    };
    const uint32_t n = (uint32_t)(sizeof(v) / sizeof(v[0]));
    for (uint32_t i = 0; i < n && i < cap; i++) out[i] = v[i];
    return n;
}
