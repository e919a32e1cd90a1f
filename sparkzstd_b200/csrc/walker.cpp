// walker.cpp -- host-side header walk: frames and blocks -> descriptor tables.
//
// C++ twin of the Go walker (go/structure, go/decompression): the reference parses these
// headers inside FrameDecompressor (decompression/framedecompressor.go:130-150, :270-303,
// :306-374) with structure/frame.go:23-127, structure/block.go:33-55, and the header halves
// of structure/literals.go:67-204 and structure/sequences.go:228-269.  Nothing here touches
// entropy-coded payload: that is the GPU's job.
#include <algorithm>
#include <cstdlib>
#include <numeric>
#include <queue>
#include <thread>
#include <cstring>
#include <new>
#include <vector>

#include "../../include/szb200.h"
#include "abi_layout.h"

struct szb_walk {
    std::vector<szb_frame_desc> frames;
    std::vector<szb_block_desc> blocks;
    uint64_t literal_bytes = 0;
    uint64_t sequences = 0;
    // a dictionary's table-only row (szb_walk_create_dict): what a frame's first block carries over from "the previous block"
    uint32_t dict_huf = SZB_NONE, dict_seq = SZB_NONE;
    bool dict = false;
    uint32_t dict_id = 0;
};

namespace {

constexpr uint32_t kMaxBlock = 128 * 1024;

struct Cursor {
    const uint8_t *p;
    uint64_t len;
    uint64_t pos;
    bool need(uint64_t n) const { return len - pos >= n; }
};

// structure/literals.go:67-204 (DecodeType, BytesNeededToDecodeSizes, DecodeSizes)
int parse_literals_header(const uint8_t *b, uint64_t avail, szb_block_desc &d) {
    if (avail < 1) return SZB_ERR_UNEXPECTED_EOF;
    uint8_t b0 = b[0];
    d.lit_type = b0 & 3;
    int sf = (b0 >> 2) & 3;
    int need;
    if (d.lit_type <= 1)
        need = sf == 1 ? 2 : (sf == 3 ? 3 : 1);
    else
        need = sf == 2 ? 4 : (sf == 3 ? 5 : 3);
    if (avail < (uint64_t)need) return SZB_ERR_UNEXPECTED_EOF;
    d.lit_hdr_bytes = (uint8_t)need;
    if (d.lit_type <= 1) {
        d.lit_streams = 1;
        uint32_t regen;
        if (sf == 0 || sf == 2)
            regen = b0 >> 3;
        else if (sf == 1)
            regen = (b0 >> 4) + ((uint32_t)b[1] << 4);
        else
            regen = (b0 >> 4) + ((uint32_t)b[1] << 4) + ((uint32_t)b[2] << 12);
        d.lit_regen = regen;
        d.lit_comp = d.lit_type == 0 ? regen : 1;
    } else {
        uint32_t v = ((uint32_t)b[0] | ((uint32_t)b[1] << 8) | ((uint32_t)b[2] << 16) |
                      ((uint32_t)(need > 3 ? b[3] : 0) << 24)) >> 4;
        d.lit_streams = 4;
        if (sf == 0 || sf == 1) {
            if (sf == 0) d.lit_streams = 1;
            d.lit_regen = v & 0x3FF;
            d.lit_comp = (v >> 10) & 0x3FF;
        } else if (sf == 2) {
            d.lit_regen = v & 0x3FFF;
            d.lit_comp = (v >> 14) & 0x3FFF;
        } else {
            d.lit_regen = v & 0x3FFFF;
            d.lit_comp = ((v >> 18) & 0x3FFFF) + ((uint32_t)b[4] << 10);
        }
    }
    // the reference slices 128 KiB scratch buffers with these sizes and panics beyond (literals.go:283,296)
    if (d.lit_regen > kMaxBlock || d.lit_comp > kMaxBlock) return SZB_ERR_PANIC;
    return SZB_OK;
}

// structure/sequences.go:228-269 (count + modes byte)
int parse_sequences_header(const uint8_t *b, uint64_t avail, szb_block_desc &d) {
    if (avail < 1) return SZB_ERR_UNEXPECTED_EOF;
    uint8_t b0 = b[0];
    uint32_t need = b0 < 128 ? 1 : (b0 < 255 ? 2 : 3);
    if (avail < need) return SZB_ERR_UNEXPECTED_EOF;
    if (b0 < 128)
        d.nseq = b0;
    else if (b0 < 255)
        d.nseq = ((uint32_t)(b0 - 128) << 8) + b[1];
    else
        d.nseq = (uint32_t)b[1] + ((uint32_t)b[2] << 8) + 0x7F00;
    if (b0 == 0) {  // sequences.go:395-400: the section is exactly this one byte
        d.nseq = 0;
        d.seq_hdr_bytes = 1;
        d.seq_modes = 0;
        return SZB_OK;
    }
    if (avail < need + 1) return SZB_ERR_UNEXPECTED_EOF;
    d.seq_modes = b[need];
    d.seq_hdr_bytes = (uint8_t)(need + 1);
    return SZB_OK;
}

struct Carry {
    uint32_t huf = SZB_NONE, ll = SZB_NONE, of = SZB_NONE, ml = SZB_NONE;
};

// One frame starting at src[off].  Appends its rows; returns bytes consumed through *used.
void walk_frame(szb_walk &w, const uint8_t *src, uint64_t off, uint64_t len, uint64_t *used) {
    szb_frame_desc f;
    std::memset(&f, 0, sizeof(f));
    f.src_off = off;
    f.first_block = (uint32_t)w.blocks.size();
    f.content_size = SZB_CONTENT_SIZE_UNKNOWN;
    Cursor c{src + off, len, 0};
    int rc = SZB_OK;
    Carry carry;
    carry.huf = w.dict_huf;  // SZB_NONE without a dictionary (framedecompressor.go:42-52: a frame starts with nothing to carry over)
    carry.ll = carry.of = carry.ml = w.dict_seq;
    uint32_t frame_idx = (uint32_t)w.frames.size();

    do {
        // CheckMagicnum, framedecompressor.go:130-150
        if (!c.need(4)) { rc = SZB_ERR_UNEXPECTED_EOF; break; }
        if (!(c.p[0] == 0x28 && c.p[1] == 0xB5 && c.p[2] == 0x2F && c.p[3] == 0xFD)) { rc = SZB_ERR_WRONG_MAGICNUMBER; break; }
        c.pos = 4;
        // DecodeFrameHeader, framedecompressor.go:306-374
        if (!c.need(1)) { rc = SZB_ERR_UNEXPECTED_EOF; break; }
        uint8_t fhd = c.p[c.pos++];
        f.descriptor = fhd;
        f.single_segment = (fhd >> 5) & 1;   // frame.go:101-103
        f.has_checksum = (fhd >> 2) & 1;     // frame.go:106-108 (the reference never acts on it)
        uint32_t dict_flag = fhd & 3;        // frame.go:113-127
        uint32_t dict_bytes = dict_flag == 3 ? 4 : dict_flag;
        uint32_t fcs_flag = fhd >> 6;        // frame.go:79-98
        uint32_t fcs_bytes = fcs_flag == 0 ? (f.single_segment ? 1 : 0) : (fcs_flag == 1 ? 2 : (fcs_flag == 2 ? 4 : 8));
        uint32_t hdr = (f.single_segment ? 0 : 1) + dict_bytes + fcs_bytes;
        if (!c.need(hdr)) { rc = SZB_ERR_UNEXPECTED_EOF; break; }
        if (!f.single_segment) {             // frame.go:28-36
            uint8_t wd = c.p[c.pos++];
            uint64_t base = 1ull << (10 + (wd >> 3));
            f.window_size = base + (base / 8) * (wd & 7);
        }
        for (uint32_t i = 0; i < dict_bytes; i++) f.dictionary_id |= (uint64_t)c.p[c.pos + i] << (8 * i);
        c.pos += dict_bytes;
        if (fcs_bytes) {                     // frame.go:49-61
            uint64_t v = 0;
            for (uint32_t i = 0; i < fcs_bytes; i++) v |= (uint64_t)c.p[c.pos + i] << (8 * i);
            if (fcs_bytes == 2) v += 256;
            c.pos += fcs_bytes;
            f.content_size = v;
            f.has_content_size = 1;
            if (f.single_segment) f.window_size = v;  // framedecompressor.go:358-360
        }
        // decodeAllBlocks, framedecompressor.go:246-267
        bool last = false;
        while (!last) {
            if (!c.need(3)) { rc = SZB_ERR_UNEXPECTED_EOF; break; }
            const uint8_t *h = c.p + c.pos;
            c.pos += 3;
            szb_block_desc d;
            std::memset(&d, 0, sizeof(d));
            d.frame = frame_idx;
            d.huf_origin = d.ll_origin = d.of_origin = d.ml_origin = SZB_NONE;
            last = (h[0] & 1) != 0;                     // block.go:38
            d.last = last;
            d.type = (h[0] >> 1) & 3;                   // block.go:39
            d.block_size = (uint32_t)(h[0] >> 3) + ((uint32_t)h[1] << 5) + ((uint32_t)h[2] << 13);
            if (d.type >= 3) { rc = SZB_ERR_ILLEGAL_BLOCK_TYPE; break; }      // block.go:45-47
            if (d.block_size > kMaxBlock) { rc = SZB_ERR_ILLEGAL_BLOCK_SIZE; break; }  // block.go:50-52
            d.src_off = off + c.pos;
            uint32_t self = (uint32_t)w.blocks.size();
            if (d.type == 0) {                          // Raw, framedecompressor.go:211-215
                if (!c.need(d.block_size)) { rc = SZB_ERR_UNEXPECTED_EOF; break; }
                c.pos += d.block_size;
            } else if (d.type == 1) {                   // RLE, framedecompressor.go:229-241
                if (!c.need(1)) { rc = SZB_ERR_UNEXPECTED_EOF; break; }
                c.pos += 1;
            } else {                                    // Compressed
                if (!c.need(d.block_size)) { rc = SZB_ERR_UNEXPECTED_EOF; break; }
                const uint8_t *bp = c.p + c.pos;
                rc = parse_literals_header(bp, d.block_size, d);
                if (rc) break;
                if (d.lit_type == 3) {                  // literals.go:247-252
                    if (carry.huf == SZB_NONE) { rc = SZB_ERR_NO_HUFF_TABLE_TO_CARRY_OVER; break; }
                    d.huf_origin = carry.huf;
                } else if (d.lit_type == 2) {
                    d.huf_origin = self;
                }
                uint64_t lit_total = (uint64_t)d.lit_hdr_bytes + d.lit_comp;
                if (lit_total > d.block_size) { rc = SZB_ERR_UNEXPECTED_EOF; break; }
                d.seq_off = (uint32_t)lit_total;
                // From here on the reference has already decoded the block's literals (DecodeNextBlockContent,
                // framedecompressor.go:93-126: literals section first, then the sequences header): an error in this header
                // does not hide an error in the literals.  The block stays in the table -- without sequences, with the
                // error in hdr_status -- so that the device decodes its literals and the first error wins; the walk ends.
                int hs = parse_sequences_header(bp + lit_total, d.block_size - lit_total, d);
                if (!hs && d.nseq == 0 && lit_total + 1 != d.block_size) hs = SZB_ERR_CORRUPT_SIZES;  // framedecompressor.go:114-123
                if (!hs && d.nseq > 0) {                // DecodeTables order LL, OF, ML (sequences.go:275-369)
                    uint32_t llm = d.seq_modes >> 6, ofm = (d.seq_modes >> 4) & 3, mlm = (d.seq_modes >> 2) & 3;
                    if (llm == 3) { if (carry.ll == SZB_NONE) hs = SZB_ERR_NO_LL_TABLE_TO_CARRY_OVER; else d.ll_origin = carry.ll; } else d.ll_origin = self;
                    if (!hs) { if (ofm == 3) { if (carry.of == SZB_NONE) hs = SZB_ERR_NO_OF_TABLE_TO_CARRY_OVER; else d.of_origin = carry.of; } else d.of_origin = self; }
                    if (!hs) { if (mlm == 3) { if (carry.ml == SZB_NONE) hs = SZB_ERR_NO_ML_TABLE_TO_CARRY_OVER; else d.ml_origin = carry.ml; } else d.ml_origin = self; }
                }
                if (hs) {
                    d.hdr_status = hs;
                    d.nseq = 0;
                    d.seq_hdr_bytes = 0;
                    d.seq_modes = 0;
                    d.ll_origin = d.of_origin = d.ml_origin = SZB_NONE;
                    if (d.lit_type >= 2) {
                        d.lit_buf_off = w.literal_bytes;
                        w.literal_bytes += ((uint64_t)d.lit_regen + 15) & ~15ull;
                    }
                    d.seq_buf_off = w.sequences;
                    w.blocks.push_back(d);
                    rc = hs;
                    break;
                }
                // carry rules, framedecompressor.go:283-294
                if (d.lit_type >= 2) carry.huf = d.huf_origin;
                if (d.nseq > 0) { carry.ll = d.ll_origin; carry.of = d.of_origin; carry.ml = d.ml_origin; }
                if (d.lit_type >= 2) {
                    d.lit_buf_off = w.literal_bytes;
                    w.literal_bytes += ((uint64_t)d.lit_regen + 15) & ~15ull;
                }
                d.seq_buf_off = w.sequences;
                w.sequences += ((uint64_t)d.nseq + 31) & ~31ull;
                c.pos += d.block_size;
            }
            w.blocks.push_back(d);
        }
        if (rc) break;
        if (f.has_checksum && c.need(4)) {
            f.checksum = (uint32_t)c.p[c.pos] | ((uint32_t)c.p[c.pos + 1] << 8) | ((uint32_t)c.p[c.pos + 2] << 16) |
                         ((uint32_t)c.p[c.pos + 3] << 24);
            f.checksum_valid = 1;
        }
    } while (false);

    // with a dictionary: a frame that names ANOTHER dictionary is not decoded with this one (RFC 8878 3.1.1.1.3)
    if (rc == SZB_OK && w.dict && f.dictionary_id != 0 && f.dictionary_id != w.dict_id) rc = SZB_ERR_WRONG_DICTIONARY;
    f.status = rc;
    f.src_len = c.pos;
    f.nblocks = (uint32_t)w.blocks.size() - f.first_block;
    w.frames.push_back(f);
    if (used) *used = c.pos + ((rc == SZB_OK && f.has_checksum && c.need(4)) ? 4 : 0);
}

}  // namespace

extern "C" {

uint32_t szb_abi_layout(uint32_t *out, uint32_t cap) { return szb_abi_layout_impl(out, out ? cap : 0); }

// Frames share no state (framedecompressor.go:42-61): multi-GPU decode is a partition of the frame list.  Greedy
// longest-processing-time binning: frames by descending weight (ties: lower index first), each to the least loaded shard
// (ties: lower shard first).  Deterministic, so every rank computes the same partition from the same weights.
int szb_shard_frames(const uint64_t *weight, uint32_t nframes, uint32_t nshards, uint32_t *shard_of, uint64_t *shard_load) {
    if (nshards == 0 || (nframes && (!weight || !shard_of))) return SZB_ERR_INVALID_ARGUMENT;
    std::vector<uint64_t> load(nshards, 0);
    if (nframes) {
        std::vector<uint32_t> order(nframes);
        std::iota(order.begin(), order.end(), 0u);
        std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return weight[a] > weight[b]; });
        typedef std::pair<uint64_t, uint32_t> Slot;  // (load, shard): the smallest pair is the least loaded, lowest shard
        std::priority_queue<Slot, std::vector<Slot>, std::greater<Slot>> heap;
        for (uint32_t r = 0; r < nshards; r++) heap.push(Slot(0, r));
        for (uint32_t i : order) {
            Slot s = heap.top();
            heap.pop();
            shard_of[i] = s.second;
            s.first += weight[i];
            load[s.second] = s.first;
            heap.push(s);
        }
    }
    if (shard_load)
        for (uint32_t r = 0; r < nshards; r++) shard_load[r] = load[r];
    return SZB_OK;
}

static int walk_create_impl(const uint8_t *src, size_t src_len, const uint64_t *frame_off, const uint64_t *frame_len,
                            uint32_t nframes, const szb_block_desc *dict_block, uint32_t dict_id, bool dict, szb_walk **out);

int szb_walk_create(const uint8_t *src, size_t src_len, const uint64_t *frame_off, const uint64_t *frame_len,
                    uint32_t nframes, szb_walk **out) {
    return walk_create_impl(src, src_len, frame_off, frame_len, nframes, nullptr, 0, false, out);
}

// The walk for a batch that is decoded with a dictionary (szb200.h, dictionaries).  dict_block (may be NULL: raw-content
// dictionary) becomes row 0 of the block table: the origin of Treeless literals and Repeat modes in every frame's first
// blocks.  It belongs to a pseudo frame appended AFTER the caller's frames (no blocks of its own, status
// SZB_ERR_INVALID_ARGUMENT): szb_walk_nframes counts it.
int szb_walk_create_dict(const uint8_t *src, size_t src_len, const uint64_t *frame_off, const uint64_t *frame_len,
                         uint32_t nframes, const szb_block_desc *dict_block, uint32_t dict_id, szb_walk **out) {
    return walk_create_impl(src, src_len, frame_off, frame_len, nframes, dict_block, dict_id, true, out);
}

static int walk_create_impl(const uint8_t *src, size_t src_len, const uint64_t *frame_off, const uint64_t *frame_len,
                            uint32_t nframes, const szb_block_desc *dict_block, uint32_t dict_id, bool dict, szb_walk **out) {
    if (!out || (!src && src_len)) return SZB_ERR_INVALID_ARGUMENT;
    szb_walk *w = new (std::nothrow) szb_walk();
    if (!w) return SZB_ERR_NOMEM;
    try {
        w->dict = dict;
        w->dict_id = dict_id;
        if (dict_block) {
            szb_block_desc d = *dict_block;
            d.flags |= SZB_BLOCK_TABLES_ONLY;
            d.lit_buf_off = 0;
            d.seq_buf_off = 0;
            w->literal_bytes = ((uint64_t)d.lit_regen + 15) & ~15ull;
            w->sequences = ((uint64_t)d.nseq + 31) & ~31ull;
            if (d.lit_type == 2) w->dict_huf = 0;
            if (d.nseq) w->dict_seq = 0;
            w->blocks.push_back(d);
        }
        if (frame_off) {
            for (uint32_t i = 0; i < nframes; i++) {
                const uint64_t off = frame_off[i];
                const uint64_t len = frame_len ? frame_len[i] : (off <= src_len ? src_len - off : 0);
                if (off > src_len || len > src_len - off) {
                    delete w;
                    return SZB_ERR_INVALID_ARGUMENT;
                }
            }
            // Frames with given extents are independent: with many of them the walk runs on several threads, each over a
            // contiguous range of frames into tables of its own, which are appended in frame order with their block indices
            // (a frame's first block, the origins of Treeless / Repeat tables) and scratch offsets moved to their final place.
            // The result is the table a single thread makes.  SZB_WALK_THREADS caps the threads (one process per GPU).
            unsigned nthreads = std::thread::hardware_concurrency();
            nthreads = nthreads < 2 ? 1 : (nthreads > 8 ? 8 : nthreads);
            if (const char *wt = getenv("SZB_WALK_THREADS")) {
                const unsigned cap = (unsigned)strtoul(wt, nullptr, 10);
                if (cap >= 1 && cap < nthreads) nthreads = cap;
            }
            if (nframes < 4096) nthreads = 1;
            auto walk_range = [&](szb_walk &part, uint32_t lo, uint32_t hi) {
                part.frames.reserve(hi - lo);
                part.blocks.reserve(hi - lo);
                for (uint32_t i = lo; i < hi; i++) {
                    const uint64_t off = frame_off[i];
                    walk_frame(part, src, off, frame_len ? frame_len[i] : src_len - off, nullptr);
                }
            };
            if (nthreads == 1) {
                w->frames.reserve(w->frames.size() + nframes);
                w->blocks.reserve(w->blocks.size() + nframes);
                walk_range(*w, 0, nframes);
            } else {
                std::vector<szb_walk> parts(nthreads);
                std::vector<std::thread> th;
                std::vector<int> failed(nthreads, 0);
                const uint32_t per = (nframes + nthreads - 1) / nthreads;
                for (unsigned t = 0; t < nthreads; t++) {
                    szb_walk &part = parts[t];
                    part.dict = w->dict;
                    part.dict_id = w->dict_id;
                    // what every frame starts with is the same in every part: the dictionary's row of the FINAL table.  Bit 31
                    // marks such an origin as absolute; the part's own origins are relative to its first block (below)
                    part.dict_huf = w->dict_huf == SZB_NONE ? SZB_NONE : (w->dict_huf | 0x80000000u);
                    part.dict_seq = w->dict_seq == SZB_NONE ? SZB_NONE : (w->dict_seq | 0x80000000u);
                    const uint32_t lo = t * per < nframes ? t * per : nframes, hi = lo + per < nframes ? lo + per : nframes;
                    auto job = [&, t, lo, hi]() {
                        try {
                            walk_range(parts[t], lo, hi);
                        } catch (...) {
                            failed[t] = 1;
                        }
                    };
                    bool started = false;
                    if (t + 1 < nthreads) {
                        try {
                            th.emplace_back(job);
                            started = true;
                        } catch (...) {
                        }
                    }
                    if (!started) job();  // the last range, or no thread to be had: right here
                }
                for (auto &x : th) x.join();
                for (unsigned t = 0; t < nthreads; t++)
                    if (failed[t]) throw std::bad_alloc();
                for (unsigned t = 0; t < nthreads; t++) {
                    const szb_walk &part = parts[t];
                    const uint32_t fbase = (uint32_t)w->frames.size(), bbase = (uint32_t)w->blocks.size();
                    const uint64_t lbase = w->literal_bytes, sbase = w->sequences;
                    for (szb_frame_desc f : part.frames) {
                        f.first_block += bbase;
                        w->frames.push_back(f);
                    }
                    for (szb_block_desc d : part.blocks) {
                        d.frame += fbase;
                        w->blocks.push_back(d);
                    }
                    for (size_t k = bbase; k < w->blocks.size(); k++) {
                        szb_block_desc &d = w->blocks[k];
                        auto move = [&](uint32_t &o) {
                            if (o == SZB_NONE) return;
                            if (o & 0x80000000u)
                                o &= 0x7FFFFFFFu;  // the dictionary's row: an absolute index
                            else
                                o += bbase;
                        };
                        move(d.huf_origin);
                        move(d.ll_origin);
                        move(d.of_origin);
                        move(d.ml_origin);
                        if (d.type == 2 && d.lit_type >= 2) d.lit_buf_off += lbase;
                        if (d.type == 2) d.seq_buf_off += sbase;
                    }
                    w->literal_bytes += part.literal_bytes;
                    w->sequences += part.sequences;
                }
            }
        } else {
            // concatenated frames: discover boundaries (not a reference behaviour)
            uint64_t pos = 0;
            while (pos < src_len) {
                if (src_len - pos >= 8 && (src[pos] & 0xF0) == 0x50 && src[pos + 1] == 0x2A && src[pos + 2] == 0x4D &&
                    src[pos + 3] == 0x18) {  // skippable frame 0x184D2A5?
                    uint64_t sz = (uint64_t)src[pos + 4] | ((uint64_t)src[pos + 5] << 8) | ((uint64_t)src[pos + 6] << 16) |
                                  ((uint64_t)src[pos + 7] << 24);
                    if (src_len - pos - 8 < sz) break;
                    pos += 8 + sz;
                    continue;
                }
                uint64_t used = 0;
                walk_frame(*w, src, pos, src_len - pos, &used);
                if (w->frames.back().status != SZB_OK || used == 0) break;
                pos += used;
            }
        }
        if (dict_block) {  // the pseudo frame the table-only row belongs to: after the caller's frames, no blocks to execute
            szb_frame_desc f;
            std::memset(&f, 0, sizeof(f));
            f.content_size = SZB_CONTENT_SIZE_UNKNOWN;
            f.status = SZB_ERR_INVALID_ARGUMENT;
            w->blocks[0].frame = (uint32_t)w->frames.size();
            w->frames.push_back(f);
        }
    } catch (const std::bad_alloc &) {
        delete w;
        return SZB_ERR_NOMEM;
    }
    *out = w;
    return SZB_OK;
}

void szb_walk_destroy(szb_walk *w) { delete w; }
uint32_t szb_walk_nframes(const szb_walk *w) { return (uint32_t)w->frames.size(); }
uint32_t szb_walk_nblocks(const szb_walk *w) { return (uint32_t)w->blocks.size(); }
const szb_frame_desc *szb_walk_frames(const szb_walk *w) { return w->frames.data(); }
const szb_block_desc *szb_walk_blocks(const szb_walk *w) { return w->blocks.data(); }
uint64_t szb_walk_literal_bytes(const szb_walk *w) { return w->literal_bytes; }
uint64_t szb_walk_sequences(const szb_walk *w) { return w->sequences; }
uint64_t szb_walk_known_output_size(const szb_walk *w) {
    uint64_t total = 0;
    for (const auto &f : w->frames) {
        if (f.status != SZB_OK) continue;
        if (!f.has_content_size) return SZB_CONTENT_SIZE_UNKNOWN;
        total += f.content_size;
    }
    return total;
}

}  // extern "C"
