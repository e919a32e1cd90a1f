// hostsim.cpp -- TEST INFRASTRUCTURE: runs the device code of the CUDA path on the CPU, driven by the product's own header
// walker, so that `pytest -m "not gpu"` can check it against the oracle without a GPU:
//   * stages 1-3: the *serial* device functions (bits.cuh, fse.cuh, huffman.cuh, sequences.cuh -- the code a single lane
//     executes).  Their warp-cooperative wrappers (table builds with ballots, cp.async bit rings) only run on the device and
//     are covered by the -m gpu tests;
//   * stage 4: the kernels themselves (execute.cuh, execute_long.cuh), thread by thread on an emulated CTA (warpsim.h).
// This file is never linked into libszb200.so.
#include <cstdint>
#include <cstring>
#include <vector>

#include <algorithm>
#include <random>

#include "warpsim.h"

#include "../../include/szb200.h"
#include "../../sparkzstd_b200/csrc/batch.cuh"
#include "../../sparkzstd_b200/csrc/long_tables.h"
#include "../../sparkzstd_b200/csrc/huffman.cuh"
#include "../../sparkzstd_b200/csrc/sequences.cuh"

namespace szb {
#include "../../sparkzstd_b200/csrc/execute.cuh"  // stage 4, execute_long.cuh included
}

using namespace szb;

namespace {

struct Tables {
    uint32_t tll[512], tof[256], tml[512];
    uint32_t al[3];
};

int build_one(const TableSource &ts, int kind, uint32_t *table, uint32_t *al, uint32_t *used) {
    int16_t norm[64];
    uint16_t next[64];
    if (ts.mode == 0) {
        uint32_t n = kind == KIND_LL ? 36 : (kind == KIND_OF ? 29 : 53);
        const int8_t *src = kind == KIND_LL ? kLLDefaultNorm : (kind == KIND_OF ? kOFDefaultNorm : kMLDefaultNorm);
        for (uint32_t i = 0; i < n; i++) norm[i] = src[i];
        *al = kind == KIND_OF ? 5 : 6;
        *used = 0;
        return fse_build_serial(norm, n, *al, kind, table, next);
    }
    if (ts.mode == 1) {
        if (ts.avail < 1) return SZB_ERR_UNEXPECTED_EOF;
        uint32_t code = ts.p[0];
        if ((kind == KIND_LL && code >= 36) || (kind == KIND_ML && code >= 53)) return SZB_ERR_PANIC;
        if (kind == KIND_OF && code > 31) return SZB_ERR_UNSUPPORTED;
        table[0] = fse_pack(0, 0, extra_bits_for(kind, code), code);
        *al = 0;
        *used = 1;
        return SZB_OK;
    }
    uint32_t nsym;
    uint32_t max_al = kind == KIND_OF ? kMaxALOF : (kind == KIND_LL ? kMaxALLL : kMaxALML);
    int rc = fse_read_description(ts.p, ts.avail, max_al, norm, &nsym, al, used);
    if (rc) return rc;
    if (kind == KIND_OF && nsym > 32) return SZB_ERR_UNSUPPORTED;
    return fse_build_serial(norm, nsym, *al, kind, table, next);
}

int huffman_block(const uint8_t *src, const szb_block_desc *blocks, uint32_t b, uint8_t *out) {
    const szb_block_desc &d = blocks[b];
    const szb_block_desc &o = blocks[d.huf_origin];
    const uint8_t *tree = src + o.src_off + o.lit_hdr_bytes;
    uint32_t tree_avail = o.lit_comp;
    if (tree_avail < 1) return SZB_ERR_UNEXPECTED_EOF;
    uint32_t hb = tree[0], nw = 0, tree_bytes;
    uint8_t weights[256];
    if (hb < 128) {
        if (1 + hb > tree_avail) return SZB_ERR_UNEXPECTED_EOF;
        int16_t norm[64];
        uint16_t next[64];
        uint32_t table[512];
        uint32_t nsym, al, used;
        int rc = fse_read_description(tree + 1, tree_avail - 1, kMaxALHufW, norm, &nsym, &al, &used);
        if (rc) return rc;
        rc = fse_build_serial(norm, nsym, al, KIND_HUFW, table, next);
        if (rc) return rc;
        if (used > hb) return SZB_ERR_PANIC;
        rc = fse_decode_weights(table, al, tree + 1 + used, hb - used, weights, &nw);
        if (rc) return rc;
        tree_bytes = 1 + hb;
    } else {
        nw = hb - 127;
        int rc = huf_read_direct_weights(tree + 1, tree_avail - 1, nw, weights);
        if (rc) return rc;
        tree_bytes = 1 + ((nw + 1) >> 1);
    }
    std::vector<uint16_t> huf(1 << kMaxHufBits);
    uint32_t max_bits;
    int rc = huf_build_serial(weights, nw, huf.data(), &max_bits);
    if (rc) return rc;
    const uint8_t *payload = src + d.src_off;
    uint32_t skip = d.lit_hdr_bytes + (d.lit_type == 2 ? tree_bytes : 0);
    int32_t comp = (int32_t)d.lit_comp - (int32_t)(d.lit_type == 2 ? tree_bytes : 0);
    if (d.lit_streams == 1) {
        if (comp < 0) return SZB_ERR_PANIC;
        return huf_decode_stream(huf.data(), max_bits, payload + skip, (uint32_t)comp, out, d.lit_regen);
    }
    comp -= 6;
    if (comp < 0) return SZB_ERR_PANIC;
    const uint8_t *jt = payload + skip;
    uint32_t s[4] = {(uint32_t)(jt[0] | (jt[1] << 8)), (uint32_t)(jt[2] | (jt[3] << 8)), (uint32_t)(jt[4] | (jt[5] << 8)), 0};
    uint32_t normal = (d.lit_regen + 3) / 4;
    int32_t last = (int32_t)d.lit_regen - 3 * (int32_t)normal;
    if (s[0] + s[1] + s[2] > (uint32_t)comp) return SZB_ERR_CORRUPTED_JUMPTABLE;
    if (last < 0) return SZB_ERR_PANIC;
    s[3] = (uint32_t)comp - (s[0] + s[1] + s[2]);
    uint32_t start = 0;
    for (int k = 0; k < 4; k++) {
        rc = huf_decode_stream(huf.data(), max_bits, jt + 6 + start, s[k], out + k * normal, k < 3 ? normal : (uint32_t)last);
        if (rc) return rc;
        start += s[k];
    }
    return SZB_OK;
}

int sequences_block(const uint8_t *src, const szb_block_desc *blocks, uint32_t b, std::vector<uint32_t> &ll,
                    std::vector<uint32_t> &ml, std::vector<uint32_t> &of) {
    const szb_block_desc &d = blocks[b];
    const uint8_t *tables = src + d.src_off + d.seq_off + d.seq_hdr_bytes;
    uint32_t tables_avail = d.block_size - d.seq_off - d.seq_hdr_bytes;
    Tables t;
    uint32_t cursor = 0;
    int16_t norm[64];
    for (int kind = 0; kind < 3; kind++) {
        uint32_t *table = kind == KIND_LL ? t.tll : (kind == KIND_OF ? t.tof : t.tml);
        uint32_t mode = field_mode(d.seq_modes, kind);
        TableSource ts;
        if (mode == 3) {
            uint32_t ob = kind == KIND_LL ? d.ll_origin : (kind == KIND_OF ? d.of_origin : d.ml_origin);
            const szb_block_desc &o = blocks[ob];
            int rc = locate_field(src + o.src_off + o.seq_off + o.seq_hdr_bytes, o.block_size - o.seq_off - o.seq_hdr_bytes,
                                  o.seq_modes, kind, norm, &ts);
            if (rc) return rc;
        } else {
            ts.p = tables + cursor;
            ts.avail = tables_avail - cursor;
            ts.mode = mode;
        }
        uint32_t used = 0;
        int rc = build_one(ts, kind, table, &t.al[kind], &used);
        if (rc) return rc;
        if (mode != 3) cursor += used;
        if (cursor > tables_avail) return SZB_ERR_UNEXPECTED_EOF;
    }
    RevBits r;
    if (!rev_init(r, tables + cursor, (int32_t)(tables_avail - cursor)) || !rev_skip_padding(r)) return SZB_ERR_BAD_PADDING;
    rev_refill(r);
    SeqStates st;
    st.ll = rev_read(r, t.al[KIND_LL]);
    st.of = rev_read(r, t.al[KIND_OF]);
    st.ml = rev_read(r, t.al[KIND_ML]);
    ll.resize(d.nseq);
    ml.resize(d.nseq);
    of.resize(d.nseq);
    for (uint32_t i = 0; i < d.nseq; i++)
        decode_one_sequence(r, t.tll, t.tof, t.tml, st, i + 1 < d.nseq, &ll[i], &ml[i], &of[i]);
    if (r.remaining != 0) return SZB_ERR_NOT_ALL_BITS_USED;
    return SZB_OK;
}

}  // namespace

extern "C" {
int hostsim_stage4_at(const uint8_t *src, size_t len, uint8_t *out, size_t cap, size_t *out_len, int path, int order, int two_frames,
                      int verify_checksum, int *exec_err);


// Decodes ONE frame with walker + serial device functions + a plain serial executor.
// Optional per-block capture: lit_out (concatenated Huffman-decoded literals, block order).
int hostsim_decode_frame(const uint8_t *src, size_t len, uint8_t *out, size_t cap, size_t *out_len) {
    uint64_t off = 0, flen = len;
    szb_walk *w = nullptr;
    int rc = szb_walk_create(src, len, &off, &flen, 1, &w);
    if (rc) return rc;
    const szb_frame_desc fr = szb_walk_frames(w)[0];
    const szb_block_desc *blocks = szb_walk_blocks(w);
    uint32_t nb = szb_walk_nblocks(w);
    size_t pos = 0;
    uint32_t h[3] = {1, 4, 8};
    std::vector<uint8_t> lit(128 * 1024);
    std::vector<uint32_t> ll, ml, of;
    rc = SZB_OK;
    for (uint32_t b = 0; b < nb && rc == SZB_OK; b++) {
        const szb_block_desc &d = blocks[b];
        const uint8_t *payload = src + d.src_off;
        if (d.type == 0) {
            if (pos + d.block_size > cap) { rc = SZB_ERR_DST_TOO_SMALL; break; }
            memcpy(out + pos, payload, d.block_size);
            pos += d.block_size;
            continue;
        }
        if (d.type == 1) {
            if (pos + d.block_size > cap) { rc = SZB_ERR_DST_TOO_SMALL; break; }
            memset(out + pos, payload[0], d.block_size);
            pos += d.block_size;
            continue;
        }
        const uint8_t *lp = nullptr;
        if (d.lit_type >= 2) {
            rc = huffman_block(src, blocks, b, lit.data());
            if (rc) break;
            lp = lit.data();
        } else if (d.lit_type == 0) {
            lp = payload + d.lit_hdr_bytes;
        }
        uint8_t rle = d.lit_type == 1 ? payload[d.lit_hdr_bytes] : 0;
        ll.clear(); ml.clear(); of.clear();
        if (d.nseq) {
            rc = sequences_block(src, blocks, b, ll, ml, of);
            if (rc) break;
        }
        uint32_t lit_pos = 0;
        for (uint32_t i = 0; i < d.nseq && rc == SZB_OK; i++) {
            if (lit_pos + ll[i] > d.lit_regen) { rc = SZB_ERR_DIDNT_COPY_ALL_LITERAL_BYTES; break; }
            if (pos + ll[i] + ml[i] > cap) { rc = SZB_ERR_DST_TOO_SMALL; break; }
            if (d.lit_type == 1) memset(out + pos, rle, ll[i]); else memcpy(out + pos, lp + lit_pos, ll[i]);
            pos += ll[i];
            lit_pos += ll[i];
            uint32_t o;
            uint32_t v = of[i];
            if (v > 3) { o = v - 3; h[2] = h[1]; h[1] = h[0]; h[0] = o; }
            else {
                uint32_t idx = v - 1 + (ll[i] == 0 ? 1 : 0);
                if (idx == 0) o = h[0];
                else if (idx == 1) { o = h[1]; h[1] = h[0]; h[0] = o; }
                else { o = idx == 2 ? h[2] : h[0] - 1; h[2] = h[1]; h[1] = h[0]; h[0] = o; }
            }
            if (o == 0 || o > pos) { rc = SZB_ERR_CANT_REPEAT_BYTES; break; }
            for (uint32_t k = 0; k < ml[i]; k++) out[pos + k] = out[pos - o + k];
            pos += ml[i];
        }
        if (rc) break;
        uint32_t rest = d.lit_regen - lit_pos;
        if (pos + rest > cap) { rc = SZB_ERR_DST_TOO_SMALL; break; }
        if (d.lit_type == 1) memset(out + pos, rle, rest); else memcpy(out + pos, lp + lit_pos, rest);
        pos += rest;
    }
    if (rc == SZB_OK) rc = fr.status;
    szb_walk_destroy(w);
    *out_len = pos;
    return rc;
}


// Decodes ONE frame with the product's own stage 4 on an emulated CTA (warpsim.h): walker, the serial device functions for
// stages 1-3, then the kernels of execute.cuh exactly as launch_execute (api.cu) queues them -- k_scan_blocks,
// k_frame_verdict, k_execute_bodies and, by `path`,
//   0: k_execute (one warp per frame)   1: k_execute_pair (producer warp + consumer warp)
//   2: the block-parallel kernels of execute_long.cuh, with k_execute_pair launched beside them as the fallback.
//   3: k_resolve + k_place (place.cuh), with k_execute launched behind them as the fallback, as launch_execute does;
//      4: the same with bitmaps too small for the output, so that every frame falls back to k_execute.
//   5: k_execute2 (exec2.cuh), with k_execute launched behind it as launch_execute does.
//      (lines are aligned to memory, not to the output: callers also pass an `out` that is not 128-byte aligned)
//   6: k_execute_pair2 (exec2.cuh: k_execute2's producer in one warp, its consumer in a second), with k_execute_pair behind it.
//   7: k_execute_team (exec2.cuh: the consumer on kX2Team warps that meet at a named barrier), with k_execute_pair behind it.
// order: 0 = CTAs in launch order, 1 = reversed (the worst case for k_long_jump), >= 2 = shuffled with that seed.
// With two_frames the same frame is decoded twice in one batch and both copies must agree.
// verify_checksum: 1 = also run k_verify_checksums; 2 = then flip an output byte and expect the mismatch (returns 3; 2 when the
// frame carries no checksum).
// Returns the frame's status; 1 when path 2 left the frame to its fallback (and the fallback decoded it).
int hostsim_stage4(const uint8_t *src, size_t len, uint8_t *out, size_t cap, size_t *out_len, int path, int order, int two_frames,
                   int verify_checksum) {
    return hostsim_stage4_at(src, len, out, cap, out_len, path, order, two_frames, verify_checksum, nullptr);
}

// The same; exec_err (when given) receives the status the frame ends with even when it is an error (statuses are compared
// with the oracle's codes).
int hostsim_stage4_at(const uint8_t *src, size_t len, uint8_t *out, size_t cap, size_t *out_len, int path, int order, int two_frames,
                      int verify_checksum, int *exec_err) {
    uint64_t off = 0, flen = len;
    szb_walk *w = nullptr;
    *out_len = 0;
    int rc = szb_walk_create(src, len, &off, &flen, 1, &w);
    if (rc) return rc;
    const szb_frame_desc fr0 = szb_walk_frames(w)[0];
    const uint32_t nb1 = szb_walk_nblocks(w);
    if (fr0.status != SZB_OK || fr0.nblocks != nb1) {
        szb_walk_destroy(w);
        return fr0.status ? fr0.status : SZB_ERR_INVALID_ARGUMENT;
    }
    const uint32_t copies = two_frames ? 2 : 1;
    const uint32_t nb = nb1 * copies;
    std::vector<szb_block_desc> blocks(nb);
    std::vector<szb_frame_desc> frames(copies, fr0);
    uint64_t lit_bytes = 0, seqs = 0;
    for (uint32_t b = 0; b < nb1; b++) {
        const szb_block_desc &d = szb_walk_blocks(w)[b];
        if (d.type == 2 && d.lit_type >= 2) lit_bytes = std::max<uint64_t>(lit_bytes, d.lit_buf_off + ((d.lit_regen + 15) & ~15u));
        if (d.type == 2) seqs = std::max<uint64_t>(seqs, d.seq_buf_off + ((d.nseq + 31) & ~31u));
    }
    for (uint32_t c = 0; c < copies; c++) {
        frames[c].first_block = c * nb1;
        for (uint32_t b = 0; b < nb1; b++) {
            szb_block_desc d = szb_walk_blocks(w)[b];
            d.frame = c;
            d.lit_buf_off += c * lit_bytes;
            d.seq_buf_off += c * seqs;
            if (d.huf_origin != SZB_NONE) d.huf_origin += c * nb1;
            if (d.ll_origin != SZB_NONE) d.ll_origin += c * nb1;
            if (d.of_origin != SZB_NONE) d.of_origin += c * nb1;
            if (d.ml_origin != SZB_NONE) d.ml_origin += c * nb1;
            blocks[c * nb1 + b] = d;
        }
    }
    szb_walk_destroy(w);
    // stages 1-3 with the serial device functions
    const uint64_t stride = seqs * copies + 32;
    std::vector<uint8_t> litbuf(lit_bytes * copies + 512, 0xEE);
    std::vector<uint32_t> seq(3 * stride, 0xDEADBEEFu);  // padding lanes hold garbage, as on the device
    std::vector<uint64_t> out_size(nb), out_off(nb, ~0ull);
    std::vector<int32_t> lit_status(nb, SZB_OK), seq_status(nb, SZB_OK);
    std::vector<uint32_t> ll, ml, of, body_list;
    for (uint32_t b = 0; b < nb; b++) {
        const szb_block_desc &d = blocks[b];
        out_size[b] = d.type == 2 ? d.lit_regen : d.block_size;
        if ((d.type != 2 && d.block_size > 0) || (d.type == 2 && d.nseq == 0 && d.lit_regen > 0)) body_list.push_back(b);  // api.cu
        if (d.type != 2) continue;
        if (d.lit_type >= 2) {
            rc = huffman_block(src, blocks.data(), b, litbuf.data() + d.lit_buf_off);
            if (rc) return rc;
        }
        if (d.nseq) {
            rc = sequences_block(src, blocks.data(), b, ll, ml, of);
            if (rc) return rc;
            for (uint32_t i = 0; i < d.nseq; i++) {
                seq[d.seq_buf_off + i] = ll[i];
                seq[stride + d.seq_buf_off + i] = ml[i];
                seq[2 * stride + d.seq_buf_off + i] = of[i];
                out_size[b] += ml[i];
            }
        }
    }
    uint64_t total = 0, dev_total = 0;
    for (uint32_t b = 0; b < nb; b++) total += out_size[b];
    if (total > cap) return SZB_ERR_DST_TOO_SMALL;
    std::vector<uint64_t> frame_out_off(copies, ~0ull), frame_out_len(copies, ~0ull);
    std::vector<int32_t> frame_status(copies, -99);
    std::vector<uint8_t> bytefill(256 * 256);
    for (int v = 0; v < 256; v++) memset(bytefill.data() + 256 * v, v, 256);
    // the long-frame tables, as batch_upload_tables (api.cu) builds them (SZB_LONG_SLICE as there)
    std::vector<uint32_t> exec_list(copies);
    for (uint32_t c = 0; c < copies; c++) exec_list[c] = copies - 1 - c;  // slots need not be in frame order
    const uint32_t slice = getenv("SZB_LONG_SLICE") ? (uint32_t)((strtoul(getenv("SZB_LONG_SLICE"), nullptr, 10) + 31) / 32 * 32) : 0;
    LongTables lt;
    lt.clear();
    if (path == 2) build_long_tables(frames.data(), blocks.data(), exec_list.data(), copies, slice, kJumpTile, lt);
    const std::vector<uint32_t> &lb_block = lt.lb_block, &lb_slot = lt.lb_slot, &long_first_lb = lt.long_first_lb;
    const std::vector<uint64_t> &long_dbase = lt.long_dbase;
    const uint32_t n_lb = (uint32_t)lb_block.size();
    std::vector<uint32_t> dist(long_dbase.back() + 1, 0xCDCDCDCDu), long_hist(3 * (size_t)n_lb + 3);
    std::vector<uint64_t> long_T(3 * (size_t)n_lb + 3), ls_T(3 * lt.ls_lb.size() + 3, 0xABABABABABABABABull), ls_sum(2 * lt.ls_lb.size() + 2);
    std::vector<unsigned long long> long_err(copies, kLongNoError);
    unsigned long long ticket = 0;
    DeviceBatch a{};
    a.src = src;
    a.blocks = blocks.data();
    a.frames = frames.data();
    a.nblocks = nb;
    a.nframes = copies;
    a.litbuf = litbuf.data();
    a.seq_ll = seq.data();
    a.seq_ml = seq.data() + stride;
    a.seq_of = seq.data() + 2 * stride;
    a.seq_stride = stride;
    a.out_size = out_size.data();
    a.out_off = out_off.data();
    a.lit_status = lit_status.data();
    a.seq_status = seq_status.data();
    a.total = &dev_total;
    a.bytefill = bytefill.data();
    a.dst = out;
    a.dst_cap = cap;
    a.frame_out_off = frame_out_off.data();
    a.frame_out_len = frame_out_len.data();
    a.frame_status = frame_status.data();
    std::vector<uint32_t> frame_nexec(copies, 0xDEADu);
    a.frame_nexec = frame_nexec.data();
    a.exec_list = exec_list.data();
    a.body_list = body_list.data();
    a.n_body = (uint32_t)body_list.size();
    a.n_long = (path == 0 || path == 5) ? 0 : copies;
    a.exec2 = path == 5 || path == 6 || path == 7 ? 1 : 0;
    a.pair2 = path == 6 ? 1 : (path == 7 ? 2 : 0);
    a.n_lb = n_lb;
    a.lb_block = lb_block.data();
    a.lb_slot = lb_slot.data();
    a.long_first_lb = long_first_lb.data();
    a.long_dbase = long_dbase.data();
    a.dist = path == 2 ? dist.data() : nullptr;
    a.long_T = long_T.data();
    a.long_hist = long_hist.data();
    a.long_err = long_err.data();
    a.long_ticket = &ticket;
    a.n_ls = (uint32_t)lt.ls_lb.size();
    a.long_slice = slice;
    a.ls_lb = lt.ls_lb.data();
    a.ls_seq0 = lt.ls_seq0.data();
    a.lb_first_ls = lt.lb_first_ls.data();
    a.ls_T = ls_T.data();
    a.ls_sum = ls_sum.data();
    // place.cuh, as batch_upload_tables / make_args (api.cu)
    std::vector<uint64_t> rec_off(nb ? nb : 1, 0);
    uint64_t rec_entries = 0, bm_bound = 0;
    for (uint32_t c = 0; c < copies; c++) {
        uint64_t fbound = 0;
        for (uint32_t i = c * nb1; i < (c + 1) * nb1; i++) {
            const szb_block_desc &d = blocks[i];
            fbound += d.type != 2 ? d.block_size : (d.nseq ? 128u * 1024u : d.lit_regen);
            if (d.type == 2 && d.nseq) {
                rec_off[i] = rec_entries;
                rec_entries += ((2 * (uint64_t)d.nseq + 1 + 3) & ~3ull) + 4;
            }
        }
        if (fr0.has_content_size && fr0.content_size < fbound) fbound = fr0.content_size;
        bm_bound += fbound;
    }
    if (path == 4) bm_bound = total ? total - 1 : 0;
    rec_entries += 64;
    const uint64_t bm_words = ((bm_bound + 256) / 32 + 4ull * copies + 64 + 3) & ~3ull;
    std::vector<uint32_t> place_mem(rec_entries + bm_words, 0xA5A5A5A5u);
    std::vector<int32_t> place_state(copies, -77);
    if (path == 3 || path == 4) {
        a.rec = place_mem.data();
        a.rec_off = rec_off.data();
        a.bm = a.rec + rec_entries;
        a.bm_bound = bm_bound;
        a.place_state = place_state.data();
        a.n_long = 0;
    }

    auto cta_order = [&](unsigned grid) {
        std::vector<unsigned> o(grid);
        for (unsigned i = 0; i < grid; i++) o[i] = i;
        if (order == 1) std::reverse(o.begin(), o.end());
        if (order >= 2) std::shuffle(o.begin(), o.end(), std::mt19937(order));
        return o;
    };
    // launch_entropy's last kernel, then launch_execute (api.cu)
    warpsim::launch(1, kScanThreads, [&] { k_scan_blocks(a); });
    if (dev_total != total) return SZB_ERR_INVALID_ARGUMENT;
    if (a.rec) {
        warpsim::launch(3, 256, [&] { k_place_zero(a); });
        warpsim::launch((copies + kResolveWarps * 32 - 1) / (kResolveWarps * 32), kResolveWarps * 32, [&] { k_resolve(a); });
    }
    warpsim::launch((copies + 3) / 4, 128, [&] { k_frame_verdict(a); });
    if (a.n_body) {
        const unsigned g = (a.n_body + kWarpsPerCta - 1) / kWarpsPerCta;
        auto o = cta_order(g);
        warpsim::launch(g, kCtaThreads, [&] { k_execute_bodies(a); }, &o);
    }
    bool fallback = false;
    if (path == 3 || path == 4) {
        for (uint32_t c = 0; c < copies; c++)
            if (place_state[c] != (total > bm_bound ? kPlaceFallback : (frame_status[c] == SZB_ERR_DST_TOO_SMALL ? 0 : frame_status[c]))) return SZB_ERR_INVALID_ARGUMENT;
        warpsim::launch((copies + kPlaceWarps - 1) / kPlaceWarps, kPlaceWarps * 32, [&] { k_place(a, 0, copies); });
        warpsim::launch((copies + kWarpsPerCta - 1) / kWarpsPerCta, kCtaThreads, [&] { k_execute(a, 0, copies); });
    } else if (path == 5) {  // k_execute2 (exec2.cuh), with k_execute launched behind it for the frames it does not take
        warpsim::launch((copies + kX2Warps - 1) / kX2Warps, kX2Warps * 32, [&] { k_execute2<false>(a, 0, copies); });
        warpsim::launch((copies + kWarpsPerCta - 1) / kWarpsPerCta, kCtaThreads, [&] { k_execute(a, 0, copies); });
    } else if (path == 7) {
        warpsim::launch(copies, (kX2Team + 1) * 32, [&] { k_execute_team(a, 0, copies); });
        warpsim::launch(copies, 64, [&] { k_execute_pair(a, 0, copies); });
    } else if (path == 6) {
        warpsim::launch(copies, 64, [&] { k_execute_pair2(a, 0, copies); });
        warpsim::launch(copies, 64, [&] { k_execute_pair(a, 0, copies); });
    } else if (path == 0) {
        warpsim::launch((copies + kWarpsPerCta - 1) / kWarpsPerCta, kCtaThreads, [&] { k_execute(a, 0, copies); });
    } else {
        warpsim::launch(copies, 64, [&] { k_execute_pair(a, 0, copies); });
    }
    if (path == 2) {
        const unsigned g_blocks = (n_lb + kWarpsPerCta - 1) / kWarpsPerCta, g_slices = (a.n_ls + kWarpsPerCta - 1) / kWarpsPerCta;
        {
            auto o = cta_order(g_slices);
            warpsim::launch(g_slices, kCtaThreads, [&] { k_long_hist(a); }, &o);
        }
        {
            auto o = cta_order(g_blocks);
            warpsim::launch(g_blocks, kCtaThreads, [&] { k_long_blockscan(a); }, &o);
        }
        warpsim::launch(copies, 32, [&] { k_long_compose(a); });
        {
            auto o = cta_order(g_slices);
            warpsim::launch(g_slices, kCtaThreads, [&] { k_long_emit(a); }, &o);
        }
        {
            const unsigned tiles = (unsigned)(long_dbase.back() / kJumpTile);
            const unsigned grid = order == 0 ? std::max(1u, tiles / 24) : std::max(1u, tiles / 8 + 1);  // order 0: several tiles per warp
            auto o = cta_order(grid);
            warpsim::launch(grid, kJumpThreads, [&] { k_long_jump(a); }, &o);
        }
        warpsim::launch(1, 128, [&] { k_long_verdict(a); });
        // a frame beyond its scratch bound is not this path's: no kernel of it may have touched the frame's cells
        for (uint32_t slot = 0; slot < copies; slot++) {
            if (frame_status[exec_list[slot]] != SZB_OK) continue;
            if (frame_out_len[exec_list[slot]] <= long_dbase[slot + 1] - long_dbase[slot]) continue;
            fallback = true;
            for (uint64_t i = long_dbase[slot]; i < long_dbase[slot + 1]; i++)
                if (dist[i] != 0xCDCDCDCDu) return SZB_ERR_INVALID_ARGUMENT;
        }
    }
    if (exec_err) *exec_err = frame_status[0];
    for (uint32_t c = 0; c < copies; c++)
        if (frame_status[c] != SZB_OK) return frame_status[c];
    if (verify_checksum) {  // SZB_FLAG_VERIFY_CHECKSUM: XXH64 of the output against the frame's trailer
        warpsim::launch((copies + 63) / 64, 64, [&] { k_verify_checksums(a); });
        for (uint32_t c = 0; c < copies; c++)
            if (frame_status[c] != SZB_OK) return frame_status[c];
        if (verify_checksum == 2) {  // and a flipped output byte must be noticed
            if (!fr0.checksum_valid || total == 0) return 2;
            out[total / copies / 2] ^= 0x40;
            warpsim::launch((copies + 63) / 64, 64, [&] { k_verify_checksums(a); });
            out[total / copies / 2] ^= 0x40;
            return frame_status[0] == SZB_ERR_CHECKSUM_MISMATCH && frame_out_len[0] == 0 ? 3 : SZB_ERR_INVALID_ARGUMENT;
        }
    }
    for (uint32_t c = 0; c < copies; c++)
        if (frame_out_off[c] != c * (total / copies) || frame_out_len[c] != total / copies) return SZB_ERR_INVALID_ARGUMENT;
    if (copies == 2 && memcmp(out, out + total / 2, total / 2) != 0) return SZB_ERR_INVALID_ARGUMENT;
    *out_len = total / copies;
    return fallback ? 1 : SZB_OK;
}


// k_long_compose alone: random transfer functions for `nblocks` blocks of one frame; the warp scan must give every block
// the history that applying the functions one after the other gives.  Returns the number of mismatching blocks.
int hostsim_compose_selftest(uint32_t nblocks, uint32_t seed) {
    std::mt19937 rng(seed);
    std::vector<uint64_t> T(3 * (size_t)nblocks);
    for (auto &t : T) {
        const uint32_t kind = rng() % 4;
        if (kind == 0) t = (uint64_t)(1 + rng() % 100000);            // a constant
        else t = sym_entry(rng() % 3) + (kind == 1 ? 0 : rng() % 5);  // entry i, minus 0..4
    }
    for (uint32_t b = 0; b < nblocks; b += 7)                          // some blocks without sequences: the identity
        for (int k = 0; k < 3; k++) T[3 * (size_t)b + k] = sym_entry(k);
    std::vector<uint32_t> want(3 * (size_t)nblocks), got(3 * (size_t)nblocks, 0);
    uint32_t h[3] = {1, 4, 8};
    for (uint32_t b = 0; b < nblocks; b++) {
        for (int k = 0; k < 3; k++) want[3 * (size_t)b + k] = h[k];
        const uint32_t n0 = hist_apply(T[3 * (size_t)b], h[0], h[1], h[2]), n1 = hist_apply(T[3 * (size_t)b + 1], h[0], h[1], h[2]),
                       n2 = hist_apply(T[3 * (size_t)b + 2], h[0], h[1], h[2]);
        h[0] = n0, h[1] = n1, h[2] = n2;
    }
    const uint32_t first_lb[2] = {0, nblocks};
    DeviceBatch a{};
    a.n_long = 1;
    a.long_first_lb = first_lb;
    a.long_T = T.data();
    a.long_hist = got.data();
    warpsim::launch(1, 32, [&] { k_long_compose(a); });
    int bad = 0;
    for (uint32_t b = 0; b < nblocks; b++)
        if (memcmp(&want[3 * (size_t)b], &got[3 * (size_t)b], 12) != 0) bad++;
    return bad;
}

}  // extern "C"
