// batch.cuh -- the argument block every kernel of the decode path takes (DeviceBatch) and the launch constants.
// Plain structs: compiles for the host as well (tests/host_sim runs device code on the CPU).
#pragma once
#include <stdint.h>

#include "../../include/szb200.h"

namespace szb {

constexpr int kWarpsPerCta = 4;
constexpr int kCtaThreads = kWarpsPerCta * 32;
constexpr uint32_t kFull = 0xFFFFFFFFu;

#ifndef SZB_SEQ_LANES
#define SZB_SEQ_LANES 21
#endif
constexpr int kSeqLanes = SZB_SEQ_LANES;  // blocks per warp in k_decode_sequences: 21 x 2.6 KB (tables + bit ring), 4 warps per SM
constexpr uint32_t kTabSlotWords = 1280;  // LL 512 | ML 512 | OF 256

struct SeqInfo {
    uint8_t al_ll, al_of, al_ml, pad;
    uint32_t stream_off;  // offset of the backward bitstream after the sequences-section header
};

struct HufInfo {
    uint8_t max_bits, pad;
    uint16_t tree_bytes;  // size of the tree description inside the origin block's literals section
    int32_t status;
};

struct DeviceBatch {
    const uint8_t *src;
    const szb_block_desc *blocks;
    const szb_frame_desc *frames;
    uint32_t nblocks, nframes;
    const uint32_t *huf_list;  // blocks with Huffman-coded literals
    uint32_t n_huf;
    const uint32_t *hufo_list; // blocks that carry a Huffman tree description (literals type Compressed)
    uint32_t n_hufo;
    const uint32_t *huf_slot;  // per huf_list entry: position of its origin block in hufo_list
    uint16_t *huf_tabs;        // per hufo_list entry: 2^kMaxHufBits decode-table cells
    HufInfo *huf_info;         // per hufo_list entry
    const uint32_t *seq_list;  // blocks with nseq > 0
    uint32_t n_seq;
    uint8_t *litbuf;
    uint32_t *seq_ll, *seq_ml, *seq_of;  // one allocation: seq_ml = seq_ll + seq_stride, seq_of = seq_ll + 2 * seq_stride
    uint64_t seq_stride;
    uint16_t *seq_tabs;     // per seq_list entry: LL(512) | ML(512) | OF(256) 16-bit decode-table cells
    SeqInfo *seq_info;      // per seq_list entry
    uint64_t *out_size;     // per block regenerated size (host-initialised for Raw/RLE/zero-sequence blocks)
    uint64_t *out_off;      // per block exclusive prefix
    int32_t *lit_status;    // per block
    int32_t *seq_status;    // per block
    uint64_t *total;        // [0] = total output bytes
    const uint32_t *predef; // predefined LL(64) | OF(32) | ML(64) decode tables
    const uint8_t *bytefill; // 256 rows of 256 equal bytes (row v holds v)
    uint8_t *dst;
    uint64_t dst_cap;
    uint64_t *frame_out_off, *frame_out_len;
    int32_t *frame_status;
    uint32_t *frame_nexec;      // per frame: how many of its leading blocks stage 4 executes -- all of them when the frame is fine;
                                // the blocks in front of the first one that failed in stages 1-3 otherwise (the reference would have
                                // executed those before it met the failing block: an error in THEIR execution comes first); 0 when
                                // nothing may be written (destination too small)
    const uint32_t *exec_list;  // frames in the order k_execute starts them (most sequences first)
    const uint32_t *body_list;  // Raw / RLE / zero-sequence blocks: output independent of earlier output
    uint32_t n_body;
    // long frames: exec_list[0, n_long); their stage 4 is parallel over blocks and bytes (execute_long.cuh)
    uint32_t n_long;
    uint32_t n_lb;
    const uint32_t *lb_block;       // all blocks of the long frames, frame after frame in block order
    const uint32_t *lb_slot;        // per lb_block entry: its frame's position in exec_list
    const uint32_t *long_first_lb;  // per long frame: its first lb_block entry (n_long + 1 entries)
    const uint64_t *long_dbase;     // per long frame: its first distance cell (n_long + 1 entries, multiples of kJumpTile)
    uint32_t n_ls;                  // SLICES of those blocks: k_long_hist and k_long_emit run one warp per slice
    uint32_t long_slice;            // sequences per slice (a multiple of 32); 0: every block is one slice
    const uint32_t *ls_lb;          // per slice: its block's lb_block entry
    const uint32_t *ls_seq0;        // per slice: its first sequence in the block
    const uint32_t *lb_first_ls;    // per lb_block entry: its first slice (n_lb + 1 entries)
    uint64_t *ls_T;                 // per slice: its history transfer function (3 entries); after k_long_blockscan: the function
                                    // from the block's start to the slice's start
    uint64_t *ls_sum;               // per slice: sum of literal lengths, sum of literal + match lengths; after k_long_blockscan:
                                    // the exclusive prefixes within the block
    uint32_t *dist;                 // one distance cell per output byte of the long frames; nullptr: k_execute_pair takes them
    uint64_t *long_T;               // per lb_block entry: the block's history transfer function (3 entries)
    uint32_t *long_hist;            // per lb_block entry: the history the block starts with (3 entries)
    unsigned long long *long_ticket;  // k_long_jump hands out its tiles in address order
    unsigned long long *long_err;   // per long frame: the first error found while emitting (block << 40 | round << 8 | -code)
    // dictionary (szb_decode_batch_dict): history in front of every frame that uses it, and the repeat offsets it starts with
    const uint8_t *dict_content;    // nullptr: no dictionary
    uint32_t dict_len;
    uint32_t dict_rep[3];           // 1, 4, 8 for a raw-content dictionary
    const uint8_t *frame_dict;      // per frame: non-zero = decoded with the dictionary
    uint32_t exec2;                 // non-zero: k_execute2 (exec2.cuh) executes the frames one warp executes; 0: k_execute
    uint32_t pair2;                 // the long frames k_execute2 could take: 1 = k_execute_pair2, 2 = k_execute_team (exec2.cuh); 0: k_execute_pair
    // frames one warp executes (place.cuh): k_resolve -> k_place; nullptr: k_execute takes them all
    uint32_t *rec;                  // per block with sequences, at rec_off[block]: one entry per segment (literal run or match) in
                                    // output order: a match's offset (through the repeat history), or bit 31 | the match bytes
                                    // of the block in front of the literal run
    const uint64_t *rec_off;        // per block: its first entry in rec (multiples of 4)
    uint32_t *bm;                   // bitmap over output positions: a segment starts here
    uint64_t bm_bound;              // output bytes the bitmap was sized for (the host's upper bound)
    int32_t *place_state;           // per frame: 0 = k_place executes it, < 0 = the error k_resolve found, 1 = k_execute's
    uint32_t n_noplace;             // exec_list[0, n_noplace): the long frames and the frames with too many sequences for one lane
                                    // of k_resolve to walk (k_execute's)
};

#if defined(__CUDACC__) || defined(SZB_WARPSIM)  // device code, or the host-side warp emulator of tests/host_sim
// the first lane's non-zero code, 0 when every lane is fine
__device__ __forceinline__ int warp_first_error(int rc) {
    uint32_t bad = __ballot_sync(kFull, rc != 0);
    if (!bad) return 0;
    return __shfl_sync(kFull, rc, __ffs(bad) - 1);
}
#endif

}  // namespace szb
