// long_tables.h -- host side of the block-parallel stage 4 for long frames (execute_long.cuh): the index tables the kernels
// walk.  Plain C++ (api.cu builds them per batch; tests/host_sim builds the same tables for the emulated kernels).
#pragma once
#include <stdint.h>

#include <vector>

#include "../../include/szb200.h"

namespace szb {

struct LongTables {
    std::vector<uint32_t> lb_block, lb_slot;  // every block of the long frames, frame after frame: block index, frame's slot
    std::vector<uint32_t> long_first_lb;      // per long frame (slot): first lb entry; n_long + 1 entries
    std::vector<uint64_t> long_dbase;         // per long frame: first distance cell; n_long + 1 entries, multiples of `tile`
    std::vector<uint32_t> ls_lb, ls_seq0;     // SLICES (one warp of k_long_hist / k_long_emit each): lb entry, first sequence
    std::vector<uint32_t> lb_first_ls;        // per lb entry: first slice; n_lb + 1 entries
    void clear() {
        lb_block.clear();
        lb_slot.clear();
        ls_lb.clear();
        ls_seq0.clear();
        long_first_lb.assign(1, 0);
        long_dbase.assign(1, 0);
        lb_first_ls.assign(1, 0);
    }
};

// exec_list[0, n_long): the long frames.  slice: sequences per slice (a multiple of 32), 0 = one slice per block.
// The cell bound of a frame: a Raw/RLE block regenerates Block_Size bytes, a compressed block with sequences at most
// Block_Maximum_Size = 128 KiB when the frame is valid (a frame that regenerates more is left to k_execute_pair).
inline void build_long_tables(const szb_frame_desc *frames, const szb_block_desc *blocks, const uint32_t *exec_list, uint32_t n_long,
                              uint32_t slice, uint32_t tile, LongTables &t) {
    t.clear();
    uint64_t cells = 0;
    for (uint32_t slot = 0; slot < n_long; slot++) {
        const szb_frame_desc &fr = frames[exec_list[slot]];
        uint64_t bound = 0;
        for (uint32_t i = 0; i < fr.nblocks; i++) {
            const szb_block_desc &d = blocks[fr.first_block + i];
            const uint32_t lb = (uint32_t)t.lb_block.size();
            t.lb_block.push_back(fr.first_block + i);
            t.lb_slot.push_back(slot);
            const uint32_t nseq = d.type == 2 ? d.nseq : 0;
            bound += d.type == 2 ? (nseq ? 128 * 1024 : d.lit_regen) : d.block_size;
            uint32_t s0 = 0;
            do {  // a block without sequences is one slice too (its cells are zeroed)
                t.ls_lb.push_back(lb);
                t.ls_seq0.push_back(s0);
                s0 += slice ? slice : nseq;
            } while (slice && s0 < nseq);
            t.lb_first_ls.push_back((uint32_t)t.ls_lb.size());
        }
        cells += (bound + tile - 1) / tile * tile;
        t.long_first_lb.push_back((uint32_t)t.lb_block.size());
        t.long_dbase.push_back(cells);
    }
}

}  // namespace szb
