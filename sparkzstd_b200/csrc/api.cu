// api.cu -- the C ABI of libszb200.so (include/szb200.h): context, batch objects, launches.
//
// Everything here is host plumbing around the four kernels in kernels.cuh.  There is no CPU
// decode path: without a CUDA device szb_ctx_create fails and nothing else can be called.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/szb200.h"
#include "kernels.cuh"
#include "long_tables.h"

using namespace szb;

struct szb_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    cudaEvent_t ev[13] = {};
    uint32_t *d_predef = nullptr;
    uint8_t *d_bytefill = nullptr;  // 256 rows of 256 equal bytes: the source of RLE literal runs (kernels.cuh, stage 4)
    std::string last_error;
    float timing[10] = {};
    uint64_t launches = 0;
    int sm_count = 148;
    // grow-only staging for the host-pointer entry points
    uint8_t *d_src = nullptr;
    size_t d_src_cap = 0;
    uint8_t *d_dst = nullptr;
    size_t d_dst_cap = 0;
    // copy streams for the pipelined host-buffer path
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaStream_t s_lit = nullptr;          // the literal chain runs beside the sequence chain (they meet at stage 4)
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    // Pinned staging for callers whose buffers are pageable (a Go slice, a Python bytes object): the copy engines only run
    // asynchronously from / to page-locked memory.  Two rings of kPinBytes slots, allocated on first use.
    static constexpr int kPinIn = 2, kPinOut = 3;
    static constexpr size_t kPinBytes = (size_t)32 << 20;
    szb_frame_desc last_frame = {};  // the header row of the frame szb_decompress_reader decoded last
    uint8_t *pin_in[kPinIn] = {}, *pin_out[kPinOut] = {};
    cudaEvent_t pin_in_ev[kPinIn] = {}, pin_out_ev[kPinOut] = {};
};

// A dictionary (szb200.h): parsed once on the host.  `payload` is a synthetic Compressed block that carries the dictionary's
// entropy tables in the places a real block keeps them -- a literals section with the Huffman tree description and no streams,
// a sequences section whose three FSE table descriptions follow the modes byte in the block order LL, OF, ML (the dictionary
// stores them OF, ML, LL) -- so that the table builders, and the Repeat / Treeless rules of the blocks that name it as their
// origin, work on it unchanged.  `blk` is its row of the block table (src_off is filled in per batch).
struct szb_dict {
    szb_ctx *ctx = nullptr;
    uint32_t id = 0;
    bool has_tables = false;
    uint32_t rep[3] = {1, 4, 8};
    std::vector<uint8_t> payload;
    szb_block_desc blk = {};
    uint8_t *d_content = nullptr;  // the dictionary's content (history in front of a frame), + 16 readable bytes
    uint32_t content_len = 0;
};

static cudaError_t pool_alloc(szb_ctx *ctx, void **p, size_t bytes);
#define CUDA_TRY(ctx, expr)                                                                          \
    do {                                                                                             \
        cudaError_t e__ = (expr);                                                                    \
        if (e__ != cudaSuccess) {                                                                    \
            (ctx)->last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);                 \
            return SZB_ERR_CUDA;                                                                     \
        }                                                                                            \
    } while (0)

struct szb_batch {
    szb_ctx *ctx = nullptr;
    uint32_t nframes = 0, nblocks = 0;
    size_t src_len = 0;
    std::vector<szb_frame_desc> frames;
    std::vector<szb_block_desc> blocks;
    std::vector<uint32_t> huf_list, seq_list, hufo_list, huf_slot, body_list, exec_list;
    uint64_t literal_bytes = 0, sequences = 0;
    std::vector<uint8_t> stage;  // host image of the descriptor tables (kept until the upload has certainly happened)
    // device
    void *d_tables = nullptr;  // one allocation: frames | blocks | lists | out_size_init
    szb_frame_desc *d_frames = nullptr;
    szb_block_desc *d_blocks = nullptr;
    uint32_t *d_huf_list = nullptr, *d_seq_list = nullptr, *d_hufo_list = nullptr, *d_huf_slot = nullptr;
    uint16_t *d_huf_tabs = nullptr;
    HufInfo *d_huf_info = nullptr;
    uint32_t *d_body_list = nullptr;
    uint32_t *d_exec_list = nullptr;
    uint32_t n_long = 0;  // leading entries of exec_list: long frames (execute_long.cuh, else k_execute_pair)
    uint32_t n_noplace = 0;  // leading entries of exec_list that k_resolve / k_place leave to the other kernels
    LongTables lt;            // index tables of the block-parallel path (long_tables.h)
    uint32_t long_slice = 0;  // sequences per slice of a block (0: the whole block)
    bool long_jump = false;  // the block-parallel path is on for this batch
    uint32_t *d_lb_block = nullptr, *d_lb_slot = nullptr, *d_long_first_lb = nullptr, *d_ls_lb = nullptr, *d_ls_seq0 = nullptr,
             *d_lb_first_ls = nullptr;
    uint64_t *d_long_dbase = nullptr;
    void *d_long = nullptr;  // one allocation: long_err, long_ticket | long_T | long_hist | ls_T | ls_sum | dist
    uint64_t *d_out_size_init = nullptr;
    void *d_state = nullptr;  // one allocation: out_size | out_off | total | frame_out_off | frame_out_len | statuses
    uint64_t *d_out_size = nullptr, *d_out_off = nullptr, *d_total = nullptr, *d_frame_out_off = nullptr,
             *d_frame_out_len = nullptr;
    int32_t *d_lit_status = nullptr, *d_seq_status = nullptr, *d_frame_status = nullptr;
    uint32_t *d_frame_nexec = nullptr;
    size_t status_bytes = 0;
    uint8_t *d_litbuf = nullptr;
    uint32_t *d_seq = nullptr;  // ll | ml | of
    uint16_t *d_seq_tabs = nullptr;  // FSE decode-table arena, kTabSlotWords 16-bit cells per block with sequences
    SeqInfo *d_seq_info = nullptr;
    // frames one warp executes (place.cuh)
    bool place = false;
    bool exec2 = false;      // k_execute2 (exec2.cuh) for the frames one warp executes
    const szb_dict *dict = nullptr;  // the batch is decoded with this dictionary (szb_decode_batch_dict)
    uint8_t *d_frame_dict = nullptr; // per frame: decoded with the dictionary
    std::vector<uint64_t> rec_off;
    uint64_t rec_entries = 0, bm_bound = 0, bm_words = 0;
    uint64_t *d_rec_off = nullptr;
    void *d_place = nullptr;  // one allocation: rec | bm
    int32_t *d_place_state = nullptr;
    bool entropy_done = false;
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Scratch arenas come from the device's stream-ordered memory pool with an unlimited release
// threshold: after the first batch of a given shape, creating a batch costs no cudaMalloc.
static cudaError_t pool_alloc(szb_ctx *ctx, void **p, size_t bytes) { return cudaMallocAsync(p, bytes ? bytes : 256, ctx->stream); }
static void pool_free(szb_ctx *ctx, void *p) {
    if (p) cudaFreeAsync(p, ctx->stream);
}

extern "C" {

const char *szb_version(void) { return "sparkzstd-b200 0.1 (sm_100a)"; }

const char *szb_strerror(int code) {
    switch (code) {
    case SZB_OK: return "ok";
    case SZB_ERR_IO: return "read or write callback failed";
    case SZB_ERR_WRONG_DICTIONARY: return "the frame names another dictionary than the one given";
    case SZB_ERR_WRONG_MAGICNUMBER: return "Magicnum is not correct";
    case SZB_ERR_CORRUPT_SIZES: return "The sizes of literal and sequence section did not add up to blocksize";
    case SZB_ERR_OUT_OF_BLOCKS: return "No blocks left in frame";
    case SZB_ERR_ILLEGAL_CONTENT_SIZE_FLAG: return "The SizeFlag for the Field ContentSize has an illegal value bigger than 3";
    case SZB_ERR_ILLEGAL_DICTIONARY_ID_FLAG: return "The SizeFlag for the Field DictionaryID has an illegal value bigger than 3";
    case SZB_ERR_NOT_ENOUGH_BYTES_FOR_BLOCK_HEADER: return "Not enough / too much bytes to decode the blockheader. Must be 3.";
    case SZB_ERR_ILLEGAL_BLOCK_TYPE: return "Illegal BlockType. Must be smaller than 3.";
    case SZB_ERR_ILLEGAL_BLOCK_SIZE: return "Illegal block-size. Must be lower than 128kb";
    case SZB_ERR_WRONG_JUMPTABLE_BYTES: return "Not enough bytes for jumptable dacoding. Must be 6";
    case SZB_ERR_CORRUPTED_JUMPTABLE: return "Bad jump table. Sizes dont add up to compressed size";
    case SZB_ERR_ILLEGAL_LITERAL_SECTION_TYPE: return "Illegal LiteralSectionType. Must be between 0 to 3";
    case SZB_ERR_ILLEGAL_LITERAL_SECTION_SIZE_FORMAT: return "Illegal LiteralSectionSizeformat. Must be between 0 to 3";
    case SZB_ERR_WRONG_SIZES_BYTES: return "Not enough bytes to decode sizes";
    case SZB_ERR_NO_HUFF_TABLE_TO_CARRY_OVER: return "No previous Huffmantree available";
    case SZB_ERR_STREAM_DIDNT_DECODE_TO_RIGHT_LENGTH: return "Huffstream did not decode to the correct length";
    case SZB_ERR_WRONG_SUM_OF_WEIGHTS: return "The weights didnt leave a power of two for the last weight";
    case SZB_ERR_CORRUPTED_HUFF_TREE: return "The tree in the description is corrupted";
    case SZB_ERR_BAD_PADDING: return "The padding at the end of the stream was more than a byte. Data is likely corrupted";
    case SZB_ERR_DIDNT_USE_ALL_BITS_TO_DECODE_HUFFMAN: return "Didnt read all bits to decode huffman stream. Data is likely corrupted";
    case SZB_ERR_NOT_ALL_BITS_USED: return "Did not read all bits to decode sequences. Likely data is corrupted.";
    case SZB_ERR_NO_LL_TABLE_TO_CARRY_OVER: return "Needed to copy old LiteralLenghts table but there was none";
    case SZB_ERR_NO_ML_TABLE_TO_CARRY_OVER: return "Needed to copy old MathcLenghts table but there was none";
    case SZB_ERR_NO_OF_TABLE_TO_CARRY_OVER: return "Needed to copy old Offsets table but there was none";
    case SZB_ERR_NOT_ALL_BYTES_USED_WHILE_SEQUENCE_DECODING: return "Didnt use all bytes from the sequence stream. Data is likely corrupted";
    case SZB_ERR_DIDNT_READ_ALL_PROBABILITIES: return "The probabilities didnt add up to the expected total sum";
    case SZB_ERR_NO_SYMBOL_FOR_STATE: return "Probably a bad state?";
    case SZB_ERR_CANT_UNWIND: return "Cant unwind more bits";
    case SZB_ERR_DIDNT_COPY_ALL_LITERAL_BYTES: return "Not enough bytes read to execute literals copy";
    case SZB_ERR_IDX_OUT_OF_BOUNDS: return "Index is out of bounds";
    case SZB_ERR_CANT_REPEAT_BYTES: return "You cant repeat bytes from before the first one";
    case SZB_ERR_DIDNT_DUMP_ALL: return "Did not write all bytes. Output will likely be corrupted";
    case SZB_ERR_UNEXPECTED_EOF: return "unexpected EOF";
    case SZB_ERR_PANIC: return "corrupt input (the reference decoder panics on it)";
    case SZB_ERR_NOMEM: return "out of memory";
    case SZB_ERR_UNSUPPORTED: return "input exceeds the zstd format limits this engine enforces";
    case SZB_ERR_DST_TOO_SMALL: return "destination buffer too small";
    case SZB_ERR_CUDA: return "CUDA error (see szb_ctx_last_error)";
    case SZB_ERR_INVALID_ARGUMENT: return "invalid argument";
    case SZB_ERR_NO_DEVICE: return "no usable CUDA device: this engine has no CPU fallback";
    case SZB_ERR_CHECKSUM_MISMATCH: return "content checksum mismatch";
    default: return "unknown error";
    }
}

int szb_ctx_create(int device, void *stream, szb_ctx **out) {
    if (!out) return SZB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        (void)cudaGetLastError();
        return SZB_ERR_NO_DEVICE;
    }
    szb_ctx *ctx = new (std::nothrow) szb_ctx();
    if (!ctx) return SZB_ERR_NOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete ctx;
        return SZB_ERR_NO_DEVICE;
    }
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
            delete ctx;
            return SZB_ERR_CUDA;
        }
        ctx->own_stream = true;
    }
    if (cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device) != cudaSuccess || ctx->sm_count <= 0) ctx->sm_count = 148;
    {
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            uint64_t keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    for (auto &e : ctx->ev) {
        if (cudaEventCreate(&e) != cudaSuccess) {
            szb_ctx_destroy(ctx);
            return SZB_ERR_CUDA;
        }
    }
    // predefined LL / OF / ML decode tables (fse/predefined.go:22-78), built once with the same
    // builder the kernels use and kept in device memory
    uint32_t predef[160];
    int16_t norm[64];
    uint16_t next[64];
    for (int i = 0; i < 36; i++) norm[i] = kLLDefaultNorm[i];
    fse_build_serial(norm, 36, 6, KIND_LL, predef, next);
    for (int i = 0; i < 29; i++) norm[i] = kOFDefaultNorm[i];
    fse_build_serial(norm, 29, 5, KIND_OF, predef + 64, next);
    for (int i = 0; i < 53; i++) norm[i] = kMLDefaultNorm[i];
    fse_build_serial(norm, 53, 6, KIND_ML, predef + 96, next);
    if (cudaMalloc(&ctx->d_predef, sizeof(predef)) != cudaSuccess ||
        cudaMemcpy(ctx->d_predef, predef, sizeof(predef), cudaMemcpyHostToDevice) != cudaSuccess) {
        szb_ctx_destroy(ctx);
        return SZB_ERR_CUDA;
    }
    {
        std::vector<uint8_t> fill(256 * 256);
        for (int v = 0; v < 256; v++) memset(fill.data() + 256 * v, v, 256);
        if (cudaMalloc(&ctx->d_bytefill, fill.size()) != cudaSuccess ||
            cudaMemcpy(ctx->d_bytefill, fill.data(), fill.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
            szb_ctx_destroy(ctx);
            return SZB_ERR_CUDA;
        }
    }
    cudaFuncSetAttribute(k_build_huf_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(HufSmem) * kWarpsPerCta));
    cudaFuncSetAttribute(k_build_seq_tables, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(sizeof(SeqSmem) * kWarpsPerCta));
    cudaFuncSetAttribute(k_decode_sequences, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSeqDecodeSmemBytes);
    cudaFuncSetAttribute(k_decode_sequences_multi, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSeqDecodeSmemBytes);
    cudaFuncSetAttribute(k_decode_literals, cudaFuncAttributePreferredSharedMemoryCarveout, 100);  // streams arrive by cp.async: L1 is not needed, resident warps are
    *out = ctx;
    return SZB_OK;
}

void szb_ctx_destroy(szb_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (auto &e : ctx->ev)
        if (e) cudaEventDestroy(e);
    if (ctx->d_predef) cudaFree(ctx->d_predef);
    if (ctx->d_bytefill) cudaFree(ctx->d_bytefill);
    if (ctx->d_src) cudaFree(ctx->d_src);
    if (ctx->d_dst) cudaFree(ctx->d_dst);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    if (ctx->s_lit) cudaStreamDestroy(ctx->s_lit);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    for (int i = 0; i < szb_ctx::kPinIn; i++) {
        if (ctx->pin_in[i]) cudaFreeHost(ctx->pin_in[i]);
        if (ctx->pin_in_ev[i]) cudaEventDestroy(ctx->pin_in_ev[i]);
    }
    for (int i = 0; i < szb_ctx::kPinOut; i++) {
        if (ctx->pin_out[i]) cudaFreeHost(ctx->pin_out[i]);
        if (ctx->pin_out_ev[i]) cudaEventDestroy(ctx->pin_out_ev[i]);
    }
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

void *szb_ctx_stream(szb_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
int szb_ctx_last_frame(szb_ctx *ctx, szb_frame_desc *out) {
    if (!ctx || !out) return SZB_ERR_INVALID_ARGUMENT;
    *out = ctx->last_frame;
    return SZB_OK;
}
const char *szb_ctx_last_error(szb_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }
uint64_t szb_launch_count(szb_ctx *ctx) { return ctx ? ctx->launches : 0; }

int szb_last_timing(szb_ctx *ctx, float *ms, int n) {
    if (!ctx || !ms) return 0;
    int k = n < 10 ? n : 10;
    for (int i = 0; i < k; i++) ms[i] = ctx->timing[i];
    return k;
}

void szb_free(void *p) { free(p); }

// ---------------------------------------------------------------------------------------------
// Frames with at least this many sequences (about 0.7 MiB of text) are executed by k_execute_pair, at most
// kMaxLongFrames of them (16 per SM: half the warps an SM holds): with more long frames than that the SMs are full of
// frames anyway and two warps per frame only cost (all 65 536 text frames as pairs: 10.2 -> 13.2 ms).
// SZB_LONG_SEQS overrides the threshold (tests force every frame through the pair kernel with it).
constexpr uint64_t kLongFrameSequences = 65536;
constexpr uint32_t kMaxLongFrames = 2368;
// measured on text-like frames: k_execute_pair 0.26 GB/s per frame at 10.7 bytes per sequence (profiles/r01d_bench_single256m_1gpu.json);
// the block-parallel path 38-49 GB/s over everything it is given (profiles/r01k_bench_single*.json)
// (one warp of k_long_emit needs ~3 ms for a 128 KiB block whatever the number of blocks: r01l_long_ncu_full_summary.txt)
constexpr double kPairSeqPerMs = 24000.0, kJumpCellsPerMs = 40e6, kJumpFloorMs = 3.0;
constexpr uint64_t kPlaceMaxSeqs = 16384;  // per frame, for k_resolve / k_place (place.cuh)
#ifndef SZB_DEFAULT_EXEC_PLACE
#define SZB_DEFAULT_EXEC_PLACE 0
#endif
constexpr bool kDefaultExecPlace = SZB_DEFAULT_EXEC_PLACE != 0;
#ifndef SZB_DEFAULT_SEQ
#define SZB_DEFAULT_SEQ 1
#endif

static int batch_upload_tables(szb_batch *b) {
    szb_ctx *ctx = b->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const uint32_t nb = b->nblocks, nf = b->nframes;
    std::vector<uint64_t> out_size_init(nb ? nb : 1);
    for (uint32_t i = 0; i < nb; i++) {
        const szb_block_desc &d = b->blocks[i];
        // Raw / RLE blocks regenerate Block_Size bytes; a compressed block regenerates its literals
        // plus the match lengths k_sequences adds later
        out_size_init[i] = d.type == 2 ? d.lit_regen : d.block_size;
        if (d.type == 2 && d.lit_type >= 2) b->huf_list.push_back(i);
        if (d.type == 2 && d.nseq > 0) b->seq_list.push_back(i);
        if ((d.type != 2 && d.block_size > 0) || (d.type == 2 && d.nseq == 0 && d.lit_regen > 0)) b->body_list.push_back(i);
    }
    {   // blocks that carry a tree description get a table slot; every Huffman block points at its origin's slot
        std::vector<uint32_t> slot_of_block(nb ? nb : 1, SZB_NONE);
        for (uint32_t i = 0; i < nb; i++) {
            const szb_block_desc &d = b->blocks[i];
            if (d.type == 2 && d.lit_type == 2) {
                slot_of_block[i] = (uint32_t)b->hufo_list.size();
                b->hufo_list.push_back(i);
            }
        }
        // k_decode_literals runs eight blocks (one lane per stream) per warp for as long as the longest stream takes: blocks
        // that share a tree stay neighbours (they share one copy of the table), such families ordered by their longest stream,
        // longest first (what scripts/lit_group_model.py measures on the mixed corpus: 33 % -> 43 % of the lanes busy)
        auto stream_len = [&](uint32_t i) {
            const szb_block_desc &d = b->blocks[i];
            return d.lit_streams == 1 ? d.lit_regen : (d.lit_regen + 3) / 4;
        };
        std::vector<uint32_t> fam_len(b->hufo_list.size() + 1, 0);
        auto fam_of = [&](uint32_t i) {
            const uint32_t o = b->blocks[i].huf_origin;
            const uint32_t sl = o < nb ? slot_of_block[o] : SZB_NONE;
            return sl == SZB_NONE ? (uint32_t)b->hufo_list.size() : sl;
        };
        for (uint32_t i : b->huf_list) fam_len[fam_of(i)] = std::max(fam_len[fam_of(i)], stream_len(i));
        static const bool lit_sort = !(getenv("SZB_LIT_SORT") && atoi(getenv("SZB_LIT_SORT")) == 0);
        if (lit_sort)
            std::stable_sort(b->huf_list.begin(), b->huf_list.end(), [&](uint32_t x, uint32_t y) {
                const uint32_t fx = fam_of(x), fy = fam_of(y);
                if (fam_len[fx] != fam_len[fy]) return fam_len[fx] > fam_len[fy];
                if (fx != fy) return fx < fy;
                return stream_len(x) > stream_len(y);
            });
        for (uint32_t i : b->huf_list) b->huf_slot.push_back(slot_of_block[b->blocks[i].huf_origin]);
    }
    // k_decode_sequences runs kSeqLanes blocks per warp in lock step: neighbours should have similar
    // sequence counts, and the longest blocks should start first
    std::stable_sort(b->seq_list.begin(), b->seq_list.end(),
                     [&](uint32_t x, uint32_t y) { return b->blocks[x].nseq > b->blocks[y].nseq; });
    // k_execute runs one warp per frame and a frame is sequential: the frames with the most sequences start first,
    // so that the longest one does not begin when the rest of the grid is already draining
    {
        std::vector<uint64_t> work(nf ? nf : 1, 0);
        for (uint32_t i : b->seq_list) work[b->blocks[i].frame] += b->blocks[i].nseq;
        b->exec_list.resize(nf);
        for (uint32_t f = 0; f < nf; f++) b->exec_list[f] = f;
        std::stable_sort(b->exec_list.begin(), b->exec_list.end(), [&](uint32_t x, uint32_t y) { return work[x] > work[y]; });
        // Long frames finish last, and a frame is sequential: they get two warps each (k_execute_pair).
        const uint64_t long_seqs = getenv("SZB_LONG_SEQS") ? strtoull(getenv("SZB_LONG_SEQS"), nullptr, 10) : kLongFrameSequences;
        uint32_t n_long = 0;
        while (n_long < nf && n_long < kMaxLongFrames && work[b->exec_list[n_long]] >= long_seqs && work[b->exec_list[n_long]] > 0) n_long++;
        if (b->dict) n_long = 0;  // dictionary batches are k_execute2's (below)
        b->n_long = n_long;
        // The block-parallel path (execute_long.cuh) wants every block of the long frames, and one distance cell per
        // output byte: the host knows an upper bound (a Raw/RLE block regenerates Block_Size bytes, a compressed one at
        // most Block_Maximum_Size = 128 KiB when the frame is valid; a frame that regenerates more stays on k_execute_pair).
        const char *mode = getenv("SZB_LONG_MODE");
        b->long_jump = n_long > 0 && !(mode && strcmp(mode, "pair") == 0);
        b->lt.clear();
        if (b->long_jump) {
            // SZB_LONG_SLICE (sequences, rounded up to whole rounds; 0 = one warp per block): k_long_hist and k_long_emit
            // run one warp per slice -- more, shorter warps for frames of few blocks -- at the price of cells that are only
            // resolved inside a slice when they leave k_long_emit.
            // Measured (profiles/README.md, r02a / r02j; one frame): 64 MiB 14.9 -> 27.5 GB/s and 256 MiB 39 -> 53 GB/s with slices of
            // 512, 1 GiB 58.7 -> 64.5 GB/s with 1 024 (61.4 with 4 096), 4 GiB 61.4 -> 64.9 GB/s with 1 024 (r02s).  Few blocks want
            // many short warps.
            uint64_t long_blocks = 0;
            for (uint32_t k = 0; k < n_long; k++) long_blocks += b->frames[b->exec_list[k]].nblocks;
            const uint32_t auto_slice = long_blocks <= 4096 ? 512u : 1024u;
            const uint32_t slice = getenv("SZB_LONG_SLICE") ? (uint32_t)((strtoul(getenv("SZB_LONG_SLICE"), nullptr, 10) + 31) / 32 * 32) : auto_slice;
            b->long_slice = slice;
            build_long_tables(b->frames.data(), b->blocks.data(), b->exec_list.data(), n_long, slice, kJumpTile, b->lt);
            const uint64_t cells = b->lt.long_dbase.back();
            static const uint64_t max_cells = (getenv("SZB_LONG_MAX_GIB") ? strtoull(getenv("SZB_LONG_MAX_GIB"), nullptr, 10) : 64) << 28;
            if (cells > max_cells) b->long_jump = false;  // 4 bytes per cell
            // Which path is faster depends on the batch.  k_execute_pair runs every long frame on its own two warps at
            // kPairSeqPerMs sequences per millisecond, all frames side by side (up to 16 per SM): its time is that of the
            // frame with the most sequences.  The block-parallel path works on all cells of all long frames together at
            // kJumpCellsPerMs: its time follows the sum.  A few huge frames: jump; thousands of 2 MiB frames: pair.
            // SZB_LONG_MODE=jump|pair overrides.
            if (b->long_jump && !(mode && strcmp(mode, "jump") == 0)) {
                const uint64_t most = work[b->exec_list[0]];  // exec_list is sorted by sequences, most first
                const uint64_t waves = (n_long + (uint64_t)ctx->sm_count * 16 - 1) / ((uint64_t)ctx->sm_count * 16);
                const double t_pair = (double)most * (double)waves / kPairSeqPerMs, t_jump = kJumpFloorMs + (double)cells / kJumpCellsPerMs;
                if (t_pair <= t_jump) b->long_jump = false;
            }
        }
        if (!b->long_jump) b->lt.clear();
    }
    // k_resolve / k_place (place.cuh): one entry per segment, one bitmap over the output positions.  The host
    // only knows an upper bound of the output: a Raw/RLE block regenerates Block_Size bytes, a compressed one at most
    // Block_Maximum_Size when the frame is valid, a frame that declares its content size no more than that; when the
    // frames regenerate more, k_execute takes them (place_on).
    {
        const char *em = getenv("SZB_EXEC");
        // k_resolve walks a frame with one lane: frames of more than kPlaceMaxSeqs sequences stay with k_execute
        // (exec_list is sorted by sequences, most first)
        const uint64_t max_seqs = getenv("SZB_PLACE_MAX_SEQS") ? strtoull(getenv("SZB_PLACE_MAX_SEQS"), nullptr, 10) : kPlaceMaxSeqs;
        b->n_noplace = b->n_long;
        {
            std::vector<uint64_t> work(nf ? nf : 1, 0);
            for (uint32_t i : b->seq_list) work[b->blocks[i].frame] += b->blocks[i].nseq;
            while (b->n_noplace < nf && work[b->exec_list[b->n_noplace]] > max_seqs) b->n_noplace++;
        }
        // SZB_EXEC=place|legacy picks stage 4 for the frames one warp executes; the default is what measured faster on the
        // headline workload (profiles/README.md, r02)
        const bool want_place = em ? strcmp(em, "place") == 0 : kDefaultExecPlace;
        b->place = want_place && nf > b->n_noplace;
        // SZB_EXEC=legacy: k_execute for every frame; exec2 (the default): k_execute2, k_execute for frames of 2 GiB and more
        b->exec2 = !b->place && !(em && strcmp(em, "legacy") == 0);
        if (b->dict) {  // only k_execute2 reaches into a dictionary's content: no long-frame paths, no k_place
            b->place = false;
            b->exec2 = true;
        }
        b->rec_off.assign(nb ? nb : 1, 0);
        uint64_t entries = 0, bound = 0;
        if (b->place) {
            for (uint32_t f = 0; f < nf; f++) {
                const szb_frame_desc &fr = b->frames[f];
                uint64_t fb = 0;
                for (uint32_t i = fr.first_block; i < fr.first_block + fr.nblocks; i++) {
                    const szb_block_desc &d = b->blocks[i];
                    fb += d.type != 2 ? d.block_size : (d.nseq ? 128u * 1024u : d.lit_regen);
                    if (d.type == 2 && d.nseq) {
                        b->rec_off[i] = entries;
                        entries += ((2 * (uint64_t)d.nseq + 1 + 3) & ~3ull) + 4;  // two segments per sequence + the literals after the last
                    }
                }
                if (fr.has_content_size && fr.content_size < fb) fb = fr.content_size;
                bound += fb;
            }
        }
        b->rec_entries = entries + 64;
        b->bm_bound = bound;
        b->bm_words = ((bound + 256) / 32 + 4ull * nf + 64 + 3) & ~3ull;  // a line's four words are one 16-byte load
    }
    // descriptor tables: one allocation, one H2D copy
    size_t o_frames = 0;
    size_t o_blocks = align_up(o_frames + sizeof(szb_frame_desc) * (size_t)nf, 256);
    size_t o_huf = align_up(o_blocks + sizeof(szb_block_desc) * (size_t)nb, 256);
    size_t o_seq = align_up(o_huf + 4 * b->huf_list.size(), 256);
    size_t o_hufo = align_up(o_seq + 4 * b->seq_list.size(), 256);
    size_t o_slot = align_up(o_hufo + 4 * b->hufo_list.size(), 256);
    size_t o_body = align_up(o_slot + 4 * b->huf_slot.size(), 256);
    size_t o_exec = align_up(o_body + 4 * b->body_list.size(), 256);
    size_t o_init = align_up(o_exec + 4 * b->exec_list.size(), 256);
    size_t o_lbb = align_up(o_init + 8 * (size_t)nb, 256);
    const LongTables &lt = b->lt;
    size_t o_lbs = align_up(o_lbb + 4 * lt.lb_block.size(), 256);
    size_t o_lfl = align_up(o_lbs + 4 * lt.lb_slot.size(), 256);
    size_t o_ldb = align_up(o_lfl + 4 * lt.long_first_lb.size(), 256);
    size_t o_lsl = align_up(o_ldb + 8 * lt.long_dbase.size(), 256);
    size_t o_lss = align_up(o_lsl + 4 * lt.ls_lb.size(), 256);
    size_t o_lfs = align_up(o_lss + 4 * lt.ls_seq0.size(), 256);
    size_t o_rec = align_up(o_lfs + 4 * lt.lb_first_ls.size(), 256);
    size_t total = align_up(o_rec + 8 * (size_t)(nb ? nb : 1), 256) + 256;
    std::vector<uint8_t> &stage = b->stage;
    stage.assign(total, 0);
    if (nf) memcpy(stage.data() + o_frames, b->frames.data(), sizeof(szb_frame_desc) * (size_t)nf);
    if (nb) memcpy(stage.data() + o_blocks, b->blocks.data(), sizeof(szb_block_desc) * (size_t)nb);
    if (!b->huf_list.empty()) memcpy(stage.data() + o_huf, b->huf_list.data(), 4 * b->huf_list.size());
    if (!b->seq_list.empty()) memcpy(stage.data() + o_seq, b->seq_list.data(), 4 * b->seq_list.size());
    if (!b->hufo_list.empty()) memcpy(stage.data() + o_hufo, b->hufo_list.data(), 4 * b->hufo_list.size());
    if (!b->huf_slot.empty()) memcpy(stage.data() + o_slot, b->huf_slot.data(), 4 * b->huf_slot.size());
    if (!b->body_list.empty()) memcpy(stage.data() + o_body, b->body_list.data(), 4 * b->body_list.size());
    if (!b->exec_list.empty()) memcpy(stage.data() + o_exec, b->exec_list.data(), 4 * b->exec_list.size());
    if (nb) memcpy(stage.data() + o_init, out_size_init.data(), 8 * (size_t)nb);
    if (!lt.lb_block.empty()) memcpy(stage.data() + o_lbb, lt.lb_block.data(), 4 * lt.lb_block.size());
    if (!lt.lb_slot.empty()) memcpy(stage.data() + o_lbs, lt.lb_slot.data(), 4 * lt.lb_slot.size());
    memcpy(stage.data() + o_lfl, lt.long_first_lb.data(), 4 * lt.long_first_lb.size());
    memcpy(stage.data() + o_ldb, lt.long_dbase.data(), 8 * lt.long_dbase.size());
    if (!lt.ls_lb.empty()) memcpy(stage.data() + o_lsl, lt.ls_lb.data(), 4 * lt.ls_lb.size());
    if (!lt.ls_seq0.empty()) memcpy(stage.data() + o_lss, lt.ls_seq0.data(), 4 * lt.ls_seq0.size());
    memcpy(stage.data() + o_lfs, lt.lb_first_ls.data(), 4 * lt.lb_first_ls.size());
    memcpy(stage.data() + o_rec, b->rec_off.data(), 8 * b->rec_off.size());
    CUDA_TRY(ctx, pool_alloc(ctx, (void **)&b->d_tables, total));
    CUDA_TRY(ctx, cudaMemcpyAsync(b->d_tables, stage.data(), total, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t *base = (uint8_t *)b->d_tables;
    b->d_frames = (szb_frame_desc *)(base + o_frames);
    b->d_blocks = (szb_block_desc *)(base + o_blocks);
    b->d_huf_list = (uint32_t *)(base + o_huf);
    b->d_seq_list = (uint32_t *)(base + o_seq);
    b->d_hufo_list = (uint32_t *)(base + o_hufo);
    b->d_huf_slot = (uint32_t *)(base + o_slot);
    b->d_body_list = (uint32_t *)(base + o_body);
    b->d_exec_list = (uint32_t *)(base + o_exec);
    b->d_out_size_init = (uint64_t *)(base + o_init);
    b->d_lb_block = (uint32_t *)(base + o_lbb);
    b->d_lb_slot = (uint32_t *)(base + o_lbs);
    b->d_long_first_lb = (uint32_t *)(base + o_lfl);
    b->d_long_dbase = (uint64_t *)(base + o_ldb);
    b->d_ls_lb = (uint32_t *)(base + o_lsl);
    b->d_ls_seq0 = (uint32_t *)(base + o_lss);
    b->d_lb_first_ls = (uint32_t *)(base + o_lfs);
    b->d_rec_off = (uint64_t *)(base + o_rec);
    // mutable state
    size_t s_out_size = 0;
    size_t s_out_off = align_up(s_out_size + 8 * (size_t)nb, 256);
    size_t s_total = align_up(s_out_off + 8 * (size_t)nb, 256);
    size_t s_foff = s_total + 256;
    size_t s_flen = align_up(s_foff + 8 * (size_t)nf, 256);
    size_t s_status = align_up(s_flen + 8 * (size_t)nf, 256);
    size_t s_lit = s_status;
    size_t s_seqs = s_lit + 4 * (size_t)nb;
    size_t s_fst = s_seqs + 4 * (size_t)nb;
    size_t s_pst = s_fst + 4 * (size_t)nf;
    b->status_bytes = 4 * (2 * (size_t)nb + 2 * (size_t)nf);
    size_t s_nexec = s_pst + 4 * (size_t)nf;
    size_t s_end = align_up(s_nexec + 4 * (size_t)nf, 256) + 256;
    CUDA_TRY(ctx, pool_alloc(ctx, (void **)&b->d_state, s_end));
    CUDA_TRY(ctx, cudaMemsetAsync(b->d_state, 0, s_end, ctx->stream));
    base = (uint8_t *)b->d_state;
    b->d_out_size = (uint64_t *)(base + s_out_size);
    b->d_out_off = (uint64_t *)(base + s_out_off);
    b->d_total = (uint64_t *)(base + s_total);
    b->d_frame_out_off = (uint64_t *)(base + s_foff);
    b->d_frame_out_len = (uint64_t *)(base + s_flen);
    b->d_lit_status = (int32_t *)(base + s_lit);
    b->d_seq_status = (int32_t *)(base + s_seqs);
    b->d_frame_status = (int32_t *)(base + s_fst);
    b->d_place_state = (int32_t *)(base + s_pst);
    b->d_frame_nexec = (uint32_t *)(base + s_nexec);
    // scratch arenas
    CUDA_TRY(ctx, pool_alloc(ctx, (void **)&b->d_litbuf, (size_t)b->literal_bytes + 256));
    CUDA_TRY(ctx, pool_alloc(ctx, (void **)&b->d_seq, (size_t)(b->sequences * 3 + 64) * 4));
    CUDA_TRY(ctx, pool_alloc(ctx, (void **)&b->d_seq_tabs, (b->seq_list.size() + 1) * (size_t)kTabSlotWords * 2 + 256));
    CUDA_TRY(ctx, pool_alloc(ctx, (void **)&b->d_seq_info, (b->seq_list.size() + 1) * sizeof(SeqInfo)));
    CUDA_TRY(ctx, pool_alloc(ctx, (void **)&b->d_huf_tabs, (b->hufo_list.size() + 1) * (size_t)(2u << kMaxHufBits)));
    CUDA_TRY(ctx, pool_alloc(ctx, (void **)&b->d_huf_info, (b->hufo_list.size() + 1) * sizeof(HufInfo)));
    if (b->place) CUDA_TRY(ctx, pool_alloc(ctx, &b->d_place, 4 * (size_t)(b->rec_entries + b->bm_words) + 256));
    return SZB_OK;
}

static int batch_create_from_tables_impl(szb_ctx *ctx, size_t src_len, const szb_frame_desc *frames, uint32_t nframes,
                                         const szb_block_desc *blocks, uint32_t nblocks, const szb_dict *dict, szb_batch **out);
int szb_batch_create_from_tables(szb_ctx *ctx, size_t src_len, const szb_frame_desc *frames, uint32_t nframes,
                                 const szb_block_desc *blocks, uint32_t nblocks, szb_batch **out) {
    return batch_create_from_tables_impl(ctx, src_len, frames, nframes, blocks, nblocks, nullptr, out);
}
static int batch_create_from_tables_impl(szb_ctx *ctx, size_t src_len, const szb_frame_desc *frames, uint32_t nframes,
                                         const szb_block_desc *blocks, uint32_t nblocks, const szb_dict *dict, szb_batch **out) {
    if (!ctx || !out || (nframes && !frames) || (nblocks && !blocks)) return SZB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    szb_batch *b = new (std::nothrow) szb_batch();
    if (!b) return SZB_ERR_NOMEM;
    b->ctx = ctx;
    b->dict = dict;
    b->nframes = nframes;
    b->nblocks = nblocks;
    b->src_len = src_len;
    try {
        b->frames.assign(frames, frames + nframes);
        b->blocks.assign(blocks, blocks + nblocks);
    } catch (const std::bad_alloc &) {
        delete b;
        return SZB_ERR_NOMEM;
    }
    // validate what the device will trust
    for (uint32_t f = 0; f < nframes; f++) {
        const szb_frame_desc &fr = b->frames[f];
        if ((uint64_t)fr.first_block + fr.nblocks > nblocks) {
            delete b;
            return SZB_ERR_INVALID_ARGUMENT;
        }
    }
    // The scratch ranges (literal bytes, sequence rows) must ascend without overlapping, as the walkers hand them out:
    // kernels of different blocks write them at the same time.
    uint64_t lit = 0, seq = 0;
    for (uint32_t i = 0; i < nblocks; i++) {
        const szb_block_desc &d = b->blocks[i];
        uint64_t payload = d.type == 1 ? 1 : d.block_size;
        bool bad = d.type > 2 || d.frame >= nframes || d.src_off > src_len || payload > src_len - d.src_off || d.block_size > 128 * 1024;
        if (!bad && d.type == 2) {
            bad = d.lit_type > 3 || (uint64_t)d.lit_hdr_bytes + d.lit_comp > d.block_size || d.lit_regen > 128 * 1024 ||
                  (uint64_t)d.seq_off + d.seq_hdr_bytes > d.block_size || d.seq_off != (uint32_t)d.lit_hdr_bytes + d.lit_comp ||
                  (d.lit_streams != 1 && d.lit_streams != 4);
            // Raw literals are read from the payload (lit_regen bytes after the header), RLE literals are its one byte
            if (!bad && d.lit_type == 0) bad = d.lit_comp != d.lit_regen;
            if (!bad && d.lit_type == 1) bad = d.lit_comp != 1;
            if (!bad && d.lit_type >= 2) {
                bad = d.huf_origin >= nblocks || d.huf_origin > i || b->blocks[d.huf_origin].type != 2 ||
                      b->blocks[d.huf_origin].lit_type != 2 || (d.lit_type == 2 && d.huf_origin != i) ||
                      (d.lit_buf_off & 15) != 0 || d.lit_buf_off < lit || d.lit_buf_off > (1ull << 48);
                lit = d.lit_buf_off + d.lit_regen;
            }
            if (!bad && d.nseq > 0) {
                const uint32_t org[3] = {d.ll_origin, d.of_origin, d.ml_origin};
                const int kinds[3] = {KIND_LL, KIND_OF, KIND_ML};
                for (int k = 0; k < 3 && !bad; k++) {
                    uint32_t m = field_mode(d.seq_modes, kinds[k]);
                    if (m == 3) {
                        bad = org[k] >= i || b->blocks[org[k]].type != 2 || b->blocks[org[k]].nseq == 0 ||
                              field_mode(b->blocks[org[k]].seq_modes, kinds[k]) == 3;
                    }
                }
                // k_decode_sequences stores rows 16 bytes at a time and stage 4 reads whole rounds: slices start at multiples of 32
                if (!bad) bad = (d.seq_buf_off & 31) != 0 || d.seq_buf_off < seq || d.seq_buf_off > (1ull << 40);
                seq = d.seq_buf_off + d.nseq;
            }
        }
        if (bad) {
            delete b;
            return SZB_ERR_INVALID_ARGUMENT;
        }
    }
    b->literal_bytes = lit;
    b->sequences = align_up(seq, 32);
    int rc = batch_upload_tables(b);
    if (rc) {
        szb_batch_destroy(b);
        return rc;
    }
    *out = b;
    return SZB_OK;
}

int szb_batch_create(szb_ctx *ctx, const uint8_t *h_src, size_t src_len, const uint64_t *frame_off,
                     const uint64_t *frame_len, uint32_t nframes, szb_batch **out) {
    if (!ctx || !out) return SZB_ERR_INVALID_ARGUMENT;
    szb_walk *w = nullptr;
    int rc = szb_walk_create(h_src, src_len, frame_off, frame_len, nframes, &w);
    if (rc) return rc;
    rc = szb_batch_create_from_tables(ctx, src_len, szb_walk_frames(w), szb_walk_nframes(w), szb_walk_blocks(w),
                                      szb_walk_nblocks(w), out);
    szb_walk_destroy(w);
    return rc;
}

// The stage scratch (literals, sequences, decode tables) is dead once the batch's kernels are in the stream: giving it
// back with cudaFreeAsync lets the next batch on the same stream reuse the very same memory (stream-ordered pool)
// instead of asking the driver for more.  The descriptor and status tables stay until szb_batch_destroy.
static void batch_release_scratch(szb_batch *b) {
    szb_ctx *ctx = b->ctx;
    void **p[] = {(void **)&b->d_litbuf, (void **)&b->d_seq, (void **)&b->d_seq_tabs, (void **)&b->d_seq_info, (void **)&b->d_huf_tabs,
                  (void **)&b->d_huf_info, &b->d_long, &b->d_place};
    for (void **q : p) {
        pool_free(ctx, *q);
        *q = nullptr;
    }
}

void szb_batch_destroy(szb_batch *b) {
    if (!b) return;
    cudaSetDevice(b->ctx->device);
    cudaStreamSynchronize(b->ctx->stream);
    pool_free(b->ctx, b->d_tables);
    pool_free(b->ctx, b->d_state);
    pool_free(b->ctx, b->d_litbuf);
    pool_free(b->ctx, b->d_seq);
    pool_free(b->ctx, b->d_seq_tabs);
    pool_free(b->ctx, b->d_seq_info);
    pool_free(b->ctx, b->d_huf_tabs);
    pool_free(b->ctx, b->d_huf_info);
    pool_free(b->ctx, b->d_long);
    pool_free(b->ctx, b->d_place);
    pool_free(b->ctx, b->d_frame_dict);
    delete b;
}

uint32_t szb_batch_nframes(const szb_batch *b) { return b ? b->nframes : 0; }
uint32_t szb_batch_nblocks(const szb_batch *b) { return b ? b->nblocks : 0; }

static DeviceBatch make_args(szb_batch *b, const void *d_src, void *d_dst, size_t dst_cap) {
    DeviceBatch a;
    a.src = (const uint8_t *)d_src;
    a.blocks = b->d_blocks;
    a.frames = b->d_frames;
    a.nblocks = b->nblocks;
    a.nframes = b->nframes;
    a.huf_list = b->d_huf_list;
    a.n_huf = (uint32_t)b->huf_list.size();
    a.hufo_list = b->d_hufo_list;
    a.n_hufo = (uint32_t)b->hufo_list.size();
    a.huf_slot = b->d_huf_slot;
    a.huf_tabs = b->d_huf_tabs;
    a.huf_info = b->d_huf_info;
    a.seq_list = b->d_seq_list;
    a.n_seq = (uint32_t)b->seq_list.size();
    a.litbuf = b->d_litbuf;
    a.seq_ll = b->d_seq;
    a.seq_ml = b->d_seq + b->sequences;
    a.seq_of = b->d_seq + 2 * b->sequences;
    a.seq_stride = b->sequences;
    a.seq_tabs = b->d_seq_tabs;
    a.seq_info = b->d_seq_info;
    a.out_size = b->d_out_size;
    a.out_off = b->d_out_off;
    a.lit_status = b->d_lit_status;
    a.seq_status = b->d_seq_status;
    a.total = b->d_total;
    a.predef = b->ctx->d_predef;
    a.bytefill = b->ctx->d_bytefill;
    a.dst = (uint8_t *)d_dst;
    a.dst_cap = dst_cap;
    a.frame_out_off = b->d_frame_out_off;
    a.frame_out_len = b->d_frame_out_len;
    a.frame_status = b->d_frame_status;
    a.frame_nexec = b->d_frame_nexec;
    a.exec_list = b->d_exec_list;
    a.body_list = b->d_body_list;
    a.n_body = (uint32_t)b->body_list.size();
    a.n_long = b->n_long;
    a.n_lb = (uint32_t)b->lt.lb_block.size();
    a.n_ls = (uint32_t)b->lt.ls_lb.size();
    a.long_slice = b->long_slice;
    a.ls_lb = b->d_ls_lb;
    a.ls_seq0 = b->d_ls_seq0;
    a.lb_first_ls = b->d_lb_first_ls;
    a.ls_T = nullptr;
    a.ls_sum = nullptr;
    a.lb_block = b->d_lb_block;
    a.lb_slot = b->d_lb_slot;
    a.long_first_lb = b->d_long_first_lb;
    a.long_dbase = b->d_long_dbase;
    a.dist = nullptr;
    a.long_T = nullptr;
    a.long_hist = nullptr;
    a.long_err = nullptr;
    a.long_ticket = nullptr;
    a.rec = b->place && b->d_place ? (uint32_t *)b->d_place : nullptr;
    a.rec_off = b->d_rec_off;
    a.bm = a.rec ? a.rec + b->rec_entries : nullptr;
    a.bm_bound = b->bm_bound;
    a.place_state = a.rec ? b->d_place_state : nullptr;
    a.n_noplace = b->n_noplace;
    a.exec2 = b->exec2 ? 1u : 0u;
    // The long frames that do not take the block-parallel path: SZB_PAIR2=1 (default) k_execute_pair2 (producer + one consumer
    // warp), 2 k_execute_team (producer + a team of consumer warps: measured slower, profiles/r03g_*), 0 k_execute_pair (round 1's)
    static const uint32_t pair2 = getenv("SZB_PAIR2") ? (uint32_t)atoi(getenv("SZB_PAIR2")) : 1u;
    a.pair2 = b->exec2 && !b->dict ? (pair2 > 2 ? 2u : pair2) : 0u;
    a.dict_content = b->dict ? b->dict->d_content : nullptr;
    a.dict_len = b->dict ? b->dict->content_len : 0;
    for (int k = 0; k < 3; k++) a.dict_rep[k] = b->dict ? b->dict->rep[k] : (k == 0 ? 1u : (k == 1 ? 4u : 8u));
    a.frame_dict = b->dict ? b->d_frame_dict : nullptr;
    return a;
}

static int launch_entropy(szb_batch *b, const void *d_src) {
    szb_ctx *ctx = b->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    DeviceBatch a = make_args(b, d_src, nullptr, 0);
    if (!ctx->s_lit) {
        CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->s_lit, cudaStreamNonBlocking));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
        CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    }
    cudaStream_t sl = ctx->s_lit;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[0], s));
    CUDA_TRY(ctx, cudaMemsetAsync(b->d_lit_status, 0, b->status_bytes, s));
    if (b->nblocks)
        CUDA_TRY(ctx, cudaMemcpyAsync(b->d_out_size, b->d_out_size_init, 8 * (size_t)b->nblocks, cudaMemcpyDeviceToDevice, s));
    // Stage 2 (literals) and stage 3 (sequences) are independent until stage 4.  The sequence kernel is
    // launched first and owns the SMs (shared memory); the literal chain, on its own stream, fills the
    // SMs the last, partial wave of the sequence kernel leaves idle.
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, s));
    CUDA_TRY(ctx, cudaStreamWaitEvent(sl, ctx->ev_fork, 0));
    if (a.n_seq) {
        k_build_seq_tables<<<(a.n_seq + kWarpsPerCta - 1) / kWarpsPerCta, kCtaThreads, sizeof(SeqSmem) * kWarpsPerCta, s>>>(a);
        ctx->launches++;
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[7], s));
    if (a.n_seq) {
        // SZB_SEQ=1: one lane per block (k_decode_sequences); 3: three lanes per block, one per FSE state (sequences3.cuh)
        static const int seq_mode = getenv("SZB_SEQ") ? atoi(getenv("SZB_SEQ")) : SZB_DEFAULT_SEQ;
        if (seq_mode == 3)
            k_decode_sequences3<<<(a.n_seq + kSeq3Chains - 1) / kSeq3Chains, 32, kSeq3SmemBytes, s>>>(a);
        else
        {
            // SZB_SEQ_CTAS_PER_SM=n (experiments, tests): at most n CTAs per SM; a CTA then walks several groups
            static const int cap = getenv("SZB_SEQ_CTAS_PER_SM") ? atoi(getenv("SZB_SEQ_CTAS_PER_SM")) : 0;
            uint32_t grid = (a.n_seq + kSeqLanes - 1) / kSeqLanes;
            if (cap > 0 && grid > (uint32_t)(cap * ctx->sm_count))
                k_decode_sequences_multi<<<(uint32_t)(cap * ctx->sm_count), 32, kSeqDecodeSmemBytes, s>>>(a);
            else
                k_decode_sequences<<<grid, 32, kSeqDecodeSmemBytes, s>>>(a);
        }
        ctx->launches++;
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[2], s));
    if (a.n_hufo) {
        k_build_huf_tables<<<(a.n_hufo + kWarpsPerCta - 1) / kWarpsPerCta, kCtaThreads, sizeof(HufSmem) * kWarpsPerCta, sl>>>(a);
        ctx->launches++;
    }
    if (a.n_huf) {
        const uint32_t groups = (a.n_huf + kHufGroup - 1) / kHufGroup;
        k_decode_literals<<<(groups + kWarpsPerCta - 1) / kWarpsPerCta, kCtaThreads, 0, sl>>>(a);
        ctx->launches++;
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[1], sl));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, sl));
    CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->ev_join, 0));
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[10], s));
    k_scan_blocks<<<1, kScanThreads, 0, s>>>(a);
    ctx->launches++;
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[3], s));
    CUDA_TRY(ctx, cudaGetLastError());
    b->entropy_done = true;
    return SZB_OK;
}

static int launch_execute(szb_batch *b, const void *d_src, void *d_dst, size_t dst_cap) {
    szb_ctx *ctx = b->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->stream;
    DeviceBatch a = make_args(b, d_src, d_dst, dst_cap);
    if (a.nframes && a.rec) {  // place.cuh: every frame k_place will execute, walked in order by one lane
        k_place_zero<<<ctx->sm_count * 8, 256, 0, s>>>(a);
        k_resolve<<<(a.nframes + kResolveWarps * 32 - 1) / (kResolveWarps * 32), kResolveWarps * 32, 0, s>>>(a);
        ctx->launches += 2;
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[11], s));
    if (a.nframes) {
        k_frame_verdict<<<(a.nframes + 3) / 4, 128, 0, s>>>(a);
        ctx->launches++;
    }
    if (a.n_body) {
        k_execute_bodies<<<(a.n_body + kWarpsPerCta - 1) / kWarpsPerCta, kCtaThreads, 0, s>>>(a);
        ctx->launches++;
    }
    if (a.nframes && a.n_seq) {
        const uint32_t n_long = b->n_long, n_rest = a.nframes - n_long;
        if (n_long) {  // beside the others, on the second stream
            if (!ctx->s_lit) {
                CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->s_lit, cudaStreamNonBlocking));
                CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
                CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
            }
            cudaStream_t sl = ctx->s_lit;
            // scratch of the block-parallel path: error words | transfer functions | start histories | distance cells.
            // Without it (allocation refused) every long frame stays on k_execute_pair.
            const size_t n_lb = b->lt.lb_block.size(), n_ls = b->lt.ls_lb.size();
            const size_t o_T = align_up(8 * (size_t)n_long + 8, 256);
            const size_t o_hist = align_up(o_T + 24 * n_lb, 256);
            const size_t o_lsT = align_up(o_hist + 12 * n_lb, 256);
            const size_t o_lsum = align_up(o_lsT + 24 * n_ls, 256);
            const size_t o_dist = align_up(o_lsum + 16 * n_ls, 256);
            if (b->long_jump && !b->d_long) {
                const size_t bytes = o_dist + 4 * (size_t)b->lt.long_dbase.back();
                if (cudaMallocAsync(&b->d_long, bytes, s) != cudaSuccess) {
                    cudaGetLastError();
                    b->d_long = nullptr;
                    b->long_jump = false;
                }
            }
            if (b->long_jump) {
                uint8_t *base = (uint8_t *)b->d_long;
                a.long_err = (unsigned long long *)base;
                a.long_T = (uint64_t *)(base + o_T);
                a.long_hist = (uint32_t *)(base + o_hist);
                a.ls_T = (uint64_t *)(base + o_lsT);
                a.ls_sum = (uint64_t *)(base + o_lsum);
                a.dist = (uint32_t *)(base + o_dist);
                a.long_ticket = a.long_err + n_long;
                CUDA_TRY(ctx, cudaMemsetAsync(a.long_err, 0xFF, 8 * (size_t)n_long, s));
                CUDA_TRY(ctx, cudaMemsetAsync(a.long_ticket, 0, 8, s));
            }
            CUDA_TRY(ctx, cudaEventRecord(ctx->ev_fork, s));
            CUDA_TRY(ctx, cudaStreamWaitEvent(sl, ctx->ev_fork, 0));
            // the frames the block-parallel path does not take: k_execute_pair2, and k_execute_pair for those of 2 GiB and more
            if (a.pair2 == 2) {
                k_execute_team<<<n_long, (kX2Team + 1) * 32, 0, sl>>>(a, 0, n_long);
                ctx->launches++;
            } else if (a.pair2) {
                k_execute_pair2<<<n_long, 64, 0, sl>>>(a, 0, n_long);
                ctx->launches++;
            }
            k_execute_pair<<<n_long, 64, 0, sl>>>(a, 0, n_long);
            ctx->launches++;
            if (b->long_jump) {
                const uint32_t g_lb = (a.n_lb + kWarpsPerCta - 1) / kWarpsPerCta, g_ls = (a.n_ls + kWarpsPerCta - 1) / kWarpsPerCta;
                k_long_hist<<<g_ls, kCtaThreads, 0, sl>>>(a);
                k_long_blockscan<<<g_lb, kCtaThreads, 0, sl>>>(a);
                k_long_compose<<<n_long, 32, 0, sl>>>(a);
                k_long_emit<<<g_ls, kCtaThreads, 0, sl>>>(a);
                const uint64_t tiles = (b->lt.long_dbase.back() / kJumpTile + kJumpThreads / 32 - 1) / (kJumpThreads / 32);  // a tile per warp
                const uint64_t resident = (uint64_t)ctx->sm_count * SZB_JUMP_CTAS_PER_SM;
                k_long_jump<<<(unsigned)(tiles < resident ? (tiles ? tiles : 1) : resident), kJumpThreads, 0, sl>>>(a);
                k_long_verdict<<<(n_long + 127) / 128, 128, 0, sl>>>(a);
                ctx->launches += 6;
            }
            CUDA_TRY(ctx, cudaEventRecord(ctx->ev_join, sl));
        }
        if (a.rec) {
            const uint32_t n_place = a.nframes - b->n_noplace;
            k_place<<<(n_place + kPlaceWarps - 1) / kPlaceWarps, kPlaceWarps * 32, 0, s>>>(a, b->n_noplace, n_place);
            ctx->launches++;
        }
        if (n_rest && a.exec2) {  // exec2.cuh: every frame of the range that regenerates less than 2 GiB
            // SZB_X2_CTAS_PER_SM=N (experiments): dynamic shared memory nobody uses caps the CTAs an SM holds, i.e. the frames in
            // flight and with them the output that wants to stay in L2
            static const int cap_ctas = getenv("SZB_X2_CTAS_PER_SM") ? atoi(getenv("SZB_X2_CTAS_PER_SM")) : 0;
            size_t pad = 0;
            if (cap_ctas > 0) {
                const size_t per = (size_t)227 * 1024 / (size_t)cap_ctas;
                const size_t own = sizeof(X2Smem) * kX2Warps + 1024;
                pad = per > own ? per - own : 0;
                static bool once = false;
                if (!once) {
                    once = true;
                    cudaFuncSetAttribute(k_execute2<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad);
                }
            }
            if (b->dict)
                k_execute2<true><<<(n_rest + kX2Warps - 1) / kX2Warps, kX2Warps * 32, 0, s>>>(a, n_long, n_rest);
            else
                k_execute2<false><<<(n_rest + kX2Warps - 1) / kX2Warps, kX2Warps * 32, pad, s>>>(a, n_long, n_rest);
            ctx->launches++;
        }
        if (n_rest) {  // with k_place or k_execute2 on: only the frames they could not take (place_on, x2_takes)
            k_execute<<<(n_rest + kWarpsPerCta - 1) / kWarpsPerCta, kCtaThreads, 0, s>>>(a, n_long, n_rest);
            ctx->launches++;
        }
        if (n_long) CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->ev_join, 0));
    }
    CUDA_TRY(ctx, cudaEventRecord(ctx->ev[4], s));
    CUDA_TRY(ctx, cudaGetLastError());
    return SZB_OK;
}

static int collect_timing(szb_ctx *ctx) {
    // [1] literals and [2] sequences both start at ev[0] and overlap; [3] scan starts when both are done
    float t;
    CUDA_TRY(ctx, cudaEventSynchronize(ctx->ev[4]));
    CUDA_TRY(ctx, cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[4]));
    ctx->timing[0] = t;
    CUDA_TRY(ctx, cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[1]));
    ctx->timing[1] = t;
    CUDA_TRY(ctx, cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[2]));
    ctx->timing[2] = t;
    CUDA_TRY(ctx, cudaEventElapsedTime(&t, ctx->ev[10], ctx->ev[3]));
    ctx->timing[3] = t;
    CUDA_TRY(ctx, cudaEventElapsedTime(&t, ctx->ev[3], ctx->ev[4]));
    ctx->timing[4] = t;
    CUDA_TRY(ctx, cudaEventElapsedTime(&t, ctx->ev[0], ctx->ev[7]));
    ctx->timing[7] = t;  // table construction share of [2]
    CUDA_TRY(ctx, cudaEventElapsedTime(&t, ctx->ev[3], ctx->ev[11]));
    ctx->timing[8] = t;  // k_resolve's share of [4]
    return SZB_OK;
}

int szb_batch_decode_entropy(szb_batch *b, const void *d_src) {
    if (!b || (!d_src && b->src_len)) return SZB_ERR_INVALID_ARGUMENT;
    if (reinterpret_cast<uintptr_t>(d_src) & 15) return SZB_ERR_INVALID_ARGUMENT;  // szb200.h: device source buffers
    return launch_entropy(b, d_src);
}

int szb_batch_sizes(szb_batch *b, uint64_t *total, uint64_t *out_off, uint64_t *out_len) {
    if (!b || !b->entropy_done) return SZB_ERR_INVALID_ARGUMENT;
    szb_ctx *ctx = b->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const uint32_t nb = b->nblocks;
    std::vector<uint64_t> off(nb ? nb : 1), size(nb ? nb : 1);
    uint64_t tot = 0;
    if (nb) {
        CUDA_TRY(ctx, cudaMemcpyAsync(off.data(), b->d_out_off, 8 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(size.data(), b->d_out_size, 8 * (size_t)nb, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(&tot, b->d_total, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (total) *total = tot;
    for (uint32_t f = 0; f < b->nframes; f++) {
        const szb_frame_desc &fr = b->frames[f];
        uint64_t o = fr.nblocks ? off[fr.first_block] : 0;
        uint64_t l = fr.nblocks ? off[fr.first_block + fr.nblocks - 1] + size[fr.first_block + fr.nblocks - 1] - o : 0;
        if (out_off) out_off[f] = o;
        if (out_len) out_len[f] = l;
    }
    return SZB_OK;
}

int szb_batch_execute(szb_batch *b, const void *d_src, void *d_dst, size_t dst_cap) {
    if (!b || !b->entropy_done || (!d_dst && dst_cap)) return SZB_ERR_INVALID_ARGUMENT;
    return launch_execute(b, d_src, d_dst, dst_cap);
}

int szb_batch_run(szb_batch *b, const void *d_src, void *d_dst, size_t dst_cap) {
    if (!b || (reinterpret_cast<uintptr_t>(d_src) & 15)) return SZB_ERR_INVALID_ARGUMENT;  // szb200.h: device source buffers
    int rc = launch_entropy(b, d_src);
    if (rc) return rc;
    return launch_execute(b, d_src, d_dst, dst_cap);
}

int szb_batch_verify_checksums(szb_batch *b, void *d_dst) {
    if (!b || !d_dst) return SZB_ERR_INVALID_ARGUMENT;
    szb_ctx *ctx = b->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!b->nframes) return SZB_OK;
    DeviceBatch a = make_args(b, nullptr, d_dst, 0);
    k_verify_checksums<<<(a.nframes + 63) / 64, 64, 0, ctx->stream>>>(a);
    ctx->launches++;
    CUDA_TRY(ctx, cudaGetLastError());
    return SZB_OK;
}

int szb_batch_finish(szb_batch *b, int32_t *status) {
    if (!b) return SZB_ERR_INVALID_ARGUMENT;
    szb_ctx *ctx = b->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    std::vector<int32_t> st(b->nframes ? b->nframes : 1, 0);
    if (b->nframes)
        CUDA_TRY(ctx, cudaMemcpyAsync(st.data(), b->d_frame_status, 4 * (size_t)b->nframes, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaGetLastError());
    int rc = collect_timing(ctx);
    if (rc) return rc;
    int first = SZB_OK;
    for (uint32_t f = 0; f < b->nframes; f++) {
        if (status) status[f] = st[f];
        if (first == SZB_OK && st[f] != SZB_OK) first = st[f];
    }
    return first;
}

int szb_batch_read_literals(szb_batch *b, uint32_t block, uint8_t *dst, size_t cap) {
    if (!b || block >= b->nblocks || !dst) return SZB_ERR_INVALID_ARGUMENT;
    const szb_block_desc &d = b->blocks[block];
    if (d.type != 2 || d.lit_type < 2 || cap < d.lit_regen) return SZB_ERR_INVALID_ARGUMENT;
    szb_ctx *ctx = b->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    CUDA_TRY(ctx, cudaMemcpy(dst, b->d_litbuf + d.lit_buf_off, d.lit_regen, cudaMemcpyDeviceToHost));
    return SZB_OK;
}

int szb_batch_read_sequences(szb_batch *b, uint32_t block, uint32_t *ll, uint32_t *ml, uint32_t *of, size_t cap) {
    if (!b || block >= b->nblocks || !ll || !ml || !of) return SZB_ERR_INVALID_ARGUMENT;
    const szb_block_desc &d = b->blocks[block];
    if (d.type != 2 || cap < d.nseq) return SZB_ERR_INVALID_ARGUMENT;
    szb_ctx *ctx = b->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    size_t n = 4 * (size_t)d.nseq;
    if (n) {
        CUDA_TRY(ctx, cudaMemcpy(ll, b->d_seq + d.seq_buf_off, n, cudaMemcpyDeviceToHost));
        CUDA_TRY(ctx, cudaMemcpy(ml, b->d_seq + b->sequences + d.seq_buf_off, n, cudaMemcpyDeviceToHost));
        CUDA_TRY(ctx, cudaMemcpy(of, b->d_seq + 2 * b->sequences + d.seq_buf_off, n, cudaMemcpyDeviceToHost));
    }
    return SZB_OK;
}

int szb_batch_read_block_results(szb_batch *b, uint64_t *out_size, int32_t *status, size_t cap) {
    if (!b || cap < b->nblocks) return SZB_ERR_INVALID_ARGUMENT;
    szb_ctx *ctx = b->ctx;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    const uint32_t nb = b->nblocks;
    if (!nb) return SZB_OK;
    if (out_size) CUDA_TRY(ctx, cudaMemcpy(out_size, b->d_out_size, 8 * (size_t)nb, cudaMemcpyDeviceToHost));
    if (status) {
        std::vector<int32_t> l(nb), s(nb);
        CUDA_TRY(ctx, cudaMemcpy(l.data(), b->d_lit_status, 4 * (size_t)nb, cudaMemcpyDeviceToHost));
        CUDA_TRY(ctx, cudaMemcpy(s.data(), b->d_seq_status, 4 * (size_t)nb, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < nb; i++) status[i] = l[i] ? l[i] : s[i];
    }
    return SZB_OK;
}

// ---------------------------------------------------------------------------------------------
static int ensure_dev(szb_ctx *ctx, uint8_t **p, size_t *cap, size_t need) {
    if (need <= *cap && *p) return SZB_OK;
    if (*p) CUDA_TRY(ctx, cudaFree(*p));
    *p = nullptr;
    *cap = 0;
    size_t want = align_up(need + 256, 1 << 20);
    CUDA_TRY(ctx, cudaMalloc((void **)p, want));
    *cap = want;
    return SZB_OK;
}

// ---- pinned staging ---------------------------------------------------------
static int pin_rings(szb_ctx *ctx) {
    for (int i = 0; i < szb_ctx::kPinIn; i++) {
        if (!ctx->pin_in[i]) CUDA_TRY(ctx, cudaHostAlloc((void **)&ctx->pin_in[i], szb_ctx::kPinBytes, cudaHostAllocDefault));
        if (!ctx->pin_in_ev[i]) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->pin_in_ev[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < szb_ctx::kPinOut; i++) {
        if (!ctx->pin_out[i]) CUDA_TRY(ctx, cudaHostAlloc((void **)&ctx->pin_out[i], szb_ctx::kPinBytes, cudaHostAllocDefault));
        if (!ctx->pin_out_ev[i]) CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->pin_out_ev[i], cudaEventDisableTiming));
    }
    if (!ctx->s_h2d) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    if (!ctx->s_d2h) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    return SZB_OK;
}
// Is p ordinary (pageable) host memory?  cudaMemcpyAsync from / to it is staged by the driver and does not overlap anything.
static bool host_pageable(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}
static unsigned host_threads() {
    unsigned n = std::thread::hardware_concurrency();
    n = n < 2 ? 1 : (n > 8 ? 8 : n - 1);
    if (const char *wt = getenv("SZB_WALK_THREADS")) {
        const unsigned cap = (unsigned)strtoul(wt, nullptr, 10);
        if (cap >= 1 && cap < n) n = cap;
    }
    return n;
}
// memcpy on several threads: one core moves ~10 GB/s, the PCIe link five times that
static void parallel_memcpy(uint8_t *dst, const uint8_t *src, size_t n) {
    const size_t kMin = (size_t)4 << 20;
    unsigned t = host_threads();
    if (n < 2 * kMin || t < 2) {
        memcpy(dst, src, n);
        return;
    }
    if ((size_t)t > n / kMin) t = (unsigned)(n / kMin);
    const size_t part = align_up(n / t, 4096);
    std::vector<std::thread> th;
    size_t done = 0;
    try {
        for (unsigned i = 0; i + 1 < t && done + part < n; i++, done += part) th.emplace_back(memcpy, dst + done, src + done, part);
    } catch (...) {
    }
    memcpy(dst + done, src + done, n - done);
    for (auto &x : th) x.join();
}
// Host -> device through the pinned ring (pageable callers): the host-side copy of piece k+1 runs while piece k is on the link.
static int staged_h2d(szb_ctx *ctx, uint8_t *d, const uint8_t *h, size_t n, uint32_t *slot) {
    for (size_t off = 0; off < n;) {
        const size_t m = n - off < szb_ctx::kPinBytes ? n - off : szb_ctx::kPinBytes;
        const int k = (int)((*slot)++ % szb_ctx::kPinIn);
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->pin_in_ev[k]));  // the slot's last copy has left it
        parallel_memcpy(ctx->pin_in[k], h + off, m);
        CUDA_TRY(ctx, cudaMemcpyAsync(d + off, ctx->pin_in[k], m, cudaMemcpyHostToDevice, ctx->s_h2d));
        CUDA_TRY(ctx, cudaEventRecord(ctx->pin_in_ev[k], ctx->s_h2d));
        off += m;
    }
    return SZB_OK;
}
// Device -> pageable host memory: a drain thread waits for each piece to land in its pinned slot and copies it out, so the
// thread that submits work keeps running ahead by kPinOut pieces.
struct Drain {
    struct Job {
        int slot;
        uint8_t *dst;
        size_t n;
    };
    szb_ctx *ctx;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<Job> jobs;
    bool free_slot[szb_ctx::kPinOut];
    bool quit = false, failed = false;
    std::thread th;
    explicit Drain(szb_ctx *c) : ctx(c) {
        for (bool &f : free_slot) f = true;
        th = std::thread([this] { run(); });
    }
    void run() {
        cudaSetDevice(ctx->device);
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [this] { return quit || !jobs.empty(); });
                if (jobs.empty()) return;
                j = jobs.front();
                jobs.pop_front();
            }
            if (cudaEventSynchronize(ctx->pin_out_ev[j.slot]) != cudaSuccess) failed = true;
            else parallel_memcpy(j.dst, ctx->pin_out[j.slot], j.n);
            {
                std::lock_guard<std::mutex> lk(mu);
                free_slot[j.slot] = true;
            }
            cv.notify_all();
        }
    }
    // d -> h on s_d2h (which the caller has made wait for the producer of d), piece by piece
    int copy(const uint8_t *d, uint8_t *h, size_t n) {
        for (size_t off = 0; off < n;) {
            const size_t m = n - off < szb_ctx::kPinBytes ? n - off : szb_ctx::kPinBytes;
            int k = -1;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] {
                    for (int i = 0; i < szb_ctx::kPinOut; i++)
                        if (free_slot[i]) {
                            k = i;
                            return true;
                        }
                    return false;
                });
                free_slot[k] = false;
            }
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->pin_out[k], d + off, m, cudaMemcpyDeviceToHost, ctx->s_d2h));
            CUDA_TRY(ctx, cudaEventRecord(ctx->pin_out_ev[k], ctx->s_d2h));
            {
                std::lock_guard<std::mutex> lk(mu);
                jobs.push_back(Job{k, h + off, m});
            }
            cv.notify_all();
            off += m;
        }
        return SZB_OK;
    }
    // everything handed to copy() is in the caller's memory when this returns
    bool finish() {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
        }
        cv.notify_all();
        if (th.joinable()) th.join();
        return !failed;
    }
    ~Drain() { finish(); }
};

static int decode_tables(szb_ctx *ctx, szb_batch *b, const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap,
                         uint64_t *out_off, uint64_t *out_len, int32_t *status, uint32_t flags) {
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    const void *d_src = src;
    if ((flags & SZB_FLAG_SRC_DEVICE) && (reinterpret_cast<uintptr_t>(src) & 15)) return SZB_ERR_INVALID_ARGUMENT;  // szb200.h
    float h2d_ms = 0, d2h_ms = 0;
    if (!(flags & SZB_FLAG_SRC_DEVICE)) {
        int rc = ensure_dev(ctx, &ctx->d_src, &ctx->d_src_cap, src_len + 16);
        if (rc) return rc;
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev[8], ctx->stream));
        if (src_len) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_src, src, src_len, cudaMemcpyHostToDevice, ctx->stream));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev[9], ctx->stream));
        d_src = ctx->d_src;
    }
    int rc = launch_entropy(b, d_src);
    if (rc) return rc;
    uint64_t total = 0;
    rc = szb_batch_sizes(b, &total, nullptr, nullptr);
    if (rc) return rc;
    if (!(flags & SZB_FLAG_SRC_DEVICE)) CUDA_TRY(ctx, cudaEventElapsedTime(&h2d_ms, ctx->ev[8], ctx->ev[9]));
    void *d_dst = dst;
    size_t d_cap = dst_cap;
    if (!(flags & SZB_FLAG_DST_DEVICE)) {
        // too small a host buffer still runs stage 4 so every frame reports SZB_ERR_DST_TOO_SMALL
        size_t need = total <= dst_cap ? (size_t)total : 0;
        rc = ensure_dev(ctx, &ctx->d_dst, &ctx->d_dst_cap, need + 16);
        if (rc) return rc;
        d_dst = ctx->d_dst;
    }
    rc = launch_execute(b, d_src, d_dst, d_cap);
    if (rc) return rc;
    if ((flags & SZB_FLAG_VERIFY_CHECKSUM) && total <= d_cap) {
        rc = szb_batch_verify_checksums(b, d_dst);
        if (rc) return rc;
    }
    std::vector<int32_t> st(b->nframes ? b->nframes : 1, 0);
    int first = szb_batch_finish(b, st.data());
    if (first == SZB_ERR_CUDA) return first;
    std::vector<uint64_t> off(b->nframes ? b->nframes : 1), len(b->nframes ? b->nframes : 1);
    if (b->nframes) {
        CUDA_TRY(ctx, cudaMemcpy(off.data(), b->d_frame_out_off, 8 * (size_t)b->nframes, cudaMemcpyDeviceToHost));
        CUDA_TRY(ctx, cudaMemcpy(len.data(), b->d_frame_out_len, 8 * (size_t)b->nframes, cudaMemcpyDeviceToHost));
    }
    if (!(flags & SZB_FLAG_DST_DEVICE) && total <= dst_cap && total > 0) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev[8], ctx->stream));
        CUDA_TRY(ctx, cudaMemcpyAsync(dst, ctx->d_dst, (size_t)total, cudaMemcpyDeviceToHost, ctx->stream));
        CUDA_TRY(ctx, cudaEventRecord(ctx->ev[9], ctx->stream));
        CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        CUDA_TRY(ctx, cudaEventElapsedTime(&d2h_ms, ctx->ev[8], ctx->ev[9]));
    }
    ctx->timing[5] = h2d_ms;
    ctx->timing[6] = d2h_ms;
    for (uint32_t f = 0; f < b->nframes; f++) {
        if (out_off) out_off[f] = off[f];
        if (out_len) out_len[f] = len[f];
        if (status) status[f] = st[f];
    }
    return first;
}

// Host-buffer batch decode with the copies overlapped with the kernels: the frames are
// cut into chunks of consecutive frames; chunk i+1's compressed bytes travel H2D and chunk i-1's
// output travels D2H while chunk i's kernels run.  Needs every frame's content size (so that each
// chunk's place in dst is known without a device round trip); otherwise the caller falls back to the
// size-then-decode path.  Returns 1 when it did not apply.
static int decode_batch_pipelined(szb_ctx *ctx, const uint8_t *src, size_t src_len, const uint64_t *frame_off,
                                  const uint64_t *frame_len, uint32_t nframes, uint8_t *dst, size_t dst_cap,
                                  uint64_t *out_off, uint64_t *out_len, int32_t *status, int *first_rc, bool verify) {
    constexpr size_t kChunkBytes = 96u << 20;  // compressed bytes per chunk
    if (!frame_off || !frame_len || nframes < 64 || src_len < 2 * kChunkBytes) return 1;
    // sizes from the frame headers only (frame.go:49-61); any frame without one -> not applicable
    std::vector<uint64_t> fcs(nframes);
    uint64_t total = 0;
    for (uint32_t f = 0; f < nframes; f++) {
        const uint64_t off = frame_off[f], len = frame_len[f];
        if (off > src_len || len > src_len - off || len < 6) return 1;
        const uint8_t *p = src + off;
        if (!(p[0] == 0x28 && p[1] == 0xB5 && p[2] == 0x2F && p[3] == 0xFD)) return 1;
        const uint8_t fhd = p[4];
        const bool single = (fhd >> 5) & 1;
        const uint32_t flag = fhd >> 6;
        const uint32_t fcs_bytes = flag == 0 ? (single ? 1 : 0) : (flag == 1 ? 2 : (flag == 2 ? 4 : 8));
        const uint32_t dict_bytes = (fhd & 3) == 3 ? 4 : (fhd & 3);
        const uint32_t at = 5 + (single ? 0 : 1) + dict_bytes;
        if (fcs_bytes == 0 || at + fcs_bytes > len) return 1;
        uint64_t v = 0;
        for (uint32_t i = 0; i < fcs_bytes; i++) v |= (uint64_t)p[at + i] << (8 * i);
        if (fcs_bytes == 2) v += 256;
        fcs[f] = v;
        total += v;
    }
    if (total > dst_cap) return 1;  // let the plain path report SZB_ERR_DST_TOO_SMALL per frame
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    if (!ctx->s_h2d) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
    if (!ctx->s_d2h) CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
    int rc = ensure_dev(ctx, &ctx->d_src, &ctx->d_src_cap, src_len + 16);
    if (rc) return rc;
    rc = ensure_dev(ctx, &ctx->d_dst, &ctx->d_dst_cap, (size_t)total + 16);
    if (rc) return rc;
    // pageable caller memory goes through the context's pinned rings, so that the copies still overlap the kernels
    // (SZB_NO_STAGING=1, measurements only: hand pageable pointers to cudaMemcpyAsync as they are -- the driver then stages them
    // itself, synchronously)
    static const bool no_staging = getenv("SZB_NO_STAGING") != nullptr;
    const bool src_pg = !no_staging && host_pageable(src), dst_pg = !no_staging && host_pageable(dst);
    if (src_pg || dst_pg) {
        rc = pin_rings(ctx);
        if (rc) return rc;
    }
    std::unique_ptr<Drain> drain;
    if (dst_pg) drain.reset(new Drain(ctx));
    uint32_t in_slot = 0;

    struct Chunk {
        uint32_t f0, f1;
        uint64_t src_lo, src_hi, dst_lo, dst_len;
        szb_batch *batch;
        cudaEvent_t up, done;
        szb_walk *walk;
        int walk_rc;
    };
    std::vector<Chunk> chunks;
    uint64_t covered = 0;
    for (uint32_t f = 0; f < nframes;) {
        Chunk c{};
        c.f0 = f;
        c.src_lo = ~0ull;
        uint64_t bytes = 0, dpos = 0;
        while (f < nframes && (bytes < kChunkBytes || f == c.f0)) {
            c.src_lo = frame_off[f] < c.src_lo ? frame_off[f] : c.src_lo;
            c.src_hi = frame_off[f] + frame_len[f] > c.src_hi ? frame_off[f] + frame_len[f] : c.src_hi;
            bytes += frame_len[f];
            dpos += fcs[f];
            f++;
        }
        c.f1 = f;
        c.dst_len = dpos;
        covered += c.src_hi - c.src_lo;
        chunks.push_back(c);
    }
    if (covered > src_len + src_len / 2) return 1;  // frames are scattered through src: chunk copies would overlap too much
    {
        uint64_t pos = 0;
        for (auto &c : chunks) {
            c.dst_lo = pos;
            pos += c.dst_len;
        }
    }
    // The header walks (host only, frames are independent) run on worker threads, chunk by chunk, ahead of the
    // submission loop below; otherwise the single submitting thread, not the copy engines, sets the pace.
    std::vector<std::atomic<int>> walked(chunks.size());
    for (auto &w : walked) w.store(0, std::memory_order_relaxed);
    std::atomic<size_t> next_chunk{0};
    std::atomic<bool> stop{false};
    auto worker = [&]() {
        for (;;) {
            const size_t i = next_chunk.fetch_add(1);
            if (i >= chunks.size() || stop.load()) return;
            Chunk &c = chunks[i];
            c.walk_rc = szb_walk_create(src, src_len, frame_off + c.f0, frame_len + c.f0, c.f1 - c.f0, &c.walk);
            walked[i].store(1, std::memory_order_release);
        }
    };
    // SZB_WALK_THREADS caps the workers (one process per GPU on a shared host: cores / ranks; bench.py sets it)
    unsigned nthreads = std::thread::hardware_concurrency();
    nthreads = nthreads < 2 ? 1 : (nthreads > 8 ? 8 : nthreads - 1);
    if (const char *wt = getenv("SZB_WALK_THREADS")) {
        const unsigned cap = (unsigned)strtoul(wt, nullptr, 10);
        if (cap >= 1 && cap < nthreads) nthreads = cap;
    }
    if (nthreads > chunks.size()) nthreads = (unsigned)chunks.size();
    std::vector<std::thread> pool;
    try {
        for (unsigned t = 0; t < nthreads; t++) pool.emplace_back(worker);
    } catch (...) {
        // no (more) threads to be had: whatever started keeps going, and with none at all the walks happen right here
    }
    if (pool.empty()) worker();
    struct PoolJoin {
        std::vector<std::thread> &p;
        std::atomic<bool> &stop;
        ~PoolJoin() {
            stop.store(true);
            for (auto &t : p)
                if (t.joinable()) t.join();
        }
    } pool_join{pool, stop};

    cudaEvent_t ev_begin;
    CUDA_TRY(ctx, cudaEventCreate(&ev_begin));
    CUDA_TRY(ctx, cudaEventRecord(ev_begin, ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->s_h2d, ev_begin, 0));
    int fail = SZB_OK;
    const bool trace = getenv("SZB_TRACE") != nullptr;
    double t_wait = 0, t_create = 0, t_launch = 0;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_loop0 = now();
#define CUDA_BRK(expr)                                                                    \
    {                                                                                     \
        const cudaError_t e__ = (expr);                                                   \
        if (e__ != cudaSuccess) {                                                         \
            ctx->last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);        \
            fail = SZB_ERR_CUDA;                                                          \
            break;                                                                        \
        }                                                                                 \
    }
    for (size_t ci = 0; ci < chunks.size(); ci++) {
        Chunk &c = chunks[ci];
        CUDA_BRK(cudaEventCreateWithFlags(&c.up, cudaEventDisableTiming));
        CUDA_BRK(cudaEventCreateWithFlags(&c.done, cudaEventDisableTiming));
        // compressed bytes of this chunk: H2D on the copy stream
        if (src_pg) {
            rc = staged_h2d(ctx, ctx->d_src + c.src_lo, src + c.src_lo, (size_t)(c.src_hi - c.src_lo), &in_slot);
            if (rc) {
                fail = rc;
                break;
            }
        } else {
            CUDA_BRK(cudaMemcpyAsync(ctx->d_src + c.src_lo, src + c.src_lo, (size_t)(c.src_hi - c.src_lo), cudaMemcpyHostToDevice, ctx->s_h2d));
        }
        CUDA_BRK(cudaEventRecord(c.up, ctx->s_h2d));
        // descriptor tables of this chunk (walked by a worker), their upload, the kernels
        double t0 = now();
        while (!walked[ci].load(std::memory_order_acquire)) std::this_thread::yield();
        double t1 = now();
        t_wait += t1 - t0;
        rc = c.walk_rc;
        if (!rc) {
            rc = szb_batch_create_from_tables(ctx, src_len, szb_walk_frames(c.walk), szb_walk_nframes(c.walk), szb_walk_blocks(c.walk),
                                              szb_walk_nblocks(c.walk), &c.batch);
        }
        t0 = now();
        t_create += t0 - t1;
        if (c.walk) {
            szb_walk_destroy(c.walk);
            c.walk = nullptr;
        }
        if (rc) {
            fail = rc;
            break;
        }
        CUDA_BRK(cudaStreamWaitEvent(ctx->stream, c.up, 0));
        rc = launch_entropy(c.batch, ctx->d_src);
        if (!rc) rc = launch_execute(c.batch, ctx->d_src, ctx->d_dst + c.dst_lo, (size_t)(total - c.dst_lo));
        if (!rc && verify) rc = szb_batch_verify_checksums(c.batch, ctx->d_dst + c.dst_lo);
        if (rc) {
            fail = rc;
            break;
        }
        CUDA_BRK(cudaEventRecord(c.done, ctx->stream));
        batch_release_scratch(c.batch);
        t_launch += now() - t0;
        // output of this chunk: D2H on the other copy stream
        CUDA_BRK(cudaStreamWaitEvent(ctx->s_d2h, c.done, 0));
        if (c.dst_len && dst_pg) {
            rc = drain->copy(ctx->d_dst + c.dst_lo, dst + c.dst_lo, (size_t)c.dst_len);
            if (rc) {
                fail = rc;
                break;
            }
        } else if (c.dst_len) {
            CUDA_BRK(cudaMemcpyAsync(dst + c.dst_lo, ctx->d_dst + c.dst_lo, (size_t)c.dst_len, cudaMemcpyDeviceToHost, ctx->s_d2h));
        }
    }
#undef CUDA_BRK
    stop.store(true);
    for (auto &t : pool) t.join();
    const double t_loop1 = now();
    cudaStreamSynchronize(ctx->s_h2d);
    cudaStreamSynchronize(ctx->stream);
    cudaStreamSynchronize(ctx->s_d2h);
    if (drain && !drain->finish() && fail == SZB_OK) fail = SZB_ERR_CUDA;
    if (trace)
        fprintf(stderr, "[szb] pipelined: %zu chunks, submit loop %.1f ms (walk wait %.1f, tables %.1f, launches %.1f), drain %.1f ms\n",
                chunks.size(), t_loop1 - t_loop0, t_wait, t_create, t_launch, now() - t_loop1);
    *first_rc = SZB_OK;
    for (auto &c : chunks) {
        if (c.walk) szb_walk_destroy(c.walk);
        if (c.batch && fail == SZB_OK) {
            const uint32_t n = c.f1 - c.f0;
            std::vector<int32_t> st(n);
            std::vector<uint64_t> off(n), len(n);
            cudaMemcpy(st.data(), c.batch->d_frame_status, 4 * (size_t)n, cudaMemcpyDeviceToHost);
            cudaMemcpy(off.data(), c.batch->d_frame_out_off, 8 * (size_t)n, cudaMemcpyDeviceToHost);
            cudaMemcpy(len.data(), c.batch->d_frame_out_len, 8 * (size_t)n, cudaMemcpyDeviceToHost);
            // The chunk was placed and copied back by the sizes its frame headers declare.  Every frame, failed or not, must
            // start where those sizes put it and the chunk must end where they end it: a failed frame whose blocks regenerate
            // another size would otherwise shift its (valid) neighbours past what was copied back.
            uint64_t d_total = 0, expect = 0;
            cudaMemcpy(&d_total, c.batch->d_total, 8, cudaMemcpyDeviceToHost);
            if (d_total != c.dst_len) fail = SZB_ERR_CORRUPT_SIZES;
            for (uint32_t i = 0; i < n; i++) {
                if (off[i] != expect) fail = SZB_ERR_CORRUPT_SIZES;
                expect += fcs[c.f0 + i];
                if (status) status[c.f0 + i] = st[i];
                if (out_off) out_off[c.f0 + i] = c.dst_lo + off[i];
                if (out_len) out_len[c.f0 + i] = st[i] == SZB_OK ? len[i] : 0;
                if (*first_rc == SZB_OK && st[i] != SZB_OK) *first_rc = st[i];
                // a frame whose header lied about its size would have shifted its neighbours
                if (st[i] == SZB_OK && len[i] != fcs[c.f0 + i]) fail = SZB_ERR_CORRUPT_SIZES;
            }
        }
        if (c.batch) szb_batch_destroy(c.batch);
        if (c.up) cudaEventDestroy(c.up);
        if (c.done) cudaEventDestroy(c.done);
    }
    cudaEventDestroy(ev_begin);
    if (fail == SZB_ERR_CORRUPT_SIZES) return 1;  // redo on the plain path, which places frames by their decoded sizes
    if (fail != SZB_OK) return fail;
    (void)collect_timing(ctx);
    ctx->timing[5] = ctx->timing[6] = 0;
    return SZB_OK;
}

int szb_decode_batch(szb_ctx *ctx, const uint8_t *src, size_t src_len, const uint64_t *frame_off,
                     const uint64_t *frame_len, uint32_t nframes, uint8_t *dst, size_t dst_cap, uint64_t *out_off,
                     uint64_t *out_len, int32_t *status, uint32_t flags) {
    if (!ctx || (!src && src_len) || (!dst && dst_cap)) return SZB_ERR_INVALID_ARGUMENT;
    if (flags & SZB_FLAG_SRC_DEVICE) return SZB_ERR_INVALID_ARGUMENT;  // the header walk needs host bytes: use szb_decode_blocks
    if (!(flags & SZB_FLAG_DST_DEVICE)) {
        int first = SZB_OK;
        int prc = decode_batch_pipelined(ctx, src, src_len, frame_off, frame_len, nframes, dst, dst_cap, out_off, out_len, status, &first,
                                         (flags & SZB_FLAG_VERIFY_CHECKSUM) != 0);
        if (prc == SZB_OK) return first;
        if (prc != 1) return prc;
    }
    szb_batch *b = nullptr;
    int rc = szb_batch_create(ctx, src, src_len, frame_off, frame_len, nframes, &b);
    if (rc) return rc;
    if (frame_off == nullptr && (out_off || out_len || status) && b->nframes > nframes) {
        // discovered more frames than the caller sized its arrays for
        szb_batch_destroy(b);
        return SZB_ERR_INVALID_ARGUMENT;
    }
    rc = decode_tables(ctx, b, src, src_len, dst, dst_cap, out_off, out_len, status, flags);
    szb_batch_destroy(b);
    return rc;
}

// ---- dictionaries (szb200.h) ----------------------------------------------------
int szb_dict_create(szb_ctx *ctx, const uint8_t *dict, size_t len, szb_dict **out) {
    if (!ctx || !out || (!dict && len) || len > 0x7FFF0000u) return SZB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    szb_dict *d = new (std::nothrow) szb_dict();
    if (!d) return SZB_ERR_NOMEM;
    d->ctx = ctx;
    const uint8_t *content = dict;
    size_t content_len = len;
    int rc = SZB_OK;
    if (len >= 8 && dict[0] == 0x37 && dict[1] == 0xA4 && dict[2] == 0x30 && dict[3] == 0xEC) {  // formatted (RFC 8878 section 5)
        d->id = (uint32_t)dict[4] | ((uint32_t)dict[5] << 8) | ((uint32_t)dict[6] << 16) | ((uint32_t)dict[7] << 24);
        size_t pos = 8;
        // Huffman tree description: header byte, then FSE-compressed (< 128: that many bytes) or direct 4-bit weights (huffman.go:40-104)
        size_t tree_len = 0;
        if (pos >= len) {
            rc = SZB_ERR_UNEXPECTED_EOF;
        } else {
            const uint32_t hb = dict[pos];
            tree_len = hb < 128 ? 1 + (size_t)hb : 1 + ((size_t)(hb - 127) + 1) / 2;
            if (tree_len > len - pos) rc = SZB_ERR_UNEXPECTED_EOF;
        }
        const size_t tree_at = pos;
        pos += tree_len;
        // FSE table descriptions in the dictionary's order: offsets, match lengths, literal lengths
        size_t at[3] = {0, 0, 0}, sz[3] = {0, 0, 0};
        const uint32_t max_al[3] = {kMaxALOF, kMaxALML, kMaxALLL};
        for (int k = 0; k < 3 && !rc; k++) {
            int16_t norm[kMaxFseSymbols];
            uint32_t nsym = 0, al = 0, used = 0;
            rc = fse_read_description(dict + pos, (uint32_t)(len - pos < 0x10000 ? len - pos : 0x10000), max_al[k], norm, &nsym, &al, &used);
            at[k] = pos;
            sz[k] = used;
            pos += used;
        }
        if (!rc && len - pos < 12) rc = SZB_ERR_UNEXPECTED_EOF;
        if (!rc) {
            for (int k = 0; k < 3; k++) {
                d->rep[k] = (uint32_t)dict[pos] | ((uint32_t)dict[pos + 1] << 8) | ((uint32_t)dict[pos + 2] << 16) | ((uint32_t)dict[pos + 3] << 24);
                pos += 4;
            }
            content = dict + pos;
            content_len = len - pos;
            for (int k = 0; k < 3; k++)
                if (d->rep[k] == 0 || d->rep[k] > content_len) rc = SZB_ERR_CANT_REPEAT_BYTES;  // points outside the content
        }
        if (!rc) {
            // the synthetic block: [literals header: Compressed, 1 stream, regenerates 0, "compressed size" = the tree][tree]
            //                      [1 sequence][modes FSE FSE FSE][LL][OF][ML][one byte of bitstream that nobody decodes]
            std::vector<uint8_t> &p = d->payload;
            const uint32_t v = 2u | (0u << 2) | (0u << 4) | ((uint32_t)tree_len << 14);
            p.push_back((uint8_t)v);
            p.push_back((uint8_t)(v >> 8));
            p.push_back((uint8_t)(v >> 16));
            p.insert(p.end(), dict + tree_at, dict + tree_at + tree_len);
            const uint32_t seq_off = (uint32_t)p.size();
            p.push_back(1);
            p.push_back(0xA8);
            const int order[3] = {2, 0, 1};  // LL, OF, ML out of OF, ML, LL
            for (int k : order) p.insert(p.end(), dict + at[k], dict + at[k] + sz[k]);
            p.push_back(1);
            szb_block_desc &b = d->blk;
            memset(&b, 0, sizeof(b));
            b.block_size = (uint32_t)p.size();
            b.type = 2;
            b.lit_type = 2;
            b.lit_streams = 1;
            b.lit_hdr_bytes = 3;
            b.lit_regen = 0;
            b.lit_comp = (uint32_t)tree_len;
            b.nseq = 1;
            b.seq_off = seq_off;
            b.seq_hdr_bytes = 2;
            b.seq_modes = 0xA8;
            b.huf_origin = b.ll_origin = b.of_origin = b.ml_origin = 0;  // itself: row 0 of the block table
            b.flags = SZB_BLOCK_TABLES_ONLY;
            d->has_tables = true;
        }
    }
    if (rc) {
        delete d;
        return rc;
    }
    d->content_len = (uint32_t)content_len;
    cudaSetDevice(ctx->device);
    if (cudaMalloc((void **)&d->d_content, content_len + 16) != cudaSuccess ||
        (content_len && cudaMemcpy(d->d_content, content, content_len, cudaMemcpyHostToDevice) != cudaSuccess)) {
        (void)cudaGetLastError();
        if (d->d_content) cudaFree(d->d_content);
        delete d;
        return SZB_ERR_CUDA;
    }
    *out = d;
    return SZB_OK;
}

void szb_dict_destroy(szb_dict *d) {
    if (!d) return;
    cudaSetDevice(d->ctx->device);
    cudaStreamSynchronize(d->ctx->stream);
    if (d->d_content) cudaFree(d->d_content);
    delete d;
}

uint32_t szb_dict_id(const szb_dict *d) { return d ? d->id : 0; }

int szb_decode_batch_dict(szb_ctx *ctx, const szb_dict *dict, const uint8_t *src, size_t src_len, const uint64_t *frame_off,
                          const uint64_t *frame_len, uint32_t nframes, uint8_t *dst, size_t dst_cap, uint64_t *out_off,
                          uint64_t *out_len, int32_t *status, uint32_t flags) {
    if (!ctx || !dict || dict->ctx != ctx || (!src && src_len) || (!dst && dst_cap) || !frame_off || !frame_len)
        return SZB_ERR_INVALID_ARGUMENT;
    if (flags & (SZB_FLAG_SRC_DEVICE | SZB_FLAG_DST_DEVICE)) return SZB_ERR_INVALID_ARGUMENT;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    // the dictionary's table-only block lives behind the caller's bytes in the device copy of src
    const size_t at = align_up(src_len + 16, 16);
    const size_t ext_len = at + dict->payload.size();
    szb_block_desc blk = dict->blk;
    blk.src_off = at;
    szb_walk *w = nullptr;
    int rc = szb_walk_create_dict(src, src_len, frame_off, frame_len, nframes, dict->has_tables ? &blk : nullptr, dict->id, &w);
    if (rc) return rc;
    const uint32_t nf = szb_walk_nframes(w);  // nframes, + the pseudo frame of the table-only block
    szb_batch *b = nullptr;
    rc = batch_create_from_tables_impl(ctx, ext_len, szb_walk_frames(w), nf, szb_walk_blocks(w), szb_walk_nblocks(w), dict, &b);
    szb_walk_destroy(w);
    if (rc) return rc;
    auto fail = [&](int code) {
        szb_batch_destroy(b);
        return code;
    };
    {   // which frames are decoded with the dictionary: the caller's (a wrong Dictionary_ID has failed in the walk), not the pseudo frame
        std::vector<uint8_t> use(nf ? nf : 1, 0);
        for (uint32_t f = 0; f < nframes && f < nf; f++) use[f] = 1;
        if (pool_alloc(ctx, (void **)&b->d_frame_dict, use.size()) != cudaSuccess ||
            cudaMemcpyAsync(b->d_frame_dict, use.data(), use.size(), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
            cudaStreamSynchronize(ctx->stream) != cudaSuccess)
            return fail(SZB_ERR_CUDA);
    }
    rc = ensure_dev(ctx, &ctx->d_src, &ctx->d_src_cap, ext_len + 16);
    if (rc) return fail(rc);
    if ((src_len && cudaMemcpyAsync(ctx->d_src, src, src_len, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) ||
        (!dict->payload.empty() &&
         cudaMemcpyAsync(ctx->d_src + at, dict->payload.data(), dict->payload.size(), cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess))
        return fail(SZB_ERR_CUDA);
    std::vector<uint64_t> off(nf ? nf : 1), len(nf ? nf : 1);
    std::vector<int32_t> st(nf ? nf : 1, 0);
    rc = decode_tables(ctx, b, ctx->d_src, ext_len, dst, dst_cap, off.data(), len.data(), st.data(), flags | SZB_FLAG_SRC_DEVICE);
    int first = SZB_OK;
    for (uint32_t f = 0; f < nframes && f < nf; f++) {
        if (out_off) out_off[f] = off[f];
        if (out_len) out_len[f] = len[f];
        if (status) status[f] = st[f];
        if (first == SZB_OK && st[f] != SZB_OK) first = st[f];
    }
    szb_batch_destroy(b);
    if (rc == SZB_ERR_CUDA || rc == SZB_ERR_NOMEM) return rc;
    return first;  // (decode_tables' own verdict also counts the pseudo frame, whose status is never 0)
}

int szb_decode_stream(szb_ctx *ctx, const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, uint64_t *out_off,
                      uint64_t *out_len, int32_t *status, uint32_t max_frames, uint32_t *nframes_out, uint64_t *total_out,
                      uint32_t flags) {
    if (!ctx || (!src && src_len) || (!dst && dst_cap)) return SZB_ERR_INVALID_ARGUMENT;
    if (flags & SZB_FLAG_SRC_DEVICE) return SZB_ERR_INVALID_ARGUMENT;  // the header walk needs host bytes
    if (nframes_out) *nframes_out = 0;
    if (total_out) *total_out = 0;
    szb_batch *b = nullptr;
    int rc = szb_batch_create(ctx, src, src_len, nullptr, nullptr, 0, &b);  // frame boundaries from the walk
    if (rc) return rc;
    const uint32_t n = b->nframes;
    if (nframes_out) *nframes_out = n;
    if ((out_off || out_len || status) && n > max_frames) {
        szb_batch_destroy(b);
        return SZB_ERR_INVALID_ARGUMENT;
    }
    std::vector<uint64_t> off(n ? n : 1), len(n ? n : 1);
    std::vector<int32_t> st(n ? n : 1, 0);
    rc = decode_tables(ctx, b, src, src_len, dst, dst_cap, off.data(), len.data(), st.data(), flags);
    uint64_t total = 0;
    for (uint32_t f = 0; f < n; f++) {
        if (out_off) out_off[f] = off[f];
        if (out_len) out_len[f] = len[f];
        if (status) status[f] = st[f];
        if (st[f] == SZB_OK && off[f] + len[f] > total) total = off[f] + len[f];
    }
    if (total_out) *total_out = total;
    szb_batch_destroy(b);
    return rc;
}

int szb_decode_blocks(szb_ctx *ctx, const void *d_src, size_t src_len, const szb_frame_desc *frames, uint32_t nframes,
                      const szb_block_desc *blocks, uint32_t nblocks, void *d_dst, size_t dst_cap, uint64_t *out_off,
                      uint64_t *out_len, int32_t *status) {
    if (!ctx) return SZB_ERR_INVALID_ARGUMENT;
    szb_batch *b = nullptr;
    int rc = szb_batch_create_from_tables(ctx, src_len, frames, nframes, blocks, nblocks, &b);
    if (rc) return rc;
    rc = decode_tables(ctx, b, (const uint8_t *)d_src, src_len, (uint8_t *)d_dst, dst_cap, out_off, out_len, status,
                       SZB_FLAG_SRC_DEVICE | SZB_FLAG_DST_DEVICE);
    szb_batch_destroy(b);
    return rc;
}

// One frame from a reader to a writer (szb200.h).  Host side: the source is read piece by piece into pinned slots (and kept
// in an ordinary buffer for the header walk); every piece goes H2D while the next is being read.  After the decode the output
// comes back through the pinned out-slots, two pieces ahead of the one `write` is consuming.
int szb_decompress_reader(szb_ctx *ctx, szb_read_fn read, void *read_user, szb_write_fn write, void *write_user,
                          uint64_t *in_used, uint64_t *out_total, uint32_t flags) {
    if (!ctx || !read || !write) return SZB_ERR_INVALID_ARGUMENT;
    if (in_used) *in_used = 0;
    if (out_total) *out_total = 0;
    CUDA_TRY(ctx, cudaSetDevice(ctx->device));
    int rc = pin_rings(ctx);
    if (rc) return rc;
    constexpr size_t kPiece = szb_ctx::kPinBytes;
    std::vector<uint8_t> host;
    size_t have = 0;
    // ---- input: read -> pinned slot -> device, the device buffer growing by doubling (its content moves with it) ----
    for (uint32_t k = 0;; k++) {
        const int slot = (int)(k % szb_ctx::kPinIn);
        CUDA_TRY(ctx, cudaEventSynchronize(ctx->pin_in_ev[slot]));
        size_t got = 0;
        bool eof = false;
        while (got < kPiece) {  // a reader may return short counts: fill the piece
            const int64_t n = read(read_user, ctx->pin_in[slot] + got, kPiece - got);
            if (n < 0 || (uint64_t)n > kPiece - got) return SZB_ERR_IO;
            if (n == 0) {
                eof = true;
                break;
            }
            got += (size_t)n;
        }
        if (got) {
            try {
                host.insert(host.end(), ctx->pin_in[slot], ctx->pin_in[slot] + got);
            } catch (...) {
                return SZB_ERR_NOMEM;
            }
            if (have + got + 16 > ctx->d_src_cap || !ctx->d_src) {
                size_t want = ctx->d_src_cap ? ctx->d_src_cap : ((size_t)64 << 20);
                while (want < have + got + 16) want *= 2;
                uint8_t *bigger = nullptr;
                CUDA_TRY(ctx, cudaMalloc((void **)&bigger, want));
                if (have) CUDA_TRY(ctx, cudaMemcpyAsync(bigger, ctx->d_src, have, cudaMemcpyDeviceToDevice, ctx->s_h2d));
                CUDA_TRY(ctx, cudaStreamSynchronize(ctx->s_h2d));
                if (ctx->d_src) CUDA_TRY(ctx, cudaFree(ctx->d_src));
                ctx->d_src = bigger;
                ctx->d_src_cap = want;
            }
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_src + have, ctx->pin_in[slot], got, cudaMemcpyHostToDevice, ctx->s_h2d));
            CUDA_TRY(ctx, cudaEventRecord(ctx->pin_in_ev[slot], ctx->s_h2d));
            have += got;
        }
        if (eof) break;
    }
    // ---- the walk (host bytes), then the four stages once the last piece has arrived ----
    uint64_t off = 0, len = have;
    szb_batch *b = nullptr;
    rc = szb_batch_create(ctx, host.data(), have, &off, &len, 1, &b);
    if (rc) return rc;
    if (in_used) *in_used = b->frames[0].src_len;
    ctx->last_frame = b->frames[0];
    cudaEvent_t up = nullptr;
    uint64_t total = 0;
    int32_t st = 0;
    auto fail = [&](int code) {
        if (up) cudaEventDestroy(up);
        szb_batch_destroy(b);
        return code;
    };
    if (cudaEventCreateWithFlags(&up, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(up, ctx->s_h2d) != cudaSuccess ||
        cudaStreamWaitEvent(ctx->stream, up, 0) != cudaSuccess)
        return fail(SZB_ERR_CUDA);
    if (!ctx->d_src) {  // an empty source: the walk has reported the missing magic number already; nothing to run
        rc = ensure_dev(ctx, &ctx->d_src, &ctx->d_src_cap, 16);
        if (rc) return fail(rc);
    }
    rc = launch_entropy(b, ctx->d_src);
    if (!rc) rc = szb_batch_sizes(b, &total, nullptr, nullptr);
    if (!rc) rc = ensure_dev(ctx, &ctx->d_dst, &ctx->d_dst_cap, (size_t)total + 16);
    if (!rc) rc = launch_execute(b, ctx->d_src, ctx->d_dst, (size_t)total);
    if (!rc && (flags & SZB_FLAG_VERIFY_CHECKSUM)) rc = szb_batch_verify_checksums(b, ctx->d_dst);
    if (!rc) rc = szb_batch_finish(b, &st);  // synchronises: the frame's verdict is known before a byte is handed out
    if (rc) return fail(rc);
    // ---- output: pieces through the pinned out-slots, kPinOut - 1 of them ahead of the writer ----
    const uint64_t pieces = (total + kPiece - 1) / kPiece;
    auto issue = [&](uint64_t p) -> bool {
        const int slot = (int)(p % szb_ctx::kPinOut);
        const size_t n = (size_t)(total - p * kPiece < kPiece ? total - p * kPiece : kPiece);
        return cudaMemcpyAsync(ctx->pin_out[slot], ctx->d_dst + p * kPiece, n, cudaMemcpyDeviceToHost, ctx->s_d2h) == cudaSuccess &&
               cudaEventRecord(ctx->pin_out_ev[slot], ctx->s_d2h) == cudaSuccess;
    };
    for (uint64_t p = 0; p < pieces && p + 1 < (uint64_t)szb_ctx::kPinOut; p++)
        if (!issue(p)) return fail(SZB_ERR_CUDA);
    for (uint64_t p = 0; p < pieces; p++) {
        const int slot = (int)(p % szb_ctx::kPinOut);
        const size_t n = (size_t)(total - p * kPiece < kPiece ? total - p * kPiece : kPiece);
        // the slot of piece p + kPinOut - 1 is the one piece p - 1 has just been written out of
        if (p + szb_ctx::kPinOut - 1 < pieces && !issue(p + szb_ctx::kPinOut - 1)) return fail(SZB_ERR_CUDA);
        if (cudaEventSynchronize(ctx->pin_out_ev[slot]) != cudaSuccess) return fail(SZB_ERR_CUDA);
        if (write(write_user, ctx->pin_out[slot], n) != 0) {
            cudaStreamSynchronize(ctx->s_d2h);
            return fail(SZB_ERR_IO);
        }
    }
    if (out_total) *out_total = total;
    cudaEventDestroy(up);
    szb_batch_destroy(b);
    return SZB_OK;
}

int szb_decompress_frame(szb_ctx *ctx, const uint8_t *src, size_t src_len, uint8_t **out, size_t *out_len,
                         size_t *consumed) {
    if (!ctx || !out || !out_len || (!src && src_len)) return SZB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    *out_len = 0;
    if (consumed) *consumed = 0;
    uint64_t off = 0, len = src_len;
    szb_batch *b = nullptr;
    int rc = szb_batch_create(ctx, src, src_len, &off, &len, 1, &b);
    if (rc) return rc;
    const szb_frame_desc fr = b->frames[0];
    if (consumed) *consumed = (size_t)fr.src_len;
    // entropy first: a frame without Frame_Content_Size only learns its length after stage 3
    rc = ensure_dev(ctx, &ctx->d_src, &ctx->d_src_cap, src_len + 16);
    if (!rc && src_len) {
        cudaError_t e = cudaMemcpyAsync(ctx->d_src, src, src_len, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) {
            ctx->last_error = cudaGetErrorString(e);
            rc = SZB_ERR_CUDA;
        }
    }
    uint64_t total = 0;
    if (!rc) rc = launch_entropy(b, ctx->d_src);
    if (!rc) rc = szb_batch_sizes(b, &total, nullptr, nullptr);
    uint8_t *host = nullptr;
    if (!rc) {
        host = (uint8_t *)malloc(total ? (size_t)total : 1);
        if (!host) rc = SZB_ERR_NOMEM;
    }
    if (!rc) rc = ensure_dev(ctx, &ctx->d_dst, &ctx->d_dst_cap, (size_t)total + 16);
    if (!rc) rc = launch_execute(b, ctx->d_src, ctx->d_dst, (size_t)total);
    int32_t st = 0;
    if (!rc) rc = szb_batch_finish(b, &st);
    if (!rc && total) {
        cudaError_t e = cudaMemcpy(host, ctx->d_dst, (size_t)total, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) {
            ctx->last_error = cudaGetErrorString(e);
            rc = SZB_ERR_CUDA;
        }
    }
    szb_batch_destroy(b);
    if (rc) {
        free(host);
        return rc;
    }
    *out = host;
    *out_len = (size_t)total;
    return SZB_OK;
}

}  // extern "C"
