/*
 * szb200.h -- C ABI of the B200-native zstd decode engine (libszb200.so).
 *
 * This is the drop-in boundary for the hot path of KillingSpark/sparkzstd.  The reference
 * has no FFI: its surface is the Go package API.  Each entry point
 * below names the reference interface it replaces (file:line under /root/reference) and is
 * what the Go side binds through cgo (go/).
 *
 * Conventions: int return, 0 = OK, negative = error (codes map 1:1 onto the reference's
 * Err* values); caller-owned buffers; the library never retains a caller pointer after a
 * call returns (cgo rule); one szb_ctx per host thread; no global mutable state; plain
 * pointers and sizes only -- no torch / C++ types in any signature.
 */
#ifndef SZB200_H
#define SZB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- error codes (shared numbering with oracle/szo.h) -------------------------------- */
enum {
    SZB_OK = 0,
    SZB_ERR_WRONG_MAGICNUMBER = -1,          /* decompression/framedecompressor.go:128 ErrWrongMagicnumber */
    SZB_ERR_CORRUPT_SIZES = -2,              /* framedecompressor.go:90 */
    SZB_ERR_OUT_OF_BLOCKS = -3,              /* framedecompressor.go:196 */
    SZB_ERR_ILLEGAL_CONTENT_SIZE_FLAG = -4,  /* structure/frame.go:76 */
    SZB_ERR_ILLEGAL_DICTIONARY_ID_FLAG = -5, /* frame.go:110 */
    SZB_ERR_NOT_ENOUGH_BYTES_FOR_BLOCK_HEADER = -6, /* structure/block.go:28 */
    SZB_ERR_ILLEGAL_BLOCK_TYPE = -7,         /* block.go:29 */
    SZB_ERR_ILLEGAL_BLOCK_SIZE = -8,         /* block.go:30 */
    SZB_ERR_WRONG_JUMPTABLE_BYTES = -9,      /* structure/literals.go:43 */
    SZB_ERR_CORRUPTED_JUMPTABLE = -10,       /* literals.go:44 */
    SZB_ERR_ILLEGAL_LITERAL_SECTION_TYPE = -11,
    SZB_ERR_ILLEGAL_LITERAL_SECTION_SIZE_FORMAT = -12,
    SZB_ERR_WRONG_SIZES_BYTES = -13,
    SZB_ERR_NO_HUFF_TABLE_TO_CARRY_OVER = -14,         /* literals.go:206 */
    SZB_ERR_STREAM_DIDNT_DECODE_TO_RIGHT_LENGTH = -15, /* literals.go:207 */
    SZB_ERR_WRONG_SUM_OF_WEIGHTS = -16,      /* structure/huffman.go:109 */
    SZB_ERR_CORRUPTED_HUFF_TREE = -17,       /* huffman.go:110 */
    SZB_ERR_BAD_PADDING = -18,               /* huffman.go:218, fse/fse.go:303 */
    SZB_ERR_DIDNT_USE_ALL_BITS_TO_DECODE_HUFFMAN = -19, /* huffman.go:219 */
    SZB_ERR_NOT_ALL_BITS_USED = -20,         /* structure/sequences.go:208 */
    SZB_ERR_NO_LL_TABLE_TO_CARRY_OVER = -21, /* sequences.go:271 */
    SZB_ERR_NO_ML_TABLE_TO_CARRY_OVER = -22, /* sequences.go:272 */
    SZB_ERR_NO_OF_TABLE_TO_CARRY_OVER = -23, /* sequences.go:273 */
    SZB_ERR_NOT_ALL_BYTES_USED_WHILE_SEQUENCE_DECODING = -24, /* sequences.go:452 */
    SZB_ERR_DIDNT_READ_ALL_PROBABILITIES = -25, /* fse/fse.go:132 */
    SZB_ERR_NO_SYMBOL_FOR_STATE = -26,       /* fse.go:259 */
    SZB_ERR_CANT_UNWIND = -27,               /* bitstream/bitstream.go:18 */
    SZB_ERR_DIDNT_COPY_ALL_LITERAL_BYTES = -28, /* decompression/sequence_execution.go:11 */
    SZB_ERR_IDX_OUT_OF_BOUNDS = -29,         /* decompression/ringbuffer.go:52 */
    SZB_ERR_CANT_REPEAT_BYTES = -30,         /* ringbuffer.go:189 */
    SZB_ERR_DIDNT_DUMP_ALL = -31,            /* ringbuffer.go:303 */
    SZB_ERR_UNEXPECTED_EOF = -32,            /* io.EOF / io.ErrUnexpectedEOF on truncated input */
    SZB_ERR_PANIC = -33,                     /* input on which the Go reference panics */
    SZB_ERR_NOMEM = -34,
    SZB_ERR_UNSUPPORTED = -35,  /* beyond the engine's limits: FSE accuracy log > 9 (LL/ML/weights) or > 8 (OF),
                                   > 64 FSE symbols, Huffman maxBits > 11, offset code > 31 (zstd spec limits;
                                   the reference does not enforce them) */
    /* engine-only codes */
    SZB_ERR_DST_TOO_SMALL = -64,
    SZB_ERR_CUDA = -65,
    SZB_ERR_INVALID_ARGUMENT = -66,
    SZB_ERR_NO_DEVICE = -67,
    SZB_ERR_CHECKSUM_MISMATCH = -68, /* only when SZB_FLAG_VERIFY_CHECKSUM is set (not a reference behaviour) */
    SZB_ERR_IO = -69,                /* a read / write callback of szb_decompress_reader failed */
    SZB_ERR_WRONG_DICTIONARY = -70   /* the frame names another Dictionary_ID than the dictionary given (szb_decode_batch_dict) */
};

/* replaces: the Go `error` values' Error() strings */
const char *szb_strerror(int code);

/* ---- descriptor tables: the output of the header walk -------------------------------- */
/* The reference walks headers inside FrameDecompressor (framedecompressor.go:130-150 magic,
 * :306-374 frame header via structure/frame.go:23-127, :270-303 block headers via
 * structure/block.go:33-55) and inside the literal / sequence section parsers
 * (structure/literals.go:67-204, structure/sequences.go:228-269).  The Go host keeps that
 * walk and emits these two tables; szb_walk_* is the C++ twin of that walker. */

#define SZB_NONE 0xFFFFFFFFu
#define SZB_BLOCK_TABLES_ONLY 1u
#define SZB_CONTENT_SIZE_UNKNOWN 0xFFFFFFFFFFFFFFFFull

typedef struct szb_frame_desc {
    uint64_t src_off;       /* offset of the frame's magic number inside src */
    uint64_t src_len;       /* bytes from the magic to the end of the last block (the reference
                               leaves the optional 4-byte checksum unread: SURVEY A.1) */
    uint64_t window_size;   /* frame.go:28-36; = content size for single-segment frames */
    uint64_t content_size;  /* frame.go:49-61 (+256 rule applied) or SZB_CONTENT_SIZE_UNKNOWN */
    uint64_t dictionary_id; /* frame.go:38-47; parsed and ignored, like the reference */
    uint32_t first_block;   /* index of the frame's first row in the block table */
    uint32_t nblocks;
    uint32_t checksum;      /* the 4 bytes after the last block when has_checksum (never verified by the reference) */
    int32_t status;         /* 0, or the error the header walk hit (frame is then skipped by the device) */
    uint8_t descriptor;     /* Frame_Header_Descriptor byte */
    uint8_t single_segment; /* frame.go:101-103 */
    uint8_t has_checksum;   /* frame.go:106-108 */
    uint8_t has_content_size;
    uint32_t checksum_valid; /* 1 when has_checksum and the 4 checksum bytes lay inside the frame's extent */
} szb_frame_desc;

typedef struct szb_block_desc {
    uint64_t src_off;      /* offset inside src of the block payload (after the 3-byte header) */
    uint64_t lit_buf_off;  /* byte offset of this block's literals in the device literal scratch */
    uint64_t seq_buf_off;  /* index of this block's first sequence in the device sequence scratch */
    uint32_t block_size;   /* Block_Size field (block.go:41-43); RLE blocks carry 1 payload byte */
    uint32_t frame;        /* owning frame */
    uint32_t lit_regen;    /* literals.go:85-159 RegeneratedSize */
    uint32_t lit_comp;     /* literals.go:85-159 CompressedSize (tree + jump table + streams; raw: = regen; rle: 1) */
    uint32_t nseq;         /* sequences.go:255-269 */
    uint32_t seq_off;      /* offset inside the payload of the sequences section */
    uint32_t huf_origin;   /* block whose payload holds the Huffman tree description this block decodes with:
                              itself for Compressed literals, the most recent such block of the frame for
                              Treeless (literals.go:247-252, carry rule framedecompressor.go:292-294); SZB_NONE */
    uint32_t ll_origin;    /* block whose sequences section defines the LL table: itself unless mode Repeat
                              (sequences.go:292-296; carry rule framedecompressor.go:283-291) */
    uint32_t of_origin;    /* same for offsets (sequences.go:321-325) */
    uint32_t ml_origin;    /* same for match lengths (sequences.go:352-356) */
    uint8_t type;          /* 0 Raw, 1 RLE, 2 Compressed (block.go:15-20) */
    uint8_t last;          /* block.go:38 */
    uint8_t lit_type;      /* 0 Raw, 1 RLE, 2 Compressed, 3 Treeless (literals.go:67-81) */
    uint8_t lit_streams;   /* 1 or 4 */
    uint8_t lit_hdr_bytes; /* 1..5 (literals.go:162-204) */
    uint8_t seq_hdr_bytes; /* bytes of the sequence count (1..3) plus the modes byte when nseq > 0 */
    uint8_t seq_modes;     /* raw Symbol_Compression_Modes byte (sequences.go:228-232) */
    uint8_t flags;         /* SZB_BLOCK_TABLES_ONLY: the block only carries entropy tables (a dictionary's, szb_dict_create):
                              stage 1 builds them, nothing is decoded from it and it regenerates nothing */
    int32_t hdr_status;    /* 0, or the error the walk hit in THIS block's sequences-section header (count, modes, a Repeat mode
                              with nothing to repeat, sizes that do not add up).  The reference decodes a block's literals before
                              it looks at that header (framedecompressor.go:93-126), so the device still decodes the literals of
                              such a block and an error in them wins; the block is the last of its frame's rows and nseq is 0 */
} szb_block_desc;

/* The layout of the two structs above, for bindings that mirror them instead of including this header (the Go structs of
 * go/szb200, the ctypes mirror): sizeof(szb_frame_desc), the offset of each of its fields in declaration order (src_off ..
 * checksum_valid), sizeof(szb_block_desc), the offset of each of its fields (src_off .. seq_modes, flags,
 * hdr_status).  Writes at most cap values, returns how many there are (38).  A binding compares them with its own when it loads. */
uint32_t szb_abi_layout(uint32_t *out, uint32_t cap);

/* Multi-GPU host split: frames are independent, so G GPUs decode a partition of the frame list, one process
 * (one szb_ctx) per GPU, no data-path collective.  weight[i] is frame i's cost (its content size when the header declares it,
 * else a multiple of its compressed size); shard_of[i] receives the shard (0 .. nshards-1) of frame i, shard_load[r]
 * (optional) the summed weight of shard r.  Greedy longest-processing-time binning; deterministic: every rank computes the
 * same partition from the same weights.  The reference decodes one frame per FrameDecompressor on one core
 * (cmd/sparkzstd/main.go:22-40 loops over files); this is the batch counterpart of that loop across devices. */
int szb_shard_frames(const uint64_t *weight, uint32_t nframes, uint32_t nshards, uint32_t *shard_of, uint64_t *shard_load);

typedef struct szb_walk szb_walk;

/* Walks the headers of nframes frames.  frame_off/frame_len give each frame's extent inside
 * src; pass frame_off == NULL (and nframes == 0) to treat src as a concatenation of frames whose
 * boundaries are discovered by the walk (skippable frames are skipped, a present checksum is
 * stepped over) -- the reference itself handles exactly one frame per reader.
 * replaces: CheckMagicnum + DecodeFrameHeader + the DecodeNextBlockHeader loop
 * (framedecompressor.go:130-150, :306-374, :270-303) and the header parts of
 * DecodeNextLiteralsSection / DecodeNextSequenceSection. */
int szb_walk_create(const uint8_t *src, size_t src_len, const uint64_t *frame_off, const uint64_t *frame_len,
                    uint32_t nframes, szb_walk **out);
/* The same for a batch decoded with a dictionary: dict_block (NULL for a raw-content dictionary) becomes row 0 of the block
 * table, owned by a pseudo frame appended after the caller's frames (szb_walk_nframes counts it; it has no blocks and a
 * non-zero status); a frame naming another Dictionary_ID than dict_id ends with SZB_ERR_WRONG_DICTIONARY. */
int szb_walk_create_dict(const uint8_t *src, size_t src_len, const uint64_t *frame_off, const uint64_t *frame_len,
                         uint32_t nframes, const szb_block_desc *dict_block, uint32_t dict_id, szb_walk **out);
void szb_walk_destroy(szb_walk *w);
uint32_t szb_walk_nframes(const szb_walk *w);
uint32_t szb_walk_nblocks(const szb_walk *w);
const szb_frame_desc *szb_walk_frames(const szb_walk *w);
const szb_block_desc *szb_walk_blocks(const szb_walk *w);
uint64_t szb_walk_literal_bytes(const szb_walk *w);  /* device literal scratch needed */
uint64_t szb_walk_sequences(const szb_walk *w);      /* device sequence scratch rows needed */
/* sum of content sizes when every frame declares one, else SZB_CONTENT_SIZE_UNKNOWN */
uint64_t szb_walk_known_output_size(const szb_walk *w);

/* ---- engine context -------------------------------------------------------------------- */
typedef struct szb_ctx szb_ctx;

/* device: CUDA ordinal.  stream: a cudaStream_t to enqueue on, or NULL to let the context
 * create its own.  Fails with SZB_ERR_NO_DEVICE when no CUDA device is usable: there is no
 * CPU fallback.  replaces: NewFrameDecompressor's buffer set-up (framedecompressor.go:55-61). */
int szb_ctx_create(int device, void *stream, szb_ctx **out);
void szb_ctx_destroy(szb_ctx *ctx);
void *szb_ctx_stream(szb_ctx *ctx);
const char *szb_ctx_last_error(szb_ctx *ctx); /* detail text of the last SZB_ERR_CUDA */

#define SZB_FLAG_SRC_DEVICE 1u /* src is a device pointer (descriptor tables still come from a host copy) */
#define SZB_FLAG_DST_DEVICE 2u /* dst is a device pointer; the output stays in HBM */
#define SZB_FLAG_VERIFY_CHECKSUM 4u /* verify the content checksum (XXH64 low 32 bits) on the GPU; SURVEY 8f-1.
                                       NOT a reference behaviour: the reference leaves those 4 bytes unread */

/* The batch entry point north_star asks for: decode nframes independent frames in one
 * launch sequence.  src/dst are HOST buffers unless flagged.  Frames are written back to
 * back into dst; out_off/out_len/status (each nframes long, host) receive every frame's
 * placement and result code.  Returns 0 when every frame decoded, else the first failing
 * frame's code.  replaces: a loop of NewFrameDecompressor(src_i, dst_i).Decompress()
 * (cmd/sparkzstd/main.go:22-40). */
int szb_decode_batch(szb_ctx *ctx, const uint8_t *src, size_t src_len, const uint64_t *frame_off,
                     const uint64_t *frame_len, uint32_t nframes, uint8_t *dst, size_t dst_cap, uint64_t *out_off,
                     uint64_t *out_len, int32_t *status, uint32_t flags);

/* A `.zst` stream as the zstd tools write it: frames back to back, skippable frames (magic 0x184D2A5?) in between,
 * a content checksum after a frame when its header says so.  The frames are discovered by the header walk, decoded as one
 * batch and written back to back into dst (so dst holds the stream's content).  out_off/out_len/status (each max_frames
 * long, host; any may be NULL) receive every frame's placement and result; *nframes_out the number of frames found.
 * SZB_ERR_INVALID_ARGUMENT when the stream holds more than max_frames frames and per-frame arrays were given;
 * trailing bytes that are no frame end the walk with that frame's error (SZB_ERR_WRONG_MAGICNUMBER).
 * NOT a reference behaviour: the reference decodes exactly one frame per reader (framedecompressor.go:130-150) and knows no
 * skippable frames; SURVEY.md section 8f-1. */
int szb_decode_stream(szb_ctx *ctx, const uint8_t *src, size_t src_len, uint8_t *dst, size_t dst_cap, uint64_t *out_off,
                      uint64_t *out_len, int32_t *status, uint32_t max_frames, uint32_t *nframes_out, uint64_t *total_out,
                      uint32_t flags);

/* Descriptor-level entry: what the Go walker feeds.  d_src / d_dst are DEVICE pointers,
 * frames / blocks are HOST tables (copied before return).  replaces: DecodeNextBlockContent +
 * ExecuteSequences + the Raw/RLE bodies of DecodeNextBlock (framedecompressor.go:93-126,
 * :198-244; sequence_execution.go:14-63) for every block of every frame. */
/* DEVICE SOURCE BUFFERS (d_src here, in szb_batch_decode_entropy / szb_batch_execute / szb_batch_run, and src with
 * SZB_FLAG_SRC_DEVICE): the bit readers fetch whole aligned 16-byte chunks, so d_src must be 16-byte aligned and at least 16
 * bytes past src_len must be readable (their content does not matter).  A misaligned pointer is SZB_ERR_INVALID_ARGUMENT; the
 * padding cannot be checked and is the caller's to provide (cudaMalloc'd buffers of src_len + 16 bytes satisfy both).
 * DEVICE DESTINATION BUFFERS (d_dst, and dst with SZB_FLAG_DST_DEVICE): stage 4 re-reads earlier output in whole aligned
 * 16-byte chunks, so d_dst should be 16-byte aligned with 16 readable bytes behind dst_cap (any cudaMalloc'd buffer has them:
 * allocations are padded to 256 bytes). */
int szb_decode_blocks(szb_ctx *ctx, const void *d_src, size_t src_len, const szb_frame_desc *frames, uint32_t nframes,
                      const szb_block_desc *blocks, uint32_t nblocks, void *d_dst, size_t dst_cap, uint64_t *out_off,
                      uint64_t *out_len, int32_t *status);

/* Streaming single-frame entry: what NewFrameDecompressor(source, target).Decompress() and FrameReader.Read do, with the
 * transfers overlapped instead of read-all / decode / write-all.  `read` is source.Read: it fills buf with up to cap bytes and
 * returns how many (0 = end of input, < 0 = error); `write` is target.Write: it takes n decoded bytes and returns 0 (non-zero
 * aborts).  Input pieces are staged through pinned memory and travel to the device while the next piece is being read; the
 * frame is decoded once its last byte is on the device (stage 4 needs every block's size); the output comes back in pieces,
 * piece k+1 and k+2 on the link while `write` consumes piece k -- so a reader sees its first bytes one piece after the
 * decode, not after the whole device-to-host copy.  Like the reference's bufio.Reader the source is read to its end, the
 * frame's own extent is reported through in_used (checksum excluded).  Both callbacks run on the calling thread.
 * replaces: FrameDecompressor.Decompress (framedecompressor.go:153-170) with its io.Reader source and io.Writer target, and
 * the read loop of FrameReader.Read (framereader.go:51-109).  SZB_FLAG_VERIFY_CHECKSUM is honoured. */
typedef int64_t (*szb_read_fn)(void *user, uint8_t *buf, size_t cap);
typedef int (*szb_write_fn)(void *user, const uint8_t *buf, size_t n);
int szb_decompress_reader(szb_ctx *ctx, szb_read_fn read, void *read_user, szb_write_fn write, void *write_user,
                          uint64_t *in_used, uint64_t *out_total, uint32_t flags);
/* The header row (window size, content size, number of blocks, walk status ...) of the frame szb_decompress_reader decoded
 * last on this context: what FrameDecompressor.BlockCounter and the reference's frame-header fields are filled from. */
int szb_ctx_last_frame(szb_ctx *ctx, szb_frame_desc *out);

/* ---- dictionaries (NOT a reference behaviour: the reference parses Dictionary_ID and ignores it, frame.go:38-47,
 * and lists dictionaries as missing, Readme.md:59-61).  RFC 8878 section 5: a raw-content dictionary is history in front of the
 * frame; a formatted one (magic 0xEC30A437) also brings a Huffman table, three FSE tables and three repeat offsets that act
 * as the "previous block" of the frame's first block.  szb_dict_create parses the dictionary on the host and keeps its content
 * and its tables' descriptions in device memory; the tables enter a batch as one table-only row of the block table
 * (SZB_BLOCK_TABLES_ONLY) that Treeless literals and Repeat modes of first blocks name as their origin. */
typedef struct szb_dict szb_dict;
int szb_dict_create(szb_ctx *ctx, const uint8_t *dict, size_t len, szb_dict **out);
void szb_dict_destroy(szb_dict *d);
uint32_t szb_dict_id(const szb_dict *d); /* 0 for a raw-content dictionary */
/* szb_decode_batch with a dictionary: every frame whose header names the dictionary's id, or none, is decoded with it; a frame
 * that names another id ends with SZB_ERR_WRONG_DICTIONARY.  Host buffers only (no SZB_FLAG_SRC_DEVICE / _DST_DEVICE). */
int szb_decode_batch_dict(szb_ctx *ctx, const szb_dict *dict, const uint8_t *src, size_t src_len, const uint64_t *frame_off,
                          const uint64_t *frame_len, uint32_t nframes, uint8_t *dst, size_t dst_cap, uint64_t *out_off,
                          uint64_t *out_len, int32_t *status, uint32_t flags);

/* ---- staged batch object (size-then-decode, resident inputs, timing) ------------------ */
typedef struct szb_batch szb_batch;

/* Walk + upload descriptor tables + size the scratch arenas.  h_src must stay valid only
 * for the duration of the call. */
int szb_batch_create(szb_ctx *ctx, const uint8_t *h_src, size_t src_len, const uint64_t *frame_off,
                     const uint64_t *frame_len, uint32_t nframes, szb_batch **out);
int szb_batch_create_from_tables(szb_ctx *ctx, size_t src_len, const szb_frame_desc *frames, uint32_t nframes,
                                 const szb_block_desc *blocks, uint32_t nblocks, szb_batch **out);
void szb_batch_destroy(szb_batch *b);
uint32_t szb_batch_nframes(const szb_batch *b);
uint32_t szb_batch_nblocks(const szb_batch *b);
/* Stages 1-3 (FSE tables, Huffman literals, FSE sequences, block-offset scan); asynchronous. */
int szb_batch_decode_entropy(szb_batch *b, const void *d_src);
/* Synchronises; total decompressed bytes and per-frame placement (any may be NULL). */
int szb_batch_sizes(szb_batch *b, uint64_t *total, uint64_t *out_off, uint64_t *out_len);
/* Stage 4 (sequence execution / Raw / RLE bodies) into d_dst; asynchronous. */
int szb_batch_execute(szb_batch *b, const void *d_src, void *d_dst, size_t dst_cap);
/* All four stages back to back, no host synchronisation in between. */
int szb_batch_run(szb_batch *b, const void *d_src, void *d_dst, size_t dst_cap);
/* After szb_batch_execute / szb_batch_run: recompute XXH64 of every frame that carries a content checksum and
 * mark mismatches SZB_ERR_CHECKSUM_MISMATCH; asynchronous. */
int szb_batch_verify_checksums(szb_batch *b, void *d_dst);
/* Synchronises and returns per-frame status (nframes long, may be NULL); returns the first failure. */
int szb_batch_finish(szb_batch *b, int32_t *status);
/* Stage-level scratch, for parity tests against the oracle's per-block trace (host copies). */
int szb_batch_read_literals(szb_batch *b, uint32_t block, uint8_t *dst, size_t cap);
int szb_batch_read_sequences(szb_batch *b, uint32_t block, uint32_t *ll, uint32_t *ml, uint32_t *of, size_t cap);
int szb_batch_read_block_results(szb_batch *b, uint64_t *out_size, int32_t *status, size_t cap);

/* Device-event timings of the last szb_batch_run / decode call, milliseconds:
 * [0] total, [1] Huffman literals, [2] FSE tables + sequences, [3] offset scan, [4] execution,
 * [5] H2D, [6] D2H, [7] the FSE table construction share of [2], [8] the share of [4] spent resolving offsets and
 * positions (k_resolve).  [1] and [2] are both measured from the
 * start of the step: the two chains run side by side on two streams.  Returns how many entries were written. */
int szb_last_timing(szb_ctx *ctx, float *ms, int n);
/* Kernel launches issued by the library since the context was created. */
uint64_t szb_launch_count(szb_ctx *ctx);

/* ---- single-frame helpers behind the Go API ------------------------------------------- */
/* replaces: FrameDecompressor.Decompress (framedecompressor.go:153-170) and what
 * FrameReader.Read drains (framereader.go:51-109).  Decodes the one frame at src[0..] and
 * returns a malloc'ed buffer the caller releases with szb_free. */
int szb_decompress_frame(szb_ctx *ctx, const uint8_t *src, size_t src_len, uint8_t **out, size_t *out_len,
                         size_t *consumed);
void szb_free(void *p);

const char *szb_version(void);

#ifdef __cplusplus
}
#endif
#endif
