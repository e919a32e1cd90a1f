/*
 * szo.h -- ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C, CPU-only restatement of the zstd decode path of KillingSpark/sparkzstd
 * (the Go reference under /root/reference).  Every function cites the reference
 * file:line whose behaviour it follows.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library; the product
 * (sparkzstd_b200/) never links, imports or calls it.
 *
 * Parity pinning: this oracle is checked (tests/test_oracle_*.py) against
 *   - the 100 decodecorpus_files golden pairs of the reference (tests/golden/decodecorpus),
 *   - the predefined LL decode table in fse/fse_test.go:8-41,
 *   - the bit-reader KATs in bitstream/reversebitstream_test.go and bitstream_test.go,
 *   - the match-copy KATs in decompression/ringbuffer_test.go:85-154.
 * The Go reference itself cannot be built here (no Go toolchain), so there is no
 * oracle/_ref; see DESIGN.md.
 */
#ifndef SZO_H
#define SZO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes: 0 = OK; the negative values map 1:1 to the reference's Err* values
 * (SURVEY.md A.10).  The numeric values are shared with include/szb200.h. */
enum {
    SZO_OK = 0,
    SZO_ERR_WRONG_MAGICNUMBER = -1,          /* framedecompressor.go:128 */
    SZO_ERR_CORRUPT_SIZES = -2,              /* framedecompressor.go:90  */
    SZO_ERR_OUT_OF_BLOCKS = -3,              /* framedecompressor.go:196 */
    SZO_ERR_ILLEGAL_CONTENT_SIZE_FLAG = -4,  /* frame.go:76 */
    SZO_ERR_ILLEGAL_DICTIONARY_ID_FLAG = -5, /* frame.go:110 */
    SZO_ERR_NOT_ENOUGH_BYTES_FOR_BLOCK_HEADER = -6, /* block.go:28 */
    SZO_ERR_ILLEGAL_BLOCK_TYPE = -7,         /* block.go:29 */
    SZO_ERR_ILLEGAL_BLOCK_SIZE = -8,         /* block.go:30 */
    SZO_ERR_WRONG_JUMPTABLE_BYTES = -9,      /* literals.go:43 */
    SZO_ERR_CORRUPTED_JUMPTABLE = -10,       /* literals.go:44 */
    SZO_ERR_ILLEGAL_LITERAL_SECTION_TYPE = -11,        /* literals.go:64 */
    SZO_ERR_ILLEGAL_LITERAL_SECTION_SIZE_FORMAT = -12, /* literals.go:65 */
    SZO_ERR_WRONG_SIZES_BYTES = -13,         /* literals.go:83 */
    SZO_ERR_NO_HUFF_TABLE_TO_CARRY_OVER = -14,         /* literals.go:206 */
    SZO_ERR_STREAM_DIDNT_DECODE_TO_RIGHT_LENGTH = -15, /* literals.go:207 */
    SZO_ERR_WRONG_SUM_OF_WEIGHTS = -16,      /* huffman.go:109 */
    SZO_ERR_CORRUPTED_HUFF_TREE = -17,       /* huffman.go:110 */
    SZO_ERR_BAD_PADDING = -18,               /* huffman.go:218, fse.go:303 */
    SZO_ERR_DIDNT_USE_ALL_BITS_TO_DECODE_HUFFMAN = -19, /* huffman.go:219 */
    SZO_ERR_NOT_ALL_BITS_USED = -20,         /* sequences.go:208 */
    SZO_ERR_NO_LL_TABLE_TO_CARRY_OVER = -21, /* sequences.go:271 */
    SZO_ERR_NO_ML_TABLE_TO_CARRY_OVER = -22, /* sequences.go:272 */
    SZO_ERR_NO_OF_TABLE_TO_CARRY_OVER = -23, /* sequences.go:273 */
    SZO_ERR_NOT_ALL_BYTES_USED_WHILE_SEQUENCE_DECODING = -24, /* sequences.go:452 */
    SZO_ERR_DIDNT_READ_ALL_PROBABILITIES = -25, /* fse.go:132 */
    SZO_ERR_NO_SYMBOL_FOR_STATE = -26,       /* fse.go:259 */
    SZO_ERR_CANT_UNWIND = -27,               /* bitstream.go:18 */
    SZO_ERR_DIDNT_COPY_ALL_LITERAL_BYTES = -28, /* sequence_execution.go:11 */
    SZO_ERR_IDX_OUT_OF_BOUNDS = -29,         /* ringbuffer.go:52 */
    SZO_ERR_CANT_REPEAT_BYTES = -30,         /* ringbuffer.go:189 */
    SZO_ERR_DIDNT_DUMP_ALL = -31,            /* ringbuffer.go:303 */
    SZO_ERR_UNEXPECTED_EOF = -32,            /* io.EOF / io.ErrUnexpectedEOF on truncated input */
    SZO_ERR_PANIC = -33,                     /* input on which the Go reference would panic (index out of range, "can't happen") */
    SZO_ERR_NOMEM = -34,
    SZO_ERR_UNSUPPORTED = -35                /* beyond the limits the GPU engine enforces (never returned by the oracle itself) */
};

const char *szo_strerror(int code);

/* ---- a1: reverse bit reader (bitstream/reversebitstream.go:3-88) ---- */
typedef struct {
    const uint8_t *data;
    int64_t len;
    int64_t offset; /* == remaining bits - 1, may go below -1 (over-read) */
} szo_rbits;

void szo_rbits_init(szo_rbits *r, const uint8_t *data, size_t len);
int64_t szo_rbits_bits_still_in_stream(const szo_rbits *r);
uint64_t szo_rbits_read(szo_rbits *r, int n);

/* ---- a2: forward bit reader (bitstream/bitstream.go:8-90) ---- */
typedef struct {
    const uint8_t *data;
    size_t len;
    size_t pos;      /* next byte to fetch from the source */
    uint8_t buffer;
    unsigned offset; /* bits of buffer already consumed; 8 = empty */
} szo_fbits;

void szo_fbits_init(szo_fbits *b, const uint8_t *data, size_t len);
int szo_fbits_read(szo_fbits *b, int n, uint64_t *out);
int szo_fbits_unwind_bit(szo_fbits *b);

/* ---- a3..a7: FSE (fse/fse.go, fse/predefined.go) ---- */
#define SZO_FSE_MAX_SYMBOLS 512
typedef struct {
    uint16_t baseline;
    uint8_t additional_bits;
    uint8_t number_of_bits;
    int32_t symbol;
} szo_fse_entry;

typedef struct {
    int accuracy_log;
    int nvalues;                         /* len(Values) */
    int64_t values[SZO_FSE_MAX_SYMBOLS]; /* probability + 1 */
    szo_fse_entry *table;                /* 1 << accuracy_log cells (malloc'ed) */
    int table_size;
    int64_t state;
    /* RepeatingDecodingTable (sequences.go:27-62) is modelled as is_rle = 1 */
    int is_rle;
    int rle_value;
    int rle_additional_bits;
} szo_fse_table;

void szo_fse_table_free(szo_fse_table *t);
int szo_fse_read_table_description(szo_fse_table *t, const uint8_t *src, size_t len, int *bytes_read);
int szo_fse_build_decoding_table(szo_fse_table *t, const int *symbol_translation, int ntrans,
                                 const uint8_t *extra_bits, int nextra);
uint32_t szo_highbit32(uint32_t v);
int szo_fse_build_ll_table(szo_fse_table *t);
int szo_fse_build_ml_table(szo_fse_table *t);
int szo_fse_build_of_table(szo_fse_table *t);
int szo_fse_decode_interleaved(szo_fse_table *t1, szo_fse_table *t2, const uint8_t *src, size_t len,
                               uint8_t *out, int out_cap, int *nout);

extern const int szo_ll_base[36];
extern const uint8_t szo_ll_extra[36];
extern const int szo_ll_default[36];
extern const int szo_ml_base[53];
extern const uint8_t szo_ml_extra[53];
extern const int szo_ml_default[53];
extern const int szo_of_default[29];

/* ---- a9..a11: Huffman (structure/huffman.go) ---- */
typedef struct {
    int max_bits;
    int size;
    int *number_of_bits;
    int *symbols;
} szo_huf_table;

void szo_huf_table_free(szo_huf_table *t);
int szo_huf_decode_tree_desc(const uint8_t *src, size_t len, uint8_t *weights, int *nweights, int *bytes_used);
int szo_huf_build(const uint8_t *weights, int nweights, szo_huf_table *out);
int szo_huf_decode_stream(const szo_huf_table *t, const uint8_t *data, size_t len, uint8_t *out,
                          size_t out_cap, int *nout);

/* ---- a15..a17: execution ---- */
typedef struct {
    int64_t match_length;
    int64_t literal_length;
    int64_t offset; /* raw offset_value */
} szo_sequence;

int64_t szo_next_offset(int64_t hist[3], const szo_sequence *seq);
/* flat-buffer statement of Ringbuffer.RepeatBeforeIndex (ringbuffer.go:242-277):
 * append n bytes copied byte-serially from out[pos-oldest...] */
int szo_match_copy(uint8_t *out, size_t *pos, size_t cap, int64_t n, int64_t oldest);

/* ---- whole frame, with optional stage-level trace ---- */
typedef struct {
    int type;      /* 0 raw, 1 rle, 2 compressed */
    int last;
    uint32_t block_size;
    uint64_t out_off; /* offset of this block's output inside the frame output */
    uint64_t out_len;
    /* compressed blocks only */
    int lit_type;  /* 0 raw 1 rle 2 compressed 3 treeless */
    int lit_streams;
    uint32_t lit_regen;
    uint64_t lit_off;  /* into trace->literals (RLE literals are stored expanded) */
    uint32_t nseq;
    uint64_t seq_off;  /* into trace->sequences */
    int ll_mode, of_mode, ml_mode;
    int huf_max_bits;
    int64_t hist_after[3];
} szo_block_trace;

typedef struct {
    szo_block_trace *blocks;
    size_t nblocks, blocks_cap;
    uint8_t *literals;
    size_t nliterals, literals_cap;
    szo_sequence *sequences;
    int64_t *real_offsets; /* resolved offsets, same indexing as sequences */
    size_t nsequences, sequences_cap;
    uint64_t window_size;
    uint64_t frame_content_size;
    int has_fcs;
    int single_segment;
    size_t bytes_consumed; /* compressed bytes read, excluding the (unread) checksum */
} szo_trace;

void szo_trace_free(szo_trace *t);

/* Decode ONE frame starting at src[0].  *out is malloc'ed (caller frees).  Mirrors
 * FrameDecompressor.Decompress (framedecompressor.go:153-170). */
int szo_decode_frame(const uint8_t *src, size_t len, uint8_t **out, size_t *out_len, szo_trace *trace);
/* The same with a dictionary (raw content, or formatted: magic 0xEC30A437).  NOT a reference behaviour (the reference has no
 * dictionary support): restates RFC 8878 section 5; pinned against frames libzstd compressed with the dictionary. */
int szo_decode_frame_dict(const uint8_t *src, size_t len, const uint8_t *dict, size_t dict_len, uint8_t **out, size_t *out_len,
                          szo_trace *trace);

/* Convenience for the CPU baseline: decode nframes independent frames with nthreads
 * POSIX threads (one decoder per thread, as BASELINE.md section 3 prescribes).  dst may be NULL
 * (output discarded) or an arena; out_off/out_len are filled when non-NULL. */
int szo_decode_batch_mt(const uint8_t *src, const uint64_t *frame_off, const uint64_t *frame_len,
                        uint32_t nframes, uint8_t *dst, size_t dst_cap, const uint64_t *dst_off,
                        uint64_t *out_len, int32_t *status, int nthreads);

#ifdef __cplusplus
}
#endif
#endif
