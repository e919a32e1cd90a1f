// sequences.cuh -- sequences-section table selection and the three-state FSE sequence decode
// (SURVEY.md section 8a rows a13, a14).
//
// replaces: structure/sequences.go:275-369 DecodeTables (Predefined / RLE / FSE / Repeat per
// field, in the order LL, OF, ML), :27-62 RepeatingDecodingTable, :126-206 DecodeSequences and
// :64-123 DecodeSequence.
#pragma once
#include "fse.cuh"

namespace szb {

// Where one field's table comes from, after Repeat has been chased to its origin block.
struct TableSource {
    const uint8_t *p;  // RLE: the code byte; FSE: the description; predefined: unused
    uint32_t avail;    // bytes from p to the end of that block
    uint32_t mode;     // 0 predefined, 1 RLE, 2 FSE-compressed
};

SZB_HD uint32_t field_mode(uint32_t modes_byte, int kind) {  // sequences.go:228-232
    return kind == KIND_LL ? (modes_byte >> 6) & 3 : kind == KIND_OF ? (modes_byte >> 4) & 3 : (modes_byte >> 2) & 3;
}

// Walks the table bytes of ONE block's sequences section (they appear in the order LL, OF, ML,
// sequences.go:278,308,339) up to field `kind` and returns where that field's bytes start.
// Used for the block itself and for the origin block of a Repeat-mode table.  Needs to parse
// the preceding FSE descriptions only for their length.  Serial (one lane); norm is scratch.
SZB_HD int locate_field(const uint8_t *tables, uint32_t avail, uint32_t modes_byte, int kind, int16_t *norm,
                        TableSource *out) {
    uint32_t pos = 0;
    const int order[3] = {KIND_LL, KIND_OF, KIND_ML};
    for (int i = 0; i < 3; i++) {
        int k = order[i];
        uint32_t m = field_mode(modes_byte, k);
        if (k == kind) {
            out->p = tables + pos;
            out->avail = avail - pos;
            out->mode = m;
            if (m == 1 && pos >= avail) return SZB_ERR_UNEXPECTED_EOF;
            return SZB_OK;
        }
        if (m == 1) {
            if (pos >= avail) return SZB_ERR_UNEXPECTED_EOF;
            pos += 1;
        } else if (m == 2) {
            uint32_t nsym, al, used;
            uint32_t max_al = k == KIND_OF ? kMaxALOF : (k == KIND_LL ? kMaxALLL : kMaxALML);
            int rc = fse_read_description(tables + pos, avail - pos, max_al, norm, &nsym, &al, &used);
            if (rc) return rc;
            pos += used;
        }
    }
    return SZB_ERR_INVALID_ARGUMENT;
}

struct SeqStates {
    uint32_t ll, of, ml;
};

// One sequence (DecodeSequence, sequences.go:64-123) plus the state update
// (DecodeSequences, sequences.go:178-194).  Order of bit reads: OF extra, ML extra, LL extra,
// then (unless it is the last sequence) LL state, ML state, OF state.
SZB_HD void decode_one_sequence(RevBits &r, const uint32_t *tll, const uint32_t *tof, const uint32_t *tml, SeqStates &st,
                                bool update, uint32_t *ll_out, uint32_t *ml_out, uint32_t *of_out) {
    const uint32_t eof_ = tof[st.of], ell = tll[st.ll], eml = tml[st.ml];  // peek OF, LL, ML (:67-78)
    const uint32_t ofc = fse_code(eof_);
    rev_refill(r);
    *of_out = (1u << ofc) + rev_read(r, ofc);  // :99-104
    rev_refill(r);
    *ml_out = ml_base(fse_code(eml)) + rev_read(r, fse_extra(eml));  // :106-112
    *ll_out = ll_base(fse_code(ell)) + rev_read(r, fse_extra(ell));  // :114-120
    if (update) {
        rev_refill(r);
        st.ll = fse_baseline(ell) + rev_read(r, fse_nb(ell));
        st.ml = fse_baseline(eml) + rev_read(r, fse_nb(eml));
        st.of = fse_baseline(eof_) + rev_read(r, fse_nb(eof_));
    }
}

}  // namespace szb
