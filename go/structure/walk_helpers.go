package structure

/*
#include "../../include/szb200.h"
*/
import "C"

import (
	"io"

	"github.com/killingspark/sparkzstd/szb200"
)

type C_uint8 = C.uint8_t
type C_uint32 = C.uint32_t
type C_uint64 = C.uint64_t

func setBlock(d *szb200.BlockDesc, payloadOff int, size uint32, btype byte, last bool) {
	d.src_off = C.uint64_t(payloadOff)
	d.block_size = C.uint32_t(size)
	d._type = C.uint8_t(btype)
	if last {
		d.last = 1
	}
	d.huf_origin, d.ll_origin, d.of_origin, d.ml_origin = none, none, none, none
}

// parseCompressed fills the literals / sequences header fields and resolves the origin blocks of
// Treeless literals and Repeat-mode tables (literals.go:85-159,247-252; sequences.go:228-269,
// 292-296,321-325,352-356; carry rules framedecompressor.go:283-294).
func parseCompressed(d *szb200.BlockDesc, b []byte, self uint32, carryHuf, carryLL, carryOF, carryML *uint32, litBytes, seqs *uint64) error {
	if len(b) < 1 {
		return io.ErrUnexpectedEOF
	}
	litType := b[0] & 3
	sf := (b[0] >> 2) & 3
	need := 3
	if litType <= 1 {
		need = 1
		if sf == 1 {
			need = 2
		} else if sf == 3 {
			need = 3
		}
	} else if sf == 2 {
		need = 4
	} else if sf == 3 {
		need = 5
	}
	if len(b) < need {
		return io.ErrUnexpectedEOF
	}
	var regen, comp uint32
	streams := uint8(1)
	if litType <= 1 {
		switch sf {
		case 0, 2:
			regen = uint32(b[0] >> 3)
		case 1:
			regen = uint32(b[0]>>4) + uint32(b[1])<<4
		default:
			regen = uint32(b[0]>>4) + uint32(b[1])<<4 + uint32(b[2])<<12
		}
		comp = regen
		if litType == 1 {
			comp = 1
		}
	} else {
		v := uint32(b[0]) | uint32(b[1])<<8 | uint32(b[2])<<16
		if need > 3 {
			v |= uint32(b[3]) << 24
		}
		v >>= 4
		streams = 4
		switch sf {
		case 0, 1:
			if sf == 0 {
				streams = 1
			}
			regen, comp = v&0x3FF, (v>>10)&0x3FF
		case 2:
			regen, comp = v&0x3FFF, (v>>14)&0x3FFF
		default:
			regen, comp = v&0x3FFFF, ((v>>18)&0x3FFFF)+uint32(b[4])<<10
		}
	}
	if regen > 128*1024 || comp > 128*1024 {
		return errPanic
	}
	d.lit_type, d.lit_streams, d.lit_hdr_bytes = C.uint8_t(litType), C.uint8_t(streams), C.uint8_t(need)
	d.lit_regen, d.lit_comp = C.uint32_t(regen), C.uint32_t(comp)
	if litType == 3 {
		if *carryHuf == none {
			return ErrNoHuffTableToCarryOver
		}
		d.huf_origin = C.uint32_t(*carryHuf)
	} else if litType == 2 {
		d.huf_origin = C.uint32_t(self)
	}
	litTotal := need + int(comp)
	if litTotal > len(b) {
		return io.ErrUnexpectedEOF
	}
	d.seq_off = C.uint32_t(litTotal)
	s := b[litTotal:]
	if len(s) < 1 {
		return io.ErrUnexpectedEOF
	}
	nb := 1
	if s[0] >= 128 {
		nb = 2
	}
	if s[0] == 255 {
		nb = 3
	}
	if len(s) < nb {
		return io.ErrUnexpectedEOF
	}
	var nseq uint32
	switch {
	case s[0] < 128:
		nseq = uint32(s[0])
	case s[0] < 255:
		nseq = uint32(s[0]-128)<<8 + uint32(s[1])
	default:
		nseq = uint32(s[1]) + uint32(s[2])<<8 + 0x7F00
	}
	if s[0] == 0 {
		d.seq_hdr_bytes = 1
		if litTotal+1 != len(b) {
			return ErrCorruptSizes
		}
	} else {
		if len(s) < nb+1 {
			return io.ErrUnexpectedEOF
		}
		d.nseq = C.uint32_t(nseq)
		d.seq_modes = C.uint8_t(s[nb])
		d.seq_hdr_bytes = C.uint8_t(nb + 1)
		modes := s[nb]
		pick := func(mode byte, carry *uint32, missing error) (uint32, error) {
			if mode == 3 {
				if *carry == none {
					return 0, missing
				}
				return *carry, nil
			}
			return self, nil
		}
		ll, err := pick(modes>>6, carryLL, ErrNoLLTableToCarryOver)
		if err != nil {
			return err
		}
		of, err := pick((modes>>4)&3, carryOF, ErrNoOFTableToCarryOver)
		if err != nil {
			return err
		}
		ml, err := pick((modes>>2)&3, carryML, ErrNoMLTableToCarryOver)
		if err != nil {
			return err
		}
		d.ll_origin, d.of_origin, d.ml_origin = C.uint32_t(ll), C.uint32_t(of), C.uint32_t(ml)
		*carryLL, *carryOF, *carryML = ll, of, ml
	}
	if litType >= 2 {
		*carryHuf = uint32(d.huf_origin)
		d.lit_buf_off = C.uint64_t(*litBytes)
		*litBytes += (uint64(regen) + 15) &^ 15
	}
	d.seq_buf_off = C.uint64_t(*seqs)
	*seqs += (uint64(d.nseq) + 31) &^ 31
	return nil
}
