package structure

// No cgo in this package: the descriptor rows are the plain Go structs of package szb200 (exported fields, layout checked
// against the C header when that package loads), so nothing here depends on C types.

import (
	"io"

	"github.com/killingspark/sparkzstd/szb200"
)

func setBlock(d *szb200.BlockDesc, payloadOff int, size uint32, btype byte, last bool) {
	d.SrcOff = uint64(payloadOff)
	d.BlockSize = uint32(size)
	d.Type = uint8(btype)
	if last {
		d.Last = 1
	}
	d.HufOrigin, d.LLOrigin, d.OFOrigin, d.MLOrigin = none, none, none, none
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
	d.LitType, d.LitStreams, d.LitHdrBytes = uint8(litType), uint8(streams), uint8(need)
	d.LitRegen, d.LitComp = uint32(regen), uint32(comp)
	if litType == 3 {
		if *carryHuf == none {
			return ErrNoHuffTableToCarryOver
		}
		d.HufOrigin = uint32(*carryHuf)
	} else if litType == 2 {
		d.HufOrigin = uint32(self)
	}
	litTotal := need + int(comp)
	if litTotal > len(b) {
		return io.ErrUnexpectedEOF
	}
	d.SeqOff = uint32(litTotal)
	// From here on the reference has already decoded the block's literals (DecodeNextBlockContent,
	// framedecompressor.go:93-126: literals first, then the sequences header): an error in this header must not hide an error
	// in the literals.  The block stays in the table -- without sequences, the error in HdrStatus -- so that the device decodes
	// its literals and the first error wins; the caller ends the walk (walker.cpp does the same).
	seqHeader := func() error {
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
			d.SeqHdrBytes = 1
			if litTotal+1 != len(b) {
				return ErrCorruptSizes
			}
		} else if nseq == 0 {
			// A two- or three-byte count that says zero (0x80 0x00): the reference (sequences.go:395-400 short-circuits on the
			// first byte only) goes on to decode tables and a stream of no sequences.  Both walkers agree on ONE answer instead:
			// the section is not the single zero byte, so the sizes do not add up (walker.cpp; DESIGN.md section 2).  The carry
			// origins are left alone: a block without sequences defines no tables.
			return ErrCorruptSizes
		} else {
			if len(s) < nb+1 {
				return io.ErrUnexpectedEOF
			}
			d.NSeq = uint32(nseq)
			d.SeqModes = uint8(s[nb])
			d.SeqHdrBytes = uint8(nb + 1)
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
			d.LLOrigin, d.OFOrigin, d.MLOrigin = uint32(ll), uint32(of), uint32(ml)
			*carryLL, *carryOF, *carryML = ll, of, ml
		}
		return nil
	}
	var hdrErr error
	if hdrErr = seqHeader(); hdrErr != nil {
		d.HdrStatus = CodeForError(hdrErr)
		d.NSeq, d.SeqHdrBytes, d.SeqModes = 0, 0, 0
		d.LLOrigin, d.OFOrigin, d.MLOrigin = none, none, none
	}
	if litType >= 2 {
		*carryHuf = uint32(d.HufOrigin)
		d.LitBufOff = uint64(*litBytes)
		*litBytes += (uint64(regen) + 15) &^ 15
	}
	d.SeqBufOff = uint64(*seqs)
	*seqs += (uint64(d.NSeq) + 31) &^ 31
	if hdrErr != nil {
		return &headerError{hdrErr}
	}
	return nil
}

// headerError: the block belongs in the table (its literals are decoded), the walk ends with Err.
type headerError struct{ Err error }

func (e *headerError) Error() string { return e.Err.Error() }
func (e *headerError) Unwrap() error { return e.Err }

// CodeForError is the szb200 status code of the errors a walk can end with (the inverse of ErrorForCode).
func CodeForError(err error) int32 {
	switch err {
	case ErrIllegalBlockType:
		return -7
	case ErrIllegalBlockSize:
		return -8
	case ErrNoHuffTableToCarryOver:
		return -14
	case ErrNoLLTableToCarryOver:
		return -21
	case ErrNoMLTableToCarryOver:
		return -22
	case ErrNoOFTableToCarryOver:
		return -23
	case ErrCorruptSizes:
		return -2
	case ErrWrongMagicnumber:
		return -1
	case io.ErrUnexpectedEOF, io.EOF:
		return -32
	case errPanic:
		return -33
	}
	return -33
}
