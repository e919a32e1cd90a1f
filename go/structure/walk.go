// Package structure keeps the container-level parsing of the reference (structure/frame.go,
// structure/block.go and the header halves of structure/literals.go and structure/sequences.go)
// and turns it into the two descriptor tables the GPU consumes (include/szb200.h).  This file is
// the Go twin of sparkzstd_b200/csrc/walker.cpp; both must produce identical tables.
//
// NOTE: not compiled in this image (no Go toolchain); see INTEGRATION.md.
package structure

import (
	"errors"
	"io"

	"github.com/killingspark/sparkzstd/szb200"
)

type BlockType byte

const (
	BlockTypeRaw        = BlockType(0)
	BlockTypeRLE        = BlockType(1)
	BlockTypeCompressed = BlockType(2)
	BlockTypeReserved   = BlockType(3)
)

type BlockHeader struct {
	LastBlock bool
	Type      BlockType
	BlockSize uint64
}

type Block struct{ Header BlockHeader }

var (
	ErrIllegalBlockType         = errors.New("Illegal BlockType. Must be smaller than 3.")
	ErrIllegalBlockSize         = errors.New("Illegal block-size. Must be lower than 128kb")
	ErrNoHuffTableToCarryOver   = errors.New("No previous Huffmantree available")
	ErrNoLLTableToCarryOver     = errors.New("Needed to copy old LiteralLenghts table but there was none")
	ErrNoMLTableToCarryOver     = errors.New("Needed to copy old MathcLenghts table but there was none")
	ErrNoOFTableToCarryOver     = errors.New("Needed to copy old Offsets table but there was none")
	ErrWrongMagicnumber         = errors.New("Magicnum is not correct")
	ErrCorruptSizes             = errors.New("The sizes of literal and sequence section did not add up to blocksize")
	errPanic                    = errors.New("corrupt input (the reference decoder panics on it)")
)

const none = 0xFFFFFFFF

// Walk is the header walk of one frame: the szb_frame_desc row, its szb_block_desc rows, and the
// error the walk stopped at (nil when the last block was reached).
type Walk struct {
	Frame  szb200.FrameDesc
	Blocks []szb200.BlockDesc
	Err    error
}

// ErrorForCode maps a device status code back to the identity-comparable error values.
func ErrorForCode(code int) (error, bool) {
	switch code {
	case -7:
		return ErrIllegalBlockType, true
	case -8:
		return ErrIllegalBlockSize, true
	case -14:
		return ErrNoHuffTableToCarryOver, true
	case -21:
		return ErrNoLLTableToCarryOver, true
	case -22:
		return ErrNoMLTableToCarryOver, true
	case -23:
		return ErrNoOFTableToCarryOver, true
	}
	return nil, false
}

func BlockFromDesc(d *szb200.BlockDesc) Block {
	return Block{Header: BlockHeader{LastBlock: d.Last != 0, Type: BlockType(d.Type), BlockSize: uint64(d.BlockSize)}}
}

// WalkFrame walks the frame at src[0:] (framedecompressor.go:130-150, :306-374, :270-303;
// frame.go:23-127; block.go:33-55; literals.go:67-204; sequences.go:228-269).  Mirrors
// walk_frame() in walker.cpp line for line.
func WalkFrame(src []byte) (*Walk, error) {
	w := &Walk{}
	pos := 0
	need := func(n int) bool { return len(src)-pos >= n }
	if !need(4) {
		return nil, io.ErrUnexpectedEOF
	}
	if src[0] != 0x28 || src[1] != 0xB5 || src[2] != 0x2F || src[3] != 0xFD {
		return nil, ErrWrongMagicnumber
	}
	pos = 4
	if !need(1) {
		return nil, io.ErrUnexpectedEOF
	}
	fhd := src[pos]
	pos++
	single := (fhd>>5)&1 == 1
	dictBytes := int(fhd & 3)
	if dictBytes == 3 {
		dictBytes = 4
	}
	fcsBytes := 0
	switch fhd >> 6 {
	case 0:
		if single {
			fcsBytes = 1
		}
	case 1:
		fcsBytes = 2
	case 2:
		fcsBytes = 4
	case 3:
		fcsBytes = 8
	}
	hdr := dictBytes + fcsBytes
	if !single {
		hdr++
	}
	if !need(hdr) {
		return nil, io.ErrUnexpectedEOF
	}
	f := &w.Frame
	f.Descriptor = uint8(fhd)
	f.ContentSize = ^uint64(0)
	if single {
		f.SingleSegment = 1
	} else {
		wd := src[pos]
		pos++
		base := uint64(1) << (10 + (wd >> 3))
		f.WindowSize = uint64(base + (base/8)*uint64(wd&7))
	}
	pos += dictBytes
	if fcsBytes > 0 {
		var v uint64
		for i := 0; i < fcsBytes; i++ {
			v |= uint64(src[pos+i]) << (8 * uint(i))
		}
		if fcsBytes == 2 {
			v += 256
		}
		pos += fcsBytes
		f.ContentSize = uint64(v)
		f.HasContentSize = 1
		if single {
			f.WindowSize = uint64(v)
		}
	}
	carryHuf, carryLL, carryOF, carryML := uint32(none), uint32(none), uint32(none), uint32(none)
	var litBytes, seqs uint64
	last := false
	for !last {
		if !need(3) {
			w.Err = io.ErrUnexpectedEOF
			break
		}
		h := src[pos : pos+3]
		pos += 3
		var d szb200.BlockDesc
		last = h[0]&1 == 1
		btype := (h[0] >> 1) & 3
		size := uint32(h[0]>>3) + uint32(h[1])<<5 + uint32(h[2])<<13
		if btype >= 3 {
			w.Err = ErrIllegalBlockType
			break
		}
		if size > 128*1024 {
			w.Err = ErrIllegalBlockSize
			break
		}
		self := uint32(len(w.Blocks))
		setBlock(&d, pos, size, btype, last)
		switch btype {
		case 0:
			if !need(int(size)) {
				w.Err = io.ErrUnexpectedEOF
			}
			pos += int(size)
		case 1:
			if !need(1) {
				w.Err = io.ErrUnexpectedEOF
			}
			pos++
		case 2:
			if !need(int(size)) {
				w.Err = io.ErrUnexpectedEOF
				break
			}
			if err := parseCompressed(&d, src[pos:pos+int(size)], self, &carryHuf, &carryLL, &carryOF, &carryML, &litBytes, &seqs); err != nil {
				var he *headerError
				if errors.As(err, &he) { // the block's literals are still decoded: it is the frame's last row
					w.Blocks = append(w.Blocks, d)
					w.Err = he.Err
				} else {
					w.Err = err
				}
				break
			}
			pos += int(size)
		}
		if w.Err != nil {
			break
		}
		w.Blocks = append(w.Blocks, d)
	}
	f.SrcLen = uint64(pos)
	f.NBlocks = uint32(len(w.Blocks))
	// The reference leaves the optional content checksum unread (frame.go:105-108).  It is recorded so that
	// szb200.FlagVerifyChecksum can have the GPU check it; src_len stays "up to the end of the last block".
	if (fhd>>2)&1 == 1 {
		f.HasChecksum = 1
		if w.Err == nil && need(4) {
			f.Checksum = uint32(uint32(src[pos]) | uint32(src[pos+1])<<8 | uint32(src[pos+2])<<16 | uint32(src[pos+3])<<24)
			f.ChecksumValid = 1
		}
	}
	return w, nil
}
