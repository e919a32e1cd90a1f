// Package decompression keeps sparkzstd's public API (NewFrameReader, FrameDecompressor) and runs
// the hot path -- FSE tables, Huffman literals, FSE sequences, sequence execution -- on a B200
// through package szb200.  What stays in Go is what the reference does in
// decompression/framedecompressor.go:130-150 (magic), :306-374 (frame header) and :270-303 (block
// headers): the header walk, now emitting descriptor tables (package structure) instead of
// decoding block payloads in place.
//
// NOTE: not compiled in this image (no Go toolchain); see INTEGRATION.md.
package decompression

import (
	"bufio"
	"errors"
	"io"

	"github.com/killingspark/sparkzstd/structure"
	"github.com/killingspark/sparkzstd/szb200"
)

// the reference's exported error values keep their identity
var (
	ErrWrongMagicnumber = errors.New("Magicnum is not correct")
	ErrCorruptSizes     = errors.New("The sizes of literal and sequence section did not add up to blocksize")
	ErrOutOfBlocks      = errors.New("No blocks left in frame")
)

// FrameDecompressor: source io.Reader -> target io.Writer, one frame (framedecompressor.go:14-40).
type FrameDecompressor struct {
	source *bufio.Reader
	target io.Writer
	ctx    *szb200.Ctx

	data      []byte // the whole compressed frame: the GPU decodes frames, not blocks
	out       []byte
	decodeErr error // what decodeFrame found; sticky, like a failed reference decoder
	written   int
	walk      *structure.Walk

	CurrentBlock  structure.Block
	PreviousBlock structure.Block
	BlockCounter  int
	Verbose       bool
}

// NewFrameDecompressor mirrors framedecompressor.go:55-61.
func NewFrameDecompressor(s io.Reader, t io.Writer) *FrameDecompressor {
	fd := &FrameDecompressor{target: t}
	if s != nil {
		fd.source = bufio.NewReader(s)
	}
	return fd
}

// Reset mirrors framedecompressor.go:42-52: a decoder is reused for the next frame.
func (fd *FrameDecompressor) Reset(newsource io.Reader, newtarget io.Writer) {
	fd.source = bufio.NewReader(newsource)
	fd.target = newtarget
	fd.data, fd.out, fd.walk, fd.decodeErr = nil, nil, nil, nil
	fd.written = 0
	fd.CurrentBlock = structure.Block{}
	fd.PreviousBlock = structure.Block{}
	fd.BlockCounter = 0
}

func (fd *FrameDecompressor) load() error {
	if fd.data != nil {
		return nil
	}
	b, err := io.ReadAll(fd.source) // like the reference's bufio.Reader, this reads ahead of the frame end
	if err != nil {
		return err
	}
	fd.data = b
	return nil
}

// CheckMagicnum mirrors framedecompressor.go:130-150.
func (fd *FrameDecompressor) CheckMagicnum() error {
	if err := fd.load(); err != nil {
		return err
	}
	if len(fd.data) < 4 {
		return io.ErrUnexpectedEOF
	}
	if fd.data[0] != 0x28 || fd.data[1] != 0xB5 || fd.data[2] != 0x2F || fd.data[3] != 0xFD {
		return ErrWrongMagicnumber
	}
	return nil
}

// DecodeFrameHeader mirrors framedecompressor.go:306-374; it walks every block header of the
// frame too, because the GPU wants the whole block-descriptor table up front.
func (fd *FrameDecompressor) DecodeFrameHeader() error {
	if err := fd.load(); err != nil {
		return err
	}
	w, err := structure.WalkFrame(fd.data)
	if err != nil {
		return err
	}
	fd.walk = w
	return nil
}

// DecodeNextBlockHeader mirrors framedecompressor.go:270-303.
func (fd *FrameDecompressor) DecodeNextBlockHeader() error {
	if fd.BlockCounter >= len(fd.walk.Blocks) {
		if fd.walk.Err != nil {
			return fd.walk.Err
		}
		return io.ErrUnexpectedEOF
	}
	fd.PreviousBlock = fd.CurrentBlock
	fd.CurrentBlock = structure.BlockFromDesc(&fd.walk.Blocks[fd.BlockCounter])
	return nil
}

// decodeFrame runs the four GPU stages over the whole frame, once.  The device decodes frames, not blocks: the entropy
// stages of all blocks run side by side and stage 4 needs every block's size before it can place the first byte.
func (fd *FrameDecompressor) decodeFrame() error {
	if fd.out != nil || fd.decodeErr != nil {
		return fd.decodeErr
	}
	if fd.ctx == nil {
		c, err := szb200.NewCtx(0)
		if err != nil {
			return err
		}
		fd.ctx = c
	}
	out, _, err := fd.ctx.DecompressFrame(fd.data)
	if err != nil {
		fd.decodeErr = translate(err)
		return fd.decodeErr
	}
	if out == nil {
		out = []byte{}
	}
	fd.out = out
	return nil
}

// DecodeNextBlockContent mirrors framedecompressor.go:93-126: the literals and sequences sections of the CURRENT block
// (a Compressed block whose header DecodeNextBlockHeader has read).  The section headers were walked with the frame
// (structure.WalkFrame), so a block whose sections do not add up to Block_Size has already failed with ErrCorruptSizes in
// DecodeFrameHeader -- earlier than the reference reports it, with the same error value.  The entropy decode itself happens
// for all blocks at once on the first call; later calls return what that decode found.
func (fd *FrameDecompressor) DecodeNextBlockContent() error {
	if fd.walk == nil {
		return io.ErrUnexpectedEOF
	}
	if fd.CurrentBlock.Header.Type != structure.BlockTypeCompressed {
		return nil // the reference only calls it for Compressed blocks (framedecompressor.go:219-222)
	}
	return fd.decodeFrame()
}

// ExecuteSequences mirrors sequence_execution.go:14-63 for a caller that steps through a frame itself: after it returns nil
// the CURRENT block's output exists.  Stage 4 of every block ran inside decodeFrame (one warp per frame walks the blocks in
// order, exactly the reference's order), so per block there is nothing left to do but report that decode's verdict: an
// error found while executing (ErrDidntCopyAllLiteralBytes, a match reaching before the frame) is returned here as the
// reference returns it.
func (fd *FrameDecompressor) ExecuteSequences() error {
	return fd.decodeFrame()
}

// DecodeNextBlock mirrors framedecompressor.go:198-244: header, then for a Compressed block DecodeNextBlockContent +
// ExecuteSequences; Raw and RLE bodies are written by the same GPU pass (k_execute_bodies).
func (fd *FrameDecompressor) DecodeNextBlock() error {
	if fd.CurrentBlock.Header.LastBlock {
		return ErrOutOfBlocks
	}
	if err := fd.DecodeNextBlockHeader(); err != nil {
		return err
	}
	switch fd.CurrentBlock.Header.Type {
	case structure.BlockTypeCompressed:
		if err := fd.DecodeNextBlockContent(); err != nil {
			return err
		}
		return fd.ExecuteSequences()
	default:
		return fd.decodeFrame()
	}
}

// Decompress mirrors framedecompressor.go:153-170.
func (fd *FrameDecompressor) Decompress() error {
	if err := fd.CheckMagicnum(); err != nil {
		return err
	}
	if err := fd.DecodeFrameHeader(); err != nil {
		return err
	}
	for !fd.CurrentBlock.Header.LastBlock {
		if err := fd.DecodeNextBlock(); err != nil {
			return err
		}
		fd.BlockCounter++
	}
	return fd.flush()
}

func (fd *FrameDecompressor) flush() error {
	for fd.written < len(fd.out) {
		n, err := fd.target.Write(fd.out[fd.written:])
		fd.written += n
		if err != nil {
			return err
		}
	}
	return nil
}

func translate(err error) error {
	var e szb200.Error
	if errors.As(err, &e) {
		switch int(e) {
		case -1:
			return ErrWrongMagicnumber
		case -2:
			return ErrCorruptSizes
		case -3:
			return ErrOutOfBlocks
		}
		if m, ok := structure.ErrorForCode(int(e)); ok {
			return m
		}
	}
	return err
}
