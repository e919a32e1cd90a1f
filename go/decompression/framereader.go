package decompression

import (
	"bytes"
	"io"
)

// FrameReader wraps a FrameDecompressor and provides the io.Reader interface (framereader.go:9-14).
type FrameReader struct {
	fd          *FrameDecompressor
	buffer      bytes.Buffer
	PrintStatus bool
	readTotal   int64
}

// NewFrameReader mirrors framereader.go:17-33: NewFrameReader(nil) is legal; with a source the
// magic number and the frame header are checked eagerly.
func NewFrameReader(source io.Reader) (*FrameReader, error) {
	fr := &FrameReader{}
	fr.fd = NewFrameDecompressor(source, &fr.buffer)
	if source != nil {
		if err := fr.fd.CheckMagicnum(); err != nil {
			return nil, err
		}
		if err := fr.fd.DecodeFrameHeader(); err != nil {
			return nil, err
		}
	}
	return fr, nil
}

// Reset mirrors framereader.go:35-49.
func (fr *FrameReader) Reset(source io.Reader) error {
	fr.buffer.Reset()
	fr.fd.Reset(source, &fr.buffer)
	fr.readTotal = 0
	if source != nil {
		if err := fr.fd.CheckMagicnum(); err != nil {
			return err
		}
		return fr.fd.DecodeFrameHeader()
	}
	return nil
}

// Read mirrors framereader.go:51-109.  The reference emits bytes only when its window ring evicts
// them; here the frame is decoded in one go on the first Read and then drained -- both are legal
// io.Reader behaviours (short reads, io.EOF after the last byte).
func (fr *FrameReader) Read(target []byte) (int, error) {
	for !fr.fd.CurrentBlock.Header.LastBlock {
		if err := fr.fd.DecodeNextBlock(); err != nil {
			return 0, err
		}
		fr.fd.BlockCounter++
	}
	if err := fr.fd.flush(); err != nil {
		return 0, err
	}
	if fr.buffer.Len() == 0 {
		return 0, io.EOF
	}
	n, _ := fr.buffer.Read(target)
	if fr.PrintStatus {
		print("Read bytes: ")
		println(fr.readTotal + int64(n))
	}
	fr.readTotal += int64(n)
	return n, nil
}
