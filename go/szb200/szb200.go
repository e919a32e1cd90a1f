// Package szb200 is the thin cgo layer over libszb200.so (include/szb200.h): the B200-native
// zstd decode engine.  It owns nothing but the C context; header walking lives in package
// structure, the drop-in API in package decompression.
//
// NOTE: this image has no Go toolchain, so this package is written against the C ABI that the
// Python/ctypes tests exercise on the GPU, but has not been compiled here (see INTEGRATION.md).
package szb200

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../sparkzstd_b200 -lszb200 -Wl,-rpath,${SRCDIR}/../../sparkzstd_b200
#include <stdlib.h>
#include "szb200.h"
*/
import "C"

import (
	"errors"
	"runtime"
	"unsafe"
)

// Error is a negative szb200 status code; the values map 1:1 onto the reference's Err* variables.
type Error int

func (e Error) Error() string { return C.GoString(C.szb_strerror(C.int(e))) }

// FrameDesc and BlockDesc mirror szb_frame_desc / szb_block_desc (include/szb200.h) byte for byte.  They are PLAIN Go structs
// with exported fields on purpose: a cgo type (C.szb_frame_desc) is private to the package that names it -- another package can
// neither touch its lower-case fields nor mix its own C.uint64_t with this package's -- so package structure fills these and this
// package hands them to C as unsafe.Pointer.  The layout is pinned three times: static_asserts on every offset in
// csrc/abi_layout.h (compile time, C side), init() below against szb_abi_layout() (load time, Go side), and
// tests/test_abi_and_walker.py (the ctypes mirror).  Go lays out uint64/uint32/uint8 fields in declaration order with natural
// alignment, exactly like the C compiler does for this field order.
type FrameDesc struct {
	SrcOff         uint64 // offset of the frame's magic number inside src
	SrcLen         uint64 // bytes from the magic to the end of the last block
	WindowSize     uint64 // frame.go:28-36
	ContentSize    uint64 // frame.go:49-61, or ContentSizeUnknown
	DictionaryID   uint64 // frame.go:38-47
	FirstBlock     uint32
	NBlocks        uint32
	Checksum       uint32
	Status         int32
	Descriptor     uint8
	SingleSegment  uint8
	HasChecksum    uint8
	HasContentSize uint8
	ChecksumValid  uint32
}

type BlockDesc struct {
	SrcOff      uint64 // offset inside src of the block payload (after the 3-byte header)
	LitBufOff   uint64
	SeqBufOff   uint64
	BlockSize   uint32
	Frame       uint32
	LitRegen    uint32
	LitComp     uint32
	NSeq        uint32
	SeqOff      uint32
	HufOrigin   uint32
	LLOrigin    uint32
	OFOrigin    uint32
	MLOrigin    uint32
	Type        uint8
	Last        uint8
	LitType     uint8
	LitStreams  uint8
	LitHdrBytes uint8
	SeqHdrBytes uint8
	SeqModes    uint8
	Flags       uint8 // BlockTablesOnly: a dictionary's table-only row
	HdrStatus   int32 // the error the walk hit in this block's sequences-section header (the literals are still decoded first)
}

const ContentSizeUnknown = ^uint64(0)
const None = uint32(0xFFFFFFFF) // SZB_NONE

// layout() lists sizeof and every field offset in the order szb_abi_layout() (include/szb200.h) reports them.
func layout() []uint32 {
	var f FrameDesc
	var b BlockDesc
	return []uint32{
		uint32(unsafe.Sizeof(f)),
		uint32(unsafe.Offsetof(f.SrcOff)), uint32(unsafe.Offsetof(f.SrcLen)), uint32(unsafe.Offsetof(f.WindowSize)),
		uint32(unsafe.Offsetof(f.ContentSize)), uint32(unsafe.Offsetof(f.DictionaryID)), uint32(unsafe.Offsetof(f.FirstBlock)),
		uint32(unsafe.Offsetof(f.NBlocks)), uint32(unsafe.Offsetof(f.Checksum)), uint32(unsafe.Offsetof(f.Status)),
		uint32(unsafe.Offsetof(f.Descriptor)), uint32(unsafe.Offsetof(f.SingleSegment)), uint32(unsafe.Offsetof(f.HasChecksum)),
		uint32(unsafe.Offsetof(f.HasContentSize)), uint32(unsafe.Offsetof(f.ChecksumValid)),
		uint32(unsafe.Sizeof(b)),
		uint32(unsafe.Offsetof(b.SrcOff)), uint32(unsafe.Offsetof(b.LitBufOff)), uint32(unsafe.Offsetof(b.SeqBufOff)),
		uint32(unsafe.Offsetof(b.BlockSize)), uint32(unsafe.Offsetof(b.Frame)), uint32(unsafe.Offsetof(b.LitRegen)),
		uint32(unsafe.Offsetof(b.LitComp)), uint32(unsafe.Offsetof(b.NSeq)), uint32(unsafe.Offsetof(b.SeqOff)),
		uint32(unsafe.Offsetof(b.HufOrigin)), uint32(unsafe.Offsetof(b.LLOrigin)), uint32(unsafe.Offsetof(b.OFOrigin)),
		uint32(unsafe.Offsetof(b.MLOrigin)), uint32(unsafe.Offsetof(b.Type)), uint32(unsafe.Offsetof(b.Last)),
		uint32(unsafe.Offsetof(b.LitType)), uint32(unsafe.Offsetof(b.LitStreams)), uint32(unsafe.Offsetof(b.LitHdrBytes)),
		uint32(unsafe.Offsetof(b.SeqHdrBytes)), uint32(unsafe.Offsetof(b.SeqModes)), uint32(unsafe.Offsetof(b.Flags)),
		uint32(unsafe.Offsetof(b.HdrStatus)),
	}
}

// A binding whose structs drifted from the header must not run: the device would read garbage tables.
func init() {
	want := layout()
	got := make([]C.uint32_t, len(want)+8)
	n := int(C.szb_abi_layout(&got[0], C.uint32_t(len(got))))
	if n != len(want) {
		panic("szb200: szb_abi_layout reports another number of fields than this binding mirrors")
	}
	for i := range want {
		if uint32(got[i]) != want[i] {
			panic("szb200: FrameDesc/BlockDesc do not match szb_frame_desc/szb_block_desc (include/szb200.h)")
		}
	}
}

// Ctx is one CUDA stream plus scratch arenas on one GPU.  Not goroutine-safe: one per goroutine,
// exactly like a reference FrameDecompressor.
type Ctx struct{ p *C.szb_ctx }

// NewCtx fails when no CUDA device is usable: there is no CPU fallback.
func NewCtx(device int) (*Ctx, error) {
	var p *C.szb_ctx
	if rc := C.szb_ctx_create(C.int(device), nil, &p); rc != 0 {
		return nil, Error(rc)
	}
	c := &Ctx{p: p}
	runtime.SetFinalizer(c, func(c *Ctx) { c.Close() })
	return c, nil
}

func (c *Ctx) Close() {
	if c.p != nil {
		C.szb_ctx_destroy(c.p)
		c.p = nil
	}
}

// DecompressFrame decodes the single frame at the start of src on the GPU and returns the
// decompressed bytes and how many compressed bytes the frame occupied (checksum excluded,
// like the reference, which never reads it).
func (c *Ctx) DecompressFrame(src []byte) ([]byte, int, error) {
	if len(src) == 0 {
		return nil, 0, Error(C.SZB_ERR_UNEXPECTED_EOF)
	}
	var out *C.uint8_t
	var n, used C.size_t
	rc := C.szb_decompress_frame(c.p, (*C.uint8_t)(unsafe.Pointer(&src[0])), C.size_t(len(src)), &out, &n, &used)
	if rc != 0 {
		return nil, int(used), Error(rc)
	}
	defer C.szb_free(unsafe.Pointer(out))
	return C.GoBytes(unsafe.Pointer(out), C.int(n)), int(used), nil
}

// DecodeBatch decodes independent frames in one launch sequence.  frames[i] = src[off[i]:off[i]+len[i]].
// dst receives the frames back to back; the returned slices say where.  The library copies what it
// needs before returning: no Go pointer is retained (cgo rule).
// FlagVerifyChecksum asks the GPU to verify each frame's content checksum (the reference never does).
const FlagVerifyChecksum = uint32(C.SZB_FLAG_VERIFY_CHECKSUM)

func (c *Ctx) DecodeBatch(src []byte, off, length []uint64, dst []byte, flags uint32) (outOff, outLen []uint64, status []int32, err error) {
	n := len(off)
	if n == 0 {
		return nil, nil, nil, nil
	}
	if len(length) != n || len(src) == 0 {
		return nil, nil, nil, errors.New("szb200: bad batch arguments")
	}
	outOff = make([]uint64, n)
	outLen = make([]uint64, n)
	status = make([]int32, n)
	var dp *C.uint8_t
	if len(dst) > 0 {
		dp = (*C.uint8_t)(unsafe.Pointer(&dst[0]))
	}
	rc := C.szb_decode_batch(c.p, (*C.uint8_t)(unsafe.Pointer(&src[0])), C.size_t(len(src)),
		(*C.uint64_t)(unsafe.Pointer(&off[0])), (*C.uint64_t)(unsafe.Pointer(&length[0])), C.uint32_t(n),
		dp, C.size_t(len(dst)),
		(*C.uint64_t)(unsafe.Pointer(&outOff[0])), (*C.uint64_t)(unsafe.Pointer(&outLen[0])),
		(*C.int32_t)(unsafe.Pointer(&status[0])), C.uint32_t(flags))
	if rc == C.SZB_ERR_CUDA || rc == C.SZB_ERR_INVALID_ARGUMENT || rc == C.SZB_ERR_NOMEM {
		return nil, nil, nil, Error(rc)
	}
	return outOff, outLen, status, nil
}

// DecodeBlocks is the descriptor-level entry: device pointers plus the tables the Go walker
// (package structure) emitted.
func (c *Ctx) DecodeBlocks(dSrc unsafe.Pointer, srcLen int, frames []FrameDesc, blocks []BlockDesc, dDst unsafe.Pointer, dstCap int) (outOff, outLen []uint64, status []int32, err error) {
	n := len(frames)
	if n == 0 {
		return nil, nil, nil, nil
	}
	outOff = make([]uint64, n)
	outLen = make([]uint64, n)
	status = make([]int32, n)
	// the tables are host memory the library copies before it returns; they hold no Go pointers
	fp := (*C.szb_frame_desc)(unsafe.Pointer(&frames[0]))
	var bp *C.szb_block_desc
	if len(blocks) > 0 {
		bp = (*C.szb_block_desc)(unsafe.Pointer(&blocks[0]))
	}
	rc := C.szb_decode_blocks(c.p, dSrc, C.size_t(srcLen), fp, C.uint32_t(n), bp, C.uint32_t(len(blocks)), dDst, C.size_t(dstCap),
		(*C.uint64_t)(unsafe.Pointer(&outOff[0])), (*C.uint64_t)(unsafe.Pointer(&outLen[0])), (*C.int32_t)(unsafe.Pointer(&status[0])))
	if rc == C.SZB_ERR_CUDA || rc == C.SZB_ERR_INVALID_ARGUMENT || rc == C.SZB_ERR_NOMEM {
		return nil, nil, nil, Error(rc)
	}
	return outOff, outLen, status, nil
}

// DecodeStream decodes a `.zst` stream as the zstd tools write it -- frames back to back, skippable frames in between --
// into dst and returns the number of bytes of content (szb_decode_stream; not a reference behaviour).
func (c *Ctx) DecodeStream(src, dst []byte, flags uint32) (int, error) {
	if len(src) == 0 {
		return 0, nil
	}
	var dp *C.uint8_t
	if len(dst) > 0 {
		dp = (*C.uint8_t)(unsafe.Pointer(&dst[0]))
	}
	var found C.uint32_t
	var total C.uint64_t
	rc := C.szb_decode_stream(c.p, (*C.uint8_t)(unsafe.Pointer(&src[0])), C.size_t(len(src)), dp, C.size_t(len(dst)),
		nil, nil, nil, 0, &found, &total, C.uint32_t(flags))
	if rc != 0 {
		return int(total), Error(rc)
	}
	return int(total), nil
}

// ShardFrames is the multi-GPU split (szb_shard_frames): frames are independent, so G GPUs decode a partition of the frame
// list, one process (one Ctx) per GPU.  shardOf[i] is the GPU of frame i; deterministic for equal weights on every rank.
func ShardFrames(weight []uint64, shards int) (shardOf []uint32, load []uint64, err error) {
	shardOf = make([]uint32, len(weight))
	load = make([]uint64, shards)
	if shards <= 0 {
		return nil, nil, errors.New("szb200: shards must be positive")
	}
	var wp *C.uint64_t
	var sp *C.uint32_t
	if len(weight) > 0 {
		wp = (*C.uint64_t)(unsafe.Pointer(&weight[0]))
		sp = (*C.uint32_t)(unsafe.Pointer(&shardOf[0]))
	}
	if rc := C.szb_shard_frames(wp, C.uint32_t(len(weight)), C.uint32_t(shards), sp, (*C.uint64_t)(unsafe.Pointer(&load[0]))); rc != 0 {
		return nil, nil, Error(rc)
	}
	return shardOf, load, nil
}

// Dict is a zstd dictionary (raw content, or formatted: magic 0xEC30A437) parsed and resident on the context's GPU.
// Not a reference behaviour: the reference parses Dictionary_ID and ignores it (structure/frame.go:38-47).
type Dict struct{ p *C.szb_dict }

func (c *Ctx) NewDict(dict []byte) (*Dict, error) {
	var p *C.szb_dict
	var dp *C.uint8_t
	if len(dict) > 0 {
		dp = (*C.uint8_t)(unsafe.Pointer(&dict[0]))
	}
	if rc := C.szb_dict_create(c.p, dp, C.size_t(len(dict)), &p); rc != 0 {
		return nil, Error(rc)
	}
	return &Dict{p: p}, nil
}

func (d *Dict) ID() uint32 { return uint32(C.szb_dict_id(d.p)) }

func (d *Dict) Close() {
	if d.p != nil {
		C.szb_dict_destroy(d.p)
		d.p = nil
	}
}

// DecodeBatchDict is DecodeBatch for frames that were compressed with the dictionary.
func (c *Ctx) DecodeBatchDict(d *Dict, src []byte, off, length []uint64, dst []byte, flags uint32) (outOff, outLen []uint64, status []int32, err error) {
	n := len(off)
	if n == 0 {
		return nil, nil, nil, nil
	}
	if len(length) != n || len(src) == 0 || d == nil || d.p == nil {
		return nil, nil, nil, errors.New("szb200: bad batch arguments")
	}
	outOff = make([]uint64, n)
	outLen = make([]uint64, n)
	status = make([]int32, n)
	var dp *C.uint8_t
	if len(dst) > 0 {
		dp = (*C.uint8_t)(unsafe.Pointer(&dst[0]))
	}
	rc := C.szb_decode_batch_dict(c.p, d.p, (*C.uint8_t)(unsafe.Pointer(&src[0])), C.size_t(len(src)),
		(*C.uint64_t)(unsafe.Pointer(&off[0])), (*C.uint64_t)(unsafe.Pointer(&length[0])), C.uint32_t(n),
		dp, C.size_t(len(dst)),
		(*C.uint64_t)(unsafe.Pointer(&outOff[0])), (*C.uint64_t)(unsafe.Pointer(&outLen[0])),
		(*C.int32_t)(unsafe.Pointer(&status[0])), C.uint32_t(flags))
	if rc == C.SZB_ERR_CUDA || rc == C.SZB_ERR_INVALID_ARGUMENT || rc == C.SZB_ERR_NOMEM {
		return nil, nil, nil, Error(rc)
	}
	return outOff, outLen, status, nil
}
