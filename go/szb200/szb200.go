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

// FrameDesc and BlockDesc mirror szb_frame_desc / szb_block_desc byte for byte.
type FrameDesc = C.szb_frame_desc
type BlockDesc = C.szb_block_desc

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
	outOff = make([]uint64, n)
	outLen = make([]uint64, n)
	status = make([]int32, n)
	var fp *FrameDesc
	var bp *BlockDesc
	if n > 0 {
		fp = &frames[0]
	}
	if len(blocks) > 0 {
		bp = &blocks[0]
	}
	rc := C.szb_decode_blocks(c.p, dSrc, C.size_t(srcLen), fp, C.uint32_t(n), bp, C.uint32_t(len(blocks)), dDst, C.size_t(dstCap),
		(*C.uint64_t)(unsafe.Pointer(&outOff[0])), (*C.uint64_t)(unsafe.Pointer(&outLen[0])), (*C.int32_t)(unsafe.Pointer(&status[0])))
	if rc == C.SZB_ERR_CUDA || rc == C.SZB_ERR_INVALID_ARGUMENT || rc == C.SZB_ERR_NOMEM {
		return nil, nil, nil, Error(rc)
	}
	return outOff, outLen, status, nil
}
