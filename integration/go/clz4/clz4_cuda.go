//go:build cgo && plz4cuda

// Package clz4, CUDA flavour: the same exported surface as internal/pkg/clz4/clz4.go (reference lines cited per
// function), bound to libplz4cu.so instead of the vendored liblz4.  Drop this file into internal/pkg/clz4/ and build
// with `-tags plz4cuda` (give clz4.go the constraint `cgo && !plz4cuda`).  Level-1, independent-block entry points go
// to the GPU engine; the HC and linked-block contexts of the reference stay on its own C code (they are declared in
// clz4_hc.go, a copy of the HC / linked parts of clz4.go that a maintainer keeps under the same tag), because the
// engine refuses them (PLZ4CU_Z_UNSUPPORTED) rather than degrade silently.
//
// NOT COMPILED in the repository that ships this file (its image has no Go toolchain); written against Go 1.21.
package clz4

/*
#cgo LDFLAGS: -lplz4cu
#include <stdlib.h>
#include "plz4cu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"math"
	"runtime"
	"sync"
	"unsafe"
)

var (
	ErrLz4Compress   = errors.New("lz4 fail compress; insufficient destination buffer") // clz4.go:15-18
	ErrLz4Decompress = errors.New("lz4 fail decompress")
	// ErrEngine is a CUDA / engine failure (no device, out of memory, a launch error).  It is NOT ErrLz4Compress: blk.CompressToBlk
	// (blk/blk.go:78-92) turns ErrLz4Compress into a stored block, and an engine failure must abort the stream instead.
	ErrEngine = errors.New("plz4cu engine failure")
)

var initOnce sync.Once
var initErr error

// Init binds the process to a device before the first call (plz4cu_init); called lazily with device 0 otherwise.
func Init(device int) error {
	initOnce.Do(func() {
		if r := C.plz4cu_init(C.int(device)); r < 0 {
			initErr = fmt.Errorf("%w: %s", ErrEngine, C.GoString(C.plz4cu_last_error()))
		}
	})
	return initErr
}

func ptr(b []byte) unsafe.Pointer {
	if len(b) == 0 {
		return nil
	}
	return unsafe.Pointer(&b[0])
}

func engineErr() error {
	return fmt.Errorf("%w: %s", ErrEngine, C.GoString(C.plz4cu_last_error()))
}

// CompressBound: clz4.go:27-29 (LZ4_compressBound).
func CompressBound(sz int) int {
	return int(C.plz4cu_compress_bound(C.size_t(sz)))
}

// CompressFast: clz4.go:31-45.  acceleration is ignored: the engine implements level 1 (acceleration 1).
func CompressFast(source, dest []byte, acceleration int) (int, error) {
	if err := Init(0); err != nil {
		return 0, err
	}
	r := int(C.plz4cu_compress_fast(ptr(source), C.int(len(source)), ptr(dest), C.int(len(dest))))
	switch {
	case r == math.MinInt32:
		return 0, engineErr()
	case r == 0:
		return 0, ErrLz4Compress
	}
	return r, nil
}

// DecompressSafe: clz4.go:47-60.  Negative return codes are liblz4's own (-(position)-1).
func DecompressSafe(source, dest []byte) (int, error) {
	if err := Init(0); err != nil {
		return 0, err
	}
	r := int(C.plz4cu_decompress_safe(ptr(source), C.int(len(source)), ptr(dest), C.int(len(dest))))
	switch {
	case r == math.MinInt32:
		return 0, engineErr()
	case r < 0:
		return r, fmt.Errorf("%w: code %d", ErrLz4Decompress, r)
	}
	return r, nil
}

// dictCache keeps one device-resident dictionary per distinct slice: DecompressSafeWithDict (clz4.go:62-78) takes the raw
// bytes on every call, and uploading 64 KiB per block would dominate.  Keyed by the slice's first byte address and length,
// the identity compress.DictT keeps stable for the life of a reader (compress/dict.go:5-56).
var dictCache sync.Map // [2]uintptr -> *DictCtx

// DecompressSafeWithDict: clz4.go:62-78.
func DecompressSafeWithDict(source, dest, dict []byte) (int, error) {
	if len(dict) == 0 {
		return DecompressSafe(source, dest)
	}
	key := [2]uintptr{uintptr(ptr(dict)), uintptr(len(dict))}
	v, ok := dictCache.Load(key)
	if !ok {
		v, _ = dictCache.LoadOrStore(key, NewDictCtx(dict))
	}
	d := v.(*DictCtx)
	if d.h == nil {
		return 0, engineErr()
	}
	r := int(C.plz4cu_decompress_safe_dict(d.h, ptr(source), C.int(len(source)), ptr(dest), C.int(len(dest))))
	switch {
	case r == math.MinInt32:
		return 0, engineErr()
	case r < 0:
		return r, fmt.Errorf("%w: code %d", ErrLz4Decompress, r)
	}
	return r, nil
}

// DictCtx: clz4.go:96-120 (LZ4_loadDictSlow on a private copy).  Here the last 64 KiB live on the device together with
// the encoder's table for them.
type DictCtx struct {
	data []byte
	h    *C.plz4cu_dict_t
}

func NewDictCtx(dict []byte) *DictCtx {
	_ = Init(0)
	dupe := make([]byte, len(dict))
	copy(dupe, dict)
	c := &DictCtx{data: dupe}
	c.h = C.plz4cu_dict_create(ptr(dupe), C.size_t(len(dupe)))
	runtime.SetFinalizer(c, func(c *DictCtx) {
		if c.h != nil {
			C.plz4cu_dict_destroy(c.h)
			c.h = nil
		}
	})
	return c
}

// StreamIndieCtx: clz4.go:151-179 (attach the dictionary context, compress one independent block).
type StreamIndieCtx struct {
	dict *DictCtx
}

func NewStreamIndieCtx(dict *DictCtx) *StreamIndieCtx {
	return &StreamIndieCtx{dict: dict}
}

func (c *StreamIndieCtx) Compress(src, dst []byte) (int, error) {
	if c.dict == nil || c.dict.h == nil {
		return CompressFast(src, dst, 1)
	}
	r := int(C.plz4cu_compress_fast_dict(c.dict.h, ptr(src), C.int(len(src)), ptr(dst), C.int(len(dst))))
	switch {
	case r == math.MinInt32:
		return 0, engineErr()
	case r == 0:
		return 0, ErrLz4Compress
	}
	runtime.KeepAlive(c.dict)
	return r, nil
}
