//go:build cgo && plz4cuda

// Package gpubatch replaces the per-block worker loops of the async writer and reader
// (internal/pkg/async/writer.go:232-282 compressLoop, internal/pkg/async/reader.go:192-221 _decompressLoop) by
// batchers: blocks are gathered into one pinned slab and cross the cgo boundary as ONE call per batch.  Ordering, the
// progress callback, the Flush barrier and the error state stay where they are (writeLoop, NextBlock); only what runs
// between inChan and outChan changes.  The batcher goroutines are started through opts.WorkerPool.Submit exactly like
// the loops they replace (async/writer.go:439-467), so WithWorkerPool keeps its meaning.
//
// NOT COMPILED in the repository that ships this file (its image has no Go toolchain); written against Go 1.21.
package gpubatch

/*
#cgo LDFLAGS: -lplz4cu
#include <stdint.h>
#include "plz4cu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"unsafe"
)

var ErrEngine = errors.New("plz4cu engine failure")

// WorkerPool is opts.WorkerPool (internal/pkg/opts/opts.go:43-45).
type WorkerPool interface {
	Submit(task func())
}

// Block is what flows through inChan / outChan: the block's index in the stream and its bytes.  For the writer Data is
// uncompressed input and the result is the framed record [size][payload][xxh32] of blk.CompressToBlk (blk/blk.go:87-106);
// for the reader Data is one record as FrameReader found it (blk/frame.go:54-112, hash check skipped) and the result is
// the decoded block.
type Block struct {
	Idx  int
	Data []byte
	Err  error
}

// Slab is pinned host memory from the engine's pool: what blk.BorrowBlk / ReturnBlk hand out (blk/pool.go:35-69).
type Slab struct {
	p   unsafe.Pointer
	len int
}

func BorrowSlab(n int) (*Slab, error) {
	p := C.plz4cu_host_alloc(C.size_t(n))
	if p == nil {
		return nil, fmt.Errorf("%w: %s", ErrEngine, C.GoString(C.plz4cu_last_error()))
	}
	return &Slab{p: p, len: n}, nil
}
func (s *Slab) Bytes() []byte { return unsafe.Slice((*byte)(s.p), s.len) }
func (s *Slab) Return()       { C.plz4cu_host_free(s.p); s.p = nil }

// CntBorrowed is blk.CntBorrowed (blk/pool.go:29-33): the leak gauge the reference's tests check after every case.
func CntBorrowed() int64 { return int64(C.plz4cu_host_outstanding()) }

// Config: what the loops read from opts.OptsT.
type Config struct {
	BlockSize     int  // descriptor.BlockIdxT.Size()
	BlockChecksum bool // WithBlockChecksum
	MaxBlocks     int  // blocks per engine call: nParallel x opts.CalcPending (opts/opts.go:62-95), at least 1
	Dict          unsafe.Pointer // *C.plz4cu_dict_t or nil (NewDictCtx(...).handle)
}

// CompressLoop is the body that replaces compressLoop: gather up to MaxBlocks blocks (never waiting for more than are
// already queued: a Flush must not stall), compress them with one call, emit the records in order.
func CompressLoop(cfg Config, in <-chan Block, out chan<- Block) {
	bsz := cfg.BlockSize
	src, err := BorrowSlab(cfg.MaxBlocks * bsz)
	if err != nil {
		fail(in, out, err)
		return
	}
	defer src.Return()
	packed, err := BorrowSlab(cfg.MaxBlocks * (bsz + 8))
	if err != nil {
		fail(in, out, err)
		return
	}
	defer packed.Return()
	off := make([]C.uint64_t, cfg.MaxBlocks)
	ln := make([]C.uint32_t, cfg.MaxBlocks)
	poff := make([]C.uint64_t, cfg.MaxBlocks+1)
	idx := make([]int, 0, cfg.MaxBlocks)
	sb := src.Bytes()
	for first := range in {
		idx = idx[:0]
		fill := 0
		take := func(b Block) {
			off[len(idx)] = C.uint64_t(fill)
			ln[len(idx)] = C.uint32_t(len(b.Data))
			copy(sb[fill:], b.Data)
			fill += len(b.Data)
			idx = append(idx, b.Idx)
		}
		take(first)
	gather:
		for len(idx) < cfg.MaxBlocks {
			select {
			case b, ok := <-in:
				if !ok {
					break gather
				}
				take(b)
			default:
				break gather
			}
		}
		cx := C.int(0)
		if cfg.BlockChecksum {
			cx = 1
		}
		rc := C.plz4cu_compress_batch_host(src.p, &off[0], &ln[0], C.uint32_t(len(idx)), C.uint32_t(bsz), cx, 0,
			(*C.plz4cu_dict_t)(cfg.Dict), packed.p, C.uint64_t(packed.len), &poff[0])
		if rc < 0 {
			e := fmt.Errorf("%w: %s", ErrEngine, C.GoString(C.plz4cu_last_error()))
			for _, i := range idx {
				out <- Block{Idx: i, Err: e}
			}
			continue
		}
		pb := packed.Bytes()
		for k, i := range idx {
			rec := make([]byte, int(poff[k+1]-poff[k])) // the reference hands a pooled BlkT on; copy-out keeps the slab reusable
			copy(rec, pb[poff[k]:poff[k+1]])
			out <- Block{Idx: i, Data: rec}
		}
	}
}

// Result codes of plz4cu_decompress_batch_host per block (include/plz4cu.h "Error convention").
const (
	eBlockHash = -0x7F000001
	eOverflow  = -0x7F000002
	eStall     = -0x7F000003
)

// Sentinel errors the reader maps the codes to; in plz4 these are zerr.ErrBlockHash, zerr.ErrBlockSizeOverflow,
// zerr.ErrDecompress joined with zerr.ErrCorrupted (zerr/zerr.go:11-41, compress/decompress.go:33-36).
var (
	ErrBlockHash         = errors.New("block hash mismatch")
	ErrBlockSizeOverflow = errors.New("block size overflow")
	ErrDecompress        = errors.New("lz4 fail decompress")
)

// DecompressLoop replaces _decompressLoop: records in, decoded blocks out, one engine call per batch, block checksums
// verified on the device (blk/frame.go:114-127 moves into the kernel).
func DecompressLoop(cfg Config, in <-chan Block, out chan<- Block) {
	bsz := cfg.BlockSize
	recs, err := BorrowSlab(cfg.MaxBlocks * (bsz + 8))
	if err != nil {
		fail(in, out, err)
		return
	}
	defer recs.Return()
	dst, err := BorrowSlab(cfg.MaxBlocks * bsz)
	if err != nil {
		fail(in, out, err)
		return
	}
	defer dst.Return()
	off := make([]C.uint64_t, cfg.MaxBlocks)
	res := make([]C.int32_t, cfg.MaxBlocks)
	idx := make([]int, 0, cfg.MaxBlocks)
	rb := recs.Bytes()
	for first := range in {
		idx = idx[:0]
		fill := 0
		take := func(b Block) {
			off[len(idx)] = C.uint64_t(fill)
			copy(rb[fill:], b.Data)
			fill += len(b.Data)
			idx = append(idx, b.Idx)
		}
		take(first)
	gather:
		for len(idx) < cfg.MaxBlocks {
			select {
			case b, ok := <-in:
				if !ok {
					break gather
				}
				take(b)
			default:
				break gather
			}
		}
		cx := C.int(0)
		if cfg.BlockChecksum {
			cx = 1
		}
		rc := C.plz4cu_decompress_batch_host(recs.p, C.uint64_t(fill), &off[0], nil, C.uint32_t(len(idx)), C.uint32_t(bsz), cx, 0,
			(*C.plz4cu_dict_t)(cfg.Dict), dst.p, C.uint64_t(bsz), &res[0])
		if rc < 0 {
			e := fmt.Errorf("%w: %s", ErrEngine, C.GoString(C.plz4cu_last_error()))
			for _, i := range idx {
				out <- Block{Idx: i, Err: e}
			}
			continue
		}
		db := dst.Bytes()
		for k, i := range idx {
			r := int(res[k])
			switch {
			case r == eBlockHash:
				out <- Block{Idx: i, Err: ErrBlockHash}
			case r == eOverflow:
				out <- Block{Idx: i, Err: ErrBlockSizeOverflow}
			case r == eStall:
				out <- Block{Idx: i, Err: fmt.Errorf("%w: a decode team stalled", ErrEngine)}
			case r < 0:
				out <- Block{Idx: i, Err: fmt.Errorf("%w: code %d", ErrDecompress, r)}
			default:
				blk := make([]byte, r)
				copy(blk, db[k*bsz:k*bsz+r])
				out <- Block{Idx: i, Data: blk}
			}
		}
	}
}

func fail(in <-chan Block, out chan<- Block, err error) {
	for b := range in {
		out <- Block{Idx: b.Idx, Err: err}
	}
}

// Kickoff starts n batchers the way kickoffAsync starts its loops (async/writer.go:439-467): through the caller's pool.
func Kickoff(wp WorkerPool, n int, loop func()) {
	for i := 0; i < n; i++ {
		wp.Submit(loop)
	}
}
