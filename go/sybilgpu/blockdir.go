// AddBlockDir: stage a sybil block directory without going through encoding/gob in Go.
//
// NOT COMPILED IN THIS REPOSITORY'S BUILD IMAGE (no Go toolchain).  The C++ reader behind it
// (include/sybilgob.h, libsybilgob.so) IS built and tested there (tests/test_gobread.py).
//
// LoadBlockFromDir (src/lib/table_block_io.go:225-310) spends most of its time in gob's reflection-driven
// decode of int_*.db / str_*.db; this hands the directory to the native reader, which returns the
// sg_block_desc that sg_table_add_block takes, arrays still encoded.
package sybilgpu

/*
#cgo LDFLAGS: -lsybilgob -lz
#include <stdlib.h>
#include "sybilgob.h"
*/
import "C"

import (
	"errors"
	"unsafe"
)

// AddBlockDir reads <dir>/info.db and the column files of the columns selected by load (nil = all;
// the LoadSpec) and stages the block.  names/types are the table's KeyTable/KeyTypes by slot.
func (t *Table) AddBlockDir(dir string, names []string, types []int32, load []bool, blockIndex int64) error {
	n := len(names)
	cnames := make([]*C.char, n+1)
	for i, s := range names {
		cnames[i] = C.CString(s)
		defer C.free(unsafe.Pointer(cnames[i]))
	}
	var mask *C.uint8_t
	if load != nil {
		m := make([]C.uint8_t, n+1)
		for i, b := range load {
			if b {
				m[i] = 1
			}
		}
		mask = &m[0]
	}
	cdir := C.CString(dir)
	defer C.free(unsafe.Pointer(cdir))
	var errbuf [512]C.char
	var ctypes *C.int32_t
	if n > 0 {
		ctypes = (*C.int32_t)(unsafe.Pointer(&types[0]))
	}
	b := C.sgob_read_block_dir(cdir, &cnames[0], ctypes, C.int32_t(n), mask, C.int64_t(blockIndex), &errbuf[0], C.size_t(len(errbuf)))
	if b == nil {
		return errors.New("sybilgob: " + C.GoString(&errbuf[0]))
	}
	defer C.sgob_block_free(b)
	if rc := C.sg_table_add_block(t.h, C.sgob_block_desc(b)); rc != C.SG_OK {
		return t.c.err()
	}
	return nil
}
