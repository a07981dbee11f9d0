/* sybilgob.h — reading sybil's block directories without Go (host side, no CUDA).
 *
 * sybil stores a block as one Go `encoding/gob` file per column plus info.db:
 *     <block>/int_<col>.db   gob(SavedIntColumn)    src/lib/column_store.go:46-54
 *     <block>/str_<col>.db   gob(SavedStrColumn)    src/lib/column_store.go:56-64
 *     <block>/info.db        gob(SavedColumnInfo)   src/lib/column_store.go:39-44
 * (each optionally gzip-wrapped as `.db.gz`, src/lib/file_decoder.go:35-53).  The reference host
 * reads them with gob.Decode inside LoadBlockFromDir / unpackIntCol / unpackStrCol
 * (src/lib/table_block_io.go:225-310, column_store_io.go:493-501,690-697).  This library does the same
 * decode in C++ and hands back the sg_block_desc that sg_table_add_block takes, arrays still encoded.
 *
 * Plain C ABI; the returned descriptor points into memory owned by the sgob_block.
 */
#ifndef SYBILGOB_H
#define SYBILGOB_H

#include <stddef.h>
#include <stdint.h>

#include "sybilgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgob_block sgob_block;

/* Reads the block at `dir`.  col_names/col_types: the table's KeyTable/KeyTypes (slot i = column i);
 * load_mask: ncols flags, 0 = do not open that column's file (the LoadSpec; NULL = all).  A column
 * whose file does not exist stays absent (table_block_io.go:271-277).  Returns NULL on error (missing
 * or malformed info.db, malformed column file) with a message in err. */
sgob_block* sgob_read_block_dir(const char* dir, const char* const* col_names, const int32_t* col_types, int32_t ncols,
                                const uint8_t* load_mask, int64_t block_index, char* err, size_t errlen);
const sg_block_desc* sgob_block_desc(const sgob_block* b);
/* 1: blocks read from now on keep their arrays narrow (uint16 record ids, int16 / int32 value deltas, uint16
 * local string ids: sg_column_desc::id_bits / value_bits) instead of widening every varint to Go's uint32 /
 * int64 / int32 — half to a quarter of the bytes over PCIe and HBM.  0 (default): Go's decoded types. */
void sgob_set_narrow(int on);
void sgob_block_free(sgob_block* b);

/* bytes of column data decoded (sum over the arrays of the descriptor) */
int64_t sgob_block_bytes(const sgob_block* b);

/* ---- the table directory: <dbdir>/<table>/info.db + one sub-directory per block ----------------
 * info.db = gob(Table{Name, KeyTable, KeyTypes, IntInfo, StrInfo}) (src/lib/table_io.go:66-78,128-175);
 * blocks = the sub-directories file_looks_like_block accepts (table_io.go:213-239), in name order
 * (LoadAndQueryRecords walks ioutil.ReadDir, table_query.go:22,96-111). */
typedef struct sgob_table sgob_table;
sgob_table* sgob_table_open(const char* dbdir, const char* table, char* err, size_t errlen);
void sgob_table_free(sgob_table* t);
int32_t sgob_table_num_cols(const sgob_table* t);                 /* key slots: max slot + 1 */
const char* sgob_table_col_name(const sgob_table* t, int32_t slot); /* "" for an unused slot */
int32_t sgob_table_col_type(const sgob_table* t, int32_t slot);   /* KeyTypes: 1 int, 2 str, else 0 */
/* table-level IntInfo{Min,Max} of a column (the histogram extents, hist.go:27-38); 1 if present */
int32_t sgob_table_int_info(const sgob_table* t, int32_t slot, int64_t* min_out, int64_t* max_out);
int64_t sgob_table_num_blocks(const sgob_table* t);
const char* sgob_table_block_dir(const sgob_table* t, int64_t i);

#ifdef __cplusplus
}
#endif
#endif
