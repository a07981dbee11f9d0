/*
 * blockgen.cpp — synthetic sybil blocks in the reference's post-gob column form.
 *
 * Produces, for consecutive 65,536-row blocks (CHUNK_SIZE, src/lib/table.go:44),
 * the arrays a gob decode of int_<col>.db / str_<col>.db would yield
 * (SavedIntColumn / SavedStrColumn, src/lib/column_store.go:46-64), choosing the
 * encoding exactly as the reference's digest does:
 *   - SeparateRecordsIntoColumns (column_store_io.go:366-417): value -> ascending
 *     row ids; ids delta-encoded when the column has <= CARDINALITY_THRESHOLD
 *     (5000, column_store_io.go:18) distinct values in the block (:21-38);
 *   - SaveIntsToColumns (:64-137): more distinct values than the threshold ->
 *     Values[max_r] with 0 for rows lacking the field, then delta-encoded (:99-113);
 *   - SaveStrsToColumns (:219-303): per-block string ids in first-seen order, raw
 *     int32 ids (not delta) in the value-array form, StringTable[id] = string.
 * Row values come from a counter-based generator (splitmix64 of seed, column, row)
 * so any block can be produced independently and identically on any thread.
 *
 * This is input generation for tests and bench.py (and the seed of the "block
 * writer" row of SURVEY.md §8f); it is not part of the query path and not the oracle.
 */
#include <algorithm>
#include <atomic>
#include <climits>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/sybilgpu.h"

extern "C" {

enum sbg_kind {
  SBG_UNIFORM = 0, /* lo + u % span */
  SBG_SUM4 = 1,    /* lo + (sum of the four 16-bit fields of u) * span / (4*65535) */
  SBG_TIME = 2,    /* lo + floor(row * a / b) + u % span   (a/b = time advance per row) */
  SBG_STRKEY = 3   /* prefix + decimal(u % span) */
};

typedef struct sbg_col {
  int32_t col_slot;
  int32_t col_type;      /* sg_coltype */
  int32_t kind;          /* sbg_kind */
  int32_t null_per_1024; /* rows lacking the field: second draw % 1024 < this */
  int64_t lo;
  int64_t span;
  int64_t a, b;
  char prefix[16];
} sbg_col;

typedef struct sbg_spec {
  uint64_t seed;
  int64_t total_rows;
  int32_t block_rows;
  int32_t ncols;
  int32_t cardinality_threshold; /* CARDINALITY_THRESHOLD */
  int32_t num_col_slots;
  const sbg_col* cols;
  /* 1: arrays as narrow as their values allow, the way a decoder that does not widen gob's varints hands them
   * over (sybilgpu.h, sg_column_desc::id_bits / value_bits): uint16 record ids, int32 / int16 value deltas on
   * top of value_base, uint16 local string ids.  0: Go's decoded types (uint32 / int64 / int32). */
  int32_t narrow;
  int32_t _pad;
} sbg_spec;

}  // extern "C"

namespace {

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct Arena {
  char* base = nullptr;
  size_t cap = 0;
  std::atomic<size_t> used{0};
  void* alloc(size_t bytes) {
    size_t sz = (bytes + 127) & ~(size_t)127;
    size_t off = used.fetch_add(sz);
    if (off + sz > cap) return nullptr;
    return base + off;
  }
};

struct BlockOut {
  sg_block_desc desc;
  std::vector<sg_column_desc> cols;
  std::vector<sg_int_info> info;
};

}  // namespace

struct sbg_store {
  sbg_spec spec;
  std::vector<sbg_col> cols;
  Arena arena;
  bool own_arena = false;
  std::vector<BlockOut> blocks;
  int64_t first_block = 0;
  bool overflow = false;
  int64_t encoded_bytes = 0;
};

namespace {

inline int64_t col_value(const sbg_col& c, uint64_t u, int64_t row) {
  switch (c.kind) {
    case SBG_UNIFORM:
    case SBG_STRKEY: return c.lo + (int64_t)(u % (uint64_t)c.span);
    case SBG_SUM4: {
      uint64_t s = (u & 0xffff) + ((u >> 16) & 0xffff) + ((u >> 32) & 0xffff) + ((u >> 48) & 0xffff);
      return c.lo + (int64_t)((s * (uint64_t)c.span) / (4ull * 65535ull));
    }
    case SBG_TIME: {
      __int128 adv = (__int128)row * (__int128)c.a / (__int128)c.b;
      return c.lo + (int64_t)adv + (int64_t)(c.span > 0 ? u % (uint64_t)c.span : 0);
    }
  }
  return 0;
}

// one block of one column: values + validity for rows [row0, row0+n)
void gen_rows(const sbg_spec& sp, const sbg_col& c, int64_t row0, int32_t n, std::vector<int64_t>& v,
              std::vector<uint8_t>& valid) {
  v.resize((size_t)n);
  valid.resize((size_t)n);
  for (int32_t i = 0; i < n; i++) {
    int64_t row = row0 + i;
    uint64_t u = splitmix64(sp.seed ^ ((uint64_t)c.col_slot << 40) ^ (uint64_t)row);
    v[(size_t)i] = col_value(c, u, row);
    valid[(size_t)i] = c.null_per_1024 > 0 ? ((splitmix64(u) & 1023) >= (uint64_t)c.null_per_1024) : 1;
  }
}

template <class T>
T* arena_copy(Arena& a, const std::vector<T>& src, bool& overflow) {
  if (src.empty()) return nullptr;
  void* p = a.alloc(src.size() * sizeof(T));
  if (!p) {
    overflow = true;
    return nullptr;
  }
  memcpy(p, src.data(), src.size() * sizeof(T));
  return (T*)p;
}

// Encode one column of one block following the reference's digest rules.
// `codes` are the values to bucket on (int value, or local string id).
void encode_column(sbg_store& st, const sbg_col& c, int32_t n, const std::vector<int64_t>& codes,
                   const std::vector<uint8_t>& valid, const std::vector<std::string>* dict, sg_column_desc& out,
                   int64_t& bytes) {
  memset(&out, 0, sizeof(out));
  out.col_slot = c.col_slot;
  out.col_type = c.col_type;
  const int thr = st.spec.cardinality_threshold;
  bool ovf = false;

  // distinct values among populated rows, with early exit above the threshold
  std::unordered_map<int64_t, uint32_t> seen;
  seen.reserve(8192);
  bool high_card = false;
  int32_t max_r = 0;
  int64_t npop = 0;
  for (int32_t i = 0; i < n; i++) {
    if (!valid[(size_t)i]) continue;
    npop++;
    max_r = i + 1;
    if (!high_card) {
      seen.emplace(codes[(size_t)i], 0);
      if ((int)seen.size() > thr) high_card = true;
    }
  }
  if (npop == 0) {  // no row has the field: the digest writes no file for it
    out.encoding = SG_ENC_ABSENT;
    return;
  }
  if (!high_card) {
    // bins ascending by value (Go map order is arbitrary; any order is valid input)
    std::vector<int64_t> keys;
    keys.reserve(seen.size());
    for (auto& kv : seen) keys.push_back(kv.first);
    std::sort(keys.begin(), keys.end());
    for (size_t b = 0; b < keys.size(); b++) seen[keys[b]] = (uint32_t)b;
    std::vector<uint32_t> offsets(keys.size() + 1, 0);
    std::vector<uint32_t> binof((size_t)n);
    for (int32_t i = 0; i < n; i++) {
      if (!valid[(size_t)i]) continue;
      uint32_t b = seen[codes[(size_t)i]];
      binof[(size_t)i] = b;
      offsets[b + 1]++;
    }
    for (size_t b = 0; b < keys.size(); b++) offsets[b + 1] += offsets[b];
    std::vector<uint32_t> cursor(offsets.begin(), offsets.end() - 1);
    std::vector<uint32_t> prev(keys.size(), 0);
    std::vector<uint32_t> ids((size_t)npop);
    for (int32_t i = 0; i < n; i++) {  // ascending rows -> delta_encode_col (:21-30)
      if (!valid[(size_t)i]) continue;
      uint32_t b = binof[(size_t)i];
      ids[cursor[b]++] = (uint32_t)i - prev[b];
      prev[b] = (uint32_t)i;
    }
    out.encoding = SG_ENC_BUCKET;
    out.delta_ids = 1;
    out.nbins = (uint32_t)keys.size();
    out.nrecord_ids = (uint32_t)npop;
    out.bin_values = arena_copy(st.arena, keys, ovf);
    out.bin_offsets = arena_copy(st.arena, offsets, ovf);
    if (st.spec.narrow) {  // a row id / gap inside a block is < 65,536
      std::vector<uint16_t> ids16(ids.begin(), ids.end());
      out.record_ids = reinterpret_cast<const uint32_t*>(arena_copy(st.arena, ids16, ovf));
      out.id_bits = 16;
      bytes += (int64_t)(keys.size() * 8 + offsets.size() * 4 + ids.size() * 2);
    } else {
      out.record_ids = arena_copy(st.arena, ids, ovf);
      bytes += (int64_t)(keys.size() * 8 + offsets.size() * 4 + ids.size() * 4);
    }
  } else {
    out.encoding = SG_ENC_VALUES;
    out.nvalues = (uint32_t)max_r;
    if (c.col_type == SG_COL_INT) {
      std::vector<int64_t> vals((size_t)max_r, 0);
      for (int32_t i = 0; i < max_r; i++)
        if (valid[(size_t)i]) vals[(size_t)i] = codes[(size_t)i];
      int64_t prev = 0;  // SaveIntsToColumns :109-113
      for (int32_t i = 0; i < max_r; i++) {
        int64_t val = vals[(size_t)i];
        vals[(size_t)i] = (int64_t)((uint64_t)val - (uint64_t)prev);
        prev = val;
      }
      out.delta_values = 1;
      // narrow form: the gaps as int16 / int32 when every one of them fits, on top of value_base = Values[0]
      int64_t dmin = 0, dmax = 0;
      for (int32_t i = 1; i < max_r; i++) {
        dmin = std::min(dmin, vals[(size_t)i]);
        dmax = std::max(dmax, vals[(size_t)i]);
      }
      if (st.spec.narrow && dmin >= INT16_MIN && dmax <= INT16_MAX) {
        std::vector<int16_t> d16((size_t)max_r, 0);
        for (int32_t i = 1; i < max_r; i++) d16[(size_t)i] = (int16_t)vals[(size_t)i];
        out.value_base = vals[0];
        out.value_bits = 16;
        out.values_i64 = reinterpret_cast<const int64_t*>(arena_copy(st.arena, d16, ovf));
        bytes += (int64_t)vals.size() * 2;
      } else if (st.spec.narrow && dmin >= INT32_MIN && dmax <= INT32_MAX) {
        std::vector<int32_t> d32((size_t)max_r, 0);
        for (int32_t i = 1; i < max_r; i++) d32[(size_t)i] = (int32_t)vals[(size_t)i];
        out.value_base = vals[0];
        out.value_bits = 32;
        out.values_i64 = reinterpret_cast<const int64_t*>(arena_copy(st.arena, d32, ovf));
        bytes += (int64_t)vals.size() * 4;
      } else {
        out.values_i64 = arena_copy(st.arena, vals, ovf);
        bytes += (int64_t)vals.size() * 8;
      }
    } else if (st.spec.narrow) {  // local string ids < len(StringTable) <= 65,536
      std::vector<uint16_t> vals((size_t)max_r, 0);
      for (int32_t i = 0; i < max_r; i++)
        if (valid[(size_t)i]) vals[(size_t)i] = (uint16_t)codes[(size_t)i];
      out.values_i32 = reinterpret_cast<const int32_t*>(arena_copy(st.arena, vals, ovf));
      out.value_bits = 16;
      bytes += (int64_t)vals.size() * 2;
    } else {
      std::vector<int32_t> vals((size_t)max_r, 0);
      for (int32_t i = 0; i < max_r; i++)
        if (valid[(size_t)i]) vals[(size_t)i] = (int32_t)codes[(size_t)i];
      out.values_i32 = arena_copy(st.arena, vals, ovf);
      bytes += (int64_t)vals.size() * 4;
    }
  }
  if (dict) {
    std::vector<char> bytes_;
    std::vector<uint32_t> offs(dict->size() + 1, 0);
    for (size_t k = 0; k < dict->size(); k++) {
      bytes_.insert(bytes_.end(), (*dict)[k].begin(), (*dict)[k].end());
      offs[k + 1] = (uint32_t)bytes_.size();
    }
    if (bytes_.empty()) bytes_.push_back(0);
    out.ndict = (uint32_t)dict->size();
    out.dict_bytes = arena_copy(st.arena, bytes_, ovf);
    out.dict_offsets = arena_copy(st.arena, offs, ovf);
  }
  if (ovf) st.overflow = true;
}

void gen_block(sbg_store& st, int64_t bi, BlockOut& bo, int64_t& bytes) {
  const sbg_spec& sp = st.spec;
  int64_t row0 = bi * (int64_t)sp.block_rows;
  int32_t n = (int32_t)std::min<int64_t>(sp.block_rows, sp.total_rows - row0);
  bo.cols.resize((size_t)sp.ncols);
  bo.info.clear();
  std::vector<int64_t> v;
  std::vector<uint8_t> valid;
  int ncols_out = 0;
  for (int ci = 0; ci < sp.ncols; ci++) {
    const sbg_col& c = st.cols[(size_t)ci];
    gen_rows(sp, c, row0, n, v, valid);
    sg_column_desc cd;
    if (c.col_type == SG_COL_STR) {
      // per-block dictionary in first-seen row order (get_val_id, table_column.go:27-48)
      std::unordered_map<int64_t, int32_t> local;
      std::vector<std::string> dict;
      std::vector<int64_t> codes((size_t)n, 0);
      for (int32_t i = 0; i < n; i++) {
        if (!valid[(size_t)i]) continue;
        auto it = local.find(v[(size_t)i]);
        if (it == local.end()) {
          it = local.emplace(v[(size_t)i], (int32_t)dict.size()).first;
          dict.push_back(std::string(c.prefix) + std::to_string(v[(size_t)i]));
        }
        codes[(size_t)i] = it->second;
      }
      encode_column(st, c, n, codes, valid, &dict, cd, bytes);
    } else {
      encode_column(st, c, n, v, valid, nullptr, cd, bytes);
      int64_t mn = INT64_MAX, mx = INT64_MIN;
      for (int32_t i = 0; i < n; i++)
        if (valid[(size_t)i]) {
          mn = std::min(mn, v[(size_t)i]);
          mx = std::max(mx, v[(size_t)i]);
        }
      if (mn <= mx) {  // block info.db IntInfoMap (exact extents; the reference's
                       // outlier-skipping update only ever narrows them)
        sg_int_info ii;
        memset(&ii, 0, sizeof(ii));
        ii.col_slot = c.col_slot;
        ii.min = mn;
        ii.max = mx;
        bo.info.push_back(ii);
      }
    }
    if (cd.encoding != SG_ENC_ABSENT) bo.cols[(size_t)ncols_out++] = cd;
  }
  bo.cols.resize((size_t)ncols_out);
  memset(&bo.desc, 0, sizeof(bo.desc));
  bo.desc.block_index = bi;
  bo.desc.num_records = n;
  bo.desc.ncols = ncols_out;
  bo.desc.cols = bo.cols.data();
  bo.desc.ninfo = (int32_t)bo.info.size();
  bo.desc.info = bo.info.data();
}

}  // namespace

extern "C" {

/* Generate blocks [first_block, first_block+nblocks) of the table described by spec
 * into `arena` (caller memory, e.g. pinned; NULL = malloc'd internally with
 * arena_bytes capacity).  Returns NULL if the arena is too small. */
sbg_store* sbg_generate(const sbg_spec* spec, int64_t first_block, int64_t nblocks, int nthreads, void* arena,
                        size_t arena_bytes) {
  sbg_store* st = new sbg_store();
  st->spec = *spec;
  st->cols.assign(spec->cols, spec->cols + spec->ncols);
  st->spec.cols = st->cols.data();
  if (st->spec.block_rows <= 0) st->spec.block_rows = SG_BLOCK_ROWS;
  if (st->spec.cardinality_threshold <= 0) st->spec.cardinality_threshold = 5000;
  st->first_block = first_block;
  if (arena) {
    st->arena.base = (char*)arena;
  } else {
    st->arena.base = (char*)aligned_alloc(128, (arena_bytes + 127) & ~(size_t)127);
    st->own_arena = true;
    if (!st->arena.base) {
      delete st;
      return nullptr;
    }
  }
  st->arena.cap = arena_bytes;
  st->blocks.resize((size_t)nblocks);
  std::atomic<int64_t> next(0);
  std::atomic<int64_t> total_bytes(0);
  if (nthreads < 1) nthreads = 1;
  auto worker = [&]() {
    for (;;) {
      int64_t i = next.fetch_add(1);
      if (i >= nblocks) break;
      int64_t bytes = 0;
      gen_block(*st, first_block + i, st->blocks[(size_t)i], bytes);
      total_bytes += bytes;
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; i++) th.emplace_back(worker);
  worker();
  for (auto& x : th) x.join();
  st->encoded_bytes = total_bytes.load();
  if (st->overflow) {
    if (st->own_arena) free(st->arena.base);
    delete st;
    return nullptr;
  }
  return st;
}
void sbg_free(sbg_store* st) {
  if (!st) return;
  if (st->own_arena) free(st->arena.base);
  delete st;
}
int64_t sbg_num_blocks(sbg_store* st) { return (int64_t)st->blocks.size(); }
const sg_block_desc* sbg_block(sbg_store* st, int64_t i) { return &st->blocks[(size_t)i].desc; }
int64_t sbg_arena_used(sbg_store* st) { return (int64_t)st->arena.used.load(); }
int64_t sbg_encoded_bytes(sbg_store* st) { return st->encoded_bytes; }
/* number of blocks a table of total_rows rows splits into */
int64_t sbg_total_blocks(const sbg_spec* spec) {
  int64_t br = spec->block_rows > 0 ? spec->block_rows : SG_BLOCK_ROWS;
  return (spec->total_rows + br - 1) / br;
}
/* decoded value of one cell, for spot checks (valid_out = 0 when the row lacks the field) */
int64_t sbg_cell(const sbg_spec* spec, int32_t col_index, int64_t row, int32_t* valid_out) {
  const sbg_col& c = spec->cols[col_index];
  uint64_t u = splitmix64(spec->seed ^ ((uint64_t)c.col_slot << 40) ^ (uint64_t)row);
  if (valid_out) *valid_out = c.null_per_1024 > 0 ? ((splitmix64(u) & 1023) >= (uint64_t)c.null_per_1024) : 1;
  return col_value(c, u, row);
}


/* ---- direct evaluation of a query on the generator's ROW VALUES ---------------------------------
 * Independent of the block encoding, of the GPU decode and of the oracle: every row's cells are
 * recomputed from (seed, column, row) and filtered / grouped / aggregated with plain loops.  bench.py
 * and the full-size GPU tests compare the engine's merged result with these arrays (per group: Count;
 * per aggregation: hist Count, exact sum, every bucket counter).  Columns must be null-free.
 * Semantics: IntFilter gt/lt/eq/neq on the cell value (filter.go:177-189; for a string
 * key column the value is the key's number, so eq/neq on "v3" is value 3); group slot = mixed radix
 * over (value - lo) of the group columns, then the time bucket index v / bucket - time_first
 * (aggregate.go:146-183); BasicHist.AddWeightedValue (hist_basic.go:101-151): reject v > max*10 or
 * v < min, bucket (v - min) / bsize clamped into [0, nvals-1]. */
typedef struct sbg_eval_spec {
  int32_t nfilters;
  int32_t filter_col[8]; /* index into spec->cols */
  int32_t filter_op[8];  /* sg_filter_op: GT, LT, EQ, NEQ */
  int64_t filter_val[8];
  int32_t ngroups;
  int32_t group_col[4];
  int32_t time_col; /* -1: none */
  int32_t naggs;
  int64_t time_bucket, time_first;
  int32_t time_n; /* number of time buckets on the axis */
  int32_t nvals;  /* bucket counters per (slot, aggregation); 0: no buckets (avg mode) */
  int32_t agg_col[16];
  int64_t info_min[16], info_max[16], bsize[16];
} sbg_eval_spec;

/* slots = prod(span of group cols) * (time_n or 1).  out_count[slots]; out_hcount, out_sum [naggs][slots];
 * out_buckets [naggs][slots][nvals].  Returns the number of rows that passed the filters, -1 on bad input. */
int64_t sbg_eval(const sbg_spec* spec, const sbg_eval_spec* ev, int64_t row0, int64_t row1, int nthreads, uint64_t* out_count,
                 uint64_t* out_hcount, uint64_t* out_sum, uint64_t* out_buckets) {
  const sbg_col* cols = spec->cols;
  uint64_t slots = 1;
  uint64_t gstride[4] = {0, 0, 0, 0};
  for (int g = 0; g < ev->ngroups; g++) {
    const sbg_col& c = cols[ev->group_col[g]];
    if (c.kind != SBG_UNIFORM && c.kind != SBG_STRKEY) return -1;
    gstride[g] = slots;
    slots *= (uint64_t)c.span;
  }
  const uint64_t tstride = slots;
  if (ev->time_col >= 0) slots *= (uint64_t)ev->time_n;
  const int na = ev->naggs;
  const uint64_t nv = (uint64_t)ev->nvals;
  const uint64_t words = slots * (1 + (uint64_t)na * (2 + nv));
  const bool priv = words <= ((uint64_t)4 << 20);
  if (nthreads < 1) nthreads = 1;
  memset(out_count, 0, slots * 8);
  if (na) {
    memset(out_hcount, 0, slots * (uint64_t)na * 8);
    memset(out_sum, 0, slots * (uint64_t)na * 8);
    if (nv) memset(out_buckets, 0, slots * (uint64_t)na * nv * 8);
  }
  std::atomic<int64_t> matched(0);
  std::atomic<int> bad(0);
  auto cell = [&](int ci, int64_t row) {
    const sbg_col& c = cols[ci];
    return col_value(c, splitmix64(spec->seed ^ ((uint64_t)c.col_slot << 40) ^ (uint64_t)row), row);
  };
  auto worker = [&](int tix) {
    const int64_t n = row1 - row0, per = (n + nthreads - 1) / nthreads;
    const int64_t a = row0 + per * tix, b = std::min<int64_t>(row1, a + per);
    std::vector<uint64_t> loc;
    uint64_t *cnt = out_count, *hc = out_hcount, *sm = out_sum, *bk = out_buckets;
    if (priv) {
      loc.assign(words, 0);
      cnt = loc.data();
      hc = cnt + slots;
      sm = hc + slots * (uint64_t)na;
      bk = sm + slots * (uint64_t)na;
    }
    auto add = [&](uint64_t* p, uint64_t v) {
      if (priv)
        *p += v;
      else
        __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
    };
    int64_t m = 0;
    for (int64_t row = a; row < b; row++) {
      bool pass = true;
      for (int f = 0; f < ev->nfilters && pass; f++) {
        const int64_t v = cell(ev->filter_col[f], row), lit = ev->filter_val[f];
        switch (ev->filter_op[f]) {
          case SG_OP_GT: pass = v > lit; break;
          case SG_OP_LT: pass = v < lit; break;
          case SG_OP_EQ: pass = v == lit; break;
          case SG_OP_NEQ: pass = v != lit; break;
          default: pass = false;
        }
      }
      if (!pass) continue;
      m++;
      uint64_t slot = 0;
      for (int g = 0; g < ev->ngroups; g++) slot += (uint64_t)(cell(ev->group_col[g], row) - cols[ev->group_col[g]].lo) * gstride[g];
      if (ev->time_col >= 0) {
        const int64_t q = cell(ev->time_col, row) / ev->time_bucket - ev->time_first;
        if (q < 0 || q >= ev->time_n) {
          bad = 1;
          continue;
        }
        slot += (uint64_t)q * tstride;
      }
      add(cnt + slot, 1);
      for (int ai = 0; ai < na; ai++) {
        const int64_t v = cell(ev->agg_col[ai], row);
        if (v > ev->info_max[ai] * 10 || v < ev->info_min[ai]) continue;
        add(hc + (uint64_t)ai * slots + slot, 1);
        add(sm + (uint64_t)ai * slots + slot, (uint64_t)v);
        if (nv) {
          int64_t bi = (v - ev->info_min[ai]) / ev->bsize[ai];
          if (bi >= (int64_t)nv) bi = (int64_t)nv - 1;
          if (bi < 0) bi = 0;
          add(bk + ((uint64_t)ai * slots + slot) * nv + (uint64_t)bi, 1);
        }
      }
    }
    matched += m;
    if (priv) {
      static std::mutex mu;
      std::lock_guard<std::mutex> lk(mu);
      for (uint64_t i = 0; i < slots; i++) out_count[i] += cnt[i];
      for (uint64_t i = 0; i < slots * (uint64_t)na; i++) {
        out_hcount[i] += hc[i];
        out_sum[i] += sm[i];
      }
      for (uint64_t i = 0; i < slots * (uint64_t)na * nv; i++) out_buckets[i] += bk[i];
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; i++) th.emplace_back(worker, i);
  worker(0);
  for (auto& x : th) x.join();
  return bad.load() ? -1 : matched.load();
}

}  // extern "C"
