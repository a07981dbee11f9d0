// gobread.cpp — sybil block directory -> sg_block_desc, without Go (include/sybilgob.h).
//
// A small Go `encoding/gob` reader: it parses the type definitions that precede the values in the
// stream and binds struct fields BY NAME (a receiver must not rely on field positions), then pulls out
// the fields of SavedIntColumn / SavedStrColumn / SavedColumnInfo (src/lib/column_store.go:39-64) and
// skips everything else.  Wire format: Go's encoding/gob documentation; SURVEY.md appendix A.
// The Python reader (sybil_b200/gob.py) is pinned to Go's own output by the reference's golden gob
// files; tests/test_gobread.py checks this reader against it on block directories.
#include <dirent.h>
#include <sys/stat.h>
#include <zlib.h>

#include <climits>
#include <limits>
#include <cstdint>
#include <algorithm>

#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/sybilgob.h"

namespace {

enum { T_BOOL = 1, T_INT = 2, T_UINT = 3, T_FLOAT = 4, T_BYTES = 5, T_STRING = 6, T_COMPLEX = 7, T_INTERFACE = 8,
       T_WIRETYPE = 16, T_ARRAYTYPE = 17, T_COMMONTYPE = 18, T_SLICETYPE = 19, T_STRUCTTYPE = 20, T_FIELDTYPE = 21,
       T_FIELDTYPE_SLICE = 22, T_MAPTYPE = 23, T_GOBENC = 24, T_BINMARSH = 25, T_TEXTMARSH = 26 };

struct Err : std::runtime_error {
  using std::runtime_error::runtime_error;
};

struct TypeDef {
  enum Kind { NONE, STRUCT, SLICE, ARRAY, MAP, OPAQUE } kind = NONE;
  std::vector<std::pair<std::string, int64_t>> fields;  // STRUCT
  int64_t elem = 0, key = 0;
};

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  std::map<int64_t, TypeDef> types;

  Reader(const uint8_t* b, size_t n) : p(b), end(b + n) {
    auto st = [&](int64_t id, std::initializer_list<std::pair<const char*, int64_t>> f) {
      TypeDef t;
      t.kind = TypeDef::STRUCT;
      for (auto& x : f) t.fields.emplace_back(x.first, x.second);
      types[id] = t;
    };
    st(T_WIRETYPE, {{"ArrayT", T_ARRAYTYPE}, {"SliceT", T_SLICETYPE}, {"StructT", T_STRUCTTYPE}, {"MapT", T_MAPTYPE},
                    {"GobEncoderT", T_GOBENC}, {"BinaryMarshalerT", T_BINMARSH}, {"TextMarshalerT", T_TEXTMARSH}});
    st(T_ARRAYTYPE, {{"CommonType", T_COMMONTYPE}, {"Elem", T_INT}, {"Len", T_INT}});
    st(T_COMMONTYPE, {{"Name", T_STRING}, {"Id", T_INT}});
    st(T_SLICETYPE, {{"CommonType", T_COMMONTYPE}, {"Elem", T_INT}});
    st(T_STRUCTTYPE, {{"CommonType", T_COMMONTYPE}, {"Field", T_FIELDTYPE_SLICE}});
    st(T_FIELDTYPE, {{"Name", T_STRING}, {"Id", T_INT}});
    st(T_MAPTYPE, {{"CommonType", T_COMMONTYPE}, {"Key", T_INT}, {"Elem", T_INT}});
    st(T_GOBENC, {{"CommonType", T_COMMONTYPE}});
    st(T_BINMARSH, {{"CommonType", T_COMMONTYPE}});
    st(T_TEXTMARSH, {{"CommonType", T_COMMONTYPE}});
    TypeDef fs;
    fs.kind = TypeDef::SLICE;
    fs.elem = T_FIELDTYPE;
    types[T_FIELDTYPE_SLICE] = fs;
  }

  bool eof() const { return p >= end; }
  uint64_t u() {
    if (p >= end) throw Err("gob: unexpected end of stream");
    uint8_t c = *p++;
    if (c < 128) return c;
    unsigned n = 256u - c;
    if (n > 8 || p + n > end) throw Err("gob: bad uint");
    uint64_t v = 0;
    for (unsigned i = 0; i < n; i++) v = (v << 8) | *p++;
    return v;
  }
  int64_t i() {
    uint64_t x = u();
    return (x & 1) ? (int64_t)~(x >> 1) : (int64_t)(x >> 1);
  }
  std::string str() {
    uint64_t n = u();
    if ((uint64_t)(end - p) < n) throw Err("gob: bad length");
    std::string s((const char*)p, (size_t)n);
    p += n;
    return s;
  }
  const TypeDef& def(int64_t tid) {
    auto it = types.find(tid);
    if (it == types.end()) throw Err("gob: value of undefined type " + std::to_string(tid));
    return it->second;
  }
  // struct: f(field name, field type id) must consume the value (or call skip)
  template <class F>
  void structure(int64_t tid, F f) {
    const TypeDef& t = def(tid);
    if (t.kind != TypeDef::STRUCT) throw Err("gob: expected a struct");
    int64_t idx = -1;
    for (;;) {
      uint64_t d = u();
      if (d == 0) return;
      idx += (int64_t)d;
      if (idx >= (int64_t)t.fields.size()) throw Err("gob: field number out of range");
      f(t.fields[(size_t)idx].first, t.fields[(size_t)idx].second);
    }
  }
  int depth = 0;  // nesting of skip(): a crafted stream with self-referential type ids must not overflow the stack
  struct Nest {
    int& d;
    explicit Nest(int& x) : d(x) {
      if (++d > 64) throw Err("gob: value nested too deeply");
    }
    ~Nest() { --d; }
  };
  void skip(int64_t tid) {
    Nest nest(depth);
    switch (tid) {
      case T_BOOL: case T_INT: case T_UINT: case T_FLOAT: u(); return;
      case T_BYTES: case T_STRING: str(); return;
      case T_COMPLEX: u(); u(); return;
      case T_INTERFACE: throw Err("gob: interface values are not expected in column files");
      default: break;
    }
    const TypeDef& t = def(tid);
    if (t.kind == TypeDef::STRUCT) {
      structure(tid, [&](const std::string&, int64_t ft) { skip(ft); });
    } else if (t.kind == TypeDef::SLICE || t.kind == TypeDef::ARRAY) {
      uint64_t n = u();
      for (uint64_t k = 0; k < n; k++) skip(t.elem);
    } else if (t.kind == TypeDef::MAP) {
      uint64_t n = u();
      for (uint64_t k = 0; k < n; k++) {
        skip(t.key);
        skip(t.elem);
      }
    } else if (t.kind == TypeDef::OPAQUE) {
      str();
    } else {
      throw Err("gob: unknown type kind");
    }
  }
  void define(int64_t tid) {
    TypeDef nt;
    structure(T_WIRETYPE, [&](const std::string& which, int64_t wt) {
      if (which == "GobEncoderT" || which == "BinaryMarshalerT" || which == "TextMarshalerT") {
        nt.kind = TypeDef::OPAQUE;
        skip(wt);
        return;
      }
      nt.kind = which == "StructT" ? TypeDef::STRUCT : which == "SliceT" ? TypeDef::SLICE : which == "ArrayT" ? TypeDef::ARRAY : TypeDef::MAP;
      structure(wt, [&](const std::string& fn, int64_t ft) {
        if (fn == "Elem") nt.elem = i();
        else if (fn == "Key") nt.key = i();
        else if (fn == "Field") {
          uint64_t n = u();
          for (uint64_t k = 0; k < n; k++) {
            std::string name;
            int64_t id = 0;
            structure(T_FIELDTYPE, [&](const std::string& a, int64_t at) {
              if (a == "Name") name = str();
              else if (a == "Id") id = i();
              else skip(at);
            });
            nt.fields.emplace_back(name, id);
          }
        } else {
          skip(ft);
        }
      });
    });
    if (nt.kind == TypeDef::NONE) throw Err("gob: empty wireType");
    types[tid] = nt;
  }
  // positions the reader at the first top-level value; returns its type id
  int64_t next_value() {
    for (;;) {
      u();  // message length
      int64_t tid = i();
      if (tid < 0) {
        define(-tid);
        continue;
      }
      if (def_kind(tid) != TypeDef::STRUCT && u() != 0) throw Err("gob: expected the 0 marker of a non-struct value");
      return tid;
    }
  }
  TypeDef::Kind def_kind(int64_t tid) {
    auto it = types.find(tid);
    return it == types.end() ? TypeDef::NONE : it->second.kind;
  }
  template <class T>
  void ints(int64_t tid, std::vector<T>& out, bool is_unsigned) {
    const TypeDef& t = def(tid);
    if (t.kind != TypeDef::SLICE && t.kind != TypeDef::ARRAY) throw Err("gob: expected a slice");
    uint64_t n = u();
    // every element takes at least one byte: a length prefix beyond the bytes left is corrupt (and must not reserve)
    if (n > (uint64_t)(end - p)) throw Err("gob: slice longer than the stream");
    out.reserve(out.size() + (size_t)n);
    // Go's decoder raises an overflow error when a value does not fit the destination type: so does this one
    if (t.elem == T_UINT || (is_unsigned && t.elem != T_INT)) {
      for (uint64_t k = 0; k < n; k++) {
        const uint64_t v = u();
        if (v > (uint64_t)std::numeric_limits<T>::max()) throw Err("gob: value overflows the column's element type");
        out.push_back((T)v);
      }
    } else {
      for (uint64_t k = 0; k < n; k++) {
        const int64_t v = i();
        if (v < (int64_t)std::numeric_limits<T>::min() || (v > 0 && (uint64_t)v > (uint64_t)std::numeric_limits<T>::max()))
          throw Err("gob: value overflows the column's element type");
        out.push_back((T)v);
      }
    }
  }
};

bool read_file(const std::string& path, std::vector<uint8_t>& out) {
  FILE* f = fopen(path.c_str(), "rb");
  if (f) {
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    out.resize((size_t)std::max(n, 0l));
    size_t got = n > 0 ? fread(out.data(), 1, (size_t)n, f) : 0;
    fclose(f);
    out.resize(got);
    return true;
  }
  gzFile g = gzopen((path + ".gz").c_str(), "rb");
  if (!g) return false;
  out.clear();
  uint8_t buf[1 << 16];
  int n;
  while ((n = gzread(g, buf, sizeof(buf))) > 0) out.insert(out.end(), buf, buf + n);
  gzclose(g);
  return n == 0;
}

struct Column {
  int32_t slot = 0, type = 0, encoding = SG_ENC_ABSENT;
  bool delta_ids = false, delta_values = false;
  std::vector<int64_t> bin_values;
  std::vector<uint32_t> bin_offsets{0};
  std::vector<uint32_t> record_ids;
  std::vector<int64_t> values_i64;
  std::vector<int32_t> values_i32;
  std::string dict_bytes;
  std::vector<uint32_t> dict_offsets{0};
  uint32_t ndict = 0;
  // narrow copies (sgob_set_narrow): what the varints of the file hold, not widened to Go's types
  std::vector<uint16_t> ids16, v16;
  std::vector<int32_t> d32;
  std::vector<int16_t> d16;
  int64_t vbase = 0;
  int id_bits = 0, value_bits = 0;
  uint32_t set_nvalues = 0;  // set column read from its non-bucketed form: len(Values)
};

bool g_narrow = false;

// keep the arrays as narrow as their values allow (include/sybilgpu.h: id_bits / value_bits)
void narrow_column(Column& c) {
  if (c.encoding == SG_ENC_BUCKET) {
    bool ok = true;
    for (uint32_t v : c.record_ids) ok = ok && v < 65536u;
    if (ok) {
      c.ids16.assign(c.record_ids.begin(), c.record_ids.end());
      c.id_bits = 16;
    }
  } else if (c.encoding == SG_ENC_VALUES && c.type == SG_COL_INT) {
    if (!c.delta_values || c.values_i64.empty()) return;
    int64_t lo = 0, hi = 0;
    for (size_t k = 1; k < c.values_i64.size(); k++) {
      lo = std::min(lo, c.values_i64[k]);
      hi = std::max(hi, c.values_i64[k]);
    }
    c.vbase = c.values_i64[0];
    if (lo >= INT16_MIN && hi <= INT16_MAX) {
      c.d16.assign(c.values_i64.size(), 0);
      for (size_t k = 1; k < c.values_i64.size(); k++) c.d16[k] = (int16_t)c.values_i64[k];
      c.value_bits = 16;
    } else if (lo >= INT32_MIN && hi <= INT32_MAX) {
      c.d32.assign(c.values_i64.size(), 0);
      for (size_t k = 1; k < c.values_i64.size(); k++) c.d32[k] = (int32_t)c.values_i64[k];
      c.value_bits = 32;
    }
  } else if (c.encoding == SG_ENC_VALUES) {
    bool ok = true;
    for (int32_t v : c.values_i32) ok = ok && v >= 0 && v < 65536;
    if (ok) {
      c.v16.assign(c.values_i32.begin(), c.values_i32.end());
      c.value_bits = 16;
    }
  }
}

void read_column(const std::vector<uint8_t>& raw, Column& c) {
  Reader r(raw.data(), raw.size());
  const int64_t tid = r.next_value();
  bool bucket = false;
  std::vector<std::vector<int32_t>> set_rows;
  r.structure(tid, [&](const std::string& f, int64_t ft) {
    if (f == "DeltaEncodedIDs") c.delta_ids = r.u() != 0;
    else if (f == "ValueEncoded") c.delta_values = r.u() != 0;
    else if (f == "BucketEncoded") bucket = r.u() != 0;
    else if (f == "Bins") {
      const TypeDef& st = r.def(ft);
      if (st.kind != TypeDef::SLICE && st.kind != TypeDef::ARRAY) throw Err("gob: Bins is not a slice");
      uint64_t n = r.u();
      for (uint64_t k = 0; k < n; k++) {
        int64_t value = 0;
        r.structure(st.elem, [&](const std::string& bf, int64_t bt) {
          if (bf == "Value") value = r.i();
          else if (bf == "Records") r.ints(bt, c.record_ids, true);
          else r.skip(bt);
        });
        c.bin_values.push_back(value);
        c.bin_offsets.push_back((uint32_t)c.record_ids.size());
      }
    } else if (f == "Values" && c.type == SG_COL_SET) {
      // SavedSetColumn.Values [][]int32 (more than 5,000 distinct tags, column_store_io.go:183-192): kept per row
      // here, turned into bins below (the C ABI takes set columns in the bucket form, sybilgpu.h)
      const TypeDef& st = r.def(ft);
      if (st.kind != TypeDef::SLICE && st.kind != TypeDef::ARRAY) throw Err("gob: set Values is not a slice");
      uint64_t n = r.u();
      if (n > (uint64_t)SG_BLOCK_ROWS) throw Err("gob: set Values longer than a block");
      set_rows.resize((size_t)n);
      for (uint64_t k = 0; k < n; k++) r.ints(st.elem, set_rows[(size_t)k], false);
    } else if (f == "Values") {
      if (c.type == SG_COL_INT) r.ints(ft, c.values_i64, false);
      else r.ints(ft, c.values_i32, false);
    } else if (f == "StringTable") {
      uint64_t n = r.u();
      for (uint64_t k = 0; k < n; k++) {
        c.dict_bytes += r.str();
        c.dict_offsets.push_back((uint32_t)c.dict_bytes.size());
      }
      c.ndict = (uint32_t)n;
    } else {
      r.skip(ft);
    }
  });
  c.encoding = bucket ? SG_ENC_BUCKET : SG_ENC_VALUES;
  if (!bucket) {
    c.bin_values.clear();
    c.bin_offsets.assign(1, 0);
    c.record_ids.clear();
  }
  if (c.type == SG_COL_SET && !bucket) {
    // rows -> bins (tag id ascending, rows ascending, gaps); len(Values) travels as nvalues: the reference marks
    // every listed row populated, empty set or not (column_store_io.go:672-682)
    std::map<int32_t, std::vector<uint32_t>> rows_of;
    for (size_t r2 = 0; r2 < set_rows.size(); r2++)
      for (int32_t tag : set_rows[r2]) {
        auto& v = rows_of[tag];
        if (v.empty() || v.back() != (uint32_t)r2) v.push_back((uint32_t)r2);
      }
    for (auto& kv : rows_of) {
      uint32_t prev = 0;
      for (size_t k = 0; k < kv.second.size(); k++) {
        c.record_ids.push_back(k == 0 ? kv.second[k] : kv.second[k] - prev);
        prev = kv.second[k];
      }
      c.bin_values.push_back(kv.first);
      c.bin_offsets.push_back((uint32_t)c.record_ids.size());
    }
    c.delta_ids = true;
    c.encoding = SG_ENC_BUCKET;
    c.set_nvalues = (uint32_t)set_rows.size();
  }
}

}  // namespace

struct sgob_block {
  std::vector<Column> cols;
  std::vector<sg_column_desc> descs;
  std::vector<sg_int_info> info;
  sg_block_desc desc;
  int64_t bytes = 0;
};

struct sgob_table {
  std::vector<std::string> names;
  std::vector<int32_t> types;
  std::vector<char> has_info;
  std::vector<std::pair<int64_t, int64_t>> info;
  std::vector<std::string> blocks;
};

extern "C" {

void sgob_set_narrow(int on) { g_narrow = on != 0; }

sgob_block* sgob_read_block_dir(const char* dir, const char* const* col_names, const int32_t* col_types, int32_t ncols,
                                const uint8_t* load_mask, int64_t block_index, char* err, size_t errlen) {
  auto fail = [&](const std::string& m) -> sgob_block* {
    if (err && errlen) snprintf(err, errlen, "%s", m.c_str());
    return nullptr;
  };
  if (!dir || ncols < 0 || (ncols > 0 && (!col_names || !col_types))) return fail("sgob_read_block_dir: bad arguments");
  std::unique_ptr<sgob_block> b(new sgob_block());
  try {
    std::vector<uint8_t> raw;
    const std::string d(dir);
    if (!read_file(d + "/info.db", raw)) return fail(d + "/info.db: cannot be read");
    int64_t num_records = 0;
    {
      Reader r(raw.data(), raw.size());
      const int64_t tid = r.next_value();
      r.structure(tid, [&](const std::string& f, int64_t ft) {
        if (f == "NumRecords") num_records = r.i();
        else if (f == "IntInfoMap") {
          const TypeDef& mt = r.def(ft);
          uint64_t n = r.u();
          for (uint64_t k = 0; k < n; k++) {
            if (mt.key != T_STRING) throw Err("info.db: IntInfoMap key is not a string");
            const std::string name = r.str();
            sg_int_info ii;
            memset(&ii, 0, sizeof(ii));
            ii.col_slot = -1;
            r.structure(mt.elem, [&](const std::string& a, int64_t at) {
              if (a == "Min") ii.min = r.i();
              else if (a == "Max") ii.max = r.i();
              else r.skip(at);
            });
            for (int32_t s = 0; s < ncols; s++)
              if (name == col_names[s]) ii.col_slot = s;
            if (ii.col_slot >= 0) b->info.push_back(ii);
          }
        } else {
          r.skip(ft);
        }
      });
    }
    for (int32_t s = 0; s < ncols; s++) {
      if (load_mask && !load_mask[s]) continue;
      if (col_types[s] != SG_COL_INT && col_types[s] != SG_COL_STR && col_types[s] != SG_COL_SET) continue;
      const std::string path =
          d + "/" + (col_types[s] == SG_COL_INT ? "int_" : (col_types[s] == SG_COL_STR ? "str_" : "set_")) + col_names[s] + ".db";
      if (!read_file(path, raw)) continue;  // the block does not hold that column
      Column c;
      c.slot = s;
      c.type = col_types[s];
      read_column(raw, c);
      if (g_narrow) narrow_column(c);
      b->cols.push_back(std::move(c));
    }
    b->descs.resize(b->cols.size());
    for (size_t k = 0; k < b->cols.size(); k++) {
      const Column& c = b->cols[k];
      sg_column_desc& cd = b->descs[k];
      memset(&cd, 0, sizeof(cd));
      cd.col_slot = c.slot;
      cd.col_type = c.type;
      cd.encoding = c.encoding;
      cd.delta_ids = c.delta_ids ? 1 : 0;
      cd.delta_values = c.delta_values ? 1 : 0;
      if (c.encoding == SG_ENC_BUCKET) {
        cd.nbins = (uint32_t)c.bin_values.size();
        cd.nrecord_ids = (uint32_t)c.record_ids.size();
        cd.bin_values = c.bin_values.data();
        cd.bin_offsets = c.bin_offsets.data();
        cd.record_ids = c.id_bits == 16 ? reinterpret_cast<const uint32_t*>(c.ids16.data()) : c.record_ids.data();
        cd.id_bits = c.id_bits;
        b->bytes += (int64_t)c.record_ids.size() * (c.id_bits == 16 ? 2 : 4) + (int64_t)c.bin_values.size() * 12;
        if (c.type == SG_COL_SET) cd.nvalues = c.set_nvalues;
      } else if (c.type == SG_COL_INT) {
        cd.nvalues = (uint32_t)c.values_i64.size();
        cd.value_bits = c.value_bits;
        cd.value_base = c.value_bits ? c.vbase : 0;
        cd.values_i64 = c.value_bits == 16   ? reinterpret_cast<const int64_t*>(c.d16.data())
                        : c.value_bits == 32 ? reinterpret_cast<const int64_t*>(c.d32.data())
                                             : c.values_i64.data();
        b->bytes += (int64_t)c.values_i64.size() * (c.value_bits ? c.value_bits / 8 : 8);
      } else {
        cd.nvalues = (uint32_t)c.values_i32.size();
        cd.value_bits = c.value_bits;
        cd.values_i32 = c.value_bits == 16 ? reinterpret_cast<const int32_t*>(c.v16.data()) : c.values_i32.data();
        b->bytes += (int64_t)c.values_i32.size() * (c.value_bits == 16 ? 2 : 4);
      }
      if (c.type == SG_COL_STR || c.type == SG_COL_SET) {
        cd.ndict = c.ndict;
        cd.dict_bytes = c.dict_bytes.empty() ? "" : c.dict_bytes.data();
        cd.dict_offsets = c.dict_offsets.data();
      }
    }
    memset(&b->desc, 0, sizeof(b->desc));
    b->desc.block_index = block_index;
    b->desc.num_records = (int32_t)num_records;
    b->desc.ncols = (int32_t)b->descs.size();
    b->desc.cols = b->descs.data();
    b->desc.ninfo = (int32_t)b->info.size();
    b->desc.info = b->info.data();
  } catch (const std::exception& e) {
    return fail(std::string(dir) + ": " + e.what());
  }
  return b.release();
}

sgob_table* sgob_table_open(const char* dbdir, const char* table, char* err, size_t errlen) {
  auto fail = [&](const std::string& m) -> sgob_table* {
    if (err && errlen) snprintf(err, errlen, "%s", m.c_str());
    return nullptr;
  };
  if (!dbdir || !table) return fail("sgob_table_open: bad arguments");
  std::unique_ptr<sgob_table> t(new sgob_table());
  const std::string tdir = std::string(dbdir) + "/" + table;
  try {
    std::vector<uint8_t> raw;
    if (!read_file(tdir + "/info.db", raw)) return fail(tdir + "/info.db: cannot be read");
    std::map<int64_t, std::string> by_slot;
    std::map<int64_t, int64_t> types;
    std::map<int64_t, std::pair<int64_t, int64_t>> info;
    Reader r(raw.data(), raw.size());
    const int64_t tid = r.next_value();
    r.structure(tid, [&](const std::string& f, int64_t ft) {
      if (f == "KeyTable") {
        const TypeDef& mt = r.def(ft);
        if (mt.key != T_STRING) throw Err("info.db: KeyTable key is not a string");
        uint64_t n = r.u();
        for (uint64_t k = 0; k < n; k++) {
          std::string name = r.str();
          by_slot[r.i()] = name;
        }
      } else if (f == "KeyTypes") {
        uint64_t n = r.u();
        for (uint64_t k = 0; k < n; k++) {
          int64_t slot = r.i();
          types[slot] = r.i();
        }
      } else if (f == "IntInfo") {
        const TypeDef& mt = r.def(ft);
        uint64_t n = r.u();
        for (uint64_t k = 0; k < n; k++) {
          int64_t slot = r.i();
          std::pair<int64_t, int64_t> mm(0, 0);
          r.structure(mt.elem, [&](const std::string& a, int64_t at) {
            if (a == "Min") mm.first = r.i();
            else if (a == "Max") mm.second = r.i();
            else r.skip(at);
          });
          info[slot] = mm;
        }
      } else {
        r.skip(ft);
      }
    });
    int64_t nslots = by_slot.empty() ? 0 : by_slot.rbegin()->first + 1;
    if (nslots < 0 || nslots > 32768) return fail(tdir + "/info.db: key slots out of range");
    t->names.assign((size_t)nslots, "");
    t->types.assign((size_t)nslots, 0);
    t->has_info.assign((size_t)nslots, 0);
    t->info.assign((size_t)nslots, {0, 0});
    for (auto& kv : by_slot)
      if (kv.first >= 0) t->names[(size_t)kv.first] = kv.second;
    for (auto& kv : types)
      if (kv.first >= 0 && kv.first < nslots && (kv.second == SG_COL_INT || kv.second == SG_COL_STR)) t->types[(size_t)kv.first] = (int32_t)kv.second;
    for (auto& kv : info)
      if (kv.first >= 0 && kv.first < nslots) {
        t->has_info[(size_t)kv.first] = 1;
        t->info[(size_t)kv.first] = kv.second;
      }
  } catch (const std::exception& e) {
    return fail(tdir + "/info.db: " + e.what());
  }
  DIR* d = opendir(tdir.c_str());
  if (!d) return fail(tdir + ": cannot be listed");
  auto ends_with = [](const std::string& s, const char* suf) {
    const size_t n = strlen(suf);
    return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
  };
  while (dirent* e = readdir(d)) {
    const std::string n = e->d_name;
    if (n == "." || n == ".." || n == "ingest" || n == ".ingest.temp" || n == "cache" || n.rfind("stomache", 0) == 0) continue;
    bool bad = false;
    for (const char* suf : {"info.db", "old", "broken", "lock", "export", "partial"}) bad = bad || ends_with(n, suf);
    if (bad) continue;
    struct stat st;
    if (stat((tdir + "/" + n).c_str(), &st) != 0 || !S_ISDIR(st.st_mode)) continue;
    t->blocks.push_back(tdir + "/" + n);
  }
  closedir(d);
  std::sort(t->blocks.begin(), t->blocks.end());
  return t.release();
}
void sgob_table_free(sgob_table* t) { delete t; }
int32_t sgob_table_num_cols(const sgob_table* t) { return t ? (int32_t)t->names.size() : 0; }
const char* sgob_table_col_name(const sgob_table* t, int32_t s) { return t && s >= 0 && (size_t)s < t->names.size() ? t->names[(size_t)s].c_str() : ""; }
int32_t sgob_table_col_type(const sgob_table* t, int32_t s) { return t && s >= 0 && (size_t)s < t->types.size() ? t->types[(size_t)s] : 0; }
int32_t sgob_table_int_info(const sgob_table* t, int32_t s, int64_t* mn, int64_t* mx) {
  if (!t || s < 0 || (size_t)s >= t->has_info.size() || !t->has_info[(size_t)s]) return 0;
  if (mn) *mn = t->info[(size_t)s].first;
  if (mx) *mx = t->info[(size_t)s].second;
  return 1;
}
int64_t sgob_table_num_blocks(const sgob_table* t) { return t ? (int64_t)t->blocks.size() : 0; }
const char* sgob_table_block_dir(const sgob_table* t, int64_t i) { return t && i >= 0 && (size_t)i < t->blocks.size() ? t->blocks[(size_t)i].c_str() : ""; }

const sg_block_desc* sgob_block_desc(const sgob_block* b) { return b ? &b->desc : nullptr; }
int64_t sgob_block_bytes(const sgob_block* b) { return b ? b->bytes : 0; }
void sgob_block_free(sgob_block* b) { delete b; }

}  // extern "C"
