/*
 * oracle.cpp — TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * A CPU restatement of logv/sybil's query hot path, written to be read side by
 * side with the Go sources it follows (cited per function as file:line under
 * /root/reference/src/lib).  It exists so the CUDA path has something to be
 * checked against: the Go toolchain is absent from the build image and the GPU
 * box, so the reference itself cannot run.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load this library;
 * nothing under sybil_b200/ links, imports or calls it.
 *
 * Pinned by: testdata/TestDecodeGoldenFiles/node_results.golden.json of the
 * reference (bucket layout 23470/1000 -> BucketSize 23, NumBuckets 1001,
 * len(Values) 1002; Combine; TOTAL key; sort) via tests/test_oracle_golden.py.
 * Column decode, filters, group-by on raw rows, MultiHist and the time rollup have
 * no golden vector in the reference: for those rows parity is "unpinned by
 * reference data" and rests on this line-by-line restatement plus the independent
 * Python mirror in oracle/pyoracle.py.
 *
 * What is restated (and deliberately kept, quirks included):
 *   makeRecordSlab          record_slab.go:28-122      AoS Ints/Strs/Populated slab
 *   unpackIntCol            column_store_io.go:690-780
 *   unpackStrCol            column_store_io.go:493-609
 *   TableColumn             table_column.go:5-58       per-block string ids
 *   IntFilter / StrFilter   filter.go:171-250
 *   ShouldLoadBlockFromDir  table_block_io.go:110-182
 *   FilterAndAggRecords     aggregate.go:56-282
 *   translate_group_by      aggregate.go:284-324
 *   BasicHist               hist_basic.go:34-279
 *   MultiHist               hist_multi.go:22-257
 *   Result.Combine          query_spec.go:138-193
 *   CombineResults          aggregate.go:414-467
 *   SortResults             aggregate.go:497-525 (ties broken by key, see below)
 * Not restated (documented in DESIGN.md): MultiCombineResults' lossy pruning
 * (aggregate.go:347-412), the query cache, count-distinct, set columns, tdigest.
 *
 * Where Go iterates a map (random order) this file iterates blocks in ascending
 * block order and groups in first-seen order, and says so at the site: the only
 * observable consequence is the last bits of the float running means.
 * In addition to the reference's float Avg each histogram carries ExactSum, the
 * wrapping int64 sum of the accepted values, which the GPU path must match
 * bit-for-bit.
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <regex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../include/sybilgpu.h"

namespace {

// record.go:14-19
enum : int8_t { NO_VAL = 0, INT_VAL = 1, STR_VAL = 2, SET_VAL = 3 };  // record.go:14-19
// aggregate.go:15-16,31 ; hist.go:3
const int INTERNAL_RESULT_LIMIT = 100000;
const int GROUP_BY_WIDTH = 8;
const uint64_t MISSING_VALUE = UINT64_MAX;
const int NUM_BUCKETS = 1000;

struct Flags {  // the globals of config.go the hot path reads
  bool op_hist = false;     // FLAGS.OP == "hist"
  bool log_hist = false;    // FLAGS.LOG_HIST
  int hist_bucket = 0;      // FLAGS.HIST_BUCKET
  bool weight_col = false;  // OPTS.WEIGHT_COL
  int weight_col_id = 0;    // OPTS.WEIGHT_COL_ID
  int time_col_id = -1;     // OPTS.TIME_COL_ID
  int order_by = -1;        // QuerySpec.OrderBy: -1 "$COUNT", -2 "" (no sort), >= 0 index of the aggregation
  bool order_asc = false;   // QuerySpec.OrderAsc
};

struct IntInfo {  // table_column_info.go:18-24 (Min/Max are all the path reads)
  int64_t Min = 0, Max = 0;
};

static inline int64_t wrap_mul10(int64_t v) { return (int64_t)((uint64_t)v * 10ull); }
static inline int64_t wrap_add(int64_t a, int64_t b) { return (int64_t)((uint64_t)a + (uint64_t)b); }
static inline int64_t wrap_sub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }

// ---------------------------------------------------------------------------
// BasicHist — hist_basic.go
// ---------------------------------------------------------------------------
struct BasicHist {
  int64_t NumBuckets = 0;
  int64_t BucketSize = 0;
  std::vector<int64_t> Values;
  std::vector<double> Averages;
  bool PercentileMode = false;
  std::vector<int64_t> Outliers, Underliers;
  int64_t Max = 0, Min = 0;
  int64_t Samples = 0;
  int64_t Count = 0;
  double Avg = 0;
  IntInfo Info;
  int64_t ExactSum = 0;  // not in the reference: exact wrapping sum of accepted values
  const Flags* flags = nullptr;

  // hist_basic.go:34-70
  void SetupBuckets(int64_t buckets, int64_t min, int64_t max) {
    Avg = 0;
    Count = 0;
    Min = min;
    Max = max;
    if (PercentileMode) {
      Outliers.clear();
      Underliers.clear();
      int64_t size = wrap_sub(max, min);
      NumBuckets = buckets;
      BucketSize = size / buckets;
      if (flags->hist_bucket > 0) BucketSize = flags->hist_bucket;
      if (BucketSize == 0) {
        if (size < 100) {
          BucketSize = 1;
          NumBuckets = size;
        } else {
          BucketSize = size / 100;
          NumBuckets = size / BucketSize;
        }
      }
      NumBuckets += 1;
      int64_t n = NumBuckets + 1;
      if (n < 0) n = 0;  // Go would panic on a negative make(); keep the oracle alive
      Values.assign((size_t)n, 0);
      Averages.assign((size_t)n, 0.0);
    }
  }
  // hist_basic.go:87-91
  void TrackPercentiles() {
    PercentileMode = true;
    SetupBuckets(NUM_BUCKETS, Info.Min, Info.Max);
  }
  // hist_basic.go:101-151
  void AddWeightedValue(int64_t value, int64_t weight) {
    if (value > wrap_mul10(Info.Max) || value < Info.Min) return;
    if (flags->weight_col || weight > 1) {
      Samples++;
      Count += weight;
    } else {
      Count++;
    }
    ExactSum = wrap_add(ExactSum, (int64_t)((uint64_t)value * (uint64_t)weight));
    Avg = Avg + (((double)value - Avg) / (double)Count) * (double)weight;
    if (value > Max) Max = value;
    if (value < Min) Min = value;
    if (!PercentileMode) return;
    int64_t bucket_value = wrap_sub(value, Min) / BucketSize;
    if (bucket_value >= (int64_t)Values.size()) {
      Outliers.push_back(value);
      bucket_value = (int64_t)Values.size() - 1;
    }
    if (bucket_value < 0) {
      Underliers.push_back(value);
      bucket_value = 0;
    }
    double partial = Averages[(size_t)bucket_value];
    Values[(size_t)bucket_value] += weight;
    Averages[(size_t)bucket_value] =
        partial + (((double)value - partial) / (double)Values[(size_t)bucket_value] * (double)weight);
  }
  // hist_basic.go:153-183
  std::vector<int64_t> GetPercentiles() const {
    if (Count == 0) return {};
    std::vector<int64_t> percentiles(101, 0);
    percentiles[0] = Min;
    int64_t count = 0, prev_p = 0;
    for (size_t k = 0; k < Values.size(); k++) {
      count += Values[k];
      int64_t p = (100 * count) / Count;
      for (int64_t ip = prev_p; ip <= p; ip++)
        if (ip >= 0 && ip <= 100) percentiles[(size_t)ip] = (int64_t)k * BucketSize + Min;
      if (p >= 0 && p <= 100) percentiles[(size_t)p] = (int64_t)k;
      prev_p = p;
    }
    percentiles.resize(100);
    return percentiles;
  }
  // hist_basic.go:192-219
  double GetStdDev() const {
    double sum_variance = 0;
    for (size_t b = 0; b < Values.size(); b++) {
      int64_t val = (int64_t)b * BucketSize + Min;
      double delta = (double)val - Avg;
      double ratio = (double)Values[b] / (double)Count;
      sum_variance += (delta * delta) * ratio;
    }
    for (int64_t v : Outliers) {
      double delta = std::pow((double)v - Avg, 2);
      sum_variance += delta * (1 / (double)Count);
    }
    for (int64_t v : Underliers) {
      double delta = std::pow((double)v - Avg, 2);
      sum_variance += delta * (1 / (double)Count);
    }
    return std::sqrt(sum_variance);
  }
  // hist_basic.go:221-239 (Go map -> ordered map; the order is not observable)
  std::map<int64_t, int64_t> GetSparseBuckets() const {
    std::map<int64_t, int64_t> ret;
    for (size_t k = 0; k < Values.size(); k++)
      if (Values[k] > 0) ret[(int64_t)k * BucketSize + Min] = Values[k];
    for (int64_t v : Outliers) ret[v] += 1;
    for (int64_t v : Underliers) ret[v] += 1;
    return ret;
  }
  // hist_basic.go:259-279: Averages, Outliers and Underliers are NOT merged
  void Combine(const BasicHist& o) {
    for (size_t k = 0; k < o.Values.size() && k < Values.size(); k++) Values[k] += o.Values[k];
    int64_t total = Count + o.Count;
    Avg = (Avg * ((double)Count / (double)total)) + (o.Avg * ((double)o.Count / (double)total));
    if (Min > o.Min) Min = o.Min;
    if (Max < o.Max) Max = o.Max;
    Samples = Samples + o.Samples;
    Count = total;
    ExactSum = wrap_add(ExactSum, o.ExactSum);
  }
};

// hist_basic.go:72-85
static BasicHist newBasicHist(const Flags* f, const IntInfo& info) {
  BasicHist h;
  h.flags = f;
  h.Info = info;
  if (f->op_hist) h.TrackPercentiles();
  return h;
}

// ---------------------------------------------------------------------------
// MultiHist — hist_multi.go
// ---------------------------------------------------------------------------
struct MultiHist {
  int64_t Max = 0, Min = 0;
  int64_t Samples = 0;
  int64_t Count = 0;
  double Avg = 0;
  bool PercentileMode = false;
  std::vector<BasicHist> Subhists;
  IntInfo Info;
  int64_t ExactSum = 0;
  const Flags* flags = nullptr;

  // hist_multi.go:223-257
  void TrackPercentiles() {
    PercentileMode = true;
    int64_t BucketSize = wrap_sub(Max, Min);
    int num_hists = 0;
    for (int64_t t = BucketSize; t > (int64_t)NUM_BUCKETS; t >>= 1) num_hists += 1;
    Subhists.clear();
    Subhists.resize((size_t)num_hists + 1);
    int64_t right_edge = Max;
    for (int i = 0; i < num_hists; i++) {
      BucketSize >>= 1;
      IntInfo info;
      info.Min = wrap_sub(right_edge, BucketSize);
      info.Max = right_edge;
      right_edge = info.Min;
      Subhists[(size_t)i] = newBasicHist(flags, info);
      Subhists[(size_t)i].TrackPercentiles();
    }
    IntInfo info;
    info.Min = Min;
    info.Max = right_edge;
    Subhists[(size_t)num_hists] = newBasicHist(flags, info);
    Subhists[(size_t)num_hists].TrackPercentiles();
  }
  // hist_multi.go:48-88
  void AddWeightedValue(int64_t value, int64_t weight) {
    if (value > wrap_mul10(Info.Max) || value < Info.Min) return;
    if (flags->weight_col || weight > 1) {
      Samples++;
      Count += weight;
    } else {
      Count++;
    }
    ExactSum = wrap_add(ExactSum, (int64_t)((uint64_t)value * (uint64_t)weight));
    Avg = Avg + (((double)value - Avg) / (double)Count) * (double)weight;
    if (value > Max) Max = value;
    if (value < Min) Min = value;
    if (!PercentileMode) return;
    for (auto& sh : Subhists) {
      if (value >= sh.Info.Min && value <= sh.Info.Max) {
        sh.AddWeightedValue(value, weight);
        break;
      }
    }
  }
  // hist_multi.go:184-200
  std::map<int64_t, int64_t> GetSparseBuckets() const {
    std::map<int64_t, int64_t> all;
    for (auto& sh : Subhists)
      for (auto& kv : sh.GetSparseBuckets()) all[kv.first] += kv.second;
    return all;
  }
  // hist_multi.go:90-131
  std::vector<int64_t> GetPercentiles() const {
    if (Count == 0) return {};
    auto all = GetSparseBuckets();
    std::vector<int64_t> buckets;
    int64_t total = 0;
    for (auto& kv : all)
      if (kv.second > 0) {
        buckets.push_back(kv.first);
        total += kv.second;
      }
    // sort.Ints on int(bucket): already ascending in the ordered map
    int64_t prev_p = 0, count = 0;
    std::vector<int64_t> percentiles(101, 0);
    if (total == 0) {
      percentiles.resize(100);
      return percentiles;
    }
    for (int64_t k : buckets) {
      count += all[k];
      int64_t p = (100 * count) / total;
      for (int64_t ip = prev_p; ip <= p; ip++)
        if (ip <= 100 && ip >= 0) percentiles[(size_t)ip] = k;
      if (p <= 100 && p >= 0) percentiles[(size_t)p] = k;
      prev_p = p;
    }
    percentiles.resize(100);
    return percentiles;
  }
  // hist_multi.go:144-158 (map order -> ascending bucket order)
  double GetStdDev() const {
    auto all = GetSparseBuckets();
    double sum_variance = 0;
    for (auto& kv : all) {
      double delta = (double)kv.first - Avg;
      double ratio = (double)kv.second / (double)Count;
      sum_variance += (delta * delta) * ratio;
    }
    return std::sqrt(sum_variance);
  }
  // hist_multi.go:202-221
  void Combine(const MultiHist& o) {
    for (size_t i = 0; i < Subhists.size() && i < o.Subhists.size(); i++) Subhists[i].Combine(o.Subhists[i]);
    int64_t total = Count + o.Count;
    Avg = (Avg * ((double)Count / (double)total)) + (o.Avg * ((double)o.Count / (double)total));
    if (Min > o.Min) Min = o.Min;
    if (Max < o.Max) Max = o.Max;
    Samples = Samples + o.Samples;
    Count = total;
    ExactSum = wrap_add(ExactSum, o.ExactSum);
  }
};

// hist_multi.go:22-38
static MultiHist newMultiHist(const Flags* f, const IntInfo& info) {
  MultiHist h;
  h.flags = f;
  h.Info = info;
  h.Avg = 0;
  h.Count = 0;
  h.Min = info.Min;
  h.Max = info.Max;
  if (f->op_hist) h.TrackPercentiles();
  return h;
}

// Histogram interface (hist.go:9-25) as a tagged pair
struct Hist {
  bool multi = false;
  BasicHist b;
  MultiHist m;
  // Table.NewHist, hist.go:27-38 (T_DIGEST needs a build tag: out of scope)
  static Hist New(const Flags* f, const IntInfo& info) {
    Hist h;
    h.multi = f->log_hist;
    if (h.multi)
      h.m = newMultiHist(f, info);
    else
      h.b = newBasicHist(f, info);
    return h;
  }
  Hist NewHist(const Flags* f) const { return New(f, multi ? m.Info : b.Info); }  // hist_compat.go:18-20
  void AddWeightedValue(int64_t v, int64_t w) { multi ? m.AddWeightedValue(v, w) : b.AddWeightedValue(v, w); }
  void Combine(const Hist& o) { multi ? m.Combine(o.m) : b.Combine(o.b); }
  double Mean() const { return multi ? m.Avg : b.Avg; }
  int64_t TotalCount() const { return multi ? m.Count : b.Count; }
  int64_t MinV() const { return multi ? m.Min : b.Min; }
  int64_t MaxV() const { return multi ? m.Max : b.Max; }
  int64_t ExactSum() const { return multi ? m.ExactSum : b.ExactSum; }
  int64_t Samples() const { return multi ? m.Samples : b.Samples; }
  std::vector<int64_t> GetPercentiles() const { return multi ? m.GetPercentiles() : b.GetPercentiles(); }
  double StdDev() const { return multi ? m.GetStdDev() : b.GetStdDev(); }
  std::map<int64_t, int64_t> GetSparseBuckets() const { return multi ? m.GetSparseBuckets() : b.GetSparseBuckets(); }
};

// ---------------------------------------------------------------------------
// Result / ResultMap — query_spec.go:85-193
// ---------------------------------------------------------------------------
struct Result {
  std::map<int, Hist> Hists;  // keyed by aggregation index (Go: by name)
  std::string GroupByKey;
  std::string BinaryByKey;
  int64_t Count = 0;
  int64_t Samples = 0;
};
typedef std::shared_ptr<Result> ResultP;
// insertion-ordered map so "map iteration" has a stated, reproducible order
struct ResultMap {
  std::unordered_map<std::string, ResultP> idx;
  std::vector<std::string> order;
  ResultP find(const std::string& k) const {
    auto it = idx.find(k);
    return it == idx.end() ? nullptr : it->second;
  }
  void put(const std::string& k, ResultP r) {
    if (idx.find(k) == idx.end()) order.push_back(k);
    idx[k] = r;
  }
  size_t size() const { return idx.size(); }
};

// query_spec.go:138-193 (MERGE_TABLE branch is the cross-node path: out of scope)
static void ResultCombine(const Flags* f, Result& rs, const Result& next) {
  if (next.Count == 0) return;
  int64_t total_samples = rs.Samples + next.Samples;
  int64_t total_count = rs.Count + next.Count;
  for (auto& kv : next.Hists) {
    auto it = rs.Hists.find(kv.first);
    if (it == rs.Hists.end()) {
      Hist nh = kv.second.NewHist(f);
      nh.Combine(kv.second);
      rs.Hists[kv.first] = nh;
    } else {
      it->second.Combine(kv.second);
    }
  }
  rs.Samples = total_samples;
  rs.Count = total_count;
}
// query_spec.go:107-116: the first result seen for a key is adopted by reference
static void ResultMapCombine(const Flags* f, ResultMap& master, const ResultMap& results) {
  for (auto& k : results.order) {
    ResultP v = results.idx.at(k);
    ResultP mval = master.find(k);
    if (!mval)
      master.put(k, v);
    else
      ResultCombine(f, *mval, *v);
  }
}

// ---------------------------------------------------------------------------
// blocks, columns, records
// ---------------------------------------------------------------------------
struct SavedColumn {  // SavedIntColumn / SavedStrColumn, column_store.go:46-64
  int col_slot = 0, col_type = 0, encoding = 0;
  bool delta_ids = false, delta_values = false;
  std::vector<int64_t> bin_values;
  std::vector<uint32_t> bin_offsets, record_ids;
  std::vector<int64_t> values_i64;
  std::vector<int32_t> values_i32;
  std::vector<std::string> StringTable;
  uint32_t set_nvalues = 0;  // SavedSetColumn in its non-bucketed form: len(Values) (sybilgpu.h hands it over as bins + this)
};
struct SavedBlock {
  int64_t block_index = 0;
  int32_t NumRecords = 0;
  std::vector<SavedColumn> cols;
  std::vector<sg_int_info> info;  // block info.db IntInfoMap
};

struct TableColumn {  // table_column.go:5-58
  int8_t Type = 0;
  std::unordered_map<std::string, int32_t> StringTable;
  std::vector<std::string> val_string_id_lookup;
  // table_column.go:27-48
  int32_t get_val_id(const std::string& name) {
    auto it = StringTable.find(name);
    if (it != StringTable.end()) return it->second;
    int32_t id = (int32_t)StringTable.size();
    StringTable[name] = id;
    if (StringTable.size() > val_string_id_lookup.size()) val_string_id_lookup.resize(StringTable.size() << 1);
    val_string_id_lookup[(size_t)id] = name;
    return id;
  }
  // table_column.go:50-58
  std::string get_string_for_val(int32_t id) const {
    if (id < 0 || (size_t)id >= val_string_id_lookup.size()) return "";
    return val_string_id_lookup[(size_t)id];
  }
};

struct Block {  // TableBlock with its AoS slab (record_slab.go:28-122)
  int32_t n = 0;
  int ncols = 0;
  std::vector<int64_t> Ints;      // n * ncols
  std::vector<int32_t> Strs;      // n * ncols
  std::vector<int8_t> Populated;  // n * ncols
  std::vector<TableColumn> columns;
  std::map<int, std::vector<std::vector<int32_t>>> SetMap;  // Record.SetMap (record.go): col -> per row the tag ids
};

struct StrReplace {  // config.go:102-105
  std::regex re;
  std::string Replace;
};
struct Table {
  int ncols = 0;
  std::vector<int32_t> KeyTypes;
  std::vector<SavedBlock> blocks;
  std::map<int, StrReplace> str_replacements;  // OPTS.STR_REPLACEMENTS by column (table_query.go:34-50)
};

// column_store_io.go:690-780.  Returns false on "BLOCK SIZE CHANGED DURING QUERY".
static bool unpackIntCol(Block& tb, const SavedColumn& into) {
  const int col_id = into.col_slot;
  const uint32_t num_records = (uint32_t)tb.n;
  const int K = tb.ncols;
  if (into.encoding == SG_ENC_BUCKET) {
    for (size_t b = 0; b + 1 < into.bin_offsets.size(); b++) {
      uint32_t prev = 0;
      for (uint32_t j = into.bin_offsets[b]; j < into.bin_offsets[b + 1]; j++) {
        uint32_t r = into.record_ids[j];
        if (into.delta_ids) r = r + prev;
        if (r >= num_records) return false;
        tb.Ints[(size_t)r * K + col_id] = into.bin_values[b];
        tb.Populated[(size_t)r * K + col_id] = INT_VAL;
        prev = r;
      }
    }
  } else if (into.encoding == SG_ENC_VALUES) {
    int64_t prev = 0;
    if ((uint32_t)into.values_i64.size() > num_records) return false;
    for (size_t r = 0; r < into.values_i64.size(); r++) {
      int64_t v = into.values_i64[r];
      if (into.delta_values) v = wrap_add(v, prev);
      tb.Ints[r * K + col_id] = v;
      tb.Populated[r * K + col_id] = INT_VAL;
      if (into.delta_values) prev = v;
    }
  }
  return true;
}

// column_store_io.go:493-609.  str_replace: OPTS.STR_REPLACEMENTS[into.Name] (nullptr: none).  The regexp dialect is
// std::regex ECMAScript, not Go's RE2, and the replacement template is std::regex_replace's ($1, $& ...), not Go's
// Expand ($1, ${1}, $name): the same for plain patterns and $N templates — dialect unpinned.
static bool unpackStrCol(Block& tb, const SavedColumn& into, const StrReplace* str_replace) {
  const int col_id = into.col_slot;
  const uint32_t num_records = (uint32_t)tb.n;
  const int K = tb.ncols;
  TableColumn& col = tb.columns[(size_t)col_id];
  std::vector<std::string> string_lookup((size_t)tb.n);
  std::unordered_map<int32_t, int32_t> bucket_replace;
  if ((uint32_t)into.StringTable.size() > num_records) return false;
  for (size_t k = 0; k < into.StringTable.size(); k++) {
    std::string v = into.StringTable[k];
    if (str_replace) v = std::regex_replace(v, str_replace->re, str_replace->Replace);  // re.ReplaceAllString (:531)
    auto ex = col.StringTable.find(v);
    if (ex != col.StringTable.end()) {
      bucket_replace[(int32_t)k] = ex->second;
    } else {
      bucket_replace[(int32_t)k] = (int32_t)k;
      col.StringTable[v] = (int32_t)k;
    }
    string_lookup[k] = v;
  }
  col.val_string_id_lookup = string_lookup;
  if (into.encoding == SG_ENC_BUCKET) {
    for (size_t b = 0; b + 1 < into.bin_offsets.size(); b++) {
      uint32_t prev = 0;
      int32_t value = (int32_t)into.bin_values[b];
      auto it = bucket_replace.find(value);
      // Go: new_value, should_replace := bucket_replace[value]; cast_value := StrField(new_value)
      // -> an id outside the string table decodes to 0 (map zero value)
      int32_t new_value = it == bucket_replace.end() ? 0 : it->second;
      int32_t cast_value = new_value;
      for (uint32_t j = into.bin_offsets[b]; j < into.bin_offsets[b + 1]; j++) {
        uint32_t r = into.record_ids[j];
        if (into.delta_ids) r = prev + r;
        if (r >= num_records) return false;
        prev = r;
        tb.Populated[(size_t)r * K + col_id] = STR_VAL;
        tb.Strs[(size_t)r * K + col_id] = cast_value;
      }
    }
  } else if (into.encoding == SG_ENC_VALUES) {
    if ((uint32_t)into.values_i32.size() > num_records) return false;
    for (size_t r = 0; r < into.values_i32.size(); r++) {
      int32_t v = into.values_i32[r];
      auto it = bucket_replace.find(v);
      if (it != bucket_replace.end()) v = it->second;
      tb.Strs[r * K + col_id] = v;
      tb.Populated[r * K + col_id] = STR_VAL;
    }
  }
  return true;
}

// column_store_io.go:611-688.  (A later duplicate in the string table overwrites the earlier id in col.StringTable,
// :632-635 — unlike unpackStrCol, where the first wins.)
static bool unpackSetCol(Block& tb, const SavedColumn& into) {
  const int col_id = into.col_slot;
  const uint32_t num_records = (uint32_t)tb.n;
  const int K = tb.ncols;
  TableColumn& col = tb.columns[(size_t)col_id];
  std::vector<std::string> tr_string_lookup(into.StringTable.size());
  for (size_t k = 0; k < into.StringTable.size(); k++) {
    col.StringTable[into.StringTable[k]] = (int32_t)k;
    tr_string_lookup[k] = into.StringTable[k];
  }
  col.val_string_id_lookup = tr_string_lookup;
  auto& sets = tb.SetMap[col_id];
  sets.resize((size_t)tb.n);
  // BucketEncoded
  for (size_t b = 0; b + 1 < into.bin_offsets.size(); b++) {
    uint32_t prev = 0;
    for (uint32_t j = into.bin_offsets[b]; j < into.bin_offsets[b + 1]; j++) {
      uint32_t r = into.record_ids[j];
      if (into.delta_ids) r = r + prev;
      if (r >= num_records) return false;
      sets[r].push_back((int32_t)into.bin_values[b]);
      tb.Populated[(size_t)r * K + col_id] = SET_VAL;
      prev = r;
    }
  }
  // the non-bucketed form (Values [][]int32, :670-683): every listed row is populated, empty set or not.  The tags
  // themselves arrive as bins (sybilgpu.h); the reference's "len(Values) > num_records" error is kept.
  if (into.set_nvalues > num_records) return false;
  for (uint32_t r = 0; r < into.set_nvalues; r++) tb.Populated[(size_t)r * K + col_id] = SET_VAL;
  return true;
}

// LoadBlockFromDir, table_block_io.go:225-310: only files named in the LoadSpec are unpacked
static bool LoadBlock(const Table& t, const SavedBlock& sb, const std::vector<char>& wanted, Block& tb) {
  if (sb.NumRecords <= 0) return false;
  tb.n = sb.NumRecords;
  tb.ncols = t.ncols;
  size_t cells = (size_t)tb.n * (size_t)t.ncols;
  tb.Ints.assign(cells, 0);
  tb.Strs.assign(cells, 0);
  tb.Populated.assign(cells, 0);
  tb.columns.assign((size_t)t.ncols, TableColumn());
  for (auto& c : sb.cols) {
    if (c.col_slot < 0 || c.col_slot >= t.ncols) continue;
    if (!wanted[(size_t)c.col_slot]) continue;
    bool ok = true;
    if (c.col_type == SG_COL_STR) {
      auto sr = t.str_replacements.find(c.col_slot);
      ok = unpackStrCol(tb, c, sr == t.str_replacements.end() ? nullptr : &sr->second);
    } else if (c.col_type == SG_COL_INT)
      ok = unpackIntCol(tb, c);
    else if (c.col_type == SG_COL_SET)
      ok = unpackSetCol(tb, c);
    if (!ok) return false;  // "ERROR DURING COLUMN UNPACK ... SKIPPING BLOCK"
  }
  return true;
}

// ---------------------------------------------------------------------------
// filters — filter.go
// ---------------------------------------------------------------------------
struct Filter {
  int col = 0, type = 0, op = 0;
  int64_t ivalue = 0;
  std::string svalue;
  std::regex re;
  bool has_re = false;
};
// filter.go:171-195
static bool IntFilterFilter(const Filter& f, const Block& tb, size_t r) {
  size_t K = (size_t)tb.ncols;
  if (tb.Populated[r * K + f.col] == 0) return false;
  int64_t field = tb.Ints[r * K + f.col];
  switch (f.op) {
    case SG_OP_GT: return field > f.ivalue;
    case SG_OP_LT: return field < f.ivalue;
    case SG_OP_EQ: return field == f.ivalue;
    case SG_OP_NEQ: return field != f.ivalue;
    default: return false;
  }
}
// filter.go:199-250
static bool StrFilterFilter(const Filter& f, Block& tb, size_t r) {
  size_t K = (size_t)tb.ncols;
  if (tb.Populated[r * K + f.col] == 0) return false;
  int32_t val = tb.Strs[r * K + f.col];
  TableColumn& col = tb.columns[(size_t)f.col];
  int64_t filterval = col.get_val_id(f.svalue);  // inserts the literal when absent (Q3)
  bool ret = false;
  switch (f.op) {
    case SG_OP_NRE:
    case SG_OP_RE: {
      std::string s = col.get_string_for_val(val);
      ret = std::regex_search(s, f.re);  // Go regexp.MatchString is an unanchored search
      if (f.op == SG_OP_NRE) ret = !ret;
      break;
    }
    case SG_OP_EQ: ret = (int64_t)val == filterval; break;
    case SG_OP_NEQ: ret = (int64_t)val != filterval; break;
    default: break;
  }
  return ret;
}

// filter.go:252-285
static bool SetFilterFilter(const Filter& f, Block& tb, size_t r) {
  size_t K = (size_t)tb.ncols;
  TableColumn& col = tb.columns[(size_t)f.col];
  bool ret = false;
  if (tb.Populated[r * K + f.col] != SET_VAL) return false;
  const std::vector<int32_t>& sets = tb.SetMap[f.col][r];
  int32_t val_id = col.get_val_id(f.svalue);
  switch (f.op) {
    case SG_OP_IN:
      for (int32_t tag : sets)
        if (tag == val_id) return true;
      break;
    case SG_OP_NIN:
      ret = true;
      for (int32_t tag : sets)
        if (tag == val_id) return false;
      break;
    default: break;
  }
  return ret;
}

struct QuerySpec {
  const Flags* fp = nullptr;  // the query's globals (owned by the orc_result)
  std::vector<Filter> Filters;
  std::vector<sg_group_desc> Groups;
  std::vector<sg_agg_desc> Aggregations;
  int64_t TimeBucket = 0;
  // results
  ResultMap Results;
  std::map<int64_t, ResultMap> TimeResults;
  int64_t MatchedCount = 0;
};

// table_block_io.go:110-182
static bool ShouldLoadBlock(const QuerySpec& qs, const std::vector<sg_int_info>& info) {
  if (info.empty()) return true;
  bool add = true;
  for (auto& f : qs.Filters) {
    if (f.type != SG_COL_INT) continue;
    const sg_int_info* fi = nullptr;
    for (auto& i : info)
      if (i.col_slot == f.col) fi = &i;
    if (f.op == SG_OP_GT || f.op == SG_OP_LT) {
      // min_record / max_record carry only the columns of IntInfoMap; an absent
      // column is unpopulated there, so Filter() is false on both
      bool pmin = false, pmax = false;
      if (fi) {
        pmin = f.op == SG_OP_GT ? fi->min > f.ivalue : fi->min < f.ivalue;
        pmax = f.op == SG_OP_GT ? fi->max > f.ivalue : fi->max < f.ivalue;
      }
      if (!pmin && !pmax) add = false;
    }
    if (f.op == SG_OP_EQ) {
      if (!fi) {
        add = false;
      } else if (fi->min > f.ivalue || fi->max < f.ivalue) {
        add = false;
      }
    }
  }
  return add;
}

static void put_u64le(std::string& buf, size_t off, uint64_t v) {
  for (int i = 0; i < 8; i++) buf[off + i] = (char)((v >> (8 * i)) & 0xff);
}
static uint64_t get_u64le(const std::string& buf, size_t off) {
  uint64_t v = 0;
  for (int i = 0; i < 8; i++) v |= (uint64_t)(uint8_t)buf[off + i] << (8 * i);
  return v;
}

// aggregate.go:284-324
static ResultMap translate_group_by(const ResultMap& Results, const std::vector<sg_group_desc>& Groups,
                                    const std::vector<TableColumn*>& columns) {
  ResultMap out;
  for (auto& k : Results.order) {
    ResultP r = Results.idx.at(k);
    std::string buffer;
    if (Groups.empty()) buffer += "total";
    for (size_t i = 0; i < Groups.size(); i++) {
      TableColumn* col = columns[(size_t)Groups[i].col_slot];
      if (col == nullptr) {
        buffer += "\t";
        continue;
      }
      uint64_t val = get_u64le(r->BinaryByKey, i * GROUP_BY_WIDTH);
      if (val != MISSING_VALUE) {
        if (col->Type == INT_VAL)
          buffer += std::to_string((int64_t)val);
        else if (col->Type == STR_VAL)
          buffer += col->get_string_for_val((int32_t)val);
      }
      buffer += "\t";
    }
    r->GroupByKey = buffer;
    out.put(buffer, r);
  }
  return out;
}

// aggregate.go:56-282
static int64_t FilterAndAggRecords(QuerySpec& qs, Block& tb) {
  const Flags& F = *qs.fp;
  std::string binarybuffer((size_t)GROUP_BY_WIDTH * qs.Groups.size(), '\0');
  int64_t weight = 1;  // declared outside the row loop: carries over (Q13)
  int64_t matched_records = 0;
  std::vector<TableColumn*> columns((size_t)tb.ncols, nullptr);
  ResultMap* result_map = &qs.Results;
  const size_t K = (size_t)tb.ncols;

  for (size_t i = 0; i < (size_t)tb.n; i++) {
    bool add = true;
    if (F.weight_col && tb.Populated[i * K + F.weight_col_id] == INT_VAL) weight = tb.Ints[i * K + F.weight_col_id];

    for (auto& f : qs.Filters) {
      bool ok = f.type == SG_COL_INT ? IntFilterFilter(f, tb, i)
                                     : (f.type == SG_COL_SET ? SetFilterFilter(f, tb, i) : StrFilterFilter(f, tb, i));
      if (!ok) {
        add = false;
        break;
      }
    }
    if (!add) continue;
    matched_records++;

    for (size_t g = 0; g < qs.Groups.size(); g++) {
      int id = qs.Groups[g].col_slot;
      int8_t pop = tb.Populated[i * K + id];
      if (columns[(size_t)id] == nullptr && pop != NO_VAL) {
        columns[(size_t)id] = &tb.columns[(size_t)id];
        columns[(size_t)id]->Type = pop;
      }
      uint64_t v = 0;
      switch (pop) {
        case INT_VAL: v = (uint64_t)tb.Ints[i * K + id]; break;
        case STR_VAL: v = (uint64_t)(int64_t)tb.Strs[i * K + id]; break;
        case NO_VAL: v = MISSING_VALUE; break;
      }
      put_u64le(binarybuffer, g * GROUP_BY_WIDTH, v);
    }

    if (qs.TimeBucket > 0) {
      if ((int)K <= F.time_col_id || F.time_col_id < 0) continue;
      if (tb.Populated[i * K + F.time_col_id] != INT_VAL) continue;
      int64_t val = tb.Ints[i * K + F.time_col_id];
      ResultP big_record = qs.Results.find(binarybuffer);
      if (!big_record) {
        if ((int)qs.Results.size() < INTERNAL_RESULT_LIMIT) {
          big_record = std::make_shared<Result>();
          big_record->BinaryByKey = binarybuffer;
          qs.Results.put(binarybuffer, big_record);
        }
      }
      if (big_record) {
        big_record->Samples++;
        big_record->Count += weight;
      }
      val = val / qs.TimeBucket * qs.TimeBucket;  // truncation toward zero (Q14)
      result_map = &qs.TimeResults[val];
    }

    ResultP added_record = result_map->find(binarybuffer);
    if (!added_record) {
      if ((int)result_map->size() >= INTERNAL_RESULT_LIMIT) continue;
      added_record = std::make_shared<Result>();
      added_record->BinaryByKey = binarybuffer;
      result_map->put(binarybuffer, added_record);
    }
    added_record->Samples++;
    added_record->Count += weight;

    for (size_t a = 0; a < qs.Aggregations.size(); a++) {
      int id = qs.Aggregations[a].col_slot;
      if (tb.Populated[i * K + id] == INT_VAL) {
        int64_t val = tb.Ints[i * K + id];
        auto it = added_record->Hists.find((int)a);
        if (it == added_record->Hists.end()) {
          IntInfo info;
          info.Min = qs.Aggregations[a].info_min;
          info.Max = qs.Aggregations[a].info_max;
          it = added_record->Hists.emplace((int)a, Hist::New(qs.fp, info)).first;
        }
        it->second.AddWeightedValue(val, weight);
      }
    }
  }

  for (auto& kv : qs.TimeResults) kv.second = translate_group_by(kv.second, qs.Groups, columns);
  if (qs.Results.size() > 0) qs.Results = translate_group_by(qs.Results, qs.Groups, columns);
  return matched_records;
}

struct Combined {  // resultSpec of CombineResults
  ResultP Cumulative;
  ResultMap Results;
  std::map<int64_t, ResultMap> TimeResults;
  int64_t MatchedCount = 0;
  std::vector<ResultP> Sorted;
  std::map<int64_t, std::vector<ResultP>> TimeSorted;
};

static void sort_results(const Flags* f, std::vector<ResultP>& v) {
  // QuerySpec.SortResults (aggregate.go:497-525): no sort for OrderBy == ""; SortResultsByCol.Less (:43-54)
  // orders descending by Count ("$COUNT") or by Hists[col].Mean(); sort.Sort is unstable in Go, ties are
  // broken here by GroupByKey ascending; orderAsc then reverses the whole list (:516-520).  A group without
  // the histogram (Go: nil map entry, Mean() would panic) sorts as mean = -inf.
  if (f->order_by == -2) return;
  const int col = f->order_by;
  auto mean = [col](const ResultP& r) {
    auto it = r->Hists.find(col);
    return it == r->Hists.end() || it->second.TotalCount() == 0 ? -INFINITY : it->second.Mean();
  };
  std::sort(v.begin(), v.end(), [&](const ResultP& a, const ResultP& b) {
    if (col < 0) {
      if (a->Count != b->Count) return a->Count > b->Count;
    } else {
      const double ma = mean(a), mb = mean(b);
      if (ma != mb) return ma > mb;
    }
    return a->GroupByKey < b->GroupByKey;
  });
  if (f->order_asc) std::reverse(v.begin(), v.end());
}

// aggregate.go:414-467; block_specs iterated in ascending block order
static void CombineResults(const QuerySpec& proto, std::vector<std::unique_ptr<QuerySpec>>& block_specs, Combined& out) {
  const Flags* f = proto.fp;
  out.Cumulative = std::make_shared<Result>();
  out.Cumulative->GroupByKey = "TOTAL";
  for (size_t i = 1; i < proto.Groups.size(); i++) out.Cumulative->GroupByKey += "\t";
  for (auto& specp : block_specs) {
    if (!specp) continue;
    QuerySpec& spec = *specp;
    ResultMapCombine(f, out.Results, spec.Results);
    out.MatchedCount += spec.MatchedCount;
    for (auto& k : spec.Results.order) ResultCombine(f, *out.Cumulative, *spec.Results.idx.at(k));
    for (auto& tv : spec.TimeResults) {
      auto mit = out.TimeResults.find(tv.first);
      if (mit == out.TimeResults.end()) {
        out.TimeResults[tv.first] = tv.second;
      } else {
        for (auto& k : tv.second.order) {
          ResultP r = tv.second.idx.at(k);
          ResultP mh = mit->second.find(k);
          if (mh)
            ResultCombine(f, *mh, *r);
          else
            mit->second.put(k, r);
        }
      }
    }
  }
  for (auto& k : out.Results.order) out.Sorted.push_back(out.Results.idx.at(k));
  sort_results(f, out.Sorted);
  for (auto& tv : out.TimeResults) {
    auto& v = out.TimeSorted[tv.first];
    for (auto& k : tv.second.order) v.push_back(tv.second.idx.at(k));
    sort_results(f, v);
  }
}

}  // namespace

// ===========================================================================
// C interface used by tests/ and bench.py's cpu_baseline through ctypes
// ===========================================================================
struct orc_table {
  Table t;
};
struct orc_result {
  Combined c;
  Flags flags;
  int naggs = 0;
  int64_t broken = 0, skipped = 0;
  double seconds = 0;
  std::vector<int64_t> time_keys;
  bool merged_view = false;
  // scratch kept alive for pointer-returning accessors
  std::vector<int64_t> values_scratch;
};

extern "C" {

orc_table* orc_table_create(int32_t num_col_slots, const int32_t* col_types) {
  orc_table* t = new orc_table();
  t->t.ncols = num_col_slots;
  t->t.KeyTypes.assign(col_types, col_types + num_col_slots);
  return t;
}
void orc_table_free(orc_table* t) { delete t; }
// FLAGS.STR_REPLACE for one column (table_query.go:34-50); pattern == NULL removes it
int orc_table_set_str_replace(orc_table* t, int32_t col_slot, const char* pattern, const char* replacement) {
  if (!pattern) {
    t->t.str_replacements.erase(col_slot);
    return 0;
  }
  try {
    StrReplace sr;
    sr.re = std::regex(pattern, std::regex::ECMAScript);
    sr.Replace = replacement ? replacement : "";
    t->t.str_replacements[col_slot] = sr;
  } catch (const std::regex_error&) {
    return -1;
  }
  return 0;
}

int orc_table_add_block(orc_table* t, const sg_block_desc* d) {
  SavedBlock sb;
  sb.block_index = d->block_index;
  sb.NumRecords = d->num_records;
  for (int i = 0; i < d->ninfo; i++) sb.info.push_back(d->info[i]);
  for (int i = 0; i < d->ncols; i++) {
    const sg_column_desc& c = d->cols[i];
    SavedColumn sc;
    sc.col_slot = c.col_slot;
    sc.col_type = c.col_type;
    sc.encoding = c.encoding;
    sc.delta_ids = c.delta_ids != 0;
    sc.delta_values = c.delta_values != 0;
    if (c.encoding == SG_ENC_BUCKET) {
      sc.bin_values.assign(c.bin_values, c.bin_values + c.nbins);
      sc.bin_offsets.assign(c.bin_offsets, c.bin_offsets + c.nbins + 1);
      // narrow arrays (sybilgpu.h): the oracle widens them to the post-gob form the reference decodes
      if (c.id_bits == 16) {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(c.record_ids);
        sc.record_ids.assign(p, p + c.nrecord_ids);
      } else {
        sc.record_ids.assign(c.record_ids, c.record_ids + c.nrecord_ids);
      }
    } else if (c.encoding == SG_ENC_VALUES) {
      if (c.col_type == SG_COL_INT) {
        if (c.value_bits == 32 || c.value_bits == 16) {
          // deltas relative to value_base -> Go's delta form: Values[0] absolute, then gaps
          sc.values_i64.resize(c.nvalues);
          for (uint32_t k = 0; k < c.nvalues; k++) {
            const int64_t dlt = c.value_bits == 32 ? (int64_t) reinterpret_cast<const int32_t*>(c.values_i64)[k]
                                                   : (int64_t) reinterpret_cast<const int16_t*>(c.values_i64)[k];
            sc.values_i64[k] = k == 0 ? (int64_t)((uint64_t)c.value_base + (uint64_t)dlt) : dlt;
          }
        } else {
          sc.values_i64.assign(c.values_i64, c.values_i64 + c.nvalues);
        }
      } else if (c.value_bits == 16) {
        const uint16_t* p = reinterpret_cast<const uint16_t*>(c.values_i32);
        sc.values_i32.assign(p, p + c.nvalues);
      } else {
        sc.values_i32.assign(c.values_i32, c.values_i32 + c.nvalues);
      }
    }
    if (c.col_type == SG_COL_SET) sc.set_nvalues = c.nvalues;
    if (c.col_type == SG_COL_STR || c.col_type == SG_COL_SET)
      for (uint32_t k = 0; k < c.ndict; k++)
        sc.StringTable.emplace_back(c.dict_bytes + c.dict_offsets[k], c.dict_offsets[k + 1] - c.dict_offsets[k]);
    sb.cols.push_back(std::move(sc));
  }
  t->t.blocks.push_back(std::move(sb));
  return 0;
}

// LoadAndQueryRecords (table_query.go:18-422): a task per block on nthreads
// workers (the reference's goroutine per block), then one CombineResults.
// max_blocks > 0 bounds the run to the first max_blocks blocks (cpu_baseline sample).
orc_result* orc_query(orc_table* tab, const sg_query_desc* d, int nthreads, int64_t max_blocks) {
  auto t0 = std::chrono::steady_clock::now();
  orc_result* r = new orc_result();
  r->flags.op_hist = d->op_mode == SG_MODE_HIST;
  r->flags.log_hist = d->hist_kind == SG_HIST_MULTI;
  r->flags.hist_bucket = d->hist_bucket;
  r->flags.weight_col = d->weight_col_slot >= 0;
  r->flags.weight_col_id = d->weight_col_slot >= 0 ? d->weight_col_slot : 0;
  r->flags.time_col_id = d->time_col_slot;
  r->flags.order_by = d->order_by_agg;
  r->flags.order_asc = d->order_asc != 0;
  QuerySpec proto;
  proto.fp = &r->flags;
  proto.TimeBucket = d->time_col_slot >= 0 ? d->time_bucket : 0;
  std::vector<char> wanted((size_t)tab->t.ncols, 0);  // LoadSpec.files
  for (int i = 0; i < d->nfilters; i++) {
    Filter f;
    f.col = d->filters[i].col_slot;
    f.type = d->filters[i].col_type;
    f.op = d->filters[i].op;
    f.ivalue = d->filters[i].int_value;
    if (d->filters[i].str_value) f.svalue.assign(d->filters[i].str_value, (size_t)d->filters[i].str_len);
    if (f.type == SG_COL_STR && (f.op == SG_OP_RE || f.op == SG_OP_NRE)) {
      f.re = std::regex(f.svalue, std::regex::ECMAScript);
      f.has_re = true;
    }
    proto.Filters.push_back(f);
    wanted[(size_t)f.col] = 1;
  }
  for (int i = 0; i < d->ngroups; i++) {
    proto.Groups.push_back(d->groups[i]);
    wanted[(size_t)d->groups[i].col_slot] = 1;
  }
  for (int i = 0; i < d->naggs; i++) {
    proto.Aggregations.push_back(d->aggs[i]);
    wanted[(size_t)d->aggs[i].col_slot] = 1;
  }
  if (d->time_col_slot >= 0) wanted[(size_t)d->time_col_slot] = 1;
  if (d->weight_col_slot >= 0) wanted[(size_t)d->weight_col_slot] = 1;

  size_t nblocks = tab->t.blocks.size();
  if (max_blocks > 0 && (size_t)max_blocks < nblocks) nblocks = (size_t)max_blocks;
  std::vector<std::unique_ptr<QuerySpec>> block_specs(nblocks);
  std::atomic<size_t> next(0);
  std::atomic<int64_t> broken(0), skipped(0);
  if (nthreads < 1) nthreads = 1;
  auto worker = [&]() {
    for (;;) {
      size_t i = next.fetch_add(1);
      if (i >= nblocks) break;
      const SavedBlock& sb = tab->t.blocks[i];
      if (!ShouldLoadBlock(proto, sb.info)) {
        skipped++;
        continue;
      }
      Block tb;
      if (!LoadBlock(tab->t, sb, wanted, tb)) {
        broken++;
        continue;
      }
      std::unique_ptr<QuerySpec> bq(new QuerySpec());  // CopyQuerySpec, aggregate.go:326-332
      bq->fp = proto.fp;
      bq->Filters = proto.Filters;
      bq->Groups = proto.Groups;
      bq->Aggregations = proto.Aggregations;
      bq->TimeBucket = proto.TimeBucket;
      bq->MatchedCount = FilterAndAggRecords(*bq, tb);
      block_specs[i] = std::move(bq);
    }
  };
  std::vector<std::thread> th;
  for (int i = 1; i < nthreads; i++) th.emplace_back(worker);
  worker();
  for (auto& x : th) x.join();

  r->naggs = d->naggs;
  CombineResults(proto, block_specs, r->c);
  for (auto& tv : r->c.TimeResults) r->time_keys.push_back(tv.first);
  r->broken = broken.load();
  r->skipped = skipped.load();
  r->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  return r;
}
void orc_result_free(orc_result* r) { delete r; }
double orc_result_seconds(orc_result* r) { return r->seconds; }
int64_t orc_result_matched_count(orc_result* r) { return r->c.MatchedCount; }
int64_t orc_result_num_groups(orc_result* r) { return (int64_t)r->c.Sorted.size(); }
int64_t orc_result_num_broken(orc_result* r) { return r->broken; }
int64_t orc_result_num_skipped(orc_result* r) { return r->skipped; }
int64_t orc_result_num_time_buckets(orc_result* r) { return (int64_t)r->time_keys.size(); }
int64_t orc_result_time_bucket(orc_result* r, int64_t b) { return r->time_keys[(size_t)b]; }

// tb < 0: Results; tb >= 0: TimeResults[time_keys[tb]].  i == -1: Cumulative.
static Result* pick(orc_result* r, int64_t tb, int64_t i) {
  if (tb < 0) {
    if (i < 0) return r->c.Cumulative.get();
    if ((size_t)i >= r->c.Sorted.size()) return nullptr;
    return r->c.Sorted[(size_t)i].get();
  }
  if ((size_t)tb >= r->time_keys.size()) return nullptr;
  auto& v = r->c.TimeSorted[r->time_keys[(size_t)tb]];
  if (i < 0 || (size_t)i >= v.size()) return nullptr;
  return v[(size_t)i].get();
}
int64_t orc_result_time_num_groups(orc_result* r, int64_t tb) {
  if ((size_t)tb >= r->time_keys.size()) return 0;
  return (int64_t)r->c.TimeSorted[r->time_keys[(size_t)tb]].size();
}
int orc_result_group(orc_result* r, int64_t tb, int64_t i, const char** key, int64_t* keylen, int64_t* count,
                     int64_t* samples) {
  Result* x = pick(r, tb, i);
  if (!x) return -1;
  *key = x->GroupByKey.data();
  *keylen = (int64_t)x->GroupByKey.size();
  *count = x->Count;
  *samples = x->Samples;
  return 0;
}
// 1 when the group holds a hist for aggregation #agg (Q7), else 0
int orc_result_hist(orc_result* r, int64_t tb, int64_t i, int32_t agg, int64_t* count, int64_t* exact_sum,
                    int64_t* min, int64_t* max, double* avg, int64_t* samples, int32_t* num_buckets,
                    int32_t* bucket_size, int32_t* nvalues, int32_t* nsub, int64_t* noutliers) {
  Result* x = pick(r, tb, i);
  if (!x) return -1;
  auto it = x->Hists.find(agg);
  if (it == x->Hists.end()) return 0;
  const Hist& h = it->second;
  *count = h.TotalCount();
  *exact_sum = h.ExactSum();
  *min = h.MinV();
  *max = h.MaxV();
  *avg = h.Mean();
  *samples = h.Samples();
  if (!h.multi) {
    *num_buckets = (int32_t)h.b.NumBuckets;
    *bucket_size = (int32_t)h.b.BucketSize;
    *nvalues = (int32_t)h.b.Values.size();
    *nsub = 0;
    *noutliers = (int64_t)(h.b.Outliers.size() + h.b.Underliers.size());
  } else {
    *num_buckets = 0;
    *bucket_size = 0;
    int64_t nv = 0, no = 0;
    for (auto& sh : h.m.Subhists) {
      nv += (int64_t)sh.Values.size();
      no += (int64_t)(sh.Outliers.size() + sh.Underliers.size());
    }
    *nvalues = (int32_t)nv;
    *nsub = (int32_t)h.m.Subhists.size();
    *noutliers = no;
  }
  return 1;
}
// bucket counters: Values (basic) or the subhists' Values concatenated (multi)
int64_t orc_result_hist_values(orc_result* r, int64_t tb, int64_t i, int32_t agg, int64_t* out, int64_t cap) {
  Result* x = pick(r, tb, i);
  if (!x) return -1;
  auto it = x->Hists.find(agg);
  if (it == x->Hists.end()) return 0;
  const Hist& h = it->second;
  int64_t n = 0;
  auto emit = [&](const BasicHist& b) {
    for (int64_t v : b.Values) {
      if (out && n < cap) out[n] = v;
      n++;
    }
  };
  if (!h.multi)
    emit(h.b);
  else
    for (auto& sh : h.m.Subhists) emit(sh);
  return n;
}
// Q9: a group's final Result object is the first block's (ResultMap.Combine adopts it by
// reference, query_spec.go:107-116) and so may still carry that block's Outliers, while every
// later block is merged through BasicHist.Combine, which drops them.  With the merged view on,
// the derived quantities are taken from a fresh clone + Combine — what the reference reports
// for Cumulative and for any group whose first block had no outliers.
void orc_result_set_merged_view(orc_result* r, int on) { r->merged_view = on != 0; }
static Hist view_of(orc_result* r, const Hist& h) {
  if (!r->merged_view) return h;
  Hist nh = h.NewHist(&r->flags);
  nh.Combine(h);
  return nh;
}

int orc_result_percentiles(orc_result* r, int64_t tb, int64_t i, int32_t agg, int64_t* out100) {
  Result* x = pick(r, tb, i);
  if (!x) return -1;
  auto it = x->Hists.find(agg);
  if (it == x->Hists.end()) return 0;
  auto p = view_of(r, it->second).GetPercentiles();
  for (size_t k = 0; k < p.size() && k < 100; k++) out100[k] = p[k];
  return (int)p.size();
}
double orc_result_stddev(orc_result* r, int64_t tb, int64_t i, int32_t agg) {
  Result* x = pick(r, tb, i);
  if (!x) return NAN;
  auto it = x->Hists.find(agg);
  if (it == x->Hists.end()) return NAN;
  return view_of(r, it->second).StdDev();
}
int64_t orc_result_sparse_buckets(orc_result* r, int64_t tb, int64_t i, int32_t agg, int64_t* edges, int64_t* counts,
                                  int64_t cap) {
  Result* x = pick(r, tb, i);
  if (!x) return -1;
  auto it = x->Hists.find(agg);
  if (it == x->Hists.end()) return 0;
  auto m = view_of(r, it->second).GetSparseBuckets();
  int64_t n = 0;
  for (auto& kv : m) {
    if (edges && n < cap) {
      edges[n] = kv.first;
      counts[n] = kv.second;
    }
    n++;
  }
  return n;
}

// ---- direct access to the hist arithmetic, for the golden-vector test --------
// Builds a BasicHist from IntInfo{min,max} in hist mode and reports its layout
// (hist_basic.go:34-70).
void orc_basic_layout(int64_t info_min, int64_t info_max, int32_t hist_bucket, int64_t* num_buckets,
                      int64_t* bucket_size, int64_t* nvalues) {
  Flags f;
  f.op_hist = true;
  f.hist_bucket = hist_bucket;
  IntInfo info;
  info.Min = info_min;
  info.Max = info_max;
  BasicHist h = newBasicHist(&f, info);
  *num_buckets = h.NumBuckets;
  *bucket_size = h.BucketSize;
  *nvalues = (int64_t)h.Values.size();
}
// MultiHist sub-range layout (hist_multi.go:223-257): writes up to cap entries of
// {info_min, info_max, bucket_size, nvalues}; returns the number of subhists
int64_t orc_multi_layout(int64_t info_min, int64_t info_max, int64_t* out4, int64_t cap) {
  Flags f;
  f.op_hist = true;
  f.log_hist = true;
  IntInfo info;
  info.Min = info_min;
  info.Max = info_max;
  MultiHist h = newMultiHist(&f, info);
  int64_t n = 0;
  for (auto& sh : h.Subhists) {
    if (n < cap) {
      out4[n * 4 + 0] = sh.Info.Min;
      out4[n * 4 + 1] = sh.Info.Max;
      out4[n * 4 + 2] = sh.BucketSize;
      out4[n * 4 + 3] = (int64_t)sh.Values.size();
    }
    n++;
  }
  return n;
}
// Combine n BasicHist states given as (count, avg, values[nvalues]) in order into a
// fresh hist (Result.Combine's clone-then-Combine, query_spec.go:168-176) and
// return Count/Avg/Values — pins hist_basic.go:259-279 against the golden file.
void orc_basic_combine(int64_t info_min, int64_t info_max, int64_t n, const int64_t* counts, const double* avgs,
                       const int64_t* values, int64_t nvalues, int64_t* out_count, double* out_avg,
                       int64_t* out_values) {
  Flags f;
  f.op_hist = true;
  IntInfo info;
  info.Min = info_min;
  info.Max = info_max;
  BasicHist acc = newBasicHist(&f, info);
  for (int64_t i = 0; i < n; i++) {
    BasicHist h = newBasicHist(&f, info);
    h.Count = counts[i];
    h.Avg = avgs[i];
    for (int64_t k = 0; k < nvalues && k < (int64_t)h.Values.size(); k++) h.Values[(size_t)k] = values[i * nvalues + k];
    acc.Combine(h);
  }
  *out_count = acc.Count;
  *out_avg = acc.Avg;
  for (int64_t k = 0; k < nvalues && k < (int64_t)acc.Values.size(); k++) out_values[k] = acc.Values[(size_t)k];
}
// GetPercentiles / GetStdDev of a BasicHist given its state (hist_basic.go:153-219)
int orc_basic_percentiles(int64_t info_min, int64_t info_max, int64_t count, const int64_t* values, int64_t nvalues,
                          int64_t* out100) {
  Flags f;
  f.op_hist = true;
  IntInfo info;
  info.Min = info_min;
  info.Max = info_max;
  BasicHist h = newBasicHist(&f, info);
  h.Count = count;
  for (int64_t k = 0; k < nvalues && k < (int64_t)h.Values.size(); k++) h.Values[(size_t)k] = values[k];
  auto p = h.GetPercentiles();
  for (size_t k = 0; k < p.size() && k < 100; k++) out100[k] = p[k];
  return (int)p.size();
}

int orc_hardware_threads(void) {
  unsigned n = std::thread::hardware_concurrency();
  return n ? (int)n : 1;
}

}  // extern "C"
