// sg_hist.h — histogram bucket layouts and finalisation (host side of the product).
//
// The kernels only count; everything the reference derives from the counters at
// print time is computed here from the merged counters:
//   SetupBuckets        src/lib/hist_basic.go:34-70
//   TrackPercentiles    src/lib/hist_multi.go:223-257   (MultiHist sub-ranges)
//   GetPercentiles      src/lib/hist_basic.go:153-183, hist_multi.go:90-131
//   GetStdDev           src/lib/hist_basic.go:192-219, hist_multi.go:144-158
//   GetSparseBuckets    src/lib/hist_basic.go:221-239, hist_multi.go:184-200
#pragma once
#include <cmath>
#include <cstdint>
#include <map>
#include <vector>

#include "sg_internal.h"

namespace sg {

constexpr int64_t NUM_BUCKETS = 1000;  // hist.go:3

static inline int64_t wmul10(int64_t v) { return (int64_t)((uint64_t)v * 10ull); }
static inline int64_t wsub(int64_t a, int64_t b) { return (int64_t)((uint64_t)a - (uint64_t)b); }

// One BasicHist layout from its IntInfo extents.  Returns false when the extents
// cannot give a usable layout (negative size: the reference would panic in make()).
static inline bool basic_layout(int64_t lo, int64_t hi, int32_t hist_bucket, KSubHist& out, int64_t* num_buckets) {
  int64_t size = wsub(hi, lo);
  int64_t nb = NUM_BUCKETS;
  int64_t bs = size / NUM_BUCKETS;
  if (hist_bucket > 0) bs = hist_bucket;
  if (bs == 0) {
    if (size < 100) {
      bs = 1;
      nb = size;
    } else {
      bs = size / 100;
      nb = size / bs;
    }
  }
  nb += 1;
  int64_t nvals = nb + 1;
  if (nvals < 1 || nvals > (int64_t)1 << 24 || bs <= 0) return false;
  out.lo = lo;
  out.hi = hi;
  out.reject_hi = wmul10(hi);
  out.bsize = bs;
  out.nvals = (uint32_t)nvals;
  out.base = 0;
  out.magic = (bs >= 2 && bs < ((int64_t)1 << 32)) ? (uint64_t)((((unsigned __int128)1) << 64) / (uint64_t)bs) + 1 : 0;
  if (num_buckets) *num_buckets = nb;
  return true;
}

struct HistLayout {
  bool tracked = false;  // FLAGS.OP == "hist": bucket counters exist
  bool multi = false;    // FLAGS.LOG_HIST
  int64_t info_min = 0, info_max = 0;
  int64_t num_buckets = 0;  // basic only
  std::vector<KSubHist> subs;
  uint32_t nvals_total = 0;
};

static inline bool make_layout(int64_t info_min, int64_t info_max, bool hist_mode, bool multi, int32_t hist_bucket,
                               HistLayout& L) {
  L = HistLayout();
  L.info_min = info_min;
  L.info_max = info_max;
  L.multi = multi;
  L.tracked = hist_mode;
  if (!hist_mode) return true;
  if (!multi) {
    KSubHist s;
    if (!basic_layout(info_min, info_max, hist_bucket, s, &L.num_buckets)) return false;
    L.subs.push_back(s);
    L.nvals_total = s.nvals;
    return true;
  }
  // hist_multi.go:223-257: halve the range from the right edge until <= NUM_BUCKETS
  int64_t bucket = wsub(info_max, info_min);
  int num_hists = 0;
  for (int64_t t = bucket; t > NUM_BUCKETS; t >>= 1) num_hists++;
  if (num_hists + 1 > MAX_SUBHISTS) return false;
  int64_t right_edge = info_max;
  uint32_t base = 0;
  for (int i = 0; i <= num_hists; i++) {
    int64_t lo, hi;
    if (i < num_hists) {
      bucket >>= 1;
      lo = wsub(right_edge, bucket);
      hi = right_edge;
      right_edge = lo;
    } else {
      lo = info_min;
      hi = right_edge;
    }
    KSubHist s;
    if (!basic_layout(lo, hi, hist_bucket, s, nullptr)) return false;
    s.base = base;
    base += s.nvals;
    L.subs.push_back(s);
  }
  L.nvals_total = base;
  return true;
}

// edge -> count over the merged counters (no Outliers survive Combine, hist_basic.go:259-279)
static inline std::map<int64_t, int64_t> sparse_buckets(const HistLayout& L, const int64_t* values) {
  std::map<int64_t, int64_t> m;
  for (auto& s : L.subs)
    for (uint32_t k = 0; k < s.nvals; k++) {
      int64_t c = values[s.base + k];
      if (c > 0) m[(int64_t)k * s.bsize + s.lo] += c;
    }
  return m;
}

// returns the number of entries written (0 when Count == 0), at most 100
static inline int percentiles(const HistLayout& L, const int64_t* values, int64_t count, int64_t hmin, int64_t* out100) {
  if (count == 0 || !L.tracked) return 0;
  int64_t p101[101];
  for (int i = 0; i < 101; i++) p101[i] = 0;
  if (!L.multi) {
    const KSubHist& s = L.subs[0];
    p101[0] = hmin;
    int64_t c = 0, prev_p = 0;
    for (uint32_t k = 0; k < s.nvals; k++) {
      c += values[k];
      int64_t p = (100 * c) / count;
      for (int64_t ip = prev_p; ip <= p; ip++)
        if (ip >= 0 && ip <= 100) p101[ip] = (int64_t)k * s.bsize + hmin;
      if (p >= 0 && p <= 100) p101[p] = (int64_t)k;
      prev_p = p;
    }
  } else {
    auto all = sparse_buckets(L, values);
    int64_t total = 0;
    for (auto& kv : all) total += kv.second;
    if (total > 0) {
      int64_t prev_p = 0, c = 0;
      for (auto& kv : all) {
        c += kv.second;
        int64_t p = (100 * c) / total;
        for (int64_t ip = prev_p; ip <= p; ip++)
          if (ip >= 0 && ip <= 100) p101[ip] = kv.first;
        if (p >= 0 && p <= 100) p101[p] = kv.first;
        prev_p = p;
      }
    }
  }
  for (int i = 0; i < 100; i++) out100[i] = p101[i];
  return 100;
}

static inline double stddev(const HistLayout& L, const int64_t* values, int64_t count, double avg, int64_t hmin) {
  if (!L.tracked) return std::sqrt(0.0 / 1.0 * 0.0);
  double sum_variance = 0;
  if (!L.multi) {
    const KSubHist& s = L.subs[0];
    for (uint32_t b = 0; b < s.nvals; b++) {
      int64_t val = (int64_t)b * s.bsize + hmin;
      double delta = (double)val - avg;
      double ratio = (double)values[b] / (double)count;
      sum_variance += (delta * delta) * ratio;
    }
  } else {
    auto all = sparse_buckets(L, values);
    for (auto& kv : all) {
      double delta = (double)kv.first - avg;
      double ratio = (double)kv.second / (double)count;
      sum_variance += (delta * delta) * ratio;
    }
  }
  return std::sqrt(sum_variance);
}

}  // namespace sg
