// sg_kernels.cu — the fused decode + filter + group-by + aggregate scan kernel
// for sm_100a.
//
// What it replaces in the reference, per 65,536-row block:
//   unpackIntCol / unpackStrCol     src/lib/column_store_io.go:690-780 / :493-609
//   IntFilter.Filter / StrFilter    src/lib/filter.go:171-250
//   FilterAndAggRecords             src/lib/aggregate.go:56-282
//   BasicHist/MultiHist.AddWeightedValue   src/lib/hist_basic.go:101-151, hist_multi.go:48-88
//
// Shape of the kernel:
//   * persistent grid, one CTA of 512 threads per SM; CTAs pull blocks from an
//     atomic work counter;
//   * per block the CTA keeps ONE word per row in shared memory ("slot": dense
//     group index in the low bits, number of filters passed above them) and walks
//     the referenced columns one after another.  The encoded arrays are streamed
//     from HBM exactly once with 128-bit loads; nothing decoded is written back;
//   * a bucket-encoded column (value -> delta-encoded row-id list) is decoded by
//     a segmented prefix sum over the flat id array (segment heads = bin starts,
//     kept as a 65,536-bit mask in shared memory) and scattered into the slot
//     words with plain byte/halfword stores (a row appears in one bin only);
//   * a value-array column is decoded by a block-wide int64 prefix sum in row
//     order;
//   * count / sum accumulators live in shared memory, replicated per lane so
//     that the 32-bit shared atomics of a warp never collide (64-bit shared
//     atomics are CAS loops on this architecture); sums are kept exact in two
//     32-bit limbs with explicit carry.  Histogram bucket counters go to HBM/L2
//     with 64-bit reductions (RED.ADD.64).
// No tensor cores: this is integer / indexing work bound by HBM bandwidth.
#include <cuda_runtime.h>

#include <cstdint>

#include "sg_internal.h"

namespace sg {

constexpr int THREADS = 512;
constexpr int NWARPS = THREADS / 32;
constexpr int U = 4;  // 128-bit loads in flight per lane per chunk
constexpr uint32_t FLAG = 0x80000000u;
constexpr uint32_t HEAD_WORDS = SG_BLOCK_ROWS / 32;  // 2048

int scan_threads() { return THREADS; }

// fixed shared-memory carve-out (bytes), in this order after the dynamic base:
//   headbits[2048] u32 | headprefix[2048] u16 | binpay[SMEM_BINS] u32 | wtot[2][NWARPS] u64 | misc[64] u32
constexpr uint32_t OFF_HEADBITS = 0;
constexpr uint32_t OFF_HEADPREFIX = OFF_HEADBITS + HEAD_WORDS * 4;
constexpr uint32_t OFF_BINPAY = OFF_HEADPREFIX + HEAD_WORDS * 2;
constexpr uint32_t OFF_WTOT = OFF_BINPAY + SMEM_BINS * 4;
constexpr uint32_t OFF_MISC = OFF_WTOT + 2 * NWARPS * 8;
constexpr uint32_t FIXED_SMEM = OFF_MISC + 64 * 4;
uint32_t scan_fixed_smem() { return FIXED_SMEM; }

struct Ctx {  // per-CTA view of shared memory and the current block
  uint32_t* headbits;
  uint16_t* headprefix;
  uint32_t* binpay_s;
  unsigned long long* wtot;  // [2][NWARPS]
  uint32_t* misc;            // [0] next block, [1] broken flag, [16..31] warp totals
  uint32_t* acc;             // replicated accumulators (or nullptr)
  int tid, lane, warp;
};

__device__ __forceinline__ uint4 ldg128(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}

// 4 consecutive u32 starting at element idx (idx % 4 == 0), zero beyond n
__device__ __forceinline__ uint4 load4_u32(const uint32_t* __restrict__ p, uint32_t idx, uint32_t n) {
  if (idx + 4 <= n) return ldg128(p + idx);
  uint4 r = make_uint4(0, 0, 0, 0);
  if (idx < n) r.x = p[idx];
  if (idx + 1 < n) r.y = p[idx + 1];
  if (idx + 2 < n) r.z = p[idx + 2];
  return r;
}

// ---------------------------------------------------------------------------
// bucket-encoded column: visit(row, bin) for every (bin, row) pair
// ---------------------------------------------------------------------------
template <class Visit>
__device__ __forceinline__ void scan_bucket(const Ctx& cx, const DevCol& c, uint32_t nrec, Visit visit) {
  const uint32_t n = c.nitems;
  const uint32_t* __restrict__ ids = reinterpret_cast<const uint32_t*>(c.data);
  const bool delta = (c.flags & COL_DELTA_IDS) != 0;
  const int tid = cx.tid, lane = cx.lane, warp = cx.warp;

  // segment heads: one bit per flat entry that starts a bin (bins are non-empty)
  for (uint32_t i = tid; i < HEAD_WORDS; i += THREADS) cx.headbits[i] = 0;
  __syncthreads();
  for (uint32_t b = tid; b < c.nbins; b += THREADS) {
    uint32_t o = c.bin_offsets[b];
    if (o < n) atomicOr(&cx.headbits[o >> 5], 1u << (o & 31));
  }
  __syncthreads();
  {  // exclusive prefix popcount per 32-entry word: 512 threads x 4 words
    uint4 w = reinterpret_cast<const uint4*>(cx.headbits)[tid];
    uint32_t p0 = __popc(w.x), p1 = __popc(w.y), p2 = __popc(w.z), p3 = __popc(w.w);
    uint32_t tot = p0 + p1 + p2 + p3, inc = tot;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
      if (lane >= d) inc += t;
    }
    if (lane == 31) cx.misc[16 + warp] = inc;
    __syncthreads();
    uint32_t base = 0;
    for (int i = 0; i < warp; i++) base += cx.misc[16 + i];
    uint32_t ex = base + inc - tot;
    cx.headprefix[tid * 4 + 0] = (uint16_t)ex;
    cx.headprefix[tid * 4 + 1] = (uint16_t)(ex + p0);
    cx.headprefix[tid * 4 + 2] = (uint16_t)(ex + p0 + p1);
    cx.headprefix[tid * 4 + 3] = (uint16_t)(ex + p0 + p1 + p2);
  }
  __syncthreads();

  constexpr uint32_t CH = THREADS * 4 * U;  // entries per chunk
  uint32_t carry = 0;                       // running sum of the open segment at chunk start
  int buf = 0;
  for (uint32_t base = 0; base < n; base += CH, buf ^= 1) {
    const uint32_t wbase = base + warp * (U * 128);
    uint4 d[U];
#pragma unroll
    for (int u = 0; u < U; u++) d[u] = load4_u32(ids, wbase + u * 128 + lane * 4, n);

    uint32_t y[U][4];
    uint32_t nibs[U];   // head bits of the lane's 4 entries (true heads)
    uint32_t openm[U];  // bit k: entry k still needs the cross-warp carry
    uint32_t binb[U];   // heads strictly before the lane's first entry
    uint32_t run_sum = 0, run_flag = 0;
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t idx = wbase + u * 128 + lane * 4;
      uint32_t word = 0, hp = 0;
      if (idx < n) {
        word = cx.headbits[idx >> 5];
        hp = cx.headprefix[idx >> 5];
      }
      const uint32_t sh = idx & 31;
      const uint32_t nib_true = (word >> sh) & 0xFu;
      binb[u] = hp + __popc(word & ((1u << sh) - 1u));
      nibs[u] = nib_true;
      const uint32_t nib = delta ? nib_true : 0xFu;  // absolute ids: every entry is its own segment
      uint32_t a0 = d[u].x, a1 = d[u].y, a2 = d[u].z, a3 = d[u].w;
      // a valid gap is < 65,536; anything larger marks the block broken (and is
      // zeroed so that the packed scan word cannot overflow into the flag bit)
      if ((a0 | a1 | a2 | a3) >= 0x10000u) {
        cx.misc[1] = 1;
        a0 &= 0xFFFFu; a1 &= 0xFFFFu; a2 &= 0xFFFFu; a3 &= 0xFFFFu;
      }
      const uint32_t x0 = a0;
      const uint32_t x1 = (nib & 2u) ? a1 : x0 + a1;
      const uint32_t x2 = (nib & 4u) ? a2 : x1 + a2;
      const uint32_t x3 = (nib & 8u) ? a3 : x2 + a3;
      uint32_t agg = x3 | (nib ? FLAG : 0u);
#pragma unroll
      for (int dd = 1; dd < 32; dd <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, agg, dd);
        if (lane >= dd && !(agg & FLAG)) agg += t;
      }
      uint32_t excl = __shfl_up_sync(0xffffffffu, agg, 1);
      if (lane == 0) excl = 0;
      // segment sum entering this lane (without the cross-warp carry)
      const uint32_t pre = (excl & FLAG) ? (excl & ~FLAG) : ((excl + run_sum) & ~FLAG);
      const uint32_t pre_flag = (excl & FLAG) | run_flag;
      y[u][0] = (nib & 1u) ? x0 : x0 + pre;
      y[u][1] = (nib & 3u) ? x1 : x1 + pre;
      y[u][2] = (nib & 7u) ? x2 : x2 + pre;
      y[u][3] = (nib & 15u) ? x3 : x3 + pre;
      uint32_t om = 0;
      if (!pre_flag) {
        om = (nib & 1u) ? 0u : 1u;
        om |= (nib & 3u) ? 0u : 2u;
        om |= (nib & 7u) ? 0u : 4u;
        om |= (nib & 15u) ? 0u : 8u;
      }
      openm[u] = om;
      const uint32_t last = __shfl_sync(0xffffffffu, agg, 31);
      run_sum = (last & FLAG) ? (last & ~FLAG) : ((last + run_sum) & ~FLAG);
      run_flag |= (last & FLAG);
    }
    if (lane == 0) cx.wtot[buf * NWARPS + warp] = (unsigned long long)(run_sum | run_flag);
    __syncthreads();
    uint32_t cin = carry, call = carry;
#pragma unroll
    for (int i = 0; i < NWARPS; i++) {
      const uint32_t t = (uint32_t)cx.wtot[buf * NWARPS + i];
      call = (t & FLAG) ? (t & ~FLAG) : ((call + t) & ~FLAG);
      if (i + 1 == warp) cin = call;
    }
    if (warp == 0) cin = carry;
    carry = call;

#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t idx = wbase + u * 128 + lane * 4;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        if (idx + k < n) {
          uint32_t row = y[u][k] + ((openm[u] >> k) & 1u ? cin : 0u);
          uint32_t bin = binb[u] + __popc(nibs[u] & ((2u << k) - 1u)) - 1u;
          if (row >= nrec || bin >= c.nbins) {
            cx.misc[1] = 1;  // "BLOCK SIZE CHANGED DURING QUERY" (column_store_io.go:733)
          } else {
            visit(row, bin);
          }
        }
      }
    }
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// value-array int column (delta-encoded int64): visit(row, value) in row order
// ---------------------------------------------------------------------------
template <class Visit>
__device__ __forceinline__ void scan_values_i64(const Ctx& cx, const DevCol& c, uint32_t nrec, Visit visit) {
  uint32_t n = c.nitems;
  if (n > nrec) n = nrec;  // staging already flags len(Values) > NumRecords as broken
  const unsigned long long* __restrict__ vals = reinterpret_cast<const unsigned long long*>(c.data);
  const bool delta = (c.flags & COL_DELTA_VALUES) != 0;
  const int lane = cx.lane, warp = cx.warp;
  constexpr uint32_t CH = THREADS * 2 * U;
  unsigned long long carry = 0;
  int buf = 0;
  for (uint32_t base = 0; base < n; base += CH, buf ^= 1) {
    const uint32_t wbase = base + warp * (U * 64);
    unsigned long long a[U][2];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t idx = wbase + u * 64 + lane * 2;
      if (idx + 2 <= n) {
        uint4 r = ldg128(vals + idx);
        a[u][0] = (unsigned long long)r.x | ((unsigned long long)r.y << 32);
        a[u][1] = (unsigned long long)r.z | ((unsigned long long)r.w << 32);
      } else {
        a[u][0] = idx < n ? vals[idx] : 0ull;
        a[u][1] = 0ull;
      }
    }
    if (delta) {
      unsigned long long run = 0;
#pragma unroll
      for (int u = 0; u < U; u++) {
        a[u][1] += a[u][0];
        unsigned long long inc = a[u][1];
#pragma unroll
        for (int dd = 1; dd < 32; dd <<= 1) {
          unsigned long long t = __shfl_up_sync(0xffffffffu, inc, dd);
          if (lane >= dd) inc += t;
        }
        unsigned long long ex = inc - a[u][1] + run;
        a[u][0] += ex;
        a[u][1] += ex;
        run += __shfl_sync(0xffffffffu, inc, 31);
      }
      if (lane == 0) cx.wtot[buf * NWARPS + warp] = run;
      __syncthreads();
      unsigned long long cin = carry, call = carry;
#pragma unroll
      for (int i = 0; i < NWARPS; i++) {
        if (i == warp) cin = call;
        call += cx.wtot[buf * NWARPS + i];
      }
      carry = call;
#pragma unroll
      for (int u = 0; u < U; u++) {
        a[u][0] += cin;
        a[u][1] += cin;
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      const uint32_t idx = wbase + u * 64 + lane * 2;
      if (idx < n) visit(idx, (long long)a[u][0]);
      if (idx + 1 < n) visit(idx + 1, (long long)a[u][1]);
    }
  }
  __syncthreads();
}

// value-array str column (raw int32 local ids): visit(row, local_id)
template <class Visit>
__device__ __forceinline__ void scan_values_i32(const Ctx& cx, const DevCol& c, uint32_t nrec, Visit visit) {
  uint32_t n = c.nitems;
  if (n > nrec) n = nrec;
  const uint32_t* __restrict__ vals = reinterpret_cast<const uint32_t*>(c.data);
  for (uint32_t idx = cx.tid * 4; idx < n; idx += THREADS * 4) {
    uint4 r = load4_u32(vals, idx, n);
    visit(idx, (int32_t)r.x);
    if (idx + 1 < n) visit(idx + 1, (int32_t)r.y);
    if (idx + 2 < n) visit(idx + 2, (int32_t)r.z);
    if (idx + 3 < n) visit(idx + 3, (int32_t)r.w);
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------
// predicates
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool int_pred(int op, long long v, long long lit) {
  // filter.go:177-189
  switch (op) {
    case SG_OP_GT: return v > lit;
    case SG_OP_LT: return v < lit;
    case SG_OP_EQ: return v == lit;
    case SG_OP_NEQ: return v != lit;
    default: return false;
  }
}
__device__ __forceinline__ bool str_pred(const KFilter& f, int32_t gid) {
  // filter.go:199-250 on global ids: EQ/NEQ against the literal's id (a literal the
  // dictionary does not hold matches nothing, Q3); RE/NRE through the host's bitset
  switch (f.op) {
    case SG_OP_EQ: return gid == f.str_gid;
    case SG_OP_NEQ: return gid != f.str_gid;
    case SG_OP_RE:
    case SG_OP_NRE: {
      bool m = false;
      if (gid >= 0 && (long long)gid < f.lut_bits) m = (f.lut[gid >> 5] >> (gid & 31)) & 1u;
      return f.op == SG_OP_RE ? m : !m;
    }
    default: return false;
  }
}
__device__ __forceinline__ int32_t str_gid(const DevCol& c, long long local) {
  if (local < 0 || local >= (long long)c.nremap) return c.oob_gid;
  return c.remap[local];
}

// time bucket code (aggregate.go:177: int(val)/TimeBucket*TimeBucket, truncating):
// dense index 1.. of trunc(val/bucket) - first; 0 = outside the planned range
__device__ __forceinline__ uint32_t time_code(const Plan& P, long long v) {
  long long q = v / P.time_bucket - P.time_first;
  if (q < 0 || q >= (long long)(P.time_radix - 1)) return 0u;
  return (uint32_t)q + 1u;
}

// ---------------------------------------------------------------------------
// accumulation of one accepted/considered value
// ---------------------------------------------------------------------------
template <bool ACC_SMEM>
__device__ __forceinline__ void agg_value(const Ctx& cx, const Plan& P, const KAgg& A, int ai, uint32_t g, long long v) {
  // BasicHist/MultiHist.AddWeightedValue (hist_basic.go:104, hist_multi.go:52)
  if (v > A.reject_hi || v < A.info_min) return;
  if (ACC_SMEM) {
    const uint32_t R = P.acc_repl;
    uint32_t* w = cx.acc + ((size_t)g * P.acc_words + 1 + 3 * ai) * R + (cx.lane & (R - 1));
    atomicAdd(w, 1u);
    const uint32_t lo = (uint32_t)(unsigned long long)v;
    uint32_t hi = (uint32_t)((unsigned long long)v >> 32);
    const uint32_t old = atomicAdd(w + R, lo);
    if (old > ~lo) hi += 1u;  // carry out of the low limb
    if (hi) atomicAdd(w + 2 * R, hi);
  } else {
    atomicAdd(reinterpret_cast<unsigned long long*>(A.hcount) + g, 1ull);
    atomicAdd(reinterpret_cast<unsigned long long*>(A.sum) + g, (unsigned long long)v);
  }
  if (v > A.info_max) atomicMax(reinterpret_cast<long long*>(A.vmax) + g, v);
  if (A.nsub > 0) {
    // BasicHist: the one layout; MultiHist: first sub-range containing v (hist_multi.go:81-86)
    for (int s = 0; s < A.nsub; s++) {
      const KSubHist& S = A.sub[s];
      if (A.nsub > 1) {
        if (v < S.lo || v > S.hi) continue;
        if (v > S.reject_hi || v < S.lo) break;  // the subhist's own reject rule
      }
      unsigned long long x = (unsigned long long)v - (unsigned long long)S.lo;
      unsigned long long b;
      if (x < 0x100000000ull && (unsigned long long)S.bsize < 0x100000000ull)
        b = (uint32_t)x / (uint32_t)S.bsize;
      else
        b = (unsigned long long)((long long)x / S.bsize);
      if (b >= S.nvals) b = S.nvals - 1;  // outlier: clamped into the last slot (:134-137)
      atomicAdd(reinterpret_cast<unsigned long long*>(A.buckets) + ((size_t)g * A.nvals_total + S.base + b), 1ull);
      break;
    }
  }
}

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
template <typename SlotT, bool ACC_SMEM>
__global__ void __launch_bounds__(THREADS, 1) scan_kernel(const LaunchParams lp) {
  extern __shared__ __align__(16) unsigned char smem[];
  const Plan& P = *lp.plan;
  Ctx cx;
  cx.tid = threadIdx.x;
  cx.lane = threadIdx.x & 31;
  cx.warp = threadIdx.x >> 5;
  cx.headbits = reinterpret_cast<uint32_t*>(smem + OFF_HEADBITS);
  cx.headprefix = reinterpret_cast<uint16_t*>(smem + OFF_HEADPREFIX);
  cx.binpay_s = reinterpret_cast<uint32_t*>(smem + OFF_BINPAY);
  cx.wtot = reinterpret_cast<unsigned long long*>(smem + OFF_WTOT);
  cx.misc = reinterpret_cast<uint32_t*>(smem + OFF_MISC);
  SlotT* slot;
  uint32_t acc_off = FIXED_SMEM;
  if (sizeof(SlotT) == 4) {
    slot = reinterpret_cast<SlotT*>(lp.gslots + (size_t)blockIdx.x * SG_BLOCK_ROWS);
  } else {
    slot = reinterpret_cast<SlotT*>(smem + FIXED_SMEM);
    acc_off = FIXED_SMEM + SG_BLOCK_ROWS * sizeof(SlotT);
  }
  cx.acc = ACC_SMEM ? reinterpret_cast<uint32_t*>(smem + acc_off) : nullptr;
  uint32_t* gbinpay = lp.gbinpay + (size_t)blockIdx.x * SG_BLOCK_ROWS;
  const uint32_t acc_total = ACC_SMEM ? P.nslots * P.acc_words * P.acc_repl : 0u;
  for (uint32_t i = cx.tid; i < acc_total; i += THREADS) cx.acc[i] = 0;
  unsigned long long matched = 0;
  const uint32_t gmask = (1u << P.gbits) - 1u;

  for (;;) {
    __syncthreads();
    if (cx.tid == 0) {
      cx.misc[0] = atomicAdd(lp.work_counter, 1u);
      cx.misc[1] = 0;
    }
    __syncthreads();
    const uint32_t wi = cx.misc[0];
    if (wi >= lp.nlist) break;
    const uint32_t bid = lp.block_list[wi];
    const uint32_t nrec = lp.blocks[bid].num_records;
    const DevCol* cols = lp.cols + (size_t)bid * P.ncolslots;

    // every row starts as: group slot 0, no filter passed
    for (uint32_t r = cx.tid; r < nrec; r += THREADS) slot[r] = 0;
    __syncthreads();

    // ---- filters (aggregate.go:105-112; unpopulated -> false, Q1) -----------------
    for (int fi = 0; fi < P.nfilters; fi++) {
      const KFilter& F = P.filters[fi];
      const DevCol& c = cols[F.col];
      const SlotT finc = (SlotT)P.finc;
      if (c.enc == SG_ENC_BUCKET) {
        uint32_t* pay = c.nbins <= SMEM_BINS ? cx.binpay_s : gbinpay;
        for (uint32_t b = cx.tid; b < c.nbins; b += THREADS) {
          const long long bv = c.bin_values[b];
          pay[b] = F.is_str ? (str_pred(F, str_gid(c, bv)) ? 1u : 0u) : (int_pred(F.op, bv, F.ival) ? 1u : 0u);
        }
        __syncthreads();
        scan_bucket(cx, c, nrec, [&](uint32_t row, uint32_t bin) {
          if (pay[bin]) slot[row] = (SlotT)(slot[row] + finc);
        });
      } else if (c.enc == SG_ENC_VALUES) {
        if (F.is_str) {
          scan_values_i32(cx, c, nrec, [&](uint32_t row, int32_t local) {
            if (str_pred(F, str_gid(c, local))) slot[row] = (SlotT)(slot[row] + finc);
          });
        } else {
          scan_values_i64(cx, c, nrec, [&](uint32_t row, long long v) {
            if (int_pred(F.op, v, F.ival)) slot[row] = (SlotT)(slot[row] + finc);
          });
        }
      }
    }

    // ---- group key (aggregate.go:125-143) as a dense mixed-radix index ---------------
    for (int gi = 0; gi < P.ngroups; gi++) {
      const KGroup& G = P.groups[gi];
      const DevCol& c = cols[G.col];
      if (c.enc == SG_ENC_BUCKET) {
        uint32_t* pay = c.nbins <= SMEM_BINS ? cx.binpay_s : gbinpay;
        for (uint32_t b = cx.tid; b < c.nbins; b += THREADS) {
          int32_t code = G.is_str ? str_gid(c, c.bin_values[b]) : c.remap[b];
          pay[b] = ((uint32_t)code + 1u) * G.stride;
        }
        __syncthreads();
        scan_bucket(cx, c, nrec, [&](uint32_t row, uint32_t bin) { slot[row] = (SlotT)(slot[row] + pay[bin]); });
      } else if (c.enc == SG_ENC_VALUES && G.is_str) {
        scan_values_i32(cx, c, nrec, [&](uint32_t row, int32_t local) {
          slot[row] = (SlotT)(slot[row] + ((uint32_t)str_gid(c, local) + 1u) * G.stride);
        });
      }
      // VALUES int group columns are routed away from this kernel by the planner
    }

    // ---- time bucket (aggregate.go:146-183) ------------------------------------------
    if (P.time_col >= 0) {
      const DevCol& c = cols[P.time_col];
      const SlotT tok = (SlotT)P.time_ok;
      if (c.enc == SG_ENC_BUCKET && !(c.flags & COL_IS_STR)) {
        uint32_t* pay = c.nbins <= SMEM_BINS ? cx.binpay_s : gbinpay;
        for (uint32_t b = cx.tid; b < c.nbins; b += THREADS) pay[b] = time_code(P, c.bin_values[b]);
        __syncthreads();
        scan_bucket(cx, c, nrec, [&](uint32_t row, uint32_t bin) {
          const uint32_t tc = pay[bin];
          if (tc)
            slot[row] = (SlotT)(slot[row] + tc * P.time_stride + tok);
          else
            atomicAdd(reinterpret_cast<unsigned long long*>(P.scalars) + 2, 1ull);
        });
      } else if (c.enc == SG_ENC_VALUES && !(c.flags & COL_IS_STR)) {
        scan_values_i64(cx, c, nrec, [&](uint32_t row, long long v) {
          const uint32_t tc = time_code(P, v);
          if (tc)
            slot[row] = (SlotT)(slot[row] + tc * P.time_stride + tok);
          else
            atomicAdd(reinterpret_cast<unsigned long long*>(P.scalars) + 2, 1ull);
        });
      }
    }
    __syncthreads();

    // ---- Count / Samples (aggregate.go:202-203) and MatchedCount (:117) --------------
    unsigned long long my_matched = 0;
    for (uint32_t r = cx.tid; r < nrec; r += THREADS) {
      const uint32_t s = (uint32_t)slot[r];
      const uint32_t hi = s >> P.gbits;
      if ((hi & P.filt_mask) == P.filt_target) my_matched++;
      if (hi == P.pass_target) {
        const uint32_t g = s & gmask;
        if (ACC_SMEM)
          atomicAdd(cx.acc + ((size_t)g * P.acc_words) * P.acc_repl + (cx.lane & (P.acc_repl - 1)), 1u);
        else
          atomicAdd(reinterpret_cast<unsigned long long*>(P.count) + g, 1ull);
      }
    }

    // ---- aggregations (aggregate.go:246-261) -------------------------------------------
    for (int ai = 0; ai < P.naggs; ai++) {
      const KAgg& A = P.aggs[ai];
      const DevCol& c = cols[A.col];
      if (c.flags & COL_IS_STR) continue;  // Populated != INT_VAL: no update
      if (c.enc == SG_ENC_BUCKET) {
        scan_bucket(cx, c, nrec, [&](uint32_t row, uint32_t bin) {
          const uint32_t s = (uint32_t)slot[row];
          if ((s >> P.gbits) == P.pass_target) agg_value<ACC_SMEM>(cx, P, A, ai, s & gmask, c.bin_values[bin]);
        });
      } else if (c.enc == SG_ENC_VALUES) {
        scan_values_i64(cx, c, nrec, [&](uint32_t row, long long v) {
          const uint32_t s = (uint32_t)slot[row];
          if ((s >> P.gbits) == P.pass_target) agg_value<ACC_SMEM>(cx, P, A, ai, s & gmask, v);
        });
      }
    }
    __syncthreads();

    // ---- end of block: publish or discard ------------------------------------------------
    const bool broken = cx.misc[1] != 0;
    if (broken) {
      if (cx.tid == 0) {
        lp.plan->block_status[bid] = 1;
        atomicAdd(reinterpret_cast<unsigned long long*>(P.scalars) + 1, 1ull);
      }
    } else {
      matched += my_matched;
    }
    if (ACC_SMEM) {
      const uint32_t R = P.acc_repl;
      const uint32_t nw = P.nslots * P.acc_words;
      for (uint32_t w = cx.tid; w < nw; w += THREADS) {
        const uint32_t g = w / P.acc_words, k = w - g * P.acc_words;
        if (k != 0 && ((k - 1) % 3) == 2) continue;  // high limbs are folded with their low limb
        unsigned long long tot = 0;
        for (uint32_t r = 0; r < R; r++) {
          tot += cx.acc[(size_t)w * R + r];
          cx.acc[(size_t)w * R + r] = 0;
        }
        if (k != 0 && ((k - 1) % 3) == 1) {
          unsigned long long hi = 0;
          for (uint32_t r = 0; r < R; r++) {
            hi += cx.acc[(size_t)(w + 1) * R + r];
            cx.acc[(size_t)(w + 1) * R + r] = 0;
          }
          tot += hi << 32;
        }
        if (tot != 0 && !broken) {
          unsigned long long* dst;
          if (k == 0) {
            dst = reinterpret_cast<unsigned long long*>(P.count) + g;
          } else {
            const KAgg& A = P.aggs[(k - 1) / 3];
            dst = ((k - 1) % 3 == 0 ? reinterpret_cast<unsigned long long*>(A.hcount)
                                    : reinterpret_cast<unsigned long long*>(A.sum)) + g;
          }
          atomicAdd(dst, tot);
        }
      }
    }
  }

  // MatchedCount
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) matched += __shfl_xor_sync(0xffffffffu, matched, d);
  if (cx.lane == 0 && matched) atomicAdd(reinterpret_cast<unsigned long long*>(P.scalars) + 0, matched);
}

template <typename SlotT, bool ACC_SMEM>
static int launch_one(const LaunchParams& lp, int grid, cudaStream_t st) {
  auto k = scan_kernel<SlotT, ACC_SMEM>;
  cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lp.smem_bytes);
  if (e != cudaSuccess) return (int)e;
  k<<<grid, THREADS, lp.smem_bytes, st>>>(lp);
  return (int)cudaGetLastError();
}

int launch_scan(const LaunchParams& lp, int grid, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const bool acc_smem = lp.acc_smem != 0;
  const uint32_t sb = lp.slot_bytes;
  const LaunchParams& p = lp;
  if (sb == 1) return acc_smem ? launch_one<uint8_t, true>(p, grid, st) : launch_one<uint8_t, false>(p, grid, st);
  if (sb == 2) return acc_smem ? launch_one<uint16_t, true>(p, grid, st) : launch_one<uint16_t, false>(p, grid, st);
  return acc_smem ? launch_one<uint32_t, true>(p, grid, st) : launch_one<uint32_t, false>(p, grid, st);
}

}  // namespace sg
