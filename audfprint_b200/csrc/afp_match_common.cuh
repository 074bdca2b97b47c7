// Pieces shared by the two matching kernels (afp_match.cu: general path, afp_match_fast.cu:
// fast path): launch geometry, kernel arguments, candidate order, block-wide helpers and the
// per-candidate time-offset histogram mode search (audfprint_match.py:284-311).
#pragma once
#include <math.h>
#include "afp_internal.cuh"

namespace {

constexpr int MT = 1024;          // threads per matching CTA (one CTA per SM, persistent over the queries)
constexpr int NW = MT / 32;
constexpr int KCAP = 1024;        // candidate depth handled by the fast path (search_depth <= KCAP)
constexpr int GCAP = 1024;        // radix select stops once the undecided set is this small
constexpr int QCAP = 16384;       // query rows sorted in shared memory to merge probes of one bucket (128 KB)
constexpr int CSEG = 32768;       // track ids counted per pass in shared memory (u32 counters, the same 128 KB)
constexpr int64_t HITS_MAX = (int64_t)1 << 30;   // per-query hit capacity (rows * depth): int indexing
constexpr int HSET_BITS = 11;     // candidate hash set: 2048 entries for <= KCAP = 1024 keys
constexpr int HSET = 1 << HSET_BITS;

struct MatchArgs {
  const int32_t* q;        // [sum nq][2]
  const int64_t* qoff;     // [nq+1] (device)
  int nqueries;
  const uint32_t* table;
  const int32_t* counts;
  const uint32_t* hpi;
  int hashbits, depth, mtb;
  int64_t nids;
  int window, thresh, sdepth, maxalign;
  // per-CTA scratch (stride in elements)
  uint2* hits;      int64_t hits_cap;     // (id, dt + bias)
  uint32_t* dlist;                        // distinct ids, hits_cap
  double* wtd;                            // weighted count per dlist entry, hits_cap
  uint32_t* dts;                          // dt + bias of the candidates' hits, grouped per candidate, hits_cap
  uint32_t* recs;                         // (id << 8 | weight) of every distinct (bucket, slot), hits_cap
  uint32_t* rawl;                         // raw count per dlist entry, hits_cap
  int32_t* hist;    int hist_len;         // dtime histogram
  int32_t* filt;                          // local-max filtered copy
  int bias;
  int32_t* rows;    int row_cap;          // [nqueries][row_cap][7]
  int32_t* row_cnt;                       // [nqueries] rows produced (may exceed row_cap)
  // sharded-table mode: publish every query's local top-sdepth candidate list
  int publish;                            // 0/1
  double* cand;                           // [nqueries][sdepth][3] = (id, raw, weight)
  int32_t* cand_cnt;                      // [nqueries][2] = (entries, n_above)
  // work list of the general kernel (NULL = every query); the fast kernel appends the queries
  // it hands over (capacity overflow, parameters outside its limits)
  int32_t* qlist;
  int* nlist;
  int32_t* fstat;                         // [nqueries] 0 = done by the fast kernel, else the reason it was not
  // fast path (afp_match_fast.cu)
  uint2* mhits;     int mh_cap;           // per-CTA list of the hits of multi-record ids: (set slot, dt + bias)
  unsigned hmin;                          // smallest hashesperid of the table; 0 = unknown (no pruning)
  int bm_exact;                           // nids <= bitmap bits: the repeat bitmap is indexed by the id itself
};

// candidate order: (weighted count desc, id desc); keys are (bits of the positive double, id)
__device__ __forceinline__ bool key_gt(unsigned long long w1, unsigned i1, unsigned long long w2, unsigned i2) {
  return w1 > w2 || (w1 == w2 && i1 > i2);
}

// scratch the block-wide helpers need, embedded in each kernel's shared-memory struct
struct ModeScratch {
  int val[NW], idx[NW];
  int nrows;
};

struct Shared {
  ModeScratch ms;
  unsigned long long a_w[KCAP + GCAP];   // gathered keys, sorted descending: the top-K' candidates
  unsigned a_id[KCAP + GCAP];
  unsigned a_raw[KCAP];
  int loff[KCAP];        // start of candidate j's dt list
  int cur[KCAP];         // fill cursor of candidate j's dt list
  unsigned char pass[KCAP];
  int rhist[256];        // radix-select digit histogram
  int wsum[NW];
  unsigned long long kw[NW];
  unsigned kid[NW];
  unsigned nhits, ndist, nabove, ngather, nrec;
  int dmin, dmax;
  int sel_digit, sel_need, sel_m;
  int segoff[514];       // record range of every id segment (nids < 2^24 -> <= 512 segments)
  int segcur[512];
};

// 96-bit composite key (weight bits, id), 12 digits of 8 bits from the top
__device__ __forceinline__ unsigned key_digit(unsigned long long w, unsigned id, int p) {
  return p < 8 ? (unsigned)(w >> (56 - 8 * p)) & 0xffu : (id >> (24 - 8 * (p - 8))) & 0xffu;
}
// compare the top `nfix` digits of (w,id) with those of the prefix: -1 below, 0 equal, +1 above
__device__ __forceinline__ int prefix_cmp(unsigned long long w, unsigned id, unsigned long long pw, unsigned pid,
                                          int nfix) {
  if (nfix == 0) return 0;
  if (nfix <= 8) {
    const int sh = 64 - 8 * nfix;
    const unsigned long long a = w >> sh, b = pw >> sh;
    return a > b ? 1 : (a < b ? -1 : 0);
  }
  if (w != pw) return w > pw ? 1 : -1;
  const int sh = 32 - 8 * (nfix - 8);
  const unsigned a = sh ? id >> sh : id, b = sh ? pid >> sh : pid;
  return a > b ? 1 : (a < b ? -1 : 0);
}

// inclusive scan of one int per thread over the CTA (MT threads)
__device__ __forceinline__ int block_scan_incl(int v, int* wsum) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o) v += t;
  }
  __syncthreads();
  if (lane == 31) wsum[warp] = v;
  __syncthreads();
  if (warp == 0) {
    int w = wsum[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += t;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  return v + (warp ? wsum[warp - 1] : 0);
}

// (value desc, index asc) arg-max over f[lo..hi] == np.argmax (first max)
__device__ inline void block_argmax(const int32_t* f, int lo, int hi, ModeScratch& sh, int& best_v, int& best_i) {
  const int tid = threadIdx.x;
  int v = -1, ix = 0x7fffffff;
  for (int i = lo + tid; i <= hi; i += MT) {
    const int x = f[i];
    if (x > v) { v = x; ix = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, ix, o);
    if (ov > v || (ov == v && oi < ix)) { v = ov; ix = oi; }
  }
  __syncthreads();
  if ((tid & 31) == 0) { sh.val[tid >> 5] = v; sh.idx[tid >> 5] = ix; }
  __syncthreads();
  best_v = sh.val[0];
  best_i = sh.idx[0];
  for (int w = 1; w < NW; ++w)
    if (sh.val[w] > best_v || (sh.val[w] == best_v && sh.idx[w] < best_i)) { best_v = sh.val[w]; best_i = sh.idx[w]; }
}

// Histogram-mode search of one candidate (audfprint_match.py:284-311) given lo/hi of its
// (already filled) dense histogram; emits rows, restores hist to zero.
__device__ inline void candidate_modes(const MatchArgs& a, ModeScratch& sh, int32_t* hist, int32_t* filt, int lo, int hi,
                                unsigned id, int raw, int rank, int32_t* qrows) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // keep_local_maxes (:70-75, locmax :51-67); zero-extended ends are equivalent
  for (int i = lo + tid; i <= hi; i += MT) {
    const int v = __ldcg(hist + i), l = __ldcg(hist + i - 1), r = __ldcg(hist + i + 1);
    filt[i] = (v >= l && r < v) ? v : 0;
  }
  __syncthreads();
  int found = 0;
  while (true) {
    int bv, bi;
    block_argmax(filt, lo, hi, sh, bv, bi);   // :290 np.argmax = first max
    if (bv <= a.thresh) break;                // :291
    // :295 count over +-window (hist is zero outside the touched range)
    int part = 0;
    for (int t2 = tid; t2 <= 2 * a.window; t2 += MT) part += __ldcg(hist + bi - a.window + t2);
    part = __reduce_add_sync(0xffffffffu, part);
    __syncthreads();
    if (lane == 0) sh.val[warp] = part;
    __syncthreads();
    if (tid == 0) {
      int count = 0;
      for (int w = 0; w < NW; ++w) count += sh.val[w];
      const int nr = sh.nrows;
      if (nr < a.row_cap) {
        int32_t* row = qrows + (size_t)nr * 7;
        row[0] = (int32_t)id; row[1] = count; row[2] = bi - a.bias; row[3] = raw;
        row[4] = rank; row[5] = 0; row[6] = 0;                      // :300-301
      }
      sh.nrows = nr + 1;
    }
    for (int t2 = tid; t2 <= 2 * a.window; t2 += MT) {              // :307-308
      const int i = bi - a.window + t2;
      if (i >= lo && i <= hi) filt[i] = 0;
    }
    __syncthreads();
    ++found;
    if (found > a.maxalign) break;                                   // :309-311
  }
  __syncthreads();
  for (int i = lo + tid; i <= hi; i += MT) hist[i] = 0;              // restore the scratch
  __syncthreads();
}

}  // namespace
