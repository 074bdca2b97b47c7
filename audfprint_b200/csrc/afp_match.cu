// K4 — device-resident hash table: bucket probe, per-track raw counts,
// weighted candidate ranking and per-candidate time-offset histogram modes.
//
// Replaces HashTable.get_hits (hash_table.py:150-176),
// Matcher._best_count_ids (audfprint_match.py:124-147) and
// Matcher._approx_match_counts (:241-312, find_time_range off).
//
// One CTA per query (persistent over the batch).  Each CTA owns a private
// scratch region in HBM (L2-resident in practice): the hit list, a dense
// per-track counter array that is cleared by replaying the list of distinct ids
// (never memset), and a dense dtime histogram cleared over the touched range.
// The results do not depend on hit order, so hits are appended with atomics.
//
// Tie rule (documented deviation, see oracle/afp_oracle.py::rank_candidates):
// the reference reverses an unstable argsort, so the order of equal weighted
// counts is implementation-defined there; here it is (weight desc, id desc).
#include <math.h>
#include "afp_internal.cuh"

namespace {

constexpr int MT = 256;   // threads per matching CTA

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
  uint32_t* counters;                     // nids
  int32_t* hist;    int hist_len;         // dtime histogram
  int32_t* filt;                          // local-max filtered copy
  int bias;
  int32_t* rows;    int row_cap;          // [nqueries][row_cap][7]
  int32_t* row_cnt;                       // [nqueries] rows produced (may exceed row_cap)
};

struct Key {
  unsigned long long w;   // bit pattern of the (positive) weighted count
  unsigned id;
  bool valid;
};
__device__ __forceinline__ bool key_less(const Key& a, const Key& b) {   // a < b
  if (!a.valid) return b.valid;
  if (!b.valid) return false;
  return a.w < b.w || (a.w == b.w && a.id < b.id);
}

__device__ Key block_max_key(Key k, Key* s_keys) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    Key other;
    other.w = __shfl_xor_sync(0xffffffffu, k.w, o);
    other.id = __shfl_xor_sync(0xffffffffu, k.id, o);
    other.valid = __shfl_xor_sync(0xffffffffu, (int)k.valid, o) != 0;
    if (key_less(k, other)) k = other;
  }
  __syncthreads();
  if ((tid & 31) == 0) s_keys[tid >> 5] = k;
  __syncthreads();
  Key best = s_keys[0];
  for (int w = 1; w < MT / 32; ++w)
    if (key_less(best, s_keys[w])) best = s_keys[w];
  return best;
}

// (value desc, index asc) arg-max over the filtered histogram == np.argmax (first max)
__device__ void block_argmax(const int32_t* f, int lo, int hi, int* s_val, int* s_idx, int& best_v,
                             int& best_i) {
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
  if ((tid & 31) == 0) { s_val[tid >> 5] = v; s_idx[tid >> 5] = ix; }
  __syncthreads();
  best_v = s_val[0];
  best_i = s_idx[0];
  for (int w = 1; w < MT / 32; ++w)
    if (s_val[w] > best_v || (s_val[w] == best_v && s_idx[w] < best_i)) { best_v = s_val[w]; best_i = s_idx[w]; }
}

__global__ void __launch_bounds__(MT) afp_match_kernel(MatchArgs a) {
  __shared__ Key s_keys[MT / 32];
  __shared__ int s_val[MT / 32], s_idx[MT / 32];
  __shared__ unsigned s_nhits, s_ndist, s_nabove;
  __shared__ int s_dmin, s_dmax, s_nrows;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint2* hits = a.hits + (size_t)blockIdx.x * a.hits_cap;
  uint32_t* dlist = a.dlist + (size_t)blockIdx.x * a.hits_cap;
  double* wtd = a.wtd + (size_t)blockIdx.x * a.hits_cap;
  uint32_t* cnt = a.counters + (size_t)blockIdx.x * a.nids;
  int32_t* hist = a.hist + (size_t)blockIdx.x * a.hist_len;
  int32_t* filt = a.filt + (size_t)blockIdx.x * a.hist_len;
  const uint32_t hmask = (1u << a.hashbits) - 1u, tmask = (1u << a.mtb) - 1u;

  for (int qi = blockIdx.x; qi < a.nqueries; qi += gridDim.x) {
    const int64_t q0 = a.qoff[qi];
    const int nq = (int)(a.qoff[qi + 1] - q0);
    if (tid == 0) { s_nhits = 0; s_ndist = 0; s_nabove = 0; s_nrows = 0; }
    __syncthreads();
    // ---- probe (hash_table.py:162-173): one warp per query row, lanes over slots
    for (int r = warp; r < nq; r += MT / 32) {
      const int qt = a.q[2 * (q0 + r)];
      const uint32_t b = (uint32_t)a.q[2 * (q0 + r) + 1] & hmask;
      const int n = min(a.depth, a.counts[b]);
      unsigned basepos = 0;
      if (lane == 0 && n > 0) basepos = atomicAdd(&s_nhits, (unsigned)n);
      basepos = __shfl_sync(0xffffffffu, basepos, 0);
      const uint32_t* row = a.table + (size_t)b * a.depth;
      for (int s = lane; s < n; s += 32) {
        const uint32_t v = row[s];
        const uint32_t id = (v >> a.mtb) - 1u;
        const int dt = (int)(v & tmask) - qt;
        hits[basepos + s] = make_uint2(id, (unsigned)(dt + a.bias));
        if (id < (uint32_t)a.nids) {
          const unsigned old = atomicAdd(&cnt[id], 1u);
          if (old == 0) dlist[atomicAdd(&s_ndist, 1u)] = id;
        }
      }
    }
    __syncthreads();
    const int nhits = (int)s_nhits, ndist = (int)s_ndist;
    // ---- weighted counts, number of ids above threshold (audfprint_match.py:132-144)
    {
      unsigned above = 0;
      for (int i = tid; i < ndist; i += MT) {
        const uint32_t id = dlist[i];
        const uint32_t raw = cnt[id];
        wtd[i] = (double)raw / (double)a.hpi[id];
        above += raw > (uint32_t)a.thresh ? 1u : 0u;
      }
      above = __reduce_add_sync(0xffffffffu, above);
      if (lane == 0 && above) atomicAdd(&s_nabove, above);
    }
    __syncthreads();
    const int maxdepth = min((int)s_nabove, a.sdepth);
    int32_t* qrows = a.rows + (size_t)qi * a.row_cap * 7;
    Key prev;
    prev.valid = false; prev.w = ~0ull; prev.id = ~0u;
    for (int rank = 0; rank < maxdepth; ++rank) {
      // ---- next candidate by (weight desc, id desc)
      Key best;
      best.valid = false; best.w = 0; best.id = 0;
      for (int i = tid; i < ndist; i += MT) {
        Key k;
        k.w = (unsigned long long)__double_as_longlong(wtd[i]);
        k.id = dlist[i];
        k.valid = true;
        const bool below_prev = !prev.valid || (k.w < prev.w || (k.w == prev.w && k.id < prev.id));
        if (below_prev && key_less(best, k)) best = k;
      }
      best = block_max_key(best, s_keys);
      if (!best.valid) break;
      prev = best;
      const uint32_t id = best.id;
      const int raw = (int)cnt[id];
      // ---- dtime histogram of this id (audfprint_match.py:284)
      if (tid == 0) { s_dmin = 0x7fffffff; s_dmax = -1; }
      __syncthreads();
      {
        int dmin = 0x7fffffff, dmax = -1;
        for (int i = tid; i < nhits; i += MT) {
          const uint2 h = hits[i];
          if (h.x == id) {
            atomicAdd(&hist[h.y], 1);
            dmin = min(dmin, (int)h.y);
            dmax = max(dmax, (int)h.y);
          }
        }
        dmin = __reduce_min_sync(0xffffffffu, dmin);
        dmax = __reduce_max_sync(0xffffffffu, dmax);
        if (lane == 0 && dmax >= 0) { atomicMin(&s_dmin, dmin); atomicMax(&s_dmax, dmax); }
      }
      __syncthreads();
      const int lo = s_dmin, hi = s_dmax;
      // keep_local_maxes (:70-75, locmax :51-67); zero-extended ends are equivalent
      for (int i = lo + tid; i <= hi; i += MT) {
        const int v = hist[i], l = hist[i - 1], r = hist[i + 1];
        filt[i] = (v >= l && r < v) ? v : 0;
      }
      __syncthreads();
      int found = 0;
      while (true) {
        int bv, bi;
        block_argmax(filt, lo, hi, s_val, s_idx, bv, bi);   // :290 np.argmax = first max
        if (bv <= a.thresh) break;                          // :291
        // :295 count over +-window (hist is zero outside the touched range)
        int part = 0;
        if (tid <= 2 * a.window) part = hist[bi - a.window + tid];
        for (int t2 = tid + MT; t2 <= 2 * a.window; t2 += MT) part += hist[bi - a.window + t2];
        part = __reduce_add_sync(0xffffffffu, part);
        __syncthreads();
        if (lane == 0) s_val[warp] = part;
        __syncthreads();
        if (tid == 0) {
          int count = 0;
          for (int w = 0; w < MT / 32; ++w) count += s_val[w];
          const int nr = s_nrows;
          if (nr < a.row_cap) {
            int32_t* row = qrows + (size_t)nr * 7;
            row[0] = (int32_t)id; row[1] = count; row[2] = bi - a.bias; row[3] = raw;
            row[4] = rank; row[5] = 0; row[6] = 0;                      // :300-301
          }
          s_nrows = nr + 1;
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
    // ---- restore the counter array by replaying the distinct ids
    __syncthreads();
    for (int i = tid; i < ndist; i += MT) cnt[dlist[i]] = 0;
    if (tid == 0) a.row_cnt[qi] = s_nrows;
    __syncthreads();
  }
}

// ---- get_hits (hash_table.py:150-176) -----------------------------------------
__global__ void afp_hit_count_kernel(const int32_t* q, int64_t nq, const int32_t* counts, int hashbits,
                                     int depth, int32_t* n_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const uint32_t b = (uint32_t)q[2 * i + 1] & ((1u << hashbits) - 1u);
  n_out[i] = min(depth, counts[b]);
}

__global__ void afp_hit_write_kernel(const int32_t* q, int64_t nq, const uint32_t* table, const int32_t* counts,
                                     int hashbits, int depth, int mtb, const int64_t* off, int32_t* hits) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  if (r >= nq) return;
  const int lane = threadIdx.x & 31;
  const int qt = q[2 * r];
  const uint32_t b = (uint32_t)q[2 * r + 1] & ((1u << hashbits) - 1u);
  const int n = min(depth, counts[b]);
  const uint32_t tmask = (1u << mtb) - 1u;
  const uint32_t* row = table + (size_t)b * depth;
  int32_t* out = hits + 4 * off[r];
  for (int s = lane; s < n; s += 32) {
    const uint32_t v = row[s];
    out[4 * s + 0] = (int32_t)((v >> mtb) - 1u);
    out[4 * s + 1] = (int32_t)(v & tmask) - qt;
    out[4 * s + 2] = (int32_t)b;
    out[4 * s + 3] = qt;
  }
}

__global__ void afp_restrict_ids_kernel(uint32_t* table, int32_t* counts, int64_t nbuckets, int depth, int mtb,
                                        uint32_t id_lo, uint32_t id_hi) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= nbuckets) return;
  uint32_t* row = table + (size_t)b * depth;
  const int n = min(depth, counts[b]);
  int k = 0;
  for (int s = 0; s < n; ++s) {
    const uint32_t v = row[s];
    const uint32_t id = (v >> mtb) - 1u;
    if (id >= id_lo && id < id_hi) row[k++] = v;
  }
  for (int s = k; s < n; ++s) row[s] = 0;
  counts[b] = k;
}

__global__ void afp_qmax_kernel(const int32_t* q, int64_t nq, int* out_max, int* out_min) {
  int mx = 0, mn = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += (int64_t)gridDim.x * blockDim.x) {
    mx = max(mx, q[2 * i]);
    mn = min(mn, q[2 * i]);
  }
  mx = __reduce_max_sync(0xffffffffu, mx);
  mn = __reduce_min_sync(0xffffffffu, mn);
  if ((threadIdx.x & 31) == 0) { atomicMax(out_max, mx); atomicMin(out_min, mn); }
}

__global__ void afp_pack_rows_kernel(const int32_t* rows, const int32_t* row_cnt, const int64_t* off, int row_cap,
                                     int32_t* packed) {
  const int qi = blockIdx.x;
  const int n = min(row_cnt[qi], row_cap);
  const int32_t* src = rows + (size_t)qi * row_cap * 7;
  int32_t* dst = packed + 7 * off[qi];
  for (int i = threadIdx.x; i < n * 7; i += blockDim.x) dst[i] = src[i];
}

__global__ void afp_clamp_kernel(const int32_t* in, int n, int cap, int32_t* out, int* overflow) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (in[i] > cap) atomicExch(overflow, 1);
  out[i] = min(in[i], cap);
}

}  // namespace

extern "C" {

int afp_table_upload(afp_ctx* c, const uint32_t* table, const int32_t* counts, int32_t hashbits, int32_t depth,
                     int32_t maxtimebits, const uint32_t* hashesperid, int64_t nids, int on_host) {
  if (!c || !table || !counts || (nids > 0 && !hashesperid)) return AFP_ERR_INVALID;
  if (hashbits < 1 || hashbits > 28 || depth < 1 || maxtimebits < 1 || maxtimebits > 24 || nids < 0)
    AFP_FAIL(c, AFP_ERR_INVALID, "bad table geometry");
  AFP_CUDA(c, cudaSetDevice(c->device));
  const size_t nb = (size_t)1 << hashbits;
  const cudaMemcpyKind kind = on_host ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
  c->tab.loaded = false;
  AFP_CUDA(c, c->tab.table.reserve(nb * (size_t)depth * sizeof(uint32_t)));
  AFP_CUDA(c, c->tab.counts.reserve(nb * sizeof(int32_t)));
  AFP_CUDA(c, c->tab.hashesperid.reserve(((size_t)nids + 1) * sizeof(uint32_t)));
  AFP_CUDA(c, cudaMemcpyAsync(c->tab.table.p, table, nb * (size_t)depth * sizeof(uint32_t), kind, c->stream));
  AFP_CUDA(c, cudaMemcpyAsync(c->tab.counts.p, counts, nb * sizeof(int32_t), kind, c->stream));
  if (nids > 0)
    AFP_CUDA(c, cudaMemcpyAsync(c->tab.hashesperid.p, hashesperid, (size_t)nids * sizeof(uint32_t), kind,
                                c->stream));
  AFP_CUDA(c, cudaStreamSynchronize(c->stream));
  c->tab.hashbits = hashbits;
  c->tab.depth = depth;
  c->tab.maxtimebits = maxtimebits;
  c->tab.nids = nids;
  c->tab.loaded = true;
  return AFP_OK;
}

int afp_table_restrict_ids(afp_ctx* c, int64_t id_lo, int64_t id_hi) {
  if (!c) return AFP_ERR_INVALID;
  if (!c->tab.loaded) AFP_FAIL(c, AFP_ERR_STATE, "no table uploaded");
  if (id_lo < 0 || id_hi < id_lo) AFP_FAIL(c, AFP_ERR_INVALID, "bad id range");
  AFP_CUDA(c, cudaSetDevice(c->device));
  const int64_t nb = (int64_t)1 << c->tab.hashbits;
  afp_restrict_ids_kernel<<<(unsigned)((nb + 255) / 256), 256, 0, c->stream>>>(
      c->tab.table.as<uint32_t>(), c->tab.counts.as<int32_t>(), nb, c->tab.depth, c->tab.maxtimebits,
      (uint32_t)id_lo, (uint32_t)id_hi);
  AFP_CUDA(c, cudaGetLastError());
  c->launches++;
  return AFP_OK;
}

static int stage_queries(afp_ctx* c, const int32_t* q, int64_t nrows, int on_host, const int32_t** dq) {
  *dq = q;
  if (on_host && nrows > 0) {
    AFP_CUDA(c, c->d_q.reserve(sizeof(int32_t) * 2 * (size_t)nrows));
    AFP_CUDA(c, cudaMemcpyAsync(c->d_q.p, q, sizeof(int32_t) * 2 * (size_t)nrows, cudaMemcpyHostToDevice,
                                c->stream));
    *dq = c->d_q.as<int32_t>();
  }
  return AFP_OK;
}

int afp_get_hits(afp_ctx* c, const int32_t* q_rows, int64_t nq, int q_on_host, int64_t* nhits) {
  if (!c || nq < 0 || (nq > 0 && !q_rows)) return AFP_ERR_INVALID;
  if (!c->tab.loaded) AFP_FAIL(c, AFP_ERR_STATE, "no table uploaded");
  AFP_CUDA(c, cudaSetDevice(c->device));
  c->nhits = -1;
  const int32_t* dq = nullptr;
  int rc = stage_queries(c, q_rows, nq, q_on_host, &dq);
  if (rc) return rc;
  AFP_CUDA(c, c->d_tmp.reserve(sizeof(int32_t) * (size_t)(nq + 1)));
  AFP_CUDA(c, c->d_hit_off.reserve(sizeof(int64_t) * (size_t)(nq + 1)));
  if (nq > 0) {
    afp_hit_count_kernel<<<(unsigned)((nq + 255) / 256), 256, 0, c->stream>>>(
        dq, nq, c->tab.counts.as<int32_t>(), c->tab.hashbits, c->tab.depth, c->d_tmp.as<int32_t>());
    AFP_CUDA(c, cudaGetLastError());
    c->launches++;
  }
  if ((rc = afp_launch_scan_i32_to_i64(c, c->d_tmp.as<int32_t>(), c->d_hit_off.as<int64_t>(), nq))) return rc;
  int64_t total = 0;
  AFP_CUDA(c, cudaMemcpyAsync(&total, c->d_hit_off.as<int64_t>() + nq, sizeof(int64_t), cudaMemcpyDeviceToHost,
                              c->stream));
  AFP_CUDA(c, cudaStreamSynchronize(c->stream));
  AFP_CUDA(c, c->d_hits.reserve(sizeof(int32_t) * 4 * (size_t)(total + 1)));
  if (total > 0) {
    afp_hit_write_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, c->stream>>>(
        dq, nq, c->tab.table.as<uint32_t>(), c->tab.counts.as<int32_t>(), c->tab.hashbits, c->tab.depth,
        c->tab.maxtimebits, c->d_hit_off.as<int64_t>(), c->d_hits.as<int32_t>());
    AFP_CUDA(c, cudaGetLastError());
    c->launches++;
  }
  c->nhits = total;
  c->hits_nq = nq;
  if (nhits) *nhits = total;
  return AFP_OK;
}

int afp_fetch_hits(afp_ctx* c, int32_t* hits, int hits_on_host) {
  if (!c) return AFP_ERR_INVALID;
  if (c->nhits < 0) AFP_FAIL(c, AFP_ERR_STATE, "afp_get_hits has not been called");
  AFP_CUDA(c, cudaSetDevice(c->device));
  if (hits && c->nhits > 0)
    AFP_CUDA(c, cudaMemcpyAsync(hits, c->d_hits.p, sizeof(int32_t) * 4 * (size_t)c->nhits,
                                hits_on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
  AFP_CUDA(c, cudaStreamSynchronize(c->stream));
  return AFP_OK;
}

int afp_match_batch(afp_ctx* c, const int32_t* q_rows, int q_on_host, int32_t nqueries, const int64_t* q_offsets,
                    const afp_matcher_params* p, int64_t* total_rows) {
  if (!c || !p || nqueries < 0 || (nqueries > 0 && !q_offsets)) return AFP_ERR_INVALID;
  if (!c->tab.loaded) AFP_FAIL(c, AFP_ERR_STATE, "no table uploaded");
  if (p->window < 0 || p->window > 4096 || p->search_depth < 0 || p->max_alignments_per_id < 0)
    AFP_FAIL(c, AFP_ERR_INVALID, "bad matcher parameters");
  AFP_CUDA(c, cudaSetDevice(c->device));
  c->match_total_rows = -1;
  c->match_nq = nqueries;
  int64_t maxnq = 0;
  for (int i = 0; i < nqueries; ++i) {
    if (q_offsets[i + 1] < q_offsets[i]) AFP_FAIL(c, AFP_ERR_INVALID, "q_offsets must be non-decreasing");
    maxnq = std::max<int64_t>(maxnq, q_offsets[i + 1] - q_offsets[i]);
  }
  const int64_t nrows_in = nqueries ? q_offsets[nqueries] : 0;
  if (nqueries && q_offsets[0] != 0) AFP_FAIL(c, AFP_ERR_INVALID, "q_offsets[0] must be 0");
  if (nrows_in > 0 && !q_rows) return AFP_ERR_INVALID;
  const int32_t* dq = nullptr;
  int rc = stage_queries(c, q_rows, nrows_in, q_on_host, &dq);
  if (rc) return rc;
  AFP_CUDA(c, c->d_qoff.reserve(sizeof(int64_t) * (size_t)(nqueries + 1)));
  AFP_CUDA(c, c->d_mrow_cnt.reserve(sizeof(int32_t) * (size_t)(nqueries + 2)));
  AFP_CUDA(c, c->d_mrow_off.reserve(sizeof(int64_t) * (size_t)(nqueries + 1)));
  if (nqueries == 0) {
    c->match_total_rows = 0;
    if (total_rows) *total_rows = 0;
    return AFP_OK;
  }
  AFP_CUDA(c, cudaMemcpyAsync(c->d_qoff.p, q_offsets, sizeof(int64_t) * (size_t)(nqueries + 1),
                              cudaMemcpyHostToDevice, c->stream));
  // largest / smallest query time (sizes the dtime histogram)
  int h_mm[2] = {0, 0};
  AFP_CUDA(c, c->d_tmp.reserve(sizeof(int32_t) * (size_t)(nqueries + 8)));
  int* d_mm = c->d_tmp.as<int>();
  AFP_CUDA(c, cudaMemsetAsync(d_mm, 0, 2 * sizeof(int), c->stream));
  if (nrows_in > 0) {
    afp_qmax_kernel<<<296, 256, 0, c->stream>>>(dq, nrows_in, d_mm, d_mm + 1);
    AFP_CUDA(c, cudaGetLastError());
    c->launches++;
  }
  AFP_CUDA(c, cudaMemcpyAsync(h_mm, d_mm, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  AFP_CUDA(c, cudaStreamSynchronize(c->stream));
  if (h_mm[1] < 0) AFP_FAIL(c, AFP_ERR_INVALID, "negative query time");

  cudaDeviceProp prop;
  AFP_CUDA(c, cudaGetDeviceProperties(&prop, c->device));
  const int nctas = (int)std::min<int64_t>(nqueries, (int64_t)prop.multiProcessorCount * 2);
  MatchArgs a;
  a.q = dq;
  a.qoff = c->d_qoff.as<int64_t>();
  a.nqueries = nqueries;
  a.table = c->tab.table.as<uint32_t>();
  a.counts = c->tab.counts.as<int32_t>();
  a.hpi = c->tab.hashesperid.as<uint32_t>();
  a.hashbits = c->tab.hashbits;
  a.depth = c->tab.depth;
  a.mtb = c->tab.maxtimebits;
  a.nids = std::max<int64_t>(c->tab.nids, 1);
  a.window = p->window;
  a.thresh = p->threshcount;
  a.sdepth = p->search_depth;
  a.maxalign = p->max_alignments_per_id;
  a.hits_cap = std::max<int64_t>(maxnq * c->tab.depth, 1);
  a.bias = h_mm[0] + p->window + 2;
  a.hist_len = (1 << c->tab.maxtimebits) + a.bias + p->window + 4;
  a.row_cap = 256;
  const char* env = getenv("AFP_MATCH_ROW_CAP");
  if (env && atoi(env) > 0) a.row_cap = atoi(env);
  const size_t per_cta = (size_t)a.hits_cap * (sizeof(uint2) + sizeof(uint32_t) + sizeof(double)) +
                         (size_t)a.nids * sizeof(uint32_t) + (size_t)a.hist_len * 2 * sizeof(int32_t) + 64;
  const size_t before = c->d_mscratch.cap;
  AFP_CUDA(c, c->d_mscratch.reserve(per_cta * (size_t)nctas + 1024));
  AFP_CUDA(c, c->d_mrows.reserve(sizeof(int32_t) * 7 * (size_t)a.row_cap * (size_t)nqueries));
  // carve the scratch; counters and histograms must start (and are left) zeroed
  char* base = c->d_mscratch.as<char>();
  auto carve = [&](size_t bytes) {
    char* p0 = base;
    base += (bytes + 15) & ~(size_t)15;
    return p0;
  };
  a.hits = (uint2*)carve(sizeof(uint2) * a.hits_cap * nctas);
  a.wtd = (double*)carve(sizeof(double) * a.hits_cap * nctas);
  a.dlist = (uint32_t*)carve(sizeof(uint32_t) * a.hits_cap * nctas);
  char* zero0 = base;
  a.counters = (uint32_t*)carve(sizeof(uint32_t) * (size_t)a.nids * nctas);
  a.hist = (int32_t*)carve(sizeof(int32_t) * (size_t)a.hist_len * nctas);
  a.filt = (int32_t*)carve(sizeof(int32_t) * (size_t)a.hist_len * nctas);
  (void)before;
  AFP_CUDA(c, cudaMemsetAsync(zero0, 0, (size_t)(base - zero0), c->stream));
  a.rows = c->d_mrows.as<int32_t>();
  a.row_cnt = c->d_mrow_cnt.as<int32_t>();
  c->match_row_cap = a.row_cap;
  afp_match_kernel<<<nctas, MT, 0, c->stream>>>(a);
  AFP_CUDA(c, cudaGetLastError());
  c->launches++;
  // clamp counts to the capacity (flagging overflow), scan, pack
  int* d_over = c->d_mrow_cnt.as<int>() + nqueries + 1;
  AFP_CUDA(c, cudaMemsetAsync(d_over, 0, sizeof(int), c->stream));
  AFP_CUDA(c, c->d_tmp.reserve(sizeof(int32_t) * (size_t)(nqueries + 8)));
  afp_clamp_kernel<<<(nqueries + 255) / 256, 256, 0, c->stream>>>(a.row_cnt, nqueries, a.row_cap,
                                                                  c->d_tmp.as<int32_t>(), d_over);
  AFP_CUDA(c, cudaGetLastError());
  c->launches++;
  if ((rc = afp_launch_scan_i32_to_i64(c, c->d_tmp.as<int32_t>(), c->d_mrow_off.as<int64_t>(), nqueries))) return rc;
  int64_t total = 0;
  int over = 0;
  AFP_CUDA(c, cudaMemcpyAsync(&total, c->d_mrow_off.as<int64_t>() + nqueries, sizeof(int64_t),
                              cudaMemcpyDeviceToHost, c->stream));
  AFP_CUDA(c, cudaMemcpyAsync(&over, d_over, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  AFP_CUDA(c, cudaStreamSynchronize(c->stream));
  if (over) AFP_FAIL(c, AFP_ERR_UNSUPPORTED, "a query produced more rows than AFP_MATCH_ROW_CAP (default 256)");
  AFP_CUDA(c, c->d_mrows_packed.reserve(sizeof(int32_t) * 7 * (size_t)(total + 1)));
  if (total > 0) {
    afp_pack_rows_kernel<<<nqueries, 64, 0, c->stream>>>(a.rows, a.row_cnt, c->d_mrow_off.as<int64_t>(), a.row_cap,
                                                         c->d_mrows_packed.as<int32_t>());
    AFP_CUDA(c, cudaGetLastError());
    c->launches++;
  }
  c->match_total_rows = total;
  if (total_rows) *total_rows = total;
  return AFP_OK;
}

int afp_fetch_match_rows(afp_ctx* c, int32_t* rows, int rows_on_host, int64_t* row_offsets) {
  if (!c) return AFP_ERR_INVALID;
  if (c->match_total_rows < 0) AFP_FAIL(c, AFP_ERR_STATE, "afp_match_batch has not been called");
  AFP_CUDA(c, cudaSetDevice(c->device));
  if (row_offsets) {
    if (c->match_nq == 0) row_offsets[0] = 0;
    else
      AFP_CUDA(c, cudaMemcpyAsync(row_offsets, c->d_mrow_off.p, sizeof(int64_t) * (size_t)(c->match_nq + 1),
                                  cudaMemcpyDeviceToHost, c->stream));
  }
  if (rows && c->match_total_rows > 0)
    AFP_CUDA(c, cudaMemcpyAsync(rows, c->d_mrows_packed.p, sizeof(int32_t) * 7 * (size_t)c->match_total_rows,
                                rows_on_host ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, c->stream));
  AFP_CUDA(c, cudaStreamSynchronize(c->stream));
  return AFP_OK;
}

}  // extern "C"
