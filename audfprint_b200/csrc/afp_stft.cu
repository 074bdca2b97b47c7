// K1 — fused framing + Hann window + 512-point real FFT + log-magnitude (FP64).
//
// Replaces stft.stft (reference stft.py:62-94: reflect pad 256, hop-256
// framing, window multiply, rfft) and the |.| / log part of
// Analyzer.find_peaks (audfprint_analyze.py:280-285).  The two whole-file
// reductions that follow (floor = max/1e6, mean, :283-286) are produced as
// per-tile partials here and finished in afp_stats_kernel; the per-bin high-pass
// (:293-295) is a time recursion and lives in the peak kernel (afp_peaks.cu).
//
// Work decomposition: a tile = 16 consecutive frames of one item (file x shift).
// PERSISTENT CTAs (2 per SM) walk the tile list; the hop-strided PCM of a tile
// is ONE contiguous run of 17*256 samples, staged into a double-buffered shared
// memory ring by 1-D TMA bulk copies (cp.async.bulk + mbarrier) issued one tile
// ahead, so the copy of tile i+1 overlaps the FFTs of tile i (edge tiles and
// unaligned files take a reflected scalar-load path).  16 threads cooperate on
// a frame: 256-point complex FFT as 16 x 16 with one padded shared-memory
// transpose, then the real-FFT split done on (k, 256-k) PAIRS so that each
// partner exchange (warp shuffle) and each W512 twiddle serves two bins.
//
// Why FP64: the peak decisions downstream compare these values bit-for-bit the
// way the reference's float64 NumPy path does; an FP32 spectrogram flips a
// decision roughly once per 10^5 frames (DESIGN.md "Precision").  The kernel is
// therefore bound by the FP64 pipe (64 lanes/clk/SM), not by HBM; the log is a
// table-driven FP64 routine (10 FP64 ops instead of libdevice's ~30).
#include <math.h>
#include <algorithm>
#include "afp_fft.cuh"
#include "afp_internal.cuh"

namespace {

constexpr int FT = AFP_FRAMES_PER_TILE;   // 16 frames per tile
constexpr int XS = 17;                    // padded row stride of the 16x16 exchange
constexpr int XF = 16 * XS;               // 272 doubles per frame per component
constexpr int K1_THREADS = 256;
constexpr int LOGTAB = 64;                // log table entries (6 mantissa bits) ...
constexpr int LOGCOPIES = 8;              // ... each stored 8 times, copy j in the 16-byte bank group j:
                                          // lane (l & 7) reads copy (l & 7), so the random lookups of a
                                          // quarter-warp never collide (4 wavefronts per LDS.128, always)

struct StftArgs {
  const void* pcm;
  const ItemDesc* items;
  const int32_t* tile_item;   // [ntiles] item of every tile (built by afp_tile_table_kernel)
  int nitems;
  int tile_begin, tile_end;   // tile range of this launch (a chunk of the batch)
  const double* window;   // 512 (pre-scaled by 2^-15 for int16 PCM)
  const double2* tw256;   // [p][r] = W256^(r*p), (cos, -sin)
  const double2* w512;    // 256: (cos, -sin)(2 pi k / 512)
  const double2* logtab;  // [64][8]: (c_i, -0.5*log(c_i)), 8 identical copies interleaved
  double* logs;           // [frames][256]
  double* nyq;            // [frames]
  double* tile_stats;     // [tiles][3]
  double* mag;            // optional [frames][257]
  // FP32 spectrogram mode (opt-in, NOT bit-identical downstream)
  const float* window_f;  // 512 (pre-scaled by 2^-15 for int16 PCM)
  const float2* tw256_f;  // [p][r]
  const float2* w512_f;   // 256
  float* logs_f;          // [frames][256]
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// np.pad(..., mode='reflect') index map (edge sample not repeated), any number
// of bounces (stft.py:88).
__device__ __forceinline__ int64_t reflect_index(int64_t j, int64_t n) {
  if (n == 1) return 0;
  const int64_t period = 2 * (n - 1);
  int64_t r = j % period;
  if (r < 0) r += period;
  return r < n ? r : period - r;
}

// exact small-int -> double without the (slow) I2F.F64 path: (1.5*2^52 + x) - 1.5*2^52
__device__ __forceinline__ double int_to_double(int x) {
  return __hiloint2double(0x43380000 + (x >> 31), x) - 6755399441055744.0;
}

// 0.5*log(0.25*v) for v > 0 normal; table-driven, |abs error| ~ 1e-16 + 0.5 ulp.
//   v = 2^e * m, m in [1,2); i = top 6 mantissa bits; r = m*c_i - 1, |r| <= 2^-7
//   0.5*log(v/4) = (e-2)*ln2/2 + t_i + 0.5*log1p(r),  t_i = -0.5*log(c_i)
// Inputs that are 0 / denormal / inf / nan give a meaningless value here; the
// caller detects them with is_special() and patches with half_log_quarter_slow().
__device__ __noinline__ double half_log_quarter_slow(double v) { return 0.5 * log(0.25 * v); }
__device__ __forceinline__ bool is_special(double v) {
  return (unsigned)(__double2hiint(v) - 0x00100000) >= 0x7fe00000u;
}

// s_logtab_lane = table base + (lane & 7): this lane's private copy (stride LOGCOPIES).
__device__ __forceinline__ double half_log_quarter(double v, const double2* s_logtab_lane) {
  const int hi = __double2hiint(v);
  const int lo = __double2loint(v);
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, lo);
  const double2 ct = s_logtab_lane[((hi >> 14) & (LOGTAB - 1)) * LOGCOPIES];
  const double r = fma(m, ct.x, -1.0);
  double p = 0.5 / 7.0;
  p = fma(p, r, -0.5 / 6.0);
  p = fma(p, r, 0.5 / 5.0);
  p = fma(p, r, -0.5 / 4.0);
  p = fma(p, r, 0.5 / 3.0);
  p = fma(p, r, -0.5 / 2.0);
  p = fma(p, r, 0.5);
  const double ed = int_to_double((hi >> 20) - 1025);
  return fma(ed, 0.34657359027997264, ct.y) + p * r;   // ln2/2
}

template <typename PcmT> struct PcmTraits;
template <> struct PcmTraits<int16_t> {
  static constexpr int NBUF = 2;
  __device__ static __forceinline__ void load2(const int16_t* p, double& a, double& b) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
    a = int_to_double((int)(short)(w & 0xffffu));
    b = int_to_double((int)w >> 16);
  }
};
template <> struct PcmTraits<float> {
  static constexpr int NBUF = 1;   // 2 x 17 KB would not leave room for two CTAs per SM
  __device__ static __forceinline__ void load2(const float* p, double& a, double& b) {
    const float2 w = *reinterpret_cast<const float2*>(p);
    a = (double)w.x;
    b = (double)w.y;
  }
};

struct TileInfo {
  int64_t frame0;     // batch-wide index of the tile's first frame
  int64_t src;        // absolute sample index of the run's first sample (sample_start + j0)
  int64_t j0;         // the same relative to the item (may be < 0)
  int64_t nsamples;   // samples of the item (reflection period)
  int nft;            // frames of the tile that exist
  int tma;            // (first << 16) | count: samples [first, first+count) of the run come by one bulk
                      // copy (count == 0: none); whatever else the run holds - the reflected head of a
                      // file's first tile, the tail of its last - is staged by scalar loads
};

template <typename PcmT>
__device__ __forceinline__ TileInfo make_tile(const StftArgs& a, const ItemDesc& it, int tile) {
  TileInfo ti;
  const int t0 = (tile - it.tile_base) * FT;
  ti.frame0 = it.frame_base + t0;
  ti.nft = min(FT, it.nframes - t0);
  ti.j0 = (int64_t)(t0 - 1) * AFP_N_HOP;   // first sample of the run (may be < 0)
  ti.src = it.sample_start + ti.j0;
  ti.nsamples = it.nsamples;
  // the part of the run that lies inside the file, trimmed to 16-byte granules; the shared-memory
  // buffer is 16-byte aligned, so source and destination agree iff the run itself starts on one
  constexpr int GR = 16 / (int)sizeof(PcmT);
  const int64_t run_n = (int64_t)(ti.nft + 1) * AFP_N_HOP;
  const int64_t lo = ti.j0 < 0 ? -ti.j0 : 0;
  const int64_t hi = (ti.j0 + run_n <= it.nsamples ? run_n : it.nsamples - ti.j0) & ~(int64_t)(GR - 1);
  const bool aligned = (reinterpret_cast<uintptr_t>(reinterpret_cast<const PcmT*>(a.pcm) + ti.src) & 15) == 0;
  ti.tma = (aligned && hi > lo) ? (int)((lo << 16) | (hi - lo)) : 0;   // lo is 0 or 256: a granule multiple
  return ti;
}

// Stage the PCM run of a tile: TMA bulk copy (one thread) or reflected loads (all).
template <typename PcmT>
__device__ __forceinline__ void stage_tile(const StftArgs& a, const TileInfo& ti, PcmT* dst,
                                           unsigned long long* bar) {
  const PcmT* pcm = reinterpret_cast<const PcmT*>(a.pcm);
  const int nsamp = (ti.nft + 1) * AFP_N_HOP;
  const int t_first = ti.tma >> 16, t_count = ti.tma & 0xffff;
  if (t_count) {
    if (threadIdx.x == 0) {
      const uint32_t bytes = t_count * sizeof(PcmT);
      // order earlier generic-proxy accesses of this buffer before the async-proxy write
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                   : "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32(dst + t_first)),
          "l"(pcm + ti.src + t_first), "r"(bytes), "r"(smem_u32(bar))
          : "memory");
    }
  }
  if (t_count != nsamp) {   // samples outside the bulk copy: [0, t_first) and [t_first + t_count, nsamp)
    const PcmT* item0 = pcm + (ti.src - ti.j0);
    for (int i = threadIdx.x; i < nsamp - t_count; i += K1_THREADS) {
      const int p = i < t_first ? i : i + t_count;
      dst[p] = item0[reflect_index(ti.j0 + p, ti.nsamples)];
    }
  }
}

__device__ __forceinline__ void wait_bar(unsigned long long* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

template <typename PcmT, bool WRITE_MAG>
__global__ void __launch_bounds__(K1_THREADS, 2) afp_stft_kernel(StftArgs a) {
  constexpr int NBUF = PcmTraits<PcmT>::NBUF;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_win = reinterpret_cast<double*>(smem_raw);                 // 512
  double2* s_tw256 = reinterpret_cast<double2*>(s_win + 512);          // 256, [p][r]
  double2* s_w512 = s_tw256 + 256;                                     // 256
  double2* s_logtab_all = s_w512 + 256;                                // LOGTAB * LOGCOPIES
  double* s_xr = reinterpret_cast<double*>(s_logtab_all + LOGTAB * LOGCOPIES);   // FT * XF
  double* s_xi = s_xr + FT * XF;                                       // FT * XF
  double* s_red = s_xi + FT * XF;                                      // 3 * 8
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_red + 24);   // 2
  PcmT* s_pcm = reinterpret_cast<PcmT*>(s_bar + 2);                    // NBUF * (FT+1)*256, 16 B aligned
  constexpr int PCM_BUF = (FT + 1) * AFP_N_HOP;

  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar + 1)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 512; i += K1_THREADS) s_win[i] = a.window[i];
  for (int i = tid; i < 256; i += K1_THREADS) {
    s_tw256[i] = a.tw256[i];
    s_w512[i] = a.w512[i];
  }
  for (int i = tid; i < LOGTAB * LOGCOPIES; i += K1_THREADS) s_logtab_all[i] = a.logtab[i];
  __syncthreads();

  const int g = tid >> 4;   // frame within the tile
  const int r = tid & 15;   // cooperating thread within the frame
  const int lane = tid & 31;
  const double2* s_logtab = s_logtab_all + (lane & (LOGCOPIES - 1));   // this lane's copy of the log table
  const int src_lane = (lane & 16) | ((16 - r) & 15);
  uint32_t phases = 0u;   // bit b = parity to wait for on barrier b

  int tile = a.tile_begin + blockIdx.x;
  if (tile >= a.tile_end) return;
  const int G = gridDim.x;
  // Tile descriptors are fetched two tiles ahead so that their (dependent) global
  // loads never sit on the critical path: item index at distance 3, ItemDesc at 2.
  TileInfo cur = make_tile<PcmT>(a, a.items[a.tile_item[tile]], tile);
  TileInfo nxt = cur;
  if (tile + G < a.tile_end) nxt = make_tile<PcmT>(a, a.items[a.tile_item[tile + G]], tile + G);
  int item_nn = (tile + 2 * G < a.tile_end) ? a.tile_item[tile + 2 * G] : 0;
  stage_tile<PcmT>(a, cur, s_pcm, s_bar);
  int buf = 0;

  for (; tile < a.tile_end; tile += G) {
    const int next = tile + G;
    // prefetch the next tile into the other buffer (free since the end-of-iteration barrier)
    if (NBUF == 2 && next < a.tile_end) stage_tile<PcmT>(a, nxt, s_pcm + (buf ^ 1) * PCM_BUF, s_bar + (buf ^ 1));
    ItemDesc desc_nn = a.items[item_nn];                                   // consumed at the end of the iteration
    const int item_n3 = (tile + 3 * G < a.tile_end) ? a.tile_item[tile + 3 * G] : 0;   // consumed next iteration
    if (cur.tma & 0xffff) {
      wait_bar(s_bar + buf, (phases >> buf) & 1u);
      phases ^= 1u << buf;
    }
    if ((cur.tma & 0xffff) != (cur.nft + 1) * AFP_N_HOP) __syncthreads();   // scalar-staged samples are visible

    const bool active = g < cur.nft;
    const int64_t frame = cur.frame0 + g;
    double vmax = 0.0, vsum = 0.0;
    int hmin = 0x7ff00000;   // min over the high words of 4|X|^2 (positive doubles order like their bits)
    double zr[16], zi[16];
    if (active) {
      // step A: z[16q + r] = (x[2n] w[2n], x[2n+1] w[2n+1]), n = 16q + r
      const PcmT* fr = s_pcm + buf * PCM_BUF + g * AFP_N_HOP;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int i0 = 2 * (16 * q + r);
        const double2 w = *reinterpret_cast<const double2*>(s_win + i0);
        double x0, x1;
        PcmTraits<PcmT>::load2(fr + i0, x0, x1);
        zr[q] = x0 * w.x;
        zi[q] = x1 * w.y;
      }
      afp_fft16(zr, zi);
      double* xr = s_xr + g * XF;
      double* xi = s_xi + g * XF;
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const double2 w = s_tw256[p * 16 + r];
        xr[p * XS + r] = zr[p] * w.x - zi[p] * w.y;
        xi[p * XS + r] = zr[p] * w.y + zi[p] * w.x;
      }
    }
    __syncwarp();
    if (active) {
      // step B: thread p = r transforms column p: Z[p + 16 s]
      const double* xr = s_xr + g * XF + r * XS;
      const double* xi = s_xi + g * XF + r * XS;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        zr[q] = xr[q];
        zi[q] = xi[q];
      }
      afp_fft16(zr, zi);
    }
    // real-FFT split on pairs (k, 256-k), k = r + 16 s, s < 8: the partner
    // Z[(256-k)&255] sits in lane (16-r)&15, register 15-s (r > 0) or (16-s)&15 (r == 0).
    //   2Xe = Z[k] + conj(Zp), 2Xo = -i (Z[k] - conj(Zp)), P = W512^k * 2Xo
    //   4|X[k]|^2 = |2Xe + P|^2,  4|X[256-k]|^2 = |2Xe - P|^2
    {
      double* out = a.logs + frame * AFP_NBINS;
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        double c = __shfl_sync(0xffffffffu, zr[15 - s], src_lane);
        double d = __shfl_sync(0xffffffffu, zi[15 - s], src_lane);
        if (r == 0) {
          c = zr[(16 - s) & 15];
          d = zi[(16 - s) & 15];
        }
        if (active) {
          const int k = r + 16 * s;
          const double2 w = s_w512[k];
          const double er = zr[s] + c, ei = zi[s] - d, orr = zi[s] + d, oi = c - zr[s];
          const double pr = w.x * orr - w.y * oi, pi = w.x * oi + w.y * orr;
          const double ar = er + pr, ai = ei + pi, br = er - pr, bi = ei - pi;
          const double ssa = ar * ar + ai * ai;   // 4 |X[k]|^2
          const double ssb = br * br + bi * bi;   // 4 |X[256-k]|^2
          double la = half_log_quarter(ssa, s_logtab);
          double lb = half_log_quarter(ssb, s_logtab);
          if (is_special(ssa) || is_special(ssb)) {   // digital silence etc.: rare, off the hot path
            la = half_log_quarter_slow(ssa);
            lb = half_log_quarter_slow(ssb);
          }
          out[k] = la;
          if (k != 0) out[256 - k] = lb; else a.nyq[frame] = lb;   // k == 0 pairs with the Nyquist bin
          if (WRITE_MAG) {
            a.mag[frame * 257 + k] = sqrt(0.25 * ssa);
            a.mag[frame * 257 + 256 - k] = sqrt(0.25 * ssb);
          }
          vmax = fmax(vmax, fmax(ssa, ssb));
          hmin = min(hmin, min(__double2hiint(ssa), __double2hiint(ssb)));
          vsum += la + lb;
        }
      }
      if (active && r == 0) {   // bin 128 pairs with itself
        const double2 w = s_w512[128];
        const double er = 2.0 * zr[8], orr = 2.0 * zi[8];   // ei = 0, oi = 0
        const double ar = er + w.x * orr, ai = w.y * orr;
        const double ss = ar * ar + ai * ai;
        double lg = half_log_quarter(ss, s_logtab);
        if (is_special(ss)) lg = half_log_quarter_slow(ss);
        out[128] = lg;
        if (WRITE_MAG) a.mag[frame * 257 + 128] = sqrt(0.25 * ss);
        vmax = fmax(vmax, ss);
        hmin = min(hmin, __double2hiint(ss));
        vsum += lg;
      }
    }
    // deterministic CTA reduction of (max |S|^2, min log, sum log)
    // a LOWER bound of the smallest log: the high word alone (low word zeroed).  The
    // statistics pass only asks "is anything below the floor?"; a false alarm just takes its
    // exact path.
    hmin = __reduce_min_sync(0xffffffffu, hmin);
    const double v_lo = __hiloint2double(hmin, 0);
    double vmin = hmin >= 0x7ff00000 ? INFINITY      // this warp had no frame in the tile
                  : (hmin < 0x00100000 ? -INFINITY : half_log_quarter(v_lo, s_logtab));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
      vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
    }
    if (lane == 0) {
      s_red[(tid >> 5) * 3 + 0] = vmax;
      s_red[(tid >> 5) * 3 + 1] = vmin;
      s_red[(tid >> 5) * 3 + 2] = vsum;
    }
    __syncthreads();   // also: every thread is done with s_pcm[buf]
    if (tid == 0) {
      double m = 0.0, mn = INFINITY, sm = 0.0;
#pragma unroll
      for (int w = 0; w < K1_THREADS / 32; ++w) {
        m = fmax(m, s_red[w * 3 + 0]);
        mn = fmin(mn, s_red[w * 3 + 1]);
        sm += s_red[w * 3 + 2];
      }
      a.tile_stats[(size_t)tile * 3 + 0] = 0.25 * m;
      a.tile_stats[(size_t)tile * 3 + 1] = mn;
      a.tile_stats[(size_t)tile * 3 + 2] = sm;
    }
    if (NBUF == 2) buf ^= 1;
    cur = nxt;
    if (tile + 2 * G < a.tile_end) nxt = make_tile<PcmT>(a, desc_nn, tile + 2 * G);
    item_nn = item_n3;
    if (NBUF == 1 && next < a.tile_end) stage_tile<PcmT>(a, cur, s_pcm, s_bar);
  }
}


// ---- FP32 variant of K1 (Analyzer.precision = 'fp32') -----------------------------------
// Same decomposition, single precision throughout (FFT, |.|^2, log via MUFU), float
// log-spectrogram out: half the output bytes, twice the FP32 lane rate, ~80 registers ->
// three CTAs per SM.  Downstream decisions then see values that differ from the reference's
// by ~1e-7 relative, so hashes are NOT guaranteed bit-identical (measured in bench.py).
template <typename PcmT> struct PcmF32;
template <> struct PcmF32<int16_t> {
  __device__ static __forceinline__ void load2(const int16_t* p, float& a, float& b) {
    const uint32_t w = *reinterpret_cast<const uint32_t*>(p);
    a = (float)(short)(w & 0xffffu);
    b = (float)((int)w >> 16);
  }
};
template <> struct PcmF32<float> {
  __device__ static __forceinline__ void load2(const float* p, float& a, float& b) {
    const float2 w = *reinterpret_cast<const float2*>(p);
    a = w.x;
    b = w.y;
  }
};

template <typename PcmT, bool WRITE_MAG>
__global__ void __launch_bounds__(K1_THREADS, 3) afp_stft_f32_kernel(StftArgs a) {
  constexpr int NBUF = PcmTraits<PcmT>::NBUF;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_win = reinterpret_cast<float*>(smem_raw);                   // 512
  float2* s_tw256 = reinterpret_cast<float2*>(s_win + 512);            // 256, [p][r]
  float2* s_w512 = s_tw256 + 256;                                      // 256
  float* s_xr = reinterpret_cast<float*>(s_w512 + 256);                // FT * XF
  float* s_xi = s_xr + FT * XF;                                        // FT * XF
  double* s_red = reinterpret_cast<double*>(s_xi + FT * XF);           // 3 * 8
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_red + 24);   // 2
  PcmT* s_pcm = reinterpret_cast<PcmT*>(s_bar + 2);                    // NBUF * (FT+1)*256
  constexpr int PCM_BUF = (FT + 1) * AFP_N_HOP;

  const int tid = threadIdx.x;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar)));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar + 1)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 512; i += K1_THREADS) s_win[i] = a.window_f[i];
  for (int i = tid; i < 256; i += K1_THREADS) {
    s_tw256[i] = a.tw256_f[i];
    s_w512[i] = a.w512_f[i];
  }
  __syncthreads();

  const int g = tid >> 4, r = tid & 15, lane = tid & 31;
  const int src_lane = (lane & 16) | ((16 - r) & 15);
  uint32_t phases = 0u;

  int tile = a.tile_begin + blockIdx.x;
  if (tile >= a.tile_end) return;
  const int G = gridDim.x;
  TileInfo cur = make_tile<PcmT>(a, a.items[a.tile_item[tile]], tile);
  TileInfo nxt = cur;
  if (tile + G < a.tile_end) nxt = make_tile<PcmT>(a, a.items[a.tile_item[tile + G]], tile + G);
  int item_nn = (tile + 2 * G < a.tile_end) ? a.tile_item[tile + 2 * G] : 0;
  stage_tile<PcmT>(a, cur, s_pcm, s_bar);
  int buf = 0;

  for (; tile < a.tile_end; tile += G) {
    const int next = tile + G;
    if (NBUF == 2 && next < a.tile_end) stage_tile<PcmT>(a, nxt, s_pcm + (buf ^ 1) * PCM_BUF, s_bar + (buf ^ 1));
    ItemDesc desc_nn = a.items[item_nn];
    const int item_n3 = (tile + 3 * G < a.tile_end) ? a.tile_item[tile + 3 * G] : 0;
    if (cur.tma & 0xffff) {
      wait_bar(s_bar + buf, (phases >> buf) & 1u);
      phases ^= 1u << buf;
    }
    if ((cur.tma & 0xffff) != (cur.nft + 1) * AFP_N_HOP) __syncthreads();
    const bool active = g < cur.nft;
    const int64_t frame = cur.frame0 + g;
    float vmax = 0.0f, vsum = 0.0f;
    int hmin = 0x7f800000;
    float zr[16], zi[16];
    if (active) {
      const PcmT* fr = s_pcm + buf * PCM_BUF + g * AFP_N_HOP;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        const int i0 = 2 * (16 * q + r);
        const float2 w = *reinterpret_cast<const float2*>(s_win + i0);
        float x0, x1;
        PcmF32<PcmT>::load2(fr + i0, x0, x1);
        zr[q] = x0 * w.x;
        zi[q] = x1 * w.y;
      }
      afp_fft16(zr, zi);
      float* xr = s_xr + g * XF;
      float* xi = s_xi + g * XF;
#pragma unroll
      for (int p = 0; p < 16; ++p) {
        const float2 w = s_tw256[p * 16 + r];
        xr[p * XS + r] = zr[p] * w.x - zi[p] * w.y;
        xi[p * XS + r] = zr[p] * w.y + zi[p] * w.x;
      }
    }
    __syncwarp();
    if (active) {
      const float* xr = s_xr + g * XF + r * XS;
      const float* xi = s_xi + g * XF + r * XS;
#pragma unroll
      for (int q = 0; q < 16; ++q) {
        zr[q] = xr[q];
        zi[q] = xi[q];
      }
      afp_fft16(zr, zi);
    }
    {
      float* out = a.logs_f + frame * AFP_NBINS;
#pragma unroll
      for (int s = 0; s < 8; ++s) {
        float c = __shfl_sync(0xffffffffu, zr[15 - s], src_lane);
        float d = __shfl_sync(0xffffffffu, zi[15 - s], src_lane);
        if (r == 0) {
          c = zr[(16 - s) & 15];
          d = zi[(16 - s) & 15];
        }
        if (active) {
          const int k = r + 16 * s;
          const float2 w = s_w512[k];
          const float er = zr[s] + c, ei = zi[s] - d, orr = zi[s] + d, oi = c - zr[s];
          const float pr = w.x * orr - w.y * oi, pi = w.x * oi + w.y * orr;
          const float ar = er + pr, ai = ei + pi, br = er - pr, bi = ei - pi;
          const float ssa = ar * ar + ai * ai, ssb = br * br + bi * bi;     // 4|X|^2
          const float la = 0.5f * __logf(0.25f * ssa), lb = 0.5f * __logf(0.25f * ssb);
          out[k] = la;
          if (k != 0) out[256 - k] = lb; else a.nyq[frame] = (double)lb;
          if (WRITE_MAG) {
            a.mag[frame * 257 + k] = (double)sqrtf(0.25f * ssa);
            a.mag[frame * 257 + 256 - k] = (double)sqrtf(0.25f * ssb);
          }
          vmax = fmaxf(vmax, fmaxf(ssa, ssb));
          hmin = min(hmin, min(__float_as_int(ssa), __float_as_int(ssb)));
          vsum += la + lb;
        }
      }
      if (active && r == 0) {
        const float2 w = s_w512[128];
        const float er = 2.0f * zr[8], orr = 2.0f * zi[8];
        const float ar = er + w.x * orr, ai = w.y * orr;
        const float ss = ar * ar + ai * ai;
        const float lg = 0.5f * __logf(0.25f * ss);
        out[128] = lg;
        if (WRITE_MAG) a.mag[frame * 257 + 128] = (double)sqrtf(0.25f * ss);
        vmax = fmaxf(vmax, ss);
        hmin = min(hmin, __float_as_int(ss));
        vsum += lg;
      }
    }
    hmin = __reduce_min_sync(0xffffffffu, hmin);
    double vmin = hmin >= 0x7f800000 ? INFINITY
                  : (hmin < 0x00800000 ? -INFINITY : (double)(0.5f * __logf(0.25f * __int_as_float(hmin))) - 1e-6);
    double dmax = (double)vmax, dsum = (double)vsum;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      dmax = fmax(dmax, __shfl_xor_sync(0xffffffffu, dmax, o));
      dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
    }
    if (lane == 0) {
      s_red[(tid >> 5) * 3 + 0] = dmax;
      s_red[(tid >> 5) * 3 + 1] = vmin;
      s_red[(tid >> 5) * 3 + 2] = dsum;
    }
    __syncthreads();
    if (tid == 0) {
      double m = 0.0, mn = INFINITY, sm = 0.0;
#pragma unroll
      for (int w = 0; w < K1_THREADS / 32; ++w) {
        m = fmax(m, s_red[w * 3 + 0]);
        mn = fmin(mn, s_red[w * 3 + 1]);
        sm += s_red[w * 3 + 2];
      }
      a.tile_stats[(size_t)tile * 3 + 0] = 0.25 * m;
      a.tile_stats[(size_t)tile * 3 + 1] = mn;
      a.tile_stats[(size_t)tile * 3 + 2] = sm;
    }
    if (NBUF == 2) buf ^= 1;
    cur = nxt;
    if (tile + 2 * G < a.tile_end) nxt = make_tile<PcmT>(a, desc_nn, tile + 2 * G);
    item_nn = item_n3;
    if (NBUF == 1 && next < a.tile_end) stage_tile<PcmT>(a, cur, s_pcm, s_bar);
  }
}

constexpr size_t k1_f32_smem_bytes(size_t pcm_elem, int nbuf) {
  return 512 * 4 + 256 * 8 * 2 + 2 * FT * XF * 4 + 24 * 8 + 16 + nbuf * (FT + 1) * 256 * pcm_elem;
}

// tile -> item table (one thread per item)
__global__ void afp_tile_table_kernel(const ItemDesc* items, int nitems, int32_t* tile_item) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nitems) return;
  const ItemDesc it = items[i];
  const int nt = (it.nframes + FT - 1) / FT;
  for (int k = 0; k < nt; ++k) tile_item[it.tile_base + k] = i;
}

constexpr size_t k1_smem_bytes(size_t pcm_elem, int nbuf) {
  return 512 * 8 + 256 * 16 * 2 + LOGTAB * LOGCOPIES * 16 + 2 * FT * XF * 8 + 24 * 8 + 16 +
         nbuf * (FT + 1) * 256 * pcm_elem;
}

// ---- per-item statistics: floor, mean (audfprint_analyze.py:283-286) ----------
// Three tiny launches.  (1) one WARP per item reduces the tile partials in a fixed order
// (deterministic): floor, all-zero flag, mean of the un-floored logs, and whether anything
// may sit below the floor.  (2) For such items only (digital silence, or the rare file whose
// smallest bin is 120 dB below its largest) the tile sums are recomputed with the floor
// applied, one CTA per tile so that a single long file cannot serialise the batch.
// (3) the means of those items are re-reduced.
__global__ void __launch_bounds__(256) afp_stats_kernel(const ItemDesc* items, int item0, int nitems,
                                                        const double* tile_stats, ItemStats* out, int phase) {
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= nitems) return;
  const int item = item0 + w;
  if (phase == 1 && !out[item].floored) return;       // only items whose tile sums were floored
  const ItemDesc it = items[item];
  const int ntiles = (it.nframes + FT - 1) / FT;
  double m = 0.0, mn = INFINITY, sm = 0.0;
  for (int i = lane; i < ntiles; i += 32) {
    const double* ts = tile_stats + (size_t)(it.tile_base + i) * 3;
    m = fmax(m, ts[0]);
    mn = fmin(mn, ts[1]);
    sm += ts[2];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    sm += __shfl_xor_sync(0xffffffffu, sm, o);
  }
  if (lane == 0) {
    const bool allzero = !(m > 0.0);
    ItemStats st;
    st.logfloor = allzero ? 0.0 : log(sqrt(m) / 1e6);
    st.mean = (allzero || it.nframes == 0) ? 0.0 : sm / ((double)it.nframes * 257.0);
    st.allzero = allzero ? 1 : 0;
    st.floored = (phase == 0 && !allzero && mn < st.logfloor) ? 1 : 0;   // needs the floored sums
    out[item] = st;
  }
}

// floored tile sums of the flagged items: grid (items, FS_SPLIT); un-flagged items leave after
// one load, a flagged item's tiles are dealt round-robin to its FS_SPLIT CTAs
constexpr int FS_SPLIT = 8;
template <typename R>
__global__ void __launch_bounds__(256) afp_floorsum_kernel(const ItemDesc* items, int item0, const ItemStats* stats,
                                                           const R* logs, const double* nyq,
                                                           double* tile_stats) {
  __shared__ double s_part[8];
  const int item = item0 + blockIdx.x;
  const ItemStats st = stats[item];
  if (!st.floored) return;                        // uniform
  const ItemDesc it = items[item];
  const int ntiles = (it.nframes + FT - 1) / FT;
  for (int k = blockIdx.y; k < ntiles; k += FS_SPLIT) {
    const int t0 = k * FT, nft = min(FT, it.nframes - t0);
    const R* L = logs + (size_t)(it.frame_base + t0) * AFP_NBINS;
    double acc = 0.0;
    for (int i = threadIdx.x; i < nft * AFP_NBINS; i += 256) acc += fmax((double)L[i], st.logfloor);
    if ((int)threadIdx.x < nft) acc += fmax(nyq[it.frame_base + t0 + threadIdx.x], st.logfloor);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < 8; ++w) t += s_part[w];
      tile_stats[(size_t)(it.tile_base + k) * 3 + 2] = t;
    }
  }
}

// ---- conditioned spectrogram for the parity entry point afp_sgram -------------
// One thread per (item, bin); serial over time.  Not on the product path (the
// peak kernel fuses this recursion); exists so that the test can compare the
// sgram itself with the oracle.
template <typename R>
__global__ void afp_sgram_kernel(const ItemDesc* items, const ItemStats* stats, const R* logs,
                                 double pole, double* out) {
  const ItemDesc it = items[blockIdx.x];
  const ItemStats st = stats[blockIdx.x];
  const int b = threadIdx.x;
  double z = 0.0;
  for (int t = 0; t < it.nframes; ++t) {
    const size_t idx = (size_t)(it.frame_base + t) * AFP_NBINS + b;
    double x = st.allzero ? 0.0 : __dsub_rn(fmax((double)logs[idx], st.logfloor), st.mean);
    const double y = __dadd_rn(z, x);
    z = __dadd_rn(-x, __dmul_rn(pole, y));
    out[idx] = y;
  }
}

}  // namespace

int afp_launch_tile_table(afp_ctx* c) {
  if (c->nitems == 0) return AFP_OK;
  afp_tile_table_kernel<<<(c->nitems + 255) / 256, 256, 0, c->stream>>>(c->d_items.as<ItemDesc>(), c->nitems,
                                                                       c->d_tile_item.as<int32_t>());
  AFP_CUDA(c, cudaGetLastError());
  c->launches++;
  return AFP_OK;
}

int afp_launch_stft(afp_ctx* c, const void* pcm, int dtype, double* mag_out, int64_t tile0, int64_t ntiles) {
  if (ntiles <= 0) return AFP_OK;
  StftArgs a;
  a.pcm = pcm;
  a.items = c->d_items.as<ItemDesc>();
  a.tile_item = c->d_tile_item.as<int32_t>();
  a.nitems = c->nitems;
  a.tile_begin = (int)tile0;
  a.tile_end = (int)(tile0 + ntiles);
  a.window = c->d_window.as<double>() + (dtype == AFP_PCM_I16 ? AFP_N_FFT : 0);
  a.tw256 = c->d_twid.as<double2>();
  a.w512 = c->d_twid.as<double2>() + 256;
  a.logtab = c->d_twid.as<double2>() + 512;
  a.logs = c->d_logs.as<double>();
  a.nyq = c->d_nyq.as<double>();
  a.tile_stats = c->d_tile_stats.as<double>();
  a.mag = mag_out;
  a.window_f = c->d_window_f.as<float>() + (dtype == AFP_PCM_I16 ? AFP_N_FFT : 0);
  a.tw256_f = c->d_twid_f.as<float2>();
  a.w512_f = c->d_twid_f.as<float2>() + 256;
  a.logs_f = c->d_logs.as<float>();
  const bool f32 = c->ap.spectrogram_fp32 != 0;
  const int nctas = (int)std::min<int64_t>(ntiles, (int64_t)c->num_sms * (f32 ? 3 : 2));
  const dim3 grid((unsigned)nctas), block(K1_THREADS);
  cudaError_t e;
#define LAUNCH_F32(T, M)                                                                              \
  do {                                                                                                \
    const size_t smem = k1_f32_smem_bytes(sizeof(T), PcmTraits<T>::NBUF);                             \
    e = cudaFuncSetAttribute(afp_stft_f32_kernel<T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             (int)smem);                                                              \
    if (e == cudaSuccess) afp_stft_f32_kernel<T, M><<<grid, block, smem, c->stream>>>(a);             \
  } while (0)
#define LAUNCH(T, M)                                                                              \
  do {                                                                                            \
    const size_t smem = k1_smem_bytes(sizeof(T), PcmTraits<T>::NBUF);                             \
    e = cudaFuncSetAttribute(afp_stft_kernel<T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             (int)smem);                                                          \
    if (e == cudaSuccess) afp_stft_kernel<T, M><<<grid, block, smem, c->stream>>>(a);             \
  } while (0)
  if (f32) {
    if (dtype == AFP_PCM_I16) {
      if (mag_out) LAUNCH_F32(int16_t, true); else LAUNCH_F32(int16_t, false);
    } else {
      if (mag_out) LAUNCH_F32(float, true); else LAUNCH_F32(float, false);
    }
  } else if (dtype == AFP_PCM_I16) {
    if (mag_out) LAUNCH(int16_t, true); else LAUNCH(int16_t, false);
  } else {
    if (mag_out) LAUNCH(float, true); else LAUNCH(float, false);
  }
#undef LAUNCH
#undef LAUNCH_F32
  AFP_CUDA(c, e);
  AFP_CUDA(c, cudaGetLastError());
  c->launches++;
  return AFP_OK;
}

int afp_launch_stats(afp_ctx* c, int item0, int nitems) {
  if (nitems <= 0) return AFP_OK;
  const ItemDesc* items = c->d_items.as<ItemDesc>();
  ItemStats* st = c->d_item_stats.as<ItemStats>();
  double* ts = c->d_tile_stats.as<double>();
  afp_stats_kernel<<<(nitems + 7) / 8, 256, 0, c->stream>>>(items, item0, nitems, ts, st, 0);
  AFP_CUDA(c, cudaGetLastError());
  if (c->ap.spectrogram_fp32)
    afp_floorsum_kernel<float><<<dim3((unsigned)nitems, FS_SPLIT), 256, 0, c->stream>>>(
        items, item0, st, c->d_logs.as<float>(), c->d_nyq.as<double>(), ts);
  else
    afp_floorsum_kernel<double><<<dim3((unsigned)nitems, FS_SPLIT), 256, 0, c->stream>>>(
        items, item0, st, c->d_logs.as<double>(), c->d_nyq.as<double>(), ts);
  AFP_CUDA(c, cudaGetLastError());
  afp_stats_kernel<<<(nitems + 7) / 8, 256, 0, c->stream>>>(items, item0, nitems, ts, st, 1);
  AFP_CUDA(c, cudaGetLastError());
  c->launches += 3;
  return AFP_OK;
}

int afp_launch_sgram(afp_ctx* c, double* sgram_out) {
  if (c->nitems == 0 || c->total_frames == 0) return AFP_OK;
  if (c->ap.spectrogram_fp32)
    afp_sgram_kernel<float><<<c->nitems, AFP_NBINS, 0, c->stream>>>(
        c->d_items.as<ItemDesc>(), c->d_item_stats.as<ItemStats>(), c->d_logs.as<float>(), c->ap.hpf_pole, sgram_out);
  else
    afp_sgram_kernel<double><<<c->nitems, AFP_NBINS, 0, c->stream>>>(
        c->d_items.as<ItemDesc>(), c->d_item_stats.as<ItemStats>(), c->d_logs.as<double>(), c->ap.hpf_pole, sgram_out);
  AFP_CUDA(c, cudaGetLastError());
  c->launches++;
  return AFP_OK;
}
