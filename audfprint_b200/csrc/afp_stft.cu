// K1 — fused framing + Hann window + 512-point real FFT + log-magnitude (FP64).
//
// Replaces stft.stft (reference stft.py:62-94: reflect pad 256, hop-256
// framing, window multiply, rfft) and the |.| / log part of
// Analyzer.find_peaks (audfprint_analyze.py:280-285).  The two whole-file
// reductions that follow (floor = max/1e6, mean, :283-286) are produced as
// per-tile partials here and finished in afp_stats_kernel; the per-bin high-pass
// (:293-295) is a time recursion and lives in the peak kernel (afp_peaks.cu).
//
// Work decomposition: one CTA = one tile of 16 consecutive frames of one item
// (file x shift); 16 threads cooperate on a frame (256-point complex FFT as
// 16 x 16, one shared-memory transpose, partner exchange by warp shuffle).
// The hop-strided PCM of a tile is ONE contiguous run of 17*256 samples, staged
// into shared memory by a 1-D TMA bulk copy (cp.async.bulk + mbarrier) when the
// tile is interior and 16-byte aligned, by reflected scalar loads otherwise.
//
// Why FP64: the peak decisions downstream compare these values bit-for-bit the
// way the reference's float64 NumPy path does; an FP32 spectrogram flips a
// decision roughly once per 10^5 frames.
#include <math.h>
#include "afp_fft.cuh"
#include "afp_internal.cuh"

namespace {

constexpr int FT = AFP_FRAMES_PER_TILE;   // 16 frames per tile
constexpr int XS = 17;                    // padded row stride of the 16x16 exchange
constexpr int XF = 16 * XS;               // 272 doubles per frame per component
constexpr int K1_THREADS = 256;

struct StftArgs {
  const void* pcm;
  const ItemDesc* items;
  int nitems;
  const double* window;   // 512
  const double2* w256;    // 256: (cos, -sin)(2 pi k / 256)
  const double2* w512;    // 256: (cos, -sin)(2 pi k / 512)
  double* logs;           // [frames][256]
  double* nyq;            // [frames]
  double* tile_stats;     // [tiles][3]
  double* mag;            // optional [frames][257]
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

__device__ __forceinline__ float pcm_to_f32(int16_t v) { return static_cast<float>(v) * (1.0f / 32768.0f); }
__device__ __forceinline__ float pcm_to_f32(float v) { return v; }

template <typename PcmT, bool WRITE_MAG>
__global__ void __launch_bounds__(K1_THREADS, 2) afp_stft_kernel(StftArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double* s_win = reinterpret_cast<double*>(smem_raw);                 // 512
  double2* s_w256 = reinterpret_cast<double2*>(s_win + 512);           // 256
  double2* s_w512 = s_w256 + 256;                                      // 256
  double* s_xr = reinterpret_cast<double*>(s_w512 + 256);              // FT * XF
  double* s_xi = s_xr + FT * XF;                                       // FT * XF
  double* s_red = s_xi + FT * XF;                                      // 3 * 8
  unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(s_red + 24);
  PcmT* s_pcm = reinterpret_cast<PcmT*>(s_bar + 2);                    // (FT+1)*256, 16 B aligned

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;

  // tile -> item (last item whose tile_base <= tile)
  int lo = 0, hi = a.nitems;
  while (hi - lo > 1) {
    int mid = (lo + hi) >> 1;
    if (a.items[mid].tile_base <= tile) lo = mid; else hi = mid;
  }
  const ItemDesc it = a.items[lo];
  const int t0 = (tile - it.tile_base) * FT;
  const int nft = min(FT, it.nframes - t0);
  const int64_t n = it.nsamples;
  const int64_t j0 = (int64_t)(t0 - 1) * AFP_N_HOP;                 // first sample of the tile (may be < 0)
  const int nsamp = (nft + 1) * AFP_N_HOP;
  const PcmT* src = reinterpret_cast<const PcmT*>(a.pcm) + it.sample_start;

  // ---- stage the PCM run -----------------------------------------------------
  const bool interior = (j0 >= 0) && (j0 + nsamp <= n);
  const bool use_tma = interior && ((reinterpret_cast<uintptr_t>(src + j0) & 15) == 0);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(s_bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (use_tma) {
    if (tid == 0) {
      const uint32_t bytes = nsamp * sizeof(PcmT);
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(s_bar)), "r"(bytes)
                   : "memory");
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              smem_u32(s_pcm)),
          "l"(src + j0), "r"(bytes), "r"(smem_u32(s_bar))
          : "memory");
    }
  } else {
    for (int i = tid; i < nsamp; i += K1_THREADS) s_pcm[i] = src[reflect_index(j0 + i, n)];
  }
  // constant tables (L2-resident) while the bulk copy is in flight
  for (int i = tid; i < 512; i += K1_THREADS) s_win[i] = a.window[i];
  for (int i = tid; i < 256; i += K1_THREADS) {
    s_w256[i] = a.w256[i];
    s_w512[i] = a.w512[i];
  }
  if (use_tma) {
    uint32_t done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done)
          : "r"(smem_u32(s_bar)), "r"(0u)
          : "memory");
    }
  }
  __syncthreads();

  const int g = tid >> 4;   // frame within the tile
  const int r = tid & 15;   // cooperating thread within the frame
  const bool active = g < nft;
  const int64_t frame = it.frame_base + t0 + g;

  double vmax = 0.0, vmin = INFINITY, vsum = 0.0;
  double zr[16], zi[16];
  if (active) {
    // step A: z[16q + r] = (x[2n] w[2n], x[2n+1] w[2n+1]), n = 16q + r
    const PcmT* fr = s_pcm + g * AFP_N_HOP;
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int i0 = 2 * (16 * q + r);
      const double2 w = *reinterpret_cast<const double2*>(s_win + i0);
      zr[q] = (double)pcm_to_f32(fr[i0]) * w.x;
      zi[q] = (double)pcm_to_f32(fr[i0 + 1]) * w.y;
    }
    afp_fft16(zr, zi);
    double* xr = s_xr + g * XF;
    double* xi = s_xi + g * XF;
#pragma unroll
    for (int p = 0; p < 16; ++p) {
      const double2 w = s_w256[(r * p) & 255];
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
  // real-FFT post-processing: partner Z[(256-k)&255] sits in lane (16-p)&15,
  // register 15-s (p > 0) or (16-s)&15 (p == 0, own registers).
  {
    const int lane = tid & 31;
    const int src_lane = (lane & 16) | ((16 - r) & 15);
    double* out = a.logs + frame * AFP_NBINS;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
      double c = __shfl_sync(0xffffffffu, zr[15 - s], src_lane);
      double d = __shfl_sync(0xffffffffu, zi[15 - s], src_lane);
      if (r == 0) {
        c = zr[(16 - s) & 15];
        d = zi[(16 - s) & 15];
      }
      if (active) {
        const int k = r + 16 * s;
        const double2 w = s_w512[k];
        double xr_, xi_;
        afp_real_post(zr[s], zi[s], c, d, w.x, w.y, xr_, xi_);
        const double ss = xr_ * xr_ + xi_ * xi_;
        const double lg = 0.5 * log(ss);
        out[k] = lg;
        if (WRITE_MAG) a.mag[frame * 257 + k] = sqrt(ss);
        vmax = fmax(vmax, ss);
        vmin = fmin(vmin, lg);
        vsum += lg;
      }
    }
    if (active && r == 0) {   // Nyquist bin: X[256] = Re Z[0] - Im Z[0]
      const double xn = zr[0] - zi[0];
      const double ss = xn * xn;
      const double lg = 0.5 * log(ss);
      a.nyq[frame] = lg;
      if (WRITE_MAG) a.mag[frame * 257 + 256] = sqrt(ss);
      vmax = fmax(vmax, ss);
      vmin = fmin(vmin, lg);
      vsum += lg;
    }
  }
  // deterministic CTA reduction of (max |S|^2, min log, sum log)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    vmax = fmax(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    vmin = fmin(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
    vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
  }
  if ((tid & 31) == 0) {
    s_red[(tid >> 5) * 3 + 0] = vmax;
    s_red[(tid >> 5) * 3 + 1] = vmin;
    s_red[(tid >> 5) * 3 + 2] = vsum;
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
    a.tile_stats[(size_t)tile * 3 + 0] = m;
    a.tile_stats[(size_t)tile * 3 + 1] = mn;
    a.tile_stats[(size_t)tile * 3 + 2] = sm;
  }
}

constexpr size_t k1_smem_bytes(size_t pcm_elem) {
  return 512 * 8 + 256 * 16 * 2 + 2 * FT * XF * 8 + 24 * 8 + 16 + (FT + 1) * 256 * pcm_elem;
}

// ---- per-item statistics: floor, mean (audfprint_analyze.py:283-286) ----------
// One CTA per item.  Fast path: no value below the floor -> mean from the tile
// partial sums (fixed order, deterministic).  Slow path (digital silence etc.):
// re-read the stored logs and sum max(L, floor).
__global__ void __launch_bounds__(256) afp_stats_kernel(const ItemDesc* items, int nitems,
                                                        const double* tile_stats, const double* logs,
                                                        const double* nyq, ItemStats* out) {
  __shared__ double s_a[256], s_b[256], s_c[256];
  const int item = blockIdx.x;
  const ItemDesc it = items[item];
  const int tid = threadIdx.x;
  const int ntiles = (it.nframes + FT - 1) / FT;
  double m = 0.0, mn = INFINITY, sm = 0.0;
  for (int i = tid; i < ntiles; i += 256) {
    const double* ts = tile_stats + (size_t)(it.tile_base + i) * 3;
    m = fmax(m, ts[0]);
    mn = fmin(mn, ts[1]);
    sm += ts[2];
  }
  s_a[tid] = m; s_b[tid] = mn; s_c[tid] = sm;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
      s_a[tid] = fmax(s_a[tid], s_a[tid + o]);
      s_b[tid] = fmin(s_b[tid], s_b[tid + o]);
      s_c[tid] += s_c[tid + o];
    }
    __syncthreads();
  }
  const double maxss = s_a[0], minlog = s_b[0];
  double total = s_c[0];
  __syncthreads();
  const bool allzero = !(maxss > 0.0);
  const double logfloor = allzero ? 0.0 : log(sqrt(maxss) / 1e6);
  if (!allzero && minlog < logfloor) {   // uniform across the CTA
    const size_t nl = (size_t)it.nframes * AFP_NBINS;
    const double* L = logs + (size_t)it.frame_base * AFP_NBINS;
    double acc = 0.0;
    for (size_t i = tid; i < nl; i += 256) acc += fmax(L[i], logfloor);
    for (int i = tid; i < it.nframes; i += 256) acc += fmax(nyq[it.frame_base + i], logfloor);
    s_c[tid] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (tid < o) s_c[tid] += s_c[tid + o];
      __syncthreads();
    }
    total = s_c[0];
  }
  if (tid == 0) {
    ItemStats st;
    st.logfloor = logfloor;
    st.mean = (allzero || it.nframes == 0) ? 0.0 : total / ((double)it.nframes * 257.0);
    st.allzero = allzero ? 1 : 0;
    st.pad = 0;
    out[item] = st;
  }
}

// ---- conditioned spectrogram for the parity entry point afp_sgram -------------
// One thread per (item, bin); serial over time.  Not on the product path (the
// peak kernel fuses this recursion); exists so that the test can compare the
// sgram itself with the oracle.
__global__ void afp_sgram_kernel(const ItemDesc* items, const ItemStats* stats, const double* logs,
                                 double pole, double* out) {
  const ItemDesc it = items[blockIdx.x];
  const ItemStats st = stats[blockIdx.x];
  const int b = threadIdx.x;
  double z = 0.0;
  for (int t = 0; t < it.nframes; ++t) {
    const size_t idx = (size_t)(it.frame_base + t) * AFP_NBINS + b;
    double x = st.allzero ? 0.0 : __dsub_rn(fmax(logs[idx], st.logfloor), st.mean);
    const double y = __dadd_rn(z, x);
    z = __dadd_rn(-x, __dmul_rn(pole, y));
    out[idx] = y;
  }
}

}  // namespace

int afp_launch_stft(afp_ctx* c, const void* pcm, int dtype, double* mag_out) {
  if (c->total_tiles == 0) return AFP_OK;
  StftArgs a;
  a.pcm = pcm;
  a.items = c->d_items.as<ItemDesc>();
  a.nitems = c->nitems;
  a.window = c->d_window.as<double>();
  a.w256 = c->d_twid.as<double2>();
  a.w512 = c->d_twid.as<double2>() + 256;
  a.logs = c->d_logs.as<double>();
  a.nyq = c->d_nyq.as<double>();
  a.tile_stats = c->d_tile_stats.as<double>();
  a.mag = mag_out;
  const dim3 grid((unsigned)c->total_tiles), block(K1_THREADS);
  cudaError_t e;
#define LAUNCH(T, M)                                                                              \
  do {                                                                                            \
    const size_t smem = k1_smem_bytes(sizeof(T));                                                 \
    e = cudaFuncSetAttribute(afp_stft_kernel<T, M>, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                             (int)smem);                                                          \
    if (e == cudaSuccess) afp_stft_kernel<T, M><<<grid, block, smem, c->stream>>>(a);             \
  } while (0)
  if (dtype == AFP_PCM_I16) {
    if (mag_out) LAUNCH(int16_t, true); else LAUNCH(int16_t, false);
  } else {
    if (mag_out) LAUNCH(float, true); else LAUNCH(float, false);
  }
#undef LAUNCH
  AFP_CUDA(c, e);
  AFP_CUDA(c, cudaGetLastError());
  c->launches++;
  return AFP_OK;
}

int afp_launch_stats(afp_ctx* c) {
  if (c->nitems == 0) return AFP_OK;
  afp_stats_kernel<<<c->nitems, 256, 0, c->stream>>>(c->d_items.as<ItemDesc>(), c->nitems,
                                                     c->d_tile_stats.as<double>(), c->d_logs.as<double>(),
                                                     c->d_nyq.as<double>(), c->d_item_stats.as<ItemStats>());
  AFP_CUDA(c, cudaGetLastError());
  c->launches++;
  return AFP_OK;
}

int afp_launch_sgram(afp_ctx* c, double* sgram_out) {
  if (c->nitems == 0 || c->total_frames == 0) return AFP_OK;
  afp_sgram_kernel<<<c->nitems, AFP_NBINS, 0, c->stream>>>(c->d_items.as<ItemDesc>(),
                                                           c->d_item_stats.as<ItemStats>(),
                                                           c->d_logs.as<double>(), c->ap.hpf_pole, sgram_out);
  AFP_CUDA(c, cudaGetLastError());
  c->launches++;
  return AFP_OK;
}
