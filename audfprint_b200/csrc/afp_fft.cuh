// FFT building blocks for K1 (real type R = double for the bit-parity path, float for the
// opt-in FP32 spectrogram mode) (512-point real FFT as a 256-point complex
// FFT split 16 x 16 across 16 cooperating threads).  Host+device so that the
// index algebra can be unit-tested on the CPU (tests/test_fft_host.py builds
// csrc/fft_host_check.cu with nvcc -x cu for the host only).
#pragma once

#ifndef __CUDACC__
#define __host__
#define __device__
#define __forceinline__ inline
#endif

#define AFP_C16 ((R)0.92387953251128673848)   // cos(pi/8)
#define AFP_S16 ((R)0.38268343236508978178)   // sin(pi/8)
#define AFP_R16 ((R)0.70710678118654752440)   // sqrt(1/2)

// y *= W16^E, W16 = exp(-2*pi*i/16); E in {0,1,2,3,4,6,9}
template <int E, typename R>
__host__ __device__ __forceinline__ void afp_mul_w16(R& yr, R& yi) {
  R r = yr, i = yi;
  if (E == 0) {
  } else if (E == 1) {
    yr = r * AFP_C16 + i * AFP_S16;
    yi = i * AFP_C16 - r * AFP_S16;
  } else if (E == 2) {
    yr = (r + i) * AFP_R16;
    yi = (i - r) * AFP_R16;
  } else if (E == 3) {
    yr = r * AFP_S16 + i * AFP_C16;
    yi = i * AFP_S16 - r * AFP_C16;
  } else if (E == 4) {
    yr = i;
    yi = -r;
  } else if (E == 6) {
    yr = (i - r) * AFP_R16;
    yi = -(r + i) * AFP_R16;
  } else if (E == 9) {
    // (-C + iS)(r + i*i) = -C r - S i + i (S r - C i)
    yr = -r * AFP_C16 - i * AFP_S16;
    yi = r * AFP_S16 - i * AFP_C16;
  }
}

// forward radix-4 butterfly on (x0..x3) -> (y0..y3), W4 = -i
#define AFP_BFLY4(x0r, x0i, x1r, x1i, x2r, x2i, x3r, x3i, y0r, y0i, y1r, y1i, y2r, y2i, y3r, y3i) \
  {                                                                                                \
    R t0r = x0r + x2r, t0i = x0i + x2i, t1r = x0r - x2r, t1i = x0i - x2i;                          \
    R t2r = x1r + x3r, t2i = x1i + x3i, t3r = x1r - x3r, t3i = x1i - x3i;                          \
    y0r = t0r + t2r; y0i = t0i + t2i;                                                              \
    y2r = t0r - t2r; y2i = t0i - t2i;                                                              \
    y1r = t1r + t3i; y1i = t1i - t3r;                                                              \
    y3r = t1r - t3i; y3i = t1i + t3r;                                                              \
  }

// In-place forward 16-point DFT, natural order in and out:
//   X[k] = sum_n x[n] * exp(-2*pi*i*n*k/16)
// n = j + 4i, k = m + 4n':  X[m+4n'] = sum_j W4^(j n') [ W16^(j m) sum_i W4^(i m) x[j+4i] ]
template <typename R>
__host__ __device__ __forceinline__ void afp_fft16(R (&xr)[16], R (&xi)[16]) {
  R br[16], bi[16];   // b[j + 4m]
  AFP_BFLY4(xr[0], xi[0], xr[4], xi[4], xr[8], xi[8], xr[12], xi[12],
            br[0], bi[0], br[4], bi[4], br[8], bi[8], br[12], bi[12]);
  AFP_BFLY4(xr[1], xi[1], xr[5], xi[5], xr[9], xi[9], xr[13], xi[13],
            br[1], bi[1], br[5], bi[5], br[9], bi[9], br[13], bi[13]);
  AFP_BFLY4(xr[2], xi[2], xr[6], xi[6], xr[10], xi[10], xr[14], xi[14],
            br[2], bi[2], br[6], bi[6], br[10], bi[10], br[14], bi[14]);
  AFP_BFLY4(xr[3], xi[3], xr[7], xi[7], xr[11], xi[11], xr[15], xi[15],
            br[3], bi[3], br[7], bi[7], br[11], bi[11], br[15], bi[15]);
  // twiddles W16^(j*m) on b[j + 4m]
  afp_mul_w16<1, R>(br[5], bi[5]);    // j=1,m=1
  afp_mul_w16<2, R>(br[9], bi[9]);    // j=1,m=2
  afp_mul_w16<3, R>(br[13], bi[13]);  // j=1,m=3
  afp_mul_w16<2, R>(br[6], bi[6]);    // j=2,m=1
  afp_mul_w16<4, R>(br[10], bi[10]);  // j=2,m=2
  afp_mul_w16<6, R>(br[14], bi[14]);  // j=2,m=3
  afp_mul_w16<3, R>(br[7], bi[7]);    // j=3,m=1
  afp_mul_w16<6, R>(br[11], bi[11]);  // j=3,m=2
  afp_mul_w16<9, R>(br[15], bi[15]);  // j=3,m=3
  // second radix-4 over j for each m: X[m + 4n']
  AFP_BFLY4(br[0], bi[0], br[1], bi[1], br[2], bi[2], br[3], bi[3],
            xr[0], xi[0], xr[4], xi[4], xr[8], xi[8], xr[12], xi[12]);
  AFP_BFLY4(br[4], bi[4], br[5], bi[5], br[6], bi[6], br[7], bi[7],
            xr[1], xi[1], xr[5], xi[5], xr[9], xi[9], xr[13], xi[13]);
  AFP_BFLY4(br[8], bi[8], br[9], bi[9], br[10], bi[10], br[11], bi[11],
            xr[2], xi[2], xr[6], xi[6], xr[10], xi[10], xr[14], xi[14]);
  AFP_BFLY4(br[12], bi[12], br[13], bi[13], br[14], bi[14], br[15], bi[15],
            xr[3], xi[3], xr[7], xi[7], xr[11], xi[11], xr[15], xi[15]);
}

// |X[k]|^2 of the 512-point real FFT from the 256-point complex FFT Z of
// z[n] = x[2n] + i x[2n+1]:  X[k] = Xe + W512^k Xo with
//   Xe = (Z[k] + conj(Z[256-k]))/2,  Xo = -i (Z[k] - conj(Z[256-k]))/2.
// (a,b) = Z[k], (c,d) = Z[(256-k) & 255], (wr,wi) = W512^k = (cos, -sin)(2 pi k/512).
template <typename R>
__host__ __device__ __forceinline__ void afp_real_post(R a, R b, R c, R d, R wr, R wi, R& xr, R& xi) {
  R er = (R)0.5 * (a + c), ei = (R)0.5 * (b - d);
  R orr = (R)0.5 * (b + d), oi = (R)-0.5 * (a - c);
  xr = er + (wr * orr - wi * oi);
  xi = ei + (wr * oi + wi * orr);
}

// ---- K1 v2: 256-point complex FFT as 8 x 8 x 4 over the 32 lanes of ONE warp ------------
// Input index m = 32a + 4b + c (a,b < 8, c < 4), output k = k1 + 8 k2 + 64 k3:
//   Z[k] = sum_c W4^(c k3) W32^(c k2) sum_b W8^(b k2) W256^((4b+c) k1) sum_a W8^(a k1) z[m]
// stage A: lane t = 4b + c holds a = 0..7   -> 8-point DFT over a, twiddle W256^(t k1), exchange 1
// stage B: lane u = k1 + 8c holds b = 0..7  -> 8-point DFT over b, twiddle W32^(c k2),  exchange 2
// stage C: lane L = k1 + 8j holds k2 in {k2a, k2b}, all c -> two radix-4 over c:
//          za[k3] = Z[L + 64 k3] (k2a = j), zb[k3] = Z[k1 + 8 k2b + 64 k3]
// k2b is chosen so that the partner Z[256 - k] of every za value is the zb[3 - k3] of ONE other
// lane (or of the lane itself): k2b = 7 - j, partner lane ((8-k1)&7) + 8j; lanes with k1 = 0
// take k2b = 8 - j (j > 0) or 4 (lane 0) and are their own partners.  Each lane then owns four
// (k, 256-k) pairs; lane 0 owns the five that involve bins 0/256, 64/192, 128, 32/224, 96/160.
// Both exchange layouts are padded so that the 64-bit accesses of a half-warp hit 16 distinct
// 8-byte bank pairs.
#define AFP_V2_XN 324   // doubles per component of one frame's exchange buffer

__host__ __device__ __forceinline__ int afp_v2_x1(int k1, int t) { return k1 * 34 + t; }
__host__ __device__ __forceinline__ int afp_v2_x2(int c, int k1, int k2) { return c * 81 + k1 * 10 + k2; }
__host__ __device__ __forceinline__ int afp_v2_k2b(int k1, int j) { return k1 ? 7 - j : (j ? 8 - j : 4); }
__host__ __device__ __forceinline__ int afp_v2_partner_lane(int k1, int j) { return ((8 - k1) & 7) + 8 * j; }

// In-place forward 8-point DFT, natural order in and out (n = j + 2i, k = m + 4n').
template <typename R>
__host__ __device__ __forceinline__ void afp_fft8(R (&xr)[8], R (&xi)[8]) {
  R er[4], ei[4], qr[4], qi[4];
  AFP_BFLY4(xr[0], xi[0], xr[2], xi[2], xr[4], xi[4], xr[6], xi[6],
            er[0], ei[0], er[1], ei[1], er[2], ei[2], er[3], ei[3]);
  AFP_BFLY4(xr[1], xi[1], xr[3], xi[3], xr[5], xi[5], xr[7], xi[7],
            qr[0], qi[0], qr[1], qi[1], qr[2], qi[2], qr[3], qi[3]);
  afp_mul_w16<2, R>(qr[1], qi[1]);   // W8^1
  afp_mul_w16<4, R>(qr[2], qi[2]);   // W8^2
  afp_mul_w16<6, R>(qr[3], qi[3]);   // W8^3
#pragma unroll
  for (int m = 0; m < 4; ++m) {
    xr[m] = er[m] + qr[m];
    xi[m] = ei[m] + qi[m];
    xr[m + 4] = er[m] - qr[m];
    xi[m + 4] = ei[m] - qi[m];
  }
}

// stage A of lane t: z[a] = windowed samples (x[2m], x[2m+1]), m = 32a + t.  V2 = double2/float2-like.
template <typename R, typename V2>
__host__ __device__ __forceinline__ void afp_v2_stage_a(int t, R (&zr)[8], R (&zi)[8], const V2* tw1, R* xr, R* xi) {
  afp_fft8(zr, zi);
  xr[afp_v2_x1(0, t)] = zr[0];
  xi[afp_v2_x1(0, t)] = zi[0];
#pragma unroll
  for (int k1 = 1; k1 < 8; ++k1) {
    const V2 w = tw1[k1 * 32 + t];
    xr[afp_v2_x1(k1, t)] = zr[k1] * w.x - zi[k1] * w.y;
    xi[afp_v2_x1(k1, t)] = zr[k1] * w.y + zi[k1] * w.x;
  }
}

// stage B of lane u: reads exchange 1, leaves B'[k2] in registers (store after a warp sync).
template <typename R, typename V2>
__host__ __device__ __forceinline__ void afp_v2_stage_b(int u, R (&zr)[8], R (&zi)[8], const V2* tw2, const R* xr,
                                                        const R* xi) {
  const int k1 = u & 7, c = u >> 3;
#pragma unroll
  for (int b = 0; b < 8; ++b) {
    zr[b] = xr[afp_v2_x1(k1, 4 * b + c)];
    zi[b] = xi[afp_v2_x1(k1, 4 * b + c)];
  }
  afp_fft8(zr, zi);
#pragma unroll
  for (int k2 = 1; k2 < 8; ++k2) {
    const V2 w = tw2[c * 8 + k2];
    const R r = zr[k2], i = zi[k2];
    zr[k2] = r * w.x - i * w.y;
    zi[k2] = r * w.y + i * w.x;
  }
}

template <typename R>
__host__ __device__ __forceinline__ void afp_v2_store_b(int u, const R (&zr)[8], const R (&zi)[8], R* xr, R* xi) {
  const int k1 = u & 7, c = u >> 3;
#pragma unroll
  for (int k2 = 0; k2 < 8; ++k2) {
    xr[afp_v2_x2(c, k1, k2)] = zr[k2];
    xi[afp_v2_x2(c, k1, k2)] = zi[k2];
  }
}

// stage C of lane L: za[k3] = Z[L + 64 k3], zb[k3] = Z[k1 + 8 k2b + 64 k3].
template <typename R>
__host__ __device__ __forceinline__ void afp_v2_stage_c(int L, R (&zar)[4], R (&zai)[4], R (&zbr)[4], R (&zbi)[4],
                                                        const R* xr, const R* xi) {
  const int k1 = L & 7, j = L >> 3;
  const int k2b = afp_v2_k2b(k1, j);
  R ir[4], ii[4];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    ir[c] = xr[afp_v2_x2(c, k1, j)];
    ii[c] = xi[afp_v2_x2(c, k1, j)];
  }
  AFP_BFLY4(ir[0], ii[0], ir[1], ii[1], ir[2], ii[2], ir[3], ii[3],
            zar[0], zai[0], zar[1], zai[1], zar[2], zai[2], zar[3], zai[3]);
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    ir[c] = xr[afp_v2_x2(c, k1, k2b)];
    ii[c] = xi[afp_v2_x2(c, k1, k2b)];
  }
  AFP_BFLY4(ir[0], ii[0], ir[1], ii[1], ir[2], ii[2], ir[3], ii[3],
            zbr[0], zbi[0], zbr[1], zbi[1], zbr[2], zbi[2], zbr[3], zbi[3]);
}

// 4|X[k]|^2 and 4|X[256-k]|^2 of the 512-point real FFT from Z[k] = (a,b), Z[(256-k)&255] = (c,d),
// W512^k = (wx, wy):  2Xe = Z[k] + conj(Zp), 2Xo = -i (Z[k] - conj(Zp)), P = W512^k * 2Xo,
// 4|X[k]|^2 = |2Xe + P|^2, 4|X[256-k]|^2 = |2Xe - P|^2.
template <typename R>
__host__ __device__ __forceinline__ void afp_pair_power(R a, R b, R c, R d, R wx, R wy, R& ssa, R& ssb) {
  const R er = a + c, ei = b - d, orr = b + d, oi = c - a;
  const R pr = wx * orr - wy * oi, pi = wx * oi + wy * orr;
  const R ar = er + pr, ai = ei + pi, br = er - pr, bi = ei - pi;
  ssa = ar * ar + ai * ai;
  ssb = br * br + bi * bi;
}

// The pairs lane L owns.  p[k3] = partner lane's zb[3 - k3] (shuffled in by the caller).
// emit(k, ss) is called exactly once for every bin k = 0..256 over the 32 lanes of a frame.
template <typename R, typename V2, class Emit>
__host__ __device__ __forceinline__ void afp_v2_pairs(int L, const R (&zar)[4], const R (&zai)[4], const R (&zbr)[4],
                                                      const R (&zbi)[4], const R (&pr)[4], const R (&pi)[4],
                                                      const V2* w512, Emit& emit) {
#pragma unroll
  for (int k3 = 0; k3 < 4; ++k3) {
    int k = L + 64 * k3;
    R a = zar[k3], b = zai[k3], c = pr[k3], d = pi[k3];
    if (L == 0) {
      c = zar[(4 - k3) & 3];
      d = zai[(4 - k3) & 3];
      if (k3 == 3) {   // (192, 64) is pair k3 = 1 again: use the slot for (32, 224)
        a = zbr[0]; b = zbi[0]; c = zbr[3]; d = zbi[3];
        k = 32;
      }
    }
    const V2 w = w512[k];
    R ssa, ssb;
    afp_pair_power<R>(a, b, c, d, w.x, w.y, ssa, ssb);
    emit(k, ssa);
    if (k != 128) emit(256 - k, ssb);
  }
  if (L == 0) {   // the 129th pair computation of the frame: (96, 160)
    const V2 w = w512[96];
    R ssa, ssb;
    afp_pair_power<R>(zbr[1], zbi[1], zbr[2], zbi[2], w.x, w.y, ssa, ssb);
    emit(96, ssa);
    emit(160, ssb);
  }
}
