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
