// Host-only check of the K1 index algebra (afp_fft.cuh): emulates the 16
// cooperating threads of one frame sequentially and compares |X[k]|^2 with a
// direct O(N^2) DFT in long double.  Built and run by tests/test_fft_host.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "afp_fft.cuh"

int main() {
  const int N = 512;
  static double x[N];
  srand(7);
  for (int i = 0; i < N; ++i) x[i] = (rand() / (double)RAND_MAX) * 2 - 1;
  // reference DFT
  static long double Xr[257], Xi[257];
  const long double PI = acosl(-1.0L);
  for (int k = 0; k <= 256; ++k) {
    long double sr = 0, si = 0;
    for (int n = 0; n < N; ++n) {
      long double a = -2 * PI * (long double)((n * k) % N) / N;
      sr += x[n] * cosl(a);
      si += x[n] * sinl(a);
    }
    Xr[k] = sr; Xi[k] = si;
  }
  // step A: thread r: z[16q + r], q = 0..15 -> fft16 over q -> A[r][p] * W256^(r p) -> exch[p][r]
  static double er[16][17], ei[16][17];
  for (int r = 0; r < 16; ++r) {
    double ar[16], ai[16];
    for (int q = 0; q < 16; ++q) { ar[q] = x[2 * (16 * q + r)]; ai[q] = x[2 * (16 * q + r) + 1]; }
    afp_fft16(ar, ai);
    for (int p = 0; p < 16; ++p) {
      double ang = -2 * M_PI * ((r * p) & 255) / 256.0;
      double wr = cos(ang), wi = sin(ang);
      er[p][r] = ar[p] * wr - ai[p] * wi;
      ei[p][r] = ar[p] * wi + ai[p] * wr;
    }
  }
  // step B: thread p: fft16 over r -> Z[p + 16 s]
  static double Zr[16][16], Zi[16][16];   // [p][s]
  for (int p = 0; p < 16; ++p) {
    double ar[16], ai[16];
    for (int r = 0; r < 16; ++r) { ar[r] = er[p][r]; ai[r] = ei[p][r]; }
    afp_fft16(ar, ai);
    for (int s = 0; s < 16; ++s) { Zr[p][s] = ar[s]; Zi[p][s] = ai[s]; }
  }
  // post: thread p, register s: k = p + 16 s; partner lane (16-p)&15, register p ? 15-s : (16-s)&15
  double worst = 0;
  for (int p = 0; p < 16; ++p)
    for (int s = 0; s < 16; ++s) {
      int k = p + 16 * s;
      int pl = (16 - p) & 15, ps = p ? 15 - s : (16 - s) & 15;
      double ang = -2 * M_PI * k / 512.0;
      double xr, xi;
      afp_real_post(Zr[p][s], Zi[p][s], Zr[pl][ps], Zi[pl][ps], cos(ang), sin(ang), xr, xi);
      double e = fabs(xr - (double)Xr[k]) + fabs(xi - (double)Xi[k]);
      if (e > worst) worst = e;
    }
  // Nyquist: X[256] = Re Z[0] - Im Z[0]
  double e = fabs((Zr[0][0] - Zi[0][0]) - (double)Xr[256]) + fabs((double)Xi[256]);
  if (e > worst) worst = e;
  printf("worst_abs_err %.3e\n", worst);
  return worst < 1e-12 ? 0 : 1;
}
