// Host-only check of the K1 index algebra (afp_fft.cuh): emulates the 16
// cooperating threads of one frame sequentially and compares |X[k]|^2 with a
// direct O(N^2) DFT in long double.  Built and run by tests/test_fft_host.py.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include "afp_fft.cuh"

// ---- K1 v2 (8 x 8 x 4 over one warp): run the very stage functions the kernel uses, lane by
// lane, with plain arrays standing in for shared memory and the partner shuffle.
struct V2d { double x, y; };
struct HostEmit {
  double ss[257];
  int seen[257];
  void operator()(int k, double v) { ss[k] = v; seen[k]++; }
};

static int check_v2(const double* x, const long double* Xr, const long double* Xi) {
  static V2d tw1[256], tw2[32], w512[256];
  for (int k1 = 0; k1 < 8; ++k1)
    for (int t = 0; t < 32; ++t) {
      double ang = -2 * M_PI * ((t * k1) & 255) / 256.0;
      tw1[k1 * 32 + t] = {cos(ang), sin(ang)};
    }
  for (int c = 0; c < 4; ++c)
    for (int k2 = 0; k2 < 8; ++k2) {
      double ang = -2 * M_PI * ((c * k2) & 31) / 32.0;
      tw2[c * 8 + k2] = {cos(ang), sin(ang)};
    }
  for (int k = 0; k < 256; ++k) w512[k] = {cos(-2 * M_PI * k / 512.0), sin(-2 * M_PI * k / 512.0)};
  static double xr[AFP_V2_XN], xi[AFP_V2_XN];
  for (int i = 0; i < AFP_V2_XN; ++i) xr[i] = xi[i] = NAN;     // any read of an unwritten slot shows up
  for (int t = 0; t < 32; ++t) {
    double zr[8], zi[8];
    for (int a = 0; a < 8; ++a) { zr[a] = x[2 * (32 * a + t)]; zi[a] = x[2 * (32 * a + t) + 1]; }
    afp_v2_stage_a<double, V2d>(t, zr, zi, tw1, xr, xi);
  }
  static double br[32][8], bi[32][8];
  for (int u = 0; u < 32; ++u) afp_v2_stage_b<double, V2d>(u, br[u], bi[u], tw2, xr, xi);
  for (int i = 0; i < AFP_V2_XN; ++i) xr[i] = xi[i] = NAN;
  for (int u = 0; u < 32; ++u) afp_v2_store_b<double>(u, br[u], bi[u], xr, xi);
  static double zar[32][4], zai[32][4], zbr[32][4], zbi[32][4];
  for (int L = 0; L < 32; ++L) afp_v2_stage_c<double>(L, zar[L], zai[L], zbr[L], zbi[L], xr, xi);
  HostEmit em;
  for (int k = 0; k <= 256; ++k) { em.seen[k] = 0; em.ss[k] = 0; }
  for (int L = 0; L < 32; ++L) {
    const int src = afp_v2_partner_lane(L & 7, L >> 3);
    double pr[4], pi[4];
    for (int k3 = 0; k3 < 4; ++k3) { pr[k3] = zbr[src][3 - k3]; pi[k3] = zbi[src][3 - k3]; }
    afp_v2_pairs<double, V2d, HostEmit>(L, zar[L], zai[L], zbr[L], zbi[L], pr, pi, w512, em);
  }
  double worst = 0;
  int bad_cover = 0;
  for (int k = 0; k <= 256; ++k) {
    if (em.seen[k] != 1) { bad_cover++; printf("bin %d emitted %d times\n", k, em.seen[k]); }
    const double want = 4.0 * (double)(Xr[k] * Xr[k] + Xi[k] * Xi[k]);
    const double e = fabs(em.ss[k] - want) / (want + 1e-30);
    if (!(e <= worst)) worst = e;
  }
  // bank-conflict audit of the padded layouts: 64-bit accesses, half-warp at a time
  int conflicts = 0;
  for (int half = 0; half < 2; ++half) {
    for (int b = 0; b < 8; ++b) {            // stage B loads of exchange 1
      int used[16] = {0};
      for (int u = 16 * half; u < 16 * half + 16; ++u) used[afp_v2_x1(u & 7, 4 * b + (u >> 3)) & 15]++;
      for (int i = 0; i < 16; ++i) conflicts += used[i] > 1;
    }
    for (int k2 = 0; k2 < 8; ++k2) {         // stage B stores into exchange 2
      int used[16] = {0};
      for (int u = 16 * half; u < 16 * half + 16; ++u) used[afp_v2_x2(u >> 3, u & 7, k2) & 15]++;
      for (int i = 0; i < 16; ++i) conflicts += used[i] > 1;
    }
    for (int c = 0; c < 4; ++c) {            // stage C loads of exchange 2, set a
      int used[16] = {0};
      for (int L = 16 * half; L < 16 * half + 16; ++L) used[afp_v2_x2(c, L & 7, L >> 3) & 15]++;
      for (int i = 0; i < 16; ++i) conflicts += used[i] > 1;
    }
  }
  printf("v2 worst_rel_err %.3e coverage_errors %d conflicting_bank_pairs(a-set) %d\n", worst, bad_cover, conflicts);
  return (worst < 1e-12 && bad_cover == 0) ? 0 : 1;
}

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
  if (!(worst < 1e-12)) return 1;
  return check_v2(x, Xr, Xi);
}
