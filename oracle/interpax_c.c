/* interpax_c.c -- plain-C restatement of the reference's tricubic / bicubic / cubic
 * evaluation, multi-threaded with OpenMP.  TEST INFRASTRUCTURE ONLY: it is the CPU
 * baseline timed by bench.py (cpu_baseline / --impl reference) and a second checker
 * for the CUDA path.  The product (interpax_b200/) never links or calls it.
 *
 * It follows the reference's algorithm as written, not the separable shortcut the GPU
 * kernels use (paths relative to the reference repo):
 *   search        clip(searchsorted(x, q, "right"), 1, n-1)      interpax/_spline.py:1051-1053
 *   local coord   dx, dxi = where(dx==0, 0, 1/dx), t = delta*dxi  :1055-1068
 *   gather        64 corner values, slope tables scaled by dx/dy/dz :1070-1091
 *   contraction   coef = A_TRICUBIC @ F  (dense 64 x 64)           :1093-1095, _coefs.py:8-163
 *   evaluation    sum coef[a,b,c] ttx[a] tty[b] ttz[c]             :1096-1099, _get_t_der :1138-1151
 *   extrap        per axis post-mask                               :1101-1103, :1180-1222
 * The periodic case is evaluated on tables the caller has already padded (the Python
 * side does _make_periodic exactly as the NumPy oracle does).
 *
 * Parity status: validated against oracle/interpax_np.py (tests/test_oracle_c.py), which
 * carries the pins listed in its header; the reference itself cannot run here (no JAX).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static const int A_CUBIC[4][4] = {{1, 0, 0, 0}, {0, 0, 1, 0}, {-3, 3, -2, -1}, {2, -2, 1, 1}};

/* dense matrices, built once as tensor powers in the reference's index order
 * (tests/test_oracle_pins.py checks the same construction element-wise against
 * interpax/_coefs.py) */
static double A3[64][64];
static double A2[16][16];
static int built = 0;

static void build_matrices(void) {
  if (built) return;
  /* 3-D table order f, fx, fy, fz, fxy, fxz, fyz, fxyz  (_spline.py:1070-1078) */
  static const int dxyz[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1},
                                 {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
  for (int ff = 0; ff < 8; ++ff)
    for (int kk = 0; kk < 2; ++kk)
      for (int jj = 0; jj < 2; ++jj)
        for (int ii = 0; ii < 2; ++ii) {
          int n = 8 * ff + 4 * kk + 2 * jj + ii;
          for (int r = 0; r < 4; ++r)
            for (int q = 0; q < 4; ++q)
              for (int p = 0; p < 4; ++p)
                A3[p + 4 * q + 16 * r][n] = (double)(A_CUBIC[p][2 * dxyz[ff][0] + ii] *
                                                     A_CUBIC[q][2 * dxyz[ff][1] + jj] *
                                                     A_CUBIC[r][2 * dxyz[ff][2] + kk]);
        }
  /* 2-D table order f, fx, fy, fxy  (_spline.py:779-783) */
  static const int dxy[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
  for (int ff = 0; ff < 4; ++ff)
    for (int jj = 0; jj < 2; ++jj)
      for (int ii = 0; ii < 2; ++ii) {
        int n = 4 * ff + 2 * jj + ii;
        for (int q = 0; q < 4; ++q)
          for (int p = 0; p < 4; ++p)
            A2[p + 4 * q][n] =
                (double)(A_CUBIC[p][2 * dxy[ff][0] + ii] * A_CUBIC[q][2 * dxy[ff][1] + jj]);
      }
  built = 1;
}

/* torchrun exports OMP_NUM_THREADS=1; the baseline wants every host thread it can use */
void ipxo_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int ipxo_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* searchsorted(side="right") then clip to [1, n-1]; NaN sorts last */
static inline int bin_index(const double* x, int n, double q) {
  int lo = 0, hi = n;
  if (q != q) return n - 1;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (x[mid] <= q) lo = mid + 1; else hi = mid;
  }
  if (lo < 1) lo = 1;
  if (lo > n - 1) lo = n - 1;
  return lo;
}

/* _get_t_der (_spline.py:1138-1151) */
static inline void t_der(double t, int order, double dxi, double tt[4]) {
  if (order <= 0) { tt[0] = 1; tt[1] = t; tt[2] = t * t; tt[3] = t * t * t; }
  else if (order == 1) { tt[0] = 0; tt[1] = dxi; tt[2] = 2 * t * dxi; tt[3] = 3 * (t * t) * dxi; }
  else if (order == 2) { double s = dxi * dxi; tt[0] = 0; tt[1] = 0; tt[2] = 2 * s; tt[3] = 6 * t * s; }
  else if (order == 3) { tt[0] = 0; tt[1] = 0; tt[2] = 0; tt[3] = 6 * (dxi * dxi * dxi); }
  else { double z = dxi * 0; tt[0] = z; tt[1] = z; tt[2] = z; tt[3] = z; }
}

typedef struct {
  int keep_lo, keep_hi;
  double fill_lo, fill_hi;
} ipxo_extrap;

static inline double apply_extrap(double v, double q, const double* x, int n, const ipxo_extrap* e) {
  if (!e->keep_lo && q < x[0]) v = e->fill_lo;
  if (!e->keep_hi && q > x[n - 1]) v = e->fill_hi;
  return v;
}

/* interp3d, cubic family, scalar-valued f (batch = 1), tables given (8 pointers in the
 * reference's order, C-order nx*ny*nz). */
void ipxo_interp3d_cubic(const double* xq, const double* yq, const double* zq, int64_t nq,
                         const double* x, int nx, const double* y, int ny, const double* z, int nz,
                         const double* const* tab, const int* deriv, const ipxo_extrap* ex,
                         double* out) {
  build_matrices();
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < nq; ++l) {
    int i = bin_index(x, nx, xq[l]), j = bin_index(y, ny, yq[l]), k = bin_index(z, nz, zq[l]);
    double dx = x[i] - x[i - 1], dy = y[j] - y[j - 1], dz = z[k] - z[k - 1];
    double dxi = dx == 0 ? 0 : 1 / dx, dyi = dy == 0 ? 0 : 1 / dy, dzi = dz == 0 ? 0 : 1 / dz;
    double tx = (xq[l] - x[i - 1]) * dxi, ty = (yq[l] - y[j - 1]) * dyi, tz = (zq[l] - z[k - 1]) * dzi;
    static const int dxyz[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1},
                                   {1, 1, 0}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
    double F[64];
    for (int ff = 0; ff < 8; ++ff)
      for (int kk = 0; kk < 2; ++kk)
        for (int jj = 0; jj < 2; ++jj)
          for (int ii = 0; ii < 2; ++ii) {
            double v = tab[ff][((int64_t)(i - 1 + ii) * ny + (j - 1 + jj)) * nz + (k - 1 + kk)];
            if (dxyz[ff][0]) v = dx * v;
            if (dxyz[ff][1]) v = dy * v;
            if (dxyz[ff][2]) v = dz * v;
            F[8 * ff + 4 * kk + 2 * jj + ii] = v;
          }
    double coef[64];
    for (int m = 0; m < 64; ++m) {
      double s = 0;
      for (int n = 0; n < 64; ++n) s += A3[m][n] * F[n];
      coef[m] = s;
    }
    double ttx[4], tty[4], ttz[4];
    t_der(tx, deriv[0], dxi, ttx);
    t_der(ty, deriv[1], dyi, tty);
    t_der(tz, deriv[2], dzi, ttz);
    double fq = 0;
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b)
        for (int c = 0; c < 4; ++c) fq += coef[a + 4 * b + 16 * c] * ttx[a] * tty[b] * ttz[c];
    fq = apply_extrap(fq, xq[l], x, nx, &ex[0]);
    fq = apply_extrap(fq, yq[l], y, ny, &ex[1]);
    fq = apply_extrap(fq, zq[l], z, nz, &ex[2]);
    out[l] = fq;
  }
}

/* interp2d, cubic family (_spline.py:757-805) */
void ipxo_interp2d_cubic(const double* xq, const double* yq, int64_t nq, const double* x, int nx,
                         const double* y, int ny, const double* const* tab, const int* deriv,
                         const ipxo_extrap* ex, double* out) {
  build_matrices();
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < nq; ++l) {
    int i = bin_index(x, nx, xq[l]), j = bin_index(y, ny, yq[l]);
    double dx = x[i] - x[i - 1], dy = y[j] - y[j - 1];
    double dxi = dx == 0 ? 0 : 1 / dx, dyi = dy == 0 ? 0 : 1 / dy;
    double tx = (xq[l] - x[i - 1]) * dxi, ty = (yq[l] - y[j - 1]) * dyi;
    static const int dxy[4][2] = {{0, 0}, {1, 0}, {0, 1}, {1, 1}};
    double F[16];
    for (int ff = 0; ff < 4; ++ff)
      for (int jj = 0; jj < 2; ++jj)
        for (int ii = 0; ii < 2; ++ii) {
          double v = tab[ff][(int64_t)(i - 1 + ii) * ny + (j - 1 + jj)];
          if (dxy[ff][0]) v = dx * v;
          if (dxy[ff][1]) v = dy * v;
          F[4 * ff + 2 * jj + ii] = v;
        }
    double coef[16];
    for (int m = 0; m < 16; ++m) {
      double s = 0;
      for (int n = 0; n < 16; ++n) s += A2[m][n] * F[n];
      coef[m] = s;
    }
    double ttx[4], tty[4];
    t_der(tx, deriv[0], dxi, ttx);
    t_der(ty, deriv[1], dyi, tty);
    double fq = 0;
    for (int a = 0; a < 4; ++a)
      for (int b = 0; b < 4; ++b) fq += coef[a + 4 * b] * ttx[a] * tty[b];
    fq = apply_extrap(fq, xq[l], x, nx, &ex[0]);
    fq = apply_extrap(fq, yq[l], y, ny, &ex[1]);
    out[l] = fq;
  }
}

/* interp1d, cubic family (_spline.py:569-592) */
void ipxo_interp1d_cubic(const double* xq, int64_t nq, const double* x, int nx,
                         const double* const* tab, const int* deriv, const ipxo_extrap* ex,
                         double* out) {
#pragma omp parallel for schedule(static)
  for (int64_t l = 0; l < nq; ++l) {
    int i = bin_index(x, nx, xq[l]);
    double dx = x[i] - x[i - 1];
    double dxi = dx == 0 ? 0 : 1 / dx;
    double t = (xq[l] - x[i - 1]) * dxi;
    double F[4] = {tab[0][i - 1], tab[0][i], tab[1][i - 1] * dx, tab[1][i] * dx};
    double tt[4];
    t_der(t, deriv[0], dxi, tt);
    double fq = 0;
    for (int a = 0; a < 4; ++a) {
      double c = 0;
      for (int n = 0; n < 4; ++n) c += A_CUBIC[a][n] * F[n];
      fq += c * tt[a];
    }
    out[l] = apply_extrap(fq, xq[l], x, nx, &ex[0]);
  }
}
