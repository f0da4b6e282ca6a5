// Block Davidson eigensolver with the host control flow in C++ (xtd_davidson): the same algorithm as xtddft_b200/davidson.py --
// the reference's solver, xtddft/utils/Davidson.py:21-298 (a fork of pyscf.lib.linalg_helper.davidson1): restart at
// max_space = 12 + 4 (nroots - 1), at most 40 new vectors per cycle, `pick` of the positive eigenvalues, convergence
// |de| < tol and |r| < tol_residual, preconditioning with e[0], projection against the subspace, linear-dependency drops --
// with the vectors resident in HBM and every O(dim) operation a kernel of this library.  What the C++ driver removes is the
// per-call interpreter / ctypes overhead that dominates the small molecules (BASELINE configs 1 and 2: ~1 ms per cycle in Python
// against a 0.5 ms sigma call); four host round trips per cycle remain (projected matrix, residual norms, two Gram matrices).
#pragma once
#include <cmath>
#include <vector>

#include "common.cuh"

namespace xtd {

// ---- symmetric eigenproblem of the projected matrix (n <= ~100): Householder tridiagonalisation + implicit QL ----------------
// a: n x n row-major symmetric, overwritten with the eigenvectors (columns); w: eigenvalues ascending.
inline int sym_eig(std::vector<double>& a, int n, std::vector<double>& w) {
  std::vector<double> e(n, 0.0);
  w.assign(n, 0.0);
  auto A = [&](int i, int j) -> double& { return a[(size_t)i * n + j]; };
  // Householder reduction to tridiagonal form, accumulating the transformation
  for (int i = n - 1; i >= 1; --i) {
    const int l = i - 1;
    double h = 0.0, scale = 0.0;
    if (l > 0) {
      for (int k = 0; k <= l; ++k) scale += std::fabs(A(i, k));
      if (scale == 0.0) {
        e[i] = A(i, l);
      } else {
        for (int k = 0; k <= l; ++k) {
          A(i, k) /= scale;
          h += A(i, k) * A(i, k);
        }
        double f = A(i, l);
        double g = f >= 0.0 ? -std::sqrt(h) : std::sqrt(h);
        e[i] = scale * g;
        h -= f * g;
        A(i, l) = f - g;
        f = 0.0;
        for (int j = 0; j <= l; ++j) {
          A(j, i) = A(i, j) / h;
          g = 0.0;
          for (int k = 0; k <= j; ++k) g += A(j, k) * A(i, k);
          for (int k = j + 1; k <= l; ++k) g += A(k, j) * A(i, k);
          e[j] = g / h;
          f += e[j] * A(i, j);
        }
        const double hh = f / (h + h);
        for (int j = 0; j <= l; ++j) {
          f = A(i, j);
          e[j] = g = e[j] - hh * f;
          for (int k = 0; k <= j; ++k) A(j, k) -= f * e[k] + g * A(i, k);
        }
      }
    } else {
      e[i] = A(i, l);
    }
    w[i] = h;
  }
  w[0] = 0.0;
  e[0] = 0.0;
  for (int i = 0; i < n; ++i) {
    const int l = i - 1;
    if (w[i] != 0.0) {
      for (int j = 0; j <= l; ++j) {
        double g = 0.0;
        for (int k = 0; k <= l; ++k) g += A(i, k) * A(k, j);
        for (int k = 0; k <= l; ++k) A(k, j) -= g * A(k, i);
      }
    }
    w[i] = A(i, i);
    A(i, i) = 1.0;
    for (int j = 0; j <= l; ++j) A(j, i) = A(i, j) = 0.0;
  }
  // implicit QL on the tridiagonal matrix (diagonal w, sub-diagonal e), rotating the eigenvector matrix
  for (int i = 1; i < n; ++i) e[i - 1] = e[i];
  e[n - 1] = 0.0;
  for (int l = 0; l < n; ++l) {
    int iter = 0, m;
    do {
      for (m = l; m < n - 1; ++m) {
        const double dd = std::fabs(w[m]) + std::fabs(w[m + 1]);
        if (std::fabs(e[m]) <= 2.3e-16 * dd) break;
      }
      if (m != l) {
        if (iter++ == 60) return -1;
        double g = (w[l + 1] - w[l]) / (2.0 * e[l]);
        double r = std::hypot(g, 1.0);
        g = w[m] - w[l] + e[l] / (g + (g >= 0.0 ? std::fabs(r) : -std::fabs(r)));
        double s = 1.0, c = 1.0, p = 0.0;
        int i;
        for (i = m - 1; i >= l; --i) {
          double f = s * e[i];
          const double b = c * e[i];
          e[i + 1] = (r = std::hypot(f, g));
          if (r == 0.0) {
            w[i + 1] -= p;
            e[m] = 0.0;
            break;
          }
          s = f / r;
          c = g / r;
          g = w[i + 1] - p;
          r = (w[i] - g) * s + 2.0 * c * b;
          w[i + 1] = g + (p = s * r);
          g = c * r - b;
          for (int k = 0; k < n; ++k) {
            f = A(k, i + 1);
            A(k, i + 1) = s * A(k, i) + c * f;
            A(k, i) = c * A(k, i) - s * f;
          }
        }
        if (r == 0.0 && i >= l) continue;
        w[l] -= p;
        e[l] = g;
        e[m] = 0.0;
      }
    } while (m != l);
  }
  // ascending order
  for (int i = 0; i < n - 1; ++i) {
    int k = i;
    for (int j = i + 1; j < n; ++j)
      if (w[j] < w[k]) k = j;
    if (k != i) {
      std::swap(w[i], w[k]);
      for (int r = 0; r < n; ++r) std::swap(A(r, i), A(r, k));
    }
  }
  return 0;
}

// Gram-Schmidt carried out on the Gram matrix g = W W^T (n x n): rows of t (nk x n) give orthonormal vectors t W, taken in order;
// a vector whose squared norm after projecting out the kept ones is <= lindep is dropped (`_qr` of the reference).
inline int gs_coefficients(const double* g, int n, double lindep, std::vector<double>& t) {
  t.clear();
  int nk = 0;
  std::vector<double> c(n), gc(n);
  for (int i = 0; i < n; ++i) {
    std::fill(c.begin(), c.end(), 0.0);
    c[i] = 1.0;
    for (int r = 0; r < nk; ++r) {
      const double* tr = &t[(size_t)r * n];
      double d = 0.0;                           // c^T g t_r
      for (int a = 0; a < n; ++a) {
        double s = 0.0;
        for (int b = 0; b < n; ++b) s += g[(size_t)a * n + b] * tr[b];
        d += c[a] * s;
      }
      for (int a = 0; a < n; ++a) c[a] -= d * tr[a];
    }
    double nrm2 = 0.0;
    for (int a = 0; a < n; ++a) {
      double s = 0.0;
      for (int b = 0; b < n; ++b) s += g[(size_t)a * n + b] * c[b];
      nrm2 += c[a] * s;
    }
    if (nrm2 > lindep) {
      const double inv = 1.0 / std::sqrt(nrm2);
      for (int a = 0; a < n; ++a) t.push_back(c[a] * inv);
      ++nk;
    }
  }
  return nk;
}

__global__ void dav_rsqrt_kernel(double* __restrict__ out, const double* __restrict__ in, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = rsqrt(fmax(in[i], 1e-300));
}
__global__ void dav_negate_kernel(double* __restrict__ a, long n) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = -a[i];
}

}  // namespace xtd
