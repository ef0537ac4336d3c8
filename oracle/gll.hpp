// ORACLE (test infrastructure only -- never linked into the product library).
// CPU restatement of the reference's Gauss-Lobatto-Legendre library.
// Follows /root/reference/SRC/gll.f90 (get_GLL_info :19-36, endw1/endw2, gammaf,
// hdgll :231-265, hgll :271-295, jacg :300-380, jacobf, pndleg, pnleg, pnormj,
// zwgjd, zwgljd :622-684).  Arithmetic order is kept statement by statement.
#pragma once
#include <cmath>
#include <vector>
#include <stdexcept>

namespace orc {
namespace gll {

inline double gammaf(double x) {
  const double pi = 3.141592653589793;
  double g = 1.0;
  if (x == -0.5) g = -2.0 * std::sqrt(pi);
  if (x == 0.5) g = std::sqrt(pi);
  if (x == 1.0) g = 1.0;
  if (x == 2.0) g = 1.0;
  if (x == 1.5) g = std::sqrt(pi) / 2.0;
  if (x == 2.5) g = 1.5 * std::sqrt(pi) / 2.0;
  if (x == 3.5) g = 2.5 * 1.5 * std::sqrt(pi) / 2.0;
  if (x == 3.0) g = 2.0;
  if (x == 4.0) g = 6.0;
  if (x == 5.0) g = 24.0;
  if (x == 6.0) g = 120.0;
  return g;
}

inline double pnormj(int n, double alpha, double beta) {
  const double one = 1.0, two = 2.0;
  double dn = (double)n;
  double cnst = alpha + beta + one;
  double prod;
  if (n <= 1) {
    prod = gammaf(dn + alpha) * gammaf(dn + beta);
    prod = prod / (gammaf(dn) * gammaf(dn + alpha + beta));
    return prod * std::pow(two, cnst) / (two * dn + cnst);
  }
  prod = gammaf(alpha + one) * gammaf(beta + one);
  prod = prod / (two * (one + cnst) * gammaf(cnst + one));
  prod = prod * (one + alpha) * (two + alpha);
  prod = prod * (one + beta) * (two + beta);
  for (int i = 3; i <= n; ++i) {
    double dindx = (double)i;
    double frac = (dindx + alpha) * (dindx + beta) / (dindx * (dindx + alpha + beta));
    prod = prod * frac;
  }
  return prod * std::pow(two, cnst) / (two * dn + cnst);
}

// jacobf: Jacobi polynomial of degree n and derivative at x (+ degree n-1, n-2)
inline void jacobf(double& poly, double& pder, double& polym1, double& pderm1, double& polym2,
                   double& pderm2, int n, double alp, double bet, double x) {
  double apb = alp + bet;
  poly = 1.0;
  pder = 0.0;
  double psave = 0.0, pdsave = 0.0;
  if (n == 0) return;
  double polyl = poly, pderl = pder;
  poly = (alp - bet + (apb + 2.0) * x) / 2.0;
  pder = (apb + 2.0) / 2.0;
  if (n == 1) return;
  for (int k = 2; k <= n; ++k) {
    double dk = (double)k;
    double a1 = 2.0 * dk * (dk + apb) * (2.0 * dk + apb - 2.0);
    double a2 = (2.0 * dk + apb - 1.0) * (alp * alp - bet * bet);
    double b3 = (2.0 * dk + apb - 2.0);
    double a3 = b3 * (b3 + 1.0) * (b3 + 2.0);
    double a4 = 2.0 * (dk + alp - 1.0) * (dk + bet - 1.0) * (2.0 * dk + apb);
    double polyn = ((a2 + a3 * x) * poly - a4 * polyl) / a1;
    double pdern = ((a2 + a3 * x) * pder - a4 * pderl + a3 * poly) / a1;
    psave = polyl;
    pdsave = pderl;
    polyl = poly;
    poly = polyn;
    pderl = pder;
    pder = pdern;
  }
  polym1 = polyl;
  pderm1 = pderl;
  polym2 = psave;
  pderm2 = pdsave;
}

// jacg: np Gauss-Jacobi points (1-based semantics mapped on xjac[0..np-1])
inline void jacg(double* xjac, int np, double alpha, double beta) {
  const int kstop = 10;
  const double eps = 1.0e-12;
  double pm1 = 0, pm2 = 0, pdm1 = 0, pdm2 = 0, xlast = 0, p = 0, pd = 0;
  int n = np - 1;
  double dth = 4.0 * std::atan(1.0) / (2.0 * (double)n + 2.0);
  double x = 0;
  for (int j = 1; j <= np; ++j) {
    if (j == 1) {
      x = std::cos((2.0 * ((double)j - 1.0) + 1.0) * dth);
    } else {
      double x1 = std::cos((2.0 * ((double)j - 1.0) + 1.0) * dth);
      double x2 = xlast;
      x = (x1 + x2) / 2.0;
    }
    for (int k = 1; k <= kstop; ++k) {
      jacobf(p, pd, pm1, pdm1, pm2, pdm2, np, alpha, beta, x);
      double recsum = 0.0;
      int jm = j - 1;
      for (int i = 1; i <= jm; ++i) recsum = recsum + 1.0 / (x - xjac[np - i + 1 - 1]);
      double delx = -p / (pd - recsum * p);
      x = x + delx;
      if (std::fabs(delx) < eps) break;
    }
    xjac[np - j + 1 - 1] = x;
    xlast = x;
  }
  int jmin = 0;
  for (int i = 1; i <= np; ++i) {
    double xmin = 2.0;
    for (int j = i; j <= np; ++j) {
      if (xjac[j - 1] < xmin) {
        xmin = xjac[j - 1];
        jmin = j;
      }
    }
    if (jmin != i) {
      double swap = xjac[i - 1];
      xjac[i - 1] = xjac[jmin - 1];
      xjac[jmin - 1] = swap;
    }
  }
}

inline void zwgjd(double* z, double* w, int np, double alpha, double beta) {
  const double one = 1.0, two = 2.0;
  double p = 0, pd = 0, pm1 = 0, pdm1 = 0, pm2 = 0, pdm2 = 0;
  int n = np - 1;
  double apb = alpha + beta;
  if (np <= 0) throw std::runtime_error("Minimum number of Gauss points is 1");
  if (np == 1) {
    z[0] = (beta - alpha) / (apb + two);
    w[0] = gammaf(alpha + one) * gammaf(beta + one) / gammaf(apb + two) * std::pow(two, apb + one);
    return;
  }
  jacg(z, np, alpha, beta);
  int np1 = n + 1, np2 = n + 2;
  double dnp1 = (double)np1, dnp2 = (double)np2;
  double fac1 = dnp1 + alpha + beta + one;
  double fac2 = fac1 + dnp1;
  double fac3 = fac2 + one;
  double fnorm = pnormj(np1, alpha, beta);
  double rcoef = (fnorm * fac2 * fac3) / (two * fac1 * dnp2);
  for (int i = 0; i < np; ++i) {
    jacobf(p, pd, pm1, pdm1, pm2, pdm2, np2, alpha, beta, z[i]);
    w[i] = -rcoef / (p * pdm1);
  }
}

inline double endw1(int n, double alpha, double beta) {
  const double zero = 0, one = 1, two = 2, three = 3, four = 4;
  double f3 = zero;
  double apb = alpha + beta;
  if (n == 0) return zero;
  double f1 = gammaf(alpha + two) * gammaf(beta + one) / gammaf(apb + three);
  f1 = f1 * (apb + two) * std::pow(two, apb + two) / two;
  if (n == 1) return f1;
  double fint1 = gammaf(alpha + two) * gammaf(beta + one) / gammaf(apb + three);
  fint1 = fint1 * std::pow(two, apb + two);
  double fint2 = gammaf(alpha + two) * gammaf(beta + two) / gammaf(apb + four);
  fint2 = fint2 * std::pow(two, apb + three);
  double f2 = (-two * (beta + two) * fint1 + (apb + four) * fint2) * (apb + three) / four;
  if (n == 2) return f2;
  for (int i = 3; i <= n; ++i) {
    double di = (double)(i - 1);
    double abn = alpha + beta + di;
    double abnn = abn + di;
    double a1 = -(two * (di + alpha) * (di + beta)) / (abn * abnn * (abnn + one));
    double a2 = (two * (alpha - beta)) / (abnn * (abnn + two));
    double a3 = (two * (abn + one)) / ((abnn + two) * (abnn + one));
    f3 = -(a2 * f2 + a1 * f1) / a3;
    f1 = f2;
    f2 = f3;
  }
  return f3;
}

inline double endw2(int n, double alpha, double beta) {
  const double zero = 0, one = 1, two = 2, three = 3, four = 4;
  double apb = alpha + beta;
  double f3 = zero;
  if (n == 0) return zero;
  double f1 = gammaf(alpha + one) * gammaf(beta + two) / gammaf(apb + three);
  f1 = f1 * (apb + two) * std::pow(two, apb + two) / two;
  if (n == 1) return f1;
  double fint1 = gammaf(alpha + one) * gammaf(beta + two) / gammaf(apb + three);
  fint1 = fint1 * std::pow(two, apb + two);
  double fint2 = gammaf(alpha + two) * gammaf(beta + two) / gammaf(apb + four);
  fint2 = fint2 * std::pow(two, apb + three);
  double f2 = (two * (alpha + two) * fint1 - (apb + four) * fint2) * (apb + three) / four;
  if (n == 2) return f2;
  for (int i = 3; i <= n; ++i) {
    double di = (double)(i - 1);
    double abn = alpha + beta + di;
    double abnn = abn + di;
    double a1 = -(two * (di + alpha) * (di + beta)) / (abn * abnn * (abnn + one));
    double a2 = (two * (alpha - beta)) / (abnn * (abnn + two));
    double a3 = (two * (abn + one)) / ((abnn + two) * (abnn + one));
    f3 = -(a2 * f2 + a1 * f1) / a3;
    f1 = f2;
    f2 = f3;
  }
  return f3;
}

inline void zwgljd(double* z, double* w, int np, double alpha, double beta) {
  const double one = 1.0, two = 2.0;
  double p = 0, pd = 0, pm1 = 0, pdm1 = 0, pm2 = 0, pdm2 = 0;
  int n = np - 1, nm1 = n - 1;
  if (np <= 1) throw std::runtime_error("Minimum number of Gauss-Lobatto points is 2");
  if (nm1 > 0) {
    double alpg = alpha + one, betg = beta + one;
    zwgjd(z + 1, w + 1, nm1, alpg, betg);
  }
  z[0] = -one;
  z[np - 1] = one;
  for (int i = 1; i <= np - 2; ++i) w[i] = w[i] / (one - z[i] * z[i]);
  jacobf(p, pd, pm1, pdm1, pm2, pdm2, n, alpha, beta, z[0]);
  w[0] = endw1(n, alpha, beta) / (two * pd);
  jacobf(p, pd, pm1, pdm1, pm2, pdm2, n, alpha, beta, z[np - 1]);
  w[np - 1] = endw2(n, alpha, beta) / (two * pd);
}

inline double pnleg(double z, int n) {
  double p1 = 1.0, p2 = z, p3 = p2;
  for (int k = 1; k <= n - 1; ++k) {
    double fk = (double)k;
    p3 = ((2.0 * fk + 1.0) * z * p2 - fk * p1) / (fk + 1.0);
    p1 = p2;
    p2 = p3;
  }
  return p3;
}

inline double pndleg(double z, int n) {
  double p1 = 1.0, p2 = z, p1d = 0.0, p2d = 1.0, p3d = 1.0, p3;
  for (int k = 1; k <= n - 1; ++k) {
    double fk = (double)k;
    p3 = ((2.0 * fk + 1.0) * z * p2 - fk * p1) / (fk + 1.0);
    p3d = ((2.0 * fk + 1.0) * p2 + (2.0 * fk + 1.0) * z * p2d - fk * p1d) / (fk + 1.0);
    p1 = p2;
    p2 = p3;
    p1d = p2d;
    p2d = p3d;
  }
  return p3d;
}

// hdgll(i,j): derivative of Lagrange interpolant i at GLL point j (0-based indices)
inline double hdgll(int i, int j, const double* zgll, int nz) {
  int idegpoly = nz - 1;
  double dn = (double)idegpoly;
  if (i == 0 && j == 0) return -dn * (dn + 1.0) / 4.0;
  if (i == idegpoly && j == idegpoly) return dn * (dn + 1.0) / 4.0;
  if (i == j) return 0.0;
  double rl1 = pnleg(zgll[j], idegpoly);
  double rl2 = pndleg(zgll[j], idegpoly);
  double rl3 = pnleg(zgll[i], idegpoly);
  return rl1 / (rl3 * (zgll[j] - zgll[i])) +
         (1.0 - zgll[j] * zgll[j]) * rl2 /
             (dn * (dn + 1.0) * rl3 * (zgll[j] - zgll[i]) * (zgll[j] - zgll[i]));
}

// hgll: Lagrange interpolant i (0-based) at z
inline double hgll(int i, double z, const double* zgll, int nz) {
  const double eps = 1.0e-5;
  double dz = z - zgll[i];
  if (std::fabs(dz) < eps) return 1.0;
  int n = nz - 1;
  double alfan = (double)n * ((double)n + 1.0);
  return -(1.0 - z * z) * pndleg(z, n) / (alfan * pnleg(zgll[i], n) * (z - zgll[i]));
}

// get_GLL_info: x, w, H(ip,ix) = h'_ip(x_ix), stored column-major H[ip + n*ix]
inline void get_GLL_info(int n, std::vector<double>& x, std::vector<double>& w,
                         std::vector<double>& H) {
  x.assign(n, 0.0);
  w.assign(n, 0.0);
  H.assign((size_t)n * n, 0.0);
  zwgljd(x.data(), w.data(), n, 0.0, 0.0);
  if (n % 2 != 0) x[(n - 1) / 2] = 0.0;
  for (int ix = 0; ix < n; ++ix)
    for (int ip = 0; ip < n; ++ip) H[ip + (size_t)n * ix] = hdgll(ip, ix, x.data(), n);
}

}  // namespace gll
}  // namespace orc
