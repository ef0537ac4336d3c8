// Source time functions of the host side: STF_read / STF_get (SRC/stf_gen.f90:41-130).  The device
// consumes one amplitude per source per step (s2d_step's src_ampli table), evaluated here at
// t = it*dt - tdelay exactly where SO_add evaluates it (SRC/src_gen.f90:300-303).
#pragma once
#include <cmath>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "namelist.hpp"

namespace sem2d {

struct stf_type {
  enum kind_t { NONE, RICKER, GAUSSIAN, BRUNE, HARMONIC, TAB, USER } kind = NONE;
  double f0 = 0, t0 = 0, ampli = 1;  // fc is held in f0 for BRUNE
  // STF_TAB_type (stf_tabulated.f90:8-11): samples and the second derivatives of their natural cubic spline
  std::vector<double> tab_t, tab_v, tab_v2;
  // STF_USER_type (stf_user.f90:20-24): ampli, onset (above), par1, par2, ipar1, ipar2
  double par1 = 0, par2 = 0;
  int ipar1 = 0, ipar2 = 0;
};

// spline (SRC/utils.f90:652-691, Numerical Recipes): second derivatives of the interpolating cubic spline;
// yp1, ypn > 0.99e30 would ask for a natural end, STF_TAB_read passes 0d0 (zero slope at both ends)
inline void spline(const std::vector<double>& x, const std::vector<double>& y, double yp1, double ypn, std::vector<double>& y2) {
  const int n = (int)x.size();
  y2.assign(n, 0.0);
  std::vector<double> u(n, 0.0);
  if (yp1 > .99e30) {
    y2[0] = 0.0;
    u[0] = 0.0;
  } else {
    y2[0] = -0.5;
    u[0] = (3.0 / (x[1] - x[0])) * ((y[1] - y[0]) / (x[1] - x[0]) - yp1);
  }
  for (int i = 1; i < n - 1; ++i) {
    const double sig = (x[i] - x[i - 1]) / (x[i + 1] - x[i - 1]);
    const double p = sig * y2[i - 1] + 2.0;
    y2[i] = (sig - 1.0) / p;
    u[i] = (6.0 * ((y[i + 1] - y[i]) / (x[i + 1] - x[i]) - (y[i] - y[i - 1]) / (x[i] - x[i - 1])) / (x[i + 1] - x[i - 1]) - sig * u[i - 1]) / p;
  }
  double qn, un;
  if (ypn > (double).99e30f) {
    qn = 0.0;
    un = 0.0;
  } else {
    qn = 0.5;
    un = (3.0 / (x[n - 1] - x[n - 2])) * (ypn - (y[n - 1] - y[n - 2]) / (x[n - 1] - x[n - 2]));
  }
  y2[n - 1] = (un - qn * u[n - 2]) / (qn * y2[n - 2] + 1.0);
  for (int k = n - 2; k >= 0; --k) y2[k] = y2[k] * y2[k + 1] + u[k];
}

// hunt (SRC/utils.f90:596-646): the interval of x in the ascending table xx, searched from the previous answer;
// jlo is 1-based and kept between calls (splint's `integer, save :: klo = 1`, shared by every tabulated function)
inline void hunt(const std::vector<double>& xx, double x, int& jlo) {
  const int n = (int)xx.size();
  auto X = [&](int k) { return xx[(size_t)k - 1]; };
  const bool ascnd = X(n) > X(1);
  int jhi;
  if (jlo <= 0 || jlo > n) {
    jlo = 0;
    jhi = n + 1;
  } else {
    int inc = 1;
    if ((x >= X(jlo)) == ascnd) {
      for (;;) {
        jhi = jlo + inc;
        if (jhi > n) {
          jhi = n + 1;
          break;
        } else if ((x >= X(jhi)) == ascnd) {
          jlo = jhi;
          inc += inc;
        } else {
          break;
        }
      }
    } else {
      jhi = jlo;
      for (;;) {
        jlo = jhi - inc;
        if (jlo < 1) {
          jlo = 0;
          break;
        } else if ((x < X(jlo)) == ascnd) {
          jhi = jlo;
          inc += inc;
        } else {
          break;
        }
      }
    }
  }
  while (jhi - jlo != 1) {
    const int jm = (jhi + jlo) / 2;
    if ((x > X(jm)) == ascnd) jlo = jm;
    else jhi = jm;
  }
}

// splint (SRC/utils.f90:697-733)
inline double splint(const std::vector<double>& xa, const std::vector<double>& ya, const std::vector<double>& y2a, double x) {
  static int klo = 1;
  hunt(xa, x, klo);
  if (klo < 1) klo = 1;                       // the reference would index xa(0) here; x is clamped to the table
  if (klo > (int)xa.size() - 1) klo = (int)xa.size() - 1;
  const int khi = klo + 1;
  const double h = xa[khi - 1] - xa[klo - 1];
  const double a = (xa[khi - 1] - x) / h, b = (x - xa[klo - 1]) / h;
  return a * ya[klo - 1] + b * ya[khi - 1] + ((a * a * a - a) * y2a[klo - 1] + (b * b * b - b) * y2a[khi - 1]) * (h * h) / 6.0;
}

// STF_read: the parameter block follows the &SRC_DEF record (forward scan)
inline stf_type STF_read(const std::string& stfname, const namelist_file& in, size_t from) {
  static const double PI = 3.141592653589793238462643383279502884197;
  (void)PI;
  stf_type s;
  auto block = [&](const char* name) -> const nml_group& {
    const long k = in.find(name, from);
    if (k < 0) throw std::runtime_error(std::string(name) + " input block not found");
    return in.at((size_t)k);
  };
  if (stfname == "RICKER") {  // f0, onset, ampli are default REALs widened by dble() (stf_ricker.f90:53,71-73)
    const nml_group& g = block("STF_RICKER");
    s.kind = stf_type::RICKER;
    s.f0 = g.real4("f0", 0.0);
    s.t0 = g.real4("onset", 0.0);
    s.ampli = g.real4("ampli", 1.0);
    if (!(s.f0 > 0.0)) throw std::runtime_error("RICKER_read: f0 must be positive");
  } else if (stfname == "GAUSSIAN") {  // stf_gaussian.f90:43-53
    const nml_group& g = block("STF_GAUSSIAN");
    s.kind = stf_type::GAUSSIAN;
    s.f0 = g.real8("f0", 1.0);
    s.t0 = g.real8("onset", 0.0);
    s.ampli = g.real8("ampli", 1.0);
  } else if (stfname == "BRUNE") {  // stf_brune.f90:28-36
    const nml_group& g = block("STF_BRUNE");
    s.kind = stf_type::BRUNE;
    s.ampli = g.real8("ampli", 1.0);
    s.f0 = g.real8("fc", 1.0);
  } else if (stfname == "HARMONIC") {  // stf_harmonic.f90:27-36
    const nml_group& g = block("STF_HARMONIC");
    s.kind = stf_type::HARMONIC;
    s.ampli = g.real8("ampli", 0.0);
    s.f0 = g.real8("f0", 0.0);
    if (!(s.f0 > 0.0)) throw std::runtime_error("STF_HARMONIC_read: f0 must be positive");
    if (s.ampli == 0.0) throw std::runtime_error("STF_HARMONIC_read: ampli must be non zero");
  } else if (stfname == "TAB") {  // STF_TAB_read (stf_tabulated.f90:41-68): two columns t, v; spline with zero end slopes
    const long k = in.find("STF_TAB", from);
    const std::string file = k >= 0 ? in.at((size_t)k).text("file", "stf.tab") : std::string("stf.tab");
    std::ifstream f(file);
    if (!f) throw std::runtime_error("STF_TAB_read: cannot open " + file);
    s.kind = stf_type::TAB;
    std::string line;
    while (std::getline(f, line)) {  // IO_file_length counts the records, each holds t and v
      for (char& c : line)
        if (c == 'd' || c == 'D') c = 'e';
      std::istringstream ls(line);
      double t, v;
      if (ls >> t >> v) {
        s.tab_t.push_back(t);
        s.tab_v.push_back(v);
      }
    }
    if (s.tab_t.size() < 2) throw std::runtime_error("STF_TAB_read: " + file + " needs at least two samples");
    spline(s.tab_t, s.tab_v, 0.0, 0.0, s.tab_v2);
  } else if (stfname == "USER") {  // STF_USER_read (stf_user.f90:33-60): default REALs widened
    const long k = in.find("STF_USER", from);
    if (k < 0) throw std::runtime_error("STF_USER_read: input block STF_USER not found");
    const nml_group& g = in.at((size_t)k);
    s.kind = stf_type::USER;
    s.t0 = g.real4("onset", 0.0);
    s.ampli = g.real4("ampli", 1.0);
    s.par1 = g.real4("par1", 0.0);
    s.par2 = g.real4("par2", 0.0);
    s.ipar1 = g.integer("ipar1", 0);
    s.ipar2 = g.integer("ipar2", 0);
    if (s.ipar1 < 0) throw std::runtime_error("STF_USER_read: ipar1 must be positive");
  } else if (stfname == "BUTTERWORTH") {  // butterworth_filter.f90:32-42: the reference itself stops here
    throw std::runtime_error("BUTTER_read: not implemented");
  } else {
    throw std::runtime_error("STF_read: unknown source time function '" + stfname + "' (stf_gen.f90:58-81)");
  }
  return s;
}

inline double STF_get(const stf_type& s, double t) {
  static const double PI = 3.141592653589793238462643383279502884197;
  switch (s.kind) {
    case stf_type::RICKER: {  // stf_ricker.f90:89-101
      double arg = PI * s.f0 * (t - s.t0);
      arg = arg * arg;
      return -s.ampli * (1.0 - 2.0 * arg) * std::exp(-arg);
    }
    case stf_type::GAUSSIAN: {  // stf_gaussian.f90:80-82
      double arg = PI * s.f0 * (t - s.t0);
      arg = arg * arg;
      return s.ampli * std::exp(-arg);
    }
    case stf_type::BRUNE: {  // stf_brune.f90:58-59
      const double arg = 2 * PI * s.f0 * std::fmax(t, 0.0);
      return s.ampli * (1.0 - (1.0 + arg) * std::exp(-arg));
    }
    case stf_type::HARMONIC:  // stf_harmonic.f90:59
      return s.ampli * std::sin(2.0 * PI * t * s.f0);
    case stf_type::TAB: {  // STF_TAB_fun (stf_tabulated.f90:76-93): clamped to the table, cubic spline
      double tb = std::fmax(t, s.tab_t.front());
      tb = std::fmin(tb, s.tab_t.back());
      return splint(s.tab_t, s.tab_v, s.tab_v2, tb);
    }
    case stf_type::USER: {  // STF_USER_fun (stf_user.f90:66-79): the template function the reference ships
      const double arg = t - s.t0;
      return s.ampli * std::sin(arg) + s.par1 * (arg * arg);
    }
    default:
      throw std::runtime_error("STF_get: unknown source time function");
  }
}

}  // namespace sem2d
