// Source time functions of the host side: STF_read / STF_get (SRC/stf_gen.f90:41-130).  The device
// consumes one amplitude per source per step (s2d_step's src_ampli table), evaluated here at
// t = it*dt - tdelay exactly where SO_add evaluates it (SRC/src_gen.f90:300-303).
#pragma once
#include <cmath>
#include <stdexcept>
#include <string>

#include "namelist.hpp"

namespace sem2d {

struct stf_type {
  enum kind_t { NONE, RICKER, GAUSSIAN, BRUNE, HARMONIC } kind = NONE;
  double f0 = 0, t0 = 0, ampli = 1;  // fc is held in f0 for BRUNE
};

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
  } else {
    throw std::runtime_error("STF_read: source time function '" + stfname + "' is not provided by this host "
                             "(RICKER, GAUSSIAN, BRUNE, HARMONIC are)");
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
    default:
      throw std::runtime_error("STF_get: unknown source time function");
  }
}

}  // namespace sem2d
