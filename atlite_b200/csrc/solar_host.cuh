// solar_host.cuh -- host-side (float64) evaluation of the time-only part of the
// Michalsky almanac the reference evaluates per cell (pv/solar_position.py:71-97):
//   tt[t] = { sin dec, cos dec, cos H0, sin H0 },  H0 = radians(lmst at lon 0) - ra
// shared by the PV and CSP operators.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cmath>
#include <vector>

namespace atl {

// {dec, H0} of one time stamp, H0 = radians(lmst at lon 0) - ra
inline void solar_almanac_step(int64_t ns, double& dec, double& H0) {
  const double PI = 3.14159265358979323846;
  const double D2R = PI / 180.0;
  const int64_t DAY = 86400LL * 1000000000LL;
  int64_t day = ns / DAY, rem = ns % DAY;
  if (rem < 0) {
    rem += DAY;
    day -= 1;
  }
  const int64_t hour = rem / 3600000000000LL;
  const int64_t minute = (rem / 60000000000LL) % 60;
  const int64_t second = (rem / 1000000000LL) % 60;
  const int64_t micro = (rem / 1000LL) % 1000000;
  const int64_t nano = rem % 1000;
  // pandas DatetimeIndex.to_julian_date: integer day count + 0.5, then + day fraction
  const double jd = ((double)day + 2440587.5) +
                    ((double)hour + (double)minute / 60.0 + (double)second / 3600.0 +
                     (double)micro / 3600.0 / 1e6 + (double)nano / 3600.0 / 1e9) /
                        24.0;
  const double n = jd - 2451545.0;                                   // :74
  const double L = 280.460 + 0.9856474 * n;                          // :86
  const double gg = (357.528 + 0.9856003 * n) * D2R;                 // :87
  const double l = (L + 1.915 * std::sin(gg) + 0.020 * std::sin(2 * gg)) * D2R;  // :88
  const double ep = (23.439 - 4e-7 * n) * D2R;                       // :89
  const double ra = std::atan2(std::cos(ep) * std::sin(l), std::cos(l));  // :91
  const double lmst0 =
      (6.697375 + ((double)hour + (double)minute / 60.0) + 0.0657098242 * n) * 15.0;  // :92
  H0 = lmst0 * D2R - ra;                                             // :95 (without lon)
  dec = std::asin(std::sin(ep) * std::sin(l));                       // :97
}

inline void solar_almanac(const int64_t* time_ns, int64_t nt, int64_t time_shift_ns,
                          std::vector<float4>& tt) {
  tt.assign((size_t)std::max<int64_t>(nt, 1), make_float4(0.f, 1.f, 1.f, 0.f));
  if (!time_ns) return;
  for (int64_t i = 0; i < nt; ++i) {
    double dec, H0;
    solar_almanac_step(time_ns[i] + time_shift_ns, dec, H0);
    tt[(size_t)i] = make_float4((float)std::sin(dec), (float)std::cos(dec), (float)std::cos(H0),
                                (float)std::sin(H0));
  }
}

// float64 tables for the materialised solar position: {sin dec, cos dec, H0} per step
inline void solar_almanac_f64(const int64_t* time_ns, int64_t nt, int64_t time_shift_ns,
                              std::vector<double>& tt) {
  tt.assign((size_t)3 * (size_t)std::max<int64_t>(nt, 1), 0.0);
  for (int64_t i = 0; i < nt; ++i) {
    double dec, H0;
    solar_almanac_step(time_ns[i] + time_shift_ns, dec, H0);
    tt[3 * (size_t)i] = std::sin(dec);
    tt[3 * (size_t)i + 1] = std::cos(dec);
    tt[3 * (size_t)i + 2] = H0;
  }
}

}  // namespace atl
