/* G4ShimTables.hh — TEST SHIM, not Geant4.  Minimal stand-ins for the handful of Geant4 classes
 * integration/GateB200TableExporter.cpp uses, so that the exporter is compiled and executed in this repository's
 * CPU test tier (tests/test_g4_exporter_shim.py) although Geant4 is not installed.  The fake G4EmCalculator
 * answers ComputeCrossSectionPerVolume from per-element tables the test registers (log-log interpolation,
 * exact at the registered nodes); G4 units follow CLHEP (mm = MeV = 1). */
#pragma once
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

typedef std::string G4String;
typedef double G4double;

namespace g4shim {
struct ElementTable {
  std::vector<double> e_MeV;                       /* ascending */
  std::map<std::string, std::vector<double>> xs;   /* process name ("phot", "compt", "Rayl", "conv") -> cm2/g */
  double density_g_cm3 = 1.0;
};
inline std::map<int, ElementTable> &tables() {
  static std::map<int, ElementTable> t;
  return t;
}
inline double interp(const ElementTable &t, const std::string &proc, double e) {
  auto it = t.xs.find(proc);
  if (it == t.xs.end()) return 0.0;
  const std::vector<double> &x = t.e_MeV, &y = it->second;
  if (e <= x.front()) return y.front();
  if (e >= x.back()) return y.back();
  size_t lo = 0, hi = x.size() - 1;
  while (hi - lo > 1) {
    const size_t mid = (lo + hi) / 2;
    if (x[mid] <= e) lo = mid; else hi = mid;
  }
  if (x[lo] == e) return y[lo];
  if (y[lo] <= 0.0 || y[hi] <= 0.0) return y[lo] + (y[hi] - y[lo]) * (e - x[lo]) / (x[hi] - x[lo]);
  const double f = (std::log(e) - std::log(x[lo])) / (std::log(x[hi]) - std::log(x[lo]));
  return std::exp(std::log(y[lo]) + f * (std::log(y[hi]) - std::log(y[lo])));
}
}  // namespace g4shim
