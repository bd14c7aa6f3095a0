/* TEST SHIM (see G4ShimTables.hh): cross-section per volume (1/mm in CLHEP units) = registered mass cross-section
 * (cm2/g) x density of the material; the real class evaluates the models of the physics list in use
 * (reference call sites: GateDoseActor.cpp:465-470, GateEmCalculatorActor.cpp:92) */
#pragma once
#include "G4Gamma.hh"
#include "G4Material.hh"
class G4EmCalculator {
public:
  double ComputeCrossSectionPerVolume(double kinEnergy, const G4ParticleDefinition *, const G4String &processName,
                                      const G4Material *mat, double /*cut*/ = 0.0) {
    const auto it = g4shim::tables().find(mat->ShimZ());
    if (it == g4shim::tables().end()) throw std::runtime_error("g4shim: no table for the element of " + mat->GetName());
    const double cm2_g = g4shim::interp(it->second, processName, kinEnergy / MeV);
    const double rho_g_cm3 = mat->GetDensity() / (g / cm3);
    return cm2_g * rho_g_cm3 / cm;   /* 1/cm -> per internal length unit (mm) */
  }
};
