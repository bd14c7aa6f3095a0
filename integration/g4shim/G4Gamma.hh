/* TEST SHIM (see G4ShimTables.hh) */
#pragma once
class G4ParticleDefinition {};
class G4Gamma : public G4ParticleDefinition {
public:
  static G4Gamma *Gamma() {
    static G4Gamma g;
    return &g;
  }
};
