// Minimal stand-in for the Geant4 header of the same name (see G4Material.hh in this directory).
#ifndef ORACLE_SHIM_G4MATERIALCUTSCOUPLE_HH
#define ORACLE_SHIM_G4MATERIALCUTSCOUPLE_HH
#include "G4Material.hh"
class G4MaterialCutsCouple {
public:
  explicit G4MaterialCutsCouple(const G4Material *m) : fMaterial(m) {}
  const G4Material *GetMaterial() const { return fMaterial; }
private:
  const G4Material *fMaterial;
};
#endif
