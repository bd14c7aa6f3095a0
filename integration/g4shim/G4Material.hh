/* TEST SHIM (see G4ShimTables.hh) */
#pragma once
#include "G4Element.hh"
#include "G4SystemOfUnits.hh"
class G4Material {
public:
  G4Material(const G4String &name, double density, int) : fName(name), fDensity(density) {}
  void AddElement(G4Element *el, double) { fZ = el->GetZasInt(); }
  double GetDensity() const { return fDensity; }
  int ShimZ() const { return fZ; }
  const G4String &GetName() const { return fName; }
private:
  G4String fName;
  double fDensity;
  int fZ = 0;
};
