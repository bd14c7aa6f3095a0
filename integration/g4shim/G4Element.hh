/* TEST SHIM (see G4ShimTables.hh) */
#pragma once
#include "G4ShimTables.hh"
class G4Element {
public:
  explicit G4Element(int z) : fZ(z), fName("Z" + std::to_string(z)) {}
  const G4String &GetName() const { return fName; }
  int GetZasInt() const { return fZ; }
private:
  int fZ;
  G4String fName;
};
