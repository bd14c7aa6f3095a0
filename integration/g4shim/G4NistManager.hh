/* TEST SHIM (see G4ShimTables.hh): elements exist when the test registered a table for them; the single-element
 * "G4_<name>" materials are built on demand with the registered density */
#pragma once
#include <memory>
#include "G4Material.hh"
class G4NistManager {
public:
  static G4NistManager *Instance() {
    static G4NistManager m;
    return &m;
  }
  const G4Element *FindOrBuildElement(int z) {
    if (!g4shim::tables().count(z)) return nullptr;
    auto &p = fElements[z];
    if (!p) p.reset(new G4Element(z));
    return p.get();
  }
  G4Material *FindOrBuildMaterial(const G4String &name) {
    for (auto &kv : fElements) {
      if ("G4_" + kv.second->GetName() != name) continue;
      auto &m = fMaterials[kv.first];
      if (!m) {
        m.reset(new G4Material(name, g4shim::tables()[kv.first].density_g_cm3 * (g / cm3), 1));
        m->AddElement(kv.second.get(), 1.0);
      }
      return m.get();
    }
    return nullptr;
  }
private:
  std::map<int, std::unique_ptr<G4Element>> fElements;
  std::map<int, std::unique_ptr<G4Material>> fMaterials;
};
