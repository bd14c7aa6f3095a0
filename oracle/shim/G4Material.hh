// Minimal stand-in for the Geant4 header of the same name, ONLY so that the reference's own
// GateMuTables.cpp (compiled unmodified from /root/reference into oracle/_ref) finds the few
// declarations it touches.  Test infrastructure; not part of the product.
#ifndef ORACLE_SHIM_G4MATERIAL_HH
#define ORACLE_SHIM_G4MATERIAL_HH
#include <cmath>   // the real header brings the C math functions into scope
#include <string>
typedef int G4int;
typedef double G4double;
namespace CLHEP {
static const double g = 1.0, cm3 = 1.0;
}
class G4Material {
public:
  explicit G4Material(double density_g_cm3) : fDensity(density_g_cm3) {}
  double GetDensity() const { return fDensity; }   // in units where g/cm3 == 1
private:
  double fDensity;
};
#endif
