/* GateB200TableExporter.cpp — Geant4 photon cross-sections for the device tables of libgate_b200.
 *
 * Belongs in core/opengate_core/opengate_lib/ of OpenGATE/opengate (registered from opengate_core.cpp like the
 * other init_* functions) and needs Geant4 there.  In this repository it is compiled and run by the CPU test tier
 * against the stand-in headers of integration/g4shim (tests/test_g4_exporter_shim.py): the code below only uses
 * G4NistManager::FindOrBuildElement / FindOrBuildMaterial, G4Material::GetDensity / AddElement and
 * G4EmCalculator::ComputeCrossSectionPerVolume.
 *
 * Pattern followed: GateMaterialMuHandler::SimulateMaterialTable (GateMaterialMuHandler.cpp:296-549, the loop
 * over processes / models at :373-428) and the G4EmCalculator use of GateDoseActor.cpp:465-470 and
 * GateEmCalculatorActor.cpp:92.  Must be called after the run manager is initialised (the physics list owns
 * the models), i.e. from SimulationEngine after initialize_g4_verbose / RunManager::Initialize.
 *
 * Returns, per element Z, the mass cross-sections (cm2/g) of the four photon processes on the requested
 * energy grid: exactly what opengate_b200.physics.build_photon_tables expects from its `element_xs`
 * callback ('phot', 'compt', 'rayl', 'conv', 'total').
 */
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "G4Element.hh"
#include "G4EmCalculator.hh"
#include "G4Gamma.hh"
#include "G4Material.hh"
#include "G4NistManager.hh"
#include "G4SystemOfUnits.hh"

namespace py = pybind11;

namespace {

/* processes as named by the Geant4 EM constructors the reference offers (physics.py:300-337):
 * "phot" (photoelectric), "compt" (Compton), "Rayl" (Rayleigh), "conv" (gamma conversion) */
const char *const kProcess[4] = {"phot", "compt", "Rayl", "conv"};
const char *const kKey[4] = {"phot", "compt", "rayl", "conv"};

py::dict ExportElementPhotonCrossSections(int z, const std::vector<double> &energies_MeV) {
  auto *nist = G4NistManager::Instance();
  const G4Element *el = nist->FindOrBuildElement(z);
  if (el == nullptr) throw std::runtime_error("GateB200TableExporter: unknown element Z=" + std::to_string(z));
  /* a single-element material: G4EmCalculator works per material / per volume */
  const G4String mname = "G4_" + el->GetName();
  G4Material *mat = nist->FindOrBuildMaterial(mname);
  if (mat == nullptr) {
    mat = new G4Material("gate_b200_" + el->GetName(), 1.0 * g / cm3, 1);
    mat->AddElement(const_cast<G4Element *>(el), 1.0);
  }
  const double rho_g_cm3 = mat->GetDensity() / (g / cm3);
  G4EmCalculator calc;
  auto *gamma = G4Gamma::Gamma();
  const size_t n = energies_MeV.size();
  py::dict out;
  py::array_t<double> total(n);
  auto tot = total.mutable_unchecked<1>();
  for (size_t i = 0; i < n; ++i) tot(i) = 0.0;
  for (int p = 0; p < 4; ++p) {
    py::array_t<double> xs(n);
    auto a = xs.mutable_unchecked<1>();
    for (size_t i = 0; i < n; ++i) {
      /* cross-section per volume (1/length in G4 units) -> 1/cm -> cm2/g */
      const double per_cm = calc.ComputeCrossSectionPerVolume(energies_MeV[i] * MeV, gamma, kProcess[p], mat) * cm;
      a(i) = per_cm / rho_g_cm3;
      tot(i) += a(i);
    }
    out[kKey[p]] = xs;
  }
  out["total"] = total;
  return out;
}

}  // namespace

void init_GateB200TableExporter(py::module &m) {
  m.def("GateB200ExportElementPhotonCrossSections", &ExportElementPhotonCrossSections, py::arg("z"),
        py::arg("energies_MeV"),
        "Mass cross-sections (cm2/g) of phot / compt / Rayl / conv for element Z from the models of the "
        "physics list in use (G4EmCalculator); call after the run manager is initialised.");
}
