/* TEST SHIM module: the exporter as it is registered in opengate_core (init_GateB200TableExporter) plus the hook
 * through which the test fills the fake G4EmCalculator's tables. */
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include "G4ShimTables.hh"

namespace py = pybind11;
void init_GateB200TableExporter(py::module &m);

PYBIND11_MODULE(opengate_core_g4shim, m) {
  init_GateB200TableExporter(m);
  m.def("ShimRegisterElement", [](int z, double density_g_cm3, const std::vector<double> &e_MeV,
                                  const std::map<std::string, std::vector<double>> &xs_cm2_g) {
    g4shim::ElementTable t;
    t.e_MeV = e_MeV;
    t.xs = xs_cm2_g;
    t.density_g_cm3 = density_g_cm3;
    g4shim::tables()[z] = t;
  });
}
