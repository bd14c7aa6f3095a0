/* pyGateB200Engine.cpp — pybind11 binding of the C-ABI of libgate_b200 (include/gate_b200.h), written in the
 * idiom of core/opengate_core/opengate_lib/py*.cpp.  NOT built in this repository (the ctypes binding
 * opengate_b200/core.py is what the tests use); a maintainer adds it to core/opengate_core, registers
 * init_GateB200Engine(m) in opengate_core.cpp and links libgate_b200.so.
 *
 * The blocking calls release the GIL exactly as StartMasterThread does (pyGateSourceManager.cpp:36-38). */
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <array>
#include <stdexcept>
#include <string>

#include "gate_b200.h"

namespace py = pybind11;

namespace {

class GateB200Engine {
public:
  GateB200Engine(int device, uint64_t seed) {
    const int rc = gate_b200_create(device, seed, &h_);
    if (rc != GATE_B200_OK) {
      const std::string msg = h_ ? gate_b200_last_error(h_) : "allocation failed";
      if (h_) gate_b200_destroy(h_);
      h_ = nullptr;
      throw std::runtime_error("gate_b200_create failed: " + msg);   // no CPU fallback
    }
  }
  ~GateB200Engine() { if (h_) gate_b200_destroy(h_); }
  GateB200Engine(const GateB200Engine &) = delete;

  void Check(int rc) const {
    if (rc != GATE_B200_OK) throw std::runtime_error(gate_b200_last_error(h_));
  }

  void SetWorld(std::array<double, 3> half, int material) { Check(gate_b200_set_world(h_, half.data(), material)); }

  /* ImageVolume: the HU image and the np.digitize table of create_label_image (volumes.py:1200-1243) */
  void SetImageVolumeHU(py::array_t<int16_t, py::array::c_style | py::array::forcecast> hu, std::array<double, 3> spacing,
                        std::array<double, 12> world_to_volume, std::vector<double> bins, std::vector<uint16_t> bin_labels,
                        int label_bytes) {
    if (hu.ndim() != 3) throw std::runtime_error("HU image must be (z, y, x)");
    const int32_t size[3] = {(int32_t)hu.shape(2), (int32_t)hu.shape(1), (int32_t)hu.shape(0)};
    Check(gate_b200_set_image_volume_hu(h_, hu.data(), GATE_B200_HU_INT16, size, spacing.data(), world_to_volume.data(),
                                        (int32_t)bins.size(), bins.data(), bin_labels.data(), label_bytes, nullptr));
  }

  void BeginRun(int run_id) { Check(gate_b200_begin_run(h_, run_id)); }
  /* GateSourceManager::StartMasterThread for one source of one run: enqueue + wait */
  void RunSource(int source_id, uint64_t first_history, uint64_t n) {
    Check(gate_b200_run_source(h_, source_id, first_history, n));
    Check(gate_b200_sync(h_));
  }

  /* fetch_from_cpp_image (doseactors.py:255-270): one accumulator as a (z, y, x) float64 array */
  py::array_t<double> GetDoseGrid(int actor, int which, std::array<int, 3> size_xyz) {
    py::array_t<double> out({size_xyz[2], size_xyz[1], size_xyz[0]});
    Check(gate_b200_get_dose_grid(h_, actor, which, out.mutable_data()));
    return out;
  }

  /* the same with the end-of-run unit conversion done on the device (doseactors.py:685-696) */
  py::array_t<double> GetDoseGridScaled(int actor, int which, double scale, std::array<int, 3> size_xyz) {
    py::array_t<double> out({size_xyz[2], size_xyz[1], size_xyz[0]});
    Check(gate_b200_get_dose_grid_scaled(h_, actor, which, scale, out.mutable_data()));
    return out;
  }
  /* relative uncertainty image of edep / dose evaluated on the device (dataitems.py:1956-1993) */
  py::array_t<double> GetUncertainty(int actor, int which, uint64_t n_events, std::array<int, 3> size_xyz) {
    py::array_t<double> out({size_xyz[2], size_xyz[1], size_xyz[0]});
    Check(gate_b200_get_uncertainty(h_, actor, which, n_events, out.mutable_data()));
    return out;
  }
  std::string KernelName() { return gate_b200_kernel_name(h_); }

  py::dict GetStats() {
    gate_b200_stats s;
    Check(gate_b200_get_stats(h_, &s));
    py::dict d;
    d["events"] = s.histories; d["tracks"] = s.tracks; d["steps"] = s.steps;
    d["tentative_collisions"] = s.tentative_collisions; d["edep_total_MeV"] = s.edep_total_MeV;
    d["kernel_ms"] = s.kernel_ms;
    return d;
  }

  gate_b200_engine *handle() { return h_; }   // set_tables / add_source / add_*_actor take the C structs:
                                              // filled on the Python side exactly as opengate_b200/abi.py does
private:
  gate_b200_engine *h_ = nullptr;
};

}  // namespace

void init_GateB200Engine(py::module &m) {
  py::class_<GateB200Engine>(m, "GateB200Engine")
      .def(py::init<int, uint64_t>(), py::arg("device"), py::arg("seed"))
      .def("SetWorld", &GateB200Engine::SetWorld)
      .def("SetImageVolumeHU", &GateB200Engine::SetImageVolumeHU)
      .def("BeginRun", &GateB200Engine::BeginRun)
      .def("RunSource", &GateB200Engine::RunSource, py::call_guard<py::gil_scoped_release>())
      .def("GetDoseGrid", &GateB200Engine::GetDoseGrid)
      .def("GetDoseGridScaled", &GateB200Engine::GetDoseGridScaled)
      .def("GetUncertainty", &GateB200Engine::GetUncertainty)
      .def("KernelName", &GateB200Engine::KernelName)
      .def("GetStats", &GateB200Engine::GetStats)
      .def("GetHandleAddress", [](GateB200Engine &e) { return reinterpret_cast<uintptr_t>(e.handle()); });
  m.def("GateB200AbiVersion", &gate_b200_abi_version);
}
