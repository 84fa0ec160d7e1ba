// bindings.cpp — pybind11 module `_dpe`, drop-in for the reference's csrc/bindings.cpp:13-43:
// same function name, keyword arguments and defaults, stdout/stderr redirected to Python,
// RuntimeError("DPE-MVS failed with code N") on a non-zero return.  Unlike the reference it
// does not include CPython-internal headers and it releases the GIL while the
// pipeline runs.
#include <pybind11/iostream.h>
#include <pybind11/pybind11.h>

#include <stdexcept>
#include <string>

#include "dpe_b200.h"

namespace py = pybind11;

static int dpe_mvs_py(const std::string& dense_folder, int gpu_index, bool verbose, bool fusion, bool viz, bool depth,
                      bool normal, bool weak, bool edge) {
  py::scoped_ostream_redirect out(std::cout);
  py::scoped_ostream_redirect err(std::cerr, py::module_::import("sys").attr("stderr"));
  int ret;
  {
    py::gil_scoped_release release;
    ret = dpe_run_pipeline(dense_folder.c_str(), gpu_index, verbose, fusion, viz, depth, normal, weak, edge);
  }
  if (ret != 0) throw std::runtime_error("DPE-MVS failed with code " + std::to_string(ret));
  return ret;
}

PYBIND11_MODULE(_dpe, m) {
  m.doc() = "B200-native DPE-MVS (drop-in for the reference's _dpe module)";
  m.def("dpe_mvs", &dpe_mvs_py, py::arg("dense_folder"), py::arg("gpu_index") = 0, py::arg("verbose") = true,
        py::arg("fusion") = false, py::arg("viz") = false, py::arg("depth") = true, py::arg("normal") = false,
        py::arg("weak") = false, py::arg("edge") = false);
}
