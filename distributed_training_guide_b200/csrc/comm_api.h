// NVLink symmetric-memory runtime + collective kernels: Python binding entry point.
#pragma once
#include <pybind11/pybind11.h>

namespace dtg {
void bind_comm(pybind11::module_& m);
void bind_attention(pybind11::module_& m);
void bind_tp(pybind11::module_& m);
void bind_dataloader(pybind11::module_& m);
void bind_symm_vmm(pybind11::module_& m);
}  // namespace dtg
