#include "comm_api.h"
namespace dtg {
void bind_comm(pybind11::module_& m) { (void)m; }
}  // namespace dtg
