// mjb_internal.h — definitions shared by the C-ABI translation units.
#pragma once
#include <string>

#include "host_model.h"

struct mjb_model {
  mjb::HostModel host;
};

namespace mjb {
void set_error(const std::string& s);
}
