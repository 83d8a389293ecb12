// Thread-local error string behind dnab_last_error().
#pragma once
#include <string>

namespace dnab {
void setLastError(const std::string& msg);
}
