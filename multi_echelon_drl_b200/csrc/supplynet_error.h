// supplynet_error.h — the one error channel of the library, shared by its translation units:
// fail(code, fmt, ...) stores the message sn_last_error() returns (thread-local) and hands back `code`.
#pragma once

namespace sn_detail {
int fail(int code, const char* fmt, ...);
const char* last_error();
}  // namespace sn_detail
