// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// Force-included (-include) when oracle/build_ref.py compiles the UNMODIFIED reference
// sources in /root/reference/lib/cuda against torch >= 2.x.  The reference writes
// AT_DISPATCH_FLOATING_TYPES(x.type(), ...) (e.g. render_utils_kernel.cu:92); modern torch's
// dispatch macro resolves the first argument through ::detail::scalar_type(), which no longer
// has an overload for at::DeprecatedTypeProperties.  Supplying that one overload restores the
// old behaviour without touching (or copying) a single reference line.
#pragma once
#include <ATen/core/DeprecatedTypeProperties.h>
#include <c10/core/ScalarType.h>
namespace detail {
inline c10::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
