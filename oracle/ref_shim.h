// TEST INFRASTRUCTURE ONLY — force-included (-include) when oracle/build_ref.py compiles the reference's OWN,
// UNMODIFIED lib/cuda/*.cu from /root/reference against torch >= 2.x.
//
// The reference dispatches with  AT_DISPATCH_FLOATING_TYPES(tensor.type(), ...)  (lib/cuda/render_utils_kernel.cu:86,
// 105,122,223,276,340,384,419,493,546; lib/cuda/adam_upd_kernel.cu:74,98,123).  Current ATen's dispatch macro calls
// ::detail::scalar_type(the_type), whose overload for the deprecated Tensor::type() result was removed; restoring that
// one overload lets the reference sources compile as they are, without a patched copy.
#pragma once
#include <ATen/ATen.h>
#include <ATen/Dispatch.h>

namespace detail {
inline at::ScalarType scalar_type(const at::DeprecatedTypeProperties& t) { return t.scalarType(); }
}  // namespace detail
