// Instantiations of the backward kernel and of the cold-start factor kernel.
#include "intopt_kernels.cuh"

namespace intopt_k {
const void* pick_backward(int threads, bool allsmem) {
#define PICK_(T, A) \
  if (threads <= T && allsmem == A) return (const void*)intopt_backward_kernel<T, A>;
  INTOPT_VARIANTS(PICK_)
#undef PICK_
  return nullptr;
}
const void* pick_cold(int threads, bool allsmem) {
#define PICK_(T, A) \
  if (threads <= T && allsmem == A) return (const void*)intopt_coldfactor_kernel<T, A>;
  INTOPT_VARIANTS(PICK_)
#undef PICK_
  return nullptr;
}
void list_backward(std::vector<const void*>& v) {
#define ADD_(T, A)                                          \
  v.push_back((const void*)intopt_backward_kernel<T, A>);   \
  v.push_back((const void*)intopt_coldfactor_kernel<T, A>);
  INTOPT_VARIANTS(ADD_)
#undef ADD_
}
}  // namespace intopt_k
