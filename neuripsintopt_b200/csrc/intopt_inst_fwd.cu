// Instantiations of the predictor-corrector forward kernel (pc=True, every reference driver).
#include "intopt_kernels.cuh"

namespace intopt_k {
const void* pick_forward_plain(int threads, bool allsmem);
const void* pick_forward(int threads, bool allsmem, bool plain) {
  if (plain) return pick_forward_plain(threads, allsmem);
#define PICK_(T, A) \
  if (threads <= T && allsmem == A) return (const void*)intopt_forward_kernel<T, A, false>;
  INTOPT_VARIANTS(PICK_)
#undef PICK_
  return nullptr;
}
void list_forward(std::vector<const void*>& v) {
#define ADD_(T, A) v.push_back((const void*)intopt_forward_kernel<T, A, false>);
  INTOPT_VARIANTS(ADD_)
#undef ADD_
}
}  // namespace intopt_k
