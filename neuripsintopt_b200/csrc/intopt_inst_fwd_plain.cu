// Instantiations of the plain path-following forward kernel (pc=False upstream, ip_model_whole.py:1392-1394).
#include "intopt_kernels.cuh"

namespace intopt_k {
const void* pick_forward_plain(int threads, bool allsmem) {
#define PICK_(T, A) \
  if (threads <= T && allsmem == A) return (const void*)intopt_forward_kernel<T, A, true>;
  INTOPT_VARIANTS(PICK_)
#undef PICK_
  return nullptr;
}
void list_forward_plain(std::vector<const void*>& v) {
#define ADD_(T, A) v.push_back((const void*)intopt_forward_kernel<T, A, true>);
  INTOPT_VARIANTS(ADD_)
#undef ADD_
}
}  // namespace intopt_k
