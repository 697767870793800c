// Instantiations of the predictor-corrector forward kernel and of the backward kernel that keep their element-wise
// n-vectors in tensor memory.
#include "intopt_kernels.cuh"

namespace intopt_k {
const void* pick_forward_tmem(bool allsmem) {
  return allsmem ? (const void*)intopt_forward_tmem_kernel<true> : (const void*)intopt_forward_tmem_kernel<false>;
}
const void* pick_backward_tmem(bool allsmem) {
  return allsmem ? (const void*)intopt_backward_tmem_kernel<true> : (const void*)intopt_backward_tmem_kernel<false>;
}
void list_tmem(std::vector<const void*>& v) {
  v.push_back((const void*)intopt_forward_tmem_kernel<true>);
  v.push_back((const void*)intopt_forward_tmem_kernel<false>);
  v.push_back((const void*)intopt_backward_tmem_kernel<true>);
  v.push_back((const void*)intopt_backward_tmem_kernel<false>);
}
}  // namespace intopt_k
