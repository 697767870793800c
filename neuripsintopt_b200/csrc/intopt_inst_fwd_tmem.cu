// Instantiations of the predictor-corrector forward kernel and of the backward kernel that keep their element-wise
// n-vectors in tensor memory (128-thread CTAs, four per SM; 256-thread CTAs, two per SM; 512-thread CTAs, one per SM).
#include "intopt_kernels.cuh"

namespace intopt_k {
#define INTOPT_TMEM_VARIANTS(F) F(128, true) F(128, false) F(256, true) F(256, false) F(512, true) F(512, false)
const void* pick_forward_tmem(int threads, bool allsmem) {
#define PICK_(T, A) \
  if (threads <= T && allsmem == A) return (const void*)intopt_forward_tmem_kernel<T, A>;
  INTOPT_TMEM_VARIANTS(PICK_)
#undef PICK_
  return nullptr;
}
const void* pick_backward_tmem(int threads, bool allsmem) {
#define PICK_(T, A) \
  if (threads <= T && allsmem == A) return (const void*)intopt_backward_tmem_kernel<T, A>;
  INTOPT_TMEM_VARIANTS(PICK_)
#undef PICK_
  return nullptr;
}
void list_tmem(std::vector<const void*>& v) {
#define ADD_(T, A)                                                   \
  v.push_back((const void*)intopt_forward_tmem_kernel<T, A>); \
  v.push_back((const void*)intopt_backward_tmem_kernel<T, A>);
  INTOPT_TMEM_VARIANTS(ADD_)
#undef ADD_
}
}  // namespace intopt_k
