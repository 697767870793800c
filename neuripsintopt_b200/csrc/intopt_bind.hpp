// Binds a HostStructure to the pointer-only ProblemDev the solver code consumes.  `Up` maps a host
// std::vector to a pointer usable by the solver: identity for the single-threaded test build,
// cudaMalloc + cudaMemcpy for the CUDA build.
#pragma once
#include "intopt_solver.cuh"
#include "intopt_structure.hpp"

namespace intopt {

inline int meta_ints(const HostStructure& S) {
  return S.mb + 2 * (S.mb + 1) + (int)S.act_blk.size() + (S.m + 1) + (int)S.csc_gptr.size() +
         2 * (int)S.asm_gptr.size();
}

template <class Up>
inline void bind_problem(const HostStructure& S, ProblemDev& D, Up&& up) {
  D.m = S.m;
  D.n = S.n;
  D.nnz = S.nnz;
  D.mb = S.mb;
  D.m8 = S.m8;
  D.mv = S.m8 + 8;
  D.g_csr_ptr = up(S.csr_ptr);
  D.csr_col = up(S.csr_col);
  D.a_unit = S.a_unit ? 1 : 0;
  D.csr_val = S.a_unit ? nullptr : up(S.csr_val);
  D.g_csc_gptr = up(S.csc_gptr);
  D.csc_row = up(S.csc_row);
  D.csc_val = S.a_unit ? nullptr : up(S.csc_val);
  D.rs_hdr = S.rs_slots ? up(S.rs_hdr) : nullptr;
  D.rs_row = S.rs_slots ? up(S.rs_row) : nullptr;
  D.rs_slots = S.rs_slots;
  D.perm = up(S.perm);
  D.g_fb = up(S.fb);
  D.g_boff = up(S.boff);
  D.bsize = S.bsize;
  D.dummy_pos = up(S.dummy_pos);
  D.n_dummy = (int)S.dummy_pos.size();
  D.g_act_ptr = up(S.act_ptr);
  D.g_act_blk = up(S.act_blk);
  D.n_act = (int)S.act_blk.size();
  D.meta_ints = meta_ints(S);
  D.n_asm = S.n_asm;
  D.g_asm_gptr = up(S.asm_gptr);
  D.asm_pos = up(S.asm_pos);
  D.asm_diag = up(S.asm_diag);
  D.asm_k16 = S.asm_k16.empty() ? nullptr : up(S.asm_k16);
  D.asm_k = (S.asm_k16.empty() && !S.asm_k.empty()) ? up(S.asm_k) : nullptr;
  D.g_asm_rptr = up(S.asm_rptr);
  D.asm_run = S.asm_run.empty() ? nullptr : up(S.asm_run);
  D.asm_unit = S.asm_unit ? 1 : 0;
  D.asm_w = S.asm_unit ? nullptr : up(S.asm_w);
  D.inst_values = S.inst_values ? 1 : 0;
  D.n_csc_pad = (int)S.csc_val.size();
  D.n_asm_pad = (int)S.asm_w.size();
  D.csr_src = S.inst_values ? up(S.csr_src) : nullptr;
  D.csc_src = S.inst_values ? up(S.csc_src) : nullptr;
  D.asm_sa = S.inst_values ? up(S.asm_sa) : nullptr;
  D.asm_sb = S.inst_values ? up(S.asm_sb) : nullptr;
}

// doubles of shared memory the solver needs beyond the (optional) factor and gather buffer
inline size_t smem_fixed_doubles(const HostStructure& S) {
  return (size_t)S_NVEC * (S.m8 + 8) + S_SCRATCH + (scan_in_factor(S.bsize, S.n) ? 0 : 2 * (size_t)scan_len(S.n));
}

}  // namespace intopt
