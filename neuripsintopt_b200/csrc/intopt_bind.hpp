// Binds a HostStructure to the pointer-only ProblemDev the solver code consumes.  `Up` maps a host
// std::vector to a pointer usable by the solver: identity for the single-threaded test build,
// cudaMalloc + cudaMemcpy for the CUDA build.
#pragma once
#include <cstdlib>

#include "intopt_solver.cuh"
#include "intopt_structure.hpp"

namespace intopt {

// the level schedule is only bound (and staged) when it shortens the chain: n_groups < mb
inline bool use_levels(const HostStructure& S) {
  const char* v = std::getenv("INTOPT_LEVELS");  // INTOPT_LEVELS=0: column-by-column elimination (A/B measurements)
  if (v && v[0] == '0') return false;
  return S.lv_ok && S.n_groups > 0 && S.n_groups < S.mb;
}

// ints of schedule arrays the kernels stage in shared memory (== ProblemDev::meta_ints after bind_problem)
inline int meta_ints(const HostStructure& S) {
  ProblemDev D{};
  static const uint32_t marker32 = 0;
  D.mb = S.mb;
  D.m = S.m;
  D.n = S.n;
  D.n_act = (int)S.act_blk.size();
  D.n_asm = S.n_asm;
  D.rs_hdr = S.rs_slots ? &marker32 : nullptr;
  D.asm_run = S.asm_run.empty() ? nullptr : &marker32;
  if (use_levels(S)) {
    D.n_groups = S.n_groups;
    D.n_rg = (int)S.rg_blk.size();
    const std::vector<int>* lists[LV_NLISTS] = {&S.lv_col_rec, &S.lv_fl_ptr, &S.lv_fl,     &S.lv_fo_ptr, &S.lv_fo, &S.lv_fsrc_ptr,
                                                &S.lv_fsrc,    &S.lv_bl_ptr, &S.lv_bl,     &S.lv_bo_ptr, &S.lv_bo, &S.lv_bsrc};
    for (int a = 0; a < LV_NLISTS; ++a) D.lv_cnt[a] = (int)lists[a]->size();
  }
  return meta_stage_ints(D);
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
  D.g_act_blk = use_levels(S) ? up(S.lv_act) : up(S.act_blk);  // (bit 16 = shared row, only read by the level code)
  D.n_act = (int)S.act_blk.size();
  D.n_groups = use_levels(S) ? S.n_groups : 0;
  D.g_grp_ptr = use_levels(S) ? up(S.grp_ptr) : nullptr;
  D.g_grp_col = use_levels(S) ? up(S.grp_col) : nullptr;
  D.g_col_group = use_levels(S) ? up(S.col_group) : nullptr;
  D.g_rg_ptr = use_levels(S) ? up(S.rg_ptr) : nullptr;
  D.g_rg_blk = use_levels(S) ? up(S.rg_blk) : nullptr;
  D.n_rg = use_levels(S) ? (int)S.rg_blk.size() : 0;
  D.g_grp_tab = use_levels(S) ? up(S.grp_tab) : nullptr;
  {
    const std::vector<int>* lists[LV_NLISTS] = {&S.lv_col_rec, &S.lv_fl_ptr, &S.lv_fl,     &S.lv_fo_ptr, &S.lv_fo, &S.lv_fsrc_ptr,
                                                &S.lv_fsrc,    &S.lv_bl_ptr, &S.lv_bl,     &S.lv_bo_ptr, &S.lv_bo, &S.lv_bsrc};
    for (int a = 0; a < LV_NLISTS; ++a) {
      D.g_lv[a] = use_levels(S) ? up(*lists[a]) : nullptr;
      D.lv_cnt[a] = use_levels(S) ? (int)lists[a]->size() : 0;
    }
  }
  D.meta_ints = 0;  // (set below, once every count it depends on is known)
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
  D.meta_ints = meta_stage_ints(D);
}

// doubles of shared memory the solver needs beyond the (optional) factor and gather buffer
inline size_t smem_fixed_doubles(const HostStructure& S) {
  return (size_t)S_NVEC * (S.m8 + 8) + S_SCRATCH + (use_levels(S) ? S_SCRATCH_LV : 0) + (scan_in_factor(S.bsize, S.n) ? 0 : 2 * (size_t)scan_len(S.n));
}

}  // namespace intopt
