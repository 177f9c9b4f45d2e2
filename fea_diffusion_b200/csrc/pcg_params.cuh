// Device-resident parameter block shared by the PCG kernels (k_pcg.cu, k_pcg_cluster.cu).
#pragma once
#include "fea_internal.cuh"

namespace fea {

// Everything the solver kernels need, resident in device memory (Ctx::d_pcg_params) instead of
// being passed by value: the kernel nodes of an instantiated CUDA graph then carry no batch
// pointers, so one graph per grid size serves every batch the context ever solves.
struct PcgPtrs {
  const int32_t* sys_of_cta;
  const int32_t* cta_first;
  const int32_t* cta_count;
  const int32_t* slice_len;
  const int64_t* slice_ptr;
  const d4* val;
  const int32_t* col;
  double2* x;
  d4* rp;            // (r.x, r.y, p.x, p.y) per block row; p = direction of the PREVIOUS iteration
  double2* q;
  double* partA;
  double* partB;
  SysScalars sc;
  double* rz_last;
  int4* active;      // compacted work list: (cta, system, first cta of system, cta count)
  int32_t n_active;  // valid entries of `active` (written by k_compact_active)
  int32_t ncta;      // CTAs of the whole batch
  int32_t ns;
  int32_t max_iter;
  int32_t two_level;
  // cluster path (k_pcg_cluster.cu): systems small enough to stay on chip.  The cluster size is
  // a function of the system alone (class c = clusters of c CTAs, 1..8: the smallest cluster whose
  // CTAs can hold the system's rows), so that its summation order -- and therefore every bit of
  // its result -- does not depend on the rest of the batch.
  const int32_t* cl_order;  // eligible systems grouped by class; largest first in each
  int32_t cl_off[9];        // first entry of each class in cl_order (index = CTAs per cluster)
  int32_t cl_cnt[9];        // entries of each class (0 = class not launched)
  int32_t* cl_counter;      // [1..8] work-queue heads, [0] restarts, [10] systems handed back to the
                            // streaming kernels (device counters, zeroed per solve); [16..21] three 64-bit
                            // counters: block reads from tensor memory / shared memory / L2
  int32_t cl_halo_cap;      // largest halo (rows gathered from other CTAs) a CTA accepts (test knob)
  // extended-precision refinement of the on-chip path (k_pcg_cluster.cu, dd_residual_rows)
  int32_t refine_dd;        // 1 = systems that hit the fp64 floor get double-double refinement rounds
  double2* xlo;             // low part of the solution of a refined system (high part: x)
  const double2* sb;        // S b of every block row, as k_pcg_init_vectors rounded it
};
static_assert(sizeof(PcgPtrs) <= kPcgParamBytes, "grow kPcgParamBytes");

}  // namespace fea
