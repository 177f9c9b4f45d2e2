// K5: lock-step Jacobi-preconditioned fp64 CG over all samples of a batch -- replaces
// Newton + ScipyDirect(SuperLU) + SimpleTimeSteppingSolver (reference
// datagen/fea_analysis.py:371-375, 425-439; SURVEY A-12..A-14, F5).
//
// The matrix is stored Jacobi-scaled (Khat = S K S), so preconditioned CG on K is plain CG
// on Khat:  two kernels per iteration, no host involvement:
//
//   spmv  : p  = r + beta * p_old        (beta = rz_cur / rz_prev, from device scalars)
//           q  = Khat p                  (block-SELL-32, one thread per 2x2 block row;
//                                         p is formed on the fly at the gathered columns, so
//                                         no separate p-update pass over HBM is needed)
//           pq = p . q                   (warp shuffle -> CTA -> last-block-per-system)
//   update: alpha = rz_cur / pq;  x += alpha p;  r -= alpha q;  rz_next = r . r
//           convergence / breakdown / max-iter flags per system.
//
// Every CTA (128 block rows) belongs to exactly one system.  Dot products are reduced in a
// fixed order (thread -> warp tree -> CTA partial -> ordered sum by the last-arriving CTA of the
// system), so results are bitwise reproducible and independent of how many systems share the
// batch or the GPU.  Finished systems cost one flag read per CTA.
#include "fea_internal.cuh"

namespace fea {

constexpr int kT = kCtaRows;  // threads per CTA

// Deterministic CTA sum of one double per thread; result valid in thread 0.
__device__ __forceinline__ double cta_sum(double v, double* sm /*[kT/32]*/) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  double t = 0.0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int w = 0; w < kT / 32; ++w) t += sm[w];
  }
  return t;
}

// Publishes this CTA's partial; returns (to all threads) whether this CTA is the last of its
// system to arrive, in which case *total (thread 0) holds the ordered sum of all partials.
__device__ __forceinline__ bool system_reduce(double cta_partial, double* __restrict__ part,
                                              unsigned int* __restrict__ cnt, int s, int first, int n,
                                              double* sm, int* sm_flag, double* total) {
  if (threadIdx.x == 0) {
    part[blockIdx.x] = cta_partial;
    __threadfence();
    const unsigned prev = atomicAdd(&cnt[s], 1u);
    *sm_flag = (prev == (unsigned)(n - 1));
  }
  __syncthreads();
  const bool last = *sm_flag != 0;
  if (!last) return false;
  __threadfence();
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += kT) acc += __ldcg(part + first + i);
  __syncthreads();  // sm reuse
  const double t = cta_sum(acc, sm);
  if (threadIdx.x == 0) {
    *total = t;
    cnt[s] = 0u;
  }
  return true;
}

struct PcgPtrs {
  const int32_t* sys_of_cta;
  const int32_t* cta_first;
  const int32_t* cta_count;
  const int32_t* slice_len;
  const int64_t* slice_ptr;
  const double2* val;
  const int32_t* col;
  double2* x;
  double2* r;
  double2* q;
  double* partA;
  double* partB;
  SysScalars sc;
  double* rz_last;
};

__global__ void __launch_bounds__(kT) k_pcg_spmv(PcgPtrs P, const double2* __restrict__ p_old,
                                                 double2* __restrict__ p_new, int parity, int check) {
  __shared__ double sm[kT / 32];
  __shared__ int sm_flag;
  const int s = P.sys_of_cta[blockIdx.x];
  if (s < 0) return;
  if (P.sc.done[s]) return;
  const double beta = P.sc.rz[parity][s] / P.sc.rz[parity ^ 1][s];
  const int64_t row = (int64_t)blockIdx.x * kT + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const int64_t slice = row >> 5;
  const int L = P.slice_len[slice];
  const int64_t base = P.slice_ptr[slice];
  const double2* __restrict__ vt = P.val + 2 * base + lane;
  const int32_t* __restrict__ cp = P.col + base + lane;
  const double2* __restrict__ r = P.r;
  double a0 = 0.0, a1 = 0.0;
#pragma unroll 4
  for (int j = 0; j < L; ++j) {
    const int c = ld_stream_i32(cp + j * 32);
    const double2 t = ld_stream_f64x2(vt + j * 64);
    const double2 b = ld_stream_f64x2(vt + j * 64 + 32);
    const double2 rj = __ldg(r + c);
    const double2 pj = __ldg(p_old + c);
    const double px = fma(beta, pj.x, rj.x);
    const double py = fma(beta, pj.y, rj.y);
    a0 = fma(t.x, px, a0);
    a0 = fma(t.y, py, a0);
    a1 = fma(b.x, px, a1);
    a1 = fma(b.y, py, a1);
  }
  const double2 ri = __ldg(r + row);
  const double2 pi = __ldg(p_old + row);
  const double2 pn = make_double2(fma(beta, pi.x, ri.x), fma(beta, pi.y, ri.y));
  p_new[row] = pn;
  P.q[row] = make_double2(a0, a1);
  const double part = cta_sum(fma(pn.x, a0, pn.y * a1), sm);
  double total;
  if (system_reduce(part, P.partA, P.sc.cntA, s, P.cta_first[s], P.cta_count[s], sm, &sm_flag, &total)) {
    if (threadIdx.x == 0) {
      P.sc.pq[s] = total;
      if (check && !(total > 0.0 && isfinite(total))) {  // not SPD on this Krylov space
        P.sc.done[s] = 1;
        P.sc.status[s] = FEA_SAMPLE_BREAKDOWN;
        atomicAdd(P.sc.n_done, 1);
      }
    }
  }
}

__global__ void __launch_bounds__(kT) k_pcg_update(PcgPtrs P, const double2* __restrict__ p, int parity,
                                                   int max_iter) {
  __shared__ double sm[kT / 32];
  __shared__ int sm_flag;
  const int s = P.sys_of_cta[blockIdx.x];
  if (s < 0) return;
  if (P.sc.done[s]) return;
  const double alpha = P.sc.rz[parity][s] / P.sc.pq[s];
  const int64_t row = (int64_t)blockIdx.x * kT + threadIdx.x;
  const double2 pv = p[row];
  const double2 qv = P.q[row];
  double2 xv = P.x[row];
  double2 rv = P.r[row];
  xv.x = fma(alpha, pv.x, xv.x);
  xv.y = fma(alpha, pv.y, xv.y);
  rv.x = fma(-alpha, qv.x, rv.x);
  rv.y = fma(-alpha, qv.y, rv.y);
  P.x[row] = xv;
  P.r[row] = rv;
  const double part = cta_sum(fma(rv.x, rv.x, rv.y * rv.y), sm);
  double total;
  if (system_reduce(part, P.partB, P.sc.cntB, s, P.cta_first[s], P.cta_count[s], sm, &sm_flag, &total)) {
    if (threadIdx.x == 0) {
      P.sc.rz[parity ^ 1][s] = total;
      P.rz_last[s] = total;
      const int it = P.sc.iters[s] + 1;
      P.sc.iters[s] = it;
      int st = -1;
      if (!isfinite(total)) st = FEA_SAMPLE_BREAKDOWN;
      else if (total <= P.sc.tol2[s]) st = FEA_SAMPLE_CONVERGED;
      else if (it >= max_iter) st = FEA_SAMPLE_MAX_ITER;
      if (st >= 0) {
        P.sc.done[s] = 1;
        P.sc.status[s] = st;
        atomicAdd(P.sc.n_done, 1);
      }
    }
  }
}

// r0 = S b, x0 = 0, p = 0; r0.r0 per system; flags.
__global__ void __launch_bounds__(kT) k_pcg_init(PcgPtrs P, const int32_t* __restrict__ vertex_of_row,
                                                 const double* __restrict__ rhs, const double* __restrict__ dscale,
                                                 double2* __restrict__ p0, double2* __restrict__ p1,
                                                 const int32_t* __restrict__ empty, double rtol) {
  __shared__ double sm[kT / 32];
  __shared__ int sm_flag;
  const int s = P.sys_of_cta[blockIdx.x];
  if (s < 0) return;
  const int64_t row = (int64_t)blockIdx.x * kT + threadIdx.x;
  const int v = vertex_of_row[row];
  double2 b = make_double2(0.0, 0.0);
  if (v >= 0) b = make_double2(dscale[2 * row] * rhs[2 * (int64_t)v], dscale[2 * row + 1] * rhs[2 * (int64_t)v + 1]);
  const double2 z = make_double2(0.0, 0.0);
  P.x[row] = z;
  P.r[row] = b;
  P.q[row] = z;
  p0[row] = z;
  p1[row] = z;
  const double part = cta_sum(fma(b.x, b.x, b.y * b.y), sm);
  double total;
  if (system_reduce(part, P.partB, P.sc.cntB, s, P.cta_first[s], P.cta_count[s], sm, &sm_flag, &total)) {
    if (threadIdx.x == 0) {
      P.sc.rz[0][s] = total;
      P.sc.rz[1][s] = __longlong_as_double(0x7ff0000000000000LL);  // +inf -> beta_0 = 0
      P.sc.rz0[s] = total;
      P.rz_last[s] = total;
      P.sc.tol2[s] = rtol * rtol * total;
      P.sc.pq[s] = 1.0;
      P.sc.iters[s] = 0;
      P.sc.cntA[s] = 0u;
      int st = FEA_SAMPLE_NOT_RUN, dn = 0;
      if (empty[s]) { st = FEA_SAMPLE_EMPTY_ROW; dn = 1; }
      else if (!(total > 0.0)) { st = isfinite(total) ? FEA_SAMPLE_CONVERGED : FEA_SAMPLE_BREAKDOWN; dn = 1; }
      P.sc.status[s] = st;
      P.sc.done[s] = dn;
      if (dn) atomicAdd(P.sc.n_done, 1);
    }
  }
}

static PcgPtrs make_ptrs(Batch& b) {
  PcgPtrs P;
  P.sys_of_cta = b.sys_of_cta;
  P.cta_first = b.cta_first;
  P.cta_count = b.cta_count;
  P.slice_len = b.slice_len;
  P.slice_ptr = b.slice_ptr;
  P.val = b.val;
  P.col = b.col;
  P.x = (double2*)b.x;
  P.r = (double2*)b.r;
  P.q = (double2*)b.q;
  P.partA = b.partA;
  P.partB = b.partB;
  P.sc = b.sc;
  P.rz_last = b.rz_last;
  return P;
}

cudaError_t launch_pcg_init(Batch& b, double rtol) {
  const int ncta = (int)(b.NBR / kCtaRows);
  cudaStream_t st = b.ctx->stream;
  cudaMemsetAsync(b.sc.n_done, 0, sizeof(int32_t), st);
  cudaMemsetAsync(b.sc.cntB, 0, sizeof(unsigned) * b.ns, st);
  if (ncta)
    k_pcg_init<<<ncta, kT, 0, st>>>(make_ptrs(b), b.vertex_of_row, b.rhs, b.dscale, (double2*)b.p0,
                                    (double2*)b.p1, b.empty, rtol);
  return cudaGetLastError();
}

// iteration parity 0 reads p0 / writes p1, parity 1 the other way round
cudaError_t launch_pcg_spmv(Batch& b, int parity, cudaStream_t st) {
  const int ncta = (int)(b.NBR / kCtaRows);
  if (!ncta) return cudaSuccess;
  const double2* po = (const double2*)(parity ? b.p1 : b.p0);
  double2* pn = (double2*)(parity ? b.p0 : b.p1);
  k_pcg_spmv<<<ncta, kT, 0, st>>>(make_ptrs(b), po, pn, parity, 1);
  return cudaGetLastError();
}

cudaError_t launch_pcg_update(Batch& b, int parity, int max_iter, cudaStream_t st) {
  const int ncta = (int)(b.NBR / kCtaRows);
  if (!ncta) return cudaSuccess;
  const double2* pn = (const double2*)(parity ? b.p0 : b.p1);
  k_pcg_update<<<ncta, kT, 0, st>>>(make_ptrs(b), pn, parity, max_iter);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// finalisation: u = S xhat scattered to the full DOF numbering (zeros at fixed DOFs, A-15),
// per-sample (min, max) of both components (ranges.txt, A-17), relative residuals.
// ---------------------------------------------------------------------------
__global__ void k_unscale_scatter(int64_t NV, const int32_t* __restrict__ vsample,
                                  const int32_t* __restrict__ row_of_vertex, const double* __restrict__ dscale,
                                  const double2* __restrict__ x, const int32_t* __restrict__ status,
                                  double2* __restrict__ u) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= NV) return;
  double2 o = make_double2(0.0, 0.0);
  if (status[vsample[v]] == FEA_SAMPLE_EMPTY_ROW) {
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    o = make_double2(nan, nan);
  } else {
    const int row = row_of_vertex[v];
    if (row >= 0) {
      const double2 xv = x[row];
      o = make_double2(dscale[2 * (int64_t)row] * xv.x, dscale[2 * (int64_t)row + 1] * xv.y);
    }
  }
  u[v] = o;
}

__global__ void k_minmax(const int64_t* __restrict__ vtx_off, const double2* __restrict__ u,
                         double* __restrict__ ranges) {
  __shared__ double sm[4][8];
  const int s = blockIdx.x;
  const int64_t v0 = vtx_off[s], v1 = vtx_off[s + 1];
  const double inf = __longlong_as_double(0x7ff0000000000000LL);
  double mnx = inf, mxx = -inf, mny = inf, mxy = -inf;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += blockDim.x) {
    const double2 w = u[v];
    mnx = fmin(mnx, w.x); mxx = fmax(mxx, w.x);
    mny = fmin(mny, w.y); mxy = fmax(mxy, w.y);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fmin(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
    mxx = fmax(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
    mny = fmin(mny, __shfl_xor_sync(0xffffffffu, mny, o));
    mxy = fmax(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
  }
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) { sm[0][w] = mnx; sm[1][w] = mxx; sm[2][w] = mny; sm[3][w] = mxy; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) {
      mnx = fmin(mnx, sm[0][i]); mxx = fmax(mxx, sm[1][i]);
      mny = fmin(mny, sm[2][i]); mxy = fmax(mxy, sm[3][i]);
    }
    ranges[4 * s + 0] = mnx; ranges[4 * s + 1] = mxx;
    ranges[4 * s + 2] = mny; ranges[4 * s + 3] = mxy;
  }
}

__global__ void k_relres(int ns, const double* __restrict__ rz_last, const double* __restrict__ rz0,
                         double* __restrict__ relres) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  relres[s] = rz0[s] > 0.0 ? sqrt(rz_last[s] / rz0[s]) : 0.0;
}

cudaError_t launch_finalize(Batch& b) {
  cudaStream_t st = b.ctx->stream;
  const int T = 256;
  if (b.NV)
    k_unscale_scatter<<<(unsigned)((b.NV + T - 1) / T), T, 0, st>>>(b.NV, b.vsample, b.row_of_vertex, b.dscale,
                                                                   (const double2*)b.x, b.sc.status, (double2*)b.u);
  k_minmax<<<b.ns, 256, 0, st>>>(b.d_vtx_off, (const double2*)b.u, b.ranges);
  k_relres<<<(b.ns + T - 1) / T, T, 0, st>>>(b.ns, b.rz_last, b.sc.rz0, b.relres);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// Test hook: y = K x for one sample through k_pcg_spmv (beta = 0 path).
// ---------------------------------------------------------------------------
__global__ void k_spmv_load(int64_t NBR, const int32_t* __restrict__ vertex_of_row,
                            const int32_t* __restrict__ vrank, int64_t v0, int64_t v1,
                            const double* __restrict__ dscale, const double* __restrict__ xin,
                            double2* __restrict__ r, double2* __restrict__ p0) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= NBR) return;
  const int v = vertex_of_row[row];
  double2 o = make_double2(0.0, 0.0);
  if (v >= v0 && v < v1) {
    const int rk = vrank[v];
    const double s0 = dscale[2 * row], s1 = dscale[2 * row + 1];
    o = make_double2(s0 > 0 ? xin[2 * rk] / s0 : 0.0, s1 > 0 ? xin[2 * rk + 1] / s1 : 0.0);
  }
  r[row] = o;
  p0[row] = make_double2(0.0, 0.0);
}
__global__ void k_spmv_store(int64_t NBR, const int32_t* __restrict__ vertex_of_row,
                             const int32_t* __restrict__ vrank, int64_t v0, int64_t v1,
                             const double* __restrict__ dscale, const double2* __restrict__ q,
                             double* __restrict__ yout) {
  const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= NBR) return;
  const int v = vertex_of_row[row];
  if (v >= v0 && v < v1) {
    const int rk = vrank[v];
    const double s0 = dscale[2 * row], s1 = dscale[2 * row + 1];
    const double2 qv = q[row];
    yout[2 * rk] = s0 > 0 ? qv.x / s0 : 0.0;
    yout[2 * rk + 1] = s1 > 0 ? qv.y / s1 : 0.0;
  }
}
__global__ void k_spmv_scalars(int ns, SysScalars sc) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns) return;
  sc.rz[0][s] = 1.0;
  sc.rz[1][s] = __longlong_as_double(0x7ff0000000000000LL);
  sc.done[s] = 0;
  sc.cntA[s] = 0u;
}

cudaError_t launch_plain_spmv(Batch& b, int32_t s, const double* d_x, double* d_y) {
  cudaStream_t st = b.ctx->stream;
  const int T = 256;
  const int ncta = (int)(b.NBR / kCtaRows);
  if (!ncta) return cudaSuccess;
  const unsigned g = (unsigned)((b.NBR + T - 1) / T);
  const int64_t v0 = b.vtx_off[s], v1 = b.vtx_off[s + 1];
  k_spmv_scalars<<<(b.ns + T - 1) / T, T, 0, st>>>(b.ns, b.sc);
  k_spmv_load<<<g, T, 0, st>>>(b.NBR, b.vertex_of_row, b.vrank, v0, v1, b.dscale, d_x, (double2*)b.r, (double2*)b.p0);
  k_pcg_spmv<<<ncta, kT, 0, st>>>(make_ptrs(b), (const double2*)b.p0, (double2*)b.p1, 0, 0);
  k_spmv_store<<<g, T, 0, st>>>(b.NBR, b.vertex_of_row, b.vrank, v0, v1, b.dscale, (const double2*)b.q, d_y);
  return cudaGetLastError();
}

}  // namespace fea
