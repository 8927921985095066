// uedge_b200/csrc/ue_gpu.cu — C ABI (include/ue_gpu.h) + kernels of the B200 pandf1 / jac_calc path.
//
// Replaces the bodies of Pandf1rhs_interface (bbb/oderhs.m:12331-12345) and
// jac_calc_interface (bbb/oderhs.m:12297-12329).  No CPU fallback: every entry
// point fails if the CUDA device is unavailable.
//
// Kernels (all FP64, -fmad=false; see ue_device.cuh for the physics):
//   k_phase0/1/2/3      full residual, one thread per cell, SoA planes in HBM
//   k_jb_stage0/p1a/p1b/p2/p3c   batched Jacobian: all perturbed unknowns ("every colour") in flight, one launch
//                       per phase, blockIdx.y = role function; private cells + candidate rows in L2-resident
//                       global memory; ordered warp compaction of each column
//   k_scan / k_fill / k_sortrows   CSC fragments -> reference CSR (csrcsc, svr/svrut4.m:1536-1608)
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

#include "ue_device.cuh"
#include "ue_gpu.h"
#include "ue_lists.hpp"


namespace {

UeStore S;
std::string g_err;
bool g_ready = false;
bool g_alloc = false;  // device state of a previous ue_gpu_init exists (freed by the next one)
int nx, ny, NXS, NC;
int64_t neq = 0;
int64_t g_launches = 0;
bool g_host_graphs = true;
bool g_fuse23 = true;  // phases 2 and 3 of the full residual in one launch (UE_GPU_NO_FUSE23=1: two launches, for comparison)
std::vector<double> g_step_host[4];  // last dtuse, ylodt, suscal, sfscal uploaded
std::set<std::pair<const void*, const void*>> g_seen_host;
std::vector<double> g_last_yldot;  // host copy of the last residual returned (d_yldot still holds it)
std::vector<double> g_base_yl;   // yl (first neq entries) for which the base planes in d_base are current
bool g_base_valid = false, g_jac_trust_base = false, g_base_dev_valid = false;
float g_jac_ms = 0.f, g_res_ms = 0.f;
cudaStream_t g_stream = nullptr;
cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;

std::vector<void*> g_static_allocs;
double *d_base = nullptr, *d_yl = nullptr, *d_yldot00 = nullptr, *d_tmp = nullptr, *d_yldot = nullptr;
double *d_dtuse = nullptr, *d_ylodt = nullptr, *d_suscal = nullptr, *d_sfscal = nullptr, *d_dtoptv = nullptr;
int* d_err = nullptr;
volatile long long* h_flags = nullptr;  // pinned + mapped: [0] error bits of the residual sequence, [1] ia(neq+1) of the last Jacobian, [2] error bits of the Jacobian sequence
long long* d_hflags = nullptr;          // device alias of h_flags
int64_t g_nnz_guess = 0;
// Jacobian work space
int64_t g_ivmin = 1, g_ivmax = 0;
std::vector<int> h_list;  // unknowns (iv) of the column range
std::vector<int64_t> h_coloff;  // per column offset into the fragment buffers (1-based iv -> h_coloff[iv-1])
std::vector<int> h_cellcand_off, h_cand_cell, h_cand_east;  // per-cell candidate lists (CSR-like)
int *d_cand_cell = nullptr, *d_cand_east = nullptr, *d_item_u = nullptr, *d_guard_items = nullptr;
int g_nguard = 0;
int* d_guard_cells = nullptr;  // all guard cells, sorted by kind and padded per kind to whole warps (-1)
int g_nguard_cells = 0;
void* d_uinfo = nullptr;
double *d_priv = nullptr, *d_jrows = nullptr, *d_rres = nullptr;
int* d_rmask = nullptr;
int g_nitems = 0;
int64_t* d_coloff = nullptr;
int *d_colcnt = nullptr, *d_colrow = nullptr;
double* d_colval = nullptr;
int *d_rowcnt = nullptr, *d_rowfill = nullptr;
int64_t *d_ia = nullptr, *d_ja = nullptr;
double* d_jac = nullptr;
int64_t g_cap_total = 0, g_nnzcap = 0;
// ---- multi-GPU: one Jacobian, columns split over the ranks, CSC fragments all-gathered over NCCL (ue_gpu_comm_init) ----
// NCCL is bound at run time (dlopen of the copy already in the process, else libnccl.so.2): a single-GPU host needs no NCCL.
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
} NC_;
ncclComm_t g_comm = nullptr;
int g_nranks = 1, g_rank = 0;
std::vector<int64_t> g_rank_lo, g_rank_hi;  // 1-based inclusive column range of every rank
int64_t g_comm_bytes = 0;                   // bytes this rank sent + received through NCCL in the last Jacobian

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      g_err = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call;                 \
      return -10;                                                                                  \
    }                                                                                              \
  } while (0)

// Launch on the library stream.  (Programmatic dependent launch was measured and rejected: with griddepcontrol
// prologues and programmatic graph edges the d3dHsm step went from 0.174 to 0.194 ms.)
template <typename... KArgs, typename... Args>
cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = g_stream;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename T>
int dev_copy(const T* h, size_t n, const T** out) {
  T* p = nullptr;
  CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
  if (n) CK(cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice));
  g_static_allocs.push_back(p);
  *out = p;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// full-residual kernels
// ------------------------------------------------------------------------------------------------
// yl may be the caller's pinned host array (zero-copy read over the bus): then yl_keep receives the device copy that
// the later phases and the Jacobian read (yl_keep == nullptr: yl already is that copy).
__global__ void k_phase0(double* base, const double* __restrict__ yl, double* __restrict__ yl_keep, int64_t neq, int NXS, int NC, int* err) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= NC) return;
  Acc<false> a; a.base = base; a.NXS = NXS; a.NC = NC;
  const int nv = NVX;
  double ycell[UE_NV] = {0., 0., 0., 0., 0.};
#pragma unroll
  for (int k = 0; k < UE_NV; ++k) if (k < nv) ycell[k] = yl[(size_t)c * nv + k];
  if (yl_keep) {
#pragma unroll
    for (int k = 0; k < UE_NV; ++k) if (k < nv) yl_keep[(size_t)c * nv + k] = ycell[k];
    if (c == 0) { yl_keep[neq] = yl[neq]; yl_keep[neq + 1] = yl[neq + 1]; }
  }
  phase0_cell<false>(a, ycell, c % NXS, c / NXS, err);
}
// phase 1: 32 cells per block, one ROLE per warp (lane = cell); 1b reads only same-cell outputs of 1a
__global__ void __launch_bounds__(160) k_phase1(double* base, int NXS, int NC) {
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  Acc<false> a; a.base = base; a.NXS = NXS; a.NC = NC;
  const Win w = make_win(D, -1, -1);
  const int ix = c % NXS, iy = c / NXS;
  if (c < NC) {
    if (role == 0) p1_xpart<false>(a, w, ix, iy);
    else if (role == 1) p1_ypart<false>(a, w, ix, iy);
    else if (role == 2) p1_visx<false>(a, w, ix, iy);
  }
  __syncthreads();
  if (c < NC) {
    if (role == 0) p1_fx<false>(a, w, ix, iy);
    else if (role == 1) p1_fy<false>(a, w, ix, iy);
    else if (role == 2) p1_exe<false>(a, w, ix, iy);
    else if (role == 3) p1_exi<false>(a, w, ix, iy);
    else p1_ey<false>(a, w, ix, iy);
  }
}
// phase 2: 32 cells per block, four role-warps (equation groups on interior cells) plus a guard-row warp (bouncon).
// The guard warp does not take the block's own cells: it takes 32 entries of a list of all guard cells sorted by kind
// (bottom, top, corner, left plate, right plate) and padded per kind to whole warps, so that a warp runs ONE kind's
// code instead of all of them one after the other.
__global__ void __launch_bounds__(160) k_phase2(double* base, double* __restrict__ tmp, int NXS, int NC, const int* __restrict__ guard_cells, int nguard) {
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  Acc<false> a; a.base = base; a.NXS = NXS; a.NC = NC;
  const Win w = make_win(D, -1, -1);
  double r[UE_NV] = {0., 0., 0., 0., 0.};
  if (role == 4) {
    const int g = blockIdx.x * 32 + lane;
    if (g >= nguard) return;
    const int c = guard_cells[g];
    if (c < 0) return;
    phase2_guard<false>(a, w, c % NXS, c / NXS, r);
    double* o = tmp + (size_t)c * UE_NV;
    for (int k = 0; k < UE_NV; ++k) o[k] = r[k];
    return;
  }
  const int c = blockIdx.x * 32 + lane;
  if (c >= NC) return;
  const int ix = c % NXS, iy = c / NXS;
  double* o = tmp + (size_t)c * UE_NV;
  if (ix >= 1 && ix <= D.nx && iy >= 1 && iy <= D.ny) {
    if (role == 0) { p2_n<false>(a, ix, iy, r, D.iseqalg); o[0] = r[0]; o[4] = r[4]; }
    else if (role == 1) { p2_m<false>(a, w, ix, iy, r, D.iseqalg); o[1] = r[1]; }
    else if (role == 2) { p2_e<false>(a, ix, iy, r, D.iseqalg); o[2] = r[2]; }
    else { p2_i<false>(a, ix, iy, r, D.iseqalg); o[3] = r[3]; }
  }
}
// Phases 2 and 3 of the full residual in one launch.  As k_phase2, plus: the four equation-group warps leave their rows
// in shared memory, and after a barrier among them the first warp applies rscalf and the time-step term to its 32 cells
// (phase3_interior; the particle balance of the east neighbour, which another block may own, is recomputed with the
// same device function) and writes yldot; the guard warp finishes its own rows.  Saves one launch per residual.
__global__ void __launch_bounds__(160) k_phase23(double* base, double* __restrict__ tmp, double* __restrict__ yldot, const double* __restrict__ yl,
                                                 const double* __restrict__ dtuse, const double* __restrict__ ylodt, int64_t neq, int NXS, int NC,
                                                 const int* __restrict__ guard_cells, int nguard, int* err, long long* hflags, double* __restrict__ yldot_host) {
  __shared__ double srow[UE_NV][32];
  const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
  Acc<false> a; a.base = base; a.NXS = NXS; a.NC = NC;
  const Win w = make_win(D, -1, -1);
  const int nv = NVX;
  if (blockIdx.x == 0 && threadIdx.x == 0) { hflags[0] = err[0]; err[0] = 0; }  // error bits of phases 0-1, as k_phase3
  double r[UE_NV] = {0., 0., 0., 0., 0.};
  if (role == 4) {
    const int g = blockIdx.x * 32 + lane;
    if (g >= nguard) return;
    const int c = guard_cells[g];
    if (c < 0) return;
    const int ix = c % NXS, iy = c / NXS;
    phase2_guard<false>(a, w, ix, iy, r);
    double* o = tmp + (size_t)c * UE_NV;
    for (int k = 0; k < UE_NV; ++k) o[k] = r[k];
    if (D.isbcwdt == 1) phase3_dt(ix, iy, r, yl + (size_t)c * nv, yl[neq], (int64_t)c * nv, dtuse, ylodt);
#pragma unroll
    for (int k = 0; k < UE_NV; ++k) if (k < nv) { yldot[(size_t)c * nv + k] = r[k]; if (yldot_host) yldot_host[(size_t)c * nv + k] = r[k]; }
    return;
  }
  const int c = blockIdx.x * 32 + lane;
  const int ix = c % NXS, iy = c / NXS;
  const bool interior = c < NC && ix >= 1 && ix <= D.nx && iy >= 1 && iy <= D.ny;
  if (interior) {
    double* o = tmp + (size_t)c * UE_NV;
    if (role == 0) { p2_n<false>(a, ix, iy, r, D.iseqalg); o[0] = r[0]; o[4] = r[4]; srow[0][lane] = r[0]; srow[4][lane] = r[4]; }
    else if (role == 1) { p2_m<false>(a, w, ix, iy, r, D.iseqalg); o[1] = r[1]; srow[1][lane] = r[1]; }
    else if (role == 2) { p2_e<false>(a, ix, iy, r, D.iseqalg); o[2] = r[2]; srow[2][lane] = r[2]; }
    else { p2_i<false>(a, ix, iy, r, D.iseqalg); o[3] = r[3]; srow[3][lane] = r[3]; }
  }
  asm volatile("bar.sync 1, 128;" ::: "memory");  // the four equation-group warps only (the guard warp may have left)
  if (role == 0 && interior) {
#pragma unroll
    for (int k = 0; k < UE_NV; ++k) r[k] = srow[k][lane];
    phase3_interior<false>(a, ix, iy, r, yl + (size_t)c * nv, yl[neq], D.iseqalg, dtuse, ylodt, true);
#pragma unroll
    for (int k = 0; k < UE_NV; ++k) if (k < nv) { yldot[(size_t)c * nv + k] = r[k]; if (yldot_host) yldot_host[(size_t)c * nv + k] = r[k]; }
  }
}
__global__ void k_phase3(double* base, const double* __restrict__ tmp, double* __restrict__ yldot, const double* __restrict__ yl,
                         const double* __restrict__ dtuse, const double* __restrict__ ylodt, int64_t neq, int NXS, int NC, int* err,
                         long long* hflags, double* __restrict__ yldot_host /* caller's pinned host array or nullptr */) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0) { hflags[0] = err[0]; err[0] = 0; }  // error bits of the residual sequence (err[0]) go straight to mapped host memory (no copy node) and are cleared
  if (c >= NC) return;
  Acc<false> a; a.base = base; a.NXS = NXS; a.NC = NC;
  const int ix = c % NXS, iy = c / NXS;
  const int nv = NVX;  // unknowns per cell in the caller's vectors (4 when isngon = 0); tmp keeps UE_NV row slots
  double r[UE_NV];
  for (int k = 0; k < UE_NV; ++k) r[k] = tmp[(size_t)c * UE_NV + k];
  if (ix >= 1 && ix <= D.nx && iy >= 1 && iy <= D.ny) phase3_interior<false>(a, ix, iy, r, yl + (size_t)c * nv, yl[neq], D.iseqalg, dtuse, ylodt);
  else if (D.isbcwdt == 1) phase3_dt(ix, iy, r, yl + (size_t)c * nv, yl[neq], (int64_t)c * nv, dtuse, ylodt);  // guard rows carry the term too
#pragma unroll
  for (int k = 0; k < UE_NV; ++k) if (k < nv) yldot[(size_t)c * nv + k] = r[k];
  if (yldot_host) {
#pragma unroll
    for (int k = 0; k < UE_NV; ++k) if (k < nv) yldot_host[(size_t)c * nv + k] = r[k];
  }
}

// ------------------------------------------------------------------------------------------------
// Batched Jacobian (jac_calc, oderhs.m:8533-8760).  Every perturbed unknown of the column range is processed at
// once.  A perturbation of cell C0 = (xc,yc) changes phase-0 fields only at C0 and phase-1 fields only at C0, its
// west/east neighbours and its south neighbour: those four cells get private copies; the residual rows that can
// change are a short CANDIDATE LIST of cells around them (host-built from the index maps, so X-point cuts need no
// special windows).  Each phase is one launch whose blocks all run the SAME role function (blockIdx.y = role) over
// 128 items: the instruction stream of a role is fetched once per SM and shared by all its warps.  Private cells
// and candidate rows live in global memory (L2-resident, ~1.6 KB per unknown).
//   item (u, k)  : unknown u, private slot k (C0, Cw, Ce, Cs)    -> phases 0, 1a, 1b
//   item (u, l)  : unknown u, candidate cell l of its list        -> phases 2, 3, compaction
// Rows outside the list cannot change and difference to exactly zero in the reference (jaccliplim = 0 drops them).
// ------------------------------------------------------------------------------------------------
struct UInfo {
  Win w;
  int iv, xc, yc, xw, xe;
  int off;   // first candidate item of this unknown (rows / rres / rmask)
  int n;     // number of candidate cells
  int coff;  // first entry of its cell's candidate list in cand_cell / cand_east
};
struct JArgs {
  const UInfo* ui;
  const int *cand_cell, *cand_east, *item_u, *guard_items;  // guard_items: candidate items whose cell is a guard cell
  int NU, nitems, nguard, role0;
  double *priv, *rows, *rres;
  int* rmask;
  double* base;
  const double *yl, *yldot00, *suscal, *sfscal, *dtuse, *ylodt;
  int64_t neq, ml, mu;
  int NXS, NC;
  const int64_t* coloff;
  int *colcnt, *colrow;
  double* colval;
  int *rowcnt, *err;
};
__device__ __forceinline__ Acc<true> jb_acc(const JArgs& A, const UInfo& q, int u) {
  Acc<true> a;
  a.base = A.base; a.NXS = A.NXS; a.NC = A.NC;
  a.priv = A.priv + (size_t)u * 4 * PL_COUNT; a.ps = 1; a.ks = PL_COUNT;
  a.xc = q.xc; a.yc = q.yc; a.xw = q.xw; a.xe = q.xe;
  a.rres = A.rres + q.off; a.rmask = A.rmask + q.off;
  a.rself = -1; a.reast = -1;
  return a;
}
__device__ __forceinline__ bool jb_slot_cell(const UInfo& q, int k, int& ix, int& iy) {  // false: slot duplicates another or does not exist
  iy = q.yc; ix = q.xc;
  if (k == 1) { ix = q.xw; return q.xw != q.xc; }
  if (k == 2) { ix = q.xe; return q.xe != q.xc && q.xe != q.xw; }
  if (k == 3) { iy = q.yc - 1; return q.yc >= 1; }
  return true;
}
__device__ __forceinline__ double jb_dyl(const JArgs& A, const UInfo& q, double& yold) {  // oderhs.m:8676-8678
  yold = A.yl[q.iv - 1];
  return D.delpert * (fabs(yold) + D.dylconst / A.suscal[q.iv - 1]);
}
// stage the private cells of 32 unknowns from the base planes, then phase 0 on their perturbed cells
__global__ void __launch_bounds__(128) k_jb_stage0(JArgs A) {
  const int u0 = blockIdx.x * 32, tid = threadIdx.x;
  {  // clear this Jacobian's counters (colcnt | rowcnt | rowfill are contiguous) and the candidate-row masks
    const int nthr = gridDim.x * 128, t0 = blockIdx.x * 128 + tid;
    for (int64_t i = t0; i < 3 * A.neq; i += nthr) A.colcnt[i] = 0;
    for (int i = t0; i < A.nitems; i += nthr) A.rmask[i] = 0;
  }
  const int u = u0 + (tid >> 2), k = tid & 3;
  if (u < A.NU) {
    const UInfo& q = A.ui[u];
    int ix, iy;
    jb_slot_cell(q, k, ix, iy);
    if (iy < 0) iy = 0;
    const int cell = ix + A.NXS * iy;
    double v[PL_COUNT];
#pragma unroll
    for (int pl = 0; pl < PL_COUNT; ++pl) v[pl] = A.base[(size_t)pl * A.NC + cell];  // all loads in flight together
    double* dst = A.priv + ((size_t)u * 4 + k) * PL_COUNT;
#pragma unroll
    for (int pl = 0; pl < PL_COUNT; ++pl) dst[pl] = v[pl];
  }
  __syncthreads();
  const int up = u0 + tid;
  if (tid < 32 && up < A.NU) {
    const UInfo& q = A.ui[up];
    const Acc<true> a = jb_acc(A, q, up);
    double ycell[UE_NV] = {0., 0., 0., 0., 0.}, yold;
    const double dyl = jb_dyl(A, q, yold);
    const int64_t c = (int64_t)(q.xc + A.NXS * q.yc) * NVX;
#pragma unroll
    for (int k2 = 0; k2 < UE_NV; ++k2) if (k2 < NVX) ycell[k2] = A.yl[c + k2];
    ycell[(q.iv - 1) - c] = yold + dyl;
    phase0_cell<true>(a, ycell, q.xc, q.yc, A.err);
  }
}
// Small grids: phases 0, 1a and 1b of 32 unknowns in ONE block (all dependencies are per unknown, so block-level
// barriers suffice): 5 role groups of P01_ITEMS (unknown, slot) items.  Saves two launches and their cold
// instruction/constant/L1 misses; the private cells stay in this SM's L1 from staging to the last phase-1 role
// (d3dHsm: Jacobian sequence 70 -> 63 us warm; no change with L2 flushed).  Large grids keep one launch per phase.
constexpr int P01_ITEMS = 64;  // (unknown, slot) items per block of k_jb_p01: 16 unknowns (8 and 32 per block measured slower)
__global__ void __launch_bounds__(5 * P01_ITEMS, 1) k_jb_p01(JArgs A) {
  const int u0 = blockIdx.x * (P01_ITEMS / 4), tid = threadIdx.x;
  {  // clear this Jacobian's counters (colcnt | rowcnt | rowfill are contiguous) and the candidate-row masks
    const int nthr = gridDim.x * 5 * P01_ITEMS, t0 = blockIdx.x * 5 * P01_ITEMS + tid;
    for (int64_t i = t0; i < 3 * A.neq; i += nthr) A.colcnt[i] = 0;
    for (int i = t0; i < A.nitems; i += nthr) A.rmask[i] = 0;
  }
  const int grp = tid / P01_ITEMS, j = tid % P01_ITEMS;          // role group, item within the block
  const int u = u0 + (j >> 2), k = j & 3;
  const bool live = u < A.NU;
  int ix = 0, iy = 0;
  bool slot_ok = false;
  if (live) slot_ok = jb_slot_cell(A.ui[u], k, ix, iy);
  if (live) {  // staging: the five groups split the planes
    const int cy = iy < 0 ? 0 : iy;
    const int cell = ix + A.NXS * cy;
    double* dst = A.priv + ((size_t)u * 4 + k) * PL_COUNT;
    for (int pl = grp; pl < PL_COUNT; pl += 5) dst[pl] = A.base[(size_t)pl * A.NC + cell];
  }
  __syncthreads();
  if (tid < P01_ITEMS / 4 && u0 + tid < A.NU) {
    const int up = u0 + tid;
    const UInfo& q = A.ui[up];
    const Acc<true> a = jb_acc(A, q, up);
    double ycell[UE_NV] = {0., 0., 0., 0., 0.}, yold;
    const double dyl = jb_dyl(A, q, yold);
    const int64_t c = (int64_t)(q.xc + A.NXS * q.yc) * NVX;
#pragma unroll
    for (int k2 = 0; k2 < UE_NV; ++k2) if (k2 < NVX) ycell[k2] = A.yl[c + k2];
    ycell[(q.iv - 1) - c] = yold + dyl;
    phase0_cell<true>(a, ycell, q.xc, q.yc, A.err);
  }
  __syncthreads();
  if (live && slot_ok) {
    const UInfo& q = A.ui[u];
    const Acc<true> a = jb_acc(A, q, u);
    if (grp == 0) p1_xpart<true>(a, q.w, ix, iy);
    else if (grp == 1) p1_ypart<true>(a, q.w, ix, iy);
    else if (grp == 2) p1_visx<true>(a, q.w, ix, iy);
  }
  __syncthreads();
  if (live && slot_ok) {
    const UInfo& q = A.ui[u];
    const Acc<true> a = jb_acc(A, q, u);
    if (grp == 0) p1_fx<true>(a, q.w, ix, iy);
    else if (grp == 1) p1_fy<true>(a, q.w, ix, iy);
    else if (grp == 2) p1_exe<true>(a, q.w, ix, iy);
    else if (grp == 3) p1_exi<true>(a, q.w, ix, iy);
    else p1_ey<true>(a, q.w, ix, iy);
  }
}
__global__ void __launch_bounds__(128) k_jb_p1a(JArgs A) {
  const int it = blockIdx.x * 128 + threadIdx.x, u = it >> 2, k = it & 3;
  if (u >= A.NU) return;
  const UInfo& q = A.ui[u];
  int ix, iy;
  if (!jb_slot_cell(q, k, ix, iy)) return;
  const Acc<true> a = jb_acc(A, q, u);
  const int role = blockIdx.y + A.role0;
  if (role == 0) p1_xpart<true>(a, q.w, ix, iy);
  else if (role == 1) p1_ypart<true>(a, q.w, ix, iy);
  else p1_visx<true>(a, q.w, ix, iy);
}
// MINB: resident blocks per SM the register allocation must allow.  Small grids run one latency-bound pass (all the
// registers the compiler wants); large grids are throughput-bound and gain from the higher occupancy despite a few spills.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_jb_p1b(JArgs A) {
  const int it = blockIdx.x * 128 + threadIdx.x, u = it >> 2, k = it & 3;
  if (u >= A.NU) return;
  const UInfo& q = A.ui[u];
  int ix, iy;
  if (!jb_slot_cell(q, k, ix, iy)) return;
  const Acc<true> a = jb_acc(A, q, u);
  const int role = blockIdx.y + A.role0;
  if (role == 0) p1_fx<true>(a, q.w, ix, iy);
  else if (role == 1) p1_fy<true>(a, q.w, ix, iy);
  else if (role == 2) p1_exe<true>(a, q.w, ix, iy);
  else if (role == 3) p1_exi<true>(a, q.w, ix, iy);
  else p1_ey<true>(a, q.w, ix, iy);
}
// phase 2 on the candidate rows; roles 0-3 = equation groups on interior rows, role 4 = guard rows.  rows[k][item], rmask[item]
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_jb_p2(JArgs A) {
  const int ry = blockIdx.y + A.role0;
  const int role = ry == 0 ? 4 : ry - 1;  // the guard role (long, divergent rows) is dispatched first, the equation groups follow
  int it = blockIdx.x * 128 + threadIdx.x;
  if (role == 4) {  // the guard role walks the (short) list of guard items only
    if (it >= A.nguard) return;
    it = A.guard_items[it];
    if (it < 0) return;  // padding between the kinds
  }
  if (it >= A.nitems) return;
  const int u = A.item_u[it];
  const UInfo& q = A.ui[u];
  const int l = it - q.off;
  const int cell = A.cand_cell[q.coff + l];
  const int ix = cell % A.NXS, iy = cell / A.NXS;
  Acc<true> a = jb_acc(A, q, u);
  a.rself = l;
  double r[UE_NV] = {0., 0., 0., 0., 0.};
  double* o = A.rows + it;
  const size_t NI = A.nitems;
  int* mk = A.rmask + it;
  if (ix >= 1 && ix <= D.nx && iy >= 1 && iy <= D.ny) {
    if (in_rng(ix, q.w.i2, q.w.i5) && in_rng(iy, q.w.j2, q.w.j5)) {
      if (role == 0) { p2_n<true>(a, ix, iy, r, D.iseqalg); o[0] = r[0]; o[4 * NI] = r[4]; atomicOr(mk, 0x111); }
      else if (role == 1) { p2_m<true>(a, q.w, ix, iy, r, D.iseqalg); o[1 * NI] = r[1]; atomicOr(mk, 0x2); }
      else if (role == 2) { p2_e<true>(a, ix, iy, r, D.iseqalg); o[2 * NI] = r[2]; atomicOr(mk, 0x4); }
      else if (role == 3) { p2_i<true>(a, ix, iy, r, D.iseqalg); o[3 * NI] = r[3]; atomicOr(mk, 0x8); }
    }
  } else if (role == 4) {
    const int m = phase2_guard<true>(a, q.w, ix, iy, r);
    for (int k = 0; k < UE_NV; ++k) o[k * NI] = r[k];
    atomicOr(mk, m);
  }
}
// phase 3 on the interior candidate rows, then difference / clip / ordered compaction into the column's CSC
// fragment (oderhs.m:8685-8719): one warp per unknown
// (Building the row pointer in the last block of this kernel instead of a separate k_scan launch was measured: no gain.)
__global__ void __launch_bounds__(128) k_jb_p3c(JArgs A) {
  const int lane = threadIdx.x & 31, u = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (u >= A.NU) return;
  const UInfo& q = A.ui[u];
  Acc<true> a = jb_acc(A, q, u);
  const size_t NI = A.nitems;
  const int NXS = A.NXS;
  const int64_t neq = A.neq;
  double yold;
  const double dyl = jb_dyl(A, q, yold);
  for (int l = lane; l < q.n; l += 32) {
    const int cell = A.cand_cell[q.coff + l];
    const int ix = cell % NXS, iy = cell / NXS;
    if (ix >= 1 && ix <= D.nx && iy >= 1 && iy <= D.ny && in_rng(ix, q.w.i2, q.w.i5) && in_rng(iy, q.w.j2, q.w.j5)) {
      double r[UE_NV], ycell[UE_NV] = {0., 0., 0., 0., 0.};
      double* o = A.rows + q.off + l;
      const int64_t c = (int64_t)cell * NVX;
      for (int k = 0; k < UE_NV; ++k) r[k] = o[k * NI];
#pragma unroll
      for (int k = 0; k < UE_NV; ++k) if (k < NVX) ycell[k] = A.yl[c + k];
      if (ix == q.xc && iy == q.yc) ycell[(q.iv - 1) - c] = yold + dyl;
      a.reast = A.cand_east[q.coff + l];
      phase3_interior<true>(a, ix, iy, r, ycell, A.yl[neq], D.iseqalg, A.dtuse, A.ylodt);
      for (int k = 0; k < UE_NV; ++k) o[k * NI] = r[k];
    }
  }
  __syncwarp();
  const int64_t iv = q.iv;
  const int64_t ii1 = max(iv - A.mu, (int64_t)1), ii2 = min(iv + A.ml, neq);
  const double sf = A.sfscal[iv - 1];
  const int nv = NVX;
  const int ncand = q.n * nv;
  const int64_t o = A.coloff[iv - 1];
  int nout = 0;
  constexpr int CH = 4;  // candidates per lane and pass: their loads are issued together, the ordered ballots follow
  for (int q0 = 0; q0 < ncand; q0 += 32 * CH) {
    bool keep[CH]; double val[CH]; int64_t ii[CH];
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const int qq = q0 + 32 * c + lane;
      keep[c] = false; val[c] = 0.; ii[c] = 0;
      if (qq < ncand) {
        const int l = nv == 5 ? qq / 5 : (nv == 4 ? qq >> 2 : qq / nv), k = qq - l * nv;  // constant divisors for the two layouts
        ii[c] = (int64_t)A.cand_cell[q.coff + l] * nv + k + 1;
        if (ii[c] >= ii1 && ii[c] <= ii2) {
          const bool written = (A.rmask[q.off + l] >> k) & 1;
          if (written || ii[c] == iv) {
            const double y00 = A.yldot00[ii[c] - 1];
            const double wk = written ? A.rows[(size_t)k * NI + q.off + l] : y00;
            double jacelem = (wk - y00) / dyl;
            if (ii[c] == iv) {
              if (D.iseqalg[iv - 1] * (1 - D.isbcwdt) == 0) jacelem = jacelem - 1 / A.dtuse[iv - 1];
              if (D.nufak > 0 && A.yl[neq] == 1) jacelem = jacelem - D.nufak;
            }
            val[c] = jacelem;
            keep[c] = fabs(jacelem * sf) > D.jaccliplim;
          }
        }
      }
    }
#pragma unroll
    for (int c = 0; c < CH; ++c) {
      const unsigned bal = __ballot_sync(0xffffffffu, keep[c]);
      if (keep[c]) {
        const int pos = nout + __popc(bal & ((1u << lane) - 1));
        A.colrow[o + pos] = (int)ii[c];
        A.colval[o + pos] = val[c];
        atomicAdd(&A.rowcnt[ii[c] - 1], 1);
      }
      nout += __popc(bal);
    }
  }
  if (lane == 0) A.colcnt[iv - 1] = nout;
}

// ---- CSC fragments -> CSR ---------------------------------------------------------------------------------------
__global__ void k_scan(const int* __restrict__ rowcnt, int64_t* __restrict__ ia, int64_t n, int* err, long long* hflags, int64_t* __restrict__ ia_host) {
  // single block of 1024 threads; ia is 1-based: ia[0] = 1, ia[i+1] = ia[i] + rowcnt[i].
  // Each thread owns a contiguous chunk: serial sum, block scan of the 1024 partials, serial write-out.
  __shared__ int64_t s[1024];
  const int t = threadIdx.x;
  const int64_t chunk = (n + 1023) / 1024, b0 = t * chunk, b1 = min(n, b0 + chunk);
  int64_t sum = 0;
  for (int64_t i = b0; i < b1; ++i) sum += rowcnt[i];
  s[t] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int64_t v = (t >= off) ? s[t - off] : 0;
    __syncthreads();
    s[t] += v;
    __syncthreads();
  }
  int64_t run = 1 + s[t] - sum;
  if (t == 0) { ia[0] = 1; if (ia_host) ia_host[0] = 1; }
  for (int64_t i = b0; i < b1; ++i) { run += rowcnt[i]; ia[i + 1] = run; if (ia_host) ia_host[i + 1] = run; }
  if (t == 1023) { hflags[1] = 1 + s[1023]; hflags[2] = err[0] | err[1]; err[0] = err[1] = 0; }  // nnz + 1 and the error bits (err[1]: Jacobian sequence) to mapped host memory
}
__global__ void k_fill(int64_t neq, int64_t ivmin, int64_t ivmax, const int64_t* __restrict__ coloff, const int* __restrict__ colcnt,
                       const int* __restrict__ colrow, const double* __restrict__ colval, const int64_t* __restrict__ ia, int* __restrict__ rowfill,
                       double* __restrict__ jac, int64_t* __restrict__ ja, int64_t nnzmx) {
  const int64_t iv = ivmin + blockIdx.x;
  if (iv > ivmax) return;
  const int64_t o = coloff[iv - 1];
  const int n = colcnt[iv - 1];
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int row = colrow[o + e];
    const int64_t p = ia[row - 1] - 1 + atomicAdd(&rowfill[row - 1], 1);
    if (p < nnzmx) { ja[p] = iv; jac[p] = colval[o + e]; }   // overflow is reported by the host after the sequence
  }
}
// one warp per row: rank sort of the row's entries by column (rows hold a few tens of entries).  jac_host/ja_host: the
// caller's page-locked arrays (or nullptr): the sorted row is also written there, so no copy follows the sequence.
// Rows longer than CAP (a dense row: the integrated core-power row, or the electron-energy rows on a half-space cut, which
// couple to every column of their window) are rank-sorted through a global scratch area (the CSC fragment buffers, free
// again once k_fill has run) instead of shared memory.
__global__ void k_sortrows(int64_t neq, const int64_t* __restrict__ ia, double* __restrict__ jac, int64_t* __restrict__ ja, int64_t nnzmx,
                           double* __restrict__ jac_host, int64_t* __restrict__ ja_host, int* __restrict__ scr_col, double* __restrict__ scr_val,
                           int64_t scr_cap) {
  constexpr int CAP = 96;
  __shared__ int64_t scol[4][CAP];
  __shared__ double sval[4][CAP];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int64_t r = (int64_t)blockIdx.x * 4 + wib;
  if (r >= neq) return;
  const int64_t b = ia[r] - 1, e = ia[r + 1] - 1;
  const int n = (int)(e - b);
  if (n <= 0 || e > nnzmx) return;
  if (n == 1) {
    if (jac_host && lane == 0) { ja_host[b] = ja[b]; jac_host[b] = jac[b]; }
  } else if (n <= CAP) {
    for (int i = lane; i < n; i += 32) { scol[wib][i] = ja[b + i]; sval[wib][i] = jac[b + i]; }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {
      const int64_t c = scol[wib][i];
      int rank = 0;
      for (int j = 0; j < n; ++j) rank += (scol[wib][j] < c);
      ja[b + rank] = c; jac[b + rank] = sval[wib][i];
    }
    if (jac_host) {  // second pass in storage order: contiguous writes over the bus
      __syncwarp();
      for (int i = lane; i < n; i += 32) { ja_host[b + i] = ja[b + i]; jac_host[b + i] = jac[b + i]; }
    }
  } else if (e <= scr_cap) {
    for (int64_t i = b + lane; i < e; i += 32) { scr_col[i] = (int)ja[i]; scr_val[i] = jac[i]; }
    __syncwarp();
    for (int64_t i = b + lane; i < e; i += 32) {
      const int c = scr_col[i];
      int rank = 0;
      for (int64_t j = b; j < e; ++j) rank += (scr_col[j] < c);
      ja[b + rank] = c; jac[b + rank] = scr_val[i];
    }
    if (jac_host) {
      __syncwarp();
      for (int64_t i = b + lane; i < e; i += 32) { ja_host[i] = ja[i]; jac_host[i] = jac[i]; }
    }
  } else {
    if (lane == 0) {  // no scratch space: serial insertion sort
      for (int64_t i = b + 1; i < e; ++i) {
        const int64_t cj = ja[i]; const double cv = jac[i];
        int64_t j = i - 1;
        while (j >= b && ja[j] > cj) { ja[j + 1] = ja[j]; jac[j + 1] = jac[j]; --j; }
        ja[j + 1] = cj; jac[j + 1] = cv;
      }
    }
    if (jac_host) {
      __syncwarp();
      for (int64_t i = b + lane; i < e; i += 32) { ja_host[i] = ja[i]; jac_host[i] = jac[i]; }
    }
  }
}

// sfsetnk scaling chain (oderhs.m:9862-9881): column scaling by 1/su (amudia, svr/svrut4.m:1104-1130), row max-norm
// (rnrms with normtype=0, svr/svrut4.m:1002-1052), sf = 1/norm, and ydt_max0 = max|yldot0*sf|.  One warp per row.
// psetnk scaling chain, one thread per row (rows are <= ~80 entries; the 1- and 2-norms are serial sums in the reference's
// order so that the factors are bit-identical): amudia, diamua, roscal (svr/svrut4.m:954-1148)
__global__ void k_rowscale(int64_t neq, const int64_t* __restrict__ ia, const int64_t* __restrict__ ja, double* __restrict__ jac,
                           const double* __restrict__ su, const double* __restrict__ sf, int isrnorm, int normtype, double* __restrict__ fac) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= neq) return;
  const int64_t k1 = ia[i] - 1, k2 = ia[i + 1] - 1;
  const double sfi = sf[i];
  double scal = 0.;
  for (int64_t k = k1; k < k2; ++k) {
    double a = jac[k] * (1. / su[ja[k] - 1]);
    a = a * sfi;
    jac[k] = a;
    if (normtype == 0) scal = fmax(scal, fabs(a));
    else if (normtype == 1) scal = scal + fabs(a);
    else scal = scal + a * a;
  }
  if (!isrnorm) { fac[i] = 1.; return; }
  if (normtype == 2) scal = sqrt(scal);
  const double d = 1.0 / scal;
  fac[i] = d;
  for (int64_t k = k1; k < k2; ++k) jac[k] = jac[k] * d;
}

// include/ue_math.h on the device (parity probe: must equal the host evaluation bit for bit)
__global__ void k_math_probe(int op, int64_t n, const double* __restrict__ x, const double* __restrict__ y, double* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double r = 0.;
  switch (op) {
    case 0: r = ue_exp(x[i]); break;
    case 1: r = ue_log(x[i]); break;
    case 2: r = ue_log10(x[i]); break;
    case 3: r = ue_pow(x[i], y[i]); break;
    case 4: r = ue_cos(x[i]); break;
    default: r = ue_sqrt(x[i]); break;
  }
  out[i] = r;
}

// yldot00 must be the residual of yl bit for bit (see ue_gpu_jac_calc)
// set_dt (oderhs.m:9886-10147), model_dt 0..3: one thread per cell; f0 is the residual just evaluated.  dtoptv persists
// between calls (a velocity row with |f0| <= cutlo keeps its previous value).
__device__ __forceinline__ double d_dtmodel(double dtopt) {
  if (D.model_dt == 0) return D.dtreal;
  if (D.model_dt == 1) return D.dtreal * dtopt / (D.dtreal + dtopt);
  if (D.model_dt == 2) return dtopt;
  return sqrt(D.dtreal * dtopt);
}
__global__ void k_set_dt(const double* __restrict__ f0, const double* __restrict__ ylodt, double* __restrict__ dtoptv, double* __restrict__ dtuse,
                         double* __restrict__ dtuse_host, int NXS, int NC) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= NC) return;
  const int nv = NVX, nx = (int)D.nx, ny = (int)D.ny;
  const int ix = c % NXS, iy = c / NXS;
  const int iym1 = max(0, iy - 1), iyp1 = min(ny + 1, iy + 1);
#pragma unroll
  for (int k = 0; k < UE_NV; ++k) {
    if (k >= nv) continue;
    const int64_t iv = (int64_t)c * nv + k;
    bool wr = true;
    if (k == 1) {
      wr = (ix != nx + 2 * D.isbcwdt);
      if (wr) {
        const int ixm1u = max(0, IXM1(ix, iy)), ixp1u = min(nx + 1, IXP1(ix, iy));
        const double up_5ca = (fabs(ylodt[iv]) + fabs(ylodt[d_iv(ixm1u, iy, 1, NXS)]) + fabs(ylodt[d_iv(ixp1u, iy, 1, NXS)]) + fabs(ylodt[d_iv(ix, iyp1, 1, NXS)]) +
                               fabs(ylodt[d_iv(ix, iym1, 1, NXS)])) / 5;
        if (fabs(f0[iv]) > D.cutlo) dtoptv[iv] = D.deldt * fabs(up_5ca / (f0[iv]));
      }
    } else dtoptv[iv] = D.deldt * fabs(ylodt[iv] / (f0[iv] + D.cutlo));
    double dt = wr ? d_dtmodel(dtoptv[iv]) : dtuse[iv];
    if (D.isbcwdt == 0 && D.iseqalg[iv] == 1) dt = 1.e20;
    dtuse[iv] = dt;
    dtuse_host[iv] = dt;
  }
}
__global__ void k_samebits(const double* __restrict__ a, const double* __restrict__ b, int64_t n, int* __restrict__ err) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && __double_as_longlong(a[i]) != __double_as_longlong(b[i])) atomicOr(err, 4);
}

__global__ void k_rownorm(int64_t neq, const int64_t* __restrict__ ia, const int64_t* __restrict__ ja, const double* __restrict__ jac,
                          const double* __restrict__ su, const double* __restrict__ yldot0, double* __restrict__ sf, unsigned long long* ydtmax_bits, int* zero_row) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= neq) return;
  double m = 0.0;
  for (int64_t k = ia[r] - 1 + lane; k < ia[r + 1] - 1; k += 32) {
    const double t = 1. / su[ja[k] - 1];
    m = fmax(m, fabs(jac[k] * t));
  }
  for (int off = 16; off > 0; off >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, off));
  if (lane == 0) {
    if (fabs(m) < 1e20 * D.cutlo) { atomicMin(zero_row, (int)(r + 1)); sf[r] = 0.; return; }
    const double s = 1. / m;
    sf[r] = s;
    const double v = fabs(yldot0[r] * s);
    atomicMax(ydtmax_bits, (unsigned long long)__double_as_longlong(v));  // v >= 0: bit pattern is monotone
  }
}

}  // namespace (kernels)
namespace {
// ------------------------------------------------------------------------------------------------
int check_switches() {
  const UeParams& P = S.p;
  { const std::string z = S.nonzero_frozen(); if (!z.empty()) { g_err = "input " + z + " must be 0: the term it switches on is outside the built hot path"; return -5; } }
  struct { const char* n; int64_t v, want; } eq[] = {
      {"nisp", P.nisp, 1}, {"nusp", P.nusp, 1}, {"ngsp", P.ngsp, 1}, {"numvar", P.numvar, 4 + (P.isngon == 1)}, {"isnonog", P.isnonog, 0}, {"isphion", P.isphion, 0},
      {"isphiofft", P.isphiofft, 0}, {"isimpon", P.isimpon, 0}, {"isupgon", P.isupgon, 0}, {"istgon", P.istgon, 0},
      {"ineudif", P.ineudif, 2}, {"isflxvar", P.isflxvar, 0}, {"ismcnon", P.ismcnon, 0}, {"ifixsrc", P.ifixsrc, 0}, {"ifixpsor", P.ifixpsor, 0},
      {"ishymol", P.ishymol, 0}, {"ishosor", P.ishosor, 0}, {"isupdrag", P.isupdrag, 0}, {"isofric", P.isofric, 0}, {"jhswitch", P.jhswitch, 0},
      {"isfeexpl0", P.isfeexpl0, 0}, {"isfeixpl0", P.isfeixpl0, 0}, {"is1D_gbx", P.is1D_gbx, 0}, {"isnglf", P.isnglf, 0}, {"isudsym", P.isudsym, 0},
      {"islimon", P.islimon, 0}, {"isdifbetap", P.isdifbetap, 0}, {"isugfm1side", P.isugfm1side, 0}, {"nxomit", P.nxomit, 0},
      {"isfixrb", P.isfixrb, 0}, {"isextrnp", P.isextrnp, 0}, {"isextrnpf", P.isextrnpf, 0}, {"isextrtpf", P.isextrtpf, 0}, {"isextrngc", P.isextrngc, 0},
      {"isextrnw", P.isextrnw, 0}, {"isextrtw", P.isextrtw, 0}, {"isnfmiy", P.isnfmiy, 0}, {"isybdrywd", P.isybdrywd, 0}, {"isnewpot", P.isnewpot, 0},
      {"isbohmms", P.isbohmms, 0}, {"isgpye", P.isgpye, 0}, {"ibctepl", P.ibctepl, 1}, {"ibctipl", P.ibctipl, 1},
      {"ibctepr", P.ibctepr, 1}, {"ibctipr", P.ibctipr, 1}, {"iskaplex", P.iskaplex, 0}, {"isnupdot1sd", P.isnupdot1sd, 0}};
  for (auto& e : eq) if (e.v != e.want) { g_err = std::string("switch outside the built hot path: ") + e.n; return -5; }
  if (P.isbohmcalc != 0 && P.isbohmcalc != 1) { g_err = "isbohmcalc must be 0/1 with facb*=0"; return -5; }
  if (P.isnicore != 0 && P.isnicore != 1) { g_err = "isnicore must be 0 or 1"; return -5; }
  if (P.isupcore < 0 || P.isupcore > 3) { g_err = "isupcore must be 0..3"; return -5; }
  if (P.iflcore < -1 || P.iflcore > 1) { g_err = "iflcore must be -1, 0 or 1"; return -5; }
  if (P.isngcore < 0 || P.isngcore > 4) { g_err = "isngcore must be 0..4"; return -5; }
  if (P.istabon != 0 && P.istabon != 7 && P.istabon != 10) { g_err = "istabon must be 0, 7 or 10"; return -5; }
  if (P.isngon != 0 && P.isngon != 1) { g_err = "isngon must be 0 or 1"; return -5; }
  if (P.isfixlb != 0 && P.isfixlb != 2) { g_err = "isfixlb must be 0 or 2"; return -5; }
  if (P.isfixlb == 2 && (P.ixpt2 < 1 || P.ixpt2 > P.nx)) { g_err = "isfixlb=2 needs the cut ixpt2 inside the mesh"; return -5; }
  // fnnuiz < 1 blends the new ionisation rate with the value left by the PREVIOUS pandf call (oderhs.m:1950-1961): the
  // reference's Jacobian then depends on the order in which the unknowns were perturbed; not reproducible in parallel
  if (P.fnnuiz != 1.) { g_err = "fnnuiz must be 1 (history-dependent rate blending is outside the built hot path)"; return -5; }
  if (P.difpr2 != 0 || P.difni2 != 0 || P.difax != 0 || P.dif4order != 0 || P.kye4order != 0 || P.kyi4order != 0) { g_err = "difpr2/difni2/difax/4th-order terms not built"; return -5; }
  if (P.l_parloss <= 1e9) { g_err = "l_parloss<=1e9 (nuvl) not built"; return -5; }
  if (P.yinc >= 6 || P.xrinc >= 20) { g_err = "yinc>=6 / xrinc>=20 windows not built"; return -5; }
  for (int m : {(int)P.methn, (int)P.methu, (int)P.methe, (int)P.methi, (int)P.methg}) {
    int mx = m % 10, my = m / 10;
    if ((mx != 2 && mx != 3) || (my != 2 && my != 3)) { g_err = "meth* must use schemes 2 (central) or 3 (upwind)"; return -5; }
  }
  for (int ix = 0; ix < NXS; ++ix) {
    if (P.fngysi[ix] != 0 || P.fngyso[ix] != 0 || P.fngyi_use[ix] != 0 || P.fngyo_use[ix] != 0) { g_err = "wall gas sources not built"; return -5; }
    for (int64_t v : {P.isnwconiix[ix], P.isnwconoix[ix]}) if (v < 0 || v > 3) { g_err = "isnwconi/o must be 0..3"; return -5; }
    for (int64_t v : {P.istepfcix[ix], P.istipfcix[ix], P.istewcix[ix], P.istiwcix[ix]}) if (v < 0 || v > 3) { g_err = "istepfc/istipfc/istewc/istiwc must be 0..3"; return -5; }
  }
  for (int iy = 0; iy < ny + 2; ++iy)
    if (P.recylb[iy] < -1. || P.recyrb[iy] < -1.) { g_err = "recylb/recyrb < -1 not built"; return -5; }
  return 0;
}

void drop_graphs();
void free_all() {
  drop_graphs();
  g_seen_host.clear(); g_last_yldot.clear(); g_base_yl.clear();
  for (auto& h : g_step_host) h.clear();
  g_base_valid = g_base_dev_valid = false;
  for (void* p : g_static_allocs) cudaFree(p);
  g_static_allocs.clear();
  void* ptrs[] = {d_base, d_yl, d_yldot00, d_tmp, d_yldot, d_dtuse, d_ylodt, d_suscal, d_sfscal, d_err, d_cand_cell, d_cand_east, d_item_u, d_guard_items, d_guard_cells, d_coloff,
                  d_colcnt, d_colrow, d_colval, d_ia, d_ja, d_jac};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (d_dtoptv) { cudaFree(d_dtoptv); d_dtoptv = nullptr; }
  for (void* p : {(void*)d_uinfo, (void*)d_priv, (void*)d_jrows, (void*)d_rres, (void*)d_rmask}) if (p) cudaFree(p);
  d_uinfo = nullptr; d_priv = d_jrows = d_rres = nullptr; d_rmask = nullptr;
  d_base = d_yl = d_yldot00 = d_tmp = d_yldot = d_dtuse = d_ylodt = d_suscal = d_sfscal = nullptr;
  d_err = nullptr; d_cand_cell = d_cand_east = d_item_u = d_guard_items = d_guard_cells = nullptr; d_coloff = nullptr; d_colcnt = d_colrow = nullptr; d_colval = nullptr;
  d_rowcnt = d_rowfill = nullptr; d_ia = d_ja = nullptr; d_jac = nullptr;
  g_ready = false;
  g_alloc = false;
}

// kind of a guard cell = which part of bouncon sets its rows; lists are sorted by kind and padded to whole warps
int guard_kind(int ix, int iy) {
  const bool yb = (iy == 0 || iy == ny + 1), xb = (ix == 0 || ix == nx + 1);
  if (yb) return xb ? 2 : (iy == 0 ? 0 : 1);
  return ix == 0 ? 3 : 4;
}
template <typename KeyOf>
std::vector<int> sort_pad_by_kind(const std::vector<int>& items, KeyOf cell_of) {
  std::vector<int> out;
  for (int kind = 0; kind < 5; ++kind) {
    for (int it : items) { const int c = cell_of(it); if (guard_kind(c % NXS, c / NXS) == kind) out.push_back(it); }
    while (out.size() % 32) out.push_back(-1);
  }
  return out;
}

int build_lists() {
  const UeParams& P = S.p;
  // per-cell candidate lists and, for each entry, the list position of its east neighbour (rscalf reads the
  // density row of ixp1, oderhs.m:8140-8160)
  h_cellcand_off.assign(NC + 1, 0); h_cand_cell.clear(); h_cand_east.clear();
  std::vector<int> cand;
  for (int c = 0; c < NC; ++c) {
    cell_candidates(P, c % NXS, c / NXS, cand);
    h_cellcand_off[c] = (int)h_cand_cell.size();
    for (int cc : cand) {
      const int e = (int)P.ixp1[cc];  // same row
      const int ecell = e + NXS * (cc / NXS);
      const auto it = std::lower_bound(cand.begin(), cand.end(), ecell);
      h_cand_cell.push_back(cc);
      h_cand_east.push_back(it != cand.end() && *it == ecell ? (int)(it - cand.begin()) : -1);
    }
  }
  h_cellcand_off[NC] = (int)h_cand_cell.size();
  // unknowns of the column range and the capacity of every column's CSC fragment
  h_list.clear();
  h_coloff.assign(neq, 0);
  int64_t off = 0;
  for (int64_t iv = 1; iv <= neq; ++iv) {
    const int c = (int)P.igyl[iv - 1] + NXS * (int)P.igyl[neq + iv - 1];
    h_coloff[iv - 1] = off;
    off += (int64_t)(h_cellcand_off[c + 1] - h_cellcand_off[c]) * UE_NV;
    if (iv >= g_ivmin && iv <= g_ivmax) h_list.push_back((int)iv);
  }
  g_cap_total = off;
  return 0;
}

int upload_lists() {
  const UeParams& P = S.p;
  const size_t NU = h_list.size();
  std::vector<UInfo> ui(NU);
  std::vector<int> item_u, guard_items;
  for (size_t u = 0; u < NU; ++u) {
    UInfo& q = ui[u];
    q.iv = h_list[u];
    q.xc = (int)P.igyl[q.iv - 1]; q.yc = (int)P.igyl[neq + q.iv - 1];
    q.w = make_win(P, q.xc, q.yc);
    const int c = q.xc + NXS * q.yc;
    q.xw = (int)P.ixm1[c]; q.xe = (int)P.ixp1[c];
    q.coff = h_cellcand_off[c]; q.n = h_cellcand_off[c + 1] - h_cellcand_off[c];
    q.off = (int)item_u.size();
    item_u.insert(item_u.end(), q.n, (int)u);
    for (int l = 0; l < q.n; ++l) {
      const int cell = h_cand_cell[q.coff + l], ix = cell % NXS, iy = cell / NXS;
      if (!(ix >= 1 && ix <= nx && iy >= 1 && iy <= ny)) guard_items.push_back(q.off + l);
    }
  }
  guard_items = sort_pad_by_kind(guard_items, [&](int it) { const UInfo& q = ui[item_u[it]]; return h_cand_cell[q.coff + (it - q.off)]; });
  g_nguard = (int)guard_items.size();
  g_nitems = (int)item_u.size();
  for (void* p : {(void*)d_uinfo, (void*)d_priv, (void*)d_jrows, (void*)d_rres, (void*)d_rmask, (void*)d_cand_cell, (void*)d_cand_east, (void*)d_item_u, (void*)d_guard_items}) if (p) cudaFree(p);
  d_uinfo = nullptr; d_priv = d_jrows = d_rres = nullptr; d_rmask = nullptr; d_cand_cell = d_cand_east = d_item_u = d_guard_items = nullptr;
  CK(cudaMalloc(&d_uinfo, std::max<size_t>(1, NU) * sizeof(UInfo)));
  if (NU) CK(cudaMemcpy(d_uinfo, ui.data(), NU * sizeof(UInfo), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_cand_cell, std::max<size_t>(1, h_cand_cell.size()) * sizeof(int)));
  CK(cudaMalloc(&d_cand_east, std::max<size_t>(1, h_cand_east.size()) * sizeof(int)));
  CK(cudaMemcpy(d_cand_cell, h_cand_cell.data(), h_cand_cell.size() * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_cand_east, h_cand_east.data(), h_cand_east.size() * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_item_u, std::max<size_t>(1, item_u.size()) * sizeof(int)));
  if (!item_u.empty()) CK(cudaMemcpy(d_item_u, item_u.data(), item_u.size() * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_guard_items, std::max<size_t>(1, guard_items.size()) * sizeof(int)));
  if (!guard_items.empty()) CK(cudaMemcpy(d_guard_items, guard_items.data(), guard_items.size() * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_priv, std::max<size_t>(1, NU) * 4 * PL_COUNT * sizeof(double)));
  CK(cudaMalloc(&d_jrows, std::max<size_t>(1, (size_t)g_nitems) * UE_NV * sizeof(double)));
  CK(cudaMalloc(&d_rres, std::max<size_t>(1, (size_t)g_nitems) * sizeof(double)));
  CK(cudaMalloc(&d_rmask, std::max<size_t>(1, (size_t)g_nitems) * sizeof(int)));
  return 0;
}

// ---- launch sequences; replayed as CUDA graphs (the sequences are launch-latency bound) -----------------
struct GKey {
  int kind; const void *p0, *p1, *p2, *p3, *p4; int64_t a, b, c; int flag;
  bool operator<(const GKey& o) const { return std::memcmp(this, &o, sizeof(GKey)) < 0; }
};
std::map<GKey, cudaGraphExec_t> g_graphs;
void drop_graphs() { for (auto& kv : g_graphs) cudaGraphExecDestroy(kv.second); g_graphs.clear(); }

// yl_src: where phase 0 reads yl (dyl itself, or the caller's device-visible host array, then copied to dyl);
// yldot_host: optional device-visible host destination written by phase 3 next to dyldot
int enqueue_residual(const double* dyl, double* dyldot, bool need_rows, const double* yl_src = nullptr, double* yldot_host = nullptr) {
  const int B = 128, G = (NC + B - 1) / B, G32 = (NC + 31) / 32;
  if (yl_src && yl_src != dyl) CK(launch(k_phase0, dim3(G), dim3(B), d_base, yl_src, const_cast<double*>(dyl), neq, NXS, NC, d_err));
  else CK(launch(k_phase0, dim3(G), dim3(B), d_base, dyl, (double*)nullptr, neq, NXS, NC, d_err));
  CK(launch(k_phase1, dim3(G32), dim3(160), d_base, NXS, NC));
  if (need_rows && g_fuse23)
    CK(launch(k_phase23, dim3(std::max(G32, (g_nguard_cells + 31) / 32)), dim3(160), d_base, d_tmp, dyldot, dyl, (const double*)d_dtuse, (const double*)d_ylodt, neq, NXS, NC,
              (const int*)d_guard_cells, g_nguard_cells, d_err, d_hflags, yldot_host));
  else {
    CK(launch(k_phase2, dim3(std::max(G32, (g_nguard_cells + 31) / 32)), dim3(160), d_base, d_tmp, NXS, NC, (const int*)d_guard_cells, g_nguard_cells));
    if (need_rows) CK(launch(k_phase3, dim3(G), dim3(B), d_base, d_tmp, dyldot, dyl, d_dtuse, d_ylodt, neq, NXS, NC, d_err, d_hflags, yldot_host));
  }
  return 0;
}
// MPICollectBroadCastJacobian (ppp/mpi_parallel.F90:262-364) on the device: the reference gathers the per-rank CSC fragments
// on rank 0 and broadcasts the concatenation; here every rank broadcasts its own fragment range in place (the fragment
// layout coloff is the same on all ranks), the per-row counts are all-reduced, and every rank transposes the full CSC.
#define NCK(call)                                                                                              \
  do {                                                                                                         \
    ncclResult_t r_ = (call);                                                                                  \
    if (r_ != ncclSuccess) { g_err = std::string("NCCL error: ") + NC_.GetErrorString(r_) + " at " #call; return -11; } \
  } while (0)
int comm_gather_columns() {
  g_comm_bytes = 0;
  NCK(NC_.GroupStart());
  for (int r = 0; r < g_nranks; ++r) {
    const int64_t lo = g_rank_lo[r], hi = g_rank_hi[r];
    if (hi < lo) continue;
    const int64_t off = h_coloff[lo - 1], end = hi < neq ? h_coloff[hi] : g_cap_total, cnt = end - off;
    NCK(NC_.Broadcast(d_colval + off, d_colval + off, (size_t)cnt, ncclFloat64, r, g_comm, g_stream));
    NCK(NC_.Broadcast(d_colrow + off, d_colrow + off, (size_t)cnt, ncclInt32, r, g_comm, g_stream));
    NCK(NC_.Broadcast(d_colcnt + (lo - 1), d_colcnt + (lo - 1), (size_t)(hi - lo + 1), ncclInt32, r, g_comm, g_stream));
    const int64_t b = cnt * 12 + (hi - lo + 1) * 4;
    g_comm_bytes += (r == g_rank) ? b * (g_nranks - 1) : b;
  }
  NCK(NC_.AllReduce(d_rowcnt, d_rowcnt, (size_t)neq, ncclInt32, ncclSum, g_comm, g_stream));
  g_comm_bytes += 2 * neq * 4;
  NCK(NC_.GroupEnd());
  return 0;
}
int res_launches() { return g_fuse23 ? 3 : 4; }  // kernels of one residual sequence with rows
int jac_launches() { return (int)h_list.size() >= 4096 ? 8 : 6; }  // kernels of one Jacobian sequence (large / small grids)
int enqueue_jac(const double* dyl, const double* dy00, int64_t ml, int64_t mu, int64_t nnzmx, double* djac, int64_t* dja, int64_t* dia, bool base_current,
                double* jac_host = nullptr, int64_t* ja_host = nullptr, int64_t* ia_host = nullptr) {
  if (!base_current) { int rc = enqueue_residual(dyl, nullptr, false); if (rc) return rc; }
  const int NU = (int)h_list.size();
  if (NU == 0) CK(cudaMemsetAsync(d_colcnt, 0, 3 * neq * sizeof(int), g_stream));  // otherwise k_jb_stage0 clears the counters
  if (NU > 0) {
    JArgs A;
    A.ui = (const UInfo*)d_uinfo; A.cand_cell = d_cand_cell; A.cand_east = d_cand_east; A.item_u = d_item_u; A.guard_items = d_guard_items; A.nguard = g_nguard;
    A.NU = NU; A.nitems = g_nitems;
    A.priv = d_priv; A.rows = d_jrows; A.rres = d_rres; A.rmask = d_rmask; A.base = d_base;
    A.yl = dyl; A.yldot00 = dy00; A.suscal = d_suscal; A.sfscal = d_sfscal; A.dtuse = d_dtuse; A.ylodt = d_ylodt;
    A.neq = neq; A.ml = ml; A.mu = mu; A.NXS = NXS; A.NC = NC;
    A.coloff = d_coloff; A.colcnt = d_colcnt; A.colrow = d_colrow; A.colval = d_colval; A.rowcnt = d_rowcnt; A.err = d_err + 1;
    const unsigned gs = (unsigned)((NU * 4 + 127) / 128), gi = (unsigned)((g_nitems + 127) / 128);
    const bool big = NU >= 4096;  // more than ~2 waves of blocks per role: throughput-bound
    A.role0 = 0;
    if (getenv("UE_DEBUG_SPLIT_ROLES")) {  // developer aid: one launch per role so that a launch list shows each role's duration
      CK(launch(k_jb_stage0, dim3((unsigned)((NU + 31) / 32)), dim3(128), A));
      for (int r = 0; r < 3; ++r) { A.role0 = r; CK(launch(k_jb_p1a, dim3(dim3(gs, 1)), dim3(128), A)); }
      for (int r = 0; r < 5; ++r) { A.role0 = r; CK(launch(k_jb_p1b<1>, dim3(dim3(gs, 1)), dim3(128), A)); }
      for (int r = 0; r < 5; ++r) { A.role0 = r; CK(launch(k_jb_p2<1>, dim3(dim3(gi, 1)), dim3(128), A)); }
      A.role0 = 0;
    } else if (big) {
      CK(launch(k_jb_stage0, dim3((unsigned)((NU + 31) / 32)), dim3(128), A));
      CK(launch(k_jb_p1a, dim3(dim3(gs, 3)), dim3(128), A));
      CK(launch(k_jb_p1b<6>, dim3(dim3(gs, 5)), dim3(128), A));
      CK(launch(k_jb_p2<6>, dim3(dim3(gi, 5)), dim3(128), A));
    } else {
      CK(launch(k_jb_p01, dim3((unsigned)((NU * 4 + P01_ITEMS - 1) / P01_ITEMS)), dim3(5 * P01_ITEMS), A));
      CK(launch(k_jb_p2<1>, dim3(dim3(gi, 5)), dim3(128), A));
    }
    CK(launch(k_jb_p3c, dim3((unsigned)((NU + 3) / 4)), dim3(128), A));
  }
  int64_t f_lo = g_ivmin, f_hi = g_ivmax;
  if (g_nranks > 1) {  // every rank receives every other rank's column fragments; the CSR is then built from all columns
    int rc = comm_gather_columns();
    if (rc) return rc;
    f_lo = 1; f_hi = neq;
  }
  CK(launch(k_scan, dim3(1), dim3(1024), d_rowcnt, dia, neq, d_err, d_hflags, ia_host));
  const int64_t ncol = f_hi - f_lo + 1;
  if (ncol > 0) {
    CK(launch(k_fill, dim3((unsigned)ncol), dim3(64), neq, f_lo, f_hi, d_coloff, d_colcnt, d_colrow, d_colval, dia, d_rowfill, djac, dja, nnzmx));
    CK(launch(k_sortrows, dim3((unsigned)((neq + 3) / 4)), dim3(128), neq, dia, djac, dja, nnzmx, jac_host, ja_host, d_colrow, d_colval, g_cap_total));
  }
  return 0;
}
template <typename F>
int replay(const GKey& key, F enqueue) {
  if (g_nranks > 1) return enqueue();  // the sequence contains NCCL calls: launched directly, not captured
  auto it = g_graphs.find(key);
  if (it == g_graphs.end()) {
    cudaGraphExec_t ex = nullptr;
    for (int attempt = 0; attempt < 1 && !ex; ++attempt) {
      cudaGraph_t graph = nullptr;
      CK(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal));
      int rc = enqueue();
      cudaError_t e = cudaStreamEndCapture(g_stream, &graph);
      if (e == cudaSuccess && rc == 0) e = cudaGraphInstantiate(&ex, graph, 0);
      if (graph) cudaGraphDestroy(graph);
      if (e == cudaSuccess && rc == 0) break;
      ex = nullptr;
      cudaGetLastError();
      if (rc) return rc;
      g_err = std::string("CUDA graph capture failed: ") + cudaGetErrorString(e);
      return -10;
    }
    if (g_graphs.size() > 64) drop_graphs();
    it = g_graphs.emplace(key, ex).first;
  }
  CK(cudaGraphLaunch(it->second, g_stream));
  return 0;
}

int run_residual_dev(const double* dyl, double* dyldot, bool need_rows) {
  GKey k; std::memset(&k, 0, sizeof k);
  k.kind = 1; k.p0 = dyl; k.p1 = dyldot; k.flag = need_rows;
  g_launches += need_rows ? res_launches() : 3;
  return replay(k, [&]() { return enqueue_residual(dyl, dyldot, need_rows); });
}

int err_of_flags() {  // after a synchronisation: error bits the last sequence posted to mapped host memory
  const long long h = h_flags[0] | h_flags[2];
  h_flags[0] = h_flags[2] = 0;
  if (h & 1) { g_err = "***  ni is negative - calculation stopped"; return -3; }
  if (h & 2) { g_err = "***  ng is negative - calculation stopped"; return -3; }
  if (h & 4) { g_err = "jac_calc: yldot00 is not pandf1(yl) as evaluated by this library (call order rhsnk -> jac_calc, oderhs.m:9466-9468)"; return -4; }
  return 0;
}
// Device-visible alias of a caller's host array if it is page-locked (cudaHostAlloc / cudaHostRegister): kernels can then
// read or write it directly and the copy nodes disappear from the sequence.  Pageable memory returns nullptr.
template <typename T>
T* device_alias(const T* host) {
  cudaPointerAttributes at;
  if (cudaPointerGetAttributes(&at, (const void*)host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (at.type != cudaMemoryTypeHost || !at.devicePointer) return nullptr;
  return (T*)at.devicePointer;
}
int check_errflag() {
  CK(cudaStreamSynchronize(g_stream));
  return err_of_flags();
}

int run_jac_dev(const double* dyl, const double* dy00, int64_t ml, int64_t mu, int64_t nnzmx, double* djac, int64_t* dja, int64_t* dia, int64_t* nnz_out,
                bool base_current) {
  GKey k; std::memset(&k, 0, sizeof k);
  k.kind = 2; k.p0 = dyl; k.p1 = dy00; k.p2 = djac; k.p3 = dja; k.p4 = dia; k.a = ml; k.b = mu; k.c = nnzmx; k.flag = base_current;
  g_launches += (base_current ? 0 : 3) + jac_launches();
  int rc = replay(k, [&]() { return enqueue_jac(dyl, dy00, ml, mu, nnzmx, djac, dja, dia, base_current); });
  if (rc) return rc;
  CK(cudaStreamSynchronize(g_stream));
  const int64_t nnz = (int64_t)h_flags[1] - 1;  // k_scan posts ia(neq+1) to mapped host memory
  *nnz_out = nnz;
  if (nnz > nnzmx) {
    g_err = "*** jac_calc -- More storage needed for Jacobian. Storage exceeded. Increase lenpfac.";
    return -2;
  }
  return 0;
}

}  // namespace

// ====================================================================================================
extern "C" {

int ue_gpu_set_int(const char* n, int64_t v) {
  // Before ue_gpu_init: stored.  After it: the device lists, derived flags and refusals were built from the switches, so a
  // changed switch needs a new ue_gpu_init -- except model_dt (read by ue_gpu_set_dt only), which is patched in place.
  if (g_ready) {
    if (S.zero_only.count(n)) {  // as ue_gpu_set_real: checked at init, a non-zero value afterwards is refused at once
      if (v != 0) { g_err = std::string("input ") + n + " must be 0: the term it switches on is outside the built hot path"; return -5; }
      return 0;
    }
    auto it = S.iscal.find(n);
    if (it != S.iscal.end() && *it->second == v) return 0;  // unchanged
    if (it != S.iscal.end() && std::string(n) == "model_dt") {
      *it->second = v;
      const size_t off = (size_t)((char*)it->second - (char*)&S.p);
      CK(cudaMemcpyToSymbol(D, &v, sizeof(int64_t), off));
      return 0;
    }
    g_ready = false;  // entry points now fail with "ue_gpu_init not called" until the caller re-initialises
  }
  g_base_valid = g_base_dev_valid = false;
  if (S.set_int(n, v)) { g_err = std::string("unknown int input ") + n; return -1; }
  return 0;
}
int ue_gpu_set_real(const char* n, double v) {
  if (S.zero_only.count(n)) {  // checked at ue_gpu_init; after it, a non-zero value is refused at once
    S.set_real(n, v);
    if (g_ready && v != 0.) { g_err = std::string("input ") + n + " must be 0: the term it switches on is outside the built hot path"; return -5; }
    return 0;
  }
  if (g_ready) {  // the shim re-sends nufak before every Jacobian: an unchanged value keeps the cached base fields
    auto it = S.rscal.find(n);
    if (it != S.rscal.end() && std::memcmp(it->second, &v, 8) == 0) return 0;
  }
  g_base_valid = g_base_dev_valid = false;
  if (S.set_real(n, v)) { g_err = std::string("unknown real input ") + n; return -1; }
  if (g_ready) {  // scalars such as nufak, dtreal may change between solves: patch the device copy in place
    const size_t off = (size_t)((char*)S.rscal[n] - (char*)&S.p);
    CK(cudaMemcpyToSymbol(D, &v, sizeof(double), off));
  }
  return 0;
}
// arrays are uploaded by ue_gpu_init: sending one afterwards disables the entry points until the next ue_gpu_init
int ue_gpu_set_real_array(const char* n, const double* d, int64_t k) { g_ready = false; if (S.set_real_array(n, d, k)) { g_err = std::string("unknown real array ") + n; return -1; } return 0; }
int ue_gpu_set_int_array(const char* n, const int64_t* d, int64_t k) { g_ready = false; if (S.set_int_array(n, d, k)) { g_err = std::string("unknown int array ") + n; return -1; } return 0; }
const char* ue_gpu_last_error(void) { return g_err.c_str(); }

int ue_gpu_init(void) {
  g_fuse23 = getenv("UE_GPU_NO_FUSE23") == nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "no CUDA device: the B200 path has no CPU fallback"; return -9; }
  if (g_alloc) free_all();  // (g_ready may already be false: a switch was changed after the previous init)
  std::string m = S.missing();
  if (!m.empty()) { g_err = "missing inputs: " + m; return -1; }
  const UeParams& P = S.p;
  nx = (int)P.nx; ny = (int)P.ny; NXS = nx + 2; NC = NXS * (ny + 2); neq = P.neq;
  std::string b = S.bad_sizes();
  if (!b.empty()) { g_err = "bad plane sizes: " + b; return -1; }
  if (neq != (int64_t)NC * S.p.numvar) { g_err = "neq != numvar*(nx+2)*(ny+2)"; return -1; }
  if (S.len("igyl") != 2 * neq || S.len("iseqalg") != neq) { g_err = "igyl/iseqalg length"; return -1; }
  int rc = check_switches();
  if (rc) return rc;
  if (!g_stream) CK(cudaStreamCreate(&g_stream));
  if (!g_ev0) { CK(cudaEventCreate(&g_ev0)); CK(cudaEventCreate(&g_ev1)); }
  // ---- static data to the device -------------------------------------------------------------------
  UeParams dp = S.p;
#define X(n) if ((rc = dev_copy(S.rdata[#n].data(), S.rdata[#n].size(), &dp.n))) return rc;
  UE_REAL_PLANES(X)
  UE_REAL_LINES(X)
#undef X
#define X(n) if ((rc = dev_copy(S.idata[#n].data(), S.idata[#n].size(), &dp.n))) return rc;
  UE_INT_PLANES(X)
  UE_INT_LINES(X)
#undef X
  CK(cudaMemcpyToSymbol(D, &dp, sizeof(UeParams)));
  DevTables t;
  std::memset(&t, 0, sizeof t);
  t.mpe = (int)P.mpe; t.mpd = (int)P.mpd;
  t.iscut = (P.isfixlb == 2 && P.iysptrx1 > 0) ? 1 : 0;
  {
    bool rare = P.isupcore >= 2 || P.iflcore == -1 || P.isngcore != 0;
    for (int ix = 0; ix < (int)P.nx + 2; ++ix)
      rare = rare || P.isnwconiix[ix] != 0 || P.isnwconoix[ix] != 0 || P.istepfcix[ix] >= 2 || P.istipfcix[ix] >= 2 || P.istewcix[ix] >= 2 || P.istiwcix[ix] >= 2 ||
             P.matwalli[ix] > 0 || P.matwallo[ix] > 0;
    t.rarebc = rare ? 1 : 0;
  }
  if (P.istabon == 10) {
    if (t.mpe < 2 || t.mpe > 64 || t.mpd < 2 || t.mpd > 16 || S.len("wsveh") != (int64_t)t.mpe * t.mpd) { g_err = "istabon=10 needs wsveh/wsveh0/welms1/welms2 (mpe<=64, mpd<=16)"; return -1; }
    t.dkpt[0] = 16.0; for (int j = 1; j < t.mpd; ++j) t.dkpt[j] = t.dkpt[j - 1] + 0.5;
    t.rldmin = t.dkpt[0]; t.rldmax = t.dkpt[t.mpd - 1]; t.deldkpt = (t.rldmax - t.rldmin) / double(t.mpd - 1);
    t.ekpt[0] = -1.2 * std::log(10.0); for (int j = 1; j < t.mpe; ++j) t.ekpt[j] = t.ekpt[j - 1] + 0.1 * std::log(10.0);
    t.rlemin = t.ekpt[0]; t.rlemax = t.ekpt[t.mpe - 1]; t.delekpt = (t.rlemax - t.rlemin) / double(t.mpe - 1);
  }
  CK(cudaMemcpyToSymbol(DT, &t, sizeof(DevTables)));
  // ---- work space -------------------------------------------------------------------------------------
  CK(cudaMalloc(&d_base, (size_t)PL_COUNT * NC * sizeof(double)));
  CK(cudaMemset(d_base, 0, (size_t)PL_COUNT * NC * sizeof(double)));
  CK(cudaMalloc(&d_yl, (neq + 2) * sizeof(double)));
  CK(cudaMalloc(&d_yldot00, (neq + 2) * sizeof(double)));
  CK(cudaMalloc(&d_tmp, (size_t)NC * UE_NV * sizeof(double)));  // UE_NV row slots per cell (numvar may be 4)
  CK(cudaMalloc(&d_yldot, neq * sizeof(double)));
  CK(cudaMalloc(&d_dtuse, neq * sizeof(double)));
  CK(cudaMalloc(&d_ylodt, neq * sizeof(double)));
  CK(cudaMalloc(&d_dtoptv, neq * sizeof(double)));
  CK(cudaMemset(d_dtoptv, 0, neq * sizeof(double)));
  CK(cudaMalloc(&d_suscal, neq * sizeof(double)));
  CK(cudaMalloc(&d_sfscal, neq * sizeof(double)));
  CK(cudaMalloc(&d_err, 2 * sizeof(int)));  // [0] residual sequence, [1] Jacobian sequence
  CK(cudaMemset(d_err, 0, 2 * sizeof(int)));  // afterwards the kernel that posts the error bits clears them
  if (!h_flags) {
    CK(cudaHostAlloc((void**)&h_flags, 4 * sizeof(long long), cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer((void**)&d_hflags, (void*)h_flags, 0));
  }
  h_flags[0] = h_flags[1] = h_flags[2] = 0; g_nnz_guess = 0;
  {
    std::vector<double> big(neq, 1e20), one(neq, 1.0), zero(neq, 0.0);
    CK(cudaMemcpy(d_dtuse, big.data(), neq * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_ylodt, zero.data(), neq * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_suscal, one.data(), neq * 8, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sfscal, one.data(), neq * 8, cudaMemcpyHostToDevice));
  }
  g_ivmin = 1; g_ivmax = neq;
  build_lists();
  if ((rc = upload_lists())) return rc;
  {  // guard cells of the full residual, sorted by kind
    std::vector<int> cells;
    for (int c = 0; c < NC; ++c) { const int ix = c % NXS, iy = c / NXS; if (!(ix >= 1 && ix <= nx && iy >= 1 && iy <= ny)) cells.push_back(c); }
    const std::vector<int> gc = sort_pad_by_kind(cells, [](int c) { return c; });
    g_nguard_cells = (int)gc.size();
    CK(cudaMalloc(&d_guard_cells, std::max<size_t>(1, gc.size()) * sizeof(int)));
    CK(cudaMemcpy(d_guard_cells, gc.data(), gc.size() * sizeof(int), cudaMemcpyHostToDevice));
  }
  CK(cudaMalloc(&d_coloff, neq * sizeof(int64_t)));
  CK(cudaMemcpy(d_coloff, h_coloff.data(), neq * sizeof(int64_t), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_colcnt, 3 * neq * sizeof(int)));  // colcnt | rowcnt | rowfill contiguous: one memset per Jacobian
  CK(cudaMalloc(&d_colrow, g_cap_total * sizeof(int)));
  CK(cudaMalloc(&d_colval, g_cap_total * sizeof(double)));
  d_rowcnt = d_colcnt + neq;
  d_rowfill = d_colcnt + 2 * neq;
  g_nnzcap = g_cap_total;
  CK(cudaMalloc(&d_ia, (neq + 1) * sizeof(int64_t)));
  CK(cudaMalloc(&d_ja, g_nnzcap * sizeof(int64_t)));
  CK(cudaMalloc(&d_jac, g_nnzcap * sizeof(double)));
  g_launches = 0;
  g_ready = true;
  g_alloc = true;
  return 0;
}

int ue_gpu_step_params(int64_t n, const double* dt, const double* yo, const double* su, const double* sf) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "step_params: neq mismatch"; return -1; }
  // The interface routines call this before every residual/Jacobian; the vectors change only
  // between nonlinear solves, so unchanged ones are recognised on the host and not uploaded again.
  const double* src[4] = {dt, yo, su, sf};
  double* dst[4] = {d_dtuse, d_ylodt, d_suscal, d_sfscal};
  bool any = false;
  for (int i = 0; i < 4; i++) {
    std::vector<double>& h = g_step_host[i];
    if ((int64_t)h.size() == n && std::memcmp(h.data(), src[i], n * 8) == 0) continue;
    h.assign(src[i], src[i] + n);
    CK(cudaMemcpyAsync(dst[i], h.data(), n * 8, cudaMemcpyHostToDevice, g_stream));
    any = true;
    if (i < 2) g_last_yldot.clear();  // dtuse / ylodt enter the residual rows: the cached yldot is stale
  }
  if (any) CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int ue_gpu_pandf1_dev(int64_t n, double time, const double* dyl, double* dyldot) {
  (void)time;
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "pandf1: neq mismatch"; return -1; }
  g_base_valid = false; g_base_dev_valid = false;
  CK(cudaEventRecord(g_ev0, g_stream));
  int rc = run_residual_dev(dyl, dyldot, true);
  if (rc) return rc;
  CK(cudaEventRecord(g_ev1, g_stream));
  rc = check_errflag();
  if (rc) return rc;
  cudaEventElapsedTime(&g_res_ms, g_ev0, g_ev1);
  g_base_dev_valid = true;
  return 0;
}

int ue_gpu_pandf1(int64_t n, double time, const double* yl, double* yldot) {
  (void)time;
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "pandf1: neq mismatch"; return -1; }
  if (g_base_valid && (int64_t)g_base_yl.size() == neq + 2 && std::memcmp(g_base_yl.data(), yl, neq * 8) == 0) {
    // Same unknowns as the previous call, only yl(neq+1)/yl(neq+2) may differ (psetnk re-evaluates f0 with the Jacobian
    // flag off right after jac_calc, oderhs.m:9470-9471).  Fluxes and guard rows do not read the flags: only the row
    // scaling / time-step phase is redone.
    CK(cudaMemcpyAsync(d_yl + neq, yl + neq, 16, cudaMemcpyHostToDevice, g_stream));
    const int B = 128, G = (NC + B - 1) / B;
    CK(launch(k_phase3, dim3(G), dim3(B), d_base, d_tmp, d_yldot, d_yl, d_dtuse, d_ylodt, neq, NXS, NC, d_err, d_hflags, (double*)nullptr));
    g_launches += 1;
    CK(cudaMemcpyAsync(yldot, d_yldot, neq * 8, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
    g_base_yl[neq] = yl[neq]; g_base_yl[neq + 1] = yl[neq + 1];
    g_last_yldot.assign(yldot, yldot + neq);
    return 0;
  }
  g_base_valid = false; g_base_dev_valid = false;
  // H2D(yl) -> phases 0-3 -> D2H(yldot), D2H(err) as ONE graph launch and ONE synchronisation
  GKey k; std::memset(&k, 0, sizeof k);
  k.kind = 3; k.p0 = yl; k.p1 = yldot;
  g_launches += res_launches();
  int rc = 0;
  // Solvers call with the same work arrays every time (NKSOL's savf/u); a pointer pair seen for the first time takes
  // the plain path so that callers with fresh buffers per call do not pay a capture each time.
  bool use_graph = g_host_graphs && (g_graphs.count(k) || !g_seen_host.insert({yl, yldot}).second);
  if (g_seen_host.size() > 256) g_seen_host.clear();
  // page-locked caller arrays: phase 0 reads yl and phase 3 writes yldot directly (no copy nodes)
  const double* yl_dev = device_alias(yl);
  double* yldot_dev = device_alias(yldot);
  auto body = [&]() {
    if (!yl_dev) CK(cudaMemcpyAsync(d_yl, yl, (neq + 2) * 8, cudaMemcpyHostToDevice, g_stream));
    int r = enqueue_residual(d_yl, d_yldot, true, yl_dev, yldot_dev);
    if (r) return r;
    if (!yldot_dev) CK(cudaMemcpyAsync(yldot, d_yldot, neq * 8, cudaMemcpyDeviceToHost, g_stream));
    return 0;
  };
  if (use_graph) {
    k.flag = (yl_dev ? 1 : 0) | (yldot_dev ? 2 : 0);
    rc = replay(k, body);
    if (rc == -10) { g_host_graphs = false; use_graph = false; cudaGetLastError(); }  // not capturable: plain path from now on
  }
  if (!use_graph) rc = body();
  if (rc) return rc;
  CK(cudaStreamSynchronize(g_stream));  // the only synchronisation of the call
  if ((rc = err_of_flags())) return rc;
  g_base_yl.assign(yl, yl + neq + 2);  // the base planes (and d_yl) now describe this yl
  g_last_yldot.assign(yldot, yldot + neq);
  g_base_valid = true; g_base_dev_valid = true;
  return 0;
}

int ue_gpu_jac_calc_dev(int64_t n, double t, const double* dyl, const double* dy00, int64_t ml, int64_t mu, int64_t nnzmx, double* djac, int64_t* dja,
                        int64_t* dia, int64_t* nnz_out) {
  (void)t;
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "jac_calc: neq mismatch"; return -1; }
  CK(cudaEventRecord(g_ev0, g_stream));
  const bool cur = g_base_dev_valid && g_jac_trust_base;  // the host-pointer bookkeeping (g_base_valid, g_base_yl) is not involved
  g_jac_trust_base = false;
  int rc = run_jac_dev(dyl, dy00, ml, mu, std::min(nnzmx, g_nnzcap), djac, dja, dia, nnz_out, cur);
  g_base_valid = false;  // device-pointer callers may change d_yl behind our back
  if (rc) return rc;
  CK(cudaEventRecord(g_ev1, g_stream));
  rc = check_errflag();
  if (rc) return rc;
  cudaEventElapsedTime(&g_jac_ms, g_ev0, g_ev1);
  return 0;
}

int ue_gpu_rhs_jac_dev(int64_t n, const double* dyl, double* dyldot00, int64_t ml, int64_t mu, int64_t nnzmx, double* djac, int64_t* dja, int64_t* dia,
                       int64_t* nnz_out, double* ms) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "rhs_jac: neq mismatch"; return -1; }
  g_base_valid = false; g_base_dev_valid = false;
  const int64_t lim = std::min(nnzmx, g_nnzcap);
  CK(cudaEventRecord(g_ev0, g_stream));
  int rc = run_residual_dev(dyl, dyldot00, true);
  if (rc) return rc;
  GKey k; std::memset(&k, 0, sizeof k);
  k.kind = 2; k.p0 = dyl; k.p1 = dyldot00; k.p2 = djac; k.p3 = dja; k.p4 = dia; k.a = ml; k.b = mu; k.c = lim; k.flag = 1;
  g_launches += jac_launches();
  rc = replay(k, [&]() { return enqueue_jac(dyl, dyldot00, ml, mu, lim, djac, dja, dia, true); });
  if (rc) return rc;
  CK(cudaEventRecord(g_ev1, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  if ((rc = err_of_flags())) return rc;
  g_base_dev_valid = true;
  float f = 0.f; cudaEventElapsedTime(&f, g_ev0, g_ev1); *ms = f;
  const int64_t nnz = (int64_t)h_flags[1] - 1;
  *nnz_out = nnz;
  if (nnz > nnzmx) { g_err = "*** jac_calc -- More storage needed for Jacobian. Storage exceeded. Increase lenpfac."; return -2; }
  return 0;
}

// Tail of the host-pointer Jacobian calls: fetch ia / jac / ja unless the kernels wrote them to the caller's arrays
// directly, one synchronisation in the usual case, error bits and nnz from mapped host memory.
static int finish_host_jac(bool direct, const double* yl, bool same_y, int64_t nnzmx, int64_t lim, double* jac, int64_t* ja, int64_t* ia, int64_t* nnz_out) {
  // copies: ia always; jac/ja speculatively with the previous call's nnz (the pattern rarely changes between Newton
  // steps); a larger nnz fetches the remainder afterwards
  const int64_t guess = direct ? 0 : std::min(g_nnz_guess, std::min(nnzmx, lim));
  if (!direct) {
    CK(cudaMemcpyAsync(ia, d_ia, (neq + 1) * 8, cudaMemcpyDeviceToHost, g_stream));
    if (guess > 0) {
      CK(cudaMemcpyAsync(jac, d_jac, guess * 8, cudaMemcpyDeviceToHost, g_stream));
      CK(cudaMemcpyAsync(ja, d_ja, guess * 8, cudaMemcpyDeviceToHost, g_stream));
    }
  }
  CK(cudaStreamSynchronize(g_stream));
  int rc = err_of_flags();
  if (rc) return rc;
  if (!same_y) { g_base_yl.assign(yl, yl + neq + 2); g_base_valid = true; g_base_dev_valid = true; }  // base fields describe this yl now
  const int64_t nnz = (int64_t)h_flags[1] - 1;
  *nnz_out = nnz;
  if (nnz > nnzmx) { g_err = "*** jac_calc -- More storage needed for Jacobian. Storage exceeded. Increase lenpfac."; return -2; }
  g_nnz_guess = nnz;
  if (!direct && nnz > guess) {
    CK(cudaMemcpyAsync(jac + guess, d_jac + guess, (nnz - guess) * 8, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaMemcpyAsync(ja + guess, d_ja + guess, (nnz - guess) * 8, cudaMemcpyDeviceToHost, g_stream));
    CK(cudaStreamSynchronize(g_stream));
  }
  return 0;
}

// Host-pointer form of the pair: upload yl once, residual + Jacobian as one stream sequence, one synchronisation.
int ue_gpu_rhs_jac(int64_t n, const double* yl, double* yldot00, int64_t ml, int64_t mu, int64_t nnzmx, double* jac, int64_t* ja, int64_t* ia,
                   int64_t* nnz_out) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "rhs_jac: neq mismatch"; return -1; }
  g_base_valid = false; g_base_dev_valid = false;
  const int64_t lim = std::min(nnzmx, g_nnzcap);
  const double* yl_dev = device_alias(yl);
  double* f_dev = device_alias(yldot00);
  double* jac_dev = device_alias(jac); int64_t* ja_dev = device_alias(ja); int64_t* ia_dev = device_alias(ia);
  const bool direct = jac_dev && ja_dev && ia_dev;
  GKey k; std::memset(&k, 0, sizeof k);
  k.kind = 5; k.p0 = yl; k.p1 = yldot00; k.p2 = jac; k.p3 = ja; k.p4 = ia; k.a = ml; k.b = mu; k.c = lim;
  k.flag = (yl_dev ? 1 : 0) | (f_dev ? 2 : 0) | (direct ? 4 : 0);
  g_launches += res_launches() + jac_launches();
  auto body = [&]() {
    if (!yl_dev) CK(cudaMemcpyAsync(d_yl, yl, (neq + 2) * 8, cudaMemcpyHostToDevice, g_stream));
    int r = enqueue_residual(d_yl, d_yldot, true, yl_dev, f_dev);
    if (r) return r;
    if (!f_dev) CK(cudaMemcpyAsync(yldot00, d_yldot, neq * 8, cudaMemcpyDeviceToHost, g_stream));
    return direct ? enqueue_jac(d_yl, d_yldot, ml, mu, lim, d_jac, d_ja, d_ia, true, jac_dev, ja_dev, ia_dev)
                  : enqueue_jac(d_yl, d_yldot, ml, mu, lim, d_jac, d_ja, d_ia, true);
  };
  // same policy as ue_gpu_pandf1: a pointer set seen for the first time runs un-captured
  const bool use_graph = g_host_graphs && (g_graphs.count(k) || !g_seen_host.insert({yl, jac}).second);
  int rc = use_graph ? replay(k, body) : body();
  if (rc) return rc;
  rc = finish_host_jac(direct, yl, false, nnzmx, lim, jac, ja, ia, nnz_out);
  if (rc) return rc;
  g_last_yldot.assign(yldot00, yldot00 + neq);
  return 0;
}

int ue_gpu_jac_calc(int64_t n, double t, const double* yl, const double* yldot00, int64_t ml, int64_t mu, int64_t nnzmx, double* jac, int64_t* ja,
                    int64_t* ia, int64_t* nnz_out) {
  (void)t;
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "jac_calc: neq mismatch"; return -1; }
  // psetnk / sfsetnk evaluate rhsnk(yl) immediately before jac_calc (oderhs.m:9466-9468, 9851-9856): if yl is
  // bit-identical to that call's, the base fields (and yl itself) are already on the device and phases 0-2 and the
  // upload are skipped; likewise yldot00 if it is the vector that call returned.
  const bool same_y = g_base_valid && (int64_t)g_base_yl.size() == neq + 2 && std::memcmp(g_base_yl.data(), yl, neq * 8) == 0;
  const bool same_flags = same_y && std::memcmp(g_base_yl.data() + neq, yl + neq, 16) == 0;
  const bool same_f = same_y && (int64_t)g_last_yldot.size() == neq && std::memcmp(g_last_yldot.data(), yldot00, neq * 8) == 0;
  if (!same_y) { g_base_valid = false; g_base_dev_valid = false; }
  if (!same_flags) CK(cudaMemcpyAsync(d_yl, yl, (neq + 2) * 8, cudaMemcpyHostToDevice, g_stream));
  const double* dy00 = d_yldot;
  bool base_current = same_y;
  if (!same_f) {
    // yldot00 did not come from the immediately preceding ue_gpu_pandf1(yl).  The dependency-pruned windows are only
    // equivalent to the reference's full windows if yldot00 IS pandf1(yl) bit for bit, so evaluate it and compare;
    // anything else is refused rather than answered with a different sparsity pattern.
    CK(cudaMemcpyAsync(d_yldot00, yldot00, neq * 8, cudaMemcpyHostToDevice, g_stream));
    if (!same_y) {
      int r = run_residual_dev(d_yl, d_yldot, true);
      if (r) return r;
    } else {
      const int B = 128, G = (NC + B - 1) / B;
      CK(launch(k_phase3, dim3(G), dim3(B), d_base, d_tmp, d_yldot, d_yl, d_dtuse, d_ylodt, neq, NXS, NC, d_err, d_hflags, (double*)nullptr));
      g_launches += 1;
    }
    CK(launch(k_samebits, dim3((unsigned)((neq + 255) / 256)), dim3(256), d_yldot, d_yldot00, neq, d_err + 1));
    g_launches += 1;
    g_last_yldot.clear();
    base_current = true;
  }
  GKey k; std::memset(&k, 0, sizeof k);
  const int64_t lim = std::min(nnzmx, g_nnzcap);
  // page-locked caller arrays: k_scan / k_sortrows write ia, jac, ja there directly and no copy follows
  double* jac_dev = device_alias(jac); int64_t* ja_dev = device_alias(ja); int64_t* ia_dev = device_alias(ia);
  const bool direct = jac_dev && ja_dev && ia_dev;
  k.kind = direct ? 4 : 2; k.p0 = d_yl; k.p1 = dy00; k.p2 = direct ? (void*)jac_dev : (void*)d_jac; k.p3 = direct ? (void*)ja_dev : (void*)d_ja;
  k.p4 = direct ? (void*)ia_dev : (void*)d_ia; k.a = ml; k.b = mu; k.c = lim; k.flag = base_current;
  g_launches += (base_current ? 0 : 3) + jac_launches();
  int rc = replay(k, [&]() {
    return direct ? enqueue_jac(d_yl, dy00, ml, mu, lim, d_jac, d_ja, d_ia, base_current, jac_dev, ja_dev, ia_dev)
                  : enqueue_jac(d_yl, dy00, ml, mu, lim, d_jac, d_ja, d_ia, base_current);
  });
  if (rc) return rc;
  return finish_host_jac(direct, yl, same_y, nnzmx, lim, jac, ja, ia, nnz_out);
}

// set_dt(neq, yl, f0) of the nksol driver (bbb/odesolve.m:299, bbb/oderhs.m:9886-10147): f0 = rhsnk(yl), then the
// per-unknown time step dtuse from ylodt (last ue_gpu_step_params), deldt, dtreal and model_dt (0..3).  dtuse stays on
// the device for the residual and Jacobian calls that follow and is returned to the caller's array as well.
int ue_gpu_set_dt(int64_t n, const double* yl, double* f0, double* dtuse) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "set_dt: neq mismatch"; return -1; }
  if (S.p.model_dt < 0 || S.p.model_dt > 3) { g_err = "model_dt must be 0..3"; return -5; }
  int rc = ue_gpu_pandf1(n, 0., yl, f0);  // leaves the residual in d_yldot
  if (rc) return rc;
  static double* h_dt = nullptr; static int64_t h_cap = 0;  // pinned staging for the returned vector
  if (h_cap < neq) { if (h_dt) cudaFreeHost(h_dt); CK(cudaHostAlloc((void**)&h_dt, neq * 8, cudaHostAllocMapped)); h_cap = neq; }
  double* h_dt_dev = nullptr;
  CK(cudaHostGetDevicePointer((void**)&h_dt_dev, h_dt, 0));
  const int B = 128, G = (NC + B - 1) / B;
  CK(launch(k_set_dt, dim3(G), dim3(B), (const double*)d_yldot, (const double*)d_ylodt, d_dtoptv, d_dtuse, h_dt_dev, NXS, NC));
  g_launches += 1;
  CK(cudaStreamSynchronize(g_stream));
  std::memcpy(dtuse, h_dt, neq * 8);
  g_step_host[0].assign(dtuse, dtuse + neq);  // what the device now holds: an identical vector in step_params is not re-sent
  g_last_yldot.clear();                       // dtuse enters the residual rows
  return 0;
}

int ue_gpu_set_column_range(int64_t ivmin, int64_t ivmax) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (ivmin < 1 || ivmax > neq) { g_err = "column range outside 1..neq"; return -1; }
  if (g_nranks > 1) { g_err = "set_column_range: the ranges are owned by ue_gpu_comm_init while a communicator is active"; return -1; }
  g_ivmin = ivmin; g_ivmax = ivmax; g_nnz_guess = 0;
  build_lists();
  drop_graphs();
  return upload_lists();
}

// Device-pointer callers: assert that d_yl has not changed since the last ue_gpu_pandf1_dev call, so the next
// ue_gpu_jac_calc_dev may reuse the base planes (the host-pointer entry points check this themselves).
int ue_gpu_assume_base_current(int64_t flag) { g_jac_trust_base = (flag != 0); return 0; }
// sfsetnk (bbb/oderhs.m:9815-9884) with the Jacobian kept on the device: f0 = pandf1(yl | flag=1), J = jac_calc,
// J <- J*diag(1/su), sf(i) = 1/max_k|J_ik|, ydt_max0 = max_i|f0_i sf_i|.  Only sf (neq doubles) returns to the host.
int ue_gpu_sfsetnk(int64_t n, const double* yl, const double* su, int64_t ml, int64_t mu, double* sf, double* ydt_max0) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "sfsetnk: neq mismatch"; return -1; }
  std::vector<double> y(yl, yl + neq + 2);
  y[neq] = 1.;  // oderhs.m:9848
  CK(cudaMemcpyAsync(d_yl, y.data(), (neq + 2) * 8, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(d_suscal, su, neq * 8, cudaMemcpyHostToDevice, g_stream));
  g_base_valid = false; g_base_dev_valid = false;
  int rc = run_residual_dev(d_yl, d_yldot00, true);
  if (rc) return rc;
  int64_t nnz = 0;
  rc = run_jac_dev(d_yl, d_yldot00, ml, mu, g_nnzcap, d_jac, d_ja, d_ia, &nnz, true);
  if (rc) return rc;
  static unsigned long long* d_bits = nullptr; static int* d_zero = nullptr;
  if (!d_bits) { CK(cudaMalloc(&d_bits, 8)); CK(cudaMalloc(&d_zero, 4)); }
  const unsigned long long cut = (unsigned long long)0;  // ydt_max0 starts at cutlo (oderhs.m:9871); applied on the host
  const int big = 0x7fffffff;
  CK(cudaMemcpyAsync(d_bits, &cut, 8, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(d_zero, &big, 4, cudaMemcpyHostToDevice, g_stream));
  CK(launch(k_rownorm, dim3((unsigned)((neq + 3) / 4)), dim3(128), neq, d_ia, d_ja, d_jac, d_suscal, d_yldot00, d_tmp, d_bits, d_zero));
  g_launches += 1;
  unsigned long long bits = 0; int zero = 0;
  CK(cudaMemcpyAsync(sf, d_tmp, neq * 8, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(&bits, d_bits, 8, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(&zero, d_zero, 4, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  rc = check_errflag();
  if (rc) return rc;
  if (zero != big) { char b[96]; snprintf(b, sizeof b, "*** Error: Jacobian row = 0 for eqn iv = %d", zero); g_err = b; return -7; }
  double v; std::memcpy(&v, &bits, 8);
  *ydt_max0 = std::max(v, S.p.cutlo);
  return 0;
}
int ue_gpu_jac_scale(int64_t n, const double* su, const double* sf, int64_t isrnorm, int64_t normtype, int64_t nnz, double* jac, double* fnormnw) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "jac_scale: neq mismatch"; return -1; }
  if (nnz != (int64_t)h_flags[1] - 1) { g_err = "jac_scale: nnz is not that of the last ue_gpu_jac_calc"; return -1; }
  if (normtype < 0 || normtype > 2) { g_err = "jac_scale: normtype must be 0, 1 or 2"; return -1; }
  static double* d_sc = nullptr; static int64_t cap = 0;  // su | sf | factors
  if (cap < 3 * neq) { if (d_sc) cudaFree(d_sc); CK(cudaMalloc(&d_sc, 3 * neq * 8)); cap = 3 * neq; }
  CK(cudaMemcpyAsync(d_sc, su, neq * 8, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(d_sc + neq, sf, neq * 8, cudaMemcpyHostToDevice, g_stream));
  CK(launch(k_rowscale, dim3((unsigned)((neq + 127) / 128)), dim3(128), neq, (const int64_t*)d_ia, (const int64_t*)d_ja, d_jac, (const double*)d_sc, (const double*)(d_sc + neq), (int)isrnorm, (int)normtype, d_sc + 2 * neq));
  g_launches += 1;
  CK(cudaMemcpyAsync(jac, d_jac, nnz * 8, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(fnormnw, d_sc + 2 * neq, neq * 8, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}
// Page-lock a caller array (the solver's work arrays, the Jacobian storage) so that the kernels can read / write it
// directly.  Optional: without it the entry points use staged copies.
int ue_gpu_pin_host_array(void* p, int64_t bytes) {
  cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
  if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return 0; }
  if (e != cudaSuccess) { g_err = std::string("cudaHostRegister: ") + cudaGetErrorString(e); cudaGetLastError(); return -10; }
  return 0;
}
int ue_gpu_unpin_host_array(void* p) {
  drop_graphs();  // captured sequences may hold the device alias of this array
  g_seen_host.clear();
  cudaError_t e = cudaHostUnregister(p);
  if (e != cudaSuccess) { g_err = std::string("cudaHostUnregister: ") + cudaGetErrorString(e); cudaGetLastError(); return -10; }
  return 0;
}
// Arrays whose terms the built hot path does not evaluate (volume sources volpsor/volmsor/pwrsore/pwrsori, user
// profiles *_use, ...): the shim passes them here once after ueinit; any non-zero element is refused by name.
int ue_gpu_assert_zero(const char* name, const double* a, int64_t n) {
  for (int64_t i = 0; i < n; ++i)
    if (a[i] != 0.) { g_err = std::string("array ") + name + " must be identically 0: the term it feeds is outside the built hot path"; return -5; }
  return 0;
}
int ue_gpu_math_probe(int64_t op, int64_t n, const double* x, const double* y, double* out) {
  if (op < 0 || op > 5 || n <= 0) { g_err = "math_probe: bad arguments"; return -1; }
  double* d = nullptr;
  CK(cudaMalloc(&d, 3 * n * 8));
  CK(cudaMemcpy(d, x, n * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d + n, y, n * 8, cudaMemcpyHostToDevice));
  k_math_probe<<<(unsigned)((n + 255) / 256), 256>>>((int)op, n, d, d + n, d + 2 * n);
  cudaError_t e = cudaMemcpy(out, d + 2 * n, n * 8, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (e != cudaSuccess) { g_err = std::string("math_probe: ") + cudaGetErrorString(e); return -10; }
  return 0;
}
int ue_gpu_kernel_launches(int64_t* n) { *n = g_launches; return 0; }
int ue_gpu_last_kernel_ms(double* jac_ms, double* res_ms) { *jac_ms = g_jac_ms; *res_ms = g_res_ms; return 0; }
// device buffers owned by the library (for callers that keep state resident, e.g. bench.py)
int ue_gpu_device_buffers(double** yl, double** yldot, double** yldot00, double** jac, int64_t** ja, int64_t** ia) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  *yl = d_yl; *yldot = d_yldot; *yldot00 = d_yldot00; *jac = d_jac; *ja = d_ja; *ia = d_ia;
  return 0;
}
// copy one intermediate base plane to the host (parity debugging; plane ids in ue_device.cuh)
int ue_gpu_get_plane(int64_t pl, double* out) {
  if (!g_ready || pl < 0 || pl >= PL_COUNT) { g_err = "bad plane"; return -1; }
  CK(cudaMemcpy(out, d_base + (size_t)pl * NC, NC * 8, cudaMemcpyDeviceToHost));
  return 0;
}
// ---- multi-GPU ---------------------------------------------------------------------------------------------------------
static int nccl_bind() {
  if (NC_.h) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the host process already loaded (e.g. torch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { g_err = std::string("cannot load libnccl.so.2: ") + dlerror(); return -11; }
#define B(f) *(void**)(&NC_.f) = dlsym(h, "nccl" #f); if (!NC_.f) { g_err = "libnccl.so.2 lacks nccl" #f; return -11; }
  B(GetUniqueId) B(CommInitRank) B(CommDestroy) B(GroupStart) B(GroupEnd) B(Broadcast) B(AllReduce) B(GetErrorString)
#undef B
  NC_.h = h;
  return 0;
}
int ue_gpu_comm_unique_id(char* id128) {
  int rc = nccl_bind();
  if (rc) return rc;
  ncclUniqueId id;
  NCK(NC_.GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  std::memcpy(id128, &id, 128);
  return 0;
}
int ue_gpu_comm_init(int64_t nranks, int64_t rank, const char* id128) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (nranks < 1 || rank < 0 || rank >= nranks) { g_err = "comm_init: bad rank / nranks"; return -1; }
  int rc = nccl_bind();
  if (rc) return rc;
  if (g_comm) { NC_.CommDestroy(g_comm); g_comm = nullptr; g_nranks = 1; g_rank = 0; }
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  NCK(NC_.CommInitRank(&g_comm, (int)nranks, id, (int)rank));
  // MPISplitIndex (ppp/mpi_parallel.F90:366-447): contiguous column ranges; the weight of a column is the size of its
  // candidate list (the work of its evaluation), known from the index maps, so no timing feedback is needed
  g_rank_lo.assign(nranks, 1); g_rank_hi.assign(nranks, 0);
  int64_t prev = 0;
  for (int r = 0; r < (int)nranks; ++r) {
    int64_t e = neq;
    if (r < (int)nranks - 1) {
      const int64_t target = (int64_t)((double)g_cap_total * (r + 1) / (double)nranks);
      e = (int64_t)(std::upper_bound(h_coloff.begin(), h_coloff.end(), target) - h_coloff.begin()) - 1;  // last column starting at or before the target
      e = std::max(prev, std::min(e, neq));
    }
    g_rank_lo[r] = prev + 1; g_rank_hi[r] = e; prev = e;
  }
  g_ivmin = g_rank_lo[rank]; g_ivmax = g_rank_hi[rank]; g_nnz_guess = 0;
  build_lists();
  drop_graphs();
  rc = upload_lists();
  if (rc) return rc;
  g_nranks = (int)nranks; g_rank = (int)rank;
  if (g_nranks > 1) {  // first collective outside any timed region: NCCL sets up its channels and buffers here
    CK(cudaMemsetAsync(d_colcnt, 0, 3 * neq * sizeof(int), g_stream));
    rc = comm_gather_columns();
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_stream));
  }
  return 0;
}
int ue_gpu_comm_info(int64_t* nranks, int64_t* rank, int64_t* ivmin, int64_t* ivmax, int64_t* bytes_last_jac) {
  *nranks = g_nranks; *rank = g_rank; *ivmin = g_ivmin; *ivmax = g_ivmax; *bytes_last_jac = g_nranks > 1 ? g_comm_bytes : 0;
  return 0;
}
int ue_gpu_comm_finalize(void) {
  if (g_comm) { cudaStreamSynchronize(g_stream); NC_.CommDestroy(g_comm); g_comm = nullptr; }
  const bool was = g_nranks > 1;
  g_nranks = 1; g_rank = 0;
  if (was && g_ready) { g_ivmin = 1; g_ivmax = neq; build_lists(); drop_graphs(); return upload_lists(); }
  return 0;
}
int ue_gpu_finalize(void) { ue_gpu_comm_finalize(); free_all(); return 0; }
}
