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
//   k_csr                          structural CSR superset -> reference CSR in one pass (csrcsc, svr/svrut4.m:1536-1608)
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <chrono>
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
// Structural CSR superset (built once from the candidate lists): every (row, column) pair a perturbation can reach, sorted by
// row then column.  k_jb_p3c writes each candidate's value and keep flag straight into its slot; one count + one write pass
// compact the kept slots into the reference's CSR (columns ascending): no atomics, no sort.
int2* d_meta = nullptr;  // per slot: (fragment index | row-start flag, column)
int* d_srow = nullptr;
unsigned long long* d_tilest = nullptr;  // per-tile scan state of k_csr + 4 counters behind it
double* d_frag = nullptr;                // column fragments: coloff[iv-1] + candidate index; NaN = entry not kept
int64_t g_nslots = 0;
constexpr int CSR_TILE = 2048;  // slots per block of the compaction kernel
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
// P2P transport (ue_gpu_comm_init_p2p): the fragment arrays of every rank are mapped into every other rank (CUDA IPC over
// NVLink); k_jb_p3c stores its column results into all of them, a flag barrier replaces the collective
bool g_p2p = false;
unsigned long long g_epoch = 0;             // Jacobians assembled since the communicator was set up
double* g_peer_frag[8] = {nullptr};
void* g_peer_base[8] = {nullptr};
unsigned long long* g_p2p_flags = nullptr;  // this rank's block: [0] epoch counter, [1..] = epoch posted by rank r-1; peers' blocks at g_peer_flags
unsigned long long* g_peer_flags[8] = {nullptr};
void* g_xchg = nullptr;                     // one allocation: 2 x fragment array + flag block, exported to the peers
size_t g_xchg_bytes = 0;

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
  double* frag;           // column fragments (this GPU): coloff[iv-1] + candidate index; NaN = not kept
  int npeer;              // > 0: also stored into the peers' fragment arrays over NVLink (ue_gpu_comm_init_p2p)
  double* frag_peer[7];
  int* err;
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
  {  // clear the candidate-row masks of this Jacobian
    const int nthr = gridDim.x * 128, t0 = blockIdx.x * 128 + tid;
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
  {  // clear the candidate-row masks of this Jacobian
    const int nthr = gridDim.x * 5 * P01_ITEMS, t0 = blockIdx.x * 5 * P01_ITEMS + tid;
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
  // difference, diagonal terms, clip (oderhs.m:8685-8706); every structural entry of this column is written: the value if it is
  // kept, NaN if not (a NaN element fails the clip test in the reference as well, so the encoding loses nothing)
  for (int qq = lane; qq < ncand; qq += 32) {
    const int l = nv == 5 ? qq / 5 : (nv == 4 ? qq >> 2 : qq / nv), k = qq - l * nv;  // constant divisors for the two layouts
    const int64_t ii = (int64_t)A.cand_cell[q.coff + l] * nv + k + 1;
    bool keep = false; double val = 0.;
    if (ii >= ii1 && ii <= ii2) {
      const bool written = (A.rmask[q.off + l] >> k) & 1;
      if (written || ii == iv) {
        const double y00 = A.yldot00[ii - 1];
        const double wk = written ? A.rows[(size_t)k * NI + q.off + l] : y00;
        double jacelem = (wk - y00) / dyl;
        if (ii == iv) {
          if (D.iseqalg[iv - 1] * (1 - D.isbcwdt) == 0) jacelem = jacelem - 1 / A.dtuse[iv - 1];
          if (D.nufak > 0 && A.yl[neq] == 1) jacelem = jacelem - D.nufak;
        }
        val = jacelem;
        keep = fabs(jacelem * sf) > D.jaccliplim;
      }
    }
    const double out = keep ? val : __longlong_as_double(0x7ff8000000000000LL);
    A.frag[o + qq] = out;
    for (int p = 0; p < A.npeer; ++p) A.frag_peer[p][o + qq] = out;  // consecutive lanes, consecutive addresses: whole NVLink packets
  }
}

// Cross-GPU barrier of the P2P transport.  Stream order guarantees that this rank's k_jb_p3c (and with it all its stores
// into the peers' fragment arrays) has completed; lane r posts the new epoch into rank r's flag block and then waits until rank r
// has posted the same epoch here.  A stalled peer is reported after ~2 s instead of hanging the GPU.
__global__ void k_xbarrier(unsigned long long* mine, int nranks, int rank, int* err) {
  __shared__ unsigned long long ep;
  if (threadIdx.x == 0) { ep = mine[0] + 1; mine[0] = ep; }
  __syncwarp();
  const int r = threadIdx.x;
  if (r < nranks && r != rank) {
    __threadfence_system();
    unsigned long long* theirs = ((unsigned long long**)(mine + 16))[r];  // peer flag blocks: pointer table at mine[16..]
    *((volatile unsigned long long*)(theirs + 1 + rank)) = ep;
    __threadfence_system();
    volatile unsigned long long* in = mine + 1 + r;
    const long long t0 = clock64();
    while (*in < ep) { if (clock64() - t0 > 4000000000LL) { atomicOr(err, 8); break; } }
  }
  __threadfence_system();
}

// ---- structural superset -> reference CSR (csrcsc, svr/svrut4.m:1536-1608: rows in order, columns ascending) -----------
// The index maps fix a superset of the pattern: slot s = (row srow[s], column scol[s]) in row-major, column-ascending order,
// its value at frag[src[s]].  ONE pass: blocks take tiles of 2048 slots in ticket order, count their kept slots, obtain the number
// of kept slots before the tile by summing the earlier tiles' published counts (state word = epoch | count; earlier
// tickets are always running or finished, so the wait is bounded), stage the kept entries in
// shared memory and write jac / ja in whole lines.  jac_host, ja_host, ia_host: the caller's page-locked arrays (or
// nullptr), written as well so that no copy follows.  The last block resets the ticket and advances the epoch.
__global__ void __launch_bounds__(256) k_csr(const double* __restrict__ frag, const int2* __restrict__ meta, const int* __restrict__ srow, int64_t nslots,
                                             unsigned long long* st, int ntiles, int64_t neq, int64_t nnzmx, double* __restrict__ jac, int64_t* __restrict__ ja,
                                             int64_t* __restrict__ ia, double* __restrict__ jac_host, int64_t* __restrict__ ja_host, int64_t* __restrict__ ia_host,
                                             int* err, long long* hflags) {
  __shared__ double sv[CSR_TILE];
  __shared__ int sc[CSR_TILE];
  __shared__ int swarp[8];
  __shared__ int stile, stot;
  __shared__ long long sbase;
  __shared__ unsigned long long sepoch;
  volatile unsigned long long* ctr = st + ntiles;  // [0] ticket, [1] finished blocks, [2] epoch of the previous launch
  const int tid = threadIdx.x, lane = tid & 31, wib = tid >> 5;
  if (tid == 0) { stile = (int)atomicAdd((unsigned long long*)&ctr[0], 1ULL); sepoch = ctr[2] + 1; }
  __syncthreads();
  const int tile = stile;
  const unsigned long long epoch = sepoch & 0x3fffffffULL;
  const int64_t s0 = (int64_t)tile * CSR_TILE + (int64_t)tid * 8;
  // meta[s] = (index of the slot's value in the fragment array | row-start flag in bit 31, column)
  int2 m[8]; double v[8]; int c = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) m[j] = (s0 + j < nslots) ? meta[s0 + j] : make_int2(-1, 0);
#pragma unroll
  for (int j = 0; j < 8; ++j) { v[j] = (s0 + j < nslots) ? frag[m[j].x & 0x7fffffff] : __longlong_as_double(0x7ff8000000000000LL); c += (v[j] == v[j]); }
  int incl = c;  // inclusive scan over the warp, then over the 8 warps
  for (int off = 1; off < 32; off <<= 1) { const int t = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += t; }
  if (lane == 31) swarp[wib] = incl;
  __syncthreads();
  int wbase = 0, total = 0;
  for (int i = 0; i < 8; ++i) { if (i < wib) wbase += swarp[i]; total += swarp[i]; }
  if (wib == 0) {  // publish this tile's count, then sum the counts of all earlier tiles (32 at a time)
    if (lane == 0) atomicExch(&st[tile], (epoch << 34) | (unsigned long long)total);
    long long run = 0;
    for (int i = lane; i < tile; i += 32) {
      unsigned long long w;
      do { w = ((volatile unsigned long long*)st)[i]; } while ((w >> 34) != epoch);
      run += (long long)(w & 0xffffffffULL);
    }
    for (int off = 16; off; off >>= 1) run += __shfl_down_sync(0xffffffffu, run, off);
    if (lane == 0) { sbase = run; stot = total; }
  }
  int lp = wbase + incl - c;  // kept slots of this tile before this thread's first slot
  __syncthreads();
  const long long base = sbase;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int64_t s = s0 + j;
    if (s < nslots) {
      if (m[j].x < 0) { const int r = srow[s]; ia[r - 1] = base + lp + 1; if (ia_host) ia_host[r - 1] = base + lp + 1; }
      if (v[j] == v[j]) { sv[lp] = v[j]; sc[lp] = m[j].y; ++lp; }
      if (s == nslots - 1) {
        ia[neq] = base + lp + 1; if (ia_host) ia_host[neq] = base + lp + 1;
        hflags[1] = base + lp + 1; hflags[2] = err[0] | err[1]; err[0] = err[1] = 0;  // nnz + 1 and the error bits to mapped host memory
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < stot; i += 256) {
    const long long pos = base + i;
    if (pos < nnzmx) {  // overflow is reported by the host after the sequence
      const double x = sv[i]; const int64_t cidx = sc[i];
      jac[pos] = x; ja[pos] = cidx;
      if (jac_host) { jac_host[pos] = x; ja_host[pos] = cidx; }
    }
  }
  if (tid == 0) {
    __threadfence();
    if (atomicAdd((unsigned long long*)&ctr[1], 1ULL) == (unsigned long long)(ntiles - 1)) { ctr[0] = 0; ctr[1] = 0; ctr[2] = sepoch; __threadfence(); }
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
  if (g_xchg) { cudaFree(g_xchg); g_xchg = nullptr; d_frag = nullptr; g_p2p_flags = nullptr; }
  g_seen_host.clear(); g_last_yldot.clear(); g_base_yl.clear();
  for (auto& h : g_step_host) h.clear();
  g_base_valid = g_base_dev_valid = false;
  for (void* p : g_static_allocs) cudaFree(p);
  g_static_allocs.clear();
  void* ptrs[] = {d_base, d_yl, d_yldot00, d_tmp, d_yldot, d_dtuse, d_ylodt, d_suscal, d_sfscal, d_err, d_cand_cell, d_cand_east, d_item_u, d_guard_items, d_guard_cells, d_coloff,
                  d_meta, d_srow, d_tilest, d_frag, d_ia, d_ja, d_jac};
  for (void* p : ptrs) if (p) cudaFree(p);
  if (d_dtoptv) { cudaFree(d_dtoptv); d_dtoptv = nullptr; }
  for (void* p : {(void*)d_uinfo, (void*)d_priv, (void*)d_jrows, (void*)d_rres, (void*)d_rmask}) if (p) cudaFree(p);
  d_uinfo = nullptr; d_priv = d_jrows = d_rres = nullptr; d_rmask = nullptr;
  d_base = d_yl = d_yldot00 = d_tmp = d_yldot = d_dtuse = d_ylodt = d_suscal = d_sfscal = nullptr;
  d_err = nullptr; d_cand_cell = d_cand_east = d_item_u = d_guard_items = d_guard_cells = nullptr; d_coloff = nullptr; d_meta = nullptr; d_srow = nullptr; d_tilest = nullptr; d_frag = nullptr;
  d_ia = d_ja = nullptr; d_jac = nullptr;
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

// structural CSR superset: slot of every (column iv, candidate) pair in row-major, column-ascending order
int build_superset() {
  const UeParams& P = S.p;
  const int nv = (int)P.numvar;
  std::vector<unsigned long long> key; std::vector<int> where;
  key.reserve((size_t)g_cap_total); where.reserve((size_t)g_cap_total);
  for (int64_t iv = 1; iv <= neq; ++iv) {
    const int c = (int)P.igyl[iv - 1] + NXS * (int)P.igyl[neq + iv - 1];
    const int n = h_cellcand_off[c + 1] - h_cellcand_off[c];
    for (int l = 0; l < n; ++l)
      for (int k = 0; k < nv; ++k) {
        const unsigned long long ii = (unsigned long long)h_cand_cell[h_cellcand_off[c] + l] * nv + k + 1;
        key.push_back((ii << 32) | (unsigned long long)iv);
        where.push_back((int)(h_coloff[iv - 1] + (int64_t)l * nv + k));
      }
  }
  const size_t n = key.size();
  std::vector<int> ord(n);
  for (size_t i = 0; i < n; ++i) ord[i] = (int)i;
  std::sort(ord.begin(), ord.end(), [&](int a, int b) { return key[a] < key[b]; });
  std::vector<int2> meta(n); std::vector<int> srow(n);
  for (size_t r = 0; r < n; ++r) {
    srow[r] = (int)(key[ord[r]] >> 32);
    meta[r].x = where[ord[r]] | ((r == 0 || srow[r - 1] != srow[r]) ? (int)0x80000000 : 0);
    meta[r].y = (int)(key[ord[r]] & 0xffffffffu);
  }
  if (g_cap_total >= 0x7fffffff) { g_err = "grid too large for 31-bit fragment indices"; return -1; }
  g_nslots = (int64_t)n;
  const size_t nt = (n + CSR_TILE - 1) / CSR_TILE;
  CK(cudaMalloc(&d_meta, n * sizeof(int2)));
  CK(cudaMalloc(&d_srow, n * sizeof(int)));
  CK(cudaMalloc(&d_tilest, (nt + 4) * sizeof(unsigned long long)));
  CK(cudaMalloc(&d_frag, std::max<size_t>(1, (size_t)g_cap_total) * sizeof(double)));
  CK(cudaMemcpy(d_meta, meta.data(), n * sizeof(int2), cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_srow, srow.data(), n * sizeof(int), cudaMemcpyHostToDevice));
  CK(cudaMemset(d_tilest, 0, (nt + 4) * sizeof(unsigned long long)));
  CK(cudaMemset(d_frag, 0xff, std::max<size_t>(1, (size_t)g_cap_total) * sizeof(double)));  // all ones = NaN = not kept
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
  int kind; const void *p0, *p1, *p2, *p3, *p4; int64_t a, b, c; int flag; int par;
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
  // the fragment layout (coloff) is the same on all ranks: every rank broadcasts its own range of columns in place
  NCK(NC_.GroupStart());
  int64_t mine = 0, total = 0;
  for (int r = 0; r < g_nranks; ++r) {
    if (g_rank_hi[r] < g_rank_lo[r]) continue;
    const int64_t b0 = h_coloff[g_rank_lo[r] - 1], b1 = g_rank_hi[r] < neq ? h_coloff[g_rank_hi[r]] : g_cap_total;
    NCK(NC_.Broadcast(d_frag + b0, d_frag + b0, (size_t)(b1 - b0), ncclFloat64, r, g_comm, g_stream));
    total += (b1 - b0) * 8;
    if (r == g_rank) mine = (b1 - b0) * 8;
  }
  NCK(NC_.GroupEnd());
  g_comm_bytes = mine * (g_nranks - 1) + (total - mine);  // sent to every peer + received from every peer
  return 0;
}
int res_launches() { return g_fuse23 ? 3 : 4; }  // kernels of one residual sequence with rows
int jac_launches() { return ((int)h_list.size() >= 4096 ? 6 : 4) + (g_p2p ? 1 : 0); }  // kernels of one Jacobian sequence (large / small grids)
int enqueue_jac(const double* dyl, const double* dy00, int64_t ml, int64_t mu, int64_t nnzmx, double* djac, int64_t* dja, int64_t* dia, bool base_current,
                double* jac_host = nullptr, int64_t* ja_host = nullptr, int64_t* ia_host = nullptr) {
  if (!base_current) { int rc = enqueue_residual(dyl, nullptr, false); if (rc) return rc; }
  const int NU = (int)h_list.size();
  if (NU > 0) {
    JArgs A;
    A.ui = (const UInfo*)d_uinfo; A.cand_cell = d_cand_cell; A.cand_east = d_cand_east; A.item_u = d_item_u; A.guard_items = d_guard_items; A.nguard = g_nguard;
    A.NU = NU; A.nitems = g_nitems;
    A.priv = d_priv; A.rows = d_jrows; A.rres = d_rres; A.rmask = d_rmask; A.base = d_base;
    A.yl = dyl; A.yldot00 = dy00; A.suscal = d_suscal; A.sfscal = d_sfscal; A.dtuse = d_dtuse; A.ylodt = d_ylodt;
    A.neq = neq; A.ml = ml; A.mu = mu; A.NXS = NXS; A.NC = NC;
    A.coloff = d_coloff; A.err = d_err + 1;
    const int par = g_p2p ? (int)(g_epoch & 1) : 0;  // P2P: two sets of fragment arrays, alternating with the Jacobian count
    A.frag = d_frag + (size_t)par * g_cap_total;
    A.npeer = 0;
    if (g_p2p) for (int r = 0; r < g_nranks; ++r) if (r != g_rank) A.frag_peer[A.npeer++] = g_peer_frag[r] + (size_t)par * g_cap_total;
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
  const int par = g_p2p ? (int)(g_epoch & 1) : 0;
  if (g_nranks > 1) {  // every rank receives every other rank's slots; the CSR is then built from all columns on every rank
    if (g_p2p) {  // the slots were stored into the peers' arrays by k_jb_p3c: signal, then wait for every peer's signal
      CK(launch(k_xbarrier, dim3(1), dim3(32), g_p2p_flags, g_nranks, g_rank, d_err + 1));
      ++g_epoch;
    } else {
      int rc = comm_gather_columns();
      if (rc) return rc;
    }
  }
  const unsigned nt = (unsigned)((g_nslots + CSR_TILE - 1) / CSR_TILE);
  CK(launch(k_csr, dim3(nt), dim3(256), (const double*)(d_frag + (size_t)par * g_cap_total), (const int2*)d_meta, (const int*)d_srow, g_nslots, d_tilest, (int)nt, neq,
            nnzmx, djac, dja, dia, jac_host, ja_host, ia_host, d_err, d_hflags));
  return 0;
}
template <typename F>
int replay(const GKey& key0, F enqueue) {
  if (g_nranks > 1 && !g_p2p) return enqueue();  // the sequence contains NCCL calls: launched directly, not captured
  GKey key = key0;
  const bool hasjac = key.kind == 2 || key.kind == 4 || key.kind == 5;
  key.par = (g_p2p && hasjac) ? (int)(g_epoch & 1) + 1 : 0;  // P2P: the fragment arrays alternate with the Jacobian count
  auto it = g_graphs.find(key);
  bool fresh = false;
  if (it == g_graphs.end()) {
    cudaGraphExec_t ex = nullptr;
    cudaGraph_t graph = nullptr;
    CK(cudaStreamBeginCapture(g_stream, cudaStreamCaptureModeThreadLocal));
    int rc = enqueue();  // (a captured Jacobian sequence advances g_epoch once: it is launched once right below)
    cudaError_t e = cudaStreamEndCapture(g_stream, &graph);
    if (e == cudaSuccess && rc == 0) e = cudaGraphInstantiate(&ex, graph, 0);
    if (graph) cudaGraphDestroy(graph);
    if (!(e == cudaSuccess && rc == 0)) {
      cudaGetLastError();
      if (rc) return rc;
      g_err = std::string("CUDA graph capture failed: ") + cudaGetErrorString(e);
      return -10;
    }
    if (g_graphs.size() > 64) drop_graphs();
    it = g_graphs.emplace(key, ex).first;
    fresh = true;
  }
  CK(cudaGraphLaunch(it->second, g_stream));
  if (!fresh && g_p2p && hasjac) ++g_epoch;
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
  if (h & 8) { g_err = "multi-GPU Jacobian: a peer GPU did not reach the exchange barrier within 2 s"; return -11; }
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
  const int64_t nnz = (int64_t)h_flags[1] - 1;  // k_csr posts ia(neq+1) to mapped host memory
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
    if (it == S.iscal.end()) { g_err = std::string("unknown int input ") + n; return -1; }  // (a typo must not disable the library)
    if (*it->second == v) return 0;  // unchanged
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
  auto it = S.rscal.find(n);
  if (it == S.rscal.end()) { g_err = std::string("unknown real input ") + n; return -1; }
  const double old = *it->second;
  S.set_real(n, v);
  if (g_ready) {  // scalars such as nufak, dtreal may change between solves: patch the device copy in place
    if (int rc = check_switches()) { *it->second = old; return rc; }  // the refusals of ue_gpu_init hold afterwards too (fnnuiz, difpr2, l_parloss, ...)
    const size_t off = (size_t)((char*)it->second - (char*)&S.p);
    CK(cudaMemcpyToSymbol(D, &v, sizeof(double), off));
  }
  g_base_valid = g_base_dev_valid = false;
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
  if (!b.empty()) { g_err = "bad array sizes (have != expected): " + b; return -1; }
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
  if ((rc = build_superset())) return rc;
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

// Timing accumulators of the reference (com/com.v:500-519): ttotfe = time in residual evaluations, ttotjf = time in Jacobian
// assemblies.  Host clock around the entry points (what the caller waits for).  ttjstor (time "storing" Jacobian elements) has no
// separate phase here - the store is fused into the assembly kernels - and is reported as 0.
double g_ttotfe = 0., g_ttotjf = 0.;
struct Stopwatch {
  double* acc; std::chrono::steady_clock::time_point t0;
  explicit Stopwatch(double* a) : acc(a), t0(std::chrono::steady_clock::now()) {}
  ~Stopwatch() { *acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};
int ue_gpu_timing(double* ttotfe, double* ttotjf, double* ttjstor, int64_t reset) {
  if (ttotfe) *ttotfe = g_ttotfe;
  if (ttotjf) *ttotjf = g_ttotjf;
  if (ttjstor) *ttjstor = 0.;
  if (reset) g_ttotfe = g_ttotjf = 0.;
  return 0;
}
int ue_gpu_pandf1(int64_t n, double time, const double* yl, double* yldot) {
  Stopwatch sw_(&g_ttotfe);
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
  Stopwatch sw_(&g_ttotjf);
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
  // page-locked caller arrays: k_csr stores ia, jac, ja there directly and no copy follows
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
  CK(cudaMemset(d_frag, 0xff, (size_t)g_cap_total * sizeof(double)));  // entries of columns outside the range are never written: NaN = not kept
  return upload_lists();
}

// Device-pointer callers: assert that d_yl has not changed since the last ue_gpu_pandf1_dev call, so the next
// ue_gpu_jac_calc_dev may reuse the base planes (the host-pointer entry points check this themselves).
int ue_gpu_assume_base_current(int64_t flag) { g_jac_trust_base = (flag != 0); return 0; }
// sfsetnk (bbb/oderhs.m:9815-9884) with the Jacobian kept on the device: f0 = pandf1(yl | flag=1), J = jac_calc,
// J <- J*diag(1/su), sf(i) = 1/max_k|J_ik|, ydt_max0 = max_i|f0_i sf_i|.  Only sf (neq doubles) returns to the host.
int ue_gpu_sfsetnk(int64_t n, const double* yl, const double* su, int64_t ml, int64_t mu, double* sf, double* ydt_max0) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq || !yl || !su || !sf || !ydt_max0) { g_err = "sfsetnk: neq mismatch or null pointer"; return -1; }
  if (g_nranks == 1 && (g_ivmin != 1 || g_ivmax != neq)) { g_err = "sfsetnk needs the full Jacobian: reset ue_gpu_set_column_range(1, neq) first"; return -1; }
  std::vector<double> y(yl, yl + neq + 2);
  y[neq] = 1.;  // oderhs.m:9848
  CK(cudaMemcpyAsync(d_yl, y.data(), (neq + 2) * 8, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(d_suscal, su, neq * 8, cudaMemcpyHostToDevice, g_stream));
  g_step_host[2].assign(su, su + neq);  // the resident suscal is now `su`: a later ue_gpu_step_params compares against this
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
// vnormnk(n, v, s) = sqrt(sum((v(i)*s(i))**2)) (svr/nksol.m:1404-1419): NKSOL's scaled norm (fnrm, unrm, pnrm).
// Fixed summation shape (1024 strided partial sums, then a binary tree), so the result does not depend on the launch.
__global__ void __launch_bounds__(1024) k_vnorm(int64_t n, const double* __restrict__ v, const double* __restrict__ s, double* __restrict__ out) {
  __shared__ double part[1024];
  double acc = 0.;
  for (int64_t i = threadIdx.x; i < n; i += 1024) { const double t = v[i] * s[i]; acc = acc + t * t; }
  part[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 512; off; off >>= 1) { if ((int)threadIdx.x < off) part[threadIdx.x] = part[threadIdx.x] + part[threadIdx.x + off]; __syncthreads(); }
  if (threadIdx.x == 0) out[0] = part[0];
}
static int vnorm_dev(const double* dv, const double* ds, double* out) {
  static double* d_sum = nullptr;
  if (!d_sum) CK(cudaMalloc(&d_sum, 8));
  CK(launch(k_vnorm, dim3(1), dim3(1024), neq, dv, ds, d_sum));
  g_launches += 1;
  double sum = 0.;
  CK(cudaMemcpyAsync(&sum, d_sum, 8, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  *out = std::sqrt(sum);
  return 0;
}
int ue_gpu_vnormnk(int64_t n, const double* v, const double* s, double* out) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq || !v || !s || !out) { g_err = "vnormnk: neq mismatch or null pointer"; return -1; }
  static double* d_vs = nullptr; static int64_t cap = 0;
  if (cap < 2 * neq) { if (d_vs) cudaFree(d_vs); CK(cudaMalloc(&d_vs, 2 * neq * 8)); cap = 2 * neq; }
  CK(cudaMemcpyAsync(d_vs, v, neq * 8, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(d_vs + neq, s, neq * 8, cudaMemcpyHostToDevice, g_stream));
  return vnorm_dev(d_vs, d_vs + neq, out);
}
int ue_gpu_fnrm(double* out) {  // fnrm = vnormnk(neq, savf, sf) of the residual the last ue_gpu_pandf1 left on the device
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (!out) { g_err = "fnrm: null pointer"; return -1; }
  if (g_last_yldot.empty()) { g_err = "fnrm: no residual has been evaluated yet"; return -1; }
  return vnorm_dev(d_yldot, d_sfscal, out);
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
// CSC copy of the last Jacobian: rcsc / icsc / jcsc of jac_calc (oderhs.m:8620-8752; read by jacmap and the ppp debug dumps).
// Built on the host from the column fragments the assembly kernel left on the device (NaN = element not kept); rows ascending.
int ue_gpu_get_csc(int64_t nnzmx, double* rcsc, int64_t* icsc, int64_t* jcsc, int64_t* nnz_out) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (!rcsc || !icsc || !jcsc || !nnz_out) { g_err = "get_csc: null pointer"; return -1; }
  if (h_flags[1] <= 0) { g_err = "get_csc: no Jacobian has been assembled yet"; return -1; }
  const int par = g_p2p ? (int)((g_epoch + 1) & 1) : 0;  // the set the last Jacobian wrote
  std::vector<double> frag((size_t)g_cap_total);
  CK(cudaStreamSynchronize(g_stream));
  CK(cudaMemcpy(frag.data(), d_frag + (size_t)par * g_cap_total, (size_t)g_cap_total * 8, cudaMemcpyDeviceToHost));
  const int nv = (int)S.p.numvar;
  int64_t nnz = 0;
  std::vector<std::pair<int64_t, double>> col;
  for (int64_t iv = 1; iv <= neq; ++iv) {
    jcsc[iv - 1] = nnz + 1;
    const int c = (int)S.p.igyl[iv - 1] + NXS * (int)S.p.igyl[neq + iv - 1];
    const int n = h_cellcand_off[c + 1] - h_cellcand_off[c];
    col.clear();
    for (int l = 0; l < n; ++l)
      for (int k = 0; k < nv; ++k) {
        const double v = frag[(size_t)h_coloff[iv - 1] + (size_t)l * nv + k];
        if (v == v) col.push_back({(int64_t)h_cand_cell[h_cellcand_off[c] + l] * nv + k + 1, v});
      }
    std::sort(col.begin(), col.end());
    for (auto& e : col) { if (nnz < nnzmx) { icsc[nnz] = e.first; rcsc[nnz] = e.second; } ++nnz; }
  }
  jcsc[neq] = nnz + 1;
  *nnz_out = nnz;
  if (nnz > nnzmx) { g_err = "get_csc: nnzmx too small"; return -2; }
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
static int comm_set_ranges(int64_t nranks, int64_t rank) {
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
  return upload_lists();
}
static void comm_release() {
  if (g_comm) { cudaStreamSynchronize(g_stream); NC_.CommDestroy(g_comm); g_comm = nullptr; }
  if (g_p2p) {
    cudaStreamSynchronize(g_stream);
    for (int r = 0; r < 8; ++r) { if (g_peer_base[r]) cudaIpcCloseMemHandle(g_peer_base[r]); g_peer_base[r] = nullptr; g_peer_frag[r] = nullptr; g_peer_flags[r] = nullptr; }
    g_p2p = false;
  }
}
int ue_gpu_comm_init(int64_t nranks, int64_t rank, const char* id128) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (nranks < 1 || rank < 0 || rank >= nranks) { g_err = "comm_init: bad rank / nranks"; return -1; }
  int rc = nccl_bind();
  if (rc) return rc;
  comm_release(); g_nranks = 1; g_rank = 0;
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  NCK(NC_.CommInitRank(&g_comm, (int)nranks, id, (int)rank));
  rc = comm_set_ranges(nranks, rank);
  if (rc) return rc;
  g_nranks = (int)nranks; g_rank = (int)rank;
  if (g_nranks > 1) {  // first collective outside any timed region: NCCL sets up its channels and buffers here
    rc = comm_gather_columns();
    if (rc) return rc;
    CK(cudaStreamSynchronize(g_stream));
  }
  return 0;
}
// ---- P2P transport: the fragment arrays of all ranks mapped into each other (CUDA IPC, NVLink peer access) ---------------------
static size_t xchg_flags_off() { return (((size_t)g_cap_total * 2 * 8) + 255) / 256 * 256; }
int ue_gpu_comm_p2p_handle(char* handle64) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  comm_release(); g_nranks = 1; g_rank = 0;
  if (!g_xchg) {  // move the fragment array into one exportable allocation: 2 x fragments | flag block
    g_xchg_bytes = xchg_flags_off() + 512;
    CK(cudaMalloc(&g_xchg, g_xchg_bytes));
    cudaFree(d_frag);
    d_frag = (double*)g_xchg;
    drop_graphs();
  }
  CK(cudaMemset(g_xchg, 0xff, xchg_flags_off()));
  CK(cudaMemset((char*)g_xchg + xchg_flags_off(), 0, 512));
  g_p2p_flags = (unsigned long long*)((char*)g_xchg + xchg_flags_off());
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, g_xchg));
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  std::memcpy(handle64, &h, 64);
  return 0;
}
int ue_gpu_comm_init_p2p(int64_t nranks, int64_t rank, const char* handles) {
  if (!g_ready || !g_xchg) { g_err = "comm_init_p2p: call ue_gpu_comm_p2p_handle on every rank first"; return -1; }
  if (nranks < 1 || nranks > 8 || rank < 0 || rank >= nranks) { g_err = "comm_init_p2p: 1..8 ranks of one node"; return -1; }
  unsigned long long table[8] = {0};
  for (int r = 0; r < (int)nranks; ++r) {
    if (r == (int)rank) { table[r] = (unsigned long long)g_p2p_flags; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handles + 64 * r, 64);
    void* base = nullptr;
    CK(cudaIpcOpenMemHandle(&base, h, cudaIpcMemLazyEnablePeerAccess));
    g_peer_base[r] = base;
    g_peer_frag[r] = (double*)base;
    g_peer_flags[r] = (unsigned long long*)((char*)base + xchg_flags_off());
    table[r] = (unsigned long long)g_peer_flags[r];
  }
  CK(cudaMemcpy(g_p2p_flags + 16, table, sizeof table, cudaMemcpyHostToDevice));
  int rc = comm_set_ranges(nranks, rank);
  if (rc) return rc;
  g_nranks = (int)nranks; g_rank = (int)rank; g_p2p = nranks > 1; g_epoch = 0; g_comm_bytes = 0;
  return 0;  // the caller synchronises the ranks (MPI_Barrier / dist.barrier) before the first Jacobian
}
int ue_gpu_comm_info(int64_t* nranks, int64_t* rank, int64_t* ivmin, int64_t* ivmax, int64_t* bytes_last_jac) {
  *nranks = g_nranks; *rank = g_rank; *ivmin = g_ivmin; *ivmax = g_ivmax; *bytes_last_jac = g_nranks > 1 ? (g_p2p ? (int64_t)9 * (g_nranks - 1) * ((g_ivmax < neq ? h_coloff[g_ivmax] : g_cap_total) - h_coloff[g_ivmin - 1]) / UE_NV * (int64_t)S.p.numvar : g_comm_bytes) : 0;
  return 0;
}
int ue_gpu_comm_finalize(void) {
  comm_release();
  const bool was = g_nranks > 1;
  g_nranks = 1; g_rank = 0;
  if (was && g_ready) { g_ivmin = 1; g_ivmax = neq; build_lists(); drop_graphs(); return upload_lists(); }
  return 0;
}
int ue_gpu_finalize(void) { ue_gpu_comm_finalize(); free_all(); return 0; }
}
