// uedge_b200/csrc/ue_gpu.cu — C ABI (include/ue_gpu.h) + kernels of the B200 pandf1 / jac_calc path.
//
// Replaces the bodies of Pandf1rhs_interface (bbb/oderhs.m:12331-12345) and
// jac_calc_interface (bbb/oderhs.m:12297-12329).  No CPU fallback: every entry
// point fails if the CUDA device is unavailable.
//
// Kernels (all FP64, -fmad=false; see ue_device.cuh for the physics):
//   k_phase0/1/2/3      full residual, one thread per cell, SoA planes in HBM
//   k_jac<BLOCK>        batched Jacobian: one thread block per perturbed unknown, all
//                       unknowns ("every colour") in flight in one launch; window box staged in
//                       shared memory; ordered in-block compaction of the column
//   k_scan / k_fill / k_sortrows   CSC fragments -> reference CSR (csrcsc, svr/svrut4.m:1536-1608)
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "ue_device.cuh"
#include "ue_gpu.h"


namespace {

UeStore S;
std::string g_err;
bool g_ready = false;
int nx, ny, NXS, NC;
int64_t neq = 0;
int64_t g_launches = 0;
float g_jac_ms = 0.f, g_res_ms = 0.f;
cudaStream_t g_stream = nullptr;
cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;

std::vector<void*> g_static_allocs;
double *d_base = nullptr, *d_yl = nullptr, *d_yldot00 = nullptr, *d_tmp = nullptr, *d_yldot = nullptr;
double *d_dtuse = nullptr, *d_ylodt = nullptr, *d_suscal = nullptr, *d_sfscal = nullptr;
int* d_err = nullptr;
// Jacobian work space
int64_t g_ivmin = 1, g_ivmax = 0;
std::vector<int> h_list_narrow, h_list_wide;
std::vector<int64_t> h_coloff;  // per column offset into the fragment buffers (1-based iv -> h_coloff[iv-1])
int *d_list_narrow = nullptr, *d_list_wide = nullptr;
int64_t* d_coloff = nullptr;
int *d_colcnt = nullptr, *d_colrow = nullptr;
double* d_colval = nullptr;
int *d_rowcnt = nullptr, *d_rowfill = nullptr;
int64_t *d_ia = nullptr, *d_ja = nullptr;
double* d_jac = nullptr;
int64_t g_cap_total = 0, g_nnzcap = 0;
int g_box_narrow = 0, g_box_wide = 0, g_ext_narrow = 0, g_ext_wide = 0;
size_t g_smem_narrow = 0, g_smem_wide = 0;

#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess) {                                                                       \
      g_err = std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " #call;                 \
      return -10;                                                                                  \
    }                                                                                              \
  } while (0)

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
__global__ void k_phase0(double* base, const double* __restrict__ yl, int NXS, int NC, int* err) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= NC) return;
  Acc<false> a{base, nullptr, NXS, NC, 0, 0, 0, 0};
  phase0_cell<false>(a, yl + (size_t)c * UE_NV, c % NXS, c / NXS, err);
}
__global__ void k_phase1(double* base, int NXS, int NC) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= NC) return;
  Acc<false> a{base, nullptr, NXS, NC, 0, 0, 0, 0};
  const Win w = make_win(D, -1, -1);
  phase1_cell<false>(a, w, c % NXS, c / NXS);
}
__global__ void k_phase2(double* base, double* __restrict__ tmp, int NXS, int NC) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= NC) return;
  Acc<false> a{base, nullptr, NXS, NC, 0, 0, 0, 0};
  const Win w = make_win(D, -1, -1);
  const int ix = c % NXS, iy = c / NXS;
  double r[UE_NV] = {0., 0., 0., 0., 0.};
  if (ix >= 1 && ix <= D.nx && iy >= 1 && iy <= D.ny) {
    phase2_interior<false>(a, w, ix, iy, r, D.iseqalg);
    double v;
    if (rightplate_up<false>(a, w, ix, iy, v)) r[1] = v;
  } else {
    phase2_guard<false>(a, w, ix, iy, r);
  }
  for (int k = 0; k < UE_NV; ++k) tmp[(size_t)c * UE_NV + k] = r[k];
}
__global__ void k_phase3(double* base, const double* __restrict__ tmp, double* __restrict__ yldot, const double* __restrict__ yl,
                         const double* __restrict__ dtuse, const double* __restrict__ ylodt, int64_t neq, int NXS, int NC) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= NC) return;
  Acc<false> a{base, nullptr, NXS, NC, 0, 0, 0, 0};
  const int ix = c % NXS, iy = c / NXS;
  double r[UE_NV];
  for (int k = 0; k < UE_NV; ++k) r[k] = tmp[(size_t)c * UE_NV + k];
  if (ix >= 1 && ix <= D.nx && iy >= 1 && iy <= D.ny) phase3_interior<false>(a, ix, iy, r, yl + (size_t)c * UE_NV, yl[neq], D.iseqalg, dtuse, ylodt);
  for (int k = 0; k < UE_NV; ++k) yldot[(size_t)c * UE_NV + k] = r[k];
}

// ------------------------------------------------------------------------------------------------
// batched Jacobian kernel: one block per perturbed unknown
// ------------------------------------------------------------------------------------------------
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_jac(const int* __restrict__ ivlist, double* base, const double* __restrict__ yl,
                                               const double* __restrict__ yldot00, const double* __restrict__ suscal,
                                               const double* __restrict__ sfscal, const double* __restrict__ dtuse,
                                               const double* __restrict__ ylodt, int64_t neq, int64_t ml, int64_t mu, int NXS, int NC,
                                               const int64_t* __restrict__ coloff, int* __restrict__ colcnt, int* __restrict__ colrow,
                                               double* __restrict__ colval, int* __restrict__ rowcnt, int* err) {
  extern __shared__ double smem[];
  __shared__ int s_scan[BLOCK];
  const int tid = threadIdx.x;
  const int64_t iv = ivlist[blockIdx.x];  // 1-based unknown
  const int xc = (int)D.igyl[iv - 1], yc = (int)D.igyl[neq + iv - 1];
  const Win w = make_win(D, xc, yc);
  const int nxx = (int)D.nx, nyy = (int)D.ny;
  Acc<true> a;
  a.base = base; a.NXS = NXS; a.NC = NC;
  a.bx0 = w.i1; a.by0 = w.j1; a.bw = w.i6 - w.i1 + 1; a.bh = w.j6 - w.j1 + 1;
  const int bsz = a.bw * a.bh;
  a.sm = smem;
  // extended box of candidate rows
  const int ex0 = max(0, w.i2 - 1), ex1 = min(nxx + 1, w.i5 + 1), ey0 = max(0, w.j2 - 1), ey1 = min(nyy + 1, w.j5 + 1);
  const int ew = ex1 - ex0 + 1, eh = ey1 - ey0 + 1, ecells = ew * eh;
  double* rows = smem + (size_t)PL_COUNT * bsz;          // [ecells][UE_NV]
  int* rmask = (int*)(rows + (size_t)ecells * UE_NV);     // [ecells]
  // ---- stage the window box of every field from the base planes -------------------------------
  for (int q = tid; q < PL_COUNT * bsz; q += BLOCK) {
    const int pl = q / bsz, l = q - pl * bsz;
    const int lx = l % a.bw, ly = l / a.bw;
    smem[q] = base[(size_t)pl * NC + (a.bx0 + lx) + NXS * (a.by0 + ly)];
  }
  // ---- perturbation (oderhs.m:8676-8678) -------------------------------------------------------
  const double yold = yl[iv - 1];
  const double dyl = D.delpert * (fabs(yold) + D.dylconst / suscal[iv - 1]);
  __syncthreads();
  if (tid == 0) {
    double ycell[UE_NV];
    const int64_t c = (int64_t)(xc + NXS * yc) * UE_NV;
    for (int k = 0; k < UE_NV; ++k) ycell[k] = yl[c + k];
    ycell[(iv - 1) - c] = yold + dyl;
    phase0_cell<true>(a, ycell, xc, yc, err);
  }
  __syncthreads();
  // ---- phase 1 over the box ----------------------------------------------------------------------
  for (int l = tid; l < bsz; l += BLOCK) phase1_cell<true>(a, w, a.bx0 + l % a.bw, a.by0 + l / a.bw);
  __syncthreads();
  // ---- phase 2 over the extended box ---------------------------------------------------------------
  for (int l = tid; l < ecells; l += BLOCK) {
    const int ix = ex0 + l % ew, iy = ey0 + l / ew;
    double r[UE_NV] = {0., 0., 0., 0., 0.};
    int m = 0;
    if (ix >= 1 && ix <= nxx && iy >= 1 && iy <= nyy) {
      if (in_rng(ix, w.i2, w.i5) && in_rng(iy, w.j2, w.j5)) {
        phase2_interior<true>(a, w, ix, iy, r, D.iseqalg);
        m = 0x1f;
        double v;
        if (rightplate_up<true>(a, w, ix, iy, v)) r[1] = v;
      }
    } else {
      m = phase2_guard<true>(a, w, ix, iy, r);
    }
    for (int k = 0; k < UE_NV; ++k) rows[(size_t)l * UE_NV + k] = r[k];
    rmask[l] = m;
  }
  __syncthreads();
  // ---- phase 3 (rscalf + dt term) on interior cells of the window --------------------------------------
  // the perturbed yl differs from yl only in entry iv; rscalf / the dt term read yl of their own cell only
  for (int l = tid; l < ecells; l += BLOCK) {
    const int ix = ex0 + l % ew, iy = ey0 + l / ew;
    if (ix >= 1 && ix <= nxx && iy >= 1 && iy <= nyy && in_rng(ix, w.i2, w.i5) && in_rng(iy, w.j2, w.j5)) {
      double r[UE_NV], ycell[UE_NV];
      const int64_t c = (int64_t)(ix + NXS * iy) * UE_NV;
      for (int k = 0; k < UE_NV; ++k) { r[k] = rows[(size_t)l * UE_NV + k]; ycell[k] = yl[c + k]; }
      if (ix == xc && iy == yc) ycell[(iv - 1) - c] = yold + dyl;
      phase3_interior<true>(a, ix, iy, r, ycell, yl[neq], D.iseqalg, dtuse, ylodt);
      for (int k = 0; k < UE_NV; ++k) rows[(size_t)l * UE_NV + k] = r[k];
    }
  }
  __syncthreads();
  // ---- difference, clip, ordered compaction into this column's CSC fragment (oderhs.m:8685-8719) -------------
  const int64_t ii1 = max(iv - mu, (int64_t)1), ii2 = min(iv + ml, neq);
  const int ncand = ecells * UE_NV;
  const int chunk = (ncand + BLOCK - 1) / BLOCK;
  const int q0 = tid * chunk, q1 = min(ncand, q0 + chunk);
  const double sf = sfscal[iv - 1];
  auto eval = [&](int q, double& val, int64_t& ii) -> bool {
    const int l = q / UE_NV, k = q - l * UE_NV;
    const int ix = ex0 + l % ew, iy = ey0 + l / ew;
    ii = ((int64_t)(ix + NXS * iy)) * UE_NV + k + 1;
    if (ii < ii1 || ii > ii2) return false;
    const bool written = (rmask[l] >> k) & 1;
    if (!written && ii != iv) return false;
    const double y00 = yldot00[ii - 1];
    const double wk = written ? rows[q] : y00;
    double jacelem = (wk - y00) / dyl;
    if (ii == iv) {
      if (D.iseqalg[iv - 1] * (1 - D.isbcwdt) == 0) jacelem = jacelem - 1 / dtuse[iv - 1];
      if (D.nufak > 0 && yl[neq] == 1) jacelem = jacelem - D.nufak;
    }
    val = jacelem;
    return fabs(jacelem * sf) > D.jaccliplim;
  };
  int cnt = 0;
  for (int q = q0; q < q1; ++q) { double v; int64_t ii; if (eval(q, v, ii)) ++cnt; }
  s_scan[tid] = cnt;
  __syncthreads();
  // exclusive scan (BLOCK <= 512; simple Hillis-Steele)
  for (int off = 1; off < BLOCK; off <<= 1) {
    int v = (tid >= off) ? s_scan[tid - off] : 0;
    __syncthreads();
    s_scan[tid] += v;
    __syncthreads();
  }
  int pos = s_scan[tid] - cnt;
  const int64_t o = coloff[iv - 1];
  for (int q = q0; q < q1; ++q) {
    double v; int64_t ii;
    if (eval(q, v, ii)) {
      colrow[o + pos] = (int)ii;
      colval[o + pos] = v;
      atomicAdd(&rowcnt[ii - 1], 1);
      ++pos;
    }
  }
  if (tid == BLOCK - 1) colcnt[iv - 1] = s_scan[tid];
}

// ---- CSC fragments -> CSR ---------------------------------------------------------------------------------------
__global__ void k_scan(const int* __restrict__ rowcnt, int64_t* __restrict__ ia, int64_t n) {
  // single block; ia is 1-based: ia[0] = 1, ia[i+1] = ia[i] + rowcnt[i]
  __shared__ int64_t s[1024];
  __shared__ int64_t carry;
  if (threadIdx.x == 0) { carry = 1; ia[0] = 1; }
  __syncthreads();
  for (int64_t b = 0; b < n; b += 1024) {
    const int64_t i = b + threadIdx.x;
    s[threadIdx.x] = (i < n) ? rowcnt[i] : 0;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      int64_t v = (threadIdx.x >= off) ? s[threadIdx.x - off] : 0;
      __syncthreads();
      s[threadIdx.x] += v;
      __syncthreads();
    }
    if (i < n) ia[i + 1] = carry + s[threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 1023) carry += s[1023];
    __syncthreads();
  }
}
__global__ void k_fill(int64_t neq, int64_t ivmin, int64_t ivmax, const int64_t* __restrict__ coloff, const int* __restrict__ colcnt,
                       const int* __restrict__ colrow, const double* __restrict__ colval, const int64_t* __restrict__ ia, int* __restrict__ rowfill,
                       double* __restrict__ jac, int64_t* __restrict__ ja) {
  const int64_t iv = ivmin + blockIdx.x;
  if (iv > ivmax) return;
  const int64_t o = coloff[iv - 1];
  const int n = colcnt[iv - 1];
  for (int e = threadIdx.x; e < n; e += blockDim.x) {
    const int row = colrow[o + e];
    const int64_t p = ia[row - 1] - 1 + atomicAdd(&rowfill[row - 1], 1);
    ja[p] = iv;
    jac[p] = colval[o + e];
  }
}
__global__ void k_sortrows(int64_t neq, const int64_t* __restrict__ ia, double* __restrict__ jac, int64_t* __restrict__ ja) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= neq) return;
  const int64_t b = ia[r] - 1, e = ia[r + 1] - 1;
  for (int64_t i = b + 1; i < e; ++i) {  // insertion sort by column: rows hold a few tens of entries
    const int64_t cj = ja[i]; const double cv = jac[i];
    int64_t j = i - 1;
    while (j >= b && ja[j] > cj) { ja[j + 1] = ja[j]; jac[j + 1] = jac[j]; --j; }
    ja[j + 1] = cj; jac[j + 1] = cv;
  }
}

// ------------------------------------------------------------------------------------------------
int check_switches() {
  const UeParams& P = S.p;
  struct { const char* n; int64_t v, want; } eq[] = {
      {"nisp", P.nisp, 1}, {"nusp", P.nusp, 1}, {"ngsp", P.ngsp, 1}, {"numvar", P.numvar, UE_NV}, {"isnonog", P.isnonog, 0}, {"isphion", P.isphion, 0},
      {"isphiofft", P.isphiofft, 0}, {"isimpon", P.isimpon, 0}, {"isupgon", P.isupgon, 0}, {"isngon", P.isngon, 1}, {"istgon", P.istgon, 0},
      {"ineudif", P.ineudif, 2}, {"isflxvar", P.isflxvar, 0}, {"ismcnon", P.ismcnon, 0}, {"ifixsrc", P.ifixsrc, 0}, {"ifixpsor", P.ifixpsor, 0},
      {"ishymol", P.ishymol, 0}, {"ishosor", P.ishosor, 0}, {"isupdrag", P.isupdrag, 0}, {"isofric", P.isofric, 0}, {"jhswitch", P.jhswitch, 0},
      {"isfeexpl0", P.isfeexpl0, 0}, {"isfeixpl0", P.isfeixpl0, 0}, {"is1D_gbx", P.is1D_gbx, 0}, {"isnglf", P.isnglf, 0}, {"isudsym", P.isudsym, 0},
      {"islimon", P.islimon, 0}, {"isdifbetap", P.isdifbetap, 0}, {"isugfm1side", P.isugfm1side, 0}, {"nxomit", P.nxomit, 0}, {"isfixlb", P.isfixlb, 0},
      {"isfixrb", P.isfixrb, 0}, {"isextrnp", P.isextrnp, 0}, {"isextrnpf", P.isextrnpf, 0}, {"isextrtpf", P.isextrtpf, 0}, {"isextrngc", P.isextrngc, 0},
      {"isextrnw", P.isextrnw, 0}, {"isextrtw", P.isextrtw, 0}, {"isnfmiy", P.isnfmiy, 0}, {"isybdrywd", P.isybdrywd, 0}, {"isnewpot", P.isnewpot, 0},
      {"isbohmms", P.isbohmms, 0}, {"isgpye", P.isgpye, 0}, {"isngcore", P.isngcore, 0}, {"ibctepl", P.ibctepl, 1}, {"ibctipl", P.ibctipl, 1},
      {"ibctepr", P.ibctepr, 1}, {"ibctipr", P.ibctipr, 1}, {"iskaplex", P.iskaplex, 0}, {"isnupdot1sd", P.isnupdot1sd, 0}};
  for (auto& e : eq) if (e.v != e.want) { g_err = std::string("switch outside the built hot path: ") + e.n; return -5; }
  if (P.isbohmcalc != 0 && P.isbohmcalc != 1) { g_err = "isbohmcalc must be 0/1 with facb*=0"; return -5; }
  if (P.isnicore != 0 && P.isnicore != 1) { g_err = "isnicore must be 0 or 1"; return -5; }
  if (P.isupcore != 0 && P.isupcore != 1) { g_err = "isupcore must be 0 or 1"; return -5; }
  if (P.iflcore != 0 && P.iflcore != 1) { g_err = "iflcore must be 0 or 1"; return -5; }
  if (P.istabon != 0 && P.istabon != 10) { g_err = "istabon must be 0 or 10"; return -5; }
  if (P.difpr2 != 0 || P.difni2 != 0 || P.difax != 0 || P.dif4order != 0 || P.kye4order != 0 || P.kyi4order != 0) { g_err = "difpr2/difni2/difax/4th-order terms not built"; return -5; }
  if (P.l_parloss <= 1e9) { g_err = "l_parloss<=1e9 (nuvl) not built"; return -5; }
  if (P.yinc >= 6 || P.xrinc >= 20) { g_err = "yinc>=6 / xrinc>=20 windows not built"; return -5; }
  for (int m : {(int)P.methn, (int)P.methu, (int)P.methe, (int)P.methi, (int)P.methg}) {
    int mx = m % 10, my = m / 10;
    if ((mx != 2 && mx != 3) || (my != 2 && my != 3)) { g_err = "meth* must use schemes 2 (central) or 3 (upwind)"; return -5; }
  }
  for (int ix = 0; ix < NXS; ++ix) {
    if (P.matwalli[ix] != 0 || P.matwallo[ix] != 0) { g_err = "matwalli/matwallo>0 not built"; return -5; }
    if (P.fngysi[ix] != 0 || P.fngyso[ix] != 0 || P.fngyi_use[ix] != 0 || P.fngyo_use[ix] != 0) { g_err = "wall gas sources not built"; return -5; }
    if (P.isnwconiix[ix] != 0 || P.isnwconoix[ix] != 0) { g_err = "isnwconi/o != 0 not built"; return -5; }
    if (P.istepfcix[ix] > 1 || P.istipfcix[ix] > 1 || P.istewcix[ix] > 1 || P.istiwcix[ix] > 1) { g_err = "istepfc/istewc > 1 not built"; return -5; }
  }
  for (int iy = 0; iy < ny + 2; ++iy)
    if (P.recylb[iy] < -1. || P.recyrb[iy] < -1.) { g_err = "recylb/recyrb < -1 not built"; return -5; }
  return 0;
}

void free_all() {
  for (void* p : g_static_allocs) cudaFree(p);
  g_static_allocs.clear();
  void* ptrs[] = {d_base, d_yl, d_yldot00, d_tmp, d_yldot, d_dtuse, d_ylodt, d_suscal, d_sfscal, d_err, d_list_narrow, d_list_wide, d_coloff,
                  d_colcnt, d_colrow, d_colval, d_rowcnt, d_rowfill, d_ia, d_ja, d_jac};
  for (void* p : ptrs) if (p) cudaFree(p);
  d_base = d_yl = d_yldot00 = d_tmp = d_yldot = d_dtuse = d_ylodt = d_suscal = d_sfscal = nullptr;
  d_err = nullptr; d_list_narrow = d_list_wide = nullptr; d_coloff = nullptr; d_colcnt = d_colrow = nullptr; d_colval = nullptr;
  d_rowcnt = d_rowfill = nullptr; d_ia = d_ja = nullptr; d_jac = nullptr;
  g_ready = false;
}

// classify unknowns by window width and lay out the per-column fragment buffers
int build_lists() {
  const UeParams& P = S.p;
  h_list_narrow.clear(); h_list_wide.clear();
  h_coloff.assign(neq, 0);
  int64_t off = 0;
  g_box_narrow = g_box_wide = g_ext_narrow = g_ext_wide = 0;
  for (int64_t iv = 1; iv <= neq; ++iv) {
    const int xc = (int)P.igyl[iv - 1], yc = (int)P.igyl[neq + iv - 1];
    const Win w = make_win(P, xc, yc);
    const int bsz = (w.i6 - w.i1 + 1) * (w.j6 - w.j1 + 1);
    const int ex0 = std::max(0, w.i2 - 1), ex1 = std::min(nx + 1, w.i5 + 1), ey0 = std::max(0, w.j2 - 1), ey1 = std::min(ny + 1, w.j5 + 1);
    const int ecells = (ex1 - ex0 + 1) * (ey1 - ey0 + 1);
    const bool wide = (w.i6 - w.i1 + 1) > 8;
    h_coloff[iv - 1] = off;
    off += (int64_t)ecells * UE_NV;
    if (iv < g_ivmin || iv > g_ivmax) continue;
    if (wide) { h_list_wide.push_back((int)iv); g_box_wide = std::max(g_box_wide, bsz); g_ext_wide = std::max(g_ext_wide, ecells); }
    else { h_list_narrow.push_back((int)iv); g_box_narrow = std::max(g_box_narrow, bsz); g_ext_narrow = std::max(g_ext_narrow, ecells); }
  }
  g_cap_total = off;
  auto smem_of = [](int bsz, int ecells) { return (size_t)PL_COUNT * bsz * 8 + (size_t)ecells * UE_NV * 8 + (size_t)ecells * 4 + 16; };
  g_smem_narrow = smem_of(g_box_narrow, g_ext_narrow);
  g_smem_wide = smem_of(g_box_wide, g_ext_wide);
  return 0;
}

int upload_lists() {
  if (d_list_narrow) { cudaFree(d_list_narrow); d_list_narrow = nullptr; }
  if (d_list_wide) { cudaFree(d_list_wide); d_list_wide = nullptr; }
  CK(cudaMalloc(&d_list_narrow, std::max<size_t>(1, h_list_narrow.size()) * sizeof(int)));
  CK(cudaMalloc(&d_list_wide, std::max<size_t>(1, h_list_wide.size()) * sizeof(int)));
  if (!h_list_narrow.empty()) CK(cudaMemcpy(d_list_narrow, h_list_narrow.data(), h_list_narrow.size() * sizeof(int), cudaMemcpyHostToDevice));
  if (!h_list_wide.empty()) CK(cudaMemcpy(d_list_wide, h_list_wide.data(), h_list_wide.size() * sizeof(int), cudaMemcpyHostToDevice));
  return 0;
}

int run_residual_dev(const double* dyl, double* dyldot, bool need_rows) {
  const int B = 128, G = (NC + B - 1) / B;
  CK(cudaMemsetAsync(d_err, 0, sizeof(int), g_stream));
  k_phase0<<<G, B, 0, g_stream>>>(d_base, dyl, NXS, NC, d_err);
  k_phase1<<<G, B, 0, g_stream>>>(d_base, NXS, NC);
  k_phase2<<<G, B, 0, g_stream>>>(d_base, d_tmp, NXS, NC);
  g_launches += 3;
  if (need_rows) { k_phase3<<<G, B, 0, g_stream>>>(d_base, d_tmp, dyldot, dyl, d_dtuse, d_ylodt, neq, NXS, NC); g_launches += 1; }
  CK(cudaGetLastError());
  return 0;
}

int check_errflag() {
  int h = 0;
  CK(cudaMemcpyAsync(&h, d_err, sizeof(int), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  if (h & 1) { g_err = "***  ni is negative - calculation stopped"; return -3; }
  if (h & 2) { g_err = "***  ng is negative - calculation stopped"; return -3; }
  return 0;
}

int run_jac_dev(const double* dyl, const double* dy00, int64_t ml, int64_t mu, int64_t nnzmx, double* djac, int64_t* dja, int64_t* dia, int64_t* nnz_out) {
  int rc = run_residual_dev(dyl, nullptr, false);  // base planes at yl
  if (rc) return rc;
  CK(cudaMemsetAsync(d_rowcnt, 0, neq * sizeof(int), g_stream));
  CK(cudaMemsetAsync(d_rowfill, 0, neq * sizeof(int), g_stream));
  CK(cudaMemsetAsync(d_colcnt, 0, neq * sizeof(int), g_stream));
  if (!h_list_narrow.empty()) {
    k_jac<64><<<(unsigned)h_list_narrow.size(), 64, g_smem_narrow, g_stream>>>(d_list_narrow, d_base, dyl, dy00, d_suscal, d_sfscal, d_dtuse, d_ylodt, neq, ml, mu,
                                                                              NXS, NC, d_coloff, d_colcnt, d_colrow, d_colval, d_rowcnt, d_err);
    g_launches += 1;
  }
  if (!h_list_wide.empty()) {
    k_jac<256><<<(unsigned)h_list_wide.size(), 256, g_smem_wide, g_stream>>>(d_list_wide, d_base, dyl, dy00, d_suscal, d_sfscal, d_dtuse, d_ylodt, neq, ml, mu,
                                                                            NXS, NC, d_coloff, d_colcnt, d_colrow, d_colval, d_rowcnt, d_err);
    g_launches += 1;
  }
  CK(cudaGetLastError());
  k_scan<<<1, 1024, 0, g_stream>>>(d_rowcnt, dia, neq);
  g_launches += 1;
  int64_t last = 0;
  CK(cudaMemcpyAsync(&last, dia + neq, sizeof(int64_t), cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  const int64_t nnz = last - 1;
  *nnz_out = nnz;
  if (nnz > nnzmx) {
    g_err = "*** jac_calc -- More storage needed for Jacobian. Storage exceeded. Increase lenpfac.";
    return -2;
  }
  const int64_t ncol = g_ivmax - g_ivmin + 1;
  if (ncol > 0) {
    k_fill<<<(unsigned)ncol, 64, 0, g_stream>>>(neq, g_ivmin, g_ivmax, d_coloff, d_colcnt, d_colrow, d_colval, dia, d_rowfill, djac, dja);
    k_sortrows<<<(unsigned)((neq + 127) / 128), 128, 0, g_stream>>>(neq, dia, djac, dja);
    g_launches += 2;
  }
  CK(cudaGetLastError());
  return 0;
}

}  // namespace

// ====================================================================================================
extern "C" {

int ue_gpu_set_int(const char* n, int64_t v) { if (S.set_int(n, v)) { g_err = std::string("unknown int input ") + n; return -1; } return 0; }
int ue_gpu_set_real(const char* n, double v) {
  if (S.set_real(n, v)) { g_err = std::string("unknown real input ") + n; return -1; }
  if (g_ready) {  // scalars such as nufak, dtreal may change between solves: patch the device copy in place
    const size_t off = (size_t)((char*)S.rscal[n] - (char*)&S.p);
    CK(cudaMemcpyToSymbol(D, &v, sizeof(double), off));
  }
  return 0;
}
int ue_gpu_set_real_array(const char* n, const double* d, int64_t k) { if (S.set_real_array(n, d, k)) { g_err = std::string("unknown real array ") + n; return -1; } return 0; }
int ue_gpu_set_int_array(const char* n, const int64_t* d, int64_t k) { if (S.set_int_array(n, d, k)) { g_err = std::string("unknown int array ") + n; return -1; } return 0; }
const char* ue_gpu_last_error(void) { return g_err.c_str(); }

int ue_gpu_init(void) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "no CUDA device: the B200 path has no CPU fallback"; return -9; }
  if (g_ready) free_all();
  std::string m = S.missing();
  if (!m.empty()) { g_err = "missing inputs: " + m; return -1; }
  const UeParams& P = S.p;
  nx = (int)P.nx; ny = (int)P.ny; NXS = nx + 2; NC = NXS * (ny + 2); neq = P.neq;
  std::string b = S.bad_sizes();
  if (!b.empty()) { g_err = "bad plane sizes: " + b; return -1; }
  if (neq != (int64_t)NC * UE_NV) { g_err = "neq != numvar*(nx+2)*(ny+2)"; return -1; }
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
  CK(cudaMalloc(&d_tmp, neq * sizeof(double)));
  CK(cudaMalloc(&d_yldot, neq * sizeof(double)));
  CK(cudaMalloc(&d_dtuse, neq * sizeof(double)));
  CK(cudaMalloc(&d_ylodt, neq * sizeof(double)));
  CK(cudaMalloc(&d_suscal, neq * sizeof(double)));
  CK(cudaMalloc(&d_sfscal, neq * sizeof(double)));
  CK(cudaMalloc(&d_err, sizeof(int)));
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
  CK(cudaMalloc(&d_coloff, neq * sizeof(int64_t)));
  CK(cudaMemcpy(d_coloff, h_coloff.data(), neq * sizeof(int64_t), cudaMemcpyHostToDevice));
  CK(cudaMalloc(&d_colcnt, neq * sizeof(int)));
  CK(cudaMalloc(&d_colrow, g_cap_total * sizeof(int)));
  CK(cudaMalloc(&d_colval, g_cap_total * sizeof(double)));
  CK(cudaMalloc(&d_rowcnt, neq * sizeof(int)));
  CK(cudaMalloc(&d_rowfill, neq * sizeof(int)));
  g_nnzcap = g_cap_total;
  CK(cudaMalloc(&d_ia, (neq + 1) * sizeof(int64_t)));
  CK(cudaMalloc(&d_ja, g_nnzcap * sizeof(int64_t)));
  CK(cudaMalloc(&d_jac, g_nnzcap * sizeof(double)));
  int dev = 0; cudaGetDevice(&dev);
  int maxsm = 0; cudaDeviceGetAttribute(&maxsm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
  if ((int64_t)g_smem_wide > maxsm || (int64_t)g_smem_narrow > maxsm) { g_err = "window box does not fit shared memory (mesh too wide for this build)"; return -6; }
  CK(cudaFuncSetAttribute(k_jac<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(g_smem_narrow, 1024)));
  CK(cudaFuncSetAttribute(k_jac<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(g_smem_wide, 1024)));
  g_launches = 0;
  g_ready = true;
  return 0;
}

int ue_gpu_step_params(int64_t n, const double* dt, const double* yo, const double* su, const double* sf) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "step_params: neq mismatch"; return -1; }
  CK(cudaMemcpyAsync(d_dtuse, dt, n * 8, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(d_ylodt, yo, n * 8, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(d_suscal, su, n * 8, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(d_sfscal, sf, n * 8, cudaMemcpyHostToDevice, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int ue_gpu_pandf1_dev(int64_t n, double time, const double* dyl, double* dyldot) {
  (void)time;
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "pandf1: neq mismatch"; return -1; }
  CK(cudaEventRecord(g_ev0, g_stream));
  int rc = run_residual_dev(dyl, dyldot, true);
  if (rc) return rc;
  CK(cudaEventRecord(g_ev1, g_stream));
  rc = check_errflag();
  if (rc) return rc;
  cudaEventElapsedTime(&g_res_ms, g_ev0, g_ev1);
  return 0;
}

int ue_gpu_pandf1(int64_t n, double time, const double* yl, double* yldot) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "pandf1: neq mismatch"; return -1; }
  CK(cudaMemcpyAsync(d_yl, yl, (neq + 2) * 8, cudaMemcpyHostToDevice, g_stream));
  int rc = ue_gpu_pandf1_dev(n, time, d_yl, d_yldot);
  if (rc) return rc;
  CK(cudaMemcpyAsync(yldot, d_yldot, neq * 8, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int ue_gpu_jac_calc_dev(int64_t n, double t, const double* dyl, const double* dy00, int64_t ml, int64_t mu, int64_t nnzmx, double* djac, int64_t* dja,
                        int64_t* dia, int64_t* nnz_out) {
  (void)t;
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "jac_calc: neq mismatch"; return -1; }
  CK(cudaEventRecord(g_ev0, g_stream));
  int rc = run_jac_dev(dyl, dy00, ml, mu, std::min(nnzmx, g_nnzcap), djac, dja, dia, nnz_out);
  if (rc) return rc;
  CK(cudaEventRecord(g_ev1, g_stream));
  rc = check_errflag();
  if (rc) return rc;
  cudaEventElapsedTime(&g_jac_ms, g_ev0, g_ev1);
  return 0;
}

int ue_gpu_jac_calc(int64_t n, double t, const double* yl, const double* yldot00, int64_t ml, int64_t mu, int64_t nnzmx, double* jac, int64_t* ja,
                    int64_t* ia, int64_t* nnz_out) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (n != neq) { g_err = "jac_calc: neq mismatch"; return -1; }
  CK(cudaMemcpyAsync(d_yl, yl, (neq + 2) * 8, cudaMemcpyHostToDevice, g_stream));
  CK(cudaMemcpyAsync(d_yldot00, yldot00, neq * 8, cudaMemcpyHostToDevice, g_stream));
  int64_t nnz = 0;
  int rc = ue_gpu_jac_calc_dev(n, t, d_yl, d_yldot00, ml, mu, nnzmx, d_jac, d_ja, d_ia, &nnz);
  *nnz_out = nnz;
  if (rc) return rc;
  CK(cudaMemcpyAsync(jac, d_jac, nnz * 8, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(ja, d_ja, nnz * 8, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaMemcpyAsync(ia, d_ia, (neq + 1) * 8, cudaMemcpyDeviceToHost, g_stream));
  CK(cudaStreamSynchronize(g_stream));
  return 0;
}

int ue_gpu_set_column_range(int64_t ivmin, int64_t ivmax) {
  if (!g_ready) { g_err = "ue_gpu_init not called"; return -1; }
  if (ivmin < 1 || ivmax > neq) { g_err = "column range outside 1..neq"; return -1; }
  g_ivmin = ivmin; g_ivmax = ivmax;
  const size_t s1 = g_smem_narrow, s2 = g_smem_wide;
  build_lists();
  g_smem_narrow = std::max(g_smem_narrow, s1); g_smem_wide = std::max(g_smem_wide, s2);
  return upload_lists();
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
int ue_gpu_finalize(void) { free_all(); return 0; }
}
