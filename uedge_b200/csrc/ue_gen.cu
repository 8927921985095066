// uedge_b200/csrc/ue_gen.cu — runtime of the GENERAL path (ue_gen_phys.h): C ABI ue_gen_* of include/ue_gen.h.
//
//   ue_gen_pandf1   : Pandf1rhs_interface (bbb/oderhs.m:8217-8254): one thread block evaluates the full domain in place
//                     on the base set of field planes (the reference's module state).
//   ue_gen_jac_calc : jac_calc_interface (bbb/oderhs.m:8533-8760): ONE WARP PER UNKNOWN.  Each warp owns a private copy
//                     of the field planes (HBM: NPL x NC doubles per unknown, 180 GB make the copies affordable), perturbs
//                     its unknown, runs the reference's windowed pandf1 (ranges i1..i8 x j1..j8 around the unknown's cell)
//                     cooperatively, differences the band against yldot00 and appends the kept elements, rows ascending,
//                     to its column fragment.  k_gen_count / k_gen_scan / k_gen_fill / k_gen_sortrows then transpose the
//                     fragments into the reference CSR (csrcsc, svr/svrut4.m:1536-1608: rows in order, columns ascending).
// The same file builds as plain C++ (-DUE_GEN_HOST, tests/hostcheck): contexts are then single threads and the loop nests
// run in the reference's order, or reversed with -DUE_GEN_REVERSE.  The host build exists for tests only; the product is
// the CUDA build and it fails loudly without a device.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "ue_gen_phys.h"
#include "ue_gen.h"

#if !defined(UE_GEN_HOST)
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#endif

namespace {

typedef std::vector<double> V;
std::map<std::string, V> IN;  // every named input (host copy)
std::string g_err;
std::string g_missing;
bool g_ready = false;

const char* kMessages[] = {"", "***  ni is negative - calculation stopped", "***  ng is negative - calculation stopped", "isnicore must be 0 or 1",
                           "recylb < -1 not built", "recyrb < -1 not built"};

// ---- memory: device (product) or host (hostcheck) ----------------------------------------------------------------------
#if defined(UE_GEN_HOST)
#define UE_PREFIX(x) ue_genh_##x
double* mem_alloc(size_t n) { return (double*)std::calloc(std::max<size_t>(n, 1), sizeof(double)); }
void mem_free(void* p) { std::free(p); }
bool mem_put(void* d, const void* h, size_t bytes) { std::memcpy(d, h, bytes); return true; }
bool mem_get(void* h, const void* d, size_t bytes) { std::memcpy(h, d, bytes); return true; }
#else
#define UE_PREFIX(x) ue_gen_##x
bool ck(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  g_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what;
  return false;
}
double* mem_alloc(size_t n) {
  void* p = nullptr;
  if (!ck(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(double)), "cudaMalloc")) return nullptr;
  cudaMemset(p, 0, std::max<size_t>(n, 1) * sizeof(double));
  return (double*)p;
}
void mem_free(void* p) { cudaFree(p); }
bool mem_put(void* d, const void* h, size_t bytes) { return ck(cudaMemcpy(d, h, bytes, cudaMemcpyHostToDevice), "cudaMemcpy H2D"); }
bool mem_get(void* h, const void* d, size_t bytes) { return ck(cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost), "cudaMemcpy D2H"); }
#endif

std::vector<void*> g_allocs;  // everything init allocated
std::map<std::string, double*> g_dev;  // uploaded inputs by name
Gen G;                  // constants + pointers as the evaluation contexts see them (device pointers in the CUDA build)
Gen* d_G = nullptr;     // the same object in context-visible memory
double* d_base = nullptr;   // base set of field planes (NPL x NC)
double* d_priv = nullptr;   // private sets, one per unknown of a chunk
size_t g_priv_cols = 0;
double *d_yl = nullptr, *d_yldot = nullptr, *d_y00 = nullptr, *d_ylp = nullptr, *d_wk = nullptr;
double* d_step = nullptr;   // dtuse | ylodt | suscal | sfscal
int *d_cnt = nullptr, *d_frow = nullptr, *d_err = nullptr;
double* d_fval = nullptr;
int64_t *d_ia = nullptr, *d_ja = nullptr, *d_ja2 = nullptr;
double *d_jac = nullptr, *d_jac2 = nullptr;
int64_t g_nnzmx = 0;
int NPL = 0;
int COLCAP = 512;
float g_comm_ms = 0.f;
float g_full_ms = 0.f, g_cols_ms = 0.f, g_csr_ms = 0.f;  // CUDA-event times of the last residual / column / CSR kernels
#define UE_GEN_BAND_DEFAULT 5
int g_band = UE_GEN_BAND_DEFAULT;  // rows of the private copy on each side of the perturbed cell (env UE_GEN_BAND; large = all rows)
int g_compact = 1;    // private planes hold the band's rows only (env UE_GEN_COMPACT=0: full-size planes)
size_t g_pad = 0;     // padding (doubles) in front of the first private slab
int g_colpad = 2;  // columns of the private copy on each side of the window i1..i6 (env UE_GEN_COLPAD; negative = whole rows)
int g_tpu = 32;  // threads per unknown in the Jacobian kernel: 32 (a warp) or 64 (a two-warp block)
int g_nslow = 0;
int g_occ1 = 0, g_occ4 = 0;  // resident blocks per SM of the two builds of the column kernel (0: not asked yet)
int g_sms = 0, g_full_grid = -1;  // SM count; blocks of the grid-mode residual (-1: by mesh size; env UE_GEN_FULL_GRID)
bool g_grid_ok = false;           // cooperative launch available and no thread can leave the evaluation on its own (see init)
int* d_gbar = nullptr;            // grid barrier ticket counter + error flags
int* d_order = nullptr;           // work list of the persistent column kernel (neq entries) + its queue counter
V g_last_yl;  // the state the base planes were last evaluated at
int64_t g_ivmin = 1, g_ivmax = 0;

#if !defined(UE_GEN_HOST)
// multi-GPU state (ue_gen_comm_init)
ncclComm_t gc_comm = nullptr;
int gc_nranks = 1, gc_rank = 0;
std::vector<int> gc_list_all, gc_list_off;   // unknowns (1-based) by rank, and the offsets of the rank segments (nranks + 1)
int *d_list_all = nullptr, *d_list_off = nullptr, *d_pk_cnt = nullptr, *d_pk_row = nullptr;
int64_t* d_pk_off = nullptr;
long long* d_pk_tot = nullptr;
double* d_pk_val = nullptr;
#endif

const V* find(const char* n) { auto it = IN.find(n); return it == IN.end() ? nullptr : &it->second; }
double SC(const char* n, int k = 0) {
  const V* v = find(n);
  if (!v || (int)v->size() <= k) { g_missing += std::string(n) + " "; return 0.; }
  return (*v)[k];
}
// context-visible copy of a named input array (uploaded once)
const double* ARR(const char* n, size_t need) {
  const V* v = find(n);
  if (!v || v->size() < need) { g_missing += std::string(n) + "[" + std::to_string(need) + "] "; return nullptr; }
  auto it = g_dev.find(n);
  if (it != g_dev.end()) return it->second;
  double* d = mem_alloc(v->size());
  if (!d) return nullptr;
  g_allocs.push_back(d);
  mem_put(d, v->data(), v->size() * sizeof(double));
  g_dev[n] = d;
  return d;
}
void VEC(const char* n, size_t need, double* dst) {
  const V* v = find(n);
  if (!v || v->size() < need) { g_missing += std::string(n) + "[" + std::to_string(need) + "] "; return; }
  for (size_t i = 0; i < need; ++i) dst[i] = (*v)[i];
}

void free_all() {
  for (void* p : g_allocs) mem_free(p);
  g_allocs.clear(); g_dev.clear();
  d_G = nullptr; d_base = d_priv = d_yl = d_yldot = d_y00 = d_ylp = d_wk = d_step = d_fval = d_jac = nullptr;
  d_cnt = d_frow = d_err = d_gbar = d_order = nullptr; d_ia = d_ja = d_ja2 = nullptr; d_jac2 = nullptr;
  g_priv_cols = 0; g_nnzmx = 0; g_ready = false; g_last_yl.clear();
#if !defined(UE_GEN_HOST)
  gc_nranks = 1; gc_rank = 0;  // (the lists belonged to the case that was just freed; the communicator itself lives until comm_finalize)
  d_list_all = d_list_off = d_pk_cnt = d_pk_row = nullptr; d_pk_off = nullptr; d_pk_tot = nullptr; d_pk_val = nullptr;
#endif
}
template <typename T> T* alloc_as(size_t n) {
  double* p = mem_alloc((n * sizeof(T) + sizeof(double) - 1) / sizeof(double));
  if (p) g_allocs.push_back(p);
  return (T*)p;
}

// rows copied on each side of the perturbed cell's row, and the number of line arrays at the end of UE_GEN_PLANES
#define UE_GEN_NLINE 14
#if defined(UE_GEN_HOST)
int g_poison = 0;  // test aid: before the copy, fill the private planes with NaN (1: a missing cell shows up as a lost entry) or with finite
                   // garbage that changes from unknown to unknown (2: a stale cell that reaches a kept row shows up as a spurious entry)
#endif
// ---- the two evaluation bodies (shared by the kernels and the host build) ----------------------------------------------
// full-domain residual in place on the context's planes
HD int eval_full(Gen& g, const double* yl, double* yldot) { return g.pandf1(-1, -1, yl, yldot); }

// one Jacobian column (oderhs.m:8600-8720).  g: private context (planes = private copy of the base set).
// ylp: private copy of yl (neq+2); wk: private residual (neq); frow/fval: the column's fragment, capacity cap.
HD int eval_column(Gen& g, const double* base, int npl, int64_t iv, const double* yl, double* ylp, double* wk, const double* yldot00, int64_t ml, int64_t mu,
                   int cap, int* frow, double* fval, int* cnt, int band, int colpad, double* priv, int compact) {
  // priv: this context's private slab - full-size planes (compact = 0: NPL x NC doubles, plane pointers already assigned) or band-rows-only
  // planes (compact = 1: (NPL - NLINE) x (2 band + 1) rows x NXS + NLINE line arrays; the plane pointers are set here for every unknown)
  const int64_t neq = g.neq;
  const int tid = g.TID();
  const int xc = (int)g.igyld[iv - 1], yc = (int)g.igyld[neq + iv - 1];
  g.sync();
  // Private copy of the base planes: only what the windowed evaluation can touch -
  //  * the band of rows yc +- band (ranges j1p-1 .. j6p+2 of oderhs.m:868-964), and in them the columns of the window
  //    i1-2 .. i6+2; windows that reach an X-point cut span all ix in the reference (xccuts, oderhs.m:960-1019) and take whole rows;
  //  * in rows 0-2, when the band holds them, all core columns: the core conditions sum fluxes and currents over the entire core
  //    boundary and, with isnewpot = 1, set the potential rows at every core column (boundary.m:229-240, 492-502, 987-1122); the
  //    corner cells of both walls (every window at a wall sets their density rows, boundary.m:246-262);
  //  * the eight cells of the X-point vertex average (convert.m:831-868, evaluated by every window);
  //  * the line arrays in full.
  // Everything else keeps whatever an earlier unknown left there: nothing in the band's result reads it
  // (tests/test_gen_hostcheck.py poisons it with NaN to show that).
  {
    const int NXS = g.NXS, NC = g.NC, nrow = g.ny + 2;
    const int r0 = mx(0, yc - band), r1 = mn(nrow - 1, yc + band);
    const int nfield = npl - UE_GEN_NLINE;
    const int brows = 2 * band + 1, nline = mx(NXS, nrow);
    const size_t bpl = (size_t)brows * NXS;  // a band plane
    g.sync();
    if (tid == 0) {
      if (compact) { g.assign_planes_band(priv, r0, brows, UE_GEN_NLINE, nline); g.rowlo = r0; g.rowhi = r1; }
      else { g.rowlo = 0; g.rowhi = nrow - 1; }
    }
    g.sync();
#if defined(UE_GEN_HOST)
    if (g_poison == 2) {  // positive, finite, different for every cell and every unknown
      const size_t n = compact ? (size_t)nfield * bpl + (size_t)UE_GEN_NLINE * nline : (size_t)npl * NC;
      for (size_t k = 0; k < n; ++k) priv[k] = 1.0e3 * (1.0 + 0.37 * (double)((k * 2654435761ull + (size_t)iv * 40503ull) % 1000003ull) / 1000003.0);
    } else if (g_poison && compact) { for (size_t k = 0; k < (size_t)nfield * bpl + (size_t)UE_GEN_NLINE * nline; ++k) priv[k] = (double)NAN; }
    else
    if (g_poison) {
      if (getenv("UE_GEN_OOB")) {  // every cell its own NaN payload: a later bit-compare finds any write, copies of poison included
        for (size_t k = 0; k < (size_t)npl * NC; ++k) { const uint64_t bits = 0x7ff8000000000000ull | (uint64_t)(k + 1); std::memcpy(&priv[k], &bits, 8); }
      } else for (size_t k = 0; k < (size_t)npl * NC; ++k) priv[k] = (double)NAN;
    }
#endif
    const auto w = g.make_win(xc, yc);
    const bool fullx = w.xccuts || colpad < 0 || (w.i1 <= colpad && w.i6 >= g.nx + 1 - colpad);
    // the copy is a union of rectangles (rows a..b) x (columns c..d); overlaps are copied twice (same values)
    int ra[9], rb[9], ca[9], cb[9], nr = 0;
    auto rect = [&](int a, int b_, int c, int d) {
      a = mx(a, 0); b_ = mn(b_, nrow - 1); c = mx(c, 0); d = mn(d, NXS - 1);
      if (a <= b_ && c <= d) { ra[nr] = a; rb[nr] = b_; ca[nr] = c; cb[nr] = d; ++nr; }
    };
    if (fullx) rect(r0, r1, 0, NXS - 1);
    else {
      rect(r0, r1, w.i1 - colpad, w.i6 + colpad);                            // the window
      if (r0 <= 2) {
        rect(r0, mn(2, r1), mn(w.i1 - colpad, g.ixpt1), mx(w.i6 + colpad, g.ixpt2 + 1));  // core boundary: sums over all core columns; with
                                                                                   // isnewpot = 1 the potential rows are set from the window to the far end of the core
        rect(r0, mn(2, r1), 0, 1); rect(r0, mn(2, r1), g.nx, g.nx + 1);       // corner cells of the inner wall
      }
      if (r1 >= nrow - 2) { rect(mx(r0, nrow - 2), r1, 0, 1); rect(mx(r0, nrow - 2), r1, g.nx, g.nx + 1); }  // ... and of the outer wall
    }
    if (g.iysptrx1 >= 0 && g.ixpt1 >= 0 && g.ixpt2 >= 0 && !(fullx && g.iysptrx1 >= r0 && g.iysptrx1 + 1 <= r1) &&
        (!compact || (g.iysptrx1 >= r0 && g.iysptrx1 + 1 <= r1))) {  // X-point vertex: 8 cells (band-rows-only planes: evaluated only inside the band)
      rect(g.iysptrx1, g.iysptrx1 + 1, g.ixpt1 - 1, g.ixpt1 + 2);
      rect(g.iysptrx1, g.iysptrx1 + 1, g.ixpt2 - 1, g.ixpt2 + 2);
    }
    for (int q = 0; q < nr; ++q) {  // one flat loop over (plane, row, column) of a rectangle: independent loads, several in flight per thread
      const int wc = cb[q] - ca[q] + 1, rw = (rb[q] - ra[q] + 1) * wc;
      const int tot = nfield * rw;
      const bool small = tot < (1 << 22);  // (the float quotients are exact below 2^22)
#pragma unroll 4
      for (int k = tid; k < tot; k += g.nth) {
        const int p = small ? UE_ROW_(k, rw) : k / rw;
        const int e = k - p * rw;
        const int r = small ? UE_ROW_(e, wc) : e / wc;
        const size_t o = (size_t)p * NC + (size_t)(ra[q] + r) * NXS + (size_t)(ca[q] + e - r * wc);
        if (compact) priv[(size_t)p * bpl + (size_t)(ra[q] + r - r0) * NXS + (size_t)(ca[q] + e - r * wc)] = base[o];
        else priv[o] = base[o];
      }
    }
    for (int p = nfield; p < npl; ++p) {
      const size_t o = (size_t)p * NC, oc = (size_t)nfield * bpl + (size_t)(p - nfield) * nline;
      for (int k = tid; k < nline; k += g.nth) priv[compact ? oc + k : o + k] = base[o + k];
    }
  }
  // private state vector and residual: the unknowns of the band's rows (the windowed evaluation converts the perturbed cell, rescales
  // and writes inside its window and at the boundary rows of the band) and whatever the differencing range ii1..ii2 reaches; the whole
  // vectors when the time-step term of pandf1 is active (it runs over every cell) or the window is the whole mesh
  int64_t ii1 = mx(iv - mu, (int64_t)1), ii2 = mn(iv + ml, neq);
  if (g.ExtendedJacPhi > 0 && g.isphion * g.isnewpot == 1 && iv % g.numvar_ == 0) {  // wider band for a potential perturbation (oderhs.m:8645-8651)
    ii1 = mx(iv - 4 * g.numvar_ * g.nx, (int64_t)1); ii2 = mn(iv + 4 * g.numvar_ * g.nx, neq);
  }
  {
    const int nrow = g.ny + 2;
    const int r0 = mx(0, yc - band), r1 = mn(nrow - 1, yc + band);
    const bool whole = (g.dtreal < 1.e15 && yl[neq] < 0) || g.yinc >= 6 || colpad < 0;
#if defined(UE_GEN_HOST)
    if (g_poison) for (int64_t k = 0; k < neq; ++k) { ylp[k] = g_poison == 2 ? 0.731 + 1e-3 * (double)(k % 977) : (double)NAN; wk[k] = g_poison == 2 ? 1.0e5 + (double)(k % 991) : (double)NAN; }
#endif
    const auto w = g.make_win(xc, yc);
    const bool fullx = w.xccuts || g.rowuniform_ == 0 || (w.i1 <= colpad && w.i6 >= g.nx + 1 - colpad);
    if (whole || fullx) {
      int64_t lo = 0, hi = neq;
      if (!whole) { lo = mn((int64_t)g.rowiv_[r0], ii1 - 1); hi = mx((int64_t)g.rowiv_[r1 + 1], ii2); }
      for (int64_t k = lo + tid; k < hi; k += g.nth) { ylp[k] = yl[k]; wk[k] = yldot00[k]; }
    } else {  // the differencing range, and in the band's rows the unknowns of the window's columns (every cell holds numvar unknowns)
      for (int64_t k = ii1 - 1 + tid; k < ii2; k += g.nth) { ylp[k] = yl[k]; wk[k] = yldot00[k]; }
      const int c0 = mx(0, w.i1 - colpad), c1 = mn(g.NXS - 1, w.i6 + colpad);
      const int per = (int)g.numvar_ * (c1 - c0 + 1), tot = per * (r1 - r0 + 1);
      for (int e = tid; e < tot; e += g.nth) {
        const int r = e / per;
        const int64_t k = (int64_t)g.rowiv_[r0 + r] + g.numvar_ * c0 + (e - r * per);
        ylp[k] = yl[k]; wk[k] = yldot00[k];
      }
    }
    if (tid == 0) { ylp[neq] = yl[neq]; ylp[neq + 1] = yl[neq + 1]; }
  }
  g.sync();
  const double yold = yl[iv - 1];
  const double dyl = g.delpert * (fabs(yold) + g.dylconst / g.suscal[iv - 1]);
  if (tid == 0) ylp[iv - 1] = yold + dyl;
  g.sync();
  const int rc = g.pandf1(xc, yc, ylp, wk);
#if defined(UE_GEN_HOST)
  if (g_poison && !compact && getenv("UE_GEN_OOB")) {  // developer aid: which planes does a windowed evaluation write outside the band's rows?
    static std::map<std::pair<int, int>, int> seen;  // (plane, row offset class) -> count
    const int NXS = g.NXS, NC = g.NC, nrow = g.ny + 2;
    const int r0 = mx(0, yc - band), r1 = mn(nrow - 1, yc + band);
    for (int p = 0; p < npl - UE_GEN_NLINE; ++p)
      for (int r = 0; r < nrow; ++r) {
        if (r >= r0 && r <= r1) continue;
        if (g.iysptrx1 >= 0 && (r == g.iysptrx1 || r == g.iysptrx1 + 1)) continue;
        bool w = false;
        for (int c = 0; c < NXS; ++c) {
          const size_t k = (size_t)p * NC + (size_t)r * NXS + c;
          uint64_t bits; std::memcpy(&bits, &priv[k], 8);
          if (bits != (0x7ff8000000000000ull | (uint64_t)(k + 1))) {
            // cells the copy itself filled (rectangles outside the band's rows: none, by construction of this scan) are excluded above
            w = true;
          }
        }
        if (w) { const int cls = r <= 2 ? r : (r >= nrow - 2 ? 100 + (r - (nrow - 2)) : 50); if (seen[{p, cls}]++ == 0) fprintf(stderr, "OOB write: plane %d row class %d (row %d, band %d..%d)\n", p, cls, r, r0, r1); }
      }
  }
#endif
  g.sync();
  if (rc) { if (tid == 0) *cnt = 0; return rc; }
  const bool isphi = g.IDXPHI(xc, yc) == iv - 1;
  // difference, diagonal terms, clip; NaN marks an element that is not kept (a NaN element fails the clip test as well)
  for (int64_t ii = ii1 + tid; ii <= ii2; ii += g.nth) {
    double jacelem = (wk[ii - 1] - yldot00[ii - 1]) / dyl;
    if (iv == ii) {
      if (g.ALG(iv - 1) * (1 - g.isbcwdt) == 0) jacelem = jacelem - 1 / g.dtuse[iv - 1];
      if (isphi) jacelem = jacelem - 1 / g.dtphi;  // oderhs.m:8694-8700
    }
    if (g.nufak > 0) if (iv == ii && yl[neq] == 1) jacelem = jacelem - g.nufak;
    wk[ii - 1] = (fabs(jacelem * g.sfscal[iv - 1]) > g.jaccliplim) ? jacelem : (double)NAN;
  }
  g.sync();
  if (tid == 0) {  // ordered append (rows ascending)
    int n = 0;
    for (int64_t ii = ii1; ii <= ii2; ++ii) {
      const double v = wk[ii - 1];
      if (v == v) { if (n < cap) { frow[n] = (int)ii; fval[n] = v; } ++n; }
    }
    *cnt = n;
  }
  return 0;
}

#if !defined(UE_GEN_HOST)
// ---- kernels ---------------------------------------------------------------------------------------------------------
extern __shared__ unsigned char g_smem[];
__device__ void load_ctx(Gen* dst, const Gen* src, int tid, int nth) {  // word-wise copy of the constant part into shared memory
  const int nw = (int)(sizeof(Gen) / sizeof(int));
  for (int k = tid; k < nw; k += nth) ((int*)dst)[k] = ((const int*)src)[k];
}
__global__ void k_gen_full(const Gen* gsrc, double* base, const double* yl, double* yldot, int* err) {
  Gen* g = (Gen*)g_smem;
  load_ctx(g, gsrc, threadIdx.x, blockDim.x);
  __syncthreads();
  if (threadIdx.x == 0) { g->nth = blockDim.x; g->gridmode = 0; g->errc = 0; g->assign_planes(base); }
  __syncthreads();
  const int rc = eval_full(*g, yl, yldot);
  if (rc && threadIdx.x == 0) { err[0] = rc; err[1] = g->errc; }
}
// the same over a co-resident grid (cooperative launch): one context = all threads, barriers between the nests are grid barriers
__global__ void __launch_bounds__(128) k_gen_full_grid(const Gen* gsrc, double* base, const double* yl, double* yldot, int* err, unsigned* gbar, int* gflag) {
  Gen* g = (Gen*)g_smem;
  load_ctx(g, gsrc, threadIdx.x, blockDim.x);
  __syncthreads();
  if (threadIdx.x == 0) { g->nth = (int)(gridDim.x * blockDim.x); g->gridmode = 1; g->gbar = gbar; g->gflag = gflag; g->errc = 0; g->assign_planes(base); }
  __syncthreads();
  const int rc = eval_full(*g, yl, yldot);
  if (rc && blockIdx.x == 0 && threadIdx.x == 0) { err[0] = rc; err[1] = g->errc; }
}
// one warp per unknown of the chunk [iv0, iv0 + ncol): every warp has its own context (shared memory) and its own planes
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_gen_cols(const Gen* gsrc, const double* base, double* priv, int npl, int64_t iv0, const int* ivlist, int ncol, const double* yl, double* ylp, double* wk,
                           const double* yldot00, int64_t ml, int64_t mu, int cap, int* frow, double* fval, int* cnt, int* err, int tpu, int band, int colpad) {
  // a unit = the threads that evaluate one unknown: a warp (4 units per block) or, with tpu > 32, the whole block
  const int unit = tpu > 32 ? 0 : (int)(threadIdx.x >> 5), lane = tpu > 32 ? (int)threadIdx.x : (int)(threadIdx.x & 31);
  Gen* g = (Gen*)g_smem + unit;
  load_ctx(g, gsrc, lane, tpu);
  if (tpu > 32) __syncthreads(); else __syncwarp();
  const int c = tpu > 32 ? (int)blockIdx.x : (int)(blockIdx.x * (blockDim.x >> 5) + unit);
  if (c >= ncol) return;
  const size_t nslab = (size_t)npl * g->NC;
  if (lane == 0) { g->nth = tpu; g->gridmode = 0; g->errc = 0; g->assign_planes(priv + (size_t)c * nslab); }
  if (tpu > 32) __syncthreads(); else __syncwarp();
  const int64_t iv = ivlist ? (int64_t)ivlist[c] : iv0 + c;  // (multi-GPU: this rank's unknowns are a list of mesh rows)
  const int64_t neq = g->neq;
  const int rc = eval_column(*g, base, npl, iv, yl, ylp + (size_t)c * (neq + 2), wk + (size_t)c * neq, yldot00, ml, mu, cap, frow + (size_t)(iv - 1) * cap,
                             fval + (size_t)(iv - 1) * cap, cnt + (iv - 1), band, colpad, priv + (size_t)c * nslab, 0);
  if (tpu > 32) __syncthreads(); else __syncwarp();
  if (rc && lane == 0) { err[0] = rc; err[1] = g->errc; }
  if (lane == 0 && cnt[iv - 1] > cap) err[2] = cnt[iv - 1];
}
// CSC fragments -> CSR: entries per row, exclusive scan, ordered fill (one thread per ROW walks the columns that can reach it)
__global__ void k_gen_count(int64_t neq, int cap, const int* cnt, const int* frow, int* rowcnt) {
  const int64_t iv = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (iv >= neq) return;
  const int n = min(cnt[iv], cap);
  for (int k = 0; k < n; ++k) atomicAdd(&rowcnt[frow[iv * cap + k] - 1], 1);
}
__global__ void k_gen_scan(int64_t neq, const int* rowcnt, int64_t* ia, int* cursor) {  // one block
  __shared__ long long carry;
  __shared__ long long part[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < neq; b0 += blockDim.x) {
    const int64_t i = b0 + threadIdx.x;
    const long long v = i < neq ? rowcnt[i] : 0;
    part[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < (int)blockDim.x; off <<= 1) {
      const long long t = threadIdx.x >= off ? part[threadIdx.x - off] : 0;
      __syncthreads();
      part[threadIdx.x] += t;
      __syncthreads();
    }
    if (i < neq) { ia[i] = carry + part[threadIdx.x] - v + 1; cursor[i] = 0; }
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += part[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) ia[neq] = carry + 1;
}
// scatter by row (cursor atomics), then every row is put in ascending column order by one thread (rows hold <= ~100 entries)
__global__ void k_gen_fill(int64_t neq, int cap, const int* cnt, const int* frow, const double* fval, const int64_t* ia, int* cursor, int64_t nnzmx, double* jac, int64_t* ja) {
  const int64_t iv = blockIdx.x + 1;
  const int n = min(cnt[iv - 1], cap);
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    const int r = frow[(iv - 1) * cap + k];
    const int64_t pos = ia[r - 1] - 1 + atomicAdd(&cursor[r - 1], 1);
    if (pos < nnzmx) { jac[pos] = fval[(iv - 1) * cap + k]; ja[pos] = iv; }
  }
}
// one warp per row: rank of every entry among the row's column numbers (they are distinct), written to the second buffer
__global__ void k_gen_sortrows(int64_t neq, const int64_t* ia, int64_t nnzmx, const double* jac_in, const int64_t* ja_in, double* jac, int64_t* ja) {
  const int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= neq) return;
  const int64_t a = ia[r] - 1, b = min(ia[r + 1] - 1, nnzmx);
  for (int64_t i = a + lane; i < b; i += 32) {
    const int64_t cj = ja_in[i];
    int rank = 0;
    for (int64_t j = a; j < b; ++j) rank += (ja_in[j] < cj);
    ja[a + rank] = cj; jac[a + rank] = jac_in[i];
  }
}
#endif

#if !defined(UE_GEN_HOST)
// Persistent form of the column kernel: as many warps as are resident on the device (or unknowns, if fewer), each with ONE private
// plane set that it reuses; the warps draw unknowns from a work queue (atomic counter) in the order of `order` - the host puts the
// unknowns whose window spans all ix (windows at an X-point cut, oderhs.m:960-1019: 10-30 x the work of the others) first, so that
// they start at once and the short ones fill in behind them.  No chunking by memory, no tail of long windows at the end of a launch.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_gen_cols_q(const Gen* gsrc, const double* base, double* priv, int npl, const int* order, int ncol, int* queue, const double* yl, double* ylp,
                             double* wk, const double* yldot00, int64_t ml, int64_t mu, int cap, int* frow, double* fval, int* cnt, int* err, int band, int colpad, size_t slab, int compact) {
  const int unit = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
  Gen* g = (Gen*)g_smem + unit;
  load_ctx(g, gsrc, lane, 32);
  __syncwarp();
  const int slot = (int)(blockIdx.x * (blockDim.x >> 5) + unit);
  if (lane == 0) { g->nth = 32; g->gridmode = 0; g->errc = 0; if (!compact) g->assign_planes(priv + (size_t)slot * slab); }
  __syncwarp();
  const int64_t neq = g->neq;
  // a block draws its warps' unknowns together: neighbours in the list are the unknowns of one cell - the same window, so the warps
  // read the same part of the base planes (L1 hits) and take equally long
  __shared__ int s_first;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_first = (*(volatile int*)err) ? ncol : atomicAdd(queue, (int)(blockDim.x >> 5));
    __syncthreads();
    const int first = s_first;
    if (first >= ncol) break;
    const int c = first + unit;
    if (c >= ncol) continue;
    const int64_t iv = (int64_t)order[c];
    const int rc = eval_column(*g, base, npl, iv, yl, ylp + (size_t)slot * (neq + 2), wk + (size_t)slot * neq, yldot00, ml, mu, cap, frow + (size_t)(iv - 1) * cap,
                               fval + (size_t)(iv - 1) * cap, cnt + (iv - 1), band, colpad, priv + (size_t)slot * slab, compact);
    __syncwarp();
    if (rc && lane == 0) { err[0] = rc; err[1] = g->errc; }
    if (lane == 0 && cnt[iv - 1] > cap) err[2] = cnt[iv - 1];
  }
}
// ---- packing of the column fragments for the exchange (multi-GPU) ---------------------------------------------------------
// segment = the unknowns of one rank in list order.  k_gen_segscan: block s scans the counts of segment s (exclusive) and leaves
// the segment total; own = 1: the counts are first gathered from cnt[iv-1] (this rank's fresh results) into pk_cnt.
__global__ void k_gen_segscan(const int* list_all, const int* list_off, int seg0, const int* cnt, int* pk_cnt, int64_t* pk_off, long long* tot, int own) {
  const int s = seg0 + blockIdx.x;
  const int a = list_off[s], b = list_off[s + 1];
  __shared__ long long carry;
  __shared__ long long part[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int b0 = a; b0 < b; b0 += blockDim.x) {
    const int j = b0 + threadIdx.x;
    long long v = 0;
    if (j < b) { if (own) pk_cnt[j] = cnt[list_all[j] - 1]; v = pk_cnt[j]; }
    part[threadIdx.x] = v;
    __syncthreads();
    for (int off = 1; off < (int)blockDim.x; off <<= 1) {
      const long long t = (int)threadIdx.x >= off ? part[threadIdx.x - off] : 0;
      __syncthreads();
      part[threadIdx.x] += t;
      __syncthreads();
    }
    if (j < b) pk_off[j] = carry + part[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry += part[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) tot[s] = carry;
}
// dir = 0: strided fragments of this rank's columns -> its packed segment; dir = 1: packed segments of all ranks -> strided fragments
__global__ void k_gen_packcopy(const int* list_all, const int* list_off, int nseg, int j0, int j1, int cap, int* cnt, int* frow, double* fval, const int* pk_cnt,
                               const int64_t* pk_off, int* pk_row, double* pk_val, int dir, int skip_seg) {
  const int j = j0 + blockIdx.x;
  if (j >= j1) return;
  int s = 0;
  while (s + 1 < nseg && list_off[s + 1] <= j) ++s;
  if (dir == 1 && s == skip_seg) return;  // (own columns are already in place)
  const int64_t iv = list_all[j];
  const int n = min(pk_cnt[j], cap);
  const size_t p = (size_t)list_off[s] * cap + (size_t)pk_off[j], q = (size_t)(iv - 1) * cap;
  if (dir == 0) { for (int k = threadIdx.x; k < n; k += blockDim.x) { pk_row[p + k] = frow[q + k]; pk_val[p + k] = fval[q + k]; } }
  else {
    for (int k = threadIdx.x; k < n; k += blockDim.x) { frow[q + k] = pk_row[p + k]; fval[q + k] = pk_val[p + k]; }
    if (threadIdx.x == 0) cnt[iv - 1] = pk_cnt[j];
  }
}
// ---- multi-GPU: the columns of ONE Jacobian split over the ranks (ppp jac_calc_mpi, ppp/mpi_parallel.F90:2-447).  The column
// kernel is throughput-bound once the mesh is large (one warp and one private plane set per unknown), so the split scales; the
// fragments (count, rows, values per column) are exchanged in place with grouped ncclBroadcast calls and every rank builds the CSR.
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
} NCG;
int nccl_bind() {
  if (NCG.h) return 0;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // the copy the host process already loaded (e.g. torch's)
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { g_err = std::string("cannot load libnccl.so.2: ") + dlerror(); return -11; }
#define B(f) *(void**)(&NCG.f) = dlsym(h, "nccl" #f); if (!NCG.f) { g_err = "libnccl.so.2 lacks nccl" #f; return -11; }
  B(CommInitRank) B(CommDestroy) B(GroupStart) B(GroupEnd) B(Broadcast) B(AllGather) B(GetErrorString)
#undef B
  NCG.h = h;
  return 0;
}
#define NCK(call)                                                                                                 \
  do {                                                                                                            \
    ncclResult_t r_ = (call);                                                                                     \
    if (r_ != ncclSuccess) { g_err = std::string("NCCL error: ") + NCG.GetErrorString(r_) + " at " #call; return -11; } \
  } while (0)
// Ownership: mesh ROWS are dealt out cyclically (row iy belongs to rank iy mod nranks).  The windows of the rows near the X-point
// span all ix (xccuts) and cost several times more than the others; a contiguous split would leave them all on the first ranks.
int build_rank_lists(int nranks) {
  const V* ig = find("igyl");
  const size_t neq = (size_t)G.neq;
  gc_list_all.clear(); gc_list_off.assign(nranks + 1, 0);
  for (int r = 0; r < nranks; ++r) {
    gc_list_off[r] = (int)gc_list_all.size();
    for (size_t iv = 1; iv <= neq; ++iv) if (((int)(*ig)[neq + iv - 1]) % nranks == r) gc_list_all.push_back((int)iv);
  }
  gc_list_off[nranks] = (int)gc_list_all.size();
  d_list_all = alloc_as<int>(neq); d_list_off = alloc_as<int>(nranks + 1); d_pk_cnt = alloc_as<int>(neq); d_pk_off = alloc_as<int64_t>(neq);
  d_pk_tot = alloc_as<long long>(2 * nranks + 2);
  if (!d_list_all || !d_list_off || !d_pk_cnt || !d_pk_off || !d_pk_tot) return -10;
  if (!mem_put(d_list_all, gc_list_all.data(), neq * sizeof(int)) || !mem_put(d_list_off, gc_list_off.data(), (nranks + 1) * sizeof(int))) return -10;
  return 0;
}
// after this rank's columns are in the strided fragment arrays: pack them, exchange the packed segments, unpack the others'
int exchange_fragments(int cap) {
  const int N = gc_nranks, me = gc_rank;
  const size_t neq = (size_t)G.neq;
  if (!d_pk_row) {
    d_pk_row = alloc_as<int>(neq * cap); d_pk_val = mem_alloc(neq * cap);
    if (!d_pk_row || !d_pk_val) return -10;
    g_allocs.push_back(d_pk_val);
  }
  const int a = gc_list_off[me], b = gc_list_off[me + 1];
  k_gen_segscan<<<1, 1024>>>(d_list_all, d_list_off, me, d_cnt, d_pk_cnt, d_pk_off, d_pk_tot, 1);
  if (b > a) k_gen_packcopy<<<b - a, 64>>>(d_list_all, d_list_off, N, a, b, cap, d_cnt, d_frow, d_fval, d_pk_cnt, d_pk_off, d_pk_row, d_pk_val, 0, -1);
  // every rank needs every segment's size on the host (the broadcast counts): all-gather of one number per rank
  NCK(NCG.AllGather(d_pk_tot + me, d_pk_tot + N + 1, 1, ncclInt64, gc_comm, 0));
  std::vector<long long> tot(N);
  if (!ck(cudaMemcpy(tot.data(), d_pk_tot + N + 1, N * sizeof(long long), cudaMemcpyDeviceToHost), "segment sizes")) return -10;
  NCK(NCG.GroupStart());
  for (int r = 0; r < N; ++r) {
    const size_t o = (size_t)gc_list_off[r], n = (size_t)(gc_list_off[r + 1] - gc_list_off[r]);
    if (n == 0) continue;
    NCK(NCG.Broadcast(d_pk_cnt + o, d_pk_cnt + o, n, ncclInt32, r, gc_comm, 0));
    if (tot[r] > 0) {
      NCK(NCG.Broadcast(d_pk_row + o * cap, d_pk_row + o * cap, (size_t)tot[r], ncclInt32, r, gc_comm, 0));
      NCK(NCG.Broadcast(d_pk_val + o * cap, d_pk_val + o * cap, (size_t)tot[r], ncclFloat64, r, gc_comm, 0));
    }
  }
  NCK(NCG.GroupEnd());
  k_gen_segscan<<<N, 1024>>>(d_list_all, d_list_off, 0, d_cnt, d_pk_cnt, d_pk_off, d_pk_tot, 0);
  k_gen_packcopy<<<(unsigned)neq, 64>>>(d_list_all, d_list_off, N, 0, (int)neq, cap, d_cnt, d_frow, d_fval, d_pk_cnt, d_pk_off, d_pk_row, d_pk_val, 1, me);
  return 0;
}
#endif

int report(int rc, int errc) {
  if (errc > 0 && errc < (int)(sizeof(kMessages) / sizeof(kMessages[0]))) g_err = kMessages[errc];
  else g_err = "pandf1 failed";
  return rc;
}

int init_all() {
  g_missing.clear();
  free_all();
  Gen& g = G;
  std::memset((void*)&g, 0, sizeof(Gen));
  auto I = [&](const char* n, int k = 0) { return (int)std::llround(SC(n, k)); };
  g.nx = I("nx"); g.ny = I("ny"); g.NXS = g.nx + 2; g.NC = g.NXS * (g.ny + 2);
  g.nisp = I("nisp"); g.nusp = I("nusp"); g.ngsp = I("ngsp"); g.nhsp = I("nhsp"); g.neq = (int64_t)std::llround(SC("neq"));
  if (!g_missing.empty()) { g_err = "missing inputs: " + g_missing; return -1; }
  if (g.NC >= (1 << 20)) { g_err = "mesh too large for the cooperative loop index arithmetic (2^20 cells)"; return -5; }
  if (g.nisp < 1 || g.nisp > 2 || g.nusp > g.nisp || g.ngsp != 1) { g_err = "nisp must be 1 or 2 (hydrogen ions + inertial atoms), ngsp 1"; return -5; }
  g.nfsp = g.nisp;
  g.ixpt1 = I("ixpt1"); g.ixpt2 = I("ixpt2"); g.iysptrx1 = I("iysptrx1"); g.iysptrx2 = I("iysptrx2"); g.iysptrx = I("iysptrx"); g.ixlb = I("ixlb"); g.ixrb = I("ixrb"); g.ixmp = I("ixmp");
  g.xlinc = I("xlinc"); g.xrinc = I("xrinc"); g.yinc = I("yinc"); g.isjaccorall = I("isjaccorall");
  g.methn = I("methn"); g.methu = I("methu"); g.methe = I("methe"); g.methi = I("methi"); g.methg = I("methg");
#define GI(n) g.n = I(#n);
  GI(isnonog) GI(isphion) GI(isphiofft) GI(ineudif) GI(isflxvar) GI(isrscalf) GI(isbcwdt) GI(icnuiz) GI(icnucx) GI(isrecmon) GI(ingb) GI(inflbg) GI(isgasdc) GI(isdifxg_aug) GI(isdifyg_aug)
  GI(isvylog) GI(isgxvon) GI(convis) GI(concap) GI(isflxlde) GI(isflxldi) GI(isplflxl) GI(inkxc) GI(isgpye) GI(ishavisy) GI(isvhyha) GI(islnlamcon) GI(isnupdot1sd) GI(iteb) GI(istabon)
  GI(ifxnsgi) GI(iflcore) GI(ifluxni) GI(isrefluxclip) GI(ibctepl) GI(ibctipl) GI(ibctepr) GI(ibctipr) GI(isbohmms) GI(isextrnp) GI(isextrnpf) GI(isextrtpf) GI(isextrngc) GI(isextrnw)
  GI(isextrtw) GI(isnfmiy) GI(isybdrywd) GI(isnewpot) GI(jhswitch) GI(isfeexpl0) GI(isfeixpl0) GI(isintlog) GI(iskaplex) GI(isexunif) GI(isfdiax) GI(isugfm1side) GI(isvisxn_old) GI(isteon) GI(istion) GI(iphibcc)
#undef GI
  g.isupgon = I("isupgon", 0); g.isngon = I("isngon", 0); g.istgon = I("istgon", 0); g.isfixlb = I("isfixlb", 0); g.isfixrb = I("isfixrb", 0); g.newbcl = I("newbcl", 0); g.newbcr = I("newbcr", 0);
  g.isngcore1 = I("isngcore", 0);
  for (int f = 0; f < 2; ++f) { g.isnicore[f] = I("isnicore", f); g.isupcore[f] = I("isupcore", f); g.isupss[f] = I("isupss", f); g.isnion[f] = I("isnion", f); g.isupon[f] = I("isupon", f); }
  g.ev = SC("ev"); g.qe = SC("qe"); g.me = SC("me"); g.mp = SC("mp"); g.pi_ = SC("pi"); g.cutlo = SC("cutlo"); g.rt8opi = SC("rt8opi");
#define GR(n) g.n = SC(#n);
  GR(temin) GR(tgmin) GR(nnorm) GR(ennorm) GR(temp0) GR(vpnorm) GR(lnlam) GR(cfnus_i) GR(cfnus_e) GR(fcdif) GR(cthe) GR(flalftf) GR(cfnetap) GR(chioniz) GR(sigvi_floor) GR(cne_sgvi) GR(cnuiz)
  GR(cfrecom) GR(cfdiss) GR(cnucx) GR(sigcx) GR(rnn2cx) GR(fnuizx) GR(fnucxx) GR(fnnuiz) GR(cvgp) GR(oldseec) GR(cpgx) GR(fracvgpgp) GR(fluxfacy) GR(coef) GR(afix) GR(flalfv) GR(flgamv) GR(kxe) GR(kxi)
  GR(ce) GR(ci) GR(rkxecore) GR(kxicore) GR(kye) GR(kyi) GR(kyet) GR(kyit) GR(ckyet) GR(ckyit) GR(lmfplim) GR(alfkxi) GR(alfkxe) GR(flalfi) GR(flalfe) GR(flalfipl) GR(flalfepl) GR(lxtimax) GR(lxtemax)
  GR(tdiflim) GR(cftiexclg) GR(cfneut) GR(cfneutsor_ei) GR(cfneutsor_ee) GR(cfneutsor_ni) GR(cfneutsor_mi) GR(cfneutdiv) GR(cfneutdiv_fng) GR(cfneutdiv_fmg) GR(kxn) GR(kyn) GR(feqp) GR(alfeqp) GR(cnfx) GR(cnfy)
  GR(cnsor) GR(cmfx) GR(cmfy) GR(cfaccony) GR(fac2sp) GR(cfmsor) GR(flgam) GR(cfcvte) GR(cfcvti) GR(cfjhf) GR(cfloye) GR(cfloyi) GR(kye4order) GR(kyi4order) GR(bcee) GR(bcei) GR(chradi) GR(chradr)
  GR(ebind) GR(ediss) GR(eion) GR(ctsor) GR(ceisor) GR(ccoldsor) GR(cfvisx) GR(cfvisy) GR(upvhflr) GR(tibg) GR(pwribkg_c) GR(cflbg) GR(difcng) GR(alftng) GR(gcfacgx) GR(gcfacgy) GR(flgamg) GR(cngsor)
  GR(nurlxn) GR(nurlxu) GR(nurlxe) GR(nurlxi) GR(nurlxg) GR(nurlxp) GR(tcoree) GR(tcorei) GR(pcoree) GR(pcorei) GR(sygytotc) GR(csfacti) GR(cfueb) GR(cgpld) GR(cmneut) GR(eedisspl) GR(eidisspl) GR(cmntgpl)
  GR(ckinfl) GR(isoldalbarea) GR(tbmin) GR(nufak) GR(dtreal) GR(dtphi) GR(dylconst) GR(jaccliplim) GR(kelhihg) GR(kelhghg) GR(lgvmax) GR(flgamvg) GR(cfvisxn) GR(cfvisyn) GR(flgamtg) GR(cfupcx) GR(cfticx)
  GR(cfnidh) GR(cfnidh2) GR(cfnidhdis) GR(cfnidhgy) GR(cfnidhg2) GR(cftgeqp) GR(flalftxy) GR(flalfgnx) GR(flalfgny) GR(nlimgx) GR(nlimgy) GR(cfloxiplt) GR(cfloygwall) GR(cfjve) GR(rsigpl) GR(rsigplcore)
  GR(bcen) GR(cfqym) GR(cfqydt) GR(cfqyao) GR(cfsigm) GR(cfyef) GR(cf2ef) GR(cfybf) GR(cf2bf) GR(cfcurv) GR(cfgradb) GR(eycore) GR(icoreelec) GR(cfniybbo) GR(cfeeybbo)
  GR(cfqybf) GR(cfq2bf) GR(cfqybbo) GR(cfqydbo) GR(cfwjdotelim) GR(tebg) GR(cfydd) GR(cf2dd) GR(cfrd) GR(cfbgt) GR(cfjpy) GR(cfjp2) GR(cfvycf) GR(cfvycr) GR(cfeta1) GR(cfrtaue) GR(cfcl_e) GR(cfcl_i) GR(omgci_taui) GR(omgce_taue) GR(nuneo) GR(cfqyn)
#undef GR
  g.erad = SC("erad"); g.delpert = SC("del");
  g.sigma1_ = SC("sigma1"); g.frfqpn_ = SC("frfqpn"); g.cffqpsat_ = SC("cffqpsat"); g.exjbdry_ = SC("exjbdry"); g.rnewpot_ = SC("rnewpot"); g.cfqyae_ = SC("cfqyae"); g.cfqyai_ = SC("cfqyai");
  g.cfgpijr_ = SC("cfgpijr"); g.sigbar0_ = SC("sigbar0"); g.r0slab_ = SC("r0slab"); g.dx0_ = SC("dx0"); g.nfqya0core_ = I("nfqya0core"); g.nfqya0pf_ = I("nfqya0pf"); g.nfqya0ow_ = I("nfqya0ow");
  g.kappamx_ = SC("kappamx"); g.cfkincor_ = SC("cfkincor"); g.gamsec_ = SC("gamsec"); g.cgengpl_ = SC("cgengpl"); g.cgmompl_ = SC("cgmompl"); g.nglfix_ = SC("nglfix"); g.ngrfix_ = SC("ngrfix");
  g.eedisspr_ = SC("eedisspr"); g.eidisspr_ = SC("eidisspr"); g.cmntgpr_ = SC("cmntgpr"); g.phintewi_ = SC("phintewi"); g.phintewo_ = SC("phintewo");
  const size_t ns = g.nisp;
#define GV(n) VEC(#n, ns, g.n);
  GV(mi) GV(zi) GV(n0) GV(vcony) GV(difpr) GV(difni) GV(difni2) GV(difpr2) GV(difax) GV(travis) GV(parvis) GV(nlimix) GV(nlimiy) GV(dif4order) GV(cpiup) GV(cfvgpx) GV(cfvgpy) GV(cfvcsx) GV(cfvcsy) GV(cfvisxy)
  GV(cngmom) GV(cmwall) GV(ncore) GV(upcore) GV(curcore) GV(csfaclb) GV(csfacrb) GV(nwimin) GV(nwomin)
#undef GV
  VEC("fnorm", g.nusp, g.fnorm); VEC("difutm", ns, g.difutm_);
  VEC("n0g", 1, g.n0g_); VEC("mg", 1, g.mg_); VEC("ngbackg", 1, g.ngbackg_); VEC("cngtgx", 1, g.cngtgx); VEC("cngtgy", 1, g.cngtgy); VEC("cdifg", 1, g.cdifg); VEC("lgmax", 1, g.lgmax); VEC("lgtmax", 2, g.lgtmax);
  VEC("rld2dxg", 1, g.rld2dxg); VEC("rld2dyg", 1, g.rld2dyg); VEC("cngflox", 1, g.cngflox); VEC("cngfloy", 1, g.cngfloy); VEC("rtg2ti", 1, g.rtg2ti); VEC("tgas", 1, g.tgas); VEC("istgcon", 1, g.istgcon);
  VEC("keligig", 1, g.keligig); VEC("cngfx", 1, g.cngfx_); VEC("cngfy", 1, g.cngfy_); VEC("ngcore", 1, g.ngcore); VEC("albedoc", 1, g.albedoc); VEC("recycp", 1, g.recycp);
  if (g.ineudif != 1 && g.ineudif != 2) { g_err = "ineudif must be 1 or 2"; return -5; }
  if (g.ineudif == 1 && (g.isnonog != 0 || g.isupgon != 0)) { g_err = "ineudif=1 is built for orthogonal meshes and diffusive atoms only"; return -5; }
  const size_t nc = g.NC, nxs = g.NXS, nys = g.ny + 2;
#define GP(n) g.n = ARR(#n, nc);
  GP(vol) GP(gx) GP(gy) GP(gxf) GP(gyf) GP(gxc) GP(gyc) GP(sx) GP(sxnp) GP(sy) GP(rr) GP(rrv) GP(volv) GP(syv) GP(dxnog) GP(dynog) GP(btot) GP(rbfbt) GP(rbfbt2) GP(lcone) GP(lconi) GP(angfx) GP(ngfix) GP(curvrby) GP(gradby) GP(curvrb2) GP(gradb2)
#undef GP
  g.b_c = ARR("b_c", nc); g.rm_c = ARR("rm_c", nc);
  g.ixm1d = ARR("ixm1", nc); g.ixp1d = ARR("ixp1", nc); g.isxptyd = ARR("isxpty", nc); g.isxptxd = ARR("isxptx", nc);
  for (int k = 0; k < 2; ++k) {
    auto S2 = [&](const char* n) { const double* p = ARR(n, 2 * nc); return p ? p + (size_t)k * nc : nullptr; };
    g.fxm[k] = S2("fxm"); g.fx0[k] = S2("fx0"); g.fxp[k] = S2("fxp"); g.fxmy[k] = S2("fxmy"); g.fxpy[k] = S2("fxpy"); g.fym[k] = S2("fym"); g.fy0[k] = S2("fy0"); g.fyp[k] = S2("fyp");
    g.fymx[k] = S2("fymx"); g.fypx[k] = S2("fypx"); g.fymv[k] = S2("fymv"); g.fy0v[k] = S2("fy0v"); g.fypv[k] = S2("fypv"); g.fymxv[k] = S2("fymxv"); g.fypxv[k] = S2("fypxv");
  }
#define GLX(n) g.n = ARR(#n, nxs);
#define GLY(n) g.n = ARR(#n, nys);
  GLX(fgtdx) GLY(fgtdy) GLX(flalfea) GLX(flalfia) GLX(flalfva) GLX(flalfgxa) GLX(flalfgxya) GLY(flalfgya) GLX(flalfvgxa) GLY(flalfvgya) GLX(flalfvgxya) GLX(flalftgxa) GLY(flalftgya) GLY(yyf)
  GLX(nwalli) GLX(nwallo) GLX(lytepf) GLX(lytewc) GLX(lytipf) GLX(lytiwc) GLX(lynipf) GLX(lyniwc) GLX(tewalli) GLX(tiwalli) GLX(tewallo) GLX(tiwallo) GLY(recylb) GLY(recyrb) GLY(alblb) GLY(albrb)
  GLX(recycwot) GLX(recycwit) GLX(fngysi) GLX(fngyso) GLX(fngyi_use) GLX(fngyo_use) GLY(fngxslb) GLY(fngxsrb) GLY(fngxlb_use) GLY(fngxrb_use) GLX(albedoi) GLX(albedoo)
  GLX(istepfcix) GLX(istipfcix) GLX(istewcix) GLX(istiwcix) GLX(matwalli) GLX(matwallo) GLX(isixcore)
  GLY(recycmlb) GLY(recycmrb) GLX(lyphiix1) GLX(lyphiix2) GLX(iphibcwoix) GLX(iphibcwiix) GLY(phi0l) GLY(phi0r) GLY(bctype)
#undef GLX
#undef GLY
  g.lyup_ = ARR("lyup", 2);
  g.isnwconiix = ARR("isnwconiix", 2 * nxs); g.isnwconoix = ARR("isnwconoix", 2 * nxs); g.isupwiix = ARR("isupwiix", 2 * nxs); g.isupwoix = ARR("isupwoix", 2 * nxs);
  g.iseqalgd = ARR("iseqalg", (size_t)g.neq); g.igyld = ARR("igyl", (size_t)(2 * g.neq));
  {  // first unknown (0-based) of every mesh row; the unknowns are numbered row by row (convert.m:25-155)
    const V* ig = find("igyl");
    if (!ig || ig->size() < (size_t)(2 * g.neq)) { g_err = "missing input igyl"; return -1; }
    std::vector<int> rowiv(g.ny + 3, (int)g.neq);
    int prev = -1;
    for (int64_t k = 0; k < g.neq; ++k) {
      const int r = (int)(*ig)[(size_t)g.neq + k];
      if (r < prev || r < 0 || r > g.ny + 1) { g_err = "igyl: the unknowns are not numbered row by row"; return -1; }
      prev = r;
      if (rowiv[r] == (int)g.neq) rowiv[r] = (int)k;
    }
    for (int r = g.ny + 1; r >= 0; --r) if (rowiv[r] == (int)g.neq) rowiv[r] = rowiv[r + 1];
    g.rowuniform_ = 1;  // every cell holds numvar unknowns, numbered along the row: unknown k of row r belongs to column (k - rowiv[r]) / numvar
    const int64_t nv = (int64_t)I("numvar");
    for (int64_t k = 0; k < g.neq && g.rowuniform_; ++k) {
      const int r = (int)(*ig)[(size_t)g.neq + k], x = (int)(*ig)[(size_t)k];
      if (nv < 1 || (k - rowiv[r]) / nv != x) g.rowuniform_ = 0;
    }
    int* d = alloc_as<int>(rowiv.size());
    if (!d || !mem_put(d, rowiv.data(), rowiv.size() * sizeof(int))) return -10;
    g.rowiv_ = d;
  }
  for (int f = 0; f < 2; ++f) {
    const double *pn = ARR("idxn", 2 * nc), *pu = ARR("idxu", 2 * nc);
    g.idxn_[f] = pn ? pn + (size_t)f * nc : nullptr; g.idxu_[f] = pu ? pu + (size_t)f * nc : nullptr;
  }
  g.idxte_ = ARR("idxte", nc); g.idxti_ = ARR("idxti", nc); g.idxg_ = ARR("idxg", nc); g.idxphi_ = ARR("idxphi", nc);
  if (!g_missing.empty()) { g_err = "missing inputs: " + g_missing; return -1; }
  if (!g_err.empty()) return -10;
  g.iigsp = -1;
  if (g.isupgon == 1) { if (g.nisp != 2 || g.zi[1] != 0.) { g_err = "isupgon=1 needs nisp=2 with zi(2)=0"; return -5; } g.iigsp = 1; }
  // switches outside what is built
  struct { const char* n; double want; } must[] = {{"isimpon", 0}, {"ismcnon", 0}, {"ishymol", 0}, {"ifixsrc", 0}, {"ifixpsor", 0}, {"isupdrag", 0}, {"isofric", 0}, {"ishosor", 0}, {"islimon", 0}, {"isudsym", 0},
                                                   {"nxomit", 0}, {"isbohmcalc", 1}, {"isdifbetap", 0}, 
                                                   {"cftef", 0}, {"cftdd", 0}, 
                                                   {"facbni", 0}, {"facbup", 0}, {"facbee", 0}, {"facbei", 0}, {"rtauxfac", 0}, {"ispsorave", 0}, {"iseesorave", 0}, {"cfvisxneov", 0},
                                                   {"cfvisxneoq", 0}, {"cfvyavis", 0}, {"cfanomvisxg", 0}, {"cfanomvisyg", 0}, {"isnfmiy", 0}, {"isfeexpl0", 0},
                                                   {"isfeixpl0", 0}, {"cfeexdbo", 0}, {"cfeixdbo", 0}, {"isextrnp", 0}, {"isextrnpf", 0},
                                                   {"isextrtpf", 0}, {"isextrngc", 0}, {"isextrnw", 0}, {"isextrtw", 0}, {"isbohmms", 0}, {"ibctepl", 1}, {"ibctipl", 1}, {"ibctepr", 1}, {"ibctipr", 1}, {"isfixrb", 0},
                                                   {"is1D_gbx", 0}, {"isnglf", 0}, {"iszeffcon", 0}, {"isup1up2", 0}, {"isflxvar", 0}, {"ikapmod", 0},
                                                   {"isphicore0", 0}, {"iskaprex", 0}, {"isrozhfac", 0}};
  for (auto& m : must) {
    const V* v = find(m.n);
    if (!v) { g_err = std::string("missing input ") + m.n; return -1; }
    if ((*v)[0] != m.want) { g_err = std::string("switch outside the built set: ") + m.n; return -5; }
  }
  if (g.isnewpot != 0 && g.isnewpot != 1) { g_err = "isnewpot must be 0 or 1"; return -5; }
  if (g.isnewpot * g.isphion == 1 && (g.iphibcc < 1 || g.iphibcc > 3)) { g_err = "only iphibcc = 1, 2, 3 available"; return -5; }
  g.ExtendedJacPhi = I("ExtendedJacPhi"); g.isphilbc = I("isphilbc"); g.isphirbc = I("isphirbc"); g.isfqpave = I("isfqpave"); g.numvar_ = I("numvar");
  g.gridmode = 0; g.gbar = nullptr; g.gflag = nullptr;
  g.rowlo = 0; g.rowhi = g.ny + 1;
  // gas energy equation (istgon = 1): the inertial atoms only
  g.idxtg_ = ARR("idxtg", nc);
  g.istgcore = I("istgcore", 0); g.istgpfc = I("istgpfc", 0); g.istgwc = I("istgwc", 0); g.istglb = I("istglb", 0); g.istgrb = I("istgrb", 0); g.isfegxyqflave = I("isfegxyqflave");
  g.tgcore = SC("tgcore"); g.cftgticore = SC("cftgticore"); g.tgwall = SC("tgwall"); g.lytg1 = SC("lytg", 0); g.lytg2 = SC("lytg", 1); g.cftgtipltl = SC("cftgtipltl"); g.cftgtipltr = SC("cftgtipltr");
  g.cftgtipfc = SC("cftgtipfc"); g.cftgtiwc = SC("cftgtiwc"); g.cgengmpl = SC("cgengmpl"); g.cgengmw = SC("cgengmw"); g.cfalbedo = SC("cfalbedo"); g.recyce = SC("recyce"); g.recycwe = SC("recycwe");
  g.cvgpg = SC("cvgpg"); g.cfcvtg = SC("cfcvtg"); g.cfegxy = SC("cfegxy"); g.flalftgxy = SC("flalftgxy");
  if (g.istgon != 0 && g.istgon != 1) { g_err = "istgon must be 0 or 1"; return -5; }
  if (g.istgon == 1) {
    if (g.isupgon != 1 || g.nisp != 2) { g_err = "istgon=1 is built for inertial atoms (isupgon=1, nisp=2) only"; return -5; }
    if (g.istgpfc < 0 || g.istgpfc > 5 || g.istgwc < 0 || g.istgwc > 5) { g_err = "invalid istgpfc / istgwc"; return -5; }
    for (int k : {g.istglb, g.istgrb}) if (k < 0 || k > 5 || k == 2) { g_err = "istglb / istgrb must be 0, 1, 3, 4 or 5"; return -5; }
    if (SC("ispfbcvsix") != 0. || SC("iswobcvsix") != 0.) { g_err = "poloidally dependent wall options with istgon=1 not built"; return -5; }
  }
  if (g.fnnuiz != 1.) { g_err = "fnnuiz must be 1"; return -5; }
  if (SC("l_parloss") <= 1e9) { g_err = "l_parloss<=1e9 (nuvl) not built"; return -5; }
  if (g.isfixlb != 0 && g.isfixlb != 2) { g_err = "isfixlb must be 0 or 2"; return -5; }
  if (g.istabon != 0 && g.istabon != 7 && g.istabon != 10) { g_err = "istabon must be 0, 7 or 10"; return -5; }
  g.mpe = I("mpe"); g.mpd = I("mpd");
  if (g.istabon == 10) {
    const size_t nt = (size_t)g.mpe * g.mpd;
    g.wsveh = ARR("wsveh", nt); g.wsveh0 = ARR("wsveh0", nt); g.welms1 = ARR("welms1", nt); g.welms2 = ARR("welms2", nt);
    V dk(g.mpd), ek(g.mpe);
    dk[0] = 16.0; for (int j = 1; j < g.mpd; ++j) dk[j] = dk[j - 1] + 0.5;
    g.rldmin = dk[0]; g.rldmax = dk[g.mpd - 1]; g.deldkpt = (g.rldmax - g.rldmin) / double(g.mpd - 1);
    ek[0] = -1.2 * ue_log(10.0); for (int j = 1; j < g.mpe; ++j) ek[j] = ek[j - 1] + 0.1 * ue_log(10.0);
    g.rlemin = ek[0]; g.rlemax = ek[g.mpe - 1]; g.delekpt = (g.rlemax - g.rlemin) / double(g.mpe - 1);
    IN["__dkpt"] = dk; IN["__ekpt"] = ek;
    g.dkpt = ARR("__dkpt", g.mpd); g.ekpt = ARR("__ekpt", g.mpe);
  }
  if (!g_missing.empty()) { g_err = "missing inputs: " + g_missing; return -1; }
  // ---- field planes: base set, initial contents (fields of equations that are off keep what the host left in them)
  NPL = Gen::nplanes();
  const size_t nslab = (size_t)NPL * nc;
  V slab(nslab, 0.0);
  g.assign_planes(slab.data());
  auto init_plane = [&](const char* n, double* dst) { const V* v = find(n); if (v && v->size() >= nc) std::copy(v->begin(), v->begin() + nc, dst); };
  init_plane("ni1_init", g.ni[0]); init_plane("ni2_init", g.ni[1]); init_plane("up1_init", g.up[0]); init_plane("up2_init", g.up[1]);
  init_plane("te_init", g.te); init_plane("ti_init", g.ti); init_plane("ng_init", g.ng); init_plane("tg_init", g.tg); init_plane("phi_init", g.phi);
  if (g.isngon != 1 && g.isupgon != 1) { const V* v = find("ngfix"); std::copy(v->begin(), v->begin() + nc, g.ng); }
  for (int f = 0; f < g.nisp; ++f) for (size_t c = 0; c < nc; ++c) g.nm[f][c] = g.ni[f][c] * g.mi[f];
  d_base = mem_alloc(nslab); if (!d_base) return -10; g_allocs.push_back(d_base);
  if (!mem_put(d_base, slab.data(), nslab * sizeof(double))) return -10;
  g.assign_planes(d_base);
  const size_t neq = (size_t)g.neq;
  d_step = mem_alloc(4 * neq); if (!d_step) return -10; g_allocs.push_back(d_step);
  {
    V st(4 * neq, 1.0);
    for (size_t i = 0; i < neq; ++i) { st[i] = 1e20; st[neq + i] = 0.; }
    mem_put(d_step, st.data(), st.size() * sizeof(double));
  }
  g.dtuse = d_step; g.ylodt = d_step + neq; g.suscal = d_step + 2 * neq; g.sfscal = d_step + 3 * neq;
  d_yl = mem_alloc(neq + 2); d_yldot = mem_alloc(neq); d_y00 = mem_alloc(neq);
  if (!d_yl || !d_yldot || !d_y00) return -10;
  g_allocs.push_back(d_yl); g_allocs.push_back(d_yldot); g_allocs.push_back(d_y00);
  d_err = alloc_as<int>(4); d_gbar = alloc_as<int>(2);
  d_G = alloc_as<Gen>(1);
  if (!d_err || !d_G) return -10;
  g.nth = 1; g.errc = 0;
  if (!mem_put(d_G, &g, sizeof(Gen))) return -10;
  if (const char* e = getenv("UE_GEN_COLCAP")) COLCAP = std::max(16, atoi(e));
  if (const char* e = getenv("UE_GEN_TPU")) g_tpu = atoi(e) > 32 ? 64 : 32;
  g_full_grid = -1;
  if (const char* e = getenv("UE_GEN_FULL_GRID")) g_full_grid = atoi(e);
#if !defined(UE_GEN_HOST)
  {  // grid-mode residual: needs cooperative launch, one resident block per SM, and switch sets whose only early exit is the
     // negative-density test (which is made uniform over the grid); the configuration errors of bouncon leave per thread
    int dev = 0, coop = 0, occ = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_gen_full_grid, 128, sizeof(Gen));
    bool cfg_ok = true;
    for (int f = 0; f < G.nisp; ++f) if (G.isnicore[f] != 0 && G.isnicore[f] != 1) cfg_ok = false;
    for (const char* nm : {"recylb", "recyrb"}) { const V* v = find(nm); if (v) for (double x : *v) if (x < -1.) cfg_ok = false; }
    g_grid_ok = coop != 0 && occ >= 1 && cfg_ok && d_gbar != nullptr;
  }
#endif
  if (const char* e = getenv("UE_GEN_BAND")) g_band = std::max(UE_GEN_BAND_DEFAULT, atoi(e));
  g_colpad = 2;
  if (const char* e = getenv("UE_GEN_COLPAD")) g_colpad = atoi(e);
  g_compact = 1;
  if (const char* e = getenv("UE_GEN_COMPACT")) g_compact = atoi(e) != 0;
  if (g_tpu > 32 || g_band > 2 * (G.ny + 2)) g_compact = 0;  // (the block-per-unknown kernel and all-rows bands keep full-size planes)

#if defined(UE_GEN_HOST)
  g_poison = getenv("UE_GEN_POISON") ? std::max(1, atoi(getenv("UE_GEN_POISON"))) : 0;
#endif
  g_ivmin = 1; g_ivmax = g.neq;
  g_ready = true;
  return 0;
}

int run_full(const double* yl_host, double* yldot_host) {
  const size_t neq = (size_t)G.neq;
  if (!mem_put(d_yl, yl_host, (neq + 2) * sizeof(double))) return -10;
#if defined(UE_GEN_HOST)
  Gen me = G; me.nth = 1; me.errc = 0;
  const int rc = eval_full(me, d_yl, d_yldot);
  if (rc) return report(rc, me.errc);
#else
  int zero[4] = {0, 0, 0, 0};
  if (!mem_put(d_err, zero, sizeof zero)) return -10;
  const int nthr = std::min(256, std::max(64, ((G.NC + 31) / 32) * 32));  // one cell per thread up to 256 threads
  static cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (!e0) { cudaEventCreate(&e0); cudaEventCreate(&e1); }
  cudaEventRecord(e0);
  // meshes beyond one block's reach: all cells at once on a co-resident grid (one cell per thread, at most one block per SM)
  const int gthr = 128;
  int nblk = std::min(g_sms, (G.NC + gthr - 1) / gthr);
  if (g_full_grid >= 0) nblk = std::min(g_sms, g_full_grid);  // (env UE_GEN_FULL_GRID: 0 = one block, n = n blocks)
  if (!g_grid_ok || (g_full_grid < 0 && G.NC <= 640)) nblk = 0;
  if (nblk > 1) {
    if (!ck(cudaMemsetAsync(d_gbar, 0, 2 * sizeof(int)), "grid barrier reset")) return -10;
    const Gen* a0 = d_G; double* a1 = d_base; const double* a2 = d_yl; double* a3 = d_yldot; int* a4 = d_err; unsigned* a5 = (unsigned*)d_gbar; int* a6 = d_gbar + 1;
    void* args[] = {&a0, &a1, &a2, &a3, &a4, &a5, &a6};
    if (!ck(cudaLaunchCooperativeKernel((const void*)k_gen_full_grid, dim3(nblk), dim3(gthr), args, sizeof(Gen), 0), "k_gen_full_grid launch")) return -10;
  } else
    k_gen_full<<<1, nthr, sizeof(Gen)>>>(d_G, d_base, d_yl, d_yldot, d_err);
  cudaEventRecord(e1);
  if (!ck(cudaGetLastError(), "k_gen_full launch") || !ck(cudaDeviceSynchronize(), "k_gen_full")) return -10;
  cudaEventElapsedTime(&g_full_ms, e0, e1);
  int e[4];
  if (!mem_get(e, d_err, sizeof e)) return -10;
  if (e[0]) return report(e[0], e[1]);
  if (nblk > 1) {
    int fl[2] = {0, 0};
    if (!mem_get(fl, d_gbar, sizeof fl)) return -10;
    if (fl[1] & 0x100) { g_err = "grid barrier of the residual kernel timed out (blocks not co-resident?): set UE_GEN_FULL_GRID=0"; return -10; }
  }
#endif
  if (yldot_host && !mem_get(yldot_host, d_yldot, neq * sizeof(double))) return -10;
  g_last_yl.assign(yl_host, yl_host + neq + 2);
  return 0;
}

}  // namespace

extern "C" {
int UE_PREFIX(clear)(void) { IN.clear(); return 0; }
int UE_PREFIX(set)(const char* n, const double* d, int64_t k) {
  if (!n || (!d && k > 0) || k < 0) { g_err = "set: null name / data"; return -1; }
  IN[n].assign(d, d + k);
  return 0;
}
const char* UE_PREFIX(last_error)(void) { return g_err.c_str(); }
int UE_PREFIX(init)(void) {
  g_err.clear();
#if !defined(UE_GEN_HOST)
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "no CUDA device: the general path has no CPU fallback"; return -10; }
#endif
  const int rc = init_all();
  if (rc) free_all();
  return rc;
}
int UE_PREFIX(finalize)(void) {
#if !defined(UE_GEN_HOST)
  if (gc_comm) { cudaDeviceSynchronize(); NCG.CommDestroy(gc_comm); gc_comm = nullptr; }
  gc_nranks = 1; gc_rank = 0;
#endif
  free_all();
  return 0;
}
int UE_PREFIX(step_params)(int64_t n, const double* dt, const double* yo, const double* su, const double* sf) {
  if (!g_ready) { g_err = "init not called"; return -1; }
  if (n != G.neq || !dt || !yo || !su || !sf) { g_err = "step_params: neq mismatch or null pointer"; return -1; }
  const size_t neq = (size_t)n;
  return (mem_put(d_step, dt, neq * 8) && mem_put(d_step + neq, yo, neq * 8) && mem_put(d_step + 2 * neq, su, neq * 8) && mem_put(d_step + 3 * neq, sf, neq * 8)) ? 0 : -10;
}
int UE_PREFIX(pandf1)(int64_t n, double time, const double* yl, double* yldot) {
  (void)time;
  if (!g_ready) { g_err = "init not called"; return -1; }
  if (n != G.neq || !yl || !yldot) { g_err = "pandf1: neq mismatch or null pointer"; return -1; }
  return run_full(yl, yldot);
}
int UE_PREFIX(set_column_range)(int64_t ivmin, int64_t ivmax) {
  if (!g_ready) { g_err = "init not called"; return -1; }
  g_ivmin = std::max<int64_t>(1, ivmin); g_ivmax = std::min<int64_t>(G.neq, ivmax);
  return 0;
}
int UE_PREFIX(jac_calc)(int64_t n, double t, const double* yl, const double* yldot00, int64_t ml, int64_t mu, int64_t nnzmx, double* jac, int64_t* ja, int64_t* ia, int64_t* nnz_out) {
  (void)t;
  if (!g_ready) { g_err = "init not called"; return -1; }
  if (n != G.neq || !yl || !yldot00 || !jac || !ja || !ia || !nnz_out) { g_err = "jac_calc: neq mismatch or null pointer"; return -1; }
  const size_t neq = (size_t)n;
  // the base planes must be those of yl (the reference calls jac_calc right after the residual at the same state)
  if (g_last_yl.size() != neq + 2 || std::memcmp(g_last_yl.data(), yl, (neq + 2) * sizeof(double)) != 0) {
    const int rc = run_full(yl, nullptr);
    if (rc) return rc;
  }
  if (!mem_put(d_yl, yl, (neq + 2) * 8) || !mem_put(d_y00, yldot00, neq * 8)) return -10;
  const int cap = COLCAP;
  // a private plane set: full-size planes, or band rows only (UE_GEN_COMPACT, the default) - then with one mesh of padding on both sides of
  // the whole allocation, because reads outside the band (harmless: see eval_column) land in neighbouring planes or slabs
  g_pad = g_compact ? (size_t)G.NC : 0;
  const size_t nslab = g_compact ? (size_t)(NPL - UE_GEN_NLINE) * (2 * g_band + 1) * G.NXS + (size_t)UE_GEN_NLINE * std::max(G.NXS, G.ny + 2) : (size_t)NPL * G.NC;
  int64_t ncols_all = std::max<int64_t>(0, g_ivmax - g_ivmin + 1);
#if !defined(UE_GEN_HOST)
  if (gc_nranks > 1) ncols_all = (ncols_all + gc_nranks - 1) / gc_nranks;
#endif
  // private plane sets: one per resident warp of the persistent column kernel (or per unknown, if fewer), within half of the free device
  // memory (host build: one set).  The allocation of the first Jacobian is kept.
  size_t chunk = 1;
  {
    size_t budget = 1ull << 30;
#if !defined(UE_GEN_HOST)
    if (g_occ1 == 0) {
      int occ1 = 1, occ4 = 1;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ1, k_gen_cols_q<3>, 128, 4 * sizeof(Gen));
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ4, k_gen_cols_q<4>, 128, 4 * sizeof(Gen));
      g_occ1 = std::max(1, occ1); g_occ4 = std::max(1, occ4);
    }
    const int64_t resident = (int64_t)g_sms * 4 * std::max(g_occ1, g_occ4);
    const int64_t want = std::max<int64_t>(4, std::min<int64_t>(ncols_all, g_tpu > 32 ? ncols_all : resident));
    chunk = (size_t)want;
    if (g_priv_cols < chunk) {  // a new allocation: within half of the free memory
      size_t fr = 0, tot = 0;
      if (cudaMemGetInfo(&fr, &tot) == cudaSuccess) budget = fr / 2;
      chunk = (size_t)std::max<int64_t>(4, std::min<int64_t>(want, (int64_t)(budget / (nslab * 8))));
    }
    if (g_priv_cols >= chunk) chunk = g_priv_cols;
#endif
  }
  if (g_priv_cols < chunk) {
    for (double* q : {d_priv, d_ylp, d_wk})  // a larger set replaces the one a narrower column range allocated
      if (q) { g_allocs.erase(std::remove(g_allocs.begin(), g_allocs.end(), (void*)q), g_allocs.end()); mem_free(q); }
    d_priv = mem_alloc(chunk * nslab + 2 * g_pad); d_ylp = mem_alloc(chunk * (neq + 2)); d_wk = mem_alloc(chunk * neq);
    if (!d_priv || !d_ylp || !d_wk) return -10;
    g_allocs.push_back(d_priv); g_allocs.push_back(d_ylp); g_allocs.push_back(d_wk);
    g_priv_cols = chunk;
  }
  if (!d_cnt) {
    d_cnt = alloc_as<int>(2 * neq + 2); d_frow = alloc_as<int>(neq * cap); d_fval = mem_alloc(neq * cap);
    d_ia = alloc_as<int64_t>(neq + 1);
    if (!d_cnt || !d_frow || !d_fval || !d_ia) return -10;
    g_allocs.push_back(d_fval);
  }
  if (g_nnzmx < nnzmx) {
    d_jac = mem_alloc((size_t)nnzmx); d_ja = alloc_as<int64_t>((size_t)nnzmx);
    d_jac2 = mem_alloc((size_t)nnzmx); d_ja2 = alloc_as<int64_t>((size_t)nnzmx);
    if (!d_jac || !d_ja || !d_jac2 || !d_ja2) return -10;
    g_allocs.push_back(d_jac); g_allocs.push_back(d_jac2);
    g_nnzmx = nnzmx;
  }
#if defined(UE_GEN_HOST)
  std::memset(d_cnt, 0, (2 * neq + 2) * sizeof(int));
  for (int64_t iv = g_ivmin; iv <= g_ivmax; ++iv) {
    Gen me = G; me.nth = 1; me.errc = 0;
    me.assign_planes(d_priv + g_pad);
    const int rc = eval_column(me, d_base, NPL, iv, d_yl, d_ylp, d_wk, d_y00, ml, mu, cap, d_frow + (size_t)(iv - 1) * cap, d_fval + (size_t)(iv - 1) * cap, d_cnt + (iv - 1), g_band, g_colpad, d_priv + g_pad, g_compact);
    if (rc) return report(rc, me.errc);
    if (d_cnt[iv - 1] > cap) { g_err = "column fragment capacity exceeded: set UE_GEN_COLCAP"; return -2; }
  }
  // csrcsc (svr/svrut4.m:1536-1608) on the fragments
  std::vector<int64_t> iao(neq + 1, 0);
  for (size_t c = 0; c < neq; ++c) for (int k = 0; k < d_cnt[c]; ++k) iao[d_frow[c * cap + k]]++;
  iao[0] = 1;
  for (size_t i = 1; i <= neq; ++i) iao[i] += iao[i - 1];
  const int64_t nnz = iao[neq] - 1;
  if (nnz > nnzmx) { g_err = "*** jac_calc -- More storage needed for Jacobian. Increase lenpfac."; return -2; }
  std::vector<int64_t> next(iao.begin(), iao.end() - 1);
  for (size_t c = 0; c < neq; ++c)
    for (int k = 0; k < d_cnt[c]; ++k) { const int r = d_frow[c * cap + k]; const int64_t p = next[r - 1]++; jac[p - 1] = d_fval[c * cap + k]; ja[p - 1] = (int64_t)c + 1; }
  std::copy(iao.begin(), iao.end(), ia);
  *nnz_out = nnz;
  return 0;
#else
  int zero[4] = {0, 0, 0, 0};
  if (!mem_put(d_err, zero, sizeof zero)) return -10;
  if (!ck(cudaMemset(d_cnt, 0, (2 * neq + 2) * sizeof(int)), "memset")) return -10;
  const int WPB = 4;
  static cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  if (!e0) { cudaEventCreate(&e0); cudaEventCreate(&e1); cudaEventCreate(&e2); }
  static cudaEvent_t ec = nullptr;
  if (!ec) cudaEventCreate(&ec);
  int64_t my_lo = g_ivmin, my_hi = g_ivmax;
  const int* my_list = nullptr;
  if (gc_nranks > 1) {
    if (g_ivmin != 1 || g_ivmax != G.neq) { g_err = "a column range and a communicator cannot be combined"; return -1; }
    my_list = d_list_all + gc_list_off[gc_rank]; my_lo = 1; my_hi = gc_list_off[gc_rank + 1] - gc_list_off[gc_rank];  // positions in the list
  }
  cudaEventRecord(e0);
  if (g_tpu <= 32) {
    // work list: this call's unknowns, those with a window over all ix first (same test as make_win's xccuts, oderhs.m:960-1019)
    const V* ig = find("igyl");
    std::vector<int> slow, fast;
    auto add = [&](int64_t iv) {
      const int xc = (int)(*ig)[(size_t)iv - 1], yc = (int)(*ig)[neq + (size_t)iv - 1];
      const bool cut = ((xc - G.xlinc <= G.ixpt1 + 1) && (xc + G.xrinc + 1 >= G.ixpt1) && (yc - G.yinc <= G.iysptrx1) && (G.iysptrx1 > 0)) ||
                       ((xc - G.xlinc <= G.ixpt2 + 1) && (xc + G.xrinc + 1 >= G.ixpt2) && (yc - G.yinc <= G.iysptrx2) && (G.iysptrx2 > 0));
      (cut ? slow : fast).push_back((int)iv);
    };
    if (gc_nranks > 1) for (int k = gc_list_off[gc_rank]; k < gc_list_off[gc_rank + 1]; ++k) add(gc_list_all[k]);
    else for (int64_t iv = g_ivmin; iv <= g_ivmax; ++iv) add(iv);
    if (const char* e = getenv("UE_GEN_DEBUG_LIST")) {  // developer switch: time the two classes separately (the Jacobian is then incomplete)
      if (e[0] == 's') fast.clear(); else if (e[0] == 'f') slow.clear();
    }
    g_nslow = (int)slow.size();
    slow.insert(slow.end(), fast.begin(), fast.end());
    const int ncol = (int)slow.size();
    if (!d_order) { d_order = alloc_as<int>(neq + 1); if (!d_order) return -10; }
    if (ncol > 0 && !mem_put(d_order, slow.data(), (size_t)ncol * sizeof(int))) return -10;
    int* d_queue = d_order + neq;
    if (!ck(cudaMemsetAsync(d_queue, 0, sizeof(int)), "work queue reset")) return -10;
    // few unknowns: every warp is alone on its scheduler, registers are free (no spills); many unknowns: 4 blocks per SM (128 registers, a
    // few spills) so that more chains overlap
    const bool many = ncol > 4096;
    const int64_t resident = (int64_t)g_sms * WPB * (many ? g_occ4 : g_occ1);
    int nslots = (int)std::max<int64_t>(1, std::min<int64_t>(std::min<int64_t>((int64_t)chunk, resident), ncol));
    if (const char* e = getenv("UE_GEN_SLOTS")) nslots = std::max(4, std::min(nslots, atoi(e)));  // developer switch: fewer resident warps
    const int nblk = (nslots + WPB - 1) / WPB;  // (a last block with idle warps never exceeds the allocation: chunk >= 4 and slots are clipped below)
    const int nblk_ok = std::min<int>(nblk, (int)(chunk / WPB));
    if (ncol > 0) {
      if (many) k_gen_cols_q<4><<<std::max(1, nblk_ok), 32 * WPB, WPB * sizeof(Gen)>>>(d_G, d_base, d_priv + g_pad, NPL, d_order, ncol, d_queue, d_yl, d_ylp, d_wk, d_y00, ml, mu, cap, d_frow, d_fval, d_cnt, d_err, g_band, g_colpad, nslab, g_compact);
      else k_gen_cols_q<3><<<std::max(1, nblk_ok), 32 * WPB, WPB * sizeof(Gen)>>>(d_G, d_base, d_priv + g_pad, NPL, d_order, ncol, d_queue, d_yl, d_ylp, d_wk, d_y00, ml, mu, cap, d_frow, d_fval, d_cnt, d_err, g_band, g_colpad, nslab, g_compact);
      if (!ck(cudaGetLastError(), "k_gen_cols_q launch")) return -10;
    }
  } else
  for (int64_t iv0 = my_lo; iv0 <= my_hi; iv0 += (int64_t)chunk) {
    const int ncol = (int)std::min<int64_t>((int64_t)chunk, my_hi - iv0 + 1);
    const int* ivl = my_list ? my_list + (iv0 - 1) : nullptr;
    // (the block-per-unknown variant, UE_GEN_TPU=64: one launch per memory-sized chunk, full-size private planes)
    k_gen_cols<1><<<ncol, g_tpu, sizeof(Gen)>>>(d_G, d_base, d_priv, NPL, iv0, ivl, ncol, d_yl, d_ylp, d_wk, d_y00, ml, mu, cap, d_frow, d_fval, d_cnt, d_err, g_tpu, g_band, g_colpad);
    if (!ck(cudaGetLastError(), "k_gen_cols launch")) return -10;
  }
  cudaEventRecord(e1);
  if (gc_nranks > 1) { const int rc = exchange_fragments(cap); if (rc) return rc; }
  cudaEventRecord(ec);
  int* rowcnt = d_cnt + neq;
  k_gen_count<<<(unsigned)((neq + 127) / 128), 128>>>((int64_t)neq, cap, d_cnt, d_frow, rowcnt);
  k_gen_scan<<<1, 1024>>>((int64_t)neq, rowcnt, d_ia, rowcnt);
  k_gen_fill<<<(unsigned)neq, 64>>>((int64_t)neq, cap, d_cnt, d_frow, d_fval, d_ia, rowcnt, nnzmx, d_jac2, d_ja2);
  k_gen_sortrows<<<(unsigned)((neq + 3) / 4), 128>>>((int64_t)neq, d_ia, nnzmx, d_jac2, d_ja2, d_jac, d_ja);
  cudaEventRecord(e2);
  if (!ck(cudaGetLastError(), "CSR kernels launch") || !ck(cudaDeviceSynchronize(), "Jacobian kernels")) return -10;
  cudaEventElapsedTime(&g_cols_ms, e0, e1); cudaEventElapsedTime(&g_comm_ms, e1, ec); cudaEventElapsedTime(&g_csr_ms, ec, e2);
  int e[4];
  if (!mem_get(e, d_err, sizeof e)) return -10;
  if (e[0]) return report(e[0], e[1]);
  if (e[2]) { g_err = "column fragment capacity exceeded (" + std::to_string(e[2]) + " entries): set UE_GEN_COLCAP"; return -2; }
  if (!mem_get(ia, d_ia, (neq + 1) * sizeof(int64_t))) return -10;
  const int64_t nnz = ia[neq] - 1;
  if (nnz > nnzmx) { g_err = "*** jac_calc -- More storage needed for Jacobian. Increase lenpfac."; return -2; }
  if (!mem_get(jac, d_jac, (size_t)nnz * 8) || !mem_get(ja, d_ja, (size_t)nnz * 8)) return -10;
  *nnz_out = nnz;
  return 0;
#endif
}
// one Jacobian over several GPUs: collective, after ue_gen_init on every rank; id128 from ue_gpu_comm_unique_id on rank 0
int UE_PREFIX(comm_init)(int64_t nranks, int64_t rank, const char* id128) {
#if defined(UE_GEN_HOST)
  (void)nranks; (void)rank; (void)id128; g_err = "host build: no communicator"; return -1;
#else
  if (!g_ready) { g_err = "init not called"; return -1; }
  if (nranks < 1 || rank < 0 || rank >= nranks || !id128) { g_err = "comm_init: bad rank / nranks"; return -1; }
  int rc = nccl_bind();
  if (rc) return rc;
  if (gc_comm) { cudaDeviceSynchronize(); NCG.CommDestroy(gc_comm); gc_comm = nullptr; }
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  NCK(NCG.CommInitRank(&gc_comm, (int)nranks, id, (int)rank));
  gc_nranks = (int)nranks; gc_rank = (int)rank;
  d_pk_row = nullptr; d_pk_val = nullptr;
  return nranks > 1 ? build_rank_lists((int)nranks) : 0;
#endif
}
int UE_PREFIX(comm_finalize)(void) {
#if !defined(UE_GEN_HOST)
  if (gc_comm) { cudaDeviceSynchronize(); NCG.CommDestroy(gc_comm); gc_comm = nullptr; }
  gc_nranks = 1; gc_rank = 0;
#endif
  return 0;
}
// CUDA-event times (ms) of the kernels of the last calls: full-domain residual, Jacobian columns, CSR transpose (0 in the host build)
int UE_PREFIX(last_kernel_ms)(double* resid_ms, double* cols_ms, double* csr_ms) {
  if (resid_ms) *resid_ms = g_full_ms;
  if (cols_ms) *cols_ms = g_cols_ms;
  if (csr_ms) *csr_ms = g_csr_ms;
  return 0;
}
int UE_PREFIX(last_comm_ms)(double* comm_ms) { if (comm_ms) *comm_ms = g_comm_ms; return 0; }
// copy a named intermediate plane of the base set out ("fnix1", "fnix2", "feex", ...)
int UE_PREFIX(get_plane)(const char* name, double* out) {
  if (!g_ready || !name || !out) { g_err = "get_plane: not initialised / null pointer"; return -1; }
  const size_t nc = (size_t)G.NC;
  size_t k = 0; long found = -1;
#define P1(x) if (found < 0 && std::strcmp(name, #x) == 0) found = (long)k; k += 1;
#define P2(x) if (found < 0 && std::strcmp(name, #x "1") == 0) found = (long)k; if (found < 0 && std::strcmp(name, #x "2") == 0) found = (long)k + 1; k += 2;
  UE_GEN_PLANES(P1, P2)
#undef P1
#undef P2
  if (found < 0) { g_err = std::string("no such plane: ") + name; return -1; }
  return mem_get(out, d_base + (size_t)found * nc, nc * sizeof(double)) ? 0 : -10;
}
}
