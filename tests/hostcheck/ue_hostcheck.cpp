// tests/hostcheck/ue_hostcheck.cpp — TEST INFRASTRUCTURE, not product code and not a CPU fallback.
//
// Compiles the DEVICE functions of the product (uedge_b200/csrc/ue_device.cuh: phase0_cell, the phase-1 role functions,
// phase2_guard, p2_*, phase3_*) for the host through a small shim and drives them with plain loops that mirror the kernels
// of ue_gpu.cu one to one (k_phase0..3 for the residual; k_jb_stage0 / p1a / p1b / p2 / p3c + the CSC->CSR transpose for
// the Jacobian, with the same private cells, candidate lists (ue_lists.hpp) and masks).  Purpose: the container that
// builds the library has no GPU; this lets `pytest -m "not gpu"` check the kernels' LOGIC (switch handling, index
// windows, dependency pruning, four-unknown layout) bit for bit against the oracle before the code reaches a B200.
// Nothing under uedge_b200/ loads this library; only tests/test_hostcheck.py does.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

// ---- shim: just enough CUDA vocabulary for the device header --------------------------------------------------------
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __constant__ static
#define __restrict__
#define __CUDACC__ 1
#define __ldg(p) (*(p))
static inline int atomicOr(int* p, int v) { const int o = *p; *p |= v; return o; }
using std::max;
using std::min;
static inline int min(int a, long b) { return a < (int)b ? a : (int)b; }
static inline int max(int a, long b) { return a > (int)b ? a : (int)b; }

#include "../../uedge_b200/csrc/ue_device.cuh"
#include "../../uedge_b200/csrc/ue_lists.hpp"

namespace {
UeStore S;
std::string g_err;
int nx, ny, NXS, NC, NV;
int64_t neq;
std::vector<double> base, tmp, dtuse, ylodt, suscal, sfscal, dtoptv;
int64_t g_ivmin = 1, g_ivmax = 0;

Acc<false> acc0() { Acc<false> a; std::memset(&a, 0, sizeof a); a.base = base.data(); a.NXS = NXS; a.NC = NC; return a; }

// k_phase0 .. k_phase3 of ue_gpu.cu
int residual(const double* yl, double* yldot, bool need_rows) {
  Acc<false> a = acc0();
  const Win w = make_win(D, -1, -1);
  int err = 0;
  for (int c = 0; c < NC; ++c) {
    double ycell[UE_NV] = {0., 0., 0., 0., 0.};
    for (int k = 0; k < NV; ++k) ycell[k] = yl[(size_t)c * NV + k];
    phase0_cell<false>(a, ycell, c % NXS, c / NXS, &err);
  }
  for (int c = 0; c < NC; ++c) { const int ix = c % NXS, iy = c / NXS; p1_xpart<false>(a, w, ix, iy); p1_ypart<false>(a, w, ix, iy); p1_visx<false>(a, w, ix, iy); }
  for (int c = 0; c < NC; ++c) {
    const int ix = c % NXS, iy = c / NXS;
    p1_fx<false>(a, w, ix, iy); p1_fy<false>(a, w, ix, iy); p1_exe<false>(a, w, ix, iy); p1_exi<false>(a, w, ix, iy); p1_ey<false>(a, w, ix, iy);
  }
  for (int c = 0; c < NC; ++c) {
    const int ix = c % NXS, iy = c / NXS;
    double r[UE_NV] = {0., 0., 0., 0., 0.};
    double* o = tmp.data() + (size_t)c * UE_NV;
    if (ix >= 1 && ix <= nx && iy >= 1 && iy <= ny) {
      p2_n<false>(a, ix, iy, r, D.iseqalg); p2_m<false>(a, w, ix, iy, r, D.iseqalg); p2_e<false>(a, ix, iy, r, D.iseqalg); p2_i<false>(a, ix, iy, r, D.iseqalg);
    } else phase2_guard<false>(a, w, ix, iy, r);
    for (int k = 0; k < UE_NV; ++k) o[k] = r[k];
  }
  if (err & 1) { g_err = "***  ni is negative - calculation stopped"; return -3; }
  if (err & 2) { g_err = "***  ng is negative - calculation stopped"; return -3; }
  if (!need_rows) return 0;
  for (int c = 0; c < NC; ++c) {
    const int ix = c % NXS, iy = c / NXS;
    double r[UE_NV];
    for (int k = 0; k < UE_NV; ++k) r[k] = tmp[(size_t)c * UE_NV + k];
    if (ix >= 1 && ix <= nx && iy >= 1 && iy <= ny) phase3_interior<false>(a, ix, iy, r, yl + (size_t)c * NV, yl[neq], D.iseqalg, dtuse.data(), ylodt.data());
    else if (D.isbcwdt == 1) phase3_dt(ix, iy, r, yl + (size_t)c * NV, yl[neq], (int64_t)c * NV, dtuse.data(), ylodt.data());
    for (int k = 0; k < NV; ++k) yldot[(size_t)c * NV + k] = r[k];
  }
  return 0;
}

bool slot_cell(int xc, int yc, int xw, int xe, int k, int& ix, int& iy) {  // jb_slot_cell
  iy = yc; ix = xc;
  if (k == 1) { ix = xw; return xw != xc; }
  if (k == 2) { ix = xe; return xe != xc && xe != xw; }
  if (k == 3) { iy = yc - 1; return yc >= 1; }
  return true;
}
}  // namespace

extern "C" {
int ue_hk_set_int(const char* n, int64_t v) { return S.set_int(n, v); }
int ue_hk_set_real(const char* n, double v) { return S.set_real(n, v); }
int ue_hk_set_real_array(const char* n, const double* d, int64_t k) { return S.set_real_array(n, d, k); }
int ue_hk_set_int_array(const char* n, const int64_t* d, int64_t k) { return S.set_int_array(n, d, k); }
const char* ue_hk_last_error(void) { return g_err.c_str(); }

int ue_hk_init(void) {
  std::string m = S.missing();
  if (!m.empty()) { g_err = "missing inputs: " + m; return -1; }
  const UeParams& P = S.p;
  nx = (int)P.nx; ny = (int)P.ny; NXS = nx + 2; NC = NXS * (ny + 2); neq = P.neq; NV = (int)P.numvar;
  D = P;  // host pointers: the shim's "constant memory"
  std::memset(&DT, 0, sizeof DT);
  DT.mpe = (int)P.mpe; DT.mpd = (int)P.mpd;
  DT.iscut = (P.isfixlb == 2 && P.iysptrx1 > 0) ? 1 : 0;
  {
    bool rare = P.isupcore >= 2 || P.iflcore == -1 || P.isngcore != 0;
    for (int ix = 0; ix < (int)P.nx + 2; ++ix)
      rare = rare || P.isnwconiix[ix] != 0 || P.isnwconoix[ix] != 0 || P.istepfcix[ix] >= 2 || P.istipfcix[ix] >= 2 || P.istewcix[ix] >= 2 || P.istiwcix[ix] >= 2 ||
             P.matwalli[ix] > 0 || P.matwallo[ix] > 0;
    DT.rarebc = rare ? 1 : 0;
  }
  if (P.istabon == 10) {  // as ue_gpu_init
    DT.dkpt[0] = 16.0; for (int j = 1; j < DT.mpd; ++j) DT.dkpt[j] = DT.dkpt[j - 1] + 0.5;
    DT.rldmin = DT.dkpt[0]; DT.rldmax = DT.dkpt[DT.mpd - 1]; DT.deldkpt = (DT.rldmax - DT.rldmin) / double(DT.mpd - 1);
    DT.ekpt[0] = -1.2 * std::log(10.0); for (int j = 1; j < DT.mpe; ++j) DT.ekpt[j] = DT.ekpt[j - 1] + 0.1 * std::log(10.0);
    DT.rlemin = DT.ekpt[0]; DT.rlemax = DT.ekpt[DT.mpe - 1]; DT.delekpt = (DT.rlemax - DT.rlemin) / double(DT.mpe - 1);
  }
  base.assign((size_t)PL_COUNT * NC, 0.); tmp.assign((size_t)NC * UE_NV, 0.);
  dtuse.assign(neq, 1e20); ylodt.assign(neq, 0.); suscal.assign(neq, 1.); sfscal.assign(neq, 1.); dtoptv.assign(neq, 0.);
  g_ivmin = 1; g_ivmax = neq;
  return 0;
}
int ue_hk_step_params(int64_t n, const double* dt, const double* yo, const double* su, const double* sf) {
  if (n != neq) return -1;
  dtuse.assign(dt, dt + n); ylodt.assign(yo, yo + n); suscal.assign(su, su + n); sfscal.assign(sf, sf + n);
  return 0;
}
int ue_hk_pandf1(int64_t n, double, const double* yl, double* yldot) { if (n != neq) return -1; D = S.p; return residual(yl, yldot, true); }
int ue_hk_set_column_range(int64_t a, int64_t b) { g_ivmin = a; g_ivmax = b; return 0; }

// k_set_dt
int ue_hk_set_dt(int64_t n, const double* yl, double* f0, double* dtuse_out) {
  D = S.p;
  int rc = residual(yl, f0, true);
  if (rc) return rc;
  for (int c = 0; c < NC; ++c) {
    const int ix = c % NXS, iy = c / NXS, iym1 = std::max(0, iy - 1), iyp1 = std::min(ny + 1, iy + 1);
    for (int k = 0; k < NV; ++k) {
      const int64_t iv = (int64_t)c * NV + k;
      bool wr = true;
      if (k == 1) {
        wr = (ix != nx + 2 * D.isbcwdt);
        if (wr) {
          const int ixm1u = std::max(0, IXM1(ix, iy)), ixp1u = std::min(nx + 1, IXP1(ix, iy));
          const double up_5ca = (fabs(ylodt[iv]) + fabs(ylodt[d_iv(ixm1u, iy, 1, NXS)]) + fabs(ylodt[d_iv(ixp1u, iy, 1, NXS)]) + fabs(ylodt[d_iv(ix, iyp1, 1, NXS)]) +
                                 fabs(ylodt[d_iv(ix, iym1, 1, NXS)])) / 5;
          if (fabs(f0[iv]) > D.cutlo) dtoptv[iv] = D.deldt * fabs(up_5ca / (f0[iv]));
        }
      } else dtoptv[iv] = D.deldt * fabs(ylodt[iv] / (f0[iv] + D.cutlo));
      double dt = dtuse[iv];
      if (wr) {
        const double o = dtoptv[iv];
        dt = D.model_dt == 0 ? D.dtreal : D.model_dt == 1 ? D.dtreal * o / (D.dtreal + o) : D.model_dt == 2 ? o : sqrt(D.dtreal * o);
      }
      if (D.isbcwdt == 0 && D.iseqalg[iv] == 1) dt = 1.e20;
      dtuse[iv] = dt;
    }
  }
  std::copy(dtuse.begin(), dtuse.end(), dtuse_out);
  (void)n;
  return 0;
}

// Batched Jacobian of ue_gpu.cu, one unknown after the other.  The base planes must describe yl (the caller evaluated
// pandf1(yl) just before, as psetnk does).
int ue_hk_jac_calc(int64_t n, double, const double* yl, const double* yldot00, int64_t ml, int64_t mu, int64_t nnzmx, double* jac, int64_t* ja, int64_t* ia,
                   int64_t* nnz_out) {
  if (n != neq) return -1;
  D = S.p;
  const UeParams& P = S.p;
  std::vector<double> rcsc; std::vector<int64_t> icsc, jcsc(neq + 1);
  std::vector<int> cand, cand_east;
  std::vector<double> priv(4 * PL_COUNT), rows, rres;
  std::vector<int> rmask;
  int err = 0;
  int64_t nnz = 1;
  for (int64_t iv = 1; iv <= neq; ++iv) {
    jcsc[iv - 1] = nnz;
    if (iv < g_ivmin || iv > g_ivmax) continue;
    const int xc = (int)P.igyl[iv - 1], yc = (int)P.igyl[neq + iv - 1];
    const Win w = make_win(P, xc, yc);
    const int c0 = xc + NXS * yc, xw = (int)P.ixm1[c0], xe = (int)P.ixp1[c0];
    cell_candidates(P, xc, yc, cand);
    const int ncnd = (int)cand.size();
    cand_east.assign(ncnd, -1);
    for (int l = 0; l < ncnd; ++l) {
      const int cc = cand[l], ecell = (int)P.ixp1[cc] + NXS * (cc / NXS);
      const auto it = std::lower_bound(cand.begin(), cand.end(), ecell);
      if (it != cand.end() && *it == ecell) cand_east[l] = (int)(it - cand.begin());
    }
    rows.assign((size_t)ncnd * UE_NV, 0.); rres.assign(ncnd, 0.); rmask.assign(ncnd, 0);
    Acc<true> a; std::memset(&a, 0, sizeof a);
    a.base = base.data(); a.NXS = NXS; a.NC = NC; a.priv = priv.data(); a.ps = 1; a.ks = PL_COUNT;
    a.xc = xc; a.yc = yc; a.xw = xw; a.xe = xe; a.rres = rres.data(); a.rmask = rmask.data(); a.rself = -1; a.reast = -1;
    // stage 0
    for (int k = 0; k < 4; ++k) {
      int ix, iy; slot_cell(xc, yc, xw, xe, k, ix, iy); if (iy < 0) iy = 0;
      for (int pl = 0; pl < PL_COUNT; ++pl) priv[(size_t)k * PL_COUNT + pl] = base[(size_t)pl * NC + ix + NXS * iy];
    }
    const double yold = yl[iv - 1];
    const double dyl = D.delpert * (fabs(yold) + D.dylconst / suscal[iv - 1]);
    {
      double ycell[UE_NV] = {0., 0., 0., 0., 0.};
      const int64_t c = (int64_t)c0 * NV;
      for (int k = 0; k < NV; ++k) ycell[k] = yl[c + k];
      ycell[(iv - 1) - c] = yold + dyl;
      phase0_cell<true>(a, ycell, xc, yc, &err);
    }
    for (int k = 0; k < 4; ++k) { int ix, iy; if (!slot_cell(xc, yc, xw, xe, k, ix, iy)) continue; p1_xpart<true>(a, w, ix, iy); p1_ypart<true>(a, w, ix, iy); p1_visx<true>(a, w, ix, iy); }
    for (int k = 0; k < 4; ++k) {
      int ix, iy; if (!slot_cell(xc, yc, xw, xe, k, ix, iy)) continue;
      p1_fx<true>(a, w, ix, iy); p1_fy<true>(a, w, ix, iy); p1_exe<true>(a, w, ix, iy); p1_exi<true>(a, w, ix, iy); p1_ey<true>(a, w, ix, iy);
    }
    // phase 2: guard role first, then the equation groups (k_jb_p2)
    for (int pass = 0; pass < 2; ++pass)
      for (int l = 0; l < ncnd; ++l) {
        const int cell = cand[l], ix = cell % NXS, iy = cell / NXS;
        const bool interior = ix >= 1 && ix <= nx && iy >= 1 && iy <= ny;
        double r[UE_NV] = {0., 0., 0., 0., 0.};
        a.rself = l;
        if (pass == 0 && !interior) {
          const int mk = phase2_guard<true>(a, w, ix, iy, r);
          for (int k = 0; k < UE_NV; ++k) rows[(size_t)k * ncnd + l] = r[k];
          rmask[l] |= mk;
        } else if (pass == 1 && interior && in_rng(ix, w.i2, w.i5) && in_rng(iy, w.j2, w.j5)) {
          p2_n<true>(a, ix, iy, r, D.iseqalg); p2_m<true>(a, w, ix, iy, r, D.iseqalg); p2_e<true>(a, ix, iy, r, D.iseqalg); p2_i<true>(a, ix, iy, r, D.iseqalg);
          for (int k = 0; k < UE_NV; ++k) rows[(size_t)k * ncnd + l] = r[k];
          rmask[l] |= 0x11f;
        }
      }
    a.rself = -1;
    // phase 3 on the interior candidate rows (k_jb_p3c)
    for (int l = 0; l < ncnd; ++l) {
      const int cell = cand[l], ix = cell % NXS, iy = cell / NXS;
      if (ix >= 1 && ix <= nx && iy >= 1 && iy <= ny && in_rng(ix, w.i2, w.i5) && in_rng(iy, w.j2, w.j5)) {
        double r[UE_NV], ycell[UE_NV] = {0., 0., 0., 0., 0.};
        const int64_t c = (int64_t)cell * NV;
        for (int k = 0; k < UE_NV; ++k) r[k] = rows[(size_t)k * ncnd + l];
        for (int k = 0; k < NV; ++k) ycell[k] = yl[c + k];
        if (ix == xc && iy == yc) ycell[(iv - 1) - c] = yold + dyl;
        a.reast = cand_east[l];
        phase3_interior<true>(a, ix, iy, r, ycell, yl[neq], D.iseqalg, dtuse.data(), ylodt.data());
        for (int k = 0; k < UE_NV; ++k) rows[(size_t)k * ncnd + l] = r[k];
      }
    }
    // difference, clip, band test, compaction (ordered)
    const int64_t ii1 = std::max(iv - mu, (int64_t)1), ii2 = std::min(iv + ml, neq);
    for (int l = 0; l < ncnd; ++l)
      for (int k = 0; k < NV; ++k) {
        const int64_t ii = (int64_t)cand[l] * NV + k + 1;
        if (ii < ii1 || ii > ii2) continue;
        const bool written = (rmask[l] >> k) & 1;
        if (!(written || ii == iv)) continue;
        const double y00 = yldot00[ii - 1];
        const double wk = written ? rows[(size_t)k * ncnd + l] : y00;
        double jacelem = (wk - y00) / dyl;
        if (ii == iv) {
          if (D.iseqalg[iv - 1] * (1 - D.isbcwdt) == 0) jacelem = jacelem - 1 / dtuse[iv - 1];
          if (D.nufak > 0 && yl[neq] == 1) jacelem = jacelem - D.nufak;
        }
        if (fabs(jacelem * sfscal[iv - 1]) > D.jaccliplim) {
          if (nnz > nnzmx) { g_err = "*** jac_calc -- More storage needed for Jacobian. Storage exceeded. Increase lenpfac."; return -2; }
          rcsc.push_back(jacelem); icsc.push_back(ii); nnz = nnz + 1;
        }
      }
  }
  jcsc[neq] = nnz;
  // CSC -> CSR (csrcsc, svr/svrut4.m:1536-1608): stable counting transpose
  std::fill(ia, ia + neq + 1, 0);
  for (int64_t k = 0; k < nnz - 1; ++k) ia[icsc[k]] += 1;   // count of row i at ia[i] (1-based rows -> shifted by one)
  ia[0] = 1;
  for (int64_t i = 1; i <= neq; ++i) ia[i] += ia[i - 1];
  std::vector<int64_t> fill(ia, ia + neq);
  for (int64_t j = 1; j <= neq; ++j)
    for (int64_t k = jcsc[j - 1]; k < jcsc[j]; ++k) {
      const int64_t i = icsc[k - 1], pos = fill[i - 1]++;
      jac[pos - 1] = rcsc[k - 1]; ja[pos - 1] = j;
    }
  *nnz_out = nnz - 1;
  return 0;
}
}
