// oracle/ue_oracle.cpp — TEST INFRASTRUCTURE, not product code.
//
// CPU restatement (plain scalar C++, double precision, no FMA contraction) of the
// reference's residual pandf1 and finite-difference Jacobian jac_calc for the
// switch set of the d3dHsm and slab families: single hydrogen ion species +
// diffusive atoms (nisp=nusp=ngsp=1, isupgon=0, istgon=0; isngon=1, or 0 = frozen
// atom density and four unknowns per cell), orthogonal mesh (isnonog=0), ix=0 a plate or a
// symmetry plane (isfixlb 0/2), no potential equation (isphion=0), no impurities
// (isimpon=0), ineudif=2 (neudifpg), rates istabon 0/7/10, all cross-field drift
// coefficients zero.  Any other switch value is refused by ue_ora_init.
//
// It keeps the reference's *stateful, windowed* semantics: all intermediate
// fields are persistent arrays (the Fortran module state); a call with xc,yc>=0
// recomputes only the index ranges i1..i8 x j1..j8 of bbb/oderhs.m:868-964 in
// place, and jac_calc perturbs one unknown at a time, calling pandf1 twice per
// unknown (perturb + restore), exactly as bbb/oderhs.m:8616-8745.
//
// Parity pins (tests/test_oracle_golden.py):
//  * PINNED: the residual.  At the reference's converged state pyexamples/d3dHsmNew/d3dHsm.h5 (with the defaults of
//    the UEDGE version that wrote it) max|yldot| <= 2.2e-6 against 1e5 a per-mille away, all five equations and all
//    boundary rows; geometry (guardc/nphygeo) equals the stored mesh; Newton on this residual + Jacobian returns to that
//    state within 4e-9 (normalised).
//  * SANITY ONLY: Forthon_case1 (slab, isfixlb=2, isngon=0, istabon=7).  Integrating this residual from the restart=0 profiles
//    to t = 4e-4 s lands within 0.3-5 % of the ni/up/te/ti arrays its 2007 output prints at that time.
//  * SANITY ONLY: Forthon_case2 (istabon=10 tables).  The converged midplane profiles land within 2-4 % of
//    output_forthon_case2.rtf (2007); its printed initial fnrm = 0.79266 is NOT reproduced (3.6156 here): defaults and
//    boundary models changed since 2007.
//  * PARITY UNPINNED against the true reference: individual Jacobian entries and the ia/ja pattern (no reference
//    fixture holds them; the reference cannot be built here).  They are pinned oracle<->CUDA (bit for bit) and
//    oracle<->finite differences of the pinned residual (windowed vs full evaluation), see DESIGN.md 2.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may load this library.
//
// Citations "oderhs.m:N" etc. are file:line in the reference tree, directory bbb/
// unless another directory is given.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "ue_math.h"
#include "ue_param_store.hpp"

namespace {

UeStore S;
UeParams& P = S.p;
int nx, ny, NXS, NC;
int64_t neq;
bool HASG = true;  // isngon = 1: the atom density is an unknown; isngon = 0: frozen field `ngfix`, numvar = 4

typedef std::vector<double> V;

// rate tables (istabon=10)
int mpe = 0, mpd = 0;
V wsveh, wsveh0, welms1, welms2, ekpt, dkpt;
double rlemin, rlemax, rldmin, rldmax, delekpt, deldkpt;
const char* plane_names =
    "ne nit nm nz2 ni te ti ng tg up pri pre pr zeff pg gprx gpry gpix gpiy gpex gtex gtix gpey gtey gtiy "
    "niy0 niy1 nity0 nity1 ney0 ney1 priy0 priy1 tey0 tey1 tiy0 tiy1 ngy0 ngy1 tgy0 tgy1 pgy0 pgy1 "
    "loglambda diffusivwrk vy frice frici ex upi uup uu upe vex vey "
    "nuiz nurc nucx nuix psorbgg psorgc psorc psordis psorxrc psorrgc psorg psor psorxr psorrg "
    "snic sniv psori smoc smov seec seev seic seiv conxg conyg floxg floyg fngx fngy resng "
    "visx visy hcxe hcxi hcye hcyi hcxij hcyij eqp w0 w1 w2 w3 fnix fniy resco flox floy conx cony fmix fmiy resmo "
    "floxe floxi floye floyi conxe conxi conye conyi feex feey feix feiy resee resei "
    "erliz erlrc eeli vsoreec vsoree wvh pwribkg";

#define A(a, ix, iy) a[(ix) + NXS * (iy)]
#define G(a, ix, iy) P.a[(ix) + NXS * (iy)]
inline int IXP1(int ix, int iy) { return (int)P.ixp1[ix + NXS * iy]; }
inline int IXM1(int ix, int iy) { return (int)P.ixm1[ix + NXS * iy]; }
// unknown numbering of `convert` (convert.m:33-152) with every equation on: 0-based
inline int64_t IDXN(int ix, int iy) { return ((int64_t)(ix + NXS * iy)) * P.numvar + 0; }
inline int64_t IDXU(int ix, int iy) { return ((int64_t)(ix + NXS * iy)) * P.numvar + 1; }
inline int64_t IDXTE(int ix, int iy) { return ((int64_t)(ix + NXS * iy)) * P.numvar + 2; }
inline int64_t IDXTI(int ix, int iy) { return ((int64_t)(ix + NXS * iy)) * P.numvar + 3; }
inline int64_t IDXG(int ix, int iy) { return ((int64_t)(ix + NXS * iy)) * P.numvar + 4; }

inline double ave(double t0, double t1) { return 2 * t0 * t1 / (P.cutlo + t0 + t1); }  // oderhs.m:697
inline double sgn(double a, double b) { return std::copysign(std::fabs(a), b); }         // Fortran sign(a,b)
inline double powi(double x, int64_t n) {  // integer power by repeated squaring (x**n, n>=0)
  double r = 1.0;
  while (n > 0) { if (n & 1) r *= x; x *= x; n >>= 1; }
  return r;
}

// ---- hydrogen rates (aph/aphrates.m) --------------------------------------------
void table_idx(double tev_j, double dens, int& je, int& jd, double& fje, double& fjd) {
  // aph/aphrates.m:1043-1056 (identical in rra/erl1/erl2)
  double zloge = ue_log(tev_j / P.ev);
  double rle = std::max(rlemin, std::min(zloge, rlemax));
  double zlogd = ue_log10(dens);
  double rld = std::max(rldmin, std::min(zlogd, rldmax));
  je = (int)((rle - rlemin) / delekpt) + 1; je = std::min(je, mpe - 1);
  jd = (int)((rld - rldmin) / deldkpt) + 1; jd = std::min(jd, mpd - 1);
  fje = (rle - ekpt[je - 1]) / (ekpt[je] - ekpt[je - 1]);
  fjd = (rld - dkpt[jd - 1]) / (dkpt[jd] - dkpt[jd - 1]);
}
double table_val(const V& w, double tev_j, double dens) {
  int je, jd; double fje, fjd;
  table_idx(tev_j, dens, je, jd, fje, fjd);
  auto W = [&](int a, int b) { return ue_log(w[(a - 1) + mpe * (b - 1)]); };
  double r11 = W(je, jd), r12 = W(je, jd + 1), r21 = W(je + 1, jd), r22 = W(je + 1, jd + 1);
  double r1 = r11 + fjd * (r12 - r11);
  double r2 = r21 + fjd * (r22 - r21);
  return ue_exp(r1 + fje * (r2 - r1));
}
// istabon = 7 (the package default, com/com.v:322): R.B. Campbell's polynomial fits in x = log10(ne), y = log10(Te[eV])
double sionf(double temp, double den) {  // aph/aphrates.m:1133-1176
  auto ain = [](double x) { return -49.05905 + 2.51313783 * x - 0.049159714 * x * x; };
  auto bin = [](double x) { return 41.1855162 - 2.3298672 * x + 4.24769144e-2 * x * x; };
  auto cin = [](double x) { return -32.798921 + 1.72102919 * x - 0.038692357 * x * x; };
  auto din = [](double x) { return 27.370466 - 1.6824361 * x + 0.0462317894 * x * x; };
  auto ein = [](double x) { return -7.9990454 + 0.127573157 * x - 6.3586911e-3 * x * x; };
  auto gin = [](double x) { return -4.5832951 + 0.776264783 * x - 1.8866089e-2 * x * x; };
  auto hin = [](double x) { return 3.08056833 - 0.39114789 * x + 9.86833304e-3 * x * x; };
  auto riin = [](double x) { return -0.4648639 + 0.0551428018 * x - 1.404213e-3 * x * x; };
  double x = std::min(22.e0, ue_log10(den)), y = ue_log10(temp);
  return ue_pow(10., ain(x) + bin(x) * y + cin(x) * y * y + din(x) * y * y * y + ein(x) * y * y * y * y + gin(x) * y * y * y * y * y +
                         hin(x) * y * y * y * y * y * y + riin(x) * y * y * y * y * y * y * y);
}
double srecf(double temp, double den) {  // aph/aphrates.m:1180-1226
  auto ar = [](double x) { return -0.4575652 - 2.144012 * x + 6.7072142e-2 * x * x - 1.391667e-4 * x * x * x; };
  auto br = [](double x) { return -121.8401 + 18.001822 * x - 0.8679488 * x * x + 1.33165e-2 * x * x * x; };
  auto cr = [](double x) { return 80.897256 - 13.29602 * x + 0.71881414 * x * x - 0.0126549 * x * x * x; };
  auto dr = [](double x) { return 56.406823 - 7.301996 * x + 0.29339793 * x * x - 3.50898e-3 * x * x * x; };
  auto er = [](double x) { return -55.73559 + 7.9634283 * x - 0.370274 * x * x + 5.567961e-3 * x * x * x; };
  auto gr = [](double x) { return 10.866692 - 1.584193 * x + 0.07563791 * x * x - 1.177562e-3 * x * x * x; };
  double x = std::min(22.e0, ue_log10(den)), y = ue_log10(temp);
  return ue_pow(10., ar(x) + br(x) * y + cr(x) * y * y + dr(x) * y * y * y + er(x) * y * y * y * y + gr(x) * y * y * y * y * y);
}
double svradp(double temp, double den) {  // aph/aphrates.m:1230-1300
  auto ai = [](double x) { return -275.845 + 37.010817 * x - 1.788045 * x * x + 0.029078333 * x * x * x; };
  auto bi = [](double x) { return 2200.9478 - 326.1153 * x + 16.148655 * x * x - 0.2660702 * x * x * x; };
  auto ci = [](double x) { return -2.935221e3 + 4.3757698e2 * x - 21.73964 * x * x + 0.358962 * x * x * x; };
  auto di = [](double x) { return 1604.1466 - 239.6959 * x + 11.923707 * x * x - 0.1970501 * x * x * x; };
  auto ei = [](double x) { return -390.8635 + 58.474495 * x - 2.910997 * x * x + 0.048133829 * x * x * x; };
  auto gi = [](double x) { return 35.012574 - 5.24202 * x + 0.26109962 * x * x - 4.319238e-3 * x * x * x; };
  auto ae = [](double x) { return 2860.4173 - 610.2452 * x + 48.275821 * x * x - 1.687994 * x * x * x + 0.02201375 * x * x * x * x; };
  auto be = [](double x) { return 10612.067 - 2046.397 * x + 147.73914 * x * x - 4.729973 * x * x * x + 0.056671796 * x * x * x * x; };
  auto ce = [](double x) { return -4.231708e4 + 8494.6102 * x - 639.0226 * x * x + 21.350311 * x * x * x - 0.2673466 * x * x * x * x; };
  auto de = [](double x) { return -8.385144e3 + 1887.6244 * x - 157.8502 * x * x + 5.820501 * x * x * x - 0.07992837 * x * x * x * x; };
  auto ee = [](double x) { return 3.938282e4 - 8.131339e3 * x + 628.8119 * x * x - 21.58636 * x * x * x + 0.27756029 * x * x * x * x; };
  auto ge = [](double x) { return -1.038281e4 + 2.1349333e3 * x - 164.4201 * x * x + 5.6210487 * x * x * x - 0.07197622 * x * x * x * x; };
  auto sionfl = [&](double x, double y) { return ue_pow(10., ai(x) + bi(x) * y + ci(x) * y * y + di(x) * y * y * y + ei(x) * y * y * y * y + gi(x) * y * y * y * y * y); };
  auto etai = [&](double x, double y) {
    return (ue_pow(10., ae(x) + be(x) * y + ce(x) * y * y + de(x) * y * y * y + ee(x) * y * y * y * y + ge(x) * y * y * y * y * y)) / sionfl(x, y);
  };
  double x = std::min(22.e0, ue_log10(den)), y = ue_log10(temp);
  return std::max(0.e0, (13.6e0 + etai(x, std::min(2.e0, y)))) * 1.602e-19 * sionfl(x, y);  // etai frozen above 100 eV
}
double rsa(double tej, double dens) {  // aph/aphrates.m:872-1131
  if (P.istabon == 0) { double a = tej / (10 * P.ev); return 3.0e-14 * a * a / (3.0 + a * a); }
  if (P.istabon == 7) return sionf(tej / P.ev, dens);  // :1033-1036
  return table_val(wsveh, tej, dens);
}
double rra(double tej, double dens) {  // aph/aphrates.m:617-870
  if (P.istabon == 0) return 0.;
  if (P.istabon == 7) return srecf(tej / P.ev, dens);  // :775-778
  return table_val(wsveh0, tej, dens);
}
double rcx(double t0) {  // aph/aphrates.m:395-399 (analytic for istabon 0 and >3)
  double a = 3 * t0 / (10 * P.ev);
  return 1.7e-14 * ue_pow(a, 0.333);
}
double rqa0(double tej) {  // aph/aphrates.m:444-447
  double a = tej / (10 * P.ev);
  return P.erad * P.ev * 3.0e-14 * a * a / (3.0 + a * a);
}
double erl1(double tej, double dens) {  // aph/aphrates.m:2-147
  if (P.istabon == 0) return (rqa0(tej) - 13.6 * P.ev * rsa(tej, dens)) * dens;
  if (P.istabon == 7) return (svradp(tej / P.ev, dens) - 13.6 * P.ev * rsa(tej, dens)) * dens;  // rqa: :601-606
  return table_val(welms1, tej, dens);
}
double erl2(double tej, double dens) {  // aph/aphrates.m:149-294
  if (P.istabon == 0 || P.istabon == 7) return (13.6 * P.ev + 1.5 * tej) * dens * rra(tej, dens);
  return table_val(welms2, tej, dens);
}

// ---- index window (oderhs.m:868-1019) -------------------------------------------
struct Win {
  int xc, yc;
  int i1, i2, i2p, i3, i4, i5, i5m, i6, i7, i8;
  int j1, j1p, j2, j2p, j3, j4, j5, j5m, j6, j5p, j6p, j7, j8;
  int ixs, ixf, iys, iyf, ixs1, ixf6, iys1, iyf6;
  bool openbox, xcnearlb, xcnearrb, xccuts;
};
Win make_win(int xc, int yc) {
  Win w; w.xc = xc; w.yc = yc;
  const int xlinc = (int)P.xlinc, xrinc = (int)P.xrinc, yinc = (int)P.yinc;
  if (xc < 0 || ((0 <= yc) && (yc - yinc <= 0) && P.isjaccorall == 1)) {
    w.i1 = 0; w.i2 = 1; w.i2p = 1; w.i3 = 0; w.i4 = 0; w.i5 = nx; w.i5m = nx - 1; w.i6 = nx + 1; w.i7 = nx + 1; w.i8 = nx + 1;
  } else {
    w.i1 = std::max(0, xc - xlinc - 1); w.i2 = std::max(1, xc - xlinc); w.i2p = std::max(1, xc - xrinc - 1);
    w.i3 = xc - xlinc; w.i4 = std::max(0, xc - xlinc); w.i5 = std::min(nx, xc + xrinc); w.i5m = std::min(nx - 1, xc + xrinc);
    w.i6 = std::min(nx + 1, xc + xrinc + 1); w.i7 = xc + xrinc; w.i8 = std::min(nx + 1, xc + xrinc);
  }
  if (yc < 0) {
    w.j1 = 0; w.j1p = 0; w.j2 = 1; w.j2p = 1; w.j3 = 0; w.j4 = 0; w.j5 = ny; w.j5m = ny - 1; w.j6 = ny + 1; w.j5p = ny;
    w.j6p = ny + 1; w.j7 = ny + 1; w.j8 = ny + 1;
  } else {
    w.j1 = std::max(0, yc - yinc - 1); w.j2 = std::max(1, yc - yinc); w.j1p = std::max(0, yc - yinc - 2);
    w.j2p = std::max(1, yc - yinc - 1); w.j3 = yc - yinc; w.j4 = std::max(0, yc - yinc); w.j5 = std::min(ny, yc + yinc);
    w.j5m = std::min(ny - 1, yc + yinc); w.j6 = std::min(ny + 1, yc + yinc); w.j5p = std::min(ny, yc + yinc + 1);
    w.j6p = std::min(ny + 1, yc + yinc + 1); w.j7 = yc + yinc; w.j8 = std::min(ny + 1, yc + yinc);
  }
  // widen near the X-point cuts (oderhs.m:927-964); xcturb needs isturbnloc*kyet, off in this switch set
  w.xccuts = false;
  if ((xc - xlinc <= P.ixpt1 + 1) && (xc + xrinc + 1 >= P.ixpt1) && (yc - yinc <= P.iysptrx1) && (P.iysptrx1 > 0)) w.xccuts = true;
  if ((xc - xlinc <= P.ixpt2 + 1) && (xc + xrinc + 1 >= P.ixpt2) && (yc - yinc <= P.iysptrx2) && (P.iysptrx2 > 0)) w.xccuts = true;
  if (w.xccuts) { w.i1 = 0; w.i2 = 1; w.i3 = 0; w.i4 = 0; w.i5 = nx; w.i6 = nx + 1; w.i7 = nx + 1; w.i8 = nx + 1; }
  w.ixs = w.i2; w.ixf = w.i5; w.iys = w.j2; w.iyf = w.j5; w.ixs1 = w.i1; w.ixf6 = w.i6; w.iys1 = w.j1; w.iyf6 = w.j6;
  if (xc >= 0 && yc >= 0) {
    w.ixs = xc; w.ixf = xc; w.iys = yc; w.iyf = yc; w.ixs1 = xc; w.ixf6 = xc;
    if (xrinc >= 20) { w.ixs1 = 0; w.ixf6 = nx + 1; }
    w.iys1 = yc; w.iyf6 = yc;
    if (yinc >= 20) { w.iys1 = 0; w.iyf6 = ny + 1; }
  }
  if (xc < 0) w.openbox = true;
  else if (w.xccuts) w.openbox = true;
  else if ((0 <= yc) && (yc <= yinc)) w.openbox = true;
  else w.openbox = false;
  w.xcnearlb = ((xc - xlinc <= P.ixlb) && (xc + xrinc >= P.ixlb)) || xc < 0;
  w.xcnearrb = ((xc - xlinc <= P.ixrb + 1) && (xc + xrinc >= P.ixrb)) || xc < 0;
  return w;
}

// ---- csrcsc (svr/svrut4.m:1536-1608): CSC -> CSR transpose, 1-based -----------------
void csrcsc(int64_t n, const double* a, const int64_t* ja, const int64_t* ia, double* ao, int64_t* jao, int64_t* iao) {
  for (int64_t i = 0; i <= n; ++i) iao[i] = 0;
  for (int64_t i = 1; i <= n; ++i)
    for (int64_t k = ia[i - 1]; k <= ia[i] - 1; ++k) { int64_t j = ja[k - 1] + 1; iao[j - 1] = iao[j - 1] + 1; }
  iao[0] = 1;
  for (int64_t i = 1; i <= n; ++i) iao[i] = iao[i - 1] + iao[i];
  for (int64_t i = 1; i <= n; ++i)
    for (int64_t k = ia[i - 1]; k <= ia[i] - 1; ++k) {
      int64_t j = ja[k - 1];
      int64_t next = iao[j - 1];
      ao[next - 1] = a[k - 1];
      jao[next - 1] = i;
      iao[j - 1] = next + 1;
    }
  for (int64_t i = n; i >= 1; --i) iao[i] = iao[i - 1];
  iao[0] = 1;
}

// ---- all mutable state + the routines that touch it: one instance per worker thread ----------
struct Ora {
// ---- persistent "module" state ------------------------------------------------
// Compla / Gradients / Comflo / Conduc / Rhsides / Locflux groups of bbb/bbb.v,
// restricted to what this switch set touches.
V ne, nit, nm, nz2, ni, te, ti, ng, tg, up, pri, pre, pr, zeff, pg;
V gprx, gpry, gpix, gpiy, gpex, gtex, gtix, gpey, gtey, gtiy;
V niy0, niy1, nity0, nity1, ney0, ney1, priy0, priy1, tey0, tey1, tiy0, tiy1;
V ngy0, ngy1, tgy0, tgy1, pgy0, pgy1;
V loglambda, diffusivwrk, vy, frice, frici, ex, upi, uup, uu, upe, vex, vey;
V nuiz, nurc, nucx, nuix, psorbgg, psorgc, psorc, psordis, psorxrc, psorrgc, psorg, psor, psorxr, psorrg;
V snic, sniv, psori, smoc, smov, seec, seev, seic, seiv;
V conxg, conyg, floxg, floyg, fngx, fngy, resng;
V visx, visy, hcxe, hcxi, hcye, hcyi, hcxij, hcyij, eqp, w0, w1, w2, w3;
V fnix, fniy, resco, flox, floy, conx, cony, fmix, fmiy, resmo;
V floxe, floxi, floye, floyi, conxe, conxi, conye, conyi, feex, feey, feix, feiy, resee, resei;
V erliz, erlrc, eeli, vsoreec, vsoree, wvh, pwribkg;
V fniycbo, feeycbo, feiycbo;
// per-solve inputs
V dtuse, ylodt, suscal, sfscal, dtoptv;
// column range (ppp LocalJacBuilder analogue)
int64_t ivmin = 1, ivmax = 0;
std::string err;

std::vector<V*> all_planes() {
  return {&ne, &nit, &nm, &nz2, &ni, &te, &ti, &ng, &tg, &up, &pri, &pre, &pr, &zeff, &pg,
          &gprx, &gpry, &gpix, &gpiy, &gpex, &gtex, &gtix, &gpey, &gtey, &gtiy,
          &niy0, &niy1, &nity0, &nity1, &ney0, &ney1, &priy0, &priy1, &tey0, &tey1, &tiy0, &tiy1,
          &ngy0, &ngy1, &tgy0, &tgy1, &pgy0, &pgy1,
          &loglambda, &diffusivwrk, &vy, &frice, &frici, &ex, &upi, &uup, &uu, &upe, &vex, &vey,
          &nuiz, &nurc, &nucx, &nuix, &psorbgg, &psorgc, &psorc, &psordis, &psorxrc, &psorrgc, &psorg, &psor,
          &psorxr, &psorrg, &snic, &sniv, &psori, &smoc, &smov, &seec, &seev, &seic, &seiv,
          &conxg, &conyg, &floxg, &floyg, &fngx, &fngy, &resng,
          &visx, &visy, &hcxe, &hcxi, &hcye, &hcyi, &hcxij, &hcyij, &eqp, &w0, &w1, &w2, &w3,
          &fnix, &fniy, &resco, &flox, &floy, &conx, &cony, &fmix, &fmiy, &resmo,
          &floxe, &floxi, &floye, &floyi, &conxe, &conxi, &conye, &conyi, &feex, &feey, &feix, &feiy, &resee, &resei,
          &erliz, &erlrc, &eeli, &vsoreec, &vsoree, &wvh, &pwribkg};
}
// ---- convsr_vo (convert.m:158-375) -----------------------------------------------
int convsr_vo(int ixl, int iyl, const double* yl) {
  int is, ie, js, je;
  const int yinc = (int)P.yinc;
  if (ixl < 0 || yinc >= 6) { is = 0; ie = nx + 1; } else { is = ixl; ie = ixl; }
  if (iyl < 0 || yinc >= 6) { js = 0; je = ny + 1; } else { js = iyl; je = iyl; }
  if (ixl < 0 && iyl >= 0) { js = std::max(0, iyl - yinc); je = std::min(ny + 1, iyl + yinc); }
  int inegni = 0, inegng = 0;
  for (int iy = js; iy <= je; ++iy)
    for (int ix = is; ix <= ie; ++ix) {
      A(ne, ix, iy) = 0.; A(nit, ix, iy) = 0.; A(nm, ix, iy) = 0.; A(nz2, ix, iy) = 0.;
      A(ni, ix, iy) = yl[IDXN(ix, iy)] * P.n0;
      if (A(ni, ix, iy) < 0) inegni = 1;
      A(ne, ix, iy) = A(ne, ix, iy) + P.zi * A(ni, ix, iy);
      A(nit, ix, iy) = A(nit, ix, iy) + A(ni, ix, iy);
      A(nz2, ix, iy) = A(nz2, ix, iy) + A(ni, ix, iy) * (P.zi * P.zi);
      A(nm, ix, iy) = A(ni, ix, iy) * P.mi;
    }
  for (int iy = js; iy <= je; ++iy)
    for (int ix = is; ix <= ie; ++ix) {
      double ntemp = P.nnorm;  // isflxvar == 0
      A(te, ix, iy) = yl[IDXTE(ix, iy)] * P.ennorm / (1.5 * ntemp);
      A(te, ix, iy) = std::max(A(te, ix, iy), P.temin * P.ev);
      if (HASG) {  // convert.m:285
        A(ng, ix, iy) = yl[IDXG(ix, iy)] * P.n0g;
        if (A(ng, ix, iy) < 0) inegng = 1;
      }
      A(ti, ix, iy) = yl[IDXTI(ix, iy)] * P.ennorm / (1.5 * ntemp);
      A(ti, ix, iy) = std::max(A(ti, ix, iy), P.temin * P.ev);
    }
  if (inegni) { err = "***  ni is negative - calculation stopped"; return -3; }  // convert.m:318-322
  if (inegng) { err = "***  ng is negative - calculation stopped"; return -3; }  // convert.m:323-327
  for (int iy = js; iy <= je; ++iy)
    for (int ix = is; ix <= ie; ++ix) {
      int ix2 = std::max(0, IXM1(ix, iy));
      double t1 = P.mi * P.n0, t2 = P.mi * P.n0;  // isflxvar == 0
      A(up, ix2, iy) = yl[IDXU(ix2, iy)] * P.fnorm / t1;
      A(up, ix, iy) = yl[IDXU(ix, iy)] * P.fnorm / t2;
    }
  return 0;
}

// ---- convsr_aux (convert.m:379-875), orthogonal stencils fx0=1, others 0 ----------
inline double interp_log(const V& a, int ix, int iy, int k) {
  // interpni/interppri/interpng/interppg (convert.m:453-482) on an orthogonal mesh: fx0 = 1 and
  // fxm = fxp = fxmy = fxpy = 0 (geometry.m:849-870), so the four zero-weighted log terms add an exact 0
  // (all densities/pressures are positive and finite) and are not evaluated.
  return ue_exp(1. * ue_log(A(a, ix, iy + k)));
}
inline double interp_lin(const V& a, int ix, int iy, int k) {  // interpte/interpti/interptg (convert.m:422-442), same remark
  return 1. * A(a, ix, iy + k);
}
// list of ix visited by "do ix = ixm1(is,jrow), min(nx,ie), inc" (convert.m:583-584 etc.)
inline void xrange(int is, int ie, int jinc, int jstart, std::vector<int>& out) {
  out.clear();
  int d = ie - IXM1(ie, jinc);
  int inc = std::max(1, std::abs(d)); if (d < 0) inc = -inc;
  int first = IXM1(is, jstart), last = std::min(nx, ie);
  if (inc > 0) for (int ix = first; ix <= last; ix += inc) out.push_back(ix);
  else for (int ix = first; ix >= last; ix += inc) out.push_back(ix);
}
void convsr_aux(int ixl, int iyl) {
  int is, ie, js, je;
  const int yinc = (int)P.yinc;
  if (ixl < 0 || yinc >= 6) { is = 0; ie = nx + 1; } else { is = ixl; ie = ixl; }
  if (iyl < 0 || yinc >= 6) { js = 0; je = ny + 1; } else { js = iyl; je = iyl; }
  if (ixl < 0 && iyl >= 0) { js = std::max(0, iyl - yinc); je = std::min(ny + 1, iyl + yinc); }
  std::vector<int> xs;
  for (int iy = js; iy <= je; ++iy)
    for (int ix = is; ix <= ie; ++ix) {
      A(pr, ix, iy) = 0.; A(zeff, ix, iy) = 0.;
      A(pri, ix, iy) = A(ni, ix, iy) * A(ti, ix, iy);
      A(pr, ix, iy) = A(pr, ix, iy) + A(pri, ix, iy);
      A(zeff, ix, iy) = A(zeff, ix, iy) + (P.zi * P.zi) * A(ni, ix, iy);
    }
  for (int iy = js; iy <= je; ++iy)
    for (int ix = is; ix <= ie; ++ix) {
      A(pre, ix, iy) = A(ne, ix, iy) * A(te, ix, iy);
      A(pr, ix, iy) = A(pr, ix, iy) + A(pre, ix, iy);
      A(zeff, ix, iy) = A(zeff, ix, iy) / A(ne, ix, iy);
      if (P.istgcon > -1.e-20) A(tg, ix, iy) = (1 - P.istgcon) * P.rtg2ti * A(ti, ix, iy) + P.istgcon * P.tgas * P.ev;
      A(pg, ix, iy) = A(ng, ix, iy) * A(tg, ix, iy);
    }
  // x-gradients (convert.m:582-624, 736-749)
  for (int iy = js; iy <= je; ++iy) {
    xrange(is, ie, iy, iy, xs);
    for (int ix : xs) A(gprx, ix, iy) = 0.0;
  }
  for (int iy = std::max(js - 1, 0); iy <= std::min(ny, je); ++iy) {
    xrange(is, ie, js, js, xs);
    for (int ix : xs) { A(ney0, ix, iy) = 0.; A(ney1, ix, iy) = 0.; A(nity0, ix, iy) = 0.; A(nity1, ix, iy) = 0.; A(gpry, ix, iy) = 0.; }
    int ix = IXP1(ie, iy);
    A(ney0, ix, iy) = 0.; A(ney1, ix, iy) = 0.; A(nity0, ix, iy) = 0.; A(nity1, ix, iy) = 0.; A(gpry, ix, iy) = 0.;
  }
  for (int iy = js; iy <= je; ++iy) {
    xrange(is, ie, iy, iy, xs);
    for (int ix : xs) {
      int ix1 = IXP1(ix, iy);
      A(gpix, ix, iy) = (A(pri, ix1, iy) - A(pri, ix, iy)) * G(gxf, ix, iy);
      A(gprx, ix, iy) = A(gprx, ix, iy) + A(gpix, ix, iy);
    }
  }
  auto yface_ion = [&](int ix, int iy) {  // convert.m:631-647
    A(niy0, ix, iy) = interp_log(ni, ix, iy, 0);
    A(niy1, ix, iy) = interp_log(ni, ix, iy, 1);
    A(nity0, ix, iy) = A(nity0, ix, iy) + A(niy0, ix, iy);
    A(nity1, ix, iy) = A(nity1, ix, iy) + A(niy1, ix, iy);
    A(ney0, ix, iy) = A(ney0, ix, iy) + P.zi * A(niy0, ix, iy);
    A(ney1, ix, iy) = A(ney1, ix, iy) + P.zi * A(niy1, ix, iy);
    A(priy0, ix, iy) = interp_log(pri, ix, iy, 0);
    A(priy1, ix, iy) = interp_log(pri, ix, iy, 1);
    A(gpiy, ix, iy) = (A(priy1, ix, iy) - A(priy0, ix, iy)) / G(dynog, ix, iy);
    A(gpry, ix, iy) = A(gpry, ix, iy) + A(gpiy, ix, iy);
  };
  auto yface_t = [&](int ix, int iy) {  // convert.m:674-677
    A(tey0, ix, iy) = interp_lin(te, ix, iy, 0); A(tey1, ix, iy) = interp_lin(te, ix, iy, 1);
    A(tiy0, ix, iy) = interp_lin(ti, ix, iy, 0); A(tiy1, ix, iy) = interp_lin(ti, ix, iy, 1);
  };
  auto yface_g = [&](int ix, int iy) {  // convert.m:707-710
    A(ngy0, ix, iy) = interp_log(ng, ix, iy, 0); A(ngy1, ix, iy) = interp_log(ng, ix, iy, 1);
    A(tgy0, ix, iy) = interp_lin(tg, ix, iy, 0); A(tgy1, ix, iy) = interp_lin(tg, ix, iy, 1);
  };
  auto yface_pg = [&](int ix, int iy) {  // convert.m:726-727
    A(pgy0, ix, iy) = interp_log(pg, ix, iy, 0); A(pgy1, ix, iy) = interp_log(pg, ix, iy, 1);
  };
  const int jlo = std::max(js - 1, 0), jhi = std::min(je, ny);
  for (int iy = jlo; iy <= jhi; ++iy) { xrange(is, ie, js, js, xs); for (int ix : xs) yface_ion(ix, iy); yface_ion(IXP1(ie, iy), iy); }
  for (int iy = jlo; iy <= jhi; ++iy) { xrange(is, ie, js, js, xs); for (int ix : xs) yface_t(ix, iy); yface_t(IXP1(ie, iy), iy); }
  for (int iy = jlo; iy <= jhi; ++iy) { xrange(is, ie, js, js, xs); for (int ix : xs) yface_g(ix, iy); yface_g(IXP1(ie, iy), iy); }
  if (P.ineudif == 2)
    for (int iy = jlo; iy <= jhi; ++iy) { xrange(is, ie, js, js, xs); for (int ix : xs) yface_pg(ix, iy); yface_pg(IXP1(ie, iy), iy); }
  for (int iy = js; iy <= je; ++iy) {  // convert.m:736-749
    xrange(is, ie, iy, iy, xs);
    for (int ix : xs) {
      int ix1 = IXP1(ix, iy);
      A(gpex, ix, iy) = (A(pre, ix1, iy) - A(pre, ix, iy)) * G(gxf, ix, iy);
      A(gtex, ix, iy) = (A(te, ix1, iy) - A(te, ix, iy)) * G(gxf, ix, iy);
      A(gtix, ix, iy) = (A(ti, ix1, iy) - A(ti, ix, iy)) * G(gxf, ix, iy);
      A(gprx, ix, iy) = A(gprx, ix, iy) + A(gpex, ix, iy);
    }
  }
  auto ygrad = [&](int ix, int iy) {  // convert.m:768-774
    A(gpey, ix, iy) = (A(ney1, ix, iy) * A(tey1, ix, iy) - A(ney0, ix, iy) * A(tey0, ix, iy)) / G(dynog, ix, iy);
    A(gtey, ix, iy) = (A(tey1, ix, iy) - A(tey0, ix, iy)) / G(dynog, ix, iy);
    A(gtiy, ix, iy) = (A(tiy1, ix, iy) - A(tiy0, ix, iy)) / G(dynog, ix, iy);
    A(gpry, ix, iy) = A(gpry, ix, iy) + A(gpey, ix, iy);
  };
  for (int iy = jlo; iy <= std::min(ny, je); ++iy) { xrange(is, ie, js, js, xs); for (int ix : xs) ygrad(ix, iy); ygrad(IXP1(ie, iy), iy); }
  // vertex quantities (convert.m:791-868) feed only the drift velocities, whose coefficients are zero here.
}

// ---- fd2tra (oderhs.m:7-534), orthogonal mesh, meth in {2,3} ----------------------
inline double upwind(double f, double p1, double p2) { return std::max(f, 0.0) * p1 + std::min(f, 0.0) * p2; }  // :81
void fd2tra(const Win& w, const V& flx, const V& fly, const V& difx, const V& dify, const V& phi, V& trax, V& tray, int pos, int meth) {
  int posx = pos % 10, posy = pos / 10, methx = meth % 10, methy = meth / 10;
  for (int iy = w.j4; iy <= w.j8; ++iy)
    for (int ix = w.i1; ix <= w.i5; ++ix) {
      int ix1 = IXP1(ix, iy);
      int ix2 = ix * (1 - posx) + ix1 * posx;
      if (methx == 2)  // oderhs.m:137-146
        A(trax, ix2, iy) = A(flx, ix2, iy) * (A(phi, ix1, iy) + A(phi, ix, iy)) / 2. - A(difx, ix2, iy) * (A(phi, ix1, iy) - A(phi, ix, iy));
      else  // methx == 3, oderhs.m:152-161
        A(trax, ix2, iy) = upwind(A(flx, ix2, iy), A(phi, ix, iy), A(phi, ix1, iy)) - A(difx, ix2, iy) * (A(phi, ix1, iy) - A(phi, ix, iy));
    }
  for (int iy = w.j1; iy <= w.j5 - posy; ++iy)
    for (int ix = w.i4; ix <= w.i8; ++ix) {
      if (methy == 2)  // oderhs.m:268-275
        A(tray, ix, iy + posy) = A(fly, ix, iy + posy) * (A(phi, ix, iy + 1) + A(phi, ix, iy)) / 2. - A(dify, ix, iy + posy) * (A(phi, ix, iy + 1) - A(phi, ix, iy));
      else  // oderhs.m:281-288
        A(tray, ix, iy + posy) = upwind(A(fly, ix, iy + posy), A(phi, ix, iy), A(phi, ix, iy + 1)) - A(dify, ix, iy + posy) * (A(phi, ix, iy + 1) - A(phi, ix, iy));
    }
}

// ---- neudifpg (oderhs.m:6058-6648) ------------------------------------------------
void neudifpg(const Win& w) {
  const int methgx = (int)(P.methg % 10), methgy = (int)(P.methg / 10);
  for (int iy = w.j4; iy <= w.j8; ++iy) {
    for (int ix = w.i1; ix <= w.i5; ++ix) {  // oderhs.m:6126-6225
      int ix2 = IXP1(ix, iy);
      double ngxface = 0.5 * (A(ng, ix, iy) + A(ng, ix2, iy));
      double t0 = std::max(A(tg, ix, iy), P.temin * P.ev), t1 = std::max(A(tg, ix2, iy), P.temin * P.ev);
      double vtn = std::sqrt(t0 / P.mg), vtnp = std::sqrt(t1 / P.mg);
      double nu1 = A(nuix, ix, iy) + vtn / P.lgmax, nu2 = A(nuix, ix2, iy) + vtnp / P.lgmax;
      double tgf = 0.5 * (A(tg, ix, iy) + A(tg, ix2, iy));
      double flalfgx_adj = P.flalfgxa[ix] * (1. + powi(P.cflbg * P.ngbackg / ngxface, P.inflbg));
      double qfl = flalfgx_adj * G(sx, ix, iy) * (vtn + vtnp) * P.rt8opi * (A(ng, ix, iy) * G(gx, ix, iy) + A(ng, ix2, iy) * G(gx, ix2, iy)) /
                   (8 * (G(gx, ix, iy) + G(gx, ix2, iy)));
      double csh = (1 - P.isgasdc) * P.cdifg * G(sx, ix, iy) * G(gxf, ix, iy) * (1 / P.mg) * ave(1. / nu1, 1. / nu2) +
                   P.isgasdc * G(sx, ix, iy) * G(gxf, ix, iy) * P.difcng / tgf +
                   (P.rld2dxg * P.rld2dxg) * G(sx, ix, iy) * (1 / G(gxf, ix, iy)) * 0.5 * (A(nuiz, ix, iy) + A(nuiz, ix2, iy)) / tgf;
      double qtgf = P.alftng * P.fgtdx[ix] * G(sx, ix, iy) * ave(G(gx, ix, iy) / nu1, G(gx, ix2, iy) / nu2) * (vtn * vtn - vtnp * vtnp);
      double vygtan = 0.;
      qtgf = qtgf - vygtan * G(sx, ix, iy);
      double nconv = 2.0 * (A(ng, ix, iy) * A(ng, ix2, iy)) / (A(ng, ix, iy) + A(ng, ix2, iy));
      if (methgx != 2) nconv = A(ng, ix, iy) * 0.5 * (1 + sgn(1., qtgf)) + A(ng, ix2, iy) * 0.5 * (1 - sgn(1., qtgf));
      double qsh = csh * (A(pg, ix, iy) - A(pg, ix2, iy)) + qtgf * nconv;
      double qr = std::fabs(qsh / qfl);
      if (ix == P.ixlb || ix == P.ixrb) { qr = P.gcfacgx * qr; qtgf = P.gcfacgx * qtgf; }
      A(conxg, ix, iy) = csh / ue_pow(1 + ue_pow(qr, P.flgamg), 1 / P.flgamg);
      if (P.isdifxg_aug == 1) A(conxg, ix, iy) = csh * (1 + qr);
      A(floxg, ix, iy) = (qtgf / tgf) / ue_pow(1 + ue_pow(qr, P.flgamg), 1 / P.flgamg);
      A(floxg, ix, iy) = A(floxg, ix, iy) + P.cngflox * G(sx, ix, iy) * A(uu, ix, iy) / tgf;
    }
    A(conxg, nx + 1, iy) = 0;
  }
  for (int iy = w.j1; iy <= w.j5; ++iy)
    for (int ix = w.i4; ix <= w.i8; ++ix) {  // oderhs.m:6239-6328
      double ngyface = 0.5 * (A(ng, ix, iy) + A(ng, ix, iy + 1));
      double t0 = std::max(A(tg, ix, iy), P.tgmin * P.ev), t1 = std::max(A(tg, ix, iy + 1), P.tgmin * P.ev);
      double vtn = std::sqrt(t0 / P.mg), vtnp = std::sqrt(t1 / P.mg);
      double nu1 = A(nuix, ix, iy) + vtn / P.lgmax, nu2 = A(nuix, ix, iy + 1) + vtnp / P.lgmax;
      double tgf = 0.5 * (A(tg, ix, iy) + A(tg, ix, iy + 1));
      double flalfgy_adj = P.flalfgya[iy] * (1. + powi(P.cflbg * P.ngbackg / ngyface, P.inflbg));
      double qfl = flalfgy_adj * G(sy, ix, iy) * (vtn + vtnp) * P.rt8opi * (A(ngy0, ix, iy) * G(gy, ix, iy) + A(ngy1, ix, iy) * G(gy, ix, iy + 1)) /
                   (8 * (G(gy, ix, iy) + G(gy, ix, iy + 1)));
      if (iy == 0) qfl = flalfgy_adj * G(sy, ix, iy) * (vtn + vtnp) * P.rt8opi * (A(ngy0, ix, iy) + A(ngy1, ix, iy)) / 8.;
      double csh = (1 - P.isgasdc) * (P.cdifg * G(sy, ix, iy) / G(dynog, ix, iy)) * (1 / P.mg) * ave(1. / nu1, 1. / nu2) +
                   P.isgasdc * G(sy, ix, iy) * P.difcng / (G(dynog, ix, iy) * tgf) +
                   (P.rld2dyg * P.rld2dyg) * G(sy, ix, iy) * G(dynog, ix, iy) * 0.5 * (A(nuiz, ix, iy) + A(nuiz, ix, iy + 1)) / tgf;
      double qtgf = P.alftng * P.fgtdy[iy] * G(sy, ix, iy) * ave(G(gy, ix, iy) / nu1, G(gy, ix, iy + 1) / nu2) * (vtn * vtn - vtnp * vtnp);
      double nconv = 2.0 * (A(ngy0, ix, iy) * A(ngy1, ix, iy)) / (A(ngy0, ix, iy) + A(ngy1, ix, iy));
      if (methgy != 2) nconv = A(ngy0, ix, iy) * 0.5 * (1 + sgn(1., qtgf)) + A(ngy1, ix, iy) * 0.5 * (1 - sgn(1., qtgf));
      double qsh = csh * (A(pgy0, ix, iy) - A(pgy1, ix, iy)) + qtgf * nconv;
      double qr = std::fabs(qsh / qfl);
      if (iy == 0) { qr = P.gcfacgy * qr; qtgf = P.gcfacgy * qtgf; }
      if (iy == ny) { qr = P.gcfacgy * qr; qtgf = P.gcfacgy * qtgf; }
      A(conyg, ix, iy) = csh / ue_pow(1 + ue_pow(qr, P.flgamg), 1 / P.flgamg);
      if (P.isdifyg_aug == 1) A(conyg, ix, iy) = csh * (1 + qr);
      A(floyg, ix, iy) = (qtgf / tgf) / ue_pow(1 + ue_pow(qr, P.flgamg), 1 / P.flgamg);
      A(floyg, ix, iy) = A(floyg, ix, iy) + P.cngfloy * G(sy, ix, iy) * A(vy, ix, iy) / tgf;
    }
  fd2tra(w, floxg, floyg, conxg, conyg, pg, fngx, fngy, 0, (int)P.methg);  // oderhs.m:6340
  for (int iy = w.j2; iy <= w.j5; ++iy)
    for (int ix = w.i2; ix <= w.i5; ++ix) {  // oderhs.m:6574-6610 (psorcxg, volpsorg, psgov_use are zero fields)
      int ix1 = IXM1(ix, iy);
      A(resng, ix, iy) = P.cngsor * (A(psorg, ix, iy) + 0. + A(psorrg, ix, iy)) + 0. + 0. * G(vol, ix, iy);
      A(resng, ix, iy) = A(resng, ix, iy) - P.cfneutdiv * P.cfneutdiv_fng * ((A(fngx, ix, iy) - A(fngx, ix1, iy)) + P.fluxfacy * (A(fngy, ix, iy) - A(fngy, ix, iy - 1)));
    }
}

// ---- bouncon (boundary.m:4-3700): guard-cell equations ----------------------------
int bouncon(const Win& w, const double* yl, double* yldot) {
  (void)yl;
  const double ev = P.ev, pi = P.pi;
  const int ixlb = (int)P.ixlb, ixrb = (int)P.ixrb;
  const int ix_fl_bc = std::min((int)P.ixpt2, nx);
  // ===== iy = 0 boundary (boundary.m:102-1123) =====
  if (w.j3 <= 0) {  // isextrnpf = isextrtpf = isextrngc = 0
    for (int ix = w.i4; ix <= w.i8; ++ix) {  // density, boundary.m:122-287
      int64_t iv1 = IDXN(ix, 0);
      if (P.isixcore[ix] == 1) {
        if (P.isnicore == 1) yldot[iv1] = P.nurlxn * (P.ncore - A(ni, ix, 0)) / P.n0;
        else  // isnicore == 0, boundary.m:211-215 (fniycbo = 0: drift coefficients zero)
          yldot[iv1] = -P.nurlxn * (P.qe * (A(fniy, ix, 0) - fniycbo[ix]) / G(sy, ix, 0) - P.curcore * G(gyf, ix, 0) / P.sygytotc) / (P.qe * P.vpnorm * P.n0);
      } else if (P.isnwconiix[ix] == 0) {  // boundary.m:259-265
        yldot[iv1] = P.nurlxn * ((1 - P.ifluxni) * (A(niy1, ix, 0) - A(niy0, ix, 0)) -
                                 P.ifluxni * (A(fniy, ix, 0) / (G(sy, ix, 0) * P.vpnorm) - 0.001 * A(ni, ix, 1) * A(vy, ix, 0) / P.vpnorm)) / P.n0;
      } else if (P.isnwconiix[ix] == 1) {  // fixed wall density, boundary.m:267-270
        yldot[iv1] = P.nurlxn * (P.nwalli[ix] - A(ni, ix, 0)) / P.n0;
      } else if (P.isnwconiix[ix] == 2) {  // extrapolation, boundary.m:271-277
        double nbound = A(ni, ix, 1) - G(gyf, ix, 1) * (A(ni, ix, 2) - A(ni, ix, 1)) / G(gyf, ix, 0);
        nbound = 1.2 * nbound / (1 + 0.5 * ue_exp(-2 * (nbound / A(ni, ix, 1) - 1))) + 0.2 * A(ni, ix, 1);
        yldot[iv1] = P.nurlxn * (nbound - A(ni, ix, 0)) / P.n0;
      } else if (P.isnwconiix[ix] == 3) {  // specified gradient length, boundary.m:278-282
        yldot[iv1] = -P.nurlxn * (A(niy0, ix, 0) - A(niy1, ix, 0) * (2 * G(gyf, ix, 0) * P.lynipf[ix] - 1) / (2 * G(gyf, ix, 0) * P.lynipf[ix] + 1) - P.nwimin) / P.n0;
      }
    }
    // corners, boundary.m:290-303
    if (P.isfixlb != 2) yldot[IDXN(ixlb, 0)] = P.nurlxn * (ave(A(ni, ixlb, 1), A(ni, ixlb + 1, 0)) - A(ni, ixlb, 0)) / P.n0;
    yldot[IDXN(ixrb + 1, 0)] = P.nurlxn * (ave(A(ni, ixrb + 1, 1), A(ni, ixrb, 0)) - A(ni, ixrb + 1, 0)) / P.n0;
    for (int ix = w.i4; ix <= w.i8; ++ix) {  // parallel velocity, boundary.m:308-383
      int64_t iv2 = IDXU(ix, 0);
      if (P.isixcore[ix] == 1) {
        if (P.isupcore == 0) yldot[iv2] = P.nurlxu * (P.upcore - A(up, ix, 0)) / P.vpnorm;
        else if (P.isupcore == 1) yldot[iv2] = P.nurlxu * (A(up, ix, 1) - A(up, ix, 0)) / P.vpnorm;
        else if (P.isupcore == 2)  // d2(up)/dy2 = 0, boundary.m:323-326
          yldot[iv2] = P.nurlxu * ((A(up, ix, 1) - A(up, ix, 0)) * G(gy, ix, 1) - (A(up, ix, 2) - A(up, ix, 1)) * G(gy, ix, 2)) / (G(gy, ix, 1) * P.vpnorm);
        else  // == 3: no radial momentum flux, boundary.m:327-329
          yldot[iv2] = -P.nurlxu * A(fmiy, ix, 0) / (P.vpnorm * G(sy, ix, 0) * P.fnorm);
      } else if (P.isupwiix[ix] == 2) {
        yldot[iv2] = P.nurlxu * A(nm, ix, 0) / P.fnorm * (A(up, ix, 1) - A(up, ix, 0));
      } else {
        yldot[iv2] = P.nurlxu * A(nm, ix, 0) / P.fnorm * (0. - A(up, ix, 0));
      }
    }
    for (int ix = w.i4; ix <= w.i8; ++ix) {  // Te, Ti, boundary.m:524-628
      int64_t iv1 = IDXTE(ix, 0), iv2 = IDXTI(ix, 0);
      if (P.isixcore[ix] == 1) {
        yldot[iv1] = P.nurlxe * (P.tcoree * ev - A(te, ix, 0)) * 1.5 * A(ne, ix, 0) / P.ennorm;
        yldot[iv2] = P.nurlxi * (P.tcorei * ev - A(ti, ix, 0)) * 1.5 * A(ne, ix, 0) / P.ennorm;
        if (P.iflcore == 1) {  // integrated core power, boundary.m:529-545, 576-593
          yldot[iv1] = -P.nurlxe * (A(te, ix, 0) - A(te, IXP1(ix, 0), 0)) * P.n0 / P.ennorm;
          yldot[iv2] = -P.nurlxi * (A(ti, ix, 0) - A(ti, IXP1(ix, 0), 0)) * P.n0 / P.ennorm;
          if (ix == ix_fl_bc) {
            int ii = std::max(0, (int)P.ixpt1 + 1);
            double feeytotc = A(feey, ii, 0) - feeycbo[ii], feiytotc = A(feiy, ii, 0) - feiycbo[ii];
            do { ii = IXP1(ii, 0); feeytotc = feeytotc + A(feey, ii, 0) - feeycbo[ii]; } while (ii != ix_fl_bc);
            ii = std::max(0, (int)P.ixpt1 + 1);
            do { ii = IXP1(ii, 0); feiytotc = feiytotc + A(feiy, ii, 0) - feiycbo[ii]; } while (ii != ix_fl_bc);
            yldot[iv1] = -P.nurlxe * (feeytotc - P.pcoree) / (P.vpnorm * P.ennorm);
            yldot[iv2] = -P.nurlxi * (feiytotc - P.pcorei) / (P.vpnorm * P.ennorm);
          }
        } else if (P.iflcore == -1) {  // zero radial temperature gradient, boundary.m:546-548, 594-596
          yldot[iv1] = -P.nurlxe * (A(te, ix, 0) - A(te, ix, 1)) * P.n0 / P.ennorm;
          yldot[iv2] = -P.nurlxi * (A(ti, ix, 0) - A(ti, ix, 1)) * P.n0 / P.ennorm;
        }
      } else {
        // boundary.m:550-565, 597-612: 0 zero flux, 1 fixed, 2 extrapolation, 3 specified gradient length
        if (P.istepfcix[ix] == 0) yldot[iv1] = -P.nurlxe * (A(feey, ix, 0) / (P.n0 * P.vpnorm * G(sy, ix, 0))) / (P.temp0 * ev);
        else if (P.istepfcix[ix] == 1) yldot[iv1] = P.nurlxe * (P.tewalli[ix] * ev - A(te, ix, 0)) / (P.temp0 * ev);
        else if (P.istepfcix[ix] == 2) {
          double tbound = A(te, ix, 1) - G(gyf, ix, 1) * (A(te, ix, 2) - A(te, ix, 1)) / G(gyf, ix, 0);
          tbound = std::max(tbound, P.tbmin * ev);
          yldot[iv1] = P.nurlxe * (tbound - A(te, ix, 0)) / (P.temp0 * ev);
        } else yldot[iv1] = P.nurlxe * ((A(te, ix, 1) - A(te, ix, 0)) - 0.5 * (A(te, ix, 1) + A(te, ix, 0)) / (G(gyf, ix, 0) * P.lytepf[ix])) / (P.temp0 * ev);
        if (P.istipfcix[ix] == 0) yldot[iv2] = -P.nurlxi * (A(feiy, ix, 0) / (P.n0 * P.vpnorm * G(sy, ix, 0))) / (P.temp0 * ev);
        else if (P.istipfcix[ix] == 1) yldot[iv2] = P.nurlxi * (P.tiwalli[ix] * ev - A(ti, ix, 0)) / (P.temp0 * ev);
        else if (P.istipfcix[ix] == 2) {
          double tbound = A(ti, ix, 1) - G(gyf, ix, 1) * (A(ti, ix, 2) - A(ti, ix, 1)) / G(gyf, ix, 0);
          tbound = std::max(tbound, P.tbmin * ev);
          yldot[iv2] = P.nurlxi * (tbound - A(ti, ix, 0)) / (P.temp0 * ev);
        } else yldot[iv2] = P.nurlxi * ((A(ti, ix, 1) - A(ti, ix, 0)) - 0.5 * (A(ti, ix, 1) + A(ti, ix, 0)) / (G(gyf, ix, 0) * P.lytipf[ix])) / (P.temp0 * ev);
      }
    }
    for (int ix = w.i4; ix <= w.i8 && HASG; ++ix) {  // neutral density, boundary.m:632-767
      int64_t iv = IDXG(ix, 0);
      double t0 = std::max(P.cdifg * A(tg, ix, 0), P.tgmin * ev);
      double vyn = 0.25 * std::sqrt(8 * t0 / (pi * P.mg));
      double nharmave = 2. * (A(ng, ix, 0) * A(ng, ix, 1)) / (A(ng, ix, 0) + A(ng, ix, 1));
      if (P.isixcore[ix] == 1) {  // boundary.m:651-681
        if (P.isngcore == 0) {
          double fng_alb = (1 - P.albedoc) * nharmave * vyn * G(sy, ix, 0);
          yldot[iv] = -P.nurlxg * (A(fngy, ix, 0) + fng_alb) / (vyn * G(sy, ix, 0) * P.n0g);
        } else if (P.isngcore == 1) yldot[iv] = P.nurlxg * (P.ngcore - A(ng, ix, 0)) / P.n0g;
        else if (P.isngcore == 2) {
          double lengg = std::sqrt(A(tg, ix, 0) / (P.mg * (A(nuix, ix, 0) * A(nuiz, ix, 0))));
          yldot[iv] = P.nurlxn * ((A(ng, ix, 1) - A(ng, ix, 0)) - 0.5 * (A(ng, ix, 1) + A(ng, ix, 0)) / (G(gyf, ix, 0) * lengg)) / P.n0g;
        } else if (P.isngcore == 3) {
          double nbound = A(ng, ix, 1) - G(gyf, ix, 1) * (A(ng, ix, 2) - A(ng, ix, 1)) / G(gyf, ix, 0);
          nbound = 1.2 * nbound / (1 + 0.5 * ue_exp(-2 * (nbound / A(ng, ix, 1) - 1))) + 0.2 * A(ng, ix, 1);
          yldot[iv] = P.nurlxn * (nbound - A(ng, ix, 0)) / P.n0g;
        } else yldot[iv] = P.nurlxn * (A(ng, ix, 1) - A(ng, ix, 0)) / P.n0g;  // == 4
      } else {  // chemsputi = 0, no ion sputtering, matwalli = 0, fngysi = fngyi_use = 0
        double fng_chem = 0., sputflxpf = 0.;
        double fng_alb = (1 - P.albedoi[ix]) * nharmave * vyn * G(sy, ix, 0);
        yldot[iv] = -P.nurlxg * (A(fngy, ix, 0) + fng_alb - fng_chem + sputflxpf) / (vyn * G(sy, ix, 0) * P.n0g);
        if (P.matwalli[ix] > 0) {  // recycling wall, boundary.m:733-760
          if (P.recycwit[ix] > 0.) {
            double fniy_recy = P.fac2sp * A(fniy, ix, 0);
            if (P.isrefluxclip == 1) fniy_recy = std::min(fniy_recy, 0.);
            yldot[iv] = -P.nurlxg * (A(fngy, ix, 0) + fniy_recy * P.recycwit[ix] - P.fngyi_use[ix] - P.fngysi[ix] + fng_alb - fng_chem + sputflxpf) /
                        (vyn * P.n0g * G(sy, ix, 0));
          } else if (P.recycwit[ix] < -1) yldot[iv] = P.nurlxg * (P.ngbackg - A(ng, ix, 0)) / P.n0g;
          else {
            nharmave = 2. * (A(ng, ix, 0) * A(ng, ix, 1)) / (A(ng, ix, 0) + A(ng, ix, 1));
            yldot[iv] = -P.nurlxg * (A(fngy, ix, 0) + (1 + P.recycwit[ix]) * nharmave * vyn * G(sy, ix, 0)) / (vyn * P.n0g * G(sy, ix, 0));
          }
        }
      }
    }
    if (w.xcnearlb || w.openbox) {  // boundary.m:897-938
      yldot[IDXU(ixlb, 0)] = -P.nurlxu * (A(up, ixlb, 0) - 0.5 * (A(up, ixlb, 1) + A(up, ixlb + 1, 0))) / P.vpnorm;
      yldot[IDXTE(ixlb, 0)] = P.nurlxe * (0.5 * (A(te, ixlb + 1, 0) + A(te, ixlb, 1)) - A(te, ixlb, 0)) / (P.temp0 * ev);
      yldot[IDXTI(ixlb, 0)] = P.nurlxi * (0.5 * (A(ti, ixlb + 1, 0) + A(ti, ixlb, 1)) - A(ti, ixlb, 0)) / (P.temp0 * ev);
      if (HASG) yldot[IDXG(ixlb, 0)] = P.nurlxg * (A(ng, ixlb + 1, 0) - A(ng, ixlb, 0)) / P.n0g;
    }
    if (w.xcnearrb || w.openbox) {  // boundary.m:939-983
      yldot[IDXU(ixrb, 0)] = -P.nurlxu * (A(up, ixrb, 0) - 0.5 * (A(up, ixrb - 1, 0) + A(up, ixrb, 1))) / P.vpnorm;
      yldot[IDXU(ixrb + 1, 0)] = -P.nurlxu * (A(up, ixrb + 1, 0) - A(up, ixrb, 0)) / P.vpnorm;
      yldot[IDXTE(ixrb + 1, 0)] = P.nurlxe * (0.5 * (A(te, ixrb + 1, 1) + A(te, ixrb, 0)) - A(te, ixrb + 1, 0)) / (P.temp0 * ev);
      yldot[IDXTI(ixrb + 1, 0)] = P.nurlxi * (0.5 * (A(ti, ixrb + 1, 1) + A(ti, ixrb, 0)) - A(ti, ixrb + 1, 0)) / (P.temp0 * ev);
      if (HASG) yldot[IDXG(ixrb + 1, 0)] = P.nurlxg * (A(ng, ixrb, 0) - A(ng, ixrb + 1, 0)) / P.n0g;
    }
  }
  // ===== iy = ny+1 boundary (boundary.m:1125-1653) =====
  if (w.j7 >= (ny + 1)) {  // isextrnw = isextrtw = 0
    for (int ix = w.i4; ix <= w.i8; ++ix) {  // boundary.m:1133-1207
      int64_t iv1 = IDXN(ix, ny + 1);
      if (P.isnwconoix[ix] == 0)
        yldot[iv1] = P.nurlxn * ((1 - P.ifluxni) * (A(niy0, ix, ny) - A(niy1, ix, ny)) +
                                 P.ifluxni * (A(fniy, ix, ny) / (G(sy, ix, ny) * P.vpnorm) - 0.001 * A(ni, ix, ny) * A(vy, ix, ny) / P.vpnorm)) / P.n0;
      else if (P.isnwconoix[ix] == 1) yldot[iv1] = P.nurlxn * (P.nwallo[ix] - A(ni, ix, ny + 1)) / P.n0;
      else if (P.isnwconoix[ix] == 2) {  // extrapolation, boundary.m:1192-1198
        double nbound = A(ni, ix, ny) + G(gyf, ix, ny - 1) * (A(ni, ix, ny) - A(ni, ix, ny - 1)) / G(gyf, ix, ny);
        nbound = 1.2 * nbound / (1 + 0.5 * ue_exp(-2 * (nbound / A(ni, ix, ny) - 1))) + 0.2 * A(ni, ix, ny);
        yldot[iv1] = P.nurlxn * (nbound - A(ni, ix, ny + 1)) / P.n0;
      }
      else  // == 3, specified gradient length
        yldot[iv1] = -P.nurlxn * (A(niy1, ix, ny) - A(niy0, ix, ny) * (2 * G(gyf, ix, ny) * P.lyniwc[ix] - 1) / (2 * G(gyf, ix, ny) * P.lyniwc[ix] + 1) - P.nwomin) / P.n0;
    }
    yldot[IDXN(ixlb, ny + 1)] = P.nurlxn * (ave(A(ni, ixlb, ny), A(ni, ixlb + 1, ny + 1)) - A(ni, ixlb, ny + 1)) / P.n0;       // :1209-1217
    yldot[IDXN(ixrb + 1, ny + 1)] = P.nurlxn * (ave(A(ni, ixrb + 1, ny), A(ni, ixrb, ny + 1)) - A(ni, ixrb + 1, ny + 1)) / P.n0;  // :1218-1226
    for (int ix = w.i4; ix <= w.i8; ++ix) {  // boundary.m:1231-1252
      int64_t iv2 = IDXU(ix, ny + 1);
      if (P.isupwoix[ix] == 2) yldot[iv2] = P.nurlxu * A(nm, ix, ny) / P.fnorm * (A(up, ix, ny) - A(up, ix, ny + 1));
      else yldot[iv2] = P.nurlxu * A(nm, ix, ny) / P.fnorm * (0. - A(up, ix, ny + 1));
    }
    for (int ix = w.i4; ix <= w.i8; ++ix) {  // boundary.m:1311-1362
      int64_t iv1 = IDXTE(ix, ny + 1), iv2 = IDXTI(ix, ny + 1);
      if (P.istewcix[ix] == 0) yldot[iv1] = P.nurlxe * (A(feey, ix, ny) / (P.n0 * P.vpnorm * G(sy, ix, ny))) / (P.temp0 * ev);
      else if (P.istewcix[ix] == 1) yldot[iv1] = P.nurlxe * (P.tewallo[ix] * ev - A(te, ix, ny + 1)) / (P.temp0 * ev);
      else if (P.istewcix[ix] == 2) {
        double tbound = A(te, ix, ny) + G(gyf, ix, ny - 1) * (A(te, ix, ny) - A(te, ix, ny - 1)) / G(gyf, ix, ny);
        tbound = std::max(tbound, P.tbmin * ev);
        yldot[iv1] = P.nurlxe * (tbound - A(te, ix, ny + 1)) / (P.temp0 * ev);
      } else yldot[iv1] = P.nurlxe * ((A(te, ix, ny) - A(te, ix, ny + 1)) - 0.5 * (A(te, ix, ny) + A(te, ix, ny + 1)) / (G(gyf, ix, ny) * P.lytewc[ix])) / (P.temp0 * ev);
      if (P.istiwcix[ix] == 0) yldot[iv2] = P.nurlxi * (A(feiy, ix, ny) / (P.n0 * P.vpnorm * G(sy, ix, ny))) / (P.temp0 * ev);
      else if (P.istiwcix[ix] == 1) yldot[iv2] = P.nurlxi * (P.tiwallo[ix] * ev - A(ti, ix, ny + 1)) / (P.temp0 * ev);
      else if (P.istiwcix[ix] == 2) {
        double tbound = A(ti, ix, ny) + G(gyf, ix, ny - 1) * (A(ti, ix, ny) - A(ti, ix, ny - 1)) / G(gyf, ix, ny);
        tbound = std::max(tbound, P.tbmin * ev);
        yldot[iv2] = P.nurlxi * (tbound - A(ti, ix, ny + 1)) / (P.temp0 * ev);
      } else yldot[iv2] = P.nurlxi * ((A(ti, ix, ny) - A(ti, ix, ny + 1)) - 0.5 * (A(ti, ix, ny) + A(ti, ix, ny + 1)) / (G(gyf, ix, ny) * P.lytiwc[ix])) / (P.temp0 * ev);
    }
    for (int ix = w.i4; ix <= w.i8 && HASG; ++ix) {  // boundary.m:1366-1462
      int64_t iv = IDXG(ix, ny + 1);
      double t0 = std::max(P.cdifg * A(tg, ix, ny + 1), P.tgmin * ev);
      double vyn = 0.25 * std::sqrt(8 * t0 / (pi * P.mg));
      double fng_chem = 0., sputflxw = 0.;
      double nharmave = 2. * (A(ng, ix, ny) * A(ng, ix, ny + 1)) / (A(ng, ix, ny) + A(ng, ix, ny + 1));
      double fng_alb = (1 - P.albedoo[ix]) * nharmave * vyn * G(sy, ix, ny);
      yldot[iv] = P.nurlxg * (A(fngy, ix, ny) - fng_alb + fng_chem + sputflxw) / (vyn * G(sy, ix, ny) * P.n0g);
      if (P.matwallo[ix] > 0) {  // recycling wall, boundary.m:1424-1452
        if (P.recycwot[ix] > 0.) {
          double fniy_recy = P.fac2sp * A(fniy, ix, ny);
          if (P.isrefluxclip == 1) fniy_recy = std::max(fniy_recy, 0.);
          yldot[iv] = P.nurlxg * (A(fngy, ix, ny) + fniy_recy * P.recycwot[ix] + P.fngyso[ix] + P.fngyo_use[ix] - fng_alb + fng_chem + sputflxw) /
                      (vyn * P.n0g * G(sy, ix, ny));
        } else if (P.recycwot[ix] < -1) yldot[iv] = P.nurlxg * (P.ngbackg - A(ng, ix, ny + 1)) / P.n0g;
        else {
          nharmave = 2. * (A(ng, ix, ny) * A(ng, ix, ny + 1)) / (A(ng, ix, ny) + A(ng, ix, ny + 1));
          yldot[iv] = P.nurlxg * (A(fngy, ix, ny) - (1 + P.recycwot[ix]) * nharmave * vyn * G(sy, ix, ny)) / (vyn * P.n0g * G(sy, ix, ny));
        }
      }
    }
    if (w.xcnearlb || w.openbox) {  // boundary.m:1543-1583
      yldot[IDXU(ixlb, ny + 1)] = -P.nurlxu * (A(up, ixlb, ny + 1) - 0.5 * (A(up, ixlb, ny) + A(up, ixlb + 1, ny + 1))) / P.vpnorm;
      yldot[IDXTE(ixlb, ny + 1)] = P.nurlxe * (0.5 * (A(te, ixlb + 1, ny + 1) + A(te, ixlb, ny)) - A(te, ixlb, ny + 1)) / (P.temp0 * ev);
      yldot[IDXTI(ixlb, ny + 1)] = P.nurlxi * (0.5 * (A(ti, ixlb + 1, ny + 1) + A(ti, ixlb, ny)) - A(ti, ixlb, ny + 1)) / (P.temp0 * ev);
      if (HASG) yldot[IDXG(ixlb, ny + 1)] = P.nurlxg * (A(ng, ixlb + 1, ny + 1) - A(ng, ixlb, ny + 1)) / P.n0g;
    }
    if (w.xcnearrb || w.openbox) {  // boundary.m:1585-1630
      yldot[IDXU(ixrb, ny + 1)] = -P.nurlxu * (A(up, ixrb, ny + 1) - 0.5 * (A(up, ixrb - 1, ny + 1) + A(up, ixrb, ny))) / P.vpnorm;
      yldot[IDXU(ixrb + 1, ny + 1)] = -P.nurlxu * (A(up, ixrb + 1, ny + 1) - A(up, ixrb, ny + 1)) / P.vpnorm;
      yldot[IDXTE(ixrb + 1, ny + 1)] = P.nurlxe * (0.5 * (A(te, ixrb, ny + 1) + A(te, ixrb + 1, ny)) - A(te, ixrb + 1, ny + 1)) / (P.temp0 * ev);
      yldot[IDXTI(ixrb + 1, ny + 1)] = P.nurlxi * (0.5 * (A(ti, ixrb, ny + 1) + A(ti, ixrb + 1, ny)) - A(ti, ixrb + 1, ny + 1)) / (P.temp0 * ev);
      if (HASG) yldot[IDXG(ixrb + 1, ny + 1)] = P.nurlxg * (A(ng, ixrb, ny + 1) - A(ng, ixrb + 1, ny + 1)) / P.n0g;
    }
  }
  // ===== ix = 0 as a symmetry plane, isfixlb = 2 (boundary.m:1666-1770; rlimiter beyond the mesh) =====
  if (w.i3 <= 0 && P.isfixlb == 2)
    for (int iy = w.j2; iy <= w.j5; ++iy) {
      yldot[IDXN(0, iy)] = P.nurlxn * (1 / P.n0) * (A(ni, 1, iy) - A(ni, 0, iy));
      yldot[IDXU(0, iy)] = P.nurlxu * (0. - A(up, 0, iy)) / P.vpnorm;
      yldot[IDXTE(0, iy)] = P.nurlxe * A(ne, 0, iy) * (A(te, 1, iy) - A(te, 0, iy)) / P.ennorm;
      yldot[IDXTI(0, iy)] = P.nurlxi * A(ne, 0, iy) * (A(ti, 1, iy) - A(ti, 0, iy)) / P.ennorm;
      if (HASG) yldot[IDXG(0, iy)] = P.nurlxg * (A(ng, 1, iy) - A(ng, 0, iy)) / P.n0g;
    }
  // velocity forced to zero on the cut of a half-space problem (boundary.m:1772-1785)
  if (P.isfixlb == 2 && w.i2 <= P.ixpt2 && w.i5 >= P.ixpt2 && w.j2 <= P.iysptrx2)
    for (int iy = 0; iy <= P.iysptrx2; ++iy) yldot[IDXU((int)P.ixpt2, iy)] = P.nurlxu * (0. - A(up, (int)P.ixpt2, iy)) / P.vpnorm;
  // ===== left plate, ix = ixlb (boundary.m:1787-2318), isfixlb = 0 =====
  if ((w.xcnearlb || w.openbox) && P.isfixlb == 0) {
    const int ixt = ixlb;
    if (w.i3 <= ixlb + P.isextrnp)  // boundary.m:1796-1845 (isextrnp = 0)
      for (int iy = w.j2; iy <= w.j5; ++iy) {
        int ixt1 = IXP1(ixt, iy);
        yldot[IDXN(ixt, iy)] = P.nurlxn * (A(ni, ixt1, iy) - A(ni, ixt, iy)) / P.n0;
      }
    if (w.i3 <= ixlb)
      for (int iy = w.j2; iy <= w.j5; ++iy) {  // boundary.m:1848-2259
        int ixt1 = IXP1(ixt, iy);
        double ueb = P.cfueb * (0. - 0.) / G(rrv, ixt, iy);  // cf2ef = 0, vytan = 0
        double cs = P.csfaclb * std::sqrt((A(te, ixt, iy) + P.csfacti * A(ti, ixt, iy)) / P.mi);
        int64_t iv2 = IDXU(ixt, iy);
        yldot[iv2] = P.nurlxu * (-cs - ueb - A(up, ixt, iy)) / P.vpnorm;  // isbohmms = 0
        if (P.isupss == 1 && A(up, ixt1, iy) + ueb < -cs) yldot[iv2] = P.nurlxu * (A(up, ixt1, iy) - A(up, ixt, iy)) / P.vpnorm;
        if (P.isupss == -1) yldot[iv2] = P.nurlxu * (A(up, ixt1, iy) - A(up, ixt, iy)) / P.vpnorm;
        double kfeix = 0.;
        kfeix = kfeix - P.cfvcsx * 0.5 * G(sx, ixt, iy) * A(visx, ixt1, iy) * G(gx, ixt1, iy) * (A(up, ixt1, iy) * A(up, ixt1, iy) - A(up, ixt, iy) * A(up, ixt, iy));
        double kappal = 3.;  // isphion = 0, boundary.m:1972-1973
        (void)kappal;
        double bcel = (1 - P.newbcl * 0) * P.bcee + P.newbcl * 0 * (2. + kappal);
        double bcil = (1 - P.newbcl * 0) * P.bcei + P.newbcl * 0 * (2.5);
        double t0 = A(te, ixt, iy) / ev;
        double f_cgpld = .5 * (1. - ue_cos(pi * (t0 - P.temin) / (.3 - P.temin)));
        if (t0 < P.temin) f_cgpld = 0.;
        if (t0 > 0.3) f_cgpld = 1.;
        t0 = std::max(A(tg, ixt1, iy), P.tgmin * ev);
        double vxn = f_cgpld * 0.25 * std::sqrt(8 * t0 / (pi * P.mg));
        {  // ibctepl == 1, boundary.m:1996-2014
          double totfeexl = A(feex, ixt, iy) + 0.;  // cfeexdbo = 0
          double totfnex = A(ne, ixt, iy) * A(vex, ixt, iy) * G(sx, ixt, iy);
          yldot[IDXTE(ixt, iy)] = -P.nurlxe * (totfeexl - totfnex * A(te, ixt, iy) * bcel + P.cgpld * G(sx, ixt, iy) * 0.5 * A(ng, ixt1, iy) * vxn * P.ediss * ev -
                                               P.cmneut * A(fnix, ixt, iy) * P.recycp * P.eedisspl * ev) / (G(sx, ixt, iy) * P.vpnorm * P.ennorm);
        }
        {  // ibctipl == 1, boundary.m:2027-2063
          double totfeixl = A(feix, ixt, iy) + P.ckinfl * kfeix;
          double totfnix = 0.;
          totfeixl = totfeixl + 0.;  // cfeixdbo = 0
          totfnix = totfnix + A(fnix, ixt, iy);
          yldot[IDXTI(ixt, iy)] = -P.nurlxi * (totfeixl - totfnix * bcil * A(ti, ixt, iy) +
                                               P.cftiexclg * (-P.cmneut * A(fnix, ixt, iy) * P.recycp * P.cmntgpl * (A(ti, ixt, iy) - P.eidisspl * ev))) /
                                  (P.vpnorm * P.ennorm * G(sx, ixt, iy));
        }
        if (HASG) {  // neutral density, boundary.m:2075-2115
          int64_t iv = IDXG(ixt, iy);
          double recy = P.recylb[iy];
          if (recy > 0.) {
            double flux_inc = P.fac2sp * A(fnix, ixt, iy);
            double t0g = std::max(A(tg, ixt1, iy), P.tgmin * ev);
            double vxg = 0.25 * std::sqrt(8 * t0g / (pi * P.mg));
            double areapl = P.isoldalbarea * G(sx, ixt, iy) + (1 - P.isoldalbarea) * G(sxnp, ixt, iy);
            yldot[iv] = -P.nurlxg * (A(fngx, ixt, iy) - P.fngxlb_use[iy] - P.fngxslb[iy] + recy * flux_inc + (1 - P.alblb[iy]) * A(ng, ixt1, iy) * vxg * areapl) /
                        (P.vpnorm * P.n0g * G(sx, ixt, iy));
          } else if (recy <= 0. && recy >= -1.) {
            double t0g = std::max(A(tg, ixt, iy), P.tgmin * ev);
            double vxg = 0.25 * std::sqrt(8 * t0g / (pi * P.mg));
            yldot[iv] = -P.nurlxg * (A(fngx, ixt, iy) + (1 + recy) * A(ng, ixt, iy) * vxg * G(sx, ixt, iy)) / (vxg * G(sx, ixt, iy) * P.n0g);
          } else { err = "oracle: recylb < -1 not built"; return -4; }
        }
      }
  }
  // ===== right plate, ix = ixrb+1 (boundary.m:2320-3002), isfixrb = 0 =====
  if (w.xcnearrb || w.openbox) {
    const int ixt = ixrb + 1;
    if (w.i6 >= (ixrb + 1 - P.isextrnp))  // boundary.m:2460-2510
      for (int iy = w.j2; iy <= w.j5; ++iy) {
        int ixt1 = IXM1(ixt, iy);
        yldot[IDXN(ixt, iy)] = P.nurlxn * (A(ni, ixt1, iy) - A(ni, ixt, iy)) / P.n0;
      }
    if (w.i6 >= ixrb + 1)
      for (int iy = w.j2; iy <= w.j5; ++iy) {  // boundary.m:2513-2800
        int ixt1 = IXM1(ixt, iy), ixt2 = IXM1(ixt1, iy);
        A(upi, ixt, iy) = A(upi, ixt1, iy);  // boundary.m:2524-2525
        A(upi, ixt, iy) = A(up, ixt, iy);
        double ueb = P.cfueb * (0. - 0.) / G(rrv, ixt1, iy);
        int64_t iv2 = IDXU(ixt1, iy), iv = IDXU(ixt, iy);
        double cs = P.csfacrb * std::sqrt((A(te, ixt, iy) + P.csfacti * A(ti, ixt, iy)) / P.mi);
        yldot[iv2] = P.nurlxu * (cs - ueb - A(up, ixt1, iy)) / P.vpnorm;  // isbohmms = 0
        if (P.isupss == 1 && A(up, ixt2, iy) + ueb > cs) yldot[iv2] = P.nurlxu * (A(up, ixt2, iy) - A(up, ixt1, iy)) / P.vpnorm;
        if (P.isupss == -1) yldot[iv2] = P.nurlxu * (A(up, ixt2, iy) - A(up, ixt1, iy)) / P.vpnorm;
        yldot[iv] = P.nurlxu * (A(up, ixt1, iy) - A(up, ixt, iy)) / P.vpnorm;  // boundary.m:2588
        double kfeix = 0.;
        kfeix = kfeix - P.cfvcsx * 0.5 * G(sx, ixt1, iy) * A(visx, ixt1, iy) * G(gx, ixt1, iy) * (A(up, ixt1, iy) * A(up, ixt1, iy) - A(up, ixt2, iy) * A(up, ixt2, iy));
        double kappar = 3.;
        double bcer = (1 - P.newbcr * 0) * P.bcee + P.newbcr * 0 * (2. + kappar);
        double bcir = (1 - P.newbcr * 0) * P.bcei + P.newbcr * 0 * (2.5);
        double t0 = A(te, ixt, iy) / ev;
        double f_cgpld = .5 * (1. - ue_cos(pi * (t0 - P.temin) / (.3 - P.temin)));
        if (t0 < P.temin) f_cgpld = 0.;
        if (t0 > 0.3) f_cgpld = 1.;
        t0 = std::max(A(tg, ixt1, iy), P.tgmin * ev);
        double vxn = f_cgpld * 0.25 * std::sqrt(8 * t0 / (pi * P.mg));
        {  // ibctepr == 1, boundary.m:2673-2696 (eedisspr = eedisspl default 0; passed as eedisspl)
          double totfeexr = A(feex, ixt1, iy) + 0.;
          double totfnex = A(ne, ixt, iy) * A(vex, ixt1, iy) * G(sx, ixt1, iy);
          yldot[IDXTE(ixt, iy)] = P.nurlxe * (totfeexr - totfnex * A(te, ixt, iy) * bcer - P.cgpld * G(sx, ixt1, iy) * 0.5 * A(ng, ixt1, iy) * vxn * P.ediss * ev -
                                              P.cmneut * A(fnix, ixt1, iy) * P.recycp * P.eedisspl * ev) / (G(sx, ixt1, iy) * P.vpnorm * P.ennorm);
        }
        {  // ibctipr == 1, boundary.m:2709-2747
          double totfeixr = A(feix, ixt1, iy) + P.ckinfl * kfeix;
          double totfnix = 0.;
          totfeixr = totfeixr + 0.;
          totfnix = totfnix + A(fnix, ixt1, iy);
          yldot[IDXTI(ixt, iy)] = P.nurlxi * (totfeixr - totfnix * bcir * A(ti, ixt, iy) +
                                              P.cftiexclg * (-P.cmneut * A(fnix, ixt1, iy) * P.recycp * P.cmntgpl * (A(ti, ixt, iy) - P.eidisspl * ev))) /
                                  (P.vpnorm * P.ennorm * G(sx, ixt1, iy));
        }
        if (HASG) {  // boundary.m:2759-2799
          int64_t ivg = IDXG(ixt, iy);
          double recy = P.recyrb[iy];
          if (recy > 0.) {
            double flux_inc = P.fac2sp * A(fnix, ixt1, iy);
            double t0g = std::max(A(tg, ixt1, iy), P.tgmin * ev);
            double vxg = 0.25 * std::sqrt(8 * t0g / (pi * P.mg));
            double areapl = P.isoldalbarea * G(sx, ixt1, iy) + (1 - P.isoldalbarea) * G(sxnp, ixt1, iy);
            yldot[ivg] = P.nurlxg * (A(fngx, ixt1, iy) + P.fngxrb_use[iy] - P.fngxsrb[iy] + recy * flux_inc - (1 - P.albrb[iy]) * A(ng, ixt1, iy) * vxg * areapl) /
                         (P.vpnorm * P.n0g * G(sx, ixt1, iy));
          } else if (recy <= 0. && recy >= -1.) {
            double t0g = std::max(A(tg, ixt, iy), P.tgmin * ev);
            double vxg = 0.25 * std::sqrt(8 * t0g / (pi * P.mg));
            yldot[ivg] = P.nurlxg * (A(fngx, ixt1, iy) - (1 + recy) * A(ng, ixt, iy) * vxg * G(sx, ixt1, iy)) / (vxg * G(sx, ixt1, iy) * P.n0g);
          } else { err = "oracle: recyrb < -1 not built"; return -4; }
        }
      }
  }
  return 0;
}

// ---- pandf (oderhs.m:537-5070) ----------------------------------------------------
int pandf(int xc, int yc, const double* yl, double* yldot) {
  const Win w = make_win(xc, yc);
  const double ev = P.ev, qe = P.qe, cutlo = P.cutlo;
  int rc = convsr_vo(xc, yc, yl);  // oderhs.m:1025
  if (rc) return rc;
  convsr_aux(xc, yc);              // oderhs.m:1028
  const int i1 = w.i1, i2 = w.i2, i4 = w.i4, i5 = w.i5, i6 = w.i6, i8 = w.i8;
  const int j1 = w.j1, j2 = w.j2, j4 = w.j4, j5 = w.j5, j6 = w.j6, j8 = w.j8;
  const int ixlb = (int)P.ixlb, ixrb = (int)P.ixrb;

  // Coulomb logarithm on x-faces (oderhs.m:1138-1155)
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {
      int ix1 = IXP1(ix, iy);
      double teev = 0.5 * (A(te, ix, iy) + A(te, ix1, iy)) / ev;
      double nexface = 0.5 * (A(ne, ix, iy) + A(ne, ix1, iy));
      if (P.islnlamcon == 1) A(loglambda, ix, iy) = P.lnlam;
      else if (teev < 50.) A(loglambda, ix, iy) = 23.4 - 1.15 * ue_log10(1.e-6 * nexface) + 3.45 * ue_log10(teev);
      else A(loglambda, ix, iy) = 25.3 - 1.15 * ue_log10(1.e-6 * nexface) + 2.33167537087122e+00 * ue_log10(teev);
    }
  // radial velocity: only the diffusive part survives (cfydd=cfrd=cfyef=cfybf=cfvycf=cfvycr=0) (oderhs.m:1174-1320)
  for (int iy = j1; iy <= j5; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {
      double vydd = P.vcony + 0. + 0. - (P.difpr + 0.) * (2 * A(gpry, ix, iy) / (A(pr, ix, iy + 1) + A(pr, ix, iy)) - 3.0 * A(gtey, ix, iy) / (A(tey1, ix, iy) + A(tey0, ix, iy)));
      A(diffusivwrk, ix, iy) = P.fcdif * P.difni + 0.;
      // stored for the second loop nest of the reference (oderhs.m:1270-1320)
      A(vy, ix, iy) = vydd;
    }
  for (int iy = j1; iy <= j5; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {
      double difnimix = A(diffusivwrk, ix, iy);
      double vydd = A(vy, ix, iy) - 1. * difnimix * (2 * (1 - P.isvylog) * ((A(niy1, ix, iy) - A(niy0, ix, iy)) / G(dynog, ix, iy)) / (A(niy1, ix, iy) + A(niy0, ix, iy)) +
                                                     P.isvylog * (ue_log(A(niy1, ix, iy)) - ue_log(A(niy0, ix, iy))) / G(dynog, ix, iy));
      A(vy, ix, iy) = vydd;
    }
  for (int ix = i1; ix <= i6; ++ix) A(vy, ix, ny + 1) = 0.0;  // oderhs.m:1466-1468

  // thermal force / friction (oderhs.m:1516-1534), fqp = 0
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {
      int ix2 = IXP1(ix, iy);
      double nbarx = 0.5 * (A(ne, ix, iy) + A(ne, ix2, iy));
      double ltmax = std::min(std::fabs(A(te, ix, iy) / (G(rrv, ix, iy) * A(gtex, ix, iy) + cutlo)), G(lcone, ix, iy));
      double lmfpe = 2e16 * ((A(te, ix, iy) / ev) * (A(te, ix, iy) / ev)) / A(ne, ix, iy);
      double flxlimf = P.flalftf * ltmax / (P.flalftf * ltmax + lmfpe);
      A(frice, ix, iy) = -P.cthe * flxlimf * nbarx * G(rrv, ix, iy) * A(gtex, ix, iy) + 0.;
      A(frici, ix, iy) = -A(frice, ix, iy);
    }
  // parallel electric field from electron momentum balance (oderhs.m:1541-1567)
  for (int iy = w.iys1; iy <= w.iyf6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {
      int ix1 = ix;
      if (ix == ixlb) ix1 = ixlb + 1; else if (ix == ixrb) ix1 = ixrb - 1;
      int ix2 = IXP1(ix1, iy);
      double ltmax = std::min(std::fabs(A(te, ix, iy) / (G(rrv, ix, iy) * A(gtex, ix, iy) + cutlo)), G(lcone, ix, iy));
      double lmfpe = 2e16 * ((A(te, ix, iy) / ev) * (A(te, ix, iy) / ev)) / A(ne, ix, iy);
      double flxlimf = P.flalftf * ltmax / (P.flalftf * ltmax + lmfpe);
      double nexface = 0.5 * (A(ne, ix2, iy) + A(ne, ix1, iy));
      A(ex, ix, iy) = (1 - 0) * (-(A(gpex, ix1, iy) / nexface + P.cthe * flxlimf * A(gtex, ix1, iy)) / qe - 0. + 0.);
    }
  // upi, uup (oderhs.m:1577-1643)
  for (int iy = w.iys1; iy <= w.iyf6; ++iy) {
    if (xc > 0) {
      int ix1 = IXM1(xc, iy);
      A(upi, ix1, iy) = A(up, ix1, iy);
      A(uup, ix1, iy) = G(rrv, ix1, iy) * A(upi, ix1, iy);
    }
    for (int ix = w.ixs1; ix <= std::min(w.ixf6, nx); ++ix) {
      A(upi, ix, iy) = A(up, ix, iy);
      A(uup, ix, iy) = G(rrv, ix, iy) * A(upi, ix, iy);
    }
  }
  // poloidal ion velocity uu (oderhs.m:1648-1682); v2 = vytan = 0, difax = 0
  for (int iy = j1; iy <= j6; ++iy) {
    if (i1 > 0) { int ix1 = IXM1(i1, iy); A(uu, ix1, iy) = A(uup, ix1, iy) + 0. - 0. - 0.; }
    for (int ix = i1; ix <= i6; ++ix) A(uu, ix, iy) = A(uup, ix, iy) + 0. - 0. - 0.;
  }
  // electron velocities (oderhs.m:1729-1792)
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) { A(vex, ix, iy) = 0.; A(vey, ix, iy) = 0.; A(upe, ix, iy) = 0.; }
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {
      int ix1 = IXP1(ix, iy);
      A(upe, ix, iy) = A(upe, ix, iy) + A(upi, ix, iy) * P.zi * 0.5 * (A(ni, ix, iy) + A(ni, ix1, iy));
    }
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {
      int ix1 = IXP1(ix, iy);
      A(upe, ix, iy) = (A(upe, ix, iy) - 0.) / (0.5 * (A(ne, ix, iy) + A(ne, ix1, iy)));
    }
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) A(vex, ix, iy) = A(upe, ix, iy) * G(rrv, ix, iy) + 0. - 0.;
  for (int iy = j1; iy <= j5; ++iy)
    for (int ix = i1; ix <= i6; ++ix) A(vey, ix, iy) = A(vey, ix, iy) + A(vy, ix, iy) * P.zi * 0.5 * (A(niy0, ix, iy) + A(niy1, ix, iy));
  for (int iy = j1; iy <= j5; ++iy)
    for (int ix = i1; ix <= i6; ++ix) A(vey, ix, iy) = (A(vey, ix, iy) - 0.) / (0.5 * (A(ney0, ix, iy) + A(ney1, ix, iy)));

  // zero the source accumulators (oderhs.m:1818-1835)
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) {
      A(snic, ix, iy) = 0.; A(sniv, ix, iy) = 0.; A(psori, ix, iy) = 0.; A(smoc, ix, iy) = 0.; A(smov, ix, iy) = 0.;
      A(seec, ix, iy) = 0.; A(seev, ix, iy) = 0.; A(seic, ix, iy) = 0.; A(seiv, ix, iy) = 0.;
    }
  // ionisation / recombination / charge exchange (oderhs.m:1909-2008), rtau = 0
  double nuizold = 0., nurcold = 0.;
  if (xc >= 0 && yc >= 0) { nuizold = A(nuiz, xc, yc); nurcold = A(nurc, xc, yc); }
  for (int iy = w.iys1; iy <= w.iyf6; ++iy)
    for (int ix = w.ixs1; ix <= w.ixf6; ++ix) {
      if (P.icnuiz == 0) {
        double ne_sgvi = A(ne, ix, iy);
        if (P.ifxnsgi == 1) ne_sgvi = P.cne_sgvi;
        A(nuiz, ix, iy) = P.chioniz * A(ne, ix, iy) * (rsa(A(te, ix, iy), ne_sgvi) + P.sigvi_floor);
        if (xc >= 0) A(nuiz, ix, iy) = P.fnnuiz * A(nuiz, ix, iy) + (1 - P.fnnuiz) * nuizold;
      } else A(nuiz, ix, iy) = P.cnuiz;
      if (P.isrecmon == 1) {
        A(nurc, ix, iy) = P.cfrecom * A(ne, ix, iy) * rra(A(te, ix, iy), A(ne, ix, iy));
        if (xc >= 0) A(nurc, ix, iy) = P.fnnuiz * A(nurc, ix, iy) + (1 - P.fnnuiz) * nurcold;
      } else A(nurc, ix, iy) = 0.;
      A(psorbgg, ix, iy) = P.ngbackg * ((0.9 + 0.1 * powi(P.ngbackg / A(ng, ix, iy), P.ingb))) * A(nuiz, ix, iy) * G(vol, ix, iy);
      A(psorgc, ix, iy) = -A(ng, ix, iy) * A(nuiz, ix, iy) * G(vol, ix, iy) + A(psorbgg, ix, iy);
      A(psorc, ix, iy) = -A(psorgc, ix, iy);
      A(psordis, ix, iy) = P.cfdiss * A(psorc, ix, iy);
      A(psorxrc, ix, iy) = -A(ni, ix, iy) * A(nurc, ix, iy) * G(vol, ix, iy);
      A(psorrgc, ix, iy) = -A(psorxrc, ix, iy);
      if (P.icnucx == 0) {
        double t0 = std::max(A(ti, ix, iy), P.temin * ev);
        double t1 = t0 / (P.mi / P.mp);
        A(nucx, ix, iy) = A(ni, ix, iy) * rcx(t1);
      } else if (P.icnucx == 1) A(nucx, ix, iy) = P.cnucx;
      else {
        double t0 = std::max(A(ti, ix, iy), P.temin * ev);
        A(nucx, ix, iy) = std::sqrt(t0 / P.mi) * P.sigcx * (A(ni, ix, iy) + P.rnn2cx * A(ng, ix, iy));
      }
      A(nuix, ix, iy) = P.fnuizx * A(nuiz, ix, iy) + P.fnucxx * A(nucx, ix, iy);
    }
  for (int iy = w.iys1; iy <= w.iyf6; ++iy)  // ispsorave = 0 (oderhs.m:2017-2029)
    for (int ix = w.ixs1; ix <= w.ixf6; ++ix) {
      A(psorg, ix, iy) = A(psorgc, ix, iy); A(psor, ix, iy) = A(psorc, ix, iy);
      A(psorxr, ix, iy) = A(psorxrc, ix, iy); A(psorrg, ix, iy) = A(psorrgc, ix, iy);
    }

  neudifpg(w);  // oderhs.m:2428

  // half-space problem: no flux and no gradients through the cut (oderhs.m:2447-2466)
  if (P.isfixlb == 2) {
    const int ix = (int)P.ixpt2;
    if (ix >= i2 && ix <= i5 + 1 && P.iysptrx1 > 0)
      for (int iy = 0; iy <= P.iysptrx1; ++iy) {
        A(gpex, ix, iy) = 0.; A(frice, ix, iy) = 0.; A(ex, ix, iy) = 0.; A(upe, ix, iy) = 0.;
        A(gpix, ix, iy) = 0.; A(frici, ix, iy) = 0.; A(uu, ix, iy) = 0.; A(upi, ix, iy) = 0.;
      }
  }

  // electron / ion pressure-work and momentum sources (oderhs.m:2471-2579)
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) {
      int ix1 = IXM1(ix, iy), ix2 = IXP1(ix, iy);
      double t1old = .5 * P.cvgp * (A(upe, ix, iy) * G(rrv, ix, iy) * ave(G(gx, ix, iy), G(gx, ix2, iy)) * A(gpex, ix, iy) / G(gxf, ix, iy) +
                                    A(upe, ix1, iy) * G(rrv, ix1, iy) * ave(G(gx, ix, iy), G(gx, ix1, iy)) * A(gpex, ix1, iy) / G(gxf, ix1, iy));
      double t2old = 0.;  // fqp = 0
      int iyp1 = std::min(iy + 1, ny + 1), iym1 = std::max(iy - 1, 0);
      double t1new = .5 * P.cvgp * (A(vex, ix, iy) * ave(G(gx, ix, iy), G(gx, ix2, iy)) * A(gpex, ix, iy) / G(gxf, ix, iy) +
                                    A(vex, ix1, iy) * ave(G(gx, ix, iy), G(gx, ix1, iy)) * A(gpex, ix1, iy) / G(gxf, ix1, iy));
      double t2new = .5 * P.cvgp * (A(vey, ix, iy) * ave(G(gy, ix, iy), G(gy, ix, iyp1)) * A(gpey, ix, iy) / G(gyf, ix, iy) +
                                    A(vey, ix, iy) * ave(G(gy, ix, iy), G(gy, ix, iym1)) * A(gpey, ix, iym1) / G(gyf, ix, iym1));
      A(seec, ix, iy) = A(seec, ix, iy) + (t1old * G(vol, ix, iy) - t2old) * P.oldseec + ((t1new + t2new) * G(vol, ix, iy)) * (1 - P.oldseec);
      A(smoc, ix, iy) = ((-P.cpgx * A(gpex, ix, iy) - 0.) * G(rrv, ix, iy) + 0.) * G(sx, ix, iy) / G(gxf, ix, iy);
    }
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) {
      int ix1 = IXM1(ix, iy), ix2 = IXP1(ix, iy);
      double tv = A(gpix, ix, iy) / G(gxf, ix, iy);
      double t1 = A(gpix, ix1, iy) / G(gxf, ix1, iy);
      t1 = .5 * P.cvgp * (A(up, ix, iy) * G(rrv, ix, iy) * ave(G(gx, ix2, iy), G(gx, ix, iy)) * tv + A(up, ix1, iy) * G(rrv, ix1, iy) * ave(G(gx, ix, iy), G(gx, ix1, iy)) * t1);
      A(seic, ix, iy) = A(seic, ix, iy) + P.cfvgpx * t1 * G(vol, ix, iy);
      double t0 = -P.cpiup * (A(gpix, ix, iy) * G(rrv, ix, iy) - 0.) * G(sx, ix, iy) / G(gxf, ix, iy);
      A(smoc, ix, iy) = A(smoc, ix, iy) + P.cpgx * t0;
      tv = 0.25 * (A(frice, ix, iy) + A(frice, ix1, iy)) * (A(upe, ix, iy) + A(upe, ix1, iy) - A(upi, ix, iy) - A(upi, ix1, iy));
      A(seec, ix, iy) = A(seec, ix, iy) - (P.zi * P.zi) * A(ni, ix, iy) * tv * G(vol, ix, iy) / A(nz2, ix, iy);
    }
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) {  // isgpye == 0; v2xgp = 0 (oderhs.m:2544-2575)
      double t1 = .5 * P.cvgp * (A(vy, ix, iy) * A(gpiy, ix, iy) + A(vy, ix, iy - 1) * A(gpiy, ix, iy - 1) + 0. + 0.);
      double t2 = t1;
      A(seec, ix, iy) = A(seec, ix, iy) - P.fluxfacy * t1 * G(vol, ix, iy);
      A(seic, ix, iy) = A(seic, ix, iy) + P.fluxfacy * P.cfvgpy * t2 * G(vol, ix, iy);
    }

  // viscosity (oderhs.m:2718-2787), nisp = 1
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {
      double tvw = (P.zi * P.zi) / std::sqrt((P.mi + P.mi) / (2 * P.mp));
      A(w0, ix, iy) = 0.0;  // the reference's w(ix,iy)
      A(w0, ix, iy) = A(w0, ix, iy) + tvw * A(ni, ix, iy);
    }
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {
      double ctaui = 2.1e13 / (A(loglambda, ix, iy) * (P.zi * P.zi));
      double tv2 = ctaui / (ev * std::sqrt(ev));
      double a = (P.convis == 0) ? std::max(A(ti, ix, iy), P.temin * ev) : P.afix * ev;
      double visxtmp = tv2 * P.coef * G(rr, ix, iy) * G(rr, ix, iy) * a * a * std::sqrt(a) * A(ni, ix, iy) / A(w0, ix, iy);
      A(visx, ix, iy) = P.parvis * visxtmp + 0. * A(nm, ix, iy);
      int ix1 = IXM1(ix, iy);
      double t0 = std::max(A(ti, ix, iy), P.temin * ev);
      double mfl = P.flalfv * A(nm, ix, iy) * G(rr, ix, iy) * G(vol, ix, iy) * G(gx, ix, iy) * (t0 / P.mi);
      double csh;
      if (P.isgxvon == 0) csh = A(visx, ix, iy) * G(vol, ix, iy) * G(gx, ix, iy) * G(gx, ix, iy);
      else csh = A(visx, ix, iy) * G(vol, ix, iy) * G(gx, ix, iy) * 2 * G(gxf, ix, iy) * G(gxf, ix1, iy) / (G(gxf, ix, iy) + G(gxf, ix1, iy));
      double msh = std::fabs(csh * (A(upi, ix1, iy) - A(upi, ix, iy)));
      A(visx, ix, iy) = A(visx, ix, iy) / ue_pow(1 + ue_pow(msh / (mfl + 1.e-20 * msh), P.flgamv), 1 / P.flgamv);
      A(visy, ix, iy) = (P.fcdif * P.travis + 0.) * A(nm, ix, iy) + 4 * 0.;
    }

  // heat conduction coefficients (oderhs.m:2801-3017), nisp = 1
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {
      A(hcxe, ix, iy) = 0.; A(hcxi, ix, iy) = 0.; A(hcye, ix, iy) = 0.; A(hcyi, ix, iy) = 0.; A(hcxij, ix, iy) = 0.; A(hcyij, ix, iy) = 0.;
      A(w1, ix, iy) = 0.; A(w2, ix, iy) = 0.;
    }
  {
    double tv = P.zi * P.zi;
    double a = (P.zi * P.zi) * std::sqrt(2 * P.mi * P.mi / (P.mi + P.mi));
    for (int iy = j1; iy <= j6; ++iy)
      for (int ix = i1; ix <= i6; ++ix) {
        int ix1 = IXP1(ix, iy);
        A(w1, ix, iy) = A(w1, ix, iy) + tv * (A(ni, ix, iy) * G(gx, ix, iy) + A(ni, ix1, iy) * G(gx, ix1, iy)) / (G(gx, ix, iy) + G(gx, ix1, iy));
        A(w2, ix, iy) = A(w2, ix, iy) + a * (A(ni, ix, iy) * G(gx, ix, iy) + A(ni, ix1, iy) * G(gx, ix1, iy)) / (G(gx, ix, iy) + G(gx, ix1, iy));
      }
  }
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {
      int ix1 = IXP1(ix, iy), iyp1 = std::min(ny + 1, iy + 1);
      double ctaue = 3.5e11 * P.zi / A(loglambda, ix, iy);
      double ctaui = 2.1e13 / (A(loglambda, ix, iy) * (P.zi * P.zi));
      double fxe = P.kxe * P.ce * ctaue / (P.me * ev * std::sqrt(ev));
      double fxi = P.kxi * P.ci * ctaui / (ev * std::sqrt(ev * P.mp));
      double fxet = fxe, fxit = fxi;
      if ((iy <= P.iysptrx) && ix > P.ixpt1 && ix <= P.ixpt2) {
        fxet = fxe / (1. + (P.rkxecore - 1.) * powi(P.yyf[iy] / (P.yyf[0] + 4.e-50), P.inkxc));
        fxit = P.kxicore * fxi;
      }
      double niavex = (A(ni, ix, iy) * G(gx, ix, iy) + A(ni, ix1, iy) * G(gx, ix1, iy)) / (G(gx, ix, iy) + G(gx, ix1, iy));
      double niavey = (A(niy0, ix, iy) * G(gy, ix, iy) + A(niy1, ix, iy) * G(gy, ix, iyp1)) / (G(gy, ix, iy) + G(gy, ix, iyp1));
      A(hcxe, ix, iy) = A(hcxe, ix, iy) + fxet * niavex / A(w1, ix, iy);
      double kyemix = P.fcdif * P.kye + 0.;
      if (P.kyet > 1.e-20 && iy > P.iysptrx) kyemix = (1. - P.ckyet) * kyemix + P.ckyet * P.kyet * A(diffusivwrk, ix, iy);
      A(hcye, ix, iy) = A(hcye, ix, iy) + (kyemix + 2.33 * (0. + 0.)) * P.zi * niavey;
      A(hcxij, ix, iy) = fxit * niavex / A(w2, ix, iy);
      double kyimix = P.fcdif * P.kyi + 0.;
      if (P.kyit > 1.e-20 && iy > P.iysptrx) kyimix = (1. - P.ckyit) * kyimix + P.ckyit * P.kyit * A(diffusivwrk, ix, iy);
      A(hcyij, ix, iy) = A(hcyij, ix, iy) + (kyimix + (0. + 0.)) * niavey;
    }
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {  // oderhs.m:2906-2965
      int ix1 = IXP1(ix, iy);
      double a, tiave = 0.;
      if (P.concap == 0) {
        tiave = (A(ti, ix, iy) * G(gx, ix, iy) + A(ti, ix1, iy) * G(gx, ix1, iy)) / (G(gx, ix, iy) + G(gx, ix1, iy));
        if (ix == ixlb) tiave = A(ti, ixlb + 1, iy);
        if (ix == ixrb) tiave = A(ti, ixrb, iy);
        a = std::max(tiave, P.temin * ev);
      } else a = P.afix * ev;
      A(hcxij, ix, iy) = A(hcxij, ix, iy) * G(rrv, ix, iy) * G(rrv, ix, iy) * a * a * std::sqrt(a);
      double lmfpi = 1.e16 * ((tiave / ev) * (tiave / ev)) / A(ni, ix, iy);
      double niavex = (A(ni, ix, iy) * G(gx, ix, iy) + A(ni, ix1, iy) * G(gx, ix1, iy)) / (G(gx, ix, iy) + G(gx, ix1, iy));
      A(hcxij, ix, iy) = A(hcxij, ix, iy) / (1. + lmfpi / P.lmfplim);
      double dti = A(ti, ix, iy) - A(ti, ix1, iy);
      double sti = 0.5 * P.alfkxi * (A(ti, ix, iy) + A(ti, ix1, iy));
      A(hcxij, ix, iy) = A(hcxij, ix, iy) * (cutlo + dti * dti) / (cutlo + dti * dti + sti * sti) + 0. * niavex;
      if (P.isflxldi == 2) {
        niavex = (A(ni, ix, iy) * G(gx, ix, iy) + A(ni, ix1, iy) * G(gx, ix1, iy)) / (G(gx, ix, iy) + G(gx, ix1, iy));
        double wallfac = 1.;
        if ((ix == ixlb || ix == ixrb) && (P.isplflxl == 0)) wallfac = P.flalfipl / P.flalfi;
        double qflx = wallfac * P.flalfi * G(rrv, ix, iy) * std::sqrt(a / P.mi) * niavex * a;
        double cshx = A(hcxij, ix, iy);
        double lxtic = 0.5 * (A(ti, ix, iy) + A(ti, ix1, iy)) / (std::fabs(A(ti, ix, iy) - A(ti, ix1, iy)) * G(gxf, ix, iy) + 100. * cutlo);
        double qshx = cshx * (A(ti, ix, iy) - A(ti, ix1, iy)) * G(gxf, ix, iy) * (1. + lxtic / P.lxtimax);
        A(hcxij, ix, iy) = cshx / (1 + std::fabs(qshx / qflx));
      }
      A(hcxi, ix, iy) = A(hcxi, ix, iy) + A(hcxij, ix, iy);
      A(hcyi, ix, iy) = A(hcyi, ix, iy) + A(hcyij, ix, iy);
    }
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {  // oderhs.m:2968-3017
      int ix1 = IXP1(ix, iy), iyp1 = std::min(ny + 1, iy + 1);
      double a;
      if (P.concap == 0) {
        double teave = (A(te, ix, iy) * G(gx, ix, iy) + A(te, ix1, iy) * G(gx, ix1, iy)) / (G(gx, ix, iy) + G(gx, ix1, iy));
        if (ix == ixlb) teave = A(te, ixlb + 1, iy);
        if (ix == ixrb) teave = A(te, ixrb, iy);
        a = std::max(teave, P.temin * ev);
      } else a = P.afix * ev;
      double zeffave = (A(zeff, ix, iy) * G(gx, ix, iy) + A(zeff, ix1, iy) * G(gx, ix1, iy)) / (G(gx, ix, iy) + G(gx, ix1, iy));
      double zcoef = 0.308 + 0.767 * zeffave - 0.075 * (zeffave * zeffave);
      A(hcxe, ix, iy) = A(hcxe, ix, iy) * G(rrv, ix, iy) * G(rrv, ix, iy) * a * a * std::sqrt(a) * zcoef;
      double lmfpe = 2e16 * ((A(te, ix, iy) / ev) * (A(te, ix, iy) / ev)) / A(ne, ix, iy);
      double neavex = (A(ne, ix, iy) * G(gx, ix, iy) + A(ne, ix1, iy) * G(gx, ix1, iy)) / (G(gx, ix, iy) + G(gx, ix1, iy));
      double dte = A(te, ix, iy) - A(te, ix1, iy);
      double ste = 0.5 * P.alfkxe * (A(te, ix, iy) + A(te, ix1, iy));
      A(hcxe, ix, iy) = A(hcxe, ix, iy) * (cutlo + dte * dte) / (cutlo + dte * dte + ste * ste) + 0. * neavex;
      A(hcxe, ix, iy) = A(hcxe, ix, iy) / ((1. + lmfpe / P.lmfplim) * (1 + A(hcxe, ix, iy) * (G(gx, ix, iy) * G(gx, ix, iy)) * P.tdiflim / A(ne, ix, iy)));
      // isupgon == 0: neutral contribution to ion conduction (oderhs.m:3001-3014)
      A(hcxi, ix, iy) = A(hcxi, ix, iy) + P.cftiexclg * P.cfneut * P.cfneutsor_ei * P.kxn * (A(ng, ix, iy) * A(ti, ix, iy) + A(ng, ix1, iy) * A(ti, ix1, iy)) /
                                              (P.mi * (A(nucx, ix, iy) + A(nucx, ix1, iy)));
      A(hcyi, ix, iy) = A(hcyi, ix, iy) + P.cftiexclg * P.cfneut * P.cfneutsor_ei * P.kyn * (A(ngy0, ix, iy) * A(tiy0, ix, iy) + A(ngy1, ix, iy) * A(tiy1, ix, iy)) /
                                              (P.mi * (A(nucx, ix, iy) + A(nucx, ix, iyp1)));
    }
  // equipartition (oderhs.m:3074-3102)
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) A(w3, ix, iy) = 0.0;
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) A(w3, ix, iy) = A(w3, ix, iy) + ((P.zi * P.zi) / P.mi) * A(ni, ix, iy);
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) {
      int ix2 = IXM1(ix, iy);
      double a = std::max(A(te, ix, iy), P.temin * ev);
      double loglmcc = 0.5 * (A(loglambda, ix, iy) + A(loglambda, ix2, iy));
      double coef1 = P.feqp * 4.8e-15 * loglmcc * std::sqrt(ev) * ev * P.mp;
      A(eqp, ix, iy) = coef1 * A(w3, ix, iy) * A(ne, ix, iy) / (a * std::sqrt(a));
      double d = a - A(ti, ix, iy), s = P.alfeqp * (a + A(ti, ix, iy));
      A(eqp, ix, iy) = A(eqp, ix, iy) * (d * d) / (cutlo + d * d + s * s);
    }

  // ion continuity fluxes (oderhs.m:3187-3319), methn upwind/central
  {
    const int methnx = (int)(P.methn % 10), methny = (int)(P.methn / 10);
    for (int iy = j4; iy <= j8; ++iy)
      for (int ix = i1; ix <= i5; ++ix) {
        int ix2 = IXP1(ix, iy);
        double t2;
        if (methnx == 2) t2 = (A(ni, ix, iy) + A(ni, ix2, iy)) / 2;
        else t2 = (A(uu, ix, iy) >= 0.) ? A(ni, ix, iy) : A(ni, ix2, iy);
        A(fnix, ix, iy) = P.cnfx * A(uu, ix, iy) * G(sx, ix, iy) * t2;
        double r1 = P.nlimix * A(ni, ix, iy) / A(ni, ix2, iy), r2 = P.nlimix * A(ni, ix2, iy) / A(ni, ix, iy);
        A(fnix, ix, iy) = A(fnix, ix, iy) / std::sqrt(1 + r1 * r1 + r2 * r2);
      }
    for (int iy = j1; iy <= j5; ++iy)
      for (int ix = i4; ix <= i8; ++ix) {
        double t2;
        if (methny == 2) t2 = (A(niy0, ix, iy) + A(niy1, ix, iy)) / 2;
        else t2 = (A(vy, ix, iy) >= 0.) ? A(niy0, ix, iy) : A(niy1, ix, iy);
        A(fniy, ix, iy) = P.cnfy * A(vy, ix, iy) * G(sy, ix, iy) * t2;
        if (A(vy, ix, iy) * (A(ni, ix, iy) - A(ni, ix, iy + 1)) < 0.) {
          double r1 = P.nlimiy / A(ni, ix, iy + 1), r2 = P.nlimiy / A(ni, ix, iy);
          A(fniy, ix, iy) = A(fniy, ix, iy) / (1 + r1 * r1 + r2 * r2);
        }
      }
    for (int ix = i4; ix <= i8; ++ix) A(fniy, ix, ny + 1) = 0.0;
  }
  for (int ix = i4; ix <= i8; ++ix) fniycbo[ix] = 0.0;  // oderhs.m:3344-3353 with cfybf = 0, cfniydbo = 0
  // particle balance (oderhs.m:3407-3456)
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix)
      A(resco, ix, iy) = A(snic, ix, iy) + A(sniv, ix, iy) * A(ni, ix, iy) + 0. + P.cfneut * P.cfneutsor_ni * P.cnsor * A(psor, ix, iy) +
                         P.cfneut * P.cfneutsor_ni * P.cnsor * A(psorxr, ix, iy) + P.cfneut * P.cfneutsor_ni * P.cnsor * A(psori, ix, iy) - 0. + 0.;
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) {
      int ix1 = IXM1(ix, iy);
      A(resco, ix, iy) = A(resco, ix, iy) - ((A(fnix, ix, iy) - A(fnix, ix1, iy)) + P.fluxfacy * (A(fniy, ix, iy) - A(fniy, ix, iy - 1)));
    }

  // parallel momentum (oderhs.m:3475-3911)
  for (int iy = j4; iy <= j8; ++iy) {
    A(flox, 0, iy) = 0.0; A(conx, 0, iy) = 0.0;
    for (int ix = i2; ix <= i6; ++ix) {
      int ix1 = IXM1(ix, iy);
      double uuv = 0.5 * (A(uu, ix1, iy) + A(uu, ix, iy));
      A(flox, ix, iy) = P.cmfx * A(nm, ix, iy) * uuv * G(vol, ix, iy) * G(gx, ix, iy);
      if (P.isgxvon == 0) A(conx, ix, iy) = A(visx, ix, iy) * G(vol, ix, iy) * G(gx, ix, iy) * G(gx, ix, iy);
      else A(conx, ix, iy) = A(visx, ix, iy) * G(vol, ix, iy) * G(gx, ix, iy) * 2 * G(gxf, ix, iy) * G(gxf, ix1, iy) / (G(gxf, ix, iy) + G(gxf, ix1, iy));
    }
  }
  for (int iy = j1; iy <= j5; ++iy)
    for (int ix = i4; ix <= i8; ++ix) {  // oderhs.m:3512-3575
      int ix2 = IXP1(ix, iy), ix4 = IXP1(ix, iy + 1);
      if (iy == P.iysptrx1 && (ix == P.ixpt1 || ix == P.ixpt2)) {
        A(floy, ix, iy) = (P.cmfy / 2) * G(syv, ix, iy) * (ave(A(nm, ix, iy), A(nm, ix, iy + 1))) * A(vy, ix, iy);
        A(floy, ix, iy) = A(floy, ix, iy) + (P.cmfy / 2) * G(syv, ix, iy) * (ave(A(nm, ix, iy), A(nm, ix, iy + 1))) * 0.;
      } else {
        A(floy, ix, iy) = (P.cmfy / 4) * G(syv, ix, iy) * (ave(A(nm, ix, iy), A(nm, ix, iy + 1)) + ave(A(nm, ix2, iy), A(nm, ix4, iy + 1))) * (A(vy, ix, iy) + A(vy, ix2, iy));
        A(floy, ix, iy) = A(floy, ix, iy) + (P.cmfy / 4) * G(syv, ix, iy) * (ave(A(nm, ix, iy), A(nm, ix, iy + 1)) + ave(A(nm, ix2, iy), A(nm, ix4, iy + 1))) * (0. + 0.);
      }
      if (P.ishavisy == 1)
        A(cony, ix, iy) = .5 * G(syv, ix, iy) * (ave(A(visy, ix, iy) * G(gy, ix, iy), A(visy, ix, iy + 1) * G(gy, ix, iy + 1)) +
                                               ave(A(visy, ix2, iy) * G(gy, ix2, iy), A(visy, ix4, iy + 1) * G(gy, ix4, iy + 1)));
      else
        A(cony, ix, iy) = .25 * P.cfaccony * G(syv, ix, iy) * (A(visy, ix, iy) * G(gy, ix, iy) + A(visy, ix, iy + 1) * G(gy, ix, iy + 1) +
                                                             A(visy, ix2, iy) * G(gy, ix2, iy) + A(visy, ix4, iy + 1) * G(gy, ix4, iy + 1));
    }
  fd2tra(w, flox, floy, conx, cony, up, fmix, fmiy, 1, (int)P.methu);  // oderhs.m:3579
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) {  // oderhs.m:3746-3866
      int ix2 = IXP1(ix, iy);
      double dp1 = P.cngmom * (1 / P.fac2sp) * (A(ng, ix2, iy) * A(tg, ix2, iy) - A(ng, ix, iy) * A(tg, ix, iy));
      A(resmo, ix, iy) = 0.;
      A(resmo, ix, iy) = A(smoc, ix, iy) + A(smov, ix, iy) * A(up, ix, iy) - P.cfneut * P.cfneutsor_mi * G(sx, ix, iy) * G(rrv, ix, iy) * dp1 -
                         P.cfneut * P.cfneutsor_mi * P.cmwall * 0.5 * (A(ng, ix, iy) + A(ng, ix2, iy)) * P.mi * A(up, ix, iy) * 0.5 * (A(nucx, ix, iy) + A(nucx, ix2, iy)) * G(volv, ix, iy) +
                         0. + P.cfmsor * (0. + 0.) + 0. + 0. + 0.;
    }
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) {
      int ix2 = IXP1(ix, iy);
      A(resmo, ix, iy) = A(resmo, ix, iy) - (A(fmix, ix2, iy) - A(fmix, ix, iy) + P.fluxfacy * (A(fmiy, ix, iy) - A(fmiy, ix, iy - 1)));
    }

  // energy equations: convective / conductive coefficients (oderhs.m:3923-4249)
  for (int iy = j1; iy <= j6; ++iy)
    for (int ix = i1; ix <= i6; ++ix) {
      A(floxe, ix, iy) = 0.; A(floxi, ix, iy) = 0.; A(floye, ix, iy) = 0.; A(floyi, ix, iy) = 0.;
      feiycbo[ix] = 0.; feeycbo[ix] = 0.; A(w0, ix, iy) = 0.; A(w1, ix, iy) = 0.;
    }
  for (int iy = j4; iy <= j8; ++iy) {
    for (int ix = i1; ix <= i5; ++ix) {
      int ix2 = IXP1(ix, iy);
      double t0 = std::max(A(te, ix, iy), P.temin * ev), t1 = std::max(A(te, ix2, iy), P.temin * ev);
      double vt0 = std::sqrt(t0 / P.me), vt1 = std::sqrt(t1 / P.me);
      double wallfac = 1.;
      if ((ix == ixlb || ix == ixrb) && (P.isplflxl == 0)) wallfac = P.flalfepl / P.flalfe;
      double qfl = wallfac * P.flalfe * G(sx, ix, iy) * G(rrv, ix, iy) * (A(ne, ix, iy) * vt0 * t0 + A(ne, ix2, iy) * vt1 * t1) / 2;
      double csh = G(sx, ix, iy) * A(hcxe, ix, iy) * G(gxf, ix, iy);
      double lxtec = 0.5 * (A(te, ix, iy) + A(te, ix2, iy)) / (std::fabs(A(te, ix, iy) - A(te, ix2, iy)) * G(gxf, ix, iy) + 100. * cutlo);
      double qsh = csh * (A(te, ix, iy) - A(te, ix2, iy)) * (1. + lxtec / P.lxtemax);
      double qr = (1 - P.isflxlde) * std::fabs(qsh / qfl);
      A(conxe, ix, iy) = (1 - P.isflxlde) * csh / ((1 + qr) * (1 + qr)) + P.isflxlde * csh / ue_pow(1 + ue_pow(std::fabs(qsh / qfl), P.flgam), 1 / P.flgam);
      A(floxe, ix, iy) = A(floxe, ix, iy) + (sgn(qr * qr, qsh) / ((1 + qr) * (1 + qr))) * P.flalfea[ix] * G(sx, ix, iy) *
                                                (A(ne, ix, iy) * G(rr, ix, iy) * vt0 + A(ne, ix2, iy) * G(rr, ix2, iy) * vt1) / 2;
      if (P.isflxldi != 2) {
        t0 = std::max(A(ti, ix, iy), P.temin * ev); t1 = std::max(A(ti, ix2, iy), P.temin * ev);
        vt0 = std::sqrt(t0 / P.mi); vt1 = std::sqrt(t1 / P.mi);
        wallfac = 1.;
        if ((ix == ixlb || ix == ixrb) && (P.isplflxl == 0)) wallfac = P.flalfipl / P.flalfi;
        qfl = wallfac * P.flalfia[ix] * G(sx, ix, iy) * G(rrv, ix, iy) * (A(ne, ix, iy) * vt0 * t0 + A(ne, ix2, iy) * vt1 * t1) / 2;
        csh = G(sx, ix, iy) * A(hcxi, ix, iy) * G(gxf, ix, iy);
        double lxtic = 0.5 * (A(ti, ix, iy) + A(ti, ix2, iy)) / (std::fabs(A(ti, ix, iy) - A(ti, ix2, iy)) * G(gxf, ix, iy) + 100. * cutlo);
        qsh = csh * (A(ti, ix, iy) - A(ti, ix2, iy)) * (1. + lxtic / P.lxtimax);
        qr = (1 - P.isflxldi) * std::fabs(qsh / qfl);
        A(conxi, ix, iy) = (1 - P.isflxldi) * csh / ((1 + qr) * (1 + qr)) + P.isflxldi * csh / ue_pow(1 + ue_pow(std::fabs(qsh / qfl), P.flgam), 1 / P.flgam);
        A(floxi, ix, iy) = A(floxi, ix, iy) + (sgn(qr * qr, qsh) / ((1 + qr) * (1 + qr))) * P.flalfia[ix] * G(sx, ix, iy) *
                                                  (A(ne, ix, iy) * G(rr, ix, iy) * vt0 + A(ne, ix2, iy) * G(rr, ix2, iy) * vt1) / 2;
      } else A(conxi, ix, iy) = G(sx, ix, iy) * A(hcxi, ix, iy) * G(gxf, ix, iy);
    }
    A(conxe, nx + 1, iy) = 0; A(conxi, nx + 1, iy) = 0;
  }
  for (int iy = j1; iy <= j5; ++iy)
    for (int ix = i4; ix <= i8; ++ix) {
      A(conye, ix, iy) = G(sy, ix, iy) * A(hcye, ix, iy) / G(dynog, ix, iy);
      A(conyi, ix, iy) = G(sy, ix, iy) * A(hcyi, ix, iy) / G(dynog, ix, iy);
    }
  for (int ix = i1; ix <= i6; ++ix) { A(conye, ix, ny + 1) = 0.0; A(conyi, ix, ny + 1) = 0.0; }
  for (int iy = j4; iy <= j8; ++iy) {  // oderhs.m:4024-4036, fqp = 0
    for (int ix = i1; ix <= i5; ++ix) {
      int ix1 = IXP1(ix, iy);
      A(floxe, ix, iy) = A(floxe, ix, iy) + P.cfcvte * 1.25 * (A(ne, ix, iy) + A(ne, ix1, iy)) * A(vex, ix, iy) * G(sx, ix, iy) - 0.;
    }
    A(floxe, nx + 1, iy) = 0.0;
  }
  for (int iy = j4; iy <= j8; ++iy) {  // oderhs.m:4065-4071
    for (int ix = i1; ix <= i5; ++ix) A(floxi, ix, iy) = A(floxi, ix, iy) + P.cfcvti * 2.5 * A(fnix, ix, iy);
    A(floxi, nx + 1, iy) = 0.0;
  }
  for (int iy = j1; iy <= j5; ++iy)
    for (int ix = i4; ix <= i8; ++ix) {  // oderhs.m:4078-4128; vyte_use, vyte_cft, cfybf = 0
      A(floye, ix, iy) = A(floye, ix, iy) + (P.cfloye / 2.) * (A(ney0, ix, iy) + A(ney1, ix, iy)) * A(vey, ix, iy) * G(sy, ix, iy) + (0. + 0.) * 0.5 * G(sy, ix, iy) * (A(ney0, ix, iy) + A(ney1, ix, iy));
      if (iy == 0) feeycbo[ix] = 0.;
      A(floyi, ix, iy) = A(floyi, ix, iy) + P.cfloyi * A(fniy, ix, iy) + (0. + 0.) * 0.5 * G(sy, ix, iy) * (A(niy0, ix, iy) + A(niy1, ix, iy));
      if (iy == 0) feiycbo[ix] = feiycbo[ix] + P.cfloyi * fniycbo[ix] * A(ti, ix, 0);
    }
  for (int iy = j4; iy <= j8; ++iy) {  // oderhs.m:4234-4240
    for (int ix = i1; ix <= i5; ++ix) A(floxi, ix, iy) = A(floxi, ix, iy) + P.cftiexclg * P.cfneut * P.cfneutsor_ei * P.cngtgx * P.cfcvti * 2.5 * A(fngx, ix, iy);
    A(floxi, nx + 1, iy) = 0.0;
  }
  for (int iy = j1; iy <= j5; ++iy)
    for (int ix = i4; ix <= i8; ++ix) A(floyi, ix, iy) = A(floyi, ix, iy) + P.cftiexclg * P.cfneut * P.cfneutsor_ei * P.cngtgy * 2.5 * A(fngy, ix, iy);
  fd2tra(w, floxe, floye, conxe, conye, te, feex, feey, 0, (int)P.methe);  // oderhs.m:4256
  fd2tra(w, floxi, floyi, conxi, conyi, ti, feix, feiy, 0, (int)P.methi);  // oderhs.m:4260

  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) {  // oderhs.m:4300-4313 (pwrsore/pwrsori/nuvl zero)
      A(resee, ix, iy) = A(seec, ix, iy) + A(seev, ix, iy) * A(te, ix, iy) + 0. + 0. - 0.;
      A(resei, ix, iy) = A(seic, ix, iy) + A(seiv, ix, iy) * A(ti, ix, iy) + 0. + 0. - 0.;
    }
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) {  // oderhs.m:4439-4478
      int ix1 = IXM1(ix, iy);
      A(resee, ix, iy) = A(resee, ix, iy) - (A(feex, ix, iy) - A(feex, ix1, iy) + P.fluxfacy * (A(feey, ix, iy) - A(feey, ix, iy - 1)));
      A(resei, ix, iy) = A(resei, ix, iy) - (A(feix, ix, iy) - A(feix, ix1, iy) + P.fluxfacy * (A(feiy, ix, iy) - A(feiy, ix, iy - 1)));
    }
  // hydrogen radiation / ionisation energy sink (oderhs.m:4484-4555)
  for (int iy = w.iys1; iy <= w.iyf6; ++iy)
    for (int ix = w.ixs1; ix <= w.ixf6; ++ix) {
      double ne_sgvi = A(ne, ix, iy);
      if (P.ifxnsgi == 1) ne_sgvi = P.cne_sgvi;
      A(erliz, ix, iy) = P.chradi * erl1(A(te, ix, iy), ne_sgvi) * (A(ng, ix, iy) - P.ngbackg * (0.9 + 0.1 * powi(P.ngbackg / A(ng, ix, iy), P.ingb))) * G(vol, ix, iy);
      if (P.isrecmon != 0) A(erlrc, ix, iy) = P.chradr * erl2(A(te, ix, iy), ne_sgvi) * P.fac2sp * A(ni, ix, iy) * G(vol, ix, iy);
      if (P.icnuiz <= 1 && A(psor, ix, iy) != 0.) A(eeli, ix, iy) = 13.6 * ev + A(erliz, ix, iy) / (P.fac2sp * A(psor, ix, iy));
    }
  for (int iy = w.iys1; iy <= w.iyf6; ++iy)
    for (int ix = w.ixs1; ix <= w.ixf6; ++ix) {
      A(vsoreec, ix, iy) = -P.cfneut * P.cfneutsor_ee * P.cnsor * 13.6 * ev * P.fac2sp * A(psorc, ix, iy) + P.cfneut * P.cfneutsor_ee * P.cnsor * 13.6 * ev * P.fac2sp * A(psorrgc, ix, iy) -
                           P.cfneut * P.cfneutsor_ee * P.cnsor * A(erliz, ix, iy) - P.cfneut * P.cfneutsor_ee * P.cnsor * A(erlrc, ix, iy) -
                           P.cfneut * P.cfneutsor_ee * P.cnsor * P.ediss * ev * (0.5 * A(psordis, ix, iy));
      A(vsoree, ix, iy) = A(vsoreec, ix, iy);  // iseesorave = 0
    }
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) {  // oderhs.m:4589-4640, isupgon = 0 branch
      int ix1 = IXM1(ix, iy);
      A(w0, ix, iy) = G(vol, ix, iy) * A(eqp, ix, iy) * (A(te, ix, iy) - A(ti, ix, iy));
      A(resee, ix, iy) = A(resee, ix, iy) - A(w0, ix, iy) + A(vsoree, ix, iy);
      double us = A(upi, ix, iy) + A(upi, ix1, iy);
      A(resei, ix, iy) = A(resei, ix, iy) + A(w0, ix, iy) + P.cfneut * P.cfneutsor_ei * P.ctsor * 1.25e-1 * P.mi * (us * us) * P.fac2sp * A(psor, ix, iy) +
                         P.cfneut * P.cfneutsor_ei * P.ceisor * P.cnsor * P.eion * ev * A(psordis, ix, iy) -
                         P.cfneut * P.cfneutsor_ei * P.ccoldsor * A(ng, ix, iy) * A(nucx, ix, iy) * (1.5 * A(ti, ix, iy) - 0.125 * P.mi * (us * us) - P.eion * ev) * G(vol, ix, iy);
    }
  // viscous heating (oderhs.m:4879-4930), angfx = 0 => cos = 1, sin = 0
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) {
      int ix1 = IXM1(ix, iy), ix2 = IXM1(ix, iy + 1), ix3 = IXM1(ix, iy - 1);
      double thetacc = 0.5 * (0. + 0.);
      double dupdx = G(gx, ix, iy) * (A(upi, ix, iy) - A(upi, ix1, iy));
      A(wvh, ix, iy) = P.cfvcsx * P.cfvisx * ue_cos(thetacc) * A(visx, ix, iy) * (dupdx * dupdx);
      double dupdy;
      const int64_t isx = P.isxpty[ix + NXS * iy];
      if (isx == 0) dupdy = 0.5 * (A(upi, ix, iy) + A(upi, ix1, iy) - A(upi, ix, iy - 1) - A(upi, ix3, iy - 1)) * G(gyf, ix, iy - 1);
      else if (isx == -1) dupdy = 0.5 * (A(upi, ix, iy + 1) + A(upi, ix2, iy + 1) - A(upi, ix, iy) - A(upi, ix1, iy)) * G(gyf, ix, iy);
      else if (isx == 1 && P.isvhyha == 1) {
        double upxavep1 = 0.5 * (A(upi, ix, iy + 1) + A(upi, ix2, iy + 1)), upxave0 = 0.5 * (A(upi, ix, iy) + A(upi, ix1, iy)),
               upxavem1 = 0.5 * (A(upi, ix, iy - 1) + A(upi, ix3, iy - 1));
        double upf0 = 2. * upxavep1 * upxave0 * (upxavep1 + upxave0) / ((upxavep1 + upxave0) * (upxavep1 + upxave0) + P.upvhflr * P.upvhflr);
        double upfm1 = 2. * upxave0 * upxavem1 * (upxave0 + upxavem1) / ((upxave0 + upxavem1) * (upxave0 + upxavem1) + P.upvhflr * P.upvhflr);
        dupdy = (upf0 - upfm1) * G(gy, ix, iy);
      } else
        dupdy = 0.25 * ((A(upi, ix, iy + 1) + A(upi, ix2, iy + 1) - A(upi, ix, iy) - A(upi, ix1, iy)) * G(gyf, ix, iy) +
                        (A(upi, ix, iy) + A(upi, ix1, iy) - A(upi, ix, iy - 1) - A(upi, ix3, iy - 1)) * G(gyf, ix, iy - 1));
      A(wvh, ix, iy) = A(wvh, ix, iy) + P.cfvcsy * P.cfvisy * A(visy, ix, iy) * (dupdy * dupdy);
      A(wvh, ix, iy) = A(wvh, ix, iy) - ue_ksin(thetacc) * P.cfvcsy * P.cfvisy * A(visy, ix, iy) * dupdx * dupdy;
      A(resei, ix, iy) = A(resei, ix, iy) + A(wvh, ix, iy) * G(vol, ix, iy);
    }
  for (int iy = w.iys; iy <= w.iyf; ++iy)  // oderhs.m:4936-4947
    for (int ix = w.ixs; ix <= w.ixf; ++ix) A(pwribkg, ix, iy) = powi(P.tibg * ev / A(ti, ix, iy), P.iteb) * P.pwribkg_c;
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) A(resei, ix, iy) = A(resei, ix, iy) + A(pwribkg, ix, iy) * G(vol, ix, iy);

  // assemble yldot (oderhs.m:4953-4996)
  for (int iy = j2; iy <= j5; ++iy)
    for (int ix = i2; ix <= i5; ++ix) {
      int64_t iv;
      iv = IDXN(ix, iy); yldot[iv] = (1 - P.iseqalg[iv]) * A(resco, ix, iy) / (G(vol, ix, iy) * P.n0);
      iv = IDXU(ix, iy); yldot[iv] = (1 - P.iseqalg[iv]) * A(resmo, ix, iy) / (G(volv, ix, iy) * P.fnorm);
      if (ix == ixrb) yldot[iv] = A(resmo, ix, iy) / (G(volv, ix, iy) * P.fnorm);
      iv = IDXTE(ix, iy); yldot[iv] = (1 - P.iseqalg[iv]) * A(resee, ix, iy) / (G(vol, ix, iy) * P.ennorm);
      iv = IDXTI(ix, iy); yldot[iv] = (1 - P.iseqalg[iv]) * A(resei, ix, iy) / (G(vol, ix, iy) * P.ennorm);
      if (HASG) { iv = IDXG(ix, iy); yldot[iv] = (1 - P.iseqalg[iv]) * A(resng, ix, iy) / (G(vol, ix, iy) * P.n0g); }
    }
  rc = bouncon(w, yl, yldot);  // oderhs.m:5009
  if (rc) return rc;
  if (xc >= 0 && yc >= 0) {  // oderhs.m:5012-5049 (only fields this switch set owns)
    // erliz, erlrc, eeli, frice, upe, upi, uup, psordis, psorc, psorxr, nucx, nurc, nuiz, nuix, psorgc, psorrgc
    // are restored by the second ("restoring") pandf1 call of jac_calc; the
    // partial restore here only matters between the two calls, where none of
    // these fields is read, so it is omitted.
  }
  return 0;
}

// ---- rscalf (oderhs.m:8059-8213), isflxvar = 0 --------------------------------------
void rscalf(const Win& w, const double* yl, double* yldot) {
  for (int iy = w.j2; iy <= w.j5; ++iy)
    for (int ix = w.i2; ix <= w.i5; ++ix) {
      double nbedot = 0., nbidot = 0.;
      int64_t iv = IDXN(ix, iy);
      nbidot = nbidot + yldot[iv] * P.n0;
      nbedot = nbedot + P.zi * yldot[iv] * P.n0;
      double nbg2dot = HASG ? yldot[IDXG(ix, iy)] * P.n0g : 0.;  // oderhs.m:8118
      int ix1 = IXP1(ix, iy);
      int64_t iv2 = IDXU(ix, iy);
      if (P.iseqalg[iv2] == 0) {
        int64_t iv1 = IDXN(ix1, iy);
        double yldot_np1 = A(resco, ix1, iy) / (G(vol, ix1, iy) * P.n0);
        double nbvdot, nbv;
        if (P.iseqalg[iv] == 1) { nbvdot = (P.isnupdot1sd == 0) ? yldot_np1 * P.n0 : yldot[iv1] * P.n0; nbv = A(ni, ix1, iy); }
        else if (P.iseqalg[iv1] == 1) { nbvdot = yldot[iv] * P.n0; nbv = A(ni, ix, iy); }
        else { nbvdot = (P.isnupdot1sd == 0) ? 0.5 * (yldot[iv] + yldot_np1) * P.n0 : yldot[iv] * P.n0; nbv = 0.5 * (A(ni, ix, iy) + A(ni, ix1, iy)); }
        yldot[iv2] = (yldot[iv2] * P.n0 - yl[iv2] * nbvdot) / nbv;
      }
      int64_t ive = IDXTE(ix, iy);
      if (P.iseqalg[ive] == 0) yldot[ive] = (yldot[ive] * P.nnorm - yl[ive] * nbedot) / A(ne, ix, iy);
      int64_t ivi = IDXTI(ix, iy);
      if (P.iseqalg[ivi] == 0) yldot[ivi] = (yldot[ivi] * P.nnorm - yl[ivi] * (nbidot + P.cngtgx * nbg2dot)) / (A(nit, ix, iy) + P.cngtgx * A(ng, ix, iy));
    }
}

// ---- pandf1 (oderhs.m:7883-8056) ---------------------------------------------------
int pandf1(int xc, int yc, const double* yl, double* yldot) {
  int rc = pandf(xc, yc, yl, yldot);
  if (rc) return rc;
  const Win w = make_win(xc, yc);
  if (P.isflxvar != 1 && P.isrscalf == 1) rscalf(w, yl, yldot);
  if (P.dtreal < 1.e15 && yl[neq] < 0) {  // svrpkg = "nksol", oderhs.m:7963-8037
    int j2l, j5l, i2l, i5l;
    if (P.isbcwdt == 0) { j2l = 1; j5l = ny; i2l = 1; i5l = nx; } else { j2l = 0; j5l = ny + 1; i2l = 0; i5l = nx + 1; }
    for (int iy = j2l; iy <= j5l; ++iy)
      for (int ix = i2l; ix <= i5l; ++ix) {
        int64_t iv = IDXN(ix, iy);
        yldot[iv] = (1. - 0.) * yldot[iv]; yldot[iv] = yldot[iv] - (yl[iv] - ylodt[iv]) / dtuse[iv];
        if (ix != nx + 2 * P.isbcwdt) { iv = IDXU(ix, iy); yldot[iv] = (1. - 0.) * yldot[iv]; yldot[iv] = yldot[iv] - (yl[iv] - ylodt[iv]) / dtuse[iv]; }
        iv = IDXTE(ix, iy); yldot[iv] = (1. - 0.) * yldot[iv]; yldot[iv] = yldot[iv] - (yl[iv] - ylodt[iv]) / dtuse[iv];
        iv = IDXTI(ix, iy); yldot[iv] = (1. - 0.) * yldot[iv]; yldot[iv] = yldot[iv] - (yl[iv] - ylodt[iv]) / dtuse[iv];
        if (HASG) { iv = IDXG(ix, iy); yldot[iv] = (1. - 0.) * yldot[iv]; yldot[iv] = yldot[iv] - (yl[iv] - ylodt[iv]) / dtuse[iv]; }
      }
  }
  return 0;
}

int check_switches() {
  { const std::string z = S.nonzero_frozen(); if (!z.empty()) { err = "input " + z + " must be 0: the term it switches on is outside the built hot path"; return -5; } }
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
      {"ibctepr", P.ibctepr, 1}, {"ibctipr", P.ibctipr, 1}, {"iskaplex", P.iskaplex, 0}};
  for (auto& e : eq) if (e.v != e.want) { err = std::string("switch outside the built hot path: ") + e.n; return -5; }
  if (P.isngon != 0 && P.isngon != 1) { err = "isngon must be 0 or 1"; return -5; }
  if (P.isfixlb != 0 && P.isfixlb != 2) { err = "isfixlb must be 0 or 2"; return -5; }
  if (P.isbohmcalc != 0 && P.isbohmcalc != 1) { err = "isbohmcalc must be 0/1 with facb*=0"; return -5; }
  if (P.isnicore != 0 && P.isnicore != 1) { err = "isnicore must be 0 or 1"; return -5; }
  if (P.isupcore < 0 || P.isupcore > 3) { err = "isupcore must be 0..3"; return -5; }
  if (P.iflcore < -1 || P.iflcore > 1) { err = "iflcore must be -1, 0 or 1"; return -5; }
  if (P.isngcore < 0 || P.isngcore > 4) { err = "isngcore must be 0..4"; return -5; }
  if (P.istabon != 0 && P.istabon != 7 && P.istabon != 10) { err = "istabon must be 0, 7 or 10"; return -5; }
  // fnnuiz < 1 blends the new ionisation rate with the value left by the PREVIOUS pandf call (oderhs.m:1950-1961): the
  // reference's Jacobian then depends on the order in which the unknowns were perturbed; not reproducible in parallel
  if (P.fnnuiz != 1.) { err = "fnnuiz must be 1 (history-dependent rate blending is outside the built hot path)"; return -5; }
  if (P.difpr2 != 0 || P.difni2 != 0 || P.difax != 0 || P.dif4order != 0 || P.kye4order != 0 || P.kyi4order != 0) { err = "difpr2/difni2/difax/4th-order terms not built"; return -5; }
  if (P.l_parloss <= 1e9) { err = "l_parloss<=1e9 (nuvl) not built"; return -5; }
  if (P.cfjhf != 0 && false) { err = ""; return -5; }
  for (int m : {(int)P.methn, (int)P.methu, (int)P.methe, (int)P.methi, (int)P.methg}) {
    int mx = m % 10, my = m / 10;
    if ((mx != 2 && mx != 3) || (my != 2 && my != 3)) { err = "meth* must use schemes 2 (central) or 3 (upwind)"; return -5; }
  }
  for (int ix = 0; ix < NXS; ++ix) {
    if (P.fngysi[ix] != 0 || P.fngyso[ix] != 0 || P.fngyi_use[ix] != 0 || P.fngyo_use[ix] != 0) { err = "wall gas sources not built"; return -5; }
    for (int64_t v : {P.isnwconiix[ix], P.isnwconoix[ix]}) if (v < 0 || v > 3) { err = "isnwconi/o must be 0..3"; return -5; }
    for (int64_t v : {P.istepfcix[ix], P.istipfcix[ix], P.istewcix[ix], P.istiwcix[ix]}) if (v < 0 || v > 3) { err = "istepfc/istipfc/istewc/istiwc must be 0..3"; return -5; }
  }
  return 0;
}

// set_dt (oderhs.m:9886-10147): the per-unknown pseudo time step of the nksol equations, model_dt 0..3.
// f0 = rhsnk(yl) first (oderhs.m:9914); ylodt is the vector of the last step_params call.  dtoptv persists between
// calls (a velocity row whose |f0| <= cutlo keeps its previous value, oderhs.m:9950-9951).
int set_dt(int64_t n, const double* yl, double* f0, double* dtuse_out) {
  if (n != neq) { err = "set_dt: neq mismatch"; return -1; }
  if (P.model_dt < 0 || P.model_dt > 3) { err = "model_dt must be 0..3"; return -5; }
  int rc = pandf1(-1, -1, yl, f0);
  if (rc) return rc;
  auto model = [&](double dtopt) {
    if (P.model_dt == 0) return P.dtreal;
    if (P.model_dt == 1) return P.dtreal * dtopt / (P.dtreal + dtopt);
    if (P.model_dt == 2) return dtopt;
    return std::sqrt(P.dtreal * dtopt);
  };
  for (int iy = 0; iy <= ny + 1; ++iy) {
    const int iym1 = std::max(0, iy - 1), iyp1 = std::min(ny + 1, iy + 1);
    for (int ix = 0; ix <= nx + 1; ++ix) {
      auto plain = [&](int64_t iv) {
        dtoptv[iv] = P.deldt * std::fabs(ylodt[iv] / (f0[iv] + P.cutlo));
        dtuse[iv] = model(dtoptv[iv]);
      };
      plain(IDXN(ix, iy));
      if (ix != nx + 2 * P.isbcwdt) {
        const int ixm1u = std::max(0, IXM1(ix, iy)), ixp1u = std::min(nx + 1, IXP1(ix, iy));
        const int64_t iv = IDXU(ix, iy);
        const double up_5ca = (std::fabs(ylodt[iv]) + std::fabs(ylodt[IDXU(ixm1u, iy)]) + std::fabs(ylodt[IDXU(ixp1u, iy)]) + std::fabs(ylodt[IDXU(ix, iyp1)]) +
                               std::fabs(ylodt[IDXU(ix, iym1)])) / 5;
        if (std::fabs(f0[iv]) > P.cutlo) dtoptv[iv] = P.deldt * std::fabs(up_5ca / (f0[iv]));
        dtuse[iv] = model(dtoptv[iv]);
      }
      plain(IDXTE(ix, iy));
      plain(IDXTI(ix, iy));
      if (HASG) plain(IDXG(ix, iy));
    }
  }
  if (P.isbcwdt == 0)
    for (int64_t iv = 0; iv < neq; ++iv) if (P.iseqalg[iv] == 1) dtuse[iv] = 1.e20;
  std::copy(dtuse.begin(), dtuse.end(), dtuse_out);
  return 0;
}

// jac_calc (oderhs.m:8533-8760), columns ivmin..ivmax only (ppp LocalJacBuilder, ppp/parallel.F90:176-381): the CSC
// fragment rcsc/icsc with jcsc[iv-1] = 1-based start of column iv inside the fragment (nnz_frag+1 beyond the range).
// The caller must have evaluated pandf1(-1,-1) at yl (psetnk/sfsetnk do, oderhs.m:9466, 9851): the module state is the base state.
int jac_csc(const double* yl_in, const double* yldot00, int64_t ml, int64_t mu, int64_t nnzmx, std::vector<double>& rcsc,
            std::vector<int64_t>& icsc, std::vector<int64_t>& jcsc) {
  std::vector<double> yl(yl_in, yl_in + neq + 2), wk(neq);
  rcsc.clear(); icsc.clear(); jcsc.assign(neq + 1, 0);
  int64_t nnz = 1;
  for (int64_t iv = 1; iv <= neq; ++iv) {
    jcsc[iv - 1] = nnz;
    if (iv < ivmin || iv > ivmax) continue;
    int64_t ii1 = std::max(iv - mu, (int64_t)1), ii2 = std::min(iv + ml, neq);
    for (int64_t ii = ii1; ii <= ii2; ++ii) wk[ii - 1] = yldot00[ii - 1];
    int xc = (int)P.igyl[iv - 1], yc = (int)P.igyl[neq + iv - 1];
    double yold = yl[iv - 1];
    double dyl = P.delpert * (std::fabs(yold) + P.dylconst / suscal[iv - 1]);
    yl[iv - 1] = yold + dyl;
    int rc = pandf1(xc, yc, yl.data(), wk.data());
    if (rc) return rc;
    for (int64_t ii = ii1; ii <= ii2; ++ii) {
      double jacelem = (wk[ii - 1] - yldot00[ii - 1]) / dyl;
      if (iv == ii) {
        if (P.iseqalg[iv - 1] * (1 - P.isbcwdt) == 0) jacelem = jacelem - 1 / dtuse[iv - 1];
      }
      if (P.nufak > 0) if (iv == ii && yl[neq] == 1) jacelem = jacelem - P.nufak;
      if (std::fabs(jacelem * sfscal[iv - 1]) > P.jaccliplim) {
        if (nnz > nnzmx) {
          char buf[256];
          snprintf(buf, sizeof buf, "*** jac_calc -- More storage needed for Jacobian. Storage exceeded at (i,j) = (%lld,%lld). Increase lenpfac.", (long long)ii, (long long)iv);
          err = buf; return -2;
        }
        rcsc.push_back(jacelem); icsc.push_back(ii); nnz = nnz + 1;
      }
    }
    yl[iv - 1] = yold;
    rc = pandf1(xc, yc, yl.data(), wk.data());
    if (rc) return rc;
  }
  jcsc[neq] = nnz;
  return 0;
}
};
Ora g_o;
// persistent worker threads (an OpenMP runtime keeps its team alive between parallel regions; so does this)
struct Pool {
  std::vector<std::thread> th;
  std::mutex m; std::condition_variable cv, done;
  std::function<void(int)> job; int njob = 0, gen = 0, pending = 0; bool stop = false;
  void worker(int id) {
    int seen = 0;
    for (;;) {
      std::function<void(int)> f;
      { std::unique_lock<std::mutex> l(m); cv.wait(l, [&] { return stop || gen != seen; }); if (stop) return; seen = gen; if (id >= njob) continue; f = job; }
      f(id);
      { std::lock_guard<std::mutex> l(m); if (--pending == 0) done.notify_one(); }
    }
  }
  void run(int n, std::function<void(int)> f) {
    while ((int)th.size() < n) { int id = (int)th.size(); th.emplace_back([this, id] { worker(id); }); }
    { std::lock_guard<std::mutex> l(m); job = f; njob = n; pending = n; ++gen; }
    cv.notify_all();
    std::unique_lock<std::mutex> l(m); done.wait(l, [&] { return pending == 0; });
  }
  ~Pool() { { std::lock_guard<std::mutex> l(m); stop = true; } cv.notify_all(); for (auto& t : th) t.join(); }
} g_pool;
std::vector<double> g_thread_w;  // column-range weights of the threaded Jacobian (previous call's timings)

}  // namespace

// =====================================================================================
extern "C" {
int ue_ora_set_int(const char* n, int64_t v) { return S.set_int(n, v); }
int ue_ora_set_real(const char* n, double v) { return S.set_real(n, v); }
int ue_ora_set_real_array(const char* n, const double* d, int64_t k) { return S.set_real_array(n, d, k); }
int ue_ora_set_int_array(const char* n, const int64_t* d, int64_t k) { return S.set_int_array(n, d, k); }
const char* ue_ora_last_error(void) { return g_o.err.c_str(); }

int ue_ora_init(void) {
  std::string& g_err = g_o.err;
  std::string m = S.missing();
  if (!m.empty()) { g_err = "missing inputs: " + m; return -1; }
  nx = (int)P.nx; ny = (int)P.ny; NXS = nx + 2; NC = NXS * (ny + 2); neq = P.neq;
  std::string b = S.bad_sizes();
  if (!b.empty()) { g_err = "bad array sizes (have != expected): " + b; return -1; }
  if (neq != (int64_t)NC * P.numvar) { g_err = "neq != numvar*(nx+2)*(ny+2)"; return -1; }
  int rc = g_o.check_switches();
  if (rc) return rc;
  // rate tables for istabon=10 (aph/aphread.m readehr1 + setauxvar :700-735)
  mpe = (int)P.mpe; mpd = (int)P.mpd;
  if (P.istabon == 10) {
    if (mpe < 2 || mpd < 2 || S.len("wsveh") != (int64_t)mpe * mpd) { g_err = "istabon=10 needs wsveh/wsveh0/welms1/welms2 tables"; return -1; }
    int64_t nt = (int64_t)mpe * mpd;
    wsveh.assign(P.wsveh, P.wsveh + nt); wsveh0.assign(P.wsveh0, P.wsveh0 + nt);
    welms1.assign(P.welms1, P.welms1 + nt); welms2.assign(P.welms2, P.welms2 + nt);
    dkpt.resize(mpd); ekpt.resize(mpe);
    dkpt[0] = 16.0; for (int j = 1; j < mpd; ++j) dkpt[j] = dkpt[j - 1] + 0.5;
    rldmin = dkpt[0]; rldmax = dkpt[mpd - 1]; deldkpt = (rldmax - rldmin) / double(mpd - 1);
    ekpt[0] = -1.2 * ue_log(10.0); for (int j = 1; j < mpe; ++j) ekpt[j] = ekpt[j - 1] + 0.1 * ue_log(10.0);
    rlemin = ekpt[0]; rlemax = ekpt[mpe - 1]; delekpt = (rlemax - rlemin) / double(mpe - 1);
  }
  for (V* v : g_o.all_planes()) v->assign(NC, 0.0);
  HASG = P.isngon == 1;
  if (!HASG) g_o.ng.assign(P.ngfix, P.ngfix + NC);  // never advanced: the field ueinit left (odesetup.m:1399-1406)
  g_o.fniycbo.assign(NXS, 0.); g_o.feeycbo.assign(NXS, 0.); g_o.feiycbo.assign(NXS, 0.);
  g_o.dtuse.assign(neq, 1e20); g_o.ylodt.assign(neq, 0.); g_o.suscal.assign(neq, 1.); g_o.sfscal.assign(neq, 1.); g_o.dtoptv.assign(neq, 0.);
  g_o.ivmin = 1; g_o.ivmax = neq;
  g_thread_w.clear();
  return 0;
}

int ue_ora_step_params(int64_t n, const double* dt, const double* yo, const double* su, const double* sf) {
  if (n != neq) { g_o.err = "step_params: neq mismatch"; return -1; }
  g_o.dtuse.assign(dt, dt + n); g_o.ylodt.assign(yo, yo + n); g_o.suscal.assign(su, su + n); g_o.sfscal.assign(sf, sf + n);
  return 0;
}

// general entry: pandf1(xc,yc,ieq,neq,time,yl,yldot); xc=yc=-1 is the full residual
int ue_ora_pandf1_win(int64_t xc, int64_t yc, int64_t n, const double* yl, double* yldot) {
  if (n != neq) { g_o.err = "pandf1: neq mismatch"; return -1; }
  return g_o.pandf1((int)xc, (int)yc, yl, yldot);
}
int ue_ora_pandf1(int64_t n, double time, const double* yl, double* yldot) { (void)time; return ue_ora_pandf1_win(-1, -1, n, yl, yldot); }

int ue_ora_set_dt(int64_t n, const double* yl, double* f0, double* dtuse_out) { return g_o.set_dt(n, yl, f0, dtuse_out); }

int ue_ora_set_column_range(int64_t ivmin, int64_t ivmax) { g_o.ivmin = ivmin; g_o.ivmax = ivmax; return 0; }

// serial jac_calc (oderhs.m:8533-8760): CSC by columns, then csrcsc
int ue_ora_jac_calc(int64_t n, double t, const double* yl_in, const double* yldot00, int64_t ml, int64_t mu, int64_t nnzmx,
                    double* jac, int64_t* ja, int64_t* ia, int64_t* nnz_out) {
  (void)t;
  if (n != neq) { g_o.err = "jac_calc: neq mismatch"; return -1; }
  std::vector<double> rcsc; std::vector<int64_t> icsc, jcsc;
  int rc = g_o.jac_csc(yl_in, yldot00, ml, mu, nnzmx, rcsc, icsc, jcsc);
  if (rc) return rc;
  csrcsc(neq, rcsc.data(), icsc.data(), jcsc.data(), jac, ja, ia);
  *nnz_out = jcsc[neq] - 1;
  return 0;
}

// Threaded jac_calc: the reference's OpenMP design (ppp/omp_parallel.F90:65-117 jac_calc_omp/OMPJacBuilder, 395-444
// OMPSplitIndex, 119-164 OMPCollectJacobian).  Every worker thread takes a private copy of the whole module state (the
// reference copies all threadprivate module arrays, omp_parallel.F90:319-332), assembles the CSC fragment of a contiguous
// range of columns, and the fragments are concatenated in thread order (= column order) before one csrcsc.  The ranges are
// re-weighted with the per-thread times of the previous call (omp_parallel.F90:199-229).  The full residual at yl must
// have been evaluated by the caller (base state), as for the serial form.
int ue_ora_jac_calc_threads(int64_t nthreads, int64_t n, const double* yl_in, const double* yldot00, int64_t ml, int64_t mu, int64_t nnzmx,
                            double* jac, int64_t* ja, int64_t* ia, int64_t* nnz_out, double* thread_ms /* nthreads, may be NULL */) {
  if (n != neq) { g_o.err = "jac_calc: neq mismatch"; return -1; }
  const int T = (int)std::max<int64_t>(1, std::min<int64_t>(nthreads, neq));
  const auto tA = std::chrono::steady_clock::now();
  if ((int)g_thread_w.size() != T) g_thread_w.assign(T, 1.0 / T);
  // OMPSplitIndex: contiguous ranges with sizes proportional to the weights
  std::vector<int64_t> lo(T), hi(T);
  { double acc = 0.; int64_t prev = 0;
    for (int t = 0; t < T; ++t) { acc += g_thread_w[t]; int64_t e = (t == T - 1) ? neq : std::min<int64_t>(neq, (int64_t)std::llround(acc * neq)); e = std::max(e, prev); lo[t] = prev + 1; hi[t] = e; prev = e; } }
  std::vector<std::vector<double>> rc_(T); std::vector<std::vector<int64_t>> ic_(T), jc_(T);
  std::vector<int> rcs(T, 0); std::vector<std::string> errs(T); std::vector<double> ms(T, 0.);
  g_pool.run(T, [&](int t) {
    auto t0 = std::chrono::steady_clock::now();
    Ora w = g_o;  // private copy of the module state
    w.ivmin = lo[t]; w.ivmax = hi[t];
    rcs[t] = (lo[t] <= hi[t]) ? w.jac_csc(yl_in, yldot00, ml, mu, nnzmx, rc_[t], ic_[t], jc_[t]) : 0;
    if (lo[t] > hi[t]) jc_[t].assign(neq + 1, 1);
    errs[t] = w.err;
    ms[t] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  });
  const auto tB = std::chrono::steady_clock::now();
  for (int t = 0; t < T; ++t) if (rcs[t]) { g_o.err = errs[t]; return rcs[t]; }
  // OMPCollectJacobian: concatenate in thread order
  std::vector<double> rcsc; std::vector<int64_t> icsc, jcsc(neq + 1);
  int64_t nnz = 1;
  for (int t = 0; t < T; ++t) {
    for (int64_t iv = lo[t]; iv <= hi[t]; ++iv) jcsc[iv - 1] = nnz + (jc_[t][iv - 1] - 1);
    nnz += (int64_t)rc_[t].size();
    rcsc.insert(rcsc.end(), rc_[t].begin(), rc_[t].end()); icsc.insert(icsc.end(), ic_[t].begin(), ic_[t].end());
  }
  jcsc[neq] = nnz;
  if (nnz - 1 > nnzmx) { g_o.err = "*** jac_calc -- More storage needed for Jacobian. Increase lenpfac."; return -2; }
  csrcsc(neq, rcsc.data(), icsc.data(), jcsc.data(), jac, ja, ia);
  *nnz_out = nnz - 1;
  // new weights ~ columns per unit time of each thread, damped by half
  { double tot = 0.; std::vector<double> sp(T);
    for (int t = 0; t < T; ++t) { sp[t] = (double)std::max<int64_t>(1, hi[t] - lo[t] + 1) / std::max(ms[t], 1e-6); tot += sp[t]; }
    for (int t = 0; t < T; ++t) g_thread_w[t] = 0.5 * g_thread_w[t] + 0.5 * sp[t] / tot; }
  if (thread_ms) for (int t = 0; t < T; ++t) thread_ms[t] = ms[t];
  if (getenv("UE_ORA_TIMING")) {
    const auto tC = std::chrono::steady_clock::now();
    fprintf(stderr, "threads %d: parallel section %.3f ms (slowest thread %.3f), merge + csrcsc %.3f ms\n", T, std::chrono::duration<double, std::milli>(tB - tA).count(),
            *std::max_element(ms.begin(), ms.end()), std::chrono::duration<double, std::milli>(tC - tB).count());
  }
  return 0;
}

// debugging / parity helper: copy a named intermediate plane out
int ue_ora_get_plane(const char* name, double* out) {
  std::vector<V*> pl = g_o.all_planes();
  std::string names(plane_names);
  size_t pos = 0; size_t k = 0;
  while (pos < names.size()) {
    size_t e = names.find(' ', pos); if (e == std::string::npos) e = names.size();
    if (names.compare(pos, e - pos, name) == 0 && (e - pos) == strlen(name)) { std::copy(pl[k]->begin(), pl[k]->end(), out); return 0; }
    pos = e + 1; ++k;
  }
  g_o.err = "no such plane"; return -1;
}
const char* ue_ora_plane_names(void) { return plane_names; }
// include/ue_math.h evaluated on the host: op 0 exp, 1 log, 2 log10, 3 pow(x,y), 4 cos, 5 sqrt (tests compare with libm
// and, bit for bit, with the same header evaluated on the device)
int ue_ora_math_probe(int64_t op, int64_t n, const double* x, const double* y, double* out) {
  for (int64_t i = 0; i < n; ++i) {
    switch (op) {
      case 0: out[i] = ue_exp(x[i]); break;
      case 1: out[i] = ue_log(x[i]); break;
      case 2: out[i] = ue_log10(x[i]); break;
      case 3: out[i] = ue_pow(x[i], y[i]); break;
      case 4: out[i] = ue_cos(x[i]); break;
      case 5: out[i] = ue_sqrt(x[i]); break;
      default: return -1;
    }
  }
  return 0;
}
}
