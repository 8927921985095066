// uedge_b200/csrc/ue_device.cuh — device-side physics of the pandf1 / jac_calc hot path (sm_100a, FP64).
//
// Design: the reference evaluates `pandf` as ~60 loop nests over
// persistent module arrays.  Here one residual is FOUR phases, each a per-cell
// device function with no intra-phase dependence between cells:
//
//   phase 0  state unpack + pointwise sources      (convsr_vo, convsr_aux pointwise part,
//                                                   oderhs.m:1909-2029, 4484-4555, 4936-4941)
//   phase 1  gradients, face interpolants, drift-free velocities, transport
//            coefficients and ALL face fluxes of the scalar equations
//                                                  (convert.m:582-784, oderhs.m:1138-1792,
//                                                   2718-3017, 3187-3319, 3923-4261, neudifpg 6126-6341)
//   phase 2  momentum fluxes, flux divergences, volume sources, guard-cell
//            (boundary-condition) rows            (oderhs.m:2471-2579, 3407-3456, 3475-3866,
//                                                   4300-4996, boundary.m:102-2800)
//   phase 3  rscalf + time-step term               (oderhs.m:7953-8037, 8059-8213)
//
// The same device functions serve two drivers through the accessor template:
//   Acc<false>  full residual: fields live in HBM planes (SoA, ix fastest), one thread per cell;
//   Acc<true>   Jacobian: private copies of the four cells a perturbation can change, every other read
//               falls through to the base planes; one launch per phase over all unknowns (ue_gpu.cu).
// Index windows, recompute ranges and the frozen-term gates are the reference's
// (oderhs.m:868-1019) so that the value-dependent sparsity pattern is reproduced.
//
// Arithmetic is written expression-for-expression like the reference and compiled
// with -fmad=false so that "unchanged input => bit-identical output" holds between the
// two drivers (exact zeros of the finite difference) and against the CPU oracle.
#pragma once
#include <cstdint>

#include "ue_math.h"
#include "ue_param_store.hpp"

#define UE_NV 5  // row slots per cell inside the kernels: ni, up, te, ti, ng (convert.m:33-152 ordering)
// Unknowns per cell in the CALLER's vectors (yl, yldot, iseqalg, dtuse, ...): 5, or 4 when isngon = 0 (no ng unknown: the
// atom density is then the frozen input plane `ngfix` and row slot 4 is computed but never stored).
#define NVX ((int)D.numvar)

enum Plane : int {
  // phase-0 outputs
  PL_NI = 0, PL_UP, PL_TE, PL_TI, PL_NG, PL_NUIZ, PL_NURC, PL_NUCX, PL_ERLIZ, PL_ERLRC, PL_PWRIBKG,
  // phase-1 outputs
  PL_GPIX, PL_GPEX, PL_GPIY, PL_GPEY, PL_NIY0, PL_NIY1, PL_VY, PL_UPE, PL_VEY, PL_FRICE, PL_VISX,
  PL_FNIX, PL_FNIY, PL_FNGX, PL_FNGY, PL_FEEX, PL_FEEY, PL_FEIX, PL_FEIY,
  // phase-2 output needed by rscalf
  PL_RESCO,
  // feex on the cut face of a half-space problem as a WINDOWED evaluation computes it (vex = 0 there, see f_upe_pre)
  PL_FEEXC,
  PL_COUNT
};

struct DevTables {  // istabon=10 abscissae (aph/aphread.m:700-735)
  double ekpt[64], dkpt[16];
  double rlemin, rlemax, rldmin, rldmax, delekpt, deldkpt;
  int mpe, mpd;
  int rarebc;  // derived at init: some column selects one of the boundary-condition options of guard_rare
  int iscut;  // derived at init: isfixlb == 2 && iysptrx1 > 0 (half-space problem with a core region: see d_cut)
};

__constant__ UeParams D;      // scalars + device pointers to static planes/lines
__constant__ DevTables DT;

// ---- index window (oderhs.m:868-1019) ---------------------------------------------------
struct Win {
  int xc, yc;
  int i1, i2, i3, i4, i5, i6, i8;
  int j1, j2, j3, j4, j5, j6, j7, j8;
  int openbox, xcnearlb, xcnearrb, xccuts;
};

__host__ __device__ inline Win make_win(const UeParams& P, int xc, int yc) {
  Win w; w.xc = xc; w.yc = yc;
  const int nx = (int)P.nx, ny = (int)P.ny;
  const int xlinc = (int)P.xlinc, xrinc = (int)P.xrinc, yinc = (int)P.yinc;
#define mx(a, b) ((a) > (b) ? (a) : (b))
#define mn(a, b) ((a) < (b) ? (a) : (b))
  if (xc < 0 || ((0 <= yc) && (yc - yinc <= 0) && P.isjaccorall == 1)) {
    w.i1 = 0; w.i2 = 1; w.i3 = 0; w.i4 = 0; w.i5 = nx; w.i6 = nx + 1; w.i8 = nx + 1;
  } else {
    w.i1 = mx(0, xc - xlinc - 1); w.i2 = mx(1, xc - xlinc); w.i3 = xc - xlinc; w.i4 = mx(0, xc - xlinc);
    w.i5 = mn(nx, xc + xrinc); w.i6 = mn(nx + 1, xc + xrinc + 1); w.i8 = mn(nx + 1, xc + xrinc);
  }
  if (yc < 0) {
    w.j1 = 0; w.j2 = 1; w.j3 = 0; w.j4 = 0; w.j5 = ny; w.j6 = ny + 1; w.j7 = ny + 1; w.j8 = ny + 1;
  } else {
    w.j1 = mx(0, yc - yinc - 1); w.j2 = mx(1, yc - yinc); w.j3 = yc - yinc; w.j4 = mx(0, yc - yinc);
    w.j5 = mn(ny, yc + yinc); w.j6 = mn(ny + 1, yc + yinc); w.j7 = yc + yinc; w.j8 = mn(ny + 1, yc + yinc);
  }
  w.xccuts = 0;
  if ((xc - xlinc <= P.ixpt1 + 1) && (xc + xrinc + 1 >= P.ixpt1) && (yc - yinc <= P.iysptrx1) && (P.iysptrx1 > 0)) w.xccuts = 1;
  if ((xc - xlinc <= P.ixpt2 + 1) && (xc + xrinc + 1 >= P.ixpt2) && (yc - yinc <= P.iysptrx2) && (P.iysptrx2 > 0)) w.xccuts = 1;
  if (w.xccuts) { w.i1 = 0; w.i2 = 1; w.i3 = 0; w.i4 = 0; w.i5 = nx; w.i6 = nx + 1; w.i8 = nx + 1; }
  if (xc < 0) w.openbox = 1;
  else if (w.xccuts) w.openbox = 1;
  else if ((0 <= yc) && (yc <= yinc)) w.openbox = 1;
  else w.openbox = 0;
  w.xcnearlb = (((xc - xlinc <= P.ixlb) && (xc + xrinc >= P.ixlb)) || xc < 0) ? 1 : 0;
  w.xcnearrb = (((xc - xlinc <= P.ixrb + 1) && (xc + xrinc >= P.ixrb)) || xc < 0) ? 1 : 0;
#undef mx
#undef mn
  return w;
}

#ifdef __CUDACC__
// ---- field accessor ------------------------------------------------------------------------
// Acc<false>: fields are the HBM planes.  Acc<true> (Jacobian): a perturbation of cell C0=(xc,yc) can
// change phase-0 fields only at C0 and phase-1 fields only at C0, its west/east neighbours (row yc
// connectivity) and its south neighbour, so only those four cells have private copies; every other
// read falls through to the base planes.  `resco` (phase 2) is kept per candidate row of the unknown.
template <bool WIN>
struct Acc {
  double* base;   // [PL_COUNT][NC] planes in HBM (base state)
  int NXS, NC;
  // WIN only
  double* priv;   // private copies of C0, Cw, Ce, Cs: plane pl of slot k at priv[pl * ps + k * ks]
  int ps, ks;
  int xc, yc, xw, xe;
  double* rres;      // resco of this unknown's candidate rows (one per candidate cell)
  const int* rmask;  // rows-written mask of the candidate cells (bit 8 <=> resco written)
  int rself, reast;  // candidate index of the row being evaluated / of its east neighbour IXP1 (-1: not a candidate);
                     // PL_RESCO is only written for the own row (p2_n) and read for the east neighbour (phase 3)
  // slot() and get() are written with selects, not branches: a read then is ONE generic load whose address was
  // chosen arithmetically, so the compiler can keep many reads of a role function in flight (with branches every
  // read became its own control-flow region and the loads serialised at L2 latency).
  __device__ __forceinline__ int slot(int ix, int iy) const {
    const bool r0 = iy == yc;
    int k = -1;
    k = (r0 && ix == xe) ? 2 : k;
    k = (r0 && ix == xw) ? 1 : k;
    k = (r0 && ix == xc) ? 0 : k;
    k = (iy == yc - 1 && ix == xc) ? 3 : k;
    return k;
  }
  __device__ __forceinline__ double get(int pl, int ix, int iy) const {
    const double* g = base + ((size_t)pl * NC + ix + NXS * iy);
    if (WIN) {
      if (pl == PL_RESCO) {
        const int idx = reast >= 0 ? reast : 0;
        const bool use = reast >= 0 && (rmask[idx] & 0x100);
        const double* p = use ? rres + idx : g;
        return *p;
      }
      const int k = slot(ix, iy);
      const double* p = k >= 0 ? priv + ((size_t)pl * ps + k * ks) : g;
      return *p;
    }
    return *g;
  }
  __device__ __forceinline__ void set(int pl, int ix, int iy, double v) const {
    if (WIN) {
      if (pl == PL_RESCO) {
        if (rself >= 0) rres[rself] = v;
      } else {
        const int k = slot(ix, iy);
        if (k >= 0) priv[(size_t)pl * ps + k * ks] = v;
      }
      return;
    }
    base[(size_t)pl * NC + ix + NXS * iy] = v;
  }
};

#define GG(a_, ix, iy) D.a_[(ix) + NXS * (iy)]
#define IXP1(ix, iy) ((int)D.ixp1[(ix) + NXS * (iy)])
#define IXM1(ix, iy) ((int)D.ixm1[(ix) + NXS * (iy)])

__device__ __forceinline__ double d_ave(double t0, double t1) { return 2 * t0 * t1 / (D.cutlo + t0 + t1); }
__device__ __forceinline__ double d_sgn(double a, double b) { return copysign(fabs(a), b); }
__device__ __forceinline__ double d_powi(double x, int64_t n) {
  double r = 1.0;
  while (n > 0) { if (n & 1) r *= x; x *= x; n >>= 1; }
  return r;
}
__device__ __forceinline__ double d_upwind(double f, double p1, double p2) { return fmax(f, 0.0) * p1 + fmin(f, 0.0) * p2; }
__device__ __forceinline__ int in_rng(int v, int lo, int hi) { return v >= lo && v <= hi; }

// ---- hydrogen rates (aph/aphrates.m) ----------------------------------------------------
__device__ inline double d_table(const double* __restrict__ w, double tej, double dens) {
  const double zloge = ue_log(tej / D.ev);
  const double rle = fmax(DT.rlemin, fmin(zloge, DT.rlemax));
  const double zlogd = ue_log10(dens);
  const double rld = fmax(DT.rldmin, fmin(zlogd, DT.rldmax));
  int je = (int)((rle - DT.rlemin) / DT.delekpt) + 1; je = min(je, DT.mpe - 1);
  int jd = (int)((rld - DT.rldmin) / DT.deldkpt) + 1; jd = min(jd, DT.mpd - 1);
  const double fje = (rle - DT.ekpt[je - 1]) / (DT.ekpt[je] - DT.ekpt[je - 1]);
  const double fjd = (rld - DT.dkpt[jd - 1]) / (DT.dkpt[jd] - DT.dkpt[jd - 1]);
  const int mpe = DT.mpe;
  const double r11 = ue_log(__ldg(&w[(je - 1) + mpe * (jd - 1)])), r12 = ue_log(__ldg(&w[(je - 1) + mpe * jd]));
  const double r21 = ue_log(__ldg(&w[je + mpe * (jd - 1)])), r22 = ue_log(__ldg(&w[je + mpe * jd]));
  const double r1 = r11 + fjd * (r12 - r11);
  const double r2 = r21 + fjd * (r22 - r21);
  return ue_exp(r1 + fje * (r2 - r1));
}
// istabon = 7 (the package default): R.B. Campbell's polynomial fits in x = log10(ne), y = log10(Te[eV]) (aph/aphrates.m:1133-1300)
__device__ __noinline__ double d_sionf(double temp, double den) {
  const double x = fmin(22.e0, ue_log10(den)), y = ue_log10(temp);
  const double ain = -49.05905 + 2.51313783 * x - 0.049159714 * x * x;
  const double bin = 41.1855162 - 2.3298672 * x + 4.24769144e-2 * x * x;
  const double cin = -32.798921 + 1.72102919 * x - 0.038692357 * x * x;
  const double din = 27.370466 - 1.6824361 * x + 0.0462317894 * x * x;
  const double ein = -7.9990454 + 0.127573157 * x - 6.3586911e-3 * x * x;
  const double gin = -4.5832951 + 0.776264783 * x - 1.8866089e-2 * x * x;
  const double hin = 3.08056833 - 0.39114789 * x + 9.86833304e-3 * x * x;
  const double riin = -0.4648639 + 0.0551428018 * x - 1.404213e-3 * x * x;
  return ue_pow(10., ain + bin * y + cin * y * y + din * y * y * y + ein * y * y * y * y + gin * y * y * y * y * y + hin * y * y * y * y * y * y +
                         riin * y * y * y * y * y * y * y);
}
__device__ __noinline__ double d_srecf(double temp, double den) {
  const double x = fmin(22.e0, ue_log10(den)), y = ue_log10(temp);
  const double ar = -0.4575652 - 2.144012 * x + 6.7072142e-2 * x * x - 1.391667e-4 * x * x * x;
  const double br = -121.8401 + 18.001822 * x - 0.8679488 * x * x + 1.33165e-2 * x * x * x;
  const double cr = 80.897256 - 13.29602 * x + 0.71881414 * x * x - 0.0126549 * x * x * x;
  const double dr = 56.406823 - 7.301996 * x + 0.29339793 * x * x - 3.50898e-3 * x * x * x;
  const double er = -55.73559 + 7.9634283 * x - 0.370274 * x * x + 5.567961e-3 * x * x * x;
  const double gr = 10.866692 - 1.584193 * x + 0.07563791 * x * x - 1.177562e-3 * x * x * x;
  return ue_pow(10., ar + br * y + cr * y * y + dr * y * y * y + er * y * y * y * y + gr * y * y * y * y * y);
}
__device__ __noinline__ double d_svradp_sionfl(double x, double y) {
  const double ai = -275.845 + 37.010817 * x - 1.788045 * x * x + 0.029078333 * x * x * x;
  const double bi = 2200.9478 - 326.1153 * x + 16.148655 * x * x - 0.2660702 * x * x * x;
  const double ci = -2.935221e3 + 4.3757698e2 * x - 21.73964 * x * x + 0.358962 * x * x * x;
  const double di = 1604.1466 - 239.6959 * x + 11.923707 * x * x - 0.1970501 * x * x * x;
  const double ei = -390.8635 + 58.474495 * x - 2.910997 * x * x + 0.048133829 * x * x * x;
  const double gi = 35.012574 - 5.24202 * x + 0.26109962 * x * x - 4.319238e-3 * x * x * x;
  return ue_pow(10., ai + bi * y + ci * y * y + di * y * y * y + ei * y * y * y * y + gi * y * y * y * y * y);
}
__device__ __noinline__ double d_svradp(double temp, double den) {
  const double x = fmin(22.e0, ue_log10(den)), y = ue_log10(temp);
  const double ym = fmin(2.e0, y);  // etai frozen above 100 eV
  const double ae = 2860.4173 - 610.2452 * x + 48.275821 * x * x - 1.687994 * x * x * x + 0.02201375 * x * x * x * x;
  const double be = 10612.067 - 2046.397 * x + 147.73914 * x * x - 4.729973 * x * x * x + 0.056671796 * x * x * x * x;
  const double ce = -4.231708e4 + 8494.6102 * x - 639.0226 * x * x + 21.350311 * x * x * x - 0.2673466 * x * x * x * x;
  const double de = -8.385144e3 + 1887.6244 * x - 157.8502 * x * x + 5.820501 * x * x * x - 0.07992837 * x * x * x * x;
  const double ee = 3.938282e4 - 8.131339e3 * x + 628.8119 * x * x - 21.58636 * x * x * x + 0.27756029 * x * x * x * x;
  const double ge = -1.038281e4 + 2.1349333e3 * x - 164.4201 * x * x + 5.6210487 * x * x * x - 0.07197622 * x * x * x * x;
  const double etai = (ue_pow(10., ae + be * ym + ce * ym * ym + de * ym * ym * ym + ee * ym * ym * ym * ym + ge * ym * ym * ym * ym * ym)) / d_svradp_sionfl(x, ym);
  return fmax(0.e0, (13.6e0 + etai)) * 1.602e-19 * d_svradp_sionfl(x, y);
}
__device__ inline double d_rsa(double tej, double dens) {
  if (D.istabon == 0) { const double a = tej / (10 * D.ev); return 3.0e-14 * a * a / (3.0 + a * a); }
  if (D.istabon == 7) return d_sionf(tej / D.ev, dens);
  return d_table(D.wsveh, tej, dens);
}
__device__ inline double d_rra(double tej, double dens) {
  if (D.istabon == 0) return 0.;
  if (D.istabon == 7) return d_srecf(tej / D.ev, dens);
  return d_table(D.wsveh0, tej, dens);
}
__device__ inline double d_rcx(double t0) { const double a = 3 * t0 / (10 * D.ev); return 1.7e-14 * ue_pow(a, 0.333); }
__device__ inline double d_rqa0(double tej) { const double a = tej / (10 * D.ev); return D.erad * D.ev * 3.0e-14 * a * a / (3.0 + a * a); }
__device__ inline double d_erl1(double tej, double dens) {
  if (D.istabon == 0) return (d_rqa0(tej) - 13.6 * D.ev * d_rsa(tej, dens)) * dens;
  if (D.istabon == 7) return (d_svradp(tej / D.ev, dens) - 13.6 * D.ev * d_rsa(tej, dens)) * dens;  // aph/aphrates.m:28-31, 601-606
  return d_table(D.welms1, tej, dens);
}
__device__ inline double d_erl2(double tej, double dens) {
  if (D.istabon == 0 || D.istabon == 7) return (13.6 * D.ev + 1.5 * tej) * dens * d_rra(tej, dens);
  return d_table(D.welms2, tej, dens);
}

// unknown index (0-based) of variable k in cell (ix,iy)
__device__ __forceinline__ int64_t d_iv(int ix, int iy, int k, int NXS) { return ((int64_t)(ix + NXS * iy)) * NVX + k; }

// ============================================================================================
// phase 0 — convsr_vo + pointwise part of convsr_aux + volumetric rates at one cell
// ============================================================================================
template <bool WIN>
__device__ void phase0_cell(const Acc<WIN>& a, const double* ycell /* the cell's UE_NV entries of yl */, int ix, int iy, int* errflag) {
  const int NXS = a.NXS;
  const double ev = D.ev;
  const double ni = ycell[0] * D.n0;                       // convert.m:250
  const double up = ycell[1] * D.fnorm / (D.mi * D.n0);    // convert.m:352-356
  double te = ycell[2] * D.ennorm / (1.5 * D.nnorm);       // convert.m:281-282
  te = fmax(te, D.temin * ev);
  const double ng = D.isngon == 1 ? ycell[4] * D.n0g : D.ngfix[ix + NXS * iy];  // convert.m:285-287; isngon = 0: never advanced
  double ti = ycell[3] * D.ennorm / (1.5 * D.nnorm);       // convert.m:310-311
  ti = fmax(ti, D.temin * ev);
  if (ni < 0) atomicOr(errflag, 1);                         // convert.m:318-322
  if (ng < 0) atomicOr(errflag, 2);                         // convert.m:323-327
  const double ne = 0. + D.zi * ni;
  const double vol = GG(vol, ix, iy);
  // ionisation, recombination, charge exchange (oderhs.m:1943-1993); rtau = 0
  double nuiz, nurc, nucx;
  if (D.icnuiz == 0) {
    double ne_sgvi = ne;
    if (D.ifxnsgi == 1) ne_sgvi = D.cne_sgvi;
    nuiz = D.chioniz * ne * (d_rsa(te, ne_sgvi) + D.sigvi_floor);
    if (WIN) nuiz = D.fnnuiz * nuiz + (1 - D.fnnuiz) * a.base[(size_t)PL_NUIZ * a.NC + ix + NXS * iy];
  } else nuiz = D.cnuiz;
  if (D.isrecmon == 1) {
    nurc = D.cfrecom * ne * d_rra(te, ne);
    if (WIN) nurc = D.fnnuiz * nurc + (1 - D.fnnuiz) * a.base[(size_t)PL_NURC * a.NC + ix + NXS * iy];
  } else nurc = 0.;
  if (D.icnucx == 0) {
    const double t0 = fmax(ti, D.temin * ev);
    const double t1 = t0 / (D.mi / D.mp);
    nucx = ni * d_rcx(t1);
  } else if (D.icnucx == 1) nucx = D.cnucx;
  else {
    const double t0 = fmax(ti, D.temin * ev);
    nucx = sqrt(t0 / D.mi) * D.sigcx * (ni + D.rnn2cx * ng);
  }
  // hydrogen radiation (oderhs.m:4486-4508)
  double ne_sgvi = ne;
  if (D.ifxnsgi == 1) ne_sgvi = D.cne_sgvi;
  const double erliz = D.chradi * d_erl1(te, ne_sgvi) * (ng - D.ngbackg * (0.9 + 0.1 * d_powi(D.ngbackg / ng, D.ingb))) * vol;
  double erlrc = 0.;
  if (D.isrecmon != 0) erlrc = D.chradr * d_erl2(te, ne_sgvi) * D.fac2sp * ni * vol;
  a.set(PL_NI, ix, iy, ni); a.set(PL_UP, ix, iy, up); a.set(PL_TE, ix, iy, te); a.set(PL_TI, ix, iy, ti); a.set(PL_NG, ix, iy, ng);
  a.set(PL_NUIZ, ix, iy, nuiz); a.set(PL_NURC, ix, iy, nurc); a.set(PL_NUCX, ix, iy, nucx);
  a.set(PL_ERLIZ, ix, iy, erliz); a.set(PL_ERLRC, ix, iy, erlrc);
  if (ix >= 1 && ix <= D.nx && iy >= 1 && iy <= D.ny)  // oderhs.m:4936-4941
    a.set(PL_PWRIBKG, ix, iy, d_powi(D.tibg * ev / ti, D.iteb) * D.pwribkg_c);
}

// pointwise derived fields (convert.m:514-557), recomputed where needed
template <bool WIN> __device__ __forceinline__ double f_ne(const Acc<WIN>& a, int ix, int iy) { return 0. + D.zi * a.get(PL_NI, ix, iy); }
template <bool WIN> __device__ __forceinline__ double f_nm(const Acc<WIN>& a, int ix, int iy) { return a.get(PL_NI, ix, iy) * D.mi; }
template <bool WIN> __device__ __forceinline__ double f_pri(const Acc<WIN>& a, int ix, int iy) { return a.get(PL_NI, ix, iy) * a.get(PL_TI, ix, iy); }
template <bool WIN> __device__ __forceinline__ double f_pre(const Acc<WIN>& a, int ix, int iy) { return f_ne(a, ix, iy) * a.get(PL_TE, ix, iy); }
template <bool WIN> __device__ __forceinline__ double f_pr(const Acc<WIN>& a, int ix, int iy) { return (0. + f_pri(a, ix, iy)) + f_pre(a, ix, iy); }
template <bool WIN> __device__ __forceinline__ double f_zeff(const Acc<WIN>& a, int ix, int iy) {
  return (0. + (D.zi * D.zi) * a.get(PL_NI, ix, iy)) / f_ne(a, ix, iy);
}
template <bool WIN> __device__ __forceinline__ double f_tg(const Acc<WIN>& a, int ix, int iy) {
  return (1 - D.istgcon) * D.rtg2ti * a.get(PL_TI, ix, iy) + D.istgcon * D.tgas * D.ev;  // convert.m:551-553 (istgcon > -1e-20)
}
template <bool WIN> __device__ __forceinline__ double f_pg(const Acc<WIN>& a, int ix, int iy) { return a.get(PL_NG, ix, iy) * f_tg(a, ix, iy); }
template <bool WIN> __device__ __forceinline__ double f_nuix(const Acc<WIN>& a, int ix, int iy) {
  return D.fnuizx * a.get(PL_NUIZ, ix, iy) + D.fnucxx * a.get(PL_NUCX, ix, iy);
}
template <bool WIN> __device__ __forceinline__ double f_uu(const Acc<WIN>& a, int ix, int iy) {  // oderhs.m:1608,1668
  const int NXS = a.NXS;
  return GG(rrv, ix, iy) * a.get(PL_UP, ix, iy) + 0. - 0. - 0.;
}
// Half-space problem with a core region (isfixlb = 2, iysptrx1 > 0): after the velocities, the electron drift velocity and
// the neutral fluxes are formed, the reference sets gpex, frice, ex, upe, gpix, frici, uu, upi to zero on the cut face
// ix = ixpt2, iy <= iysptrx1 (oderhs.m:2447-2466); every later reader sees zeros there.  (Every windowed evaluation that
// recomputes those fields there is an "xccuts" window spanning all ix, so it re-applies the zeroing: the fields are
// zero in every state a reader can see.)  gpix, gpex, frice have only later readers and are stored as zeros; upe, uu
// and upi (= up) also have earlier readers (vex, the neutral convection, upe itself), so their later readers go through
// the *_cut accessors below.
__device__ __forceinline__ bool d_cut(int ix, int iy) { return DT.iscut && ix == (int)D.ixpt2 && iy <= (int)D.iysptrx1; }
template <bool WIN> __device__ __forceinline__ double f_uu_cut(const Acc<WIN>& a, int ix, int iy) { return d_cut(ix, iy) ? 0. : f_uu(a, ix, iy); }
template <bool WIN> __device__ __forceinline__ double f_upi_cut(const Acc<WIN>& a, int ix, int iy) { return d_cut(ix, iy) ? 0. : a.get(PL_UP, ix, iy); }
template <bool WIN> __device__ __forceinline__ double f_upe_cut(const Acc<WIN>& a, int ix, int iy) { return d_cut(ix, iy) ? 0. : a.get(PL_UPE, ix, iy); }
// upe BEFORE the zeroing, i.e. the value vex = upe*rrv is formed from (oderhs.m:1744-1792).  The reference's arrays are
// persistent: upi on the cut was zeroed by the previous evaluation, and a Jacobian-mode evaluation refreshes upi only at
// (xc,yc) and (ixm1(xc,yc),yc) (oderhs.m:1577-1643) while it recomputes upe -- from that upi -- over its whole window.
// So a windowed evaluation forms vex on the cut from upi = 0 unless the cut cell is one of those two cells, whereas the
// full evaluation (yldot00) used upi = up.  With up != 0 on the cut the electron-energy rows next to it therefore differ
// from yldot00 in EVERY window that recomputes them, and the reference's Jacobian carries those finite-difference
// artefacts as entries; they are reproduced here (f_upe_pre, PL_FEEXC, and the extra candidate rows in ue_lists.hpp).
template <bool WIN> __device__ __forceinline__ double f_upe_pre(const Acc<WIN>& a, int ix, int iy) {
  if (WIN && d_cut(ix, iy)) {
    const int NXS = a.NXS;
    const bool refreshed = iy == a.yc && (ix == a.xc || (a.xc > 0 && IXM1(a.xc, a.yc) == ix));
    if (!refreshed) return 0.;
  }
  return a.get(PL_UPE, ix, iy);
}
// feex(ix,iy) for a row evaluation: on the cut face a windowed evaluation sees the vex = 0 variant unless it recomputed the
// face itself (private cell)
template <bool WIN> __device__ __forceinline__ double f_feex(const Acc<WIN>& a, int ix, int iy) {
  if (WIN && d_cut(ix, iy) && a.slot(ix, iy) < 0) return a.get(PL_FEEXC, ix, iy);
  return a.get(PL_FEEX, ix, iy);
}
template <bool WIN> __device__ __forceinline__ double f_visy(const Acc<WIN>& a, int ix, int iy) {  // oderhs.m:2784
  return (D.fcdif * D.travis + 0.) * f_nm(a, ix, iy) + 4 * 0.;
}
// psor family (oderhs.m:1966-1974, 2017-2029)
template <bool WIN> __device__ __forceinline__ void f_psor(const Acc<WIN>& a, int ix, int iy, double& psor, double& psorxr, double& psordis) {
  const int NXS = a.NXS;
  const double ng = a.get(PL_NG, ix, iy), nuiz = a.get(PL_NUIZ, ix, iy), vol = GG(vol, ix, iy);
  const double psorbgg = D.ngbackg * ((0.9 + 0.1 * d_powi(D.ngbackg / ng, D.ingb))) * nuiz * vol;
  const double psorgc = -ng * nuiz * vol + psorbgg;
  psor = -psorgc;
  psordis = D.cfdiss * psor;
  psorxr = -a.get(PL_NI, ix, iy) * a.get(PL_NURC, ix, iy) * vol;
}
// orthogonal-mesh y-face interpolants (convert.m:422-482 with fx0=1, other weights 0)
template <bool WIN, typename F> __device__ __forceinline__ double f_ilog(const Acc<WIN>& a, F f, int ix, int iy, int k) {
  // zero-weighted neighbour terms add an exact 0 for positive finite fields and are not evaluated
  return ue_exp(1. * ue_log(f(a, ix, iy + k)));
}
template <bool WIN, typename F> __device__ __forceinline__ double f_ilin(const Acc<WIN>& a, F f, int ix, int iy, int k) {
  return 1. * f(a, ix, iy + k);
}
template <bool WIN> struct Fld {
  static __device__ __forceinline__ double ni(const Acc<WIN>& a, int ix, int iy) { return a.get(PL_NI, ix, iy); }
  static __device__ __forceinline__ double te(const Acc<WIN>& a, int ix, int iy) { return a.get(PL_TE, ix, iy); }
  static __device__ __forceinline__ double ti(const Acc<WIN>& a, int ix, int iy) { return a.get(PL_TI, ix, iy); }
  static __device__ __forceinline__ double ng(const Acc<WIN>& a, int ix, int iy) { return a.get(PL_NG, ix, iy); }
  static __device__ __forceinline__ double tg(const Acc<WIN>& a, int ix, int iy) { return f_tg(a, ix, iy); }
  static __device__ __forceinline__ double pri(const Acc<WIN>& a, int ix, int iy) { return f_pri(a, ix, iy); }
  static __device__ __forceinline__ double pg(const Acc<WIN>& a, int ix, int iy) { return f_pg(a, ix, iy); }
};

// membership in the x-lists "do ix = ixm1(is,j), min(nx,ie), inc" of convsr_aux (convert.m:583-592)
__device__ inline bool in_xlist(int ix, int xc, int jinc, int jstart, int NXS) {
  const int nx = (int)D.nx;
  const int d = xc - IXM1(xc, jinc);
  int inc = max(1, abs(d)); if (d < 0) inc = -inc;
  const int first = IXM1(xc, jstart), last = min(nx, xc);
  if (inc > 0) { for (int i = first; i <= last; i += inc) if (i == ix) return true; }
  else { for (int i = first; i >= last; i += inc) if (i == ix) return true; }
  return false;
}

// ============================================================================================
// phase 1 — everything that lives on one cell / its east and north faces, split into ROLES that
// are independent of each other inside a sub-phase so that different warps can take them:
//   sub-phase 1a:  p1_xpart (x-gradients, thermal force, upe)   p1_ypart (y-face values, vy, vey)
//                  p1_visx  (parallel viscosity)
//   sub-phase 1b:  p1_fx (fngx,fnix)  p1_fy (fngy,fniy)  p1_exe (feex)  p1_exi (feix)  p1_ey (feey,feiy)
// 1b reads only same-cell outputs of 1a.  A role that needs a quantity owned by another role of the
// same sub-phase recomputes it with the same device function (bit-identical).
// ============================================================================================
struct P1Rng { bool r16, r15, rfx, rfy; };
__device__ __forceinline__ P1Rng p1_ranges(const Win& w, int ix, int iy) {
  P1Rng r;
  r.r16 = in_rng(ix, w.i1, w.i6) && in_rng(iy, w.j1, w.j6);  // [i1..i6]x[j1..j6]
  r.r15 = in_rng(ix, w.i1, w.i6) && in_rng(iy, w.j1, w.j5);  // [i1..i6]x[j1..j5]
  r.rfx = in_rng(ix, w.i1, w.i5) && in_rng(iy, w.j4, w.j8);  // x-face fluxes
  r.rfy = in_rng(ix, w.i4, w.i8) && in_rng(iy, w.j1, w.j5);  // y-face fluxes
  return r;
}
// Coulomb logarithm on the east face of (ix,iy) (oderhs.m:1138-1155)
template <bool WIN> __device__ inline double f_loglambda(const Acc<WIN>& a, int ix, int iy) {
  const int NXS = a.NXS;
  const int ix1 = IXP1(ix, iy);
  const double teev = 0.5 * (a.get(PL_TE, ix, iy) + a.get(PL_TE, ix1, iy)) / D.ev;
  const double nexface = 0.5 * (f_ne(a, ix, iy) + f_ne(a, ix1, iy));
  if (D.islnlamcon == 1) return D.lnlam;
  if (teev < 50.) return 23.4 - 1.15 * ue_log10(1.e-6 * nexface) + 3.45 * ue_log10(teev);
  return 25.3 - 1.15 * ue_log10(1.e-6 * nexface) + 2.33167537087122e+00 * ue_log10(teev);
}

// ---- 1a: x-part (convert.m:608-618, 736-749; oderhs.m:1516-1534, 1738-1762) ------------------------
template <bool WIN>
__device__ void p1_xpart(const Acc<WIN>& a, const Win& w, int ix, int iy) {
  const int NXS = a.NXS;
  const int nx = (int)D.nx;
  const double ev = D.ev, cutlo = D.cutlo;
  const int ix1 = IXP1(ix, iy);
  bool do_xg;
  if (!WIN) do_xg = (ix <= nx);
  else do_xg = (iy == w.yc) && in_xlist(ix, w.xc, iy, iy, NXS);
  const double gxf = GG(gxf, ix, iy), rrv = GG(rrv, ix, iy);
  if (do_xg) {
    a.set(PL_GPIX, ix, iy, d_cut(ix, iy) ? 0. : (f_pri(a, ix1, iy) - f_pri(a, ix, iy)) * gxf);
    a.set(PL_GPEX, ix, iy, d_cut(ix, iy) ? 0. : (f_pre(a, ix1, iy) - f_pre(a, ix, iy)) * gxf);
  }
  const P1Rng r = p1_ranges(w, ix, iy);
  if (r.r16) {
    const double ni = a.get(PL_NI, ix, iy), ni_e = a.get(PL_NI, ix1, iy), te = a.get(PL_TE, ix, iy), te_e = a.get(PL_TE, ix1, iy);
    const double ne = f_ne(a, ix, iy), ne_e = f_ne(a, ix1, iy);
    const double gtex = (te_e - te) * gxf;  // convert.m:741
    const double nbarx = 0.5 * (ne + ne_e);
    const double ltmax = fmin(fabs(te / (rrv * gtex + cutlo)), GG(lcone, ix, iy));
    const double lmfpe = 2e16 * ((te / ev) * (te / ev)) / ne;
    const double flxlimf = D.flalftf * ltmax / (D.flalftf * ltmax + lmfpe);
    a.set(PL_FRICE, ix, iy, d_cut(ix, iy) ? 0. : -D.cthe * flxlimf * nbarx * rrv * gtex + 0.);
    double upe = 0. + a.get(PL_UP, ix, iy) * D.zi * 0.5 * (ni + ni_e);
    upe = (upe - 0.) / (0.5 * (ne + ne_e));
    a.set(PL_UPE, ix, iy, upe);
  }
}

// ---- 1a: y-part (convert.m:627-784; oderhs.m:1174-1320, 1466-1468, 1729-1792) -----------------------
template <bool WIN>
__device__ void p1_ypart(const Acc<WIN>& a, const Win& w, int ix, int iy) {
  const int NXS = a.NXS;
  const int ny = (int)D.ny;
  const P1Rng r = p1_ranges(w, ix, iy);
  if (iy <= ny) {
    bool do_yf;
    if (!WIN) do_yf = true;
    else {
      const int jlo = max(w.yc - 1, 0), jhi = min(w.yc, ny);
      do_yf = (iy >= jlo && iy <= jhi) && (in_xlist(ix, w.xc, w.yc, w.yc, NXS) || ix == IXP1(w.xc, iy));
    }
    const double dynog = GG(dynog, ix, iy);
    double niy0 = a.get(PL_NIY0, ix, iy), niy1 = a.get(PL_NIY1, ix, iy), gpiy = a.get(PL_GPIY, ix, iy), gpey = a.get(PL_GPEY, ix, iy);
    const double tey0 = f_ilin(a, Fld<WIN>::te, ix, iy, 0), tey1 = f_ilin(a, Fld<WIN>::te, ix, iy, 1);
    if (do_yf) {
      niy0 = f_ilog(a, Fld<WIN>::ni, ix, iy, 0); niy1 = f_ilog(a, Fld<WIN>::ni, ix, iy, 1);
      const double priy0 = f_ilog(a, Fld<WIN>::pri, ix, iy, 0), priy1 = f_ilog(a, Fld<WIN>::pri, ix, iy, 1);
      gpiy = (priy1 - priy0) / dynog;
      const double ney0 = 0. + D.zi * niy0, ney1 = 0. + D.zi * niy1;
      gpey = (ney1 * tey1 - ney0 * tey0) / dynog;
      a.set(PL_NIY0, ix, iy, niy0); a.set(PL_NIY1, ix, iy, niy1); a.set(PL_GPIY, ix, iy, gpiy); a.set(PL_GPEY, ix, iy, gpey);
    }
    double vy = a.get(PL_VY, ix, iy);
    if (r.r15) {
      const double gpry = (0. + gpiy) + gpey;
      const double pr_n = f_pr(a, ix, iy + 1), pr_c = f_pr(a, ix, iy);
      const double gtey = (tey1 - tey0) / dynog;
      double vydd = D.vcony + 0. + 0. - (D.difpr + 0.) * (2 * gpry / (pr_n + pr_c) - 3.0 * gtey / (tey1 + tey0));
      const double difnimix = D.fcdif * D.difni + 0.;
      vydd = vydd - 1. * difnimix * (2 * (1 - D.isvylog) * ((niy1 - niy0) / dynog) / (niy1 + niy0) + D.isvylog * (ue_log(niy1) - ue_log(niy0)) / dynog);
      vy = vydd;
      a.set(PL_VY, ix, iy, vy);
    }
    double vey = a.get(PL_VEY, ix, iy);
    if (r.r16) vey = 0.;
    if (r.r15) {
      const double ney0 = 0. + D.zi * niy0, ney1 = 0. + D.zi * niy1;
      vey = 0. + vy * D.zi * 0.5 * (niy0 + niy1);
      vey = (vey - 0.) / (0.5 * (ney0 + ney1));
    }
    if (r.r16) a.set(PL_VEY, ix, iy, vey);
  } else {  // iy == ny+1
    if (in_rng(ix, w.i1, w.i6)) a.set(PL_VY, ix, iy, 0.0);     // oderhs.m:1466-1468
    if (r.r16) a.set(PL_VEY, ix, iy, 0.);                       // oderhs.m:1729-1733
    if (in_rng(ix, w.i4, w.i8)) a.set(PL_FNIY, ix, iy, 0.0);   // oderhs.m:3315-3317
  }
}

// ---- 1a: parallel viscosity (oderhs.m:2718-2787) -------------------------------------------------------
template <bool WIN>
__device__ void p1_visx(const Acc<WIN>& a, const Win& w, int ix, int iy) {
  const int NXS = a.NXS;
  const double ev = D.ev;
  const P1Rng r = p1_ranges(w, ix, iy);
  if (!r.r16) return;
  const double ni = a.get(PL_NI, ix, iy), ti = a.get(PL_TI, ix, iy);
  const double gxf = GG(gxf, ix, iy), gx = GG(gx, ix, iy);
  const double loglambda = f_loglambda(a, ix, iy);
  const double tvw = (D.zi * D.zi) / sqrt((D.mi + D.mi) / (2 * D.mp));
  const double wsum = 0.0 + tvw * ni;
  const double ctaui = 2.1e13 / (loglambda * (D.zi * D.zi));
  const double tv2 = ctaui / (ev * sqrt(ev));
  const double aa = (D.convis == 0) ? fmax(ti, D.temin * ev) : D.afix * ev;
  const double rr = GG(rr, ix, iy), vol = GG(vol, ix, iy);
  const double visxtmp = tv2 * D.coef * rr * rr * aa * aa * sqrt(aa) * ni / wsum;
  double visx = D.parvis * visxtmp + 0. * f_nm(a, ix, iy);
  const int ixw = IXM1(ix, iy);
  const double t0 = fmax(ti, D.temin * ev);
  const double mfl = D.flalfv * f_nm(a, ix, iy) * rr * vol * gx * (t0 / D.mi);
  double csh;
  if (D.isgxvon == 0) csh = visx * vol * gx * gx;
  else csh = visx * vol * gx * 2 * gxf * GG(gxf, ixw, iy) / (gxf + GG(gxf, ixw, iy));
  const double msh = fabs(csh * (f_upi_cut(a, ixw, iy) - f_upi_cut(a, ix, iy)));
  visx = visx / ue_pow(1 + ue_pow(msh / (mfl + 1.e-20 * msh), D.flgamv), 1 / D.flgamv);
  a.set(PL_VISX, ix, iy, visx);
}

// ---- x-face particle fluxes (neudifpg oderhs.m:6126-6225 + fd2tra; oderhs.m:3197-3244) -----------------
template <bool WIN> __device__ inline double f_fngx(const Acc<WIN>& a, int ix, int iy) {
  const int NXS = a.NXS;
  const double ev = D.ev;
  const int ixlb = (int)D.ixlb, ixrb = (int)D.ixrb;
  const int ix1 = IXP1(ix, iy);
  const int methgx = (int)(D.methg % 10);
  const double ng = a.get(PL_NG, ix, iy), ng_e = a.get(PL_NG, ix1, iy);
  const double gxf = GG(gxf, ix, iy), gx = GG(gx, ix, iy), gx_e = GG(gx, ix1, iy), sx = GG(sx, ix, iy);
  const double uu = f_uu(a, ix, iy);
  const double tg = f_tg(a, ix, iy), tg_e = f_tg(a, ix1, iy);
  const double ngxface = 0.5 * (ng + ng_e);
  const double t0 = fmax(tg, D.temin * ev), t1 = fmax(tg_e, D.temin * ev);
  const double vtn = sqrt(t0 / D.mg), vtnp = sqrt(t1 / D.mg);
  const double nu1 = f_nuix(a, ix, iy) + vtn / D.lgmax, nu2 = f_nuix(a, ix1, iy) + vtnp / D.lgmax;
  const double tgf = 0.5 * (tg + tg_e);
  const double flalfgx_adj = D.flalfgxa[ix] * (1. + d_powi(D.cflbg * D.ngbackg / ngxface, D.inflbg));
  const double qfl = flalfgx_adj * sx * (vtn + vtnp) * D.rt8opi * (ng * gx + ng_e * gx_e) / (8 * (gx + gx_e));
  const double csh = (1 - D.isgasdc) * D.cdifg * sx * gxf * (1 / D.mg) * d_ave(1. / nu1, 1. / nu2) + D.isgasdc * sx * gxf * D.difcng / tgf +
                     (D.rld2dxg * D.rld2dxg) * sx * (1 / gxf) * 0.5 * (a.get(PL_NUIZ, ix, iy) + a.get(PL_NUIZ, ix1, iy)) / tgf;
  double qtgf = D.alftng * D.fgtdx[ix] * sx * d_ave(gx / nu1, gx_e / nu2) * (vtn * vtn - vtnp * vtnp);
  const double vygtan = 0.;
  qtgf = qtgf - vygtan * sx;
  double nconv = 2.0 * (ng * ng_e) / (ng + ng_e);
  if (methgx != 2) nconv = ng * 0.5 * (1 + d_sgn(1., qtgf)) + ng_e * 0.5 * (1 - d_sgn(1., qtgf));
  const double pg = f_pg(a, ix, iy), pg_e = f_pg(a, ix1, iy);
  const double qsh = csh * (pg - pg_e) + qtgf * nconv;
  double qr = fabs(qsh / qfl);
  if (ix == ixlb || ix == ixrb) { qr = D.gcfacgx * qr; qtgf = D.gcfacgx * qtgf; }
  double conxg = csh / ue_pow(1 + ue_pow(qr, D.flgamg), 1 / D.flgamg);
  if (D.isdifxg_aug == 1) conxg = csh * (1 + qr);
  double floxg = (qtgf / tgf) / ue_pow(1 + ue_pow(qr, D.flgamg), 1 / D.flgamg);
  floxg = floxg + D.cngflox * sx * uu / tgf;
  if (methgx == 2) return floxg * (pg_e + pg) / 2. - conxg * (pg_e - pg);
  return d_upwind(floxg, pg, pg_e) - conxg * (pg_e - pg);
}
template <bool WIN> __device__ inline double f_fnix(const Acc<WIN>& a, int ix, int iy) {
  const int NXS = a.NXS;
  const int ix1 = IXP1(ix, iy);
  const double ni = a.get(PL_NI, ix, iy), ni_e = a.get(PL_NI, ix1, iy), sx = GG(sx, ix, iy);
  const double uu = f_uu_cut(a, ix, iy);
  const int methnx = (int)(D.methn % 10);
  double t2;
  if (methnx == 2) t2 = (ni + ni_e) / 2;
  else t2 = (uu >= 0.) ? ni : ni_e;
  double fnix = D.cnfx * uu * sx * t2;
  const double r1 = D.nlimix * ni / ni_e, r2 = D.nlimix * ni_e / ni;
  fnix = fnix / sqrt(1 + r1 * r1 + r2 * r2);
  return fnix;
}
template <bool WIN>
__device__ void p1_fx(const Acc<WIN>& a, const Win& w, int ix, int iy) {
  const P1Rng r = p1_ranges(w, ix, iy);
  if (!r.rfx) return;
  a.set(PL_FNGX, ix, iy, f_fngx(a, ix, iy));
  a.set(PL_FNIX, ix, iy, f_fnix(a, ix, iy));
}

// ---- y-face particle fluxes (oderhs.m:6239-6328 + fd2tra; oderhs.m:3264-3312) -----------------------------
template <bool WIN> __device__ inline double f_fngy(const Acc<WIN>& a, int ix, int iy) {
  const int NXS = a.NXS;
  const int ny = (int)D.ny;
  const double ev = D.ev;
  const int methgy = (int)(D.methg / 10);
  const double dynog = GG(dynog, ix, iy);
  const double vy = a.get(PL_VY, ix, iy);
  const double ng = a.get(PL_NG, ix, iy), ng_n = a.get(PL_NG, ix, iy + 1);
  const double ngy0 = f_ilog(a, Fld<WIN>::ng, ix, iy, 0), ngy1 = f_ilog(a, Fld<WIN>::ng, ix, iy, 1);
  const double tg = f_tg(a, ix, iy), tg_n = f_tg(a, ix, iy + 1);
  const double gy = GG(gy, ix, iy), gy_n = GG(gy, ix, iy + 1), sy = GG(sy, ix, iy);
  const double ngyface = 0.5 * (ng + ng_n);
  const double t0 = fmax(tg, D.tgmin * ev), t1 = fmax(tg_n, D.tgmin * ev);
  const double vtn = sqrt(t0 / D.mg), vtnp = sqrt(t1 / D.mg);
  const double nu1 = f_nuix(a, ix, iy) + vtn / D.lgmax, nu2 = f_nuix(a, ix, iy + 1) + vtnp / D.lgmax;
  const double tgf = 0.5 * (tg + tg_n);
  const double flalfgy_adj = D.flalfgya[iy] * (1. + d_powi(D.cflbg * D.ngbackg / ngyface, D.inflbg));
  double qfl = flalfgy_adj * sy * (vtn + vtnp) * D.rt8opi * (ngy0 * gy + ngy1 * gy_n) / (8 * (gy + gy_n));
  if (iy == 0) qfl = flalfgy_adj * sy * (vtn + vtnp) * D.rt8opi * (ngy0 + ngy1) / 8.;
  const double csh = (1 - D.isgasdc) * (D.cdifg * sy / dynog) * (1 / D.mg) * d_ave(1. / nu1, 1. / nu2) + D.isgasdc * sy * D.difcng / (dynog * tgf) +
                     (D.rld2dyg * D.rld2dyg) * sy * dynog * 0.5 * (a.get(PL_NUIZ, ix, iy) + a.get(PL_NUIZ, ix, iy + 1)) / tgf;
  double qtgf = D.alftng * D.fgtdy[iy] * sy * d_ave(gy / nu1, gy_n / nu2) * (vtn * vtn - vtnp * vtnp);
  double nconv = 2.0 * (ngy0 * ngy1) / (ngy0 + ngy1);
  if (methgy != 2) nconv = ngy0 * 0.5 * (1 + d_sgn(1., qtgf)) + ngy1 * 0.5 * (1 - d_sgn(1., qtgf));
  const double pgy0 = f_ilog(a, Fld<WIN>::pg, ix, iy, 0), pgy1 = f_ilog(a, Fld<WIN>::pg, ix, iy, 1);
  const double qsh = csh * (pgy0 - pgy1) + qtgf * nconv;
  double qr = fabs(qsh / qfl);
  if (iy == 0) { qr = D.gcfacgy * qr; qtgf = D.gcfacgy * qtgf; }
  if (iy == ny) { qr = D.gcfacgy * qr; qtgf = D.gcfacgy * qtgf; }
  double conyg = csh / ue_pow(1 + ue_pow(qr, D.flgamg), 1 / D.flgamg);
  if (D.isdifyg_aug == 1) conyg = csh * (1 + qr);
  double floyg = (qtgf / tgf) / ue_pow(1 + ue_pow(qr, D.flgamg), 1 / D.flgamg);
  floyg = floyg + D.cngfloy * sy * vy / tgf;
  const double pg = f_pg(a, ix, iy), pg_n = f_pg(a, ix, iy + 1);
  if (methgy == 2) return floyg * (pg_n + pg) / 2. - conyg * (pg_n - pg);
  return d_upwind(floyg, pg, pg_n) - conyg * (pg_n - pg);
}
template <bool WIN> __device__ inline double f_fniy(const Acc<WIN>& a, int ix, int iy) {
  const int NXS = a.NXS;
  const double vy = a.get(PL_VY, ix, iy), sy = GG(sy, ix, iy);
  const double niy0 = a.get(PL_NIY0, ix, iy), niy1 = a.get(PL_NIY1, ix, iy);
  const double ni = a.get(PL_NI, ix, iy), ni_n = a.get(PL_NI, ix, iy + 1);
  const int methny = (int)(D.methn / 10);
  double t2;
  if (methny == 2) t2 = (niy0 + niy1) / 2;
  else t2 = (vy >= 0.) ? niy0 : niy1;
  double fniy = D.cnfy * vy * sy * t2;
  if (vy * (ni - ni_n) < 0.) {
    const double r1 = D.nlimiy / ni_n, r2 = D.nlimiy / ni;
    fniy = fniy / (1 + r1 * r1 + r2 * r2);
  }
  return fniy;
}
template <bool WIN>
__device__ void p1_fy(const Acc<WIN>& a, const Win& w, int ix, int iy) {
  const P1Rng r = p1_ranges(w, ix, iy);
  if (!r.rfy) return;
  a.set(PL_FNGY, ix, iy, f_fngy(a, ix, iy));
  a.set(PL_FNIY, ix, iy, f_fniy(a, ix, iy));
}

// ---- x-face electron energy flux (oderhs.m:2850-2874, 2968-3000, 3941-3965, 4024-4036 + fd2tra) --------------
template <bool WIN>
__device__ void p1_exe(const Acc<WIN>& a, const Win& w, int ix, int iy) {
  const int NXS = a.NXS;
  const double ev = D.ev, cutlo = D.cutlo;
  const int ixlb = (int)D.ixlb, ixrb = (int)D.ixrb;
  const P1Rng r = p1_ranges(w, ix, iy);
  if (!r.rfx) return;
  const int ix1 = IXP1(ix, iy);
  const double ni = a.get(PL_NI, ix, iy), ni_e = a.get(PL_NI, ix1, iy), te = a.get(PL_TE, ix, iy), te_e = a.get(PL_TE, ix1, iy);
  const double ne = f_ne(a, ix, iy), ne_e = f_ne(a, ix1, iy);
  const double gxf = GG(gxf, ix, iy), gx = GG(gx, ix, iy), gx_e = GG(gx, ix1, iy), sx = GG(sx, ix, iy), rrv = GG(rrv, ix, iy);
  const double loglambda = f_loglambda(a, ix, iy);
  double w1 = 0.;
  w1 = w1 + (D.zi * D.zi) * (ni * gx + ni_e * gx_e) / (gx + gx_e);
  const double ctaue = 3.5e11 * D.zi / loglambda;
  const double fxe = D.kxe * D.ce * ctaue / (D.me * ev * sqrt(ev));
  double fxet = fxe;
  if ((iy <= D.iysptrx) && ix > D.ixpt1 && ix <= D.ixpt2) fxet = fxe / (1. + (D.rkxecore - 1.) * d_powi(D.yyf[iy] / (D.yyf[0] + 4.e-50), D.inkxc));
  const double niavex = (ni * gx + ni_e * gx_e) / (gx + gx_e);
  double hcxe = 0. + fxet * niavex / w1;
  double ae;
  if (D.concap == 0) {
    double teave = (te * gx + te_e * gx_e) / (gx + gx_e);
    if (ix == ixlb) teave = a.get(PL_TE, ixlb + 1, iy);
    if (ix == ixrb) teave = a.get(PL_TE, ixrb, iy);
    ae = fmax(teave, D.temin * ev);
  } else ae = D.afix * ev;
  const double zeffave = (f_zeff(a, ix, iy) * gx + f_zeff(a, ix1, iy) * gx_e) / (gx + gx_e);
  const double zcoef = 0.308 + 0.767 * zeffave - 0.075 * (zeffave * zeffave);
  hcxe = hcxe * rrv * rrv * ae * ae * sqrt(ae) * zcoef;
  const double lmfpe = 2e16 * ((te / ev) * (te / ev)) / ne;
  const double neavex = (ne * gx + ne_e * gx_e) / (gx + gx_e);
  const double dte = te - te_e;
  const double ste = 0.5 * D.alfkxe * (te + te_e);
  hcxe = hcxe * (cutlo + dte * dte) / (cutlo + dte * dte + ste * ste) + 0. * neavex;
  hcxe = hcxe / ((1. + lmfpe / D.lmfplim) * (1 + hcxe * (gx * gx) * D.tdiflim / ne));
  const double t0 = fmax(te, D.temin * ev), t1 = fmax(te_e, D.temin * ev);
  const double vt0 = sqrt(t0 / D.me), vt1 = sqrt(t1 / D.me);
  double wallfac = 1.;
  if ((ix == ixlb || ix == ixrb) && (D.isplflxl == 0)) wallfac = D.flalfepl / D.flalfe;
  const double qfl = wallfac * D.flalfe * sx * rrv * (ne * vt0 * t0 + ne_e * vt1 * t1) / 2;
  const double csh = sx * hcxe * gxf;
  const double lxtec = 0.5 * (te + te_e) / (fabs(te - te_e) * gxf + 100. * cutlo);
  const double qsh = csh * (te - te_e) * (1. + lxtec / D.lxtemax);
  const double qr = (1 - D.isflxlde) * fabs(qsh / qfl);
  const double conxe = (1 - D.isflxlde) * csh / ((1 + qr) * (1 + qr)) + D.isflxlde * csh / ue_pow(1 + ue_pow(fabs(qsh / qfl), D.flgam), 1 / D.flgam);
  const double rr = GG(rr, ix, iy), rr_e = GG(rr, ix1, iy);
  double floxe = 0. + (d_sgn(qr * qr, qsh) / ((1 + qr) * (1 + qr))) * D.flalfea[ix] * sx * (ne * rr * vt0 + ne_e * rr_e * vt1) / 2;
  const double floxe0 = floxe;
  const double vex = f_upe_pre(a, ix, iy) * rrv + 0. - 0.;
  floxe = floxe + D.cfcvte * 1.25 * (ne + ne_e) * vex * sx - 0.;
  double feex;
  if ((int)(D.methe % 10) == 2) feex = floxe * (te_e + te) / 2. - conxe * (te_e - te);
  else feex = d_upwind(floxe, te, te_e) - conxe * (te_e - te);
  a.set(PL_FEEX, ix, iy, feex);
  if (!WIN && d_cut(ix, iy)) {  // the variant a windowed evaluation computes on the cut face: vex = 0
    const double vex0 = 0. * rrv + 0. - 0.;
    const double fl = floxe0 + D.cfcvte * 1.25 * (ne + ne_e) * vex0 * sx - 0.;
    double f0;
    if ((int)(D.methe % 10) == 2) f0 = fl * (te_e + te) / 2. - conxe * (te_e - te);
    else f0 = d_upwind(fl, te, te_e) - conxe * (te_e - te);
    a.set(PL_FEEXC, ix, iy, f0);
  }
}

// ---- x-face ion energy flux (oderhs.m:2875-2965, 3001-3009, 3967-3993, 4065-4071, 4234-4240 + fd2tra) ---------
template <bool WIN>
__device__ void p1_exi(const Acc<WIN>& a, const Win& w, int ix, int iy) {
  const int NXS = a.NXS;
  const double ev = D.ev, cutlo = D.cutlo;
  const int ixlb = (int)D.ixlb, ixrb = (int)D.ixrb;
  const P1Rng r = p1_ranges(w, ix, iy);
  if (!r.rfx) return;
  const int ix1 = IXP1(ix, iy);
  const double ni = a.get(PL_NI, ix, iy), ni_e = a.get(PL_NI, ix1, iy), ti = a.get(PL_TI, ix, iy), ti_e = a.get(PL_TI, ix1, iy);
  const double ng = a.get(PL_NG, ix, iy), ng_e = a.get(PL_NG, ix1, iy);
  const double ne = f_ne(a, ix, iy), ne_e = f_ne(a, ix1, iy);
  const double gxf = GG(gxf, ix, iy), gx = GG(gx, ix, iy), gx_e = GG(gx, ix1, iy), sx = GG(sx, ix, iy), rrv = GG(rrv, ix, iy);
  const double loglambda = f_loglambda(a, ix, iy);
  double w2 = 0.;
  w2 = w2 + ((D.zi * D.zi) * sqrt(2 * D.mi * D.mi / (D.mi + D.mi))) * (ni * gx + ni_e * gx_e) / (gx + gx_e);
  const double ctaui = 2.1e13 / (loglambda * (D.zi * D.zi));
  const double fxi = D.kxi * D.ci * ctaui / (ev * sqrt(ev * D.mp));
  double fxit = fxi;
  if ((iy <= D.iysptrx) && ix > D.ixpt1 && ix <= D.ixpt2) fxit = D.kxicore * fxi;
  double niavex = (ni * gx + ni_e * gx_e) / (gx + gx_e);
  double hcxij = fxit * niavex / w2;
  double aa, tiave = 0.;
  if (D.concap == 0) {
    tiave = (ti * gx + ti_e * gx_e) / (gx + gx_e);
    if (ix == ixlb) tiave = a.get(PL_TI, ixlb + 1, iy);
    if (ix == ixrb) tiave = a.get(PL_TI, ixrb, iy);
    aa = fmax(tiave, D.temin * ev);
  } else aa = D.afix * ev;
  hcxij = hcxij * rrv * rrv * aa * aa * sqrt(aa);
  const double lmfpi = 1.e16 * ((tiave / ev) * (tiave / ev)) / ni;
  niavex = (ni * gx + ni_e * gx_e) / (gx + gx_e);
  hcxij = hcxij / (1. + lmfpi / D.lmfplim);
  const double dti = ti - ti_e;
  const double sti = 0.5 * D.alfkxi * (ti + ti_e);
  hcxij = hcxij * (cutlo + dti * dti) / (cutlo + dti * dti + sti * sti) + 0. * niavex;
  if (D.isflxldi == 2) {
    niavex = (ni * gx + ni_e * gx_e) / (gx + gx_e);
    double wallfac = 1.;
    if ((ix == ixlb || ix == ixrb) && (D.isplflxl == 0)) wallfac = D.flalfipl / D.flalfi;
    const double qflx = wallfac * D.flalfi * rrv * sqrt(aa / D.mi) * niavex * aa;
    const double cshx = hcxij;
    const double lxtic = 0.5 * (ti + ti_e) / (fabs(ti - ti_e) * gxf + 100. * cutlo);
    const double qshx = cshx * (ti - ti_e) * gxf * (1. + lxtic / D.lxtimax);
    hcxij = cshx / (1 + fabs(qshx / qflx));
  }
  double hcxi = 0. + hcxij;
  hcxi = hcxi + D.cftiexclg * D.cfneut * D.cfneutsor_ei * D.kxn * (ng * ti + ng_e * ti_e) / (D.mi * (a.get(PL_NUCX, ix, iy) + a.get(PL_NUCX, ix1, iy)));
  double conxi, floxi = 0.;
  if (D.isflxldi != 2) {
    const double u0 = fmax(ti, D.temin * ev), u1 = fmax(ti_e, D.temin * ev);
    const double vt0 = sqrt(u0 / D.mi), vt1 = sqrt(u1 / D.mi);
    double wallfac = 1.;
    if ((ix == ixlb || ix == ixrb) && (D.isplflxl == 0)) wallfac = D.flalfipl / D.flalfi;
    const double qfl = wallfac * D.flalfia[ix] * sx * rrv * (ne * vt0 * u0 + ne_e * vt1 * u1) / 2;
    const double csh = sx * hcxi * gxf;
    const double lxtic = 0.5 * (ti + ti_e) / (fabs(ti - ti_e) * gxf + 100. * cutlo);
    const double qsh = csh * (ti - ti_e) * (1. + lxtic / D.lxtimax);
    const double qr = (1 - D.isflxldi) * fabs(qsh / qfl);
    conxi = (1 - D.isflxldi) * csh / ((1 + qr) * (1 + qr)) + D.isflxldi * csh / ue_pow(1 + ue_pow(fabs(qsh / qfl), D.flgam), 1 / D.flgam);
    const double rr = GG(rr, ix, iy), rr_e = GG(rr, ix1, iy);
    floxi = floxi + (d_sgn(qr * qr, qsh) / ((1 + qr) * (1 + qr))) * D.flalfia[ix] * sx * (ne * rr * vt0 + ne_e * rr_e * vt1) / 2;
  } else conxi = sx * hcxi * gxf;
  floxi = floxi + D.cfcvti * 2.5 * f_fnix(a, ix, iy);
  const double cg = D.cftiexclg * D.cfneut * D.cfneutsor_ei * D.cngtgx * D.cfcvti * 2.5;
  floxi = floxi + ((cg != 0.) ? cg * f_fngx(a, ix, iy) : 0.);  // cg*fngx with cg == 0 adds an exact 0
  double feix;
  if ((int)(D.methi % 10) == 2) feix = floxi * (ti_e + ti) / 2. - conxi * (ti_e - ti);
  else feix = d_upwind(floxi, ti, ti_e) - conxi * (ti_e - ti);
  a.set(PL_FEIX, ix, iy, feix);
}

// ---- y-face energy fluxes (oderhs.m:2871-2899, 3010-3013, 4001-4006, 4078-4128, 4244-4249 + fd2tra) ------------
template <bool WIN>
__device__ void p1_ey(const Acc<WIN>& a, const Win& w, int ix, int iy) {
  const int NXS = a.NXS;
  const int ny = (int)D.ny;
  const P1Rng r = p1_ranges(w, ix, iy);
  if (!r.rfy) return;
  const int iyp1 = min(ny + 1, iy + 1);
  const double te = a.get(PL_TE, ix, iy), ti = a.get(PL_TI, ix, iy);
  const double niy0 = a.get(PL_NIY0, ix, iy), niy1 = a.get(PL_NIY1, ix, iy);
  const double ney0 = 0. + D.zi * niy0, ney1 = 0. + D.zi * niy1;
  const double dynog = GG(dynog, ix, iy), sy = GG(sy, ix, iy);
  const double gy = GG(gy, ix, iy), gy_p = GG(gy, ix, iyp1);
  const double niavey = (niy0 * gy + niy1 * gy_p) / (gy + gy_p);
  const double diffusivwrk = D.fcdif * D.difni + 0.;
  double kyemix = D.fcdif * D.kye + 0.;
  if (D.kyet > 1.e-20 && iy > D.iysptrx) kyemix = (1. - D.ckyet) * kyemix + D.ckyet * D.kyet * diffusivwrk;
  const double hcye = 0. + (kyemix + 2.33 * (0. + 0.)) * D.zi * niavey;
  double kyimix = D.fcdif * D.kyi + 0.;
  if (D.kyit > 1.e-20 && iy > D.iysptrx) kyimix = (1. - D.ckyit) * kyimix + D.ckyit * D.kyit * diffusivwrk;
  const double hcyij = 0. + (kyimix + (0. + 0.)) * niavey;
  double hcyi = 0. + hcyij;
  const double cn = D.cftiexclg * D.cfneut * D.cfneutsor_ei * D.kyn;
  if (cn != 0.) {
    const double ngy0 = f_ilog(a, Fld<WIN>::ng, ix, iy, 0), ngy1 = f_ilog(a, Fld<WIN>::ng, ix, iy, 1);
    const double tiy0 = f_ilin(a, Fld<WIN>::ti, ix, iy, 0), tiy1 = f_ilin(a, Fld<WIN>::ti, ix, iy, 1);
    hcyi = hcyi + cn * (ngy0 * tiy0 + ngy1 * tiy1) / (D.mi * (a.get(PL_NUCX, ix, iy) + a.get(PL_NUCX, ix, iyp1)));
  } else hcyi = hcyi + 0.;
  const double conye = sy * hcye / dynog, conyi = sy * hcyi / dynog;
  const double vey = a.get(PL_VEY, ix, iy);
  const double floye = 0. + (D.cfloye / 2.) * (ney0 + ney1) * vey * sy + (0. + 0.) * 0.5 * sy * (ney0 + ney1);
  double floyi = 0. + D.cfloyi * f_fniy(a, ix, iy) + (0. + 0.) * 0.5 * sy * (niy0 + niy1);
  const double cg = D.cftiexclg * D.cfneut * D.cfneutsor_ei * D.cngtgy * 2.5;
  floyi = floyi + ((cg != 0.) ? cg * f_fngy(a, ix, iy) : 0.);
  const double te_n = a.get(PL_TE, ix, iy + 1), ti_n = a.get(PL_TI, ix, iy + 1);
  double feey, feiy;
  if ((int)(D.methe / 10) == 2) feey = floye * (te_n + te) / 2. - conye * (te_n - te);
  else feey = d_upwind(floye, te, te_n) - conye * (te_n - te);
  if ((int)(D.methi / 10) == 2) feiy = floyi * (ti_n + ti) / 2. - conyi * (ti_n - ti);
  else feiy = d_upwind(floyi, ti, ti_n) - conyi * (ti_n - ti);
  a.set(PL_FEEY, ix, iy, feey); a.set(PL_FEIY, ix, iy, feiy);
}

// ---- momentum fluxes (oderhs.m:3483-3581), evaluated by the consumer cell -----------------------------------------------
template <bool WIN> __device__ inline double f_fmix(const Acc<WIN>& a, int ixc, int ixw, int iy) {
  // fmix at cell centre ixc = ixp1(ixw): upwind(flox(ixc), up(ixw), up(ixc)) - conx(ixc)*(up(ixc)-up(ixw))   (fd2tra pos=1)
  const int NXS = a.NXS;
  const int ixm = IXM1(ixc, iy);
  const double uuv = 0.5 * (f_uu_cut(a, ixm, iy) + f_uu_cut(a, ixc, iy));
  const double vol = GG(vol, ixc, iy), gx = GG(gx, ixc, iy);
  const double flox = D.cmfx * f_nm(a, ixc, iy) * uuv * vol * gx;
  double conx;
  if (D.isgxvon == 0) conx = a.get(PL_VISX, ixc, iy) * vol * gx * gx;
  else conx = a.get(PL_VISX, ixc, iy) * vol * gx * 2 * GG(gxf, ixc, iy) * GG(gxf, ixm, iy) / (GG(gxf, ixc, iy) + GG(gxf, ixm, iy));
  const double p0 = a.get(PL_UP, ixw, iy), p1 = a.get(PL_UP, ixc, iy);
  if ((int)(D.methu % 10) == 2) return flox * (p1 + p0) / 2. - conx * (p1 - p0);
  return d_upwind(flox, p0, p1) - conx * (p1 - p0);
}
template <bool WIN> __device__ inline double f_fmiy(const Acc<WIN>& a, int ix, int iy) {
  const int NXS = a.NXS;
  const int ix2 = IXP1(ix, iy), ix4 = IXP1(ix, iy + 1);
  const double syv = GG(syv, ix, iy);
  const double vy = a.get(PL_VY, ix, iy);
  double floy;
  if (iy == D.iysptrx1 && (ix == D.ixpt1 || ix == D.ixpt2)) {
    floy = (D.cmfy / 2) * syv * (d_ave(f_nm(a, ix, iy), f_nm(a, ix, iy + 1))) * vy;
    floy = floy + (D.cmfy / 2) * syv * (d_ave(f_nm(a, ix, iy), f_nm(a, ix, iy + 1))) * 0.;
  } else {
    const double s = d_ave(f_nm(a, ix, iy), f_nm(a, ix, iy + 1)) + d_ave(f_nm(a, ix2, iy), f_nm(a, ix4, iy + 1));
    floy = (D.cmfy / 4) * syv * (s) * (vy + a.get(PL_VY, ix2, iy));
    floy = floy + (D.cmfy / 4) * syv * (s) * (0. + 0.);
  }
  double cony;
  const double v00 = f_visy(a, ix, iy) * GG(gy, ix, iy), v01 = f_visy(a, ix, iy + 1) * GG(gy, ix, iy + 1);
  const double v10 = f_visy(a, ix2, iy) * GG(gy, ix2, iy), v11 = f_visy(a, ix4, iy + 1) * GG(gy, ix4, iy + 1);
  if (D.ishavisy == 1) cony = .5 * syv * (d_ave(v00, v01) + d_ave(v10, v11));
  else cony = .25 * D.cfaccony * syv * (v00 + v01 + v10 + v11);
  const double p0 = a.get(PL_UP, ix, iy), p1 = a.get(PL_UP, ix, iy + 1);
  if ((int)(D.methu / 10) == 2) return floy * (p1 + p0) / 2. - cony * (p1 - p0);
  return d_upwind(floy, p0, p1) - cony * (p1 - p0);
}

// half-space problem: the velocity rows on the cut ix = ixpt2, iy <= iysptrx2 are overwritten by up -> 0 whenever the
// row window reaches the cut (boundary.m:1772-1785)
__device__ __forceinline__ bool cut_up_gate(const Win& w) {
  return D.isfixlb == 2 && w.i2 <= (int)D.ixpt2 && w.i5 >= (int)D.ixpt2 && w.j2 <= (int)D.iysptrx2;
}
// The less common wall and core boundary-condition options (wall densities isnwconi/o 1-3, wall temperatures
// istepfc/istipfc/istewc/istiwc 2-3, recycling walls matwalli/matwallo > 0, core options isupcore 2-3, iflcore -1,
// isngcore 1-4), kept out of line so that the common path of phase2_guard stays short: DT.rarebc (derived at init)
// says whether any of them is selected anywhere.  Overwrites the rows the common path has set for (ix, iy = 0 | ny+1).
template <bool WIN>
__device__ __noinline__ void guard_rare(const Acc<WIN>& a, int ix, int iy, double out[UE_NV]) {
  const int NXS = a.NXS;
  const int ny = (int)D.ny;
  const double ev = D.ev, pi = D.pi;
  const double ni = a.get(PL_NI, ix, iy), up = a.get(PL_UP, ix, iy), te = a.get(PL_TE, ix, iy), ti = a.get(PL_TI, ix, iy), ng = a.get(PL_NG, ix, iy);
  if (iy == 0) {
    const bool core = (D.isixcore[ix] == 1);
    const double sy = GG(sy, ix, 0);
    if (!core) {
      const int64_t mn = D.isnwconiix[ix];
      if (mn == 1) out[0] = D.nurlxn * (D.nwalli[ix] - ni) / D.n0;  // fixed wall density (boundary.m:267-270)
      else if (mn == 2) {  // extrapolation (boundary.m:271-277)
        const double n1 = a.get(PL_NI, ix, 1);
        double nbound = n1 - GG(gyf, ix, 1) * (a.get(PL_NI, ix, 2) - n1) / GG(gyf, ix, 0);
        nbound = 1.2 * nbound / (1 + 0.5 * ue_exp(-2 * (nbound / n1 - 1))) + 0.2 * n1;
        out[0] = D.nurlxn * (nbound - ni) / D.n0;
      } else if (mn == 3) {  // specified gradient length (boundary.m:278-282)
        const double gyf0 = GG(gyf, ix, 0);
        out[0] = -D.nurlxn * (a.get(PL_NIY0, ix, 0) - a.get(PL_NIY1, ix, 0) * (2 * gyf0 * D.lynipf[ix] - 1) / (2 * gyf0 * D.lynipf[ix] + 1) - D.nwimin) / D.n0;
      }
      // boundary.m:550-565, 597-612: 2 extrapolation from rows 1 and 2, 3 specified gradient length
      const int64_t me = D.istepfcix[ix], mi = D.istipfcix[ix];
      if (me == 2) {
        const double t1 = a.get(PL_TE, ix, 1);
        double tbound = t1 - GG(gyf, ix, 1) * (a.get(PL_TE, ix, 2) - t1) / GG(gyf, ix, 0);
        tbound = fmax(tbound, D.tbmin * ev);
        out[2] = D.nurlxe * (tbound - te) / (D.temp0 * ev);
      } else if (me == 3) {
        const double t1 = a.get(PL_TE, ix, 1);
        out[2] = D.nurlxe * ((t1 - te) - 0.5 * (t1 + te) / (GG(gyf, ix, 0) * D.lytepf[ix])) / (D.temp0 * ev);
      }
      if (mi == 2) {
        const double t1 = a.get(PL_TI, ix, 1);
        double tbound = t1 - GG(gyf, ix, 1) * (a.get(PL_TI, ix, 2) - t1) / GG(gyf, ix, 0);
        tbound = fmax(tbound, D.tbmin * ev);
        out[3] = D.nurlxi * (tbound - ti) / (D.temp0 * ev);
      } else if (mi == 3) {
        const double t1 = a.get(PL_TI, ix, 1);
        out[3] = D.nurlxi * ((t1 - ti) - 0.5 * (t1 + ti) / (GG(gyf, ix, 0) * D.lytipf[ix])) / (D.temp0 * ev);
      }
    } else {
      if (D.isupcore == 2) {  // d2(up)/dy2 = 0 (boundary.m:323-326)
        const double u1 = a.get(PL_UP, ix, 1);
        out[1] = D.nurlxu * ((u1 - up) * GG(gy, ix, 1) - (a.get(PL_UP, ix, 2) - u1) * GG(gy, ix, 2)) / (GG(gy, ix, 1) * D.vpnorm);
      } else if (D.isupcore == 3) out[1] = -D.nurlxu * f_fmiy(a, ix, 0) / (D.vpnorm * sy * D.fnorm);  // no radial momentum flux (boundary.m:327-329)
      if (D.iflcore == -1) {  // zero radial temperature gradient (boundary.m:546-548, 594-596)
        out[2] = -D.nurlxe * (te - a.get(PL_TE, ix, 1)) * D.n0 / D.ennorm;
        out[3] = -D.nurlxi * (ti - a.get(PL_TI, ix, 1)) * D.n0 / D.ennorm;
      }
    }
    // neutral density (boundary.m:651-681, 733-760)
    const double t0 = fmax(D.cdifg * f_tg(a, ix, 0), D.tgmin * ev);
    const double vyn = 0.25 * sqrt(8 * t0 / (pi * D.mg));
    const double ng1 = a.get(PL_NG, ix, 1);
    const double nharmave = 2. * (ng * ng1) / (ng + ng1);
    if (core) {
      if (D.isngcore == 1) out[4] = D.nurlxg * (D.ngcore - ng) / D.n0g;
      else if (D.isngcore == 2) {
        const double lengg = sqrt(f_tg(a, ix, 0) / (D.mg * (f_nuix(a, ix, 0) * a.get(PL_NUIZ, ix, 0))));
        out[4] = D.nurlxn * ((ng1 - ng) - 0.5 * (ng1 + ng) / (GG(gyf, ix, 0) * lengg)) / D.n0g;
      } else if (D.isngcore == 3) {
        double nbound = ng1 - GG(gyf, ix, 1) * (a.get(PL_NG, ix, 2) - ng1) / GG(gyf, ix, 0);
        nbound = 1.2 * nbound / (1 + 0.5 * ue_exp(-2 * (nbound / ng1 - 1))) + 0.2 * ng1;
        out[4] = D.nurlxn * (nbound - ng) / D.n0g;
      } else if (D.isngcore == 4) out[4] = D.nurlxn * (ng1 - ng) / D.n0g;
    } else if (D.matwalli[ix] > 0) {  // recycling wall
      const double fng_chem = 0., sputflxpf = 0.;
      const double fng_alb = (1 - D.albedoi[ix]) * nharmave * vyn * sy;
      const double rw = D.recycwit[ix];
      if (rw > 0.) {
        double fniy_recy = D.fac2sp * a.get(PL_FNIY, ix, 0);
        if (D.isrefluxclip == 1) fniy_recy = fmin(fniy_recy, 0.);
        out[4] = -D.nurlxg * (a.get(PL_FNGY, ix, 0) + fniy_recy * rw - D.fngyi_use[ix] - D.fngysi[ix] + fng_alb - fng_chem + sputflxpf) / (vyn * D.n0g * sy);
      } else if (rw < -1) out[4] = D.nurlxg * (D.ngbackg - ng) / D.n0g;
      else {
        const double nh2 = 2. * (ng * ng1) / (ng + ng1);
        out[4] = -D.nurlxg * (a.get(PL_FNGY, ix, 0) + (1 + rw) * nh2 * vyn * sy) / (vyn * D.n0g * sy);
      }
    }
  } else {  // outer wall (boundary.m:1188-1204, 1314-1357, 1424-1452)
    const double sy = GG(sy, ix, ny);
    const int64_t mn = D.isnwconoix[ix];
    if (mn == 1) out[0] = D.nurlxn * (D.nwallo[ix] - ni) / D.n0;
    else if (mn == 2) {
      const double n1 = a.get(PL_NI, ix, ny);
      double nbound = n1 + GG(gyf, ix, ny - 1) * (n1 - a.get(PL_NI, ix, ny - 1)) / GG(gyf, ix, ny);
      nbound = 1.2 * nbound / (1 + 0.5 * ue_exp(-2 * (nbound / n1 - 1))) + 0.2 * n1;
      out[0] = D.nurlxn * (nbound - ni) / D.n0;
    } else if (mn == 3) {
      const double gyfn = GG(gyf, ix, ny);
      out[0] = -D.nurlxn * (a.get(PL_NIY1, ix, ny) - a.get(PL_NIY0, ix, ny) * (2 * gyfn * D.lyniwc[ix] - 1) / (2 * gyfn * D.lyniwc[ix] + 1) - D.nwomin) / D.n0;
    }
    const int64_t me = D.istewcix[ix], mi = D.istiwcix[ix];
    if (me == 2) {
      const double t1 = a.get(PL_TE, ix, ny);
      double tbound = t1 + GG(gyf, ix, ny - 1) * (t1 - a.get(PL_TE, ix, ny - 1)) / GG(gyf, ix, ny);
      tbound = fmax(tbound, D.tbmin * ev);
      out[2] = D.nurlxe * (tbound - te) / (D.temp0 * ev);
    } else if (me == 3) {
      const double t1 = a.get(PL_TE, ix, ny);
      out[2] = D.nurlxe * ((t1 - te) - 0.5 * (t1 + te) / (GG(gyf, ix, ny) * D.lytewc[ix])) / (D.temp0 * ev);
    }
    if (mi == 2) {
      const double t1 = a.get(PL_TI, ix, ny);
      double tbound = t1 + GG(gyf, ix, ny - 1) * (t1 - a.get(PL_TI, ix, ny - 1)) / GG(gyf, ix, ny);
      tbound = fmax(tbound, D.tbmin * ev);
      out[3] = D.nurlxi * (tbound - ti) / (D.temp0 * ev);
    } else if (mi == 3) {
      const double t1 = a.get(PL_TI, ix, ny);
      out[3] = D.nurlxi * ((t1 - ti) - 0.5 * (t1 + ti) / (GG(gyf, ix, ny) * D.lytiwc[ix])) / (D.temp0 * ev);
    }
    if (D.matwallo[ix] > 0) {
      const double t0 = fmax(D.cdifg * f_tg(a, ix, ny + 1), D.tgmin * ev);
      const double vyn = 0.25 * sqrt(8 * t0 / (pi * D.mg));
      const double fng_chem = 0., sputflxw = 0.;
      const double ngc = a.get(PL_NG, ix, ny);
      const double nharmave = 2. * (ngc * ng) / (ngc + ng);
      const double fng_alb = (1 - D.albedoo[ix]) * nharmave * vyn * sy;
      const double rw = D.recycwot[ix];
      if (rw > 0.) {
        double fniy_recy = D.fac2sp * a.get(PL_FNIY, ix, ny);
        if (D.isrefluxclip == 1) fniy_recy = fmax(fniy_recy, 0.);
        out[4] = D.nurlxg * (a.get(PL_FNGY, ix, ny) + fniy_recy * rw + D.fngyso[ix] + D.fngyo_use[ix] - fng_alb + fng_chem + sputflxw) / (vyn * D.n0g * sy);
      } else if (rw < -1) out[4] = D.nurlxg * (D.ngbackg - ng) / D.n0g;
      else {
        const double nh2 = 2. * (ngc * ng) / (ngc + ng);
        out[4] = D.nurlxg * (a.get(PL_FNGY, ix, ny) - (1 + rw) * nh2 * vyn * sy) / (vyn * D.n0g * sy);
      }
    }
  }
}

// ============================================================================================
// phase 2b — guard-cell rows (bouncon, boundary.m:102-2800).  Returns a 5-bit mask of rows written.
// `ix,iy` is a guard cell; the right-plate momentum row that lives in interior column nx is
// produced by rightplate_up().
// ============================================================================================
template <bool WIN>
__device__ int phase2_guard(const Acc<WIN>& a, const Win& w, int ix, int iy, double out[UE_NV]) {
  const int NXS = a.NXS;
  const int nx = (int)D.nx, ny = (int)D.ny;
  const double ev = D.ev, pi = D.pi;
  const int ixlb = (int)D.ixlb, ixrb = (int)D.ixrb;
  int mask = 0;
  const bool gl = (w.xcnearlb || w.openbox), gr = (w.xcnearrb || w.openbox);
  if (iy == 0 || iy == ny + 1) {
    const bool bottom = (iy == 0);
    const bool sect = bottom ? (w.j3 <= 0) : (w.j7 >= ny + 1);
    if (!sect) return 0;
    const int iyc = bottom ? 1 : ny;       // adjacent interior row
    const int iyf = bottom ? 0 : ny;       // index of the y-face between guard and interior
    if (in_rng(ix, w.i4, w.i8)) {
      const double ni = a.get(PL_NI, ix, iy), up = a.get(PL_UP, ix, iy), te = a.get(PL_TE, ix, iy), ti = a.get(PL_TI, ix, iy), ng = a.get(PL_NG, ix, iy);
      const double sy = GG(sy, ix, iyf);
      if (bottom) {
        const bool core = (D.isixcore[ix] == 1);
        // density (boundary.m:207-265)
        if (core) {
          if (D.isnicore == 1) out[0] = D.nurlxn * (D.ncore - ni) / D.n0;
          else out[0] = -D.nurlxn * (D.qe * (a.get(PL_FNIY, ix, 0) - 0.) / sy - D.curcore * GG(gyf, ix, 0) / D.sygytotc) / (D.qe * D.vpnorm * D.n0);
        } else {  // isnwconi = 0; the other options: guard_rare
          out[0] = D.nurlxn * ((1 - D.ifluxni) * (a.get(PL_NIY1, ix, 0) - a.get(PL_NIY0, ix, 0)) -
                               D.ifluxni * (a.get(PL_FNIY, ix, 0) / (sy * D.vpnorm) - 0.001 * a.get(PL_NI, ix, 1) * a.get(PL_VY, ix, 0) / D.vpnorm)) / D.n0;
        }
        // parallel velocity (boundary.m:313-380)
        if (core) {
          if (D.isupcore == 0) out[1] = D.nurlxu * (D.upcore - up) / D.vpnorm;
          else out[1] = D.nurlxu * (a.get(PL_UP, ix, 1) - up) / D.vpnorm;  // isupcore = 1; 2, 3: guard_rare
        } else if (D.isupwiix[ix] == 2) out[1] = D.nurlxu * f_nm(a, ix, 0) / D.fnorm * (a.get(PL_UP, ix, 1) - up);
        else out[1] = D.nurlxu * f_nm(a, ix, 0) / D.fnorm * (0. - up);
        // temperatures (boundary.m:524-628)
        if (core) {
          const double ne = f_ne(a, ix, 0);
          out[2] = D.nurlxe * (D.tcoree * ev - te) * 1.5 * ne / D.ennorm;
          out[3] = D.nurlxi * (D.tcorei * ev - ti) * 1.5 * ne / D.ennorm;
          if (D.iflcore == 1) {
            const int ixe = IXP1(ix, 0);
            out[2] = -D.nurlxe * (te - a.get(PL_TE, ixe, 0)) * D.n0 / D.ennorm;
            out[3] = -D.nurlxi * (ti - a.get(PL_TI, ixe, 0)) * D.n0 / D.ennorm;
            const int ix_fl_bc = min((int)D.ixpt2, nx);
            if (ix == ix_fl_bc) {  // integrated core power: serial sum in the reference's order
              int ii = max(0, (int)D.ixpt1 + 1);
              double feeytotc = a.get(PL_FEEY, ii, 0) - 0., feiytotc = a.get(PL_FEIY, ii, 0) - 0.;
              do { ii = IXP1(ii, 0); feeytotc = feeytotc + a.get(PL_FEEY, ii, 0) - 0.; feiytotc = feiytotc + a.get(PL_FEIY, ii, 0) - 0.; } while (ii != ix_fl_bc);
              out[2] = -D.nurlxe * (feeytotc - D.pcoree) / (D.vpnorm * D.ennorm);
              out[3] = -D.nurlxi * (feiytotc - D.pcorei) / (D.vpnorm * D.ennorm);
            }
          }  // iflcore = -1: guard_rare
        } else {  // boundary.m:550-565, 597-612: 0 zero flux, 1 fixed; 2 extrapolation, 3 gradient length: guard_rare
          if (D.istepfcix[ix] == 0) out[2] = -D.nurlxe * (a.get(PL_FEEY, ix, 0) / (D.n0 * D.vpnorm * sy)) / (D.temp0 * ev);
          else out[2] = D.nurlxe * (D.tewalli[ix] * ev - te) / (D.temp0 * ev);
          if (D.istipfcix[ix] == 0) out[3] = -D.nurlxi * (a.get(PL_FEIY, ix, 0) / (D.n0 * D.vpnorm * sy)) / (D.temp0 * ev);
          else out[3] = D.nurlxi * (D.tiwalli[ix] * ev - ti) / (D.temp0 * ev);
        }
        // neutral density (boundary.m:632-767)
        {
          const double t0 = fmax(D.cdifg * f_tg(a, ix, 0), D.tgmin * ev);
          const double vyn = 0.25 * sqrt(8 * t0 / (pi * D.mg));
          const double ng1 = a.get(PL_NG, ix, 1);
          const double nharmave = 2. * (ng * ng1) / (ng + ng1);
          if (core) {  // isngcore = 0 (boundary.m:651-657); 1..4: guard_rare
            const double fng_alb = (1 - D.albedoc) * nharmave * vyn * sy;
            out[4] = -D.nurlxg * (a.get(PL_FNGY, ix, 0) + fng_alb) / (vyn * sy * D.n0g);
          } else {  // recycling walls (matwalli > 0): guard_rare
            const double fng_chem = 0., sputflxpf = 0.;
            const double fng_alb = (1 - D.albedoi[ix]) * nharmave * vyn * sy;
            out[4] = -D.nurlxg * (a.get(PL_FNGY, ix, 0) + fng_alb - fng_chem + sputflxpf) / (vyn * sy * D.n0g);
          }
        }
      } else {  // outer wall (boundary.m:1133-1462)
        out[0] = D.nurlxn * ((1 - D.ifluxni) * (a.get(PL_NIY0, ix, ny) - a.get(PL_NIY1, ix, ny)) +
                             D.ifluxni * (a.get(PL_FNIY, ix, ny) / (sy * D.vpnorm) - 0.001 * a.get(PL_NI, ix, ny) * a.get(PL_VY, ix, ny) / D.vpnorm)) / D.n0;  // isnwcono = 0
        if (D.isupwoix[ix] == 2) out[1] = D.nurlxu * f_nm(a, ix, ny) / D.fnorm * (a.get(PL_UP, ix, ny) - up);
        else out[1] = D.nurlxu * f_nm(a, ix, ny) / D.fnorm * (0. - up);
        if (D.istewcix[ix] == 0) out[2] = D.nurlxe * (a.get(PL_FEEY, ix, ny) / (D.n0 * D.vpnorm * sy)) / (D.temp0 * ev);
        else out[2] = D.nurlxe * (D.tewallo[ix] * ev - te) / (D.temp0 * ev);
        if (D.istiwcix[ix] == 0) out[3] = D.nurlxi * (a.get(PL_FEIY, ix, ny) / (D.n0 * D.vpnorm * sy)) / (D.temp0 * ev);
        else out[3] = D.nurlxi * (D.tiwallo[ix] * ev - ti) / (D.temp0 * ev);
        const double t0 = fmax(D.cdifg * f_tg(a, ix, ny + 1), D.tgmin * ev);
        const double vyn = 0.25 * sqrt(8 * t0 / (pi * D.mg));
        const double fng_chem = 0., sputflxw = 0.;
        const double ngc = a.get(PL_NG, ix, ny);
        const double nharmave = 2. * (ngc * ng) / (ngc + ng);
        const double fng_alb = (1 - D.albedoo[ix]) * nharmave * vyn * sy;
        out[4] = D.nurlxg * (a.get(PL_FNGY, ix, ny) - fng_alb + fng_chem + sputflxw) / (vyn * sy * D.n0g);
      }
      if (DT.rarebc) guard_rare<WIN>(a, ix, iy, out);  // the less common wall / core options overwrite the rows above
      mask = 0x1f;
    }
    if (bottom && ix == (int)D.ixpt2 && cut_up_gate(w)) { out[1] = D.nurlxu * (0. - a.get(PL_UP, ix, 0)) / D.vpnorm; mask |= 2; }  // boundary.m:1772-1785 (iy = 0)
    // corner cells and the special rows next to them (boundary.m:290-303, 897-983, 1209-1226, 1543-1630)
    if (ix == ixlb) {
      if (!(bottom && D.isfixlb == 2)) {  // boundary.m:291: the bottom-left corner keeps the iy=0 condition on a symmetry plane
        out[0] = D.nurlxn * (d_ave(a.get(PL_NI, ixlb, iyc), a.get(PL_NI, ixlb + 1, iy)) - a.get(PL_NI, ixlb, iy)) / D.n0; mask |= 1;
      }
      if (gl) {
        out[1] = -D.nurlxu * (a.get(PL_UP, ixlb, iy) - 0.5 * (a.get(PL_UP, ixlb, iyc) + a.get(PL_UP, ixlb + 1, iy))) / D.vpnorm;
        if (bottom) {
          out[2] = D.nurlxe * (0.5 * (a.get(PL_TE, ixlb + 1, 0) + a.get(PL_TE, ixlb, 1)) - a.get(PL_TE, ixlb, 0)) / (D.temp0 * ev);
          out[3] = D.nurlxi * (0.5 * (a.get(PL_TI, ixlb + 1, 0) + a.get(PL_TI, ixlb, 1)) - a.get(PL_TI, ixlb, 0)) / (D.temp0 * ev);
        } else {
          out[2] = D.nurlxe * (0.5 * (a.get(PL_TE, ixlb + 1, ny + 1) + a.get(PL_TE, ixlb, ny)) - a.get(PL_TE, ixlb, ny + 1)) / (D.temp0 * ev);
          out[3] = D.nurlxi * (0.5 * (a.get(PL_TI, ixlb + 1, ny + 1) + a.get(PL_TI, ixlb, ny)) - a.get(PL_TI, ixlb, ny + 1)) / (D.temp0 * ev);
        }
        out[4] = D.nurlxg * (a.get(PL_NG, ixlb + 1, iy) - a.get(PL_NG, ixlb, iy)) / D.n0g;
        mask |= 0x1e;
      }
    }
    if (ix == ixrb + 1) {
      out[0] = D.nurlxn * (d_ave(a.get(PL_NI, ixrb + 1, iyc), a.get(PL_NI, ixrb, iy)) - a.get(PL_NI, ixrb + 1, iy)) / D.n0; mask |= 1;
      if (gr) {
        out[1] = -D.nurlxu * (a.get(PL_UP, ixrb + 1, iy) - a.get(PL_UP, ixrb, iy)) / D.vpnorm;
        if (bottom) {
          out[2] = D.nurlxe * (0.5 * (a.get(PL_TE, ixrb + 1, 1) + a.get(PL_TE, ixrb, 0)) - a.get(PL_TE, ixrb + 1, 0)) / (D.temp0 * ev);
          out[3] = D.nurlxi * (0.5 * (a.get(PL_TI, ixrb + 1, 1) + a.get(PL_TI, ixrb, 0)) - a.get(PL_TI, ixrb + 1, 0)) / (D.temp0 * ev);
        } else {
          out[2] = D.nurlxe * (0.5 * (a.get(PL_TE, ixrb, ny + 1) + a.get(PL_TE, ixrb + 1, ny)) - a.get(PL_TE, ixrb + 1, ny + 1)) / (D.temp0 * ev);
          out[3] = D.nurlxi * (0.5 * (a.get(PL_TI, ixrb, ny + 1) + a.get(PL_TI, ixrb + 1, ny)) - a.get(PL_TI, ixrb + 1, ny + 1)) / (D.temp0 * ev);
        }
        out[4] = D.nurlxg * (a.get(PL_NG, ixrb, iy) - a.get(PL_NG, ixrb + 1, iy)) / D.n0g;
        mask |= 0x1e;
      }
    }
    if (ix == ixrb && gr) {  // boundary.m:943-946, 1589-1592
      out[1] = -D.nurlxu * (a.get(PL_UP, ixrb, iy) - 0.5 * (a.get(PL_UP, ixrb - 1, iy) + a.get(PL_UP, ixrb, iyc))) / D.vpnorm;
      mask |= 2;
    }
    return mask;
  }
  // ---- plates --------------------------------------------------------------------------------------------
  if (!in_rng(iy, w.j2, w.j5)) return 0;
  if (ix == 0 && D.isfixlb == 2) {  // ix = 0 is a symmetry plane (boundary.m:1666-1770; rlimiter lies beyond the mesh)
    if (w.i3 <= 0) {
      out[0] = D.nurlxn * (1 / D.n0) * (a.get(PL_NI, 1, iy) - a.get(PL_NI, 0, iy));
      out[1] = D.nurlxu * (0. - a.get(PL_UP, 0, iy)) / D.vpnorm;
      out[2] = D.nurlxe * f_ne(a, 0, iy) * (a.get(PL_TE, 1, iy) - a.get(PL_TE, 0, iy)) / D.ennorm;
      out[3] = D.nurlxi * f_ne(a, 0, iy) * (a.get(PL_TI, 1, iy) - a.get(PL_TI, 0, iy)) / D.ennorm;
      out[4] = D.nurlxg * (a.get(PL_NG, 1, iy) - a.get(PL_NG, 0, iy)) / D.n0g;
      mask = 0x1f;
    }
    return mask;
  }
  if (ix == ixlb && gl) {  // boundary.m:1793-2259
    const int ixt = ixlb, ixt1 = IXP1(ixt, iy);
    if (w.i3 <= ixlb + D.isextrnp) { out[0] = D.nurlxn * (a.get(PL_NI, ixt1, iy) - a.get(PL_NI, ixt, iy)) / D.n0; mask |= 1; }
    if (w.i3 <= ixlb) {
      const double te = a.get(PL_TE, ixt, iy), ti = a.get(PL_TI, ixt, iy), up = a.get(PL_UP, ixt, iy), up1 = a.get(PL_UP, ixt1, iy);
      const double sx = GG(sx, ixt, iy);
      const double ueb = D.cfueb * (0. - 0.) / GG(rrv, ixt, iy);
      const double cs = D.csfaclb * sqrt((te + D.csfacti * ti) / D.mi);
      out[1] = D.nurlxu * (-cs - ueb - up) / D.vpnorm;
      if (D.isupss == 1 && up1 + ueb < -cs) out[1] = D.nurlxu * (up1 - up) / D.vpnorm;
      if (D.isupss == -1) out[1] = D.nurlxu * (up1 - up) / D.vpnorm;
      double kfeix = 0.;
      kfeix = kfeix - D.cfvcsx * 0.5 * sx * a.get(PL_VISX, ixt1, iy) * GG(gx, ixt1, iy) * (up1 * up1 - up * up);
      const double kappal = 3.;
      const double bcel = (1 - D.newbcl * 0) * D.bcee + D.newbcl * 0 * (2. + kappal);
      const double bcil = (1 - D.newbcl * 0) * D.bcei + D.newbcl * 0 * (2.5);
      double t0 = te / ev;
      double f_cgpld = .5 * (1. - ue_cos(pi * (t0 - D.temin) / (.3 - D.temin)));
      if (t0 < D.temin) f_cgpld = 0.;
      if (t0 > 0.3) f_cgpld = 1.;
      t0 = fmax(f_tg(a, ixt1, iy), D.tgmin * ev);
      const double vxn = f_cgpld * 0.25 * sqrt(8 * t0 / (pi * D.mg));
      const double fnix = a.get(PL_FNIX, ixt, iy);
      {
        const double totfeexl = a.get(PL_FEEX, ixt, iy) + 0.;
        const double vex = a.get(PL_UPE, ixt, iy) * GG(rrv, ixt, iy) + 0. - 0.;
        const double totfnex = f_ne(a, ixt, iy) * vex * sx;
        out[2] = -D.nurlxe * (totfeexl - totfnex * te * bcel + D.cgpld * sx * 0.5 * a.get(PL_NG, ixt1, iy) * vxn * D.ediss * ev - D.cmneut * fnix * D.recycp * D.eedisspl * ev) /
                 (sx * D.vpnorm * D.ennorm);
      }
      {
        double totfeixl = a.get(PL_FEIX, ixt, iy) + D.ckinfl * kfeix;
        double totfnix = 0.;
        totfeixl = totfeixl + 0.;
        totfnix = totfnix + fnix;
        out[3] = -D.nurlxi * (totfeixl - totfnix * bcil * ti + D.cftiexclg * (-D.cmneut * fnix * D.recycp * D.cmntgpl * (ti - D.eidisspl * ev))) / (D.vpnorm * D.ennorm * sx);
      }
      {
        const double recy = D.recylb[iy];
        if (recy > 0.) {
          const double flux_inc = D.fac2sp * fnix;
          const double t0g = fmax(f_tg(a, ixt1, iy), D.tgmin * ev);
          const double vxg = 0.25 * sqrt(8 * t0g / (pi * D.mg));
          const double areapl = D.isoldalbarea * sx + (1 - D.isoldalbarea) * GG(sxnp, ixt, iy);
          out[4] = -D.nurlxg * (a.get(PL_FNGX, ixt, iy) - D.fngxlb_use[iy] - D.fngxslb[iy] + recy * flux_inc + (1 - D.alblb[iy]) * a.get(PL_NG, ixt1, iy) * vxg * areapl) /
                   (D.vpnorm * D.n0g * sx);
        } else {
          const double t0g = fmax(f_tg(a, ixt, iy), D.tgmin * ev);
          const double vxg = 0.25 * sqrt(8 * t0g / (pi * D.mg));
          out[4] = -D.nurlxg * (a.get(PL_FNGX, ixt, iy) + (1 + recy) * a.get(PL_NG, ixt, iy) * vxg * sx) / (vxg * sx * D.n0g);
        }
      }
      mask |= 0x1e;
    }
    return mask;
  }
  if (ix == ixrb + 1 && gr) {  // boundary.m:2457-2800
    const int ixt = ixrb + 1, ixt1 = IXM1(ixt, iy), ixt2 = IXM1(ixt1, iy);
    if (w.i6 >= (ixrb + 1 - D.isextrnp)) { out[0] = D.nurlxn * (a.get(PL_NI, ixt1, iy) - a.get(PL_NI, ixt, iy)) / D.n0; mask |= 1; }
    if (w.i6 >= ixrb + 1) {
      const double te = a.get(PL_TE, ixt, iy), ti = a.get(PL_TI, ixt, iy);
      const double up1 = a.get(PL_UP, ixt1, iy), up2 = a.get(PL_UP, ixt2, iy);
      const double sx1 = GG(sx, ixt1, iy);
      out[1] = D.nurlxu * (up1 - a.get(PL_UP, ixt, iy)) / D.vpnorm;  // boundary.m:2588
      double kfeix = 0.;
      kfeix = kfeix - D.cfvcsx * 0.5 * sx1 * a.get(PL_VISX, ixt1, iy) * GG(gx, ixt1, iy) * (up1 * up1 - up2 * up2);
      const double kappar = 3.;
      const double bcer = (1 - D.newbcr * 0) * D.bcee + D.newbcr * 0 * (2. + kappar);
      const double bcir = (1 - D.newbcr * 0) * D.bcei + D.newbcr * 0 * (2.5);
      double t0 = te / ev;
      double f_cgpld = .5 * (1. - ue_cos(pi * (t0 - D.temin) / (.3 - D.temin)));
      if (t0 < D.temin) f_cgpld = 0.;
      if (t0 > 0.3) f_cgpld = 1.;
      t0 = fmax(f_tg(a, ixt1, iy), D.tgmin * ev);
      const double vxn = f_cgpld * 0.25 * sqrt(8 * t0 / (pi * D.mg));
      const double fnix = a.get(PL_FNIX, ixt1, iy);
      {
        const double totfeexr = a.get(PL_FEEX, ixt1, iy) + 0.;
        const double vex = a.get(PL_UPE, ixt1, iy) * GG(rrv, ixt1, iy) + 0. - 0.;
        const double totfnex = f_ne(a, ixt, iy) * vex * sx1;
        out[2] = D.nurlxe * (totfeexr - totfnex * te * bcer - D.cgpld * sx1 * 0.5 * a.get(PL_NG, ixt1, iy) * vxn * D.ediss * ev - D.cmneut * fnix * D.recycp * D.eedisspl * ev) /
                 (sx1 * D.vpnorm * D.ennorm);
      }
      {
        double totfeixr = a.get(PL_FEIX, ixt1, iy) + D.ckinfl * kfeix;
        double totfnix = 0.;
        totfeixr = totfeixr + 0.;
        totfnix = totfnix + fnix;
        out[3] = D.nurlxi * (totfeixr - totfnix * bcir * ti + D.cftiexclg * (-D.cmneut * fnix * D.recycp * D.cmntgpl * (ti - D.eidisspl * ev))) / (D.vpnorm * D.ennorm * sx1);
      }
      {
        const double recy = D.recyrb[iy];
        if (recy > 0.) {
          const double flux_inc = D.fac2sp * fnix;
          const double t0g = fmax(f_tg(a, ixt1, iy), D.tgmin * ev);
          const double vxg = 0.25 * sqrt(8 * t0g / (pi * D.mg));
          const double areapl = D.isoldalbarea * sx1 + (1 - D.isoldalbarea) * GG(sxnp, ixt1, iy);
          out[4] = D.nurlxg * (a.get(PL_FNGX, ixt1, iy) + D.fngxrb_use[iy] - D.fngxsrb[iy] + recy * flux_inc - (1 - D.albrb[iy]) * a.get(PL_NG, ixt1, iy) * vxg * areapl) /
                   (D.vpnorm * D.n0g * sx1);
        } else {
          const double t0g = fmax(f_tg(a, ixt, iy), D.tgmin * ev);
          const double vxg = 0.25 * sqrt(8 * t0g / (pi * D.mg));
          out[4] = D.nurlxg * (a.get(PL_FNGX, ixt1, iy) - (1 + recy) * a.get(PL_NG, ixt, iy) * vxg * sx1) / (vxg * sx1 * D.n0g);
        }
      }
      mask |= 0x1e;
    }
    return mask;
  }
  return 0;
}

// right-plate Bohm condition lives in the momentum row of interior column ixrb (boundary.m:2534-2589)
template <bool WIN>
__device__ bool rightplate_up(const Acc<WIN>& a, const Win& w, int ix, int iy, double& val) {
  const int NXS = a.NXS;
  const int ixrb = (int)D.ixrb;
  if (ix != ixrb || !(w.xcnearrb || w.openbox) || !(w.i6 >= ixrb + 1) || !in_rng(iy, w.j2, w.j5)) return false;
  const int ixt = ixrb + 1, ixt1 = IXM1(ixt, iy), ixt2 = IXM1(ixt1, iy);
  const double ueb = D.cfueb * (0. - 0.) / GG(rrv, ixt1, iy);
  const double cs = D.csfacrb * sqrt((a.get(PL_TE, ixt, iy) + D.csfacti * a.get(PL_TI, ixt, iy)) / D.mi);
  const double up1 = a.get(PL_UP, ixt1, iy), up2 = a.get(PL_UP, ixt2, iy);
  val = D.nurlxu * (cs - ueb - up1) / D.vpnorm;
  if (D.isupss == 1 && up2 + ueb > cs) val = D.nurlxu * (up2 - up1) / D.vpnorm;
  if (D.isupss == -1) val = D.nurlxu * (up2 - up1) / D.vpnorm;
  return true;
}

// ============================================================================================
// phase 2a — interior cell rows (before rscalf), one ROLE per equation group so that four warps can
// work on the same cell:  p2_n (ni, ng rows + resco)   p2_m (up row)   p2_e (te row)   p2_i (ti row)
// ============================================================================================
template <bool WIN> __device__ inline double f_eqp(const Acc<WIN>& a, int ix, int iy) {  // oderhs.m:3091-3101
  const int NXS = a.NXS;
  const double ev = D.ev;
  const int ix1 = IXM1(ix, iy);
  const double ni = a.get(PL_NI, ix, iy), te = a.get(PL_TE, ix, iy), ti = a.get(PL_TI, ix, iy);
  const double w3 = 0.0 + ((D.zi * D.zi) / D.mi) * ni;
  const double aa = fmax(te, D.temin * ev);
  const double loglmcc = 0.5 * (f_loglambda(a, ix, iy) + f_loglambda(a, ix1, iy));
  const double coef1 = D.feqp * 4.8e-15 * loglmcc * sqrt(ev) * ev * D.mp;
  double eqp = coef1 * w3 * f_ne(a, ix, iy) / (aa * sqrt(aa));
  const double d = aa - ti, s = D.alfeqp * (aa + ti);
  eqp = eqp * (d * d) / (D.cutlo + d * d + s * s);
  return eqp;
}

// particle balance of an interior cell (oderhs.m:3407-3456); also evaluated for the east neighbour by the fused
// phase-2/3 kernel of the full residual (rscalf reads resco(ixp1), oderhs.m:8140-8160)
template <bool WIN>
__device__ __forceinline__ double f_resco(const Acc<WIN>& a, int ix, int iy, double psor, double psorxr) {
  const int NXS = a.NXS;
  const int ix1 = IXM1(ix, iy);
  const double ni = a.get(PL_NI, ix, iy);
  double resco = 0. + 0. * ni + 0. + D.cfneut * D.cfneutsor_ni * D.cnsor * psor + D.cfneut * D.cfneutsor_ni * D.cnsor * psorxr +
                 D.cfneut * D.cfneutsor_ni * D.cnsor * 0. - 0. + 0.;
  resco = resco - ((a.get(PL_FNIX, ix, iy) - a.get(PL_FNIX, ix1, iy)) + D.fluxfacy * (a.get(PL_FNIY, ix, iy) - a.get(PL_FNIY, ix, iy - 1)));
  return resco;
}
template <bool WIN>
__device__ void p2_n(const Acc<WIN>& a, int ix, int iy, double out[UE_NV], const int64_t* __restrict__ iseqalg) {
  const int NXS = a.NXS;
  const int ix1 = IXM1(ix, iy);
  const double vol = GG(vol, ix, iy);
  double psor, psorxr, psordis;
  f_psor(a, ix, iy, psor, psorxr, psordis);
  const double resco = f_resco(a, ix, iy, psor, psorxr);
  a.set(PL_RESCO, ix, iy, resco);
  // neutral balance (oderhs.m:6587-6595)
  const double psorg = -psor, psorrg = -psorxr;
  double resng = D.cngsor * (psorg + 0. + psorrg) + 0. + 0. * vol;
  resng = resng - D.cfneutdiv * D.cfneutdiv_fng * ((a.get(PL_FNGX, ix, iy) - a.get(PL_FNGX, ix1, iy)) + D.fluxfacy * (a.get(PL_FNGY, ix, iy) - a.get(PL_FNGY, ix, iy - 1)));
  const int64_t c = (int64_t)(ix + NXS * iy) * NVX;
  out[0] = (1 - iseqalg[c + 0]) * resco / (vol * D.n0);
  out[4] = D.isngon == 1 ? (1 - iseqalg[c + 4]) * resng / (vol * D.n0g) : 0.;
}

template <bool WIN>
__device__ void p2_m(const Acc<WIN>& a, const Win& w, int ix, int iy, double out[UE_NV], const int64_t* __restrict__ iseqalg) {
  const int NXS = a.NXS;
  const int ix1 = IXM1(ix, iy), ix2 = IXP1(ix, iy);
  const double gxf = GG(gxf, ix, iy), rrv = GG(rrv, ix, iy), sx = GG(sx, ix, iy);
  const double ng = a.get(PL_NG, ix, iy), up = a.get(PL_UP, ix, iy);
  // momentum source (oderhs.m:2496-2498, 2521-2525)
  double smoc = ((-D.cpgx * a.get(PL_GPEX, ix, iy) - 0.) * rrv + 0.) * sx / gxf;
  const double t0 = -D.cpiup * (a.get(PL_GPIX, ix, iy) * rrv - 0.) * sx / gxf;
  smoc = smoc + D.cpgx * t0;
  // oderhs.m:3746-3866
  const double ng_e = a.get(PL_NG, ix2, iy);
  const double dp1 = D.cngmom * (1 / D.fac2sp) * (ng_e * f_tg(a, ix2, iy) - ng * f_tg(a, ix, iy));
  double resmo = smoc + 0. * up - D.cfneut * D.cfneutsor_mi * sx * rrv * dp1 -
                 D.cfneut * D.cfneutsor_mi * D.cmwall * 0.5 * (ng + ng_e) * D.mi * up * 0.5 * (a.get(PL_NUCX, ix, iy) + a.get(PL_NUCX, ix2, iy)) * GG(volv, ix, iy) +
                 0. + D.cfmsor * (0. + 0.) + 0. + 0. + 0.;
  resmo = resmo - (f_fmix(a, ix2, ix, iy) - f_fmix(a, ix, ix1, iy) + D.fluxfacy * (f_fmiy(a, ix, iy) - f_fmiy(a, ix, iy - 1)));
  const int64_t c = (int64_t)(ix + NXS * iy) * NVX;
  out[1] = (1 - iseqalg[c + 1]) * resmo / (GG(volv, ix, iy) * D.fnorm);
  if (ix == D.ixrb) out[1] = resmo / (GG(volv, ix, iy) * D.fnorm);
  double v;
  if (rightplate_up<WIN>(a, w, ix, iy, v)) out[1] = v;
  if (cut_up_gate(w) && ix == (int)D.ixpt2 && iy <= (int)D.iysptrx2) out[1] = D.nurlxu * (0. - up) / D.vpnorm;  // boundary.m:1772-1785
}

template <bool WIN>
__device__ void p2_e(const Acc<WIN>& a, int ix, int iy, double out[UE_NV], const int64_t* __restrict__ iseqalg) {
  const int NXS = a.NXS;
  const int ny = (int)D.ny;
  const double ev = D.ev;
  const int ix1 = IXM1(ix, iy), ix2 = IXP1(ix, iy);
  const double vol = GG(vol, ix, iy), gx = GG(gx, ix, iy), gxf = GG(gxf, ix, iy), gxf_w = GG(gxf, ix1, iy);
  const double rrv = GG(rrv, ix, iy), rrv_w = GG(rrv, ix1, iy);
  const double ni = a.get(PL_NI, ix, iy), te = a.get(PL_TE, ix, iy), ti = a.get(PL_TI, ix, iy);
  const double up = a.get(PL_UP, ix, iy), up_w = a.get(PL_UP, ix1, iy);
  const double gpex = a.get(PL_GPEX, ix, iy), gpex_w = a.get(PL_GPEX, ix1, iy);
  const double upe_raw = f_upe_pre(a, ix, iy), upe_raw_w = f_upe_pre(a, ix1, iy);  // vex was formed before the cut zeroing
  const double upe = f_upe_cut(a, ix, iy), upe_w = f_upe_cut(a, ix1, iy);
  const double upi = f_upi_cut(a, ix, iy), upi_w = f_upi_cut(a, ix1, iy);
  const double vey = a.get(PL_VEY, ix, iy);
  double seec = 0.;
  {  // oderhs.m:2471-2495, 2534-2538, 2557-2572
    const double gx_e = GG(gx, ix2, iy), gx_w = GG(gx, ix1, iy);
    const double t1old = .5 * D.cvgp * (upe * rrv * d_ave(gx, gx_e) * gpex / gxf + upe_w * rrv_w * d_ave(gx, gx_w) * gpex_w / gxf_w);
    const double t2old = 0.;
    const int iyp1 = min(iy + 1, ny + 1), iym1 = max(iy - 1, 0);
    const double vex = upe_raw * rrv + 0. - 0., vex_w = upe_raw_w * rrv_w + 0. - 0.;
    const double t1new = .5 * D.cvgp * (vex * d_ave(gx, gx_e) * gpex / gxf + vex_w * d_ave(gx, gx_w) * gpex_w / gxf_w);
    const double gy = GG(gy, ix, iy);
    const double t2new = .5 * D.cvgp * (vey * d_ave(gy, GG(gy, ix, iyp1)) * a.get(PL_GPEY, ix, iy) / GG(gyf, ix, iy) +
                                       vey * d_ave(gy, GG(gy, ix, iym1)) * a.get(PL_GPEY, ix, iym1) / GG(gyf, ix, iym1));
    seec = seec + (t1old * vol - t2old) * D.oldseec + ((t1new + t2new) * vol) * (1 - D.oldseec);
    const double tv = 0.25 * (a.get(PL_FRICE, ix, iy) + a.get(PL_FRICE, ix1, iy)) * (upe + upe_w - upi - upi_w);
    const double nz2 = 0. + ni * (D.zi * D.zi);
    seec = seec - (D.zi * D.zi) * ni * tv * vol / nz2;
    const double t1y = .5 * D.cvgp * (a.get(PL_VY, ix, iy) * a.get(PL_GPIY, ix, iy) + a.get(PL_VY, ix, iy - 1) * a.get(PL_GPIY, ix, iy - 1) + 0. + 0.);
    seec = seec - D.fluxfacy * t1y * vol;
  }
  double psor, psorxr, psordis;
  f_psor(a, ix, iy, psor, psorxr, psordis);
  double resee = seec + 0. * te + 0. + 0. - 0.;
  resee = resee - (f_feex(a, ix, iy) - f_feex(a, ix1, iy) + D.fluxfacy * (a.get(PL_FEEY, ix, iy) - a.get(PL_FEEY, ix, iy - 1)));
  const double psorrgc = -psorxr;
  const double vsoree = -D.cfneut * D.cfneutsor_ee * D.cnsor * 13.6 * ev * D.fac2sp * psor + D.cfneut * D.cfneutsor_ee * D.cnsor * 13.6 * ev * D.fac2sp * psorrgc -
                        D.cfneut * D.cfneutsor_ee * D.cnsor * a.get(PL_ERLIZ, ix, iy) - D.cfneut * D.cfneutsor_ee * D.cnsor * a.get(PL_ERLRC, ix, iy) -
                        D.cfneut * D.cfneutsor_ee * D.cnsor * D.ediss * ev * (0.5 * psordis);
  const double w0 = vol * f_eqp(a, ix, iy) * (te - ti);
  resee = resee - w0 + vsoree;
  const int64_t c = (int64_t)(ix + NXS * iy) * NVX;
  out[2] = (1 - iseqalg[c + 2]) * resee / (vol * D.ennorm);
}

template <bool WIN>
__device__ void p2_i(const Acc<WIN>& a, int ix, int iy, double out[UE_NV], const int64_t* __restrict__ iseqalg) {
  const int NXS = a.NXS;
  const double ev = D.ev;
  const int ix1 = IXM1(ix, iy), ix2 = IXP1(ix, iy);
  const double vol = GG(vol, ix, iy), gx = GG(gx, ix, iy), gxf = GG(gxf, ix, iy), gxf_w = GG(gxf, ix1, iy);
  const double rrv = GG(rrv, ix, iy), rrv_w = GG(rrv, ix1, iy);
  const double te = a.get(PL_TE, ix, iy), ti = a.get(PL_TI, ix, iy), ng = a.get(PL_NG, ix, iy);
  const double up = a.get(PL_UP, ix, iy), up_w = a.get(PL_UP, ix1, iy);
  double seic = 0.;
  {  // oderhs.m:2514-2520, 2557-2574
    const double gx_e = GG(gx, ix2, iy), gx_w = GG(gx, ix1, iy);
    const double tv = a.get(PL_GPIX, ix, iy) / gxf;
    double t1 = a.get(PL_GPIX, ix1, iy) / gxf_w;
    t1 = .5 * D.cvgp * (up * rrv * d_ave(gx_e, gx) * tv + up_w * rrv_w * d_ave(gx, gx_w) * t1);
    seic = seic + D.cfvgpx * t1 * vol;
    const double t1y = .5 * D.cvgp * (a.get(PL_VY, ix, iy) * a.get(PL_GPIY, ix, iy) + a.get(PL_VY, ix, iy - 1) * a.get(PL_GPIY, ix, iy - 1) + 0. + 0.);
    const double t2y = t1y;
    seic = seic + D.fluxfacy * D.cfvgpy * t2y * vol;
  }
  double psor, psorxr, psordis;
  f_psor(a, ix, iy, psor, psorxr, psordis);
  double resei = seic + 0. * ti + 0. + 0. - 0.;
  resei = resei - (a.get(PL_FEIX, ix, iy) - a.get(PL_FEIX, ix1, iy) + D.fluxfacy * (a.get(PL_FEIY, ix, iy) - a.get(PL_FEIY, ix, iy - 1)));
  const double w0 = vol * f_eqp(a, ix, iy) * (te - ti);
  const double upi = f_upi_cut(a, ix, iy), upi_w = f_upi_cut(a, ix1, iy);
  const double us = upi + upi_w;
  resei = resei + w0 + D.cfneut * D.cfneutsor_ei * D.ctsor * 1.25e-1 * D.mi * (us * us) * D.fac2sp * psor + D.cfneut * D.cfneutsor_ei * D.ceisor * D.cnsor * D.eion * ev * psordis -
          D.cfneut * D.cfneutsor_ei * D.ccoldsor * ng * a.get(PL_NUCX, ix, iy) * (1.5 * ti - 0.125 * D.mi * (us * us) - D.eion * ev) * vol;
  {  // viscous heating (oderhs.m:4879-4930)
    const int ixn = IXM1(ix, iy + 1), ixs = IXM1(ix, iy - 1);
    const double thetacc = 0.5 * (0. + 0.);
    const double dupdx = gx * (upi - upi_w);
    double wvh = D.cfvcsx * D.cfvisx * ue_cos(thetacc) * a.get(PL_VISX, ix, iy) * (dupdx * dupdx);
    double dupdy;
    const int64_t isx = D.isxpty[ix + NXS * iy];
    const double up_n = f_upi_cut(a, ix, iy + 1), up_nw = f_upi_cut(a, ixn, iy + 1), up_s = f_upi_cut(a, ix, iy - 1), up_sw = f_upi_cut(a, ixs, iy - 1);
    if (isx == 0) dupdy = 0.5 * (upi + upi_w - up_s - up_sw) * GG(gyf, ix, iy - 1);
    else if (isx == -1) dupdy = 0.5 * (up_n + up_nw - upi - upi_w) * GG(gyf, ix, iy);
    else if (isx == 1 && D.isvhyha == 1) {
      const double upxavep1 = 0.5 * (up_n + up_nw), upxave0 = 0.5 * (upi + upi_w), upxavem1 = 0.5 * (up_s + up_sw);
      const double upf0 = 2. * upxavep1 * upxave0 * (upxavep1 + upxave0) / ((upxavep1 + upxave0) * (upxavep1 + upxave0) + D.upvhflr * D.upvhflr);
      const double upfm1 = 2. * upxave0 * upxavem1 * (upxave0 + upxavem1) / ((upxave0 + upxavem1) * (upxave0 + upxavem1) + D.upvhflr * D.upvhflr);
      dupdy = (upf0 - upfm1) * GG(gy, ix, iy);
    } else
      dupdy = 0.25 * ((up_n + up_nw - upi - upi_w) * GG(gyf, ix, iy) + (upi + upi_w - up_s - up_sw) * GG(gyf, ix, iy - 1));
    const double visy = f_visy(a, ix, iy);
    wvh = wvh + D.cfvcsy * D.cfvisy * visy * (dupdy * dupdy);
    wvh = wvh - ue_ksin(thetacc) * D.cfvcsy * D.cfvisy * visy * dupdx * dupdy;
    resei = resei + wvh * vol;
  }
  resei = resei + a.get(PL_PWRIBKG, ix, iy) * vol;
  const int64_t c = (int64_t)(ix + NXS * iy) * NVX;
  out[3] = (1 - iseqalg[c + 3]) * resei / (vol * D.ennorm);
}

// time-step term of the nksol equations (oderhs.m:7963-8037): rows of interior cells, and of the guard cells too when
// isbcwdt = 1; only for calls from nksol (yl(neq+1) < 0), never for the Jacobian evaluations
__device__ __forceinline__ void phase3_dt(int ix, int iy, double r[UE_NV], const double* ycell, double ylflag, int64_t c,
                                          const double* __restrict__ dtuse, const double* __restrict__ ylodt) {
  (void)iy;
  if (D.dtreal < 1.e15 && ylflag < 0) {
#pragma unroll
    for (int k = 0; k < UE_NV; ++k) {
      if (k >= NVX) continue;
      if (k == 1 && ix == D.nx + 2 * D.isbcwdt) continue;  // oderhs.m:7991: the algebraic up row at ix = nx unless isbcwdt = 1
      r[k] = (1. - 0.) * r[k];
      r[k] = r[k] - (ycell[k] - ylodt[c + k]) / dtuse[c + k];
    }
  }
}

// ============================================================================================
// phase 3 — rscalf (oderhs.m:8096-8210) and the time-step term (oderhs.m:7963-8037) on an interior cell
// ============================================================================================
template <bool WIN>
__device__ void phase3_interior(const Acc<WIN>& a, int ix, int iy, double r[UE_NV], const double* ycell /* this cell's entries of yl */, double ylflag /* yl(neq+1) */,
                                const int64_t* __restrict__ iseqalg, const double* __restrict__ dtuse, const double* __restrict__ ylodt,
                                bool recompute_resco_east = false /* fused phase-2/3 kernel: the plane entry of the east cell may not be written yet */) {
  const int NXS = a.NXS;
  const int64_t c = (int64_t)(ix + NXS * iy) * NVX;
  if (D.isflxvar != 1 && D.isrscalf == 1) {
    const double ni = a.get(PL_NI, ix, iy);
    double nbedot = 0., nbidot = 0.;
    nbidot = nbidot + r[0] * D.n0;
    nbedot = nbedot + D.zi * r[0] * D.n0;
    const double nbg2dot = D.isngon == 1 ? r[4] * D.n0g : 0.;  // oderhs.m:8118
    const int ix1 = IXP1(ix, iy);
    if (iseqalg[c + 1] == 0) {
      const int64_t c1 = (int64_t)(ix1 + NXS * iy) * NVX;
      double resco_e;
      if (recompute_resco_east) {
        double psor, psorxr, psordis;
        f_psor(a, ix1, iy, psor, psorxr, psordis);
        resco_e = f_resco(a, ix1, iy, psor, psorxr);
      } else resco_e = a.get(PL_RESCO, ix1, iy);
      const double yldot_np1 = resco_e / (GG(vol, ix1, iy) * D.n0);
      double nbvdot, nbv;
      if (iseqalg[c + 0] == 1) { nbvdot = yldot_np1 * D.n0; nbv = a.get(PL_NI, ix1, iy); }      // isnupdot1sd = 0
      else if (iseqalg[c1 + 0] == 1) { nbvdot = r[0] * D.n0; nbv = ni; }
      else { nbvdot = 0.5 * (r[0] + yldot_np1) * D.n0; nbv = 0.5 * (ni + a.get(PL_NI, ix1, iy)); }
      r[1] = (r[1] * D.n0 - ycell[1] * nbvdot) / nbv;
    }
    if (iseqalg[c + 2] == 0) r[2] = (r[2] * D.nnorm - ycell[2] * nbedot) / f_ne(a, ix, iy);
    if (iseqalg[c + 3] == 0) r[3] = (r[3] * D.nnorm - ycell[3] * (nbidot + D.cngtgx * nbg2dot)) / ((0. + ni) + D.cngtgx * a.get(PL_NG, ix, iy));
  }
  phase3_dt(ix, iy, r, ycell, ylflag, c, dtuse, ylodt);
}
#endif  // __CUDACC__
