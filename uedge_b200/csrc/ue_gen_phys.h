// uedge_b200/csrc/ue_gen_phys.h — the GENERAL residual of the hot path: pandf1 for the hydrogen family of switch sets
//   * one or two "ion" species (hydrogen ions + INERTIAL atoms, isupgon=1, nhsp=2) or diffusive atoms (isngon),
//   * orthogonal and NON-ORTHOGONAL meshes (isnonog=1: 5-point stencils fxm..fypx, vytan, fngxy, fmixy, feexy, feixy),
//   * every fd2tra scheme, the potential equation (isphion=1: calc_currents, calc_curr_cx, poteneq, phi boundary rows, isnewpot 0/1),
//   * every cross-field drift part (ExB, grad-B / curvature, diamagnetic, resistive, classical), B x grad(T) heat flows, Joule heating,
//   * the gas energy equation engbalg (istgon=1 for the inertial atoms),
//   * any subset of equations (isnion/isupon/isteon/istion/isngon/istgon/isphion), general idx* maps.
// Reference: bbb/convert.m:158-875 (convsr_vo, convsr_aux), bbb/oderhs.m:7-534 (fd2tra), :537-5070 (pandf), :5584-6648
// (neudif, neudifpg), :7508-7878 (engbalg), :7883-8213 (pandf1, rscalf), bbb/potencur.m:39-597 (calc_currents, calc_curr_cx, poteneq),
// bbb/boundary.m:4-3002 (bouncon), aph/aphrates.m (hydrogen rates).  Impurities, molecules, dnull / limiter are refused at init.
//
// EXECUTION MODEL.  One evaluation context = one `Gen` object: the constants and input-array pointers of the case plus pointers to
// ONE set of field planes in HBM and the identity of the cooperating threads.  The reference's loop nests become cooperative loops
// (FOR2 / FOR1: every nest two-dimensional, the row walks of convsr_aux flattened with XRQ) whose iterations are dealt out to the
// threads of the context and which begin with a barrier, so that everything an earlier nest wrote is visible; code between nests
// that writes fields runs on the context's first thread (SER).  On the GPU a context is a warp (Jacobian: one perturbed unknown,
// window ranges i1..i8 x j1..j8 of oderhs.m:868-1019, private planes that hold the band's rows only: rowlo..rowhi, inrow()), a
// thread block (full-domain residual on small meshes) or a co-resident grid with grid barriers (gridmode: large meshes);
// compiled for the host (tests/hostcheck) a context is one thread and the loops run in the reference's order - or reversed
// (UE_GEN_REVERSE), which exposes any dependence between iterations of one nest.
//
// Arithmetic: no FMA contraction, ue_math.h transcendental functions - the same bits on host and device.
#pragma once
#include <cmath>
#include <cstdint>

#include "ue_math.h"

#if defined(__CUDACC__)
#define HD __host__ __device__ inline
#else
#define HD inline
#endif

#define A(a, ix, iy) a[(ix) + NXS * (iy)]
#if defined(UE_GEN_REVERSE) && !defined(__CUDA_ARCH__)
#define UE_K_(k, n) ((n) - 1 - (k))
#else
#define UE_K_(k, n) (k)
#endif
// cooperative loop over iy = j0..j1 (outer), ix = i0..i1 (inner): starts with a barrier of the context
#if defined(__CUDA_ARCH__)
// row of iteration q in a nest w wide: exact for q < 2^22 (a float division instead of the ~40-instruction integer one)
#define UE_ROW_(q, w) __float2int_rz(__fdividef((float)(q) + 0.5f, (float)(w)))
#else
#define UE_ROW_(q, w) ((q) / (w))
#endif
#define FOR2(iy, j0, j1, ix, i0, i1)                                                                                               \
  for (int _w = (i1) - (i0) + 1, _h = (j1) - (j0) + 1, _n = (sync(), (_w > 0 && _h > 0) ? _w * _h : 0), _k = TID(), _q = 0, _r = 0, ix = 0, iy = 0; \
       _k < _n && ((_q = UE_K_(_k, _n)), (_r = UE_ROW_(_q, _w)), (iy = (j0) + _r), (ix = (i0) + _q - _r * _w), true); _k += nth)
#define FOR1(v, a, b) \
  for (int _n = (sync(), (b) - (a) + 1), _k = TID(), v = 0; _k < _n && ((v = (a) + UE_K_(_k, _n)), true); _k += nth)
#define FORXS(ix, xr) \
  for (int _q = 0, ix = 0; _q < (xr).n + ((xr).extra >= 0 ? 1 : 0) && ((ix = (_q < (xr).n ? (xr).first + _q * (xr).inc : (xr).extra)), true); ++_q)
#define SER if (sync(), TID() == 0)

template <typename T> HD T mx(T a, T b) { return a < b ? b : a; }
template <typename T> HD T mn(T a, T b) { return b < a ? b : a; }
HD int iabs(int a) { return a < 0 ? -a : a; }

// every field plane of one evaluation context, in slab order (P1: one plane, P2: one plane per species 1, 2)
#define UE_GEN_PLANES(P1, P2) \
  P1(ne) P1(nit) P1(nz2) P1(te) P1(ti) P1(phi) P1(ng) P1(tg) P1(pg) P1(pr) P1(pre) P1(zeff) P1(znot) \
  P2(ni) P2(nm) P2(up) P2(pri) P2(gpix) P2(gpiy) P2(niy0) P2(niy1) P2(priy0) P2(priy1) \
  P1(gprx) P1(gpry) P1(gpex) P1(gtex) P1(gtix) P1(gpey) P1(gtey) P1(gtiy) P1(ex) P1(ey) P1(nity0) P1(nity1) P1(ney0) P1(ney1) P1(tey0) P1(tey1) \
  P1(tiy0) P1(tiy1) P1(phiy0) P1(phiy1) P1(ngy0) P1(ngy1) P1(tgy0) P1(tgy1) P1(pgy0) P1(pgy1) P1(phiv) P1(tiv) P1(tev) P1(prev) P1(prtv) P2(priv) \
  P1(loglambda) P1(diffusivwrk) P2(vy) P2(vydd) P2(vygp) P2(v2) P2(v2dd) P2(v2xgp) P2(vytan) P1(frice) P2(frici) P2(upi) P2(uup) P2(uu) P1(upe) P1(vex) P1(vey) \
  P1(nuiz) P1(nurc) P1(nucx) P1(nuix) P1(psorbgg) P1(psorgc) P2(psorc) P1(psordis) P2(psorxrc) P1(psorrgc) P1(psorg) P2(psor) P2(psorxr) P1(psorrg) \
  P2(snic) P2(sniv) P2(psori) P2(smoc) P2(smov) P1(seec) P1(seev) P1(seic) P1(seiv) \
  P1(conxg) P1(conyg) P1(floxg) P1(floyg) P1(fngx) P1(fngy) P1(fngxy) P1(vygtan) P1(uug) P1(uuxg) P1(vyg) P1(resng) \
  P2(visx) P2(visy) P1(hcxe) P1(hcxi) P1(hcye) P1(hcyi) P2(hcxij) P2(hcyij) P1(hcxn) P1(hcyn) P1(hcxg) P1(hcyg) P1(eqp) P1(eqpg) P1(w0) P1(w1) P1(w2) P1(w3) P1(w) \
  P2(fnix) P2(fniy) P2(resco) P1(flox) P1(floy) P1(conx) P1(cony) P2(fmix) P2(fmiy) P2(fmixy) P2(resmo) P2(wvh) \
  P1(floxe) P1(floxi) P1(floye) P1(floyi) P1(conxe) P1(conxi) P1(conye) P1(conyi) P1(feex) P1(feey) P1(feix) P1(feiy) P1(feexy) P1(feixy) P1(resee) P1(resei) \
  P1(erliz) P1(erlrc) P1(eeli) P1(vsoreec) P1(vsoree) P1(pwribkg) P1(pwrebkg) P1(pradhyd) \
  P1(fqp) P1(fqx) P1(fqy) P1(fq2) P1(fqxb) P1(fqyb) P1(fqyn) P1(fqym) P1(fqymi) P1(fqya) P1(fqydt) P1(fqydti) P1(fqyao) P1(fqyae) P1(fqyd) P1(fqygp) P1(fq2d) P1(netap) P1(resphi) P1(dphi_iy1) \
  P2(vyce) P2(vycb) P2(vycp) P1(veycb) P2(v2ce) P2(v2cb) P1(ve2cb) P1(ve2cd) P2(v2cd) P1(vycf) P1(vycr) P1(wjdote) P1(segc) P1(floxge) P1(floyge) P1(conxge) P1(conyge) P1(fegx) P1(fegy) P1(fegxy) P1(reseg) P2(fmity) P2(fqymi_) P2(fniycbo) P1(feeycbo) P1(feiycbo) P1(kappal) P1(kappar) P1(bcel) P1(bcer) P1(bcil) P1(bcir) \
  P1(fqpsatlb) P1(fqpsatrb) P1(fdiaxlb) P1(fdiaxrb)

struct Gen {
  // ---- cooperative-thread identity of this context -------------------------------------------------------------------
  int nth;  // threads of the context: 1 (host), 32 (a warp), the block size, or - grid mode - all threads of a co-resident grid
  // Grid mode (full-domain residual on large meshes): the context is the whole grid of a cooperative launch.  Every block holds its
  // own copy of this struct in shared memory, all working on the same planes; the barrier between two loop nests is a grid barrier
  // (one atomic ticket per block on gbar, a spin on its acquire load, fences on both sides: what cooperative_groups' grid sync does).
  int gridmode;
  // rows the planes of this context hold: all of them (full evaluation, full-size private slabs) or the band of a Jacobian column whose
  // planes are stored band-rows-only (assign_planes_band).  Statements of the reference that touch FIXED rows or whole planes regardless
  // of the window (cosmetic zeroing of row ny+1, guard rows of fqya, ...) are skipped outside: nothing in the band reads them.
  int rowlo, rowhi;
  HD bool inrow(int iy) const { return iy >= rowlo && iy <= rowhi; }
  unsigned* gbar;  // monotonic ticket counter (zeroed before the launch)
  int* gflag;      // error flags raised by any thread, read by all after a barrier (uniform returns)
  HD int TID() const {
#if defined(__CUDA_ARCH__)
    if (gridmode) return (int)(blockIdx.x * blockDim.x + threadIdx.x);
    return nth > 32 ? (int)threadIdx.x : (int)(threadIdx.x & 31);
#else
    return 0;
#endif
  }
  HD void sync() const {
#if defined(__CUDA_ARCH__)
    if (gridmode) {
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        const unsigned nb = gridDim.x, ticket = atomicAdd(gbar, 1u);
        const unsigned target = (ticket / nb + 1u) * nb;
        unsigned seen, spins = 0;
        // (a block that never arrives would hang the device: after ~2^22 polls - seconds - the barrier gives up, raises flag 0x100 and
        //  every later barrier falls through; the host reports the evaluation as failed)
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(gbar) : "memory");
          if ((++spins & 1023u) == 0 && (spins > (1u << 22) || ((*(volatile int*)gflag) & 0x100))) { atomicOr(gflag, 0x100); break; }
        } while ((int)(seen - target) < 0);
        __threadfence();
      }
      __syncthreads();
    } else if (nth > 32) __syncthreads(); else __syncwarp();
#endif
  }
  // a thread-local error seen inside a nest becomes the same answer for every thread of the context (grid mode: no thread may leave
  // the evaluation on its own; the other modes keep the plain test)
  HD int any_flag(int mine, int bit) const {
#if defined(__CUDA_ARCH__)
    if (gridmode) {
      if (mine) atomicOr(gflag, bit);
      sync();
      return (*(volatile int*)gflag) & bit;
    }
#endif
    (void)bit;
    return mine;
  }
// ---- dimensions and switches ---------------------------------------------------------------------------------------
int nx, ny, NXS, NC, nisp, nusp, ngsp, nhsp, nfsp, iigsp;  // iigsp: 0-based species index of the inertial atoms (or -1)
int64_t neq;
int ixpt1, ixpt2, iysptrx1, iysptrx2, iysptrx, ixlb, ixrb, ixmp;
int xlinc, xrinc, yinc, isjaccorall;
int methn, methu, methe, methi, methg;
int isnonog, isphion, isphiofft, isupgon, isngon, istgon, ineudif, isflxvar, isrscalf, isbcwdt, icnuiz, icnucx, isrecmon, ingb, inflbg, isgasdc,
    isdifxg_aug, isdifyg_aug, isvylog, isgxvon, convis, concap, isflxlde, isflxldi, isplflxl, inkxc, isgpye, ishavisy, isvhyha, islnlamcon,
    isnupdot1sd, iteb, istabon, ifxnsgi, iflcore, ifluxni, isrefluxclip, ibctepl, ibctipl, ibctepr, ibctipr, isbohmms, isfixlb, isfixrb, isextrnp,
    isextrnpf, isextrtpf, isextrngc, isextrnw, isextrtw, isnfmiy, isybdrywd, isnewpot, jhswitch, isfeexpl0, isfeixpl0, isintlog, newbcl, newbcr,
    iskaplex, isnewpot_, isupwi_unused, iphibcc, isutcore, iphibcwi, iphibcwo, isexunif, isfdiax, isugfm1side, isvisxn_old;
int isnicore[2], isupcore[2], isngcore1, isupss[2], isnion[2], isupon[2], isteon, istion;
double ev, qe, me, mp, pi_, cutlo, rt8opi, temin, tgmin, nnorm, ennorm, temp0, vpnorm, lnlam, cfnus_i, cfnus_e, fcdif, cthe, flalftf, cfnetap,
    chioniz, sigvi_floor, cne_sgvi, cnuiz, cfrecom, cfdiss, cnucx, sigcx, rnn2cx, fnuizx, fnucxx, fnnuiz, cvgp, oldseec, cpgx, fracvgpgp, fluxfacy,
    coef, afix, flalfv, flgamv, kxe, kxi, ce, ci, rkxecore, kxicore, kye, kyi, kyet, kyit, ckyet, ckyit, lmfplim, alfkxi, alfkxe, flalfi, flalfe,
    flalfipl, flalfepl, lxtimax, lxtemax, tdiflim, cftiexclg, cfneut, cfneutsor_ei, cfneutsor_ee, cfneutsor_ni, cfneutsor_mi, cfneutdiv,
    cfneutdiv_fng, cfneutdiv_fmg, kxn, kyn, feqp, alfeqp, cnfx, cnfy, cnsor, cmfx, cmfy, cfaccony, fac2sp, cfmsor, flgam, cfcvte, cfcvti, cfjhf, cfloye,
    cfloyi, kye4order, kyi4order, bcee, bcei, chradi, chradr, ebind, ediss, eion, ctsor, ceisor, ccoldsor, cfvisx, cfvisy, upvhflr, tibg, pwribkg_c,
    tebg, pwrbkg_c, cflbg, difcng, alftng, gcfacgx, gcfacgy, flgamg, cngsor, erad, nurlxn, nurlxu, nurlxe, nurlxi, nurlxg, nurlxp, tcoree, tcorei, pcoree,
    pcorei, sygytotc, engbsr, csfacti, cfueb, cgpld, cmneut, eedisspl, eidisspl, cmntgpl, ckinfl, isoldalbarea, tbmin, recycm, nufak, dtreal, dtphi,
    delpert, dylconst, jaccliplim, kelhihg, kelhghg, lgvmax, flgamvg, cfvisxn, cfvisyn, flgamtg, cfupcx, cfticx, cfnidh, cfnidh2, cfnidhdis, cfnidhgy,
    cfnidhg2, cftgeqp, flalftxy, flalfgnx, flalfgny, nlimgx, nlimgy, cfloxiplt, cfloygwall, cfjve, rsigpl, rsigplcore, bcen, bceew, bciew, cfqym, cfqydt,
    cfjpy, cfjp2, cfqybf, cfq2bf, cfqybbo, cfqydbo, cfydd_, cfjp2_, cfqyn, cfqyao, cfqya, cfqyae, cfjpy_, fqpsatlb_unused, lnlam_unused, phiwi0, phiwo0,
    kappamx, kappa0, cfsigm, fqsatlb_u, fupe_cur, dtphi_, cfhcxgc_u, lyphi0, lyphi1, isparmultdt_u, tewallmin_u, cfwjdotelim, cfeexdbo, cfeixdbo, cfkincor,
    cfyef, cf2ef, cfybf, cf2bf, cfcurv, cfgradb, eycore, icoreelec, cfniybbo, cfeeybbo, cfydd, cf2dd, cfrd, cfbgt, cfvycf, cfvycr, cfeta1, cfrtaue, cfcl_e, cfcl_i, omgci_taui, omgce_taue, nuneo;
int isphilbc, isphirbc, isphicore0, isfqpave;
int ExtendedJacPhi, istgcore, istgpfc, istgwc, istglb, istgrb, isfegxyqflave;
int64_t numvar_;
double tgcore, cftgticore, tgwall, lytg1, lytg2, cftgtipltl, cftgtipltr, cftgtipfc, cftgtiwc, cgengmpl, cgengmw, cfalbedo, recyce, recycwe, cvgpg, cfcvtg, cfegxy, flalftgxy;
const double* idxtg_;
int rowuniform_;    // 1: every cell holds numvar unknowns in column order (the private state copies may then be cut to the window's columns)
const int* rowiv_;  // first unknown (0-based) of every mesh row, ny+3 entries (private copies of the state vector, ue_gen.cu)
double cngfx_[2], cngfy_[2], mi[2], zi[2], n0[2], fnorm[2], n0g_[2], mg_[2], ngbackg_[2], vcony[2], difpr[2], difni[2], difni2[2], difpr2[2], difax[2], travis[2], parvis[2],
    nlimix[2], nlimiy[2], dif4order[2], cpiup[2], cfvgpx[2], cfvgpy[2], cfvcsx[2], cfvcsy[2], cfvisxy[2], cngmom[2], cmwall[2], cngtgx[2], cngtgy[2], cdifg[2], lgmax[2], lgtmax[2],
    rld2dxg[2], rld2dyg[2], cngflox[2], cngfloy[2], rtg2ti[2], tgas[2], istgcon[2], keligig[2], ncore[2], ngcore[2], upcore[2], curcore[2], albedoc[2], csfaclb[2], csfacrb[2],
    recycp[2], nwimin[2], nwomin[2], difutm_[2];
// geometry planes / lines
const double *vol, *gx, *gy, *gxf, *gyf, *gxc, *gyc, *sx, *sxnp, *sy, *rr, *rrv, *volv, *syv, *dxnog, *dynog, *btot, *rbfbt, *rbfbt2, *lcone, *lconi, *angfx,
    *ngfix, *dx_, *dy_, *curvrby, *gradby, *curvrb2, *gradb2;
const double *fxm[2], *fx0[2], *fxp[2], *fxmy[2], *fxpy[2], *fym[2], *fy0[2], *fyp[2], *fymx[2], *fypx[2], *fymv[2], *fy0v[2], *fypv[2], *fymxv[2], *fypxv[2];
const double *ixm1d, *ixp1d, *isxptyd, *isxptxd;
const double *fgtdx, *fgtdy, *flalfea, *flalfia, *flalfva, *flalfgxa, *flalfgxya, *flalfgya, *flalfvgxa, *flalfvgya, *flalfvgxya, *flalftgxa, *flalftgya, *yyf;
const double *nwalli, *nwallo, *lytepf, *lytewc, *lytipf, *lytiwc, *lynipf, *lyniwc, *tewalli, *tiwalli, *tewallo, *tiwallo, *recylb, *recyrb, *alblb, *albrb,
    *recycwot, *recycwit, *fngysi, *fngyso, *fngyi_use, *fngyo_use, *fngxslb, *fngxsrb, *fngxlb_use, *fngxrb_use, *albedoi, *albedoo;
const double *istepfcix, *istipfcix, *isnwconiix, *isupwiix, *istewcix, *istiwcix, *isnwconoix, *isupwoix, *matwalli, *matwallo, *isixcore, *iseqalgd, *igyld;
const double *idxn_[2], *idxu_[2], *idxte_, *idxti_, *idxg_, *idxphi_;  // 1-based unknown numbers, 0 = equation off at that cell
// rate tables (istabon=10)
int mpe, mpd;
const double *wsveh, *wsveh0, *welms1, *welms2, *ekpt, *dkpt;
double rlemin, rlemax, rldmin, rldmax, delekpt, deldkpt;

HD int IXP1(int ix, int iy) { return (int)ixp1d[ix + NXS * iy]; }
HD int IXM1(int ix, int iy) { return (int)ixm1d[ix + NXS * iy]; }
HD int64_t IDXN(int f, int ix, int iy) { return (int64_t)idxn_[f][ix + NXS * iy] - 1; }   // -1: off
HD int64_t IDXU(int f, int ix, int iy) { return (int64_t)idxu_[f][ix + NXS * iy] - 1; }
HD int64_t IDXTE(int ix, int iy) { return (int64_t)idxte_[ix + NXS * iy] - 1; }
HD int64_t IDXTI(int ix, int iy) { return (int64_t)idxti_[ix + NXS * iy] - 1; }
HD int64_t IDXG(int ix, int iy) { return (int64_t)idxg_[ix + NXS * iy] - 1; }
HD int64_t IDXTG(int ix, int iy) { return (int64_t)idxtg_[ix + NXS * iy] - 1; }
HD int64_t IDXPHI(int ix, int iy) { return (int64_t)idxphi_[ix + NXS * iy] - 1; }
HD int ALG(int64_t iv) { return (int)iseqalgd[iv]; }

HD double ave(double t0, double t1) { return 2 * t0 * t1 / (cutlo + t0 + t1); }  // oderhs.m:697
HD double sgn(double a, double b) { return copysign(fabs(a), b); }     // Fortran sign(a,b)
HD double sq(double x) { return x * x; }
// classical (Braginskii) collisional factors of pandf's simple model (oderhs.m:1157-1166), evaluated where they are used
HD double eta1_(int ix, int iy) const { return cfeta1 * 0.3 * A(nm[0], ix, iy) * A(ti, ix, iy) * (1 / (qe * A(btot, ix, iy))) / omgci_taui; }
HD double rtaue_(int ix, int iy) const { return cfrtaue * (1 / (qe * A(btot, ix, iy))) / omgce_taue; }
HD double dclass_i_(int ix, int iy) const { return cfcl_i == 0. ? 0. : cfcl_i * eta1_(ix, iy) / (0.3 * A(nm[0], ix, iy)); }
HD double dclass_e_(int ix, int iy) const { return cfcl_e == 0. ? 0. : cfcl_e * A(te, ix, iy) * rtaue_(ix, iy); }
// perpendicular resistivity (statement function of pandf, oderhs.m:698)
HD double etaper(int ix, int iy) const { return 3.234e-9 * A(loglambda, ix, iy) / ue_pow(mx(A(te, ix, iy), temin * ev) / (1000. * ev), 1.5); }
HD double powi(double x, int64_t n) { double r = 1.0; while (n > 0) { if (n & 1) r *= x; x *= x; n >>= 1; } return r; }

// ---- hydrogen rates (aph/aphrates.m), istabon 0 / 7 / 10: same restatement as ue_oracle.cpp ----------------------
HD void table_idx(double tev_j, double dens, int& je, int& jd, double& fje, double& fjd) {  // aph/aphrates.m:1043-1056
  double zloge = ue_log(tev_j / ev);
  double rle = mx(rlemin, mn(zloge, rlemax));
  double zlogd = ue_log10(dens);
  double rld = mx(rldmin, mn(zlogd, rldmax));
  je = (int)((rle - rlemin) / delekpt) + 1; je = mn(je, mpe - 1);
  jd = (int)((rld - rldmin) / deldkpt) + 1; jd = mn(jd, mpd - 1);
  fje = (rle - ekpt[je - 1]) / (ekpt[je] - ekpt[je - 1]);
  fjd = (rld - dkpt[jd - 1]) / (dkpt[jd] - dkpt[jd - 1]);
}
HD double table_val(const double* w, double tev_j, double dens) {
  int je, jd; double fje, fjd;
  table_idx(tev_j, dens, je, jd, fje, fjd);
  auto W = [&](int a, int b) { return ue_log(w[(a - 1) + mpe * (b - 1)]); };
  double r11 = W(je, jd), r12 = W(je, jd + 1), r21 = W(je + 1, jd), r22 = W(je + 1, jd + 1);
  double r1 = r11 + fjd * (r12 - r11);
  double r2 = r21 + fjd * (r22 - r21);
  return ue_exp(r1 + fje * (r2 - r1));
}
HD double sionf(double temp, double den) {  // aph/aphrates.m:1133-1176 (R.B. Campbell's fits, istabon=7)
  auto ain = [](double x) { return -49.05905 + 2.51313783 * x - 0.049159714 * x * x; };
  auto bin = [](double x) { return 41.1855162 - 2.3298672 * x + 4.24769144e-2 * x * x; };
  auto cin = [](double x) { return -32.798921 + 1.72102919 * x - 0.038692357 * x * x; };
  auto din = [](double x) { return 27.370466 - 1.6824361 * x + 0.0462317894 * x * x; };
  auto ein = [](double x) { return -7.9990454 + 0.127573157 * x - 6.3586911e-3 * x * x; };
  auto gin = [](double x) { return -4.5832951 + 0.776264783 * x - 1.8866089e-2 * x * x; };
  auto hin = [](double x) { return 3.08056833 - 0.39114789 * x + 9.86833304e-3 * x * x; };
  auto riin = [](double x) { return -0.4648639 + 0.0551428018 * x - 1.404213e-3 * x * x; };
  double x = mn(22.e0, ue_log10(den)), y = ue_log10(temp);
  return ue_pow(10., ain(x) + bin(x) * y + cin(x) * y * y + din(x) * y * y * y + ein(x) * y * y * y * y + gin(x) * y * y * y * y * y +
                         hin(x) * y * y * y * y * y * y + riin(x) * y * y * y * y * y * y * y);
}
HD double srecf(double temp, double den) {  // aph/aphrates.m:1180-1226
  auto ar = [](double x) { return -0.4575652 - 2.144012 * x + 6.7072142e-2 * x * x - 1.391667e-4 * x * x * x; };
  auto br = [](double x) { return -121.8401 + 18.001822 * x - 0.8679488 * x * x + 1.33165e-2 * x * x * x; };
  auto cr = [](double x) { return 80.897256 - 13.29602 * x + 0.71881414 * x * x - 0.0126549 * x * x * x; };
  auto dr = [](double x) { return 56.406823 - 7.301996 * x + 0.29339793 * x * x - 3.50898e-3 * x * x * x; };
  auto er = [](double x) { return -55.73559 + 7.9634283 * x - 0.370274 * x * x + 5.567961e-3 * x * x * x; };
  auto gr = [](double x) { return 10.866692 - 1.584193 * x + 0.07563791 * x * x - 1.177562e-3 * x * x * x; };
  double x = mn(22.e0, ue_log10(den)), y = ue_log10(temp);
  return ue_pow(10., ar(x) + br(x) * y + cr(x) * y * y + dr(x) * y * y * y + er(x) * y * y * y * y + gr(x) * y * y * y * y * y);
}
HD double svradp(double temp, double den) {  // aph/aphrates.m:1230-1300
  auto ai = [](double x) { return -275.845 + 37.010817 * x - 1.788045 * x * x + 0.029078333 * x * x * x; };
  auto bi = [](double x) { return 2200.9478 - 326.1153 * x + 16.148655 * x * x - 0.2660702 * x * x * x; };
  auto ci2 = [](double x) { return -2.935221e3 + 4.3757698e2 * x - 21.73964 * x * x + 0.358962 * x * x * x; };
  auto di = [](double x) { return 1604.1466 - 239.6959 * x + 11.923707 * x * x - 0.1970501 * x * x * x; };
  auto ei = [](double x) { return -390.8635 + 58.474495 * x - 2.910997 * x * x + 0.048133829 * x * x * x; };
  auto gi = [](double x) { return 35.012574 - 5.24202 * x + 0.26109962 * x * x - 4.319238e-3 * x * x * x; };
  auto ae = [](double x) { return 2860.4173 - 610.2452 * x + 48.275821 * x * x - 1.687994 * x * x * x + 0.02201375 * x * x * x * x; };
  auto be = [](double x) { return 10612.067 - 2046.397 * x + 147.73914 * x * x - 4.729973 * x * x * x + 0.056671796 * x * x * x * x; };
  auto ce2 = [](double x) { return -4.231708e4 + 8494.6102 * x - 639.0226 * x * x + 21.350311 * x * x * x - 0.2673466 * x * x * x * x; };
  auto de = [](double x) { return -8.385144e3 + 1887.6244 * x - 157.8502 * x * x + 5.820501 * x * x * x - 0.07992837 * x * x * x * x; };
  auto ee = [](double x) { return 3.938282e4 - 8.131339e3 * x + 628.8119 * x * x - 21.58636 * x * x * x + 0.27756029 * x * x * x * x; };
  auto ge = [](double x) { return -1.038281e4 + 2.1349333e3 * x - 164.4201 * x * x + 5.6210487 * x * x * x - 0.07197622 * x * x * x * x; };
  auto sionfl = [&](double x, double y) { return ue_pow(10., ai(x) + bi(x) * y + ci2(x) * y * y + di(x) * y * y * y + ei(x) * y * y * y * y + gi(x) * y * y * y * y * y); };
  auto etai = [&](double x, double y) {
    return (ue_pow(10., ae(x) + be(x) * y + ce2(x) * y * y + de(x) * y * y * y + ee(x) * y * y * y * y + ge(x) * y * y * y * y * y)) / sionfl(x, y);
  };
  double x = mn(22.e0, ue_log10(den)), y = ue_log10(temp);
  return mx(0.e0, (13.6e0 + etai(x, mn(2.e0, y)))) * 1.602e-19 * sionfl(x, y);
}
HD double rsa(double tej, double dens) {  // aph/aphrates.m:872-1131
  if (istabon == 0) { double a = tej / (10 * ev); return 3.0e-14 * a * a / (3.0 + a * a); }
  if (istabon == 7) return sionf(tej / ev, dens);
  return table_val(wsveh, tej, dens);
}
HD double rra(double tej, double dens) {  // aph/aphrates.m:617-870
  if (istabon == 0) return 0.;
  if (istabon == 7) return srecf(tej / ev, dens);
  return table_val(wsveh0, tej, dens);
}
HD double rcx(double t0) { double a = 3 * t0 / (10 * ev); return 1.7e-14 * ue_pow(a, 0.333); }  // aph/aphrates.m:395-399
HD double rqa0(double tej) { double a = tej / (10 * ev); return erad * ev * 3.0e-14 * a * a / (3.0 + a * a); }  // :444-447
HD double erl1(double tej, double dens) {  // aph/aphrates.m:2-147
  if (istabon == 0) return (rqa0(tej) - 13.6 * ev * rsa(tej, dens)) * dens;
  if (istabon == 7) return (svradp(tej / ev, dens) - 13.6 * ev * rsa(tej, dens)) * dens;
  return table_val(welms1, tej, dens);
}
HD double erl2(double tej, double dens) {  // aph/aphrates.m:149-294
  if (istabon == 0 || istabon == 7) return (13.6 * ev + 1.5 * tej) * dens * rra(tej, dens);
  return table_val(welms2, tej, dens);
}

// ---- index window (oderhs.m:868-1019) ----------------------------------------------------------------------------
struct Win {
  int xc, yc;
  int i1, i2, i2p, i3, i4, i5, i5m, i6, i7, i8;
  int j1, j1p, j2, j2p, j3, j4, j5, j5m, j6, j5p, j6p, j7, j8;
  int ixs, ixf, iys, iyf, ixs1, ixf6, iys1, iyf6;
  bool openbox, xcnearlb, xcnearrb, xccuts;
};
HD Win make_win(int xc, int yc) {
  Win w; w.xc = xc; w.yc = yc;
  if (xc < 0 || ((0 <= yc) && (yc - yinc <= 0) && isjaccorall == 1)) {
    w.i1 = 0; w.i2 = 1; w.i2p = 1; w.i3 = 0; w.i4 = 0; w.i5 = nx; w.i5m = nx - 1; w.i6 = nx + 1; w.i7 = nx + 1; w.i8 = nx + 1;
  } else {
    w.i1 = mx(0, xc - xlinc - 1); w.i2 = mx(1, xc - xlinc); w.i2p = mx(1, xc - xrinc - 1);
    w.i3 = xc - xlinc; w.i4 = mx(0, xc - xlinc); w.i5 = mn(nx, xc + xrinc); w.i5m = mn(nx - 1, xc + xrinc);
    w.i6 = mn(nx + 1, xc + xrinc + 1); w.i7 = xc + xrinc; w.i8 = mn(nx + 1, xc + xrinc);
  }
  if (yc < 0) {
    w.j1 = 0; w.j1p = 0; w.j2 = 1; w.j2p = 1; w.j3 = 0; w.j4 = 0; w.j5 = ny; w.j5m = ny - 1; w.j6 = ny + 1; w.j5p = ny;
    w.j6p = ny + 1; w.j7 = ny + 1; w.j8 = ny + 1;
  } else {
    w.j1 = mx(0, yc - yinc - 1); w.j2 = mx(1, yc - yinc); w.j1p = mx(0, yc - yinc - 2);
    w.j2p = mx(1, yc - yinc - 1); w.j3 = yc - yinc; w.j4 = mx(0, yc - yinc); w.j5 = mn(ny, yc + yinc);
    w.j5m = mn(ny - 1, yc + yinc); w.j6 = mn(ny + 1, yc + yinc); w.j5p = mn(ny, yc + yinc + 1);
    w.j6p = mn(ny + 1, yc + yinc + 1); w.j7 = yc + yinc; w.j8 = mn(ny + 1, yc + yinc);
  }
  w.xccuts = false;
  if ((xc - xlinc <= ixpt1 + 1) && (xc + xrinc + 1 >= ixpt1) && (yc - yinc <= iysptrx1) && (iysptrx1 > 0)) w.xccuts = true;
  if ((xc - xlinc <= ixpt2 + 1) && (xc + xrinc + 1 >= ixpt2) && (yc - yinc <= iysptrx2) && (iysptrx2 > 0)) w.xccuts = true;
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
  w.xcnearlb = ((xc - xlinc <= ixlb) && (xc + xrinc >= ixlb)) || xc < 0;
  w.xcnearrb = ((xc - xlinc <= ixrb + 1) && (xc + xrinc >= ixrb)) || xc < 0;
  return w;
}

// ix visited by "do ix = ixm1(is,jstart), min(nx,ie), inc" with inc from row jinc (convert.m:583-584 etc.), plus one optional extra
struct XR { int first, inc, n, extra; };
HD XR xrange(int is, int ie, int jinc, int jstart) {
  XR r; r.extra = -1;
  int d = ie - IXM1(ie, jinc);
  r.inc = mx(1, iabs(d)); if (d < 0) r.inc = -r.inc;
  r.first = IXM1(is, jstart);
  const int last = mn(nx, ie);
  r.n = r.inc > 0 ? (last >= r.first ? (last - r.first) / r.inc + 1 : 0) : (r.first >= last ? (r.first - last) / (-r.inc) + 1 : 0);
  return r;
}

// element q of the sequence FORXS walks (regular elements, then the extra one); false beyond its end.  XRQ opens the body of a
// FOR2 nest over (row, q): the rows' sequences are walked by different threads, q < xw bounds every row's length
HD bool xr_at(const XR& r, int q, int& ix) const {
  if (q < r.n) { ix = r.first + q * r.inc; return true; }
  if (q == r.n && r.extra >= 0) { ix = r.extra; return true; }
  return false;
}
#define XRQ(ix, xs, q) int ix = 0; if (!xr_at(xs, q, ix)) continue;

// 5-point stencil of the non-orthogonal mesh at the y-face above cell (ix,iy), side k (convert.m:422-482)
struct St5 { int c[5]; double f[5]; };
HD St5 stx(int ix, int iy, int k) {
  St5 s; const int c = ix + NXS * iy;
  s.c[0] = IXM1(ix, iy + k) + NXS * (iy + k); s.f[0] = fxm[k][c];
  s.c[1] = ix + NXS * (iy + k); s.f[1] = fx0[k][c];
  s.c[2] = IXP1(ix, iy + k) + NXS * (iy + k); s.f[2] = fxp[k][c];
  s.c[3] = IXM1(ix, iy + 1 - k) + NXS * (iy + 1 - k); s.f[3] = fxmy[k][c];
  s.c[4] = IXP1(ix, iy + 1 - k) + NXS * (iy + 1 - k); s.f[4] = fxpy[k][c];
  return s;
}
HD double st_lin(const St5& s, const double* a) { return s.f[0] * a[s.c[0]] + s.f[1] * a[s.c[1]] + s.f[2] * a[s.c[2]] + s.f[3] * a[s.c[3]] + s.f[4] * a[s.c[4]]; }
HD double st_log(const St5& s, const double* a) {
  return ue_exp(s.f[0] * ue_log(a[s.c[0]]) + s.f[1] * ue_log(a[s.c[1]]) + s.f[2] * ue_log(a[s.c[2]]) + s.f[3] * ue_log(a[s.c[3]]) + s.f[4] * ue_log(a[s.c[4]]));
}
HD double st_inv(const St5& s, const double* a) { return 1 / (s.f[0] / a[s.c[0]] + s.f[1] / a[s.c[1]] + s.f[2] / a[s.c[2]] + s.f[3] / a[s.c[3]] + s.f[4] / a[s.c[4]]); }
// difference across the x-face (ix,iy) of the y-stencil values: "grdnv" numerators (oderhs.m:1408-1419, 4333-4343, ...);
// mode 0 linear, 1 log, 2 inverse
HD double grdnv_y(const double* a, int ix, int iy, int mode) {
  const int c = ix + NXS * iy;
  const int iy1 = mx(0, iy - 1), iy2 = mn(ny + 1, iy + 1);
  const int ix2 = IXP1(ix, iy), ix4 = IXP1(ix, iy1), ix6 = IXP1(ix, iy2);
  auto g = [&](double v) { return mode == 1 ? ue_log(v) : (mode == 2 ? 1 / v : v); };
  const double hi = fym[1][c] * g(A(a, ix2, iy1)) + fy0[1][c] * g(A(a, ix2, iy)) + fyp[1][c] * g(A(a, ix2, iy2)) + fymx[1][c] * g(A(a, ix, iy1)) + fypx[1][c] * g(A(a, ix, iy2));
  const double lo = fym[0][c] * g(A(a, ix, iy1)) + fy0[0][c] * g(A(a, ix, iy)) + fyp[0][c] * g(A(a, ix, iy2)) + fymx[0][c] * g(A(a, ix4, iy1)) + fypx[0][c] * g(A(a, ix6, iy2));
  if (mode == 2) return 1 / hi - 1 / lo;
  return hi - lo;
}
HD double upwind(double f, double p1, double p2) { return mx(f, 0.0) * p1 + mn(f, 0.0) * p2; }  // oderhs.m:81

// ---- all mutable state + the routines that touch it (one instance per worker thread) ------------------------------
// ---- field arrays (one private set per evaluation context) ----
  // Compla / Gradients / Comflo / Conduc / Rhsides / Locflux groups of bbb/bbb.v.  Species-indexed fields are arrays of
  // planes [ifld-1]; gas fields exist for gas species 1 only (ngsp = 1).
  double *ne, *nit, *nz2, *te, *ti, *phi, *ng, *tg, *pg, *pr, *pre, *zeff, *znot;
  double *ni[2], *nm[2], *up[2], *pri[2], *gpix[2], *gpiy[2], *niy0[2], *niy1[2], *priy0[2], *priy1[2];
  double *gprx, *gpry, *gpex, *gtex, *gtix, *gpey, *gtey, *gtiy, *ex, *ey, *nity0, *nity1, *ney0, *ney1, *tey0, *tey1, *tiy0, *tiy1, *phiy0, *phiy1;
  double *ngy0, *ngy1, *tgy0, *tgy1, *pgy0, *pgy1, *phiv, *tiv, *tev, *prev, *prtv, *priv[2];
  double *loglambda, *diffusivwrk, *vy[2], *vydd[2], *vygp[2], *v2[2], *v2dd[2], *v2xgp[2], *vytan[2], *frice, *frici[2], *upi[2], *uup[2], *uu[2], *upe, *vex, *vey;
  double *nuiz, *nurc, *nucx, *nuix, *psorbgg, *psorgc, *psorc[2], *psordis, *psorxrc[2], *psorrgc, *psorg, *psor[2], *psorxr[2], *psorrg;
  double *snic[2], *sniv[2], *psori[2], *smoc[2], *smov[2], *seec, *seev, *seic, *seiv;
  double *conxg, *conyg, *floxg, *floyg, *fngx, *fngy, *fngxy, *vygtan, *uug, *uuxg, *vyg, *resng;
  double *visx[2], *visy[2], *hcxe, *hcxi, *hcye, *hcyi, *hcxij[2], *hcyij[2], *hcxn, *hcyn, *hcxg, *hcyg, *eqp, *eqpg, *w0, *w1, *w2, *w3, *w;
  double *fnix[2], *fniy[2], *resco[2], *flox, *floy, *conx, *cony, *fmix[2], *fmiy[2], *fmixy[2], *resmo[2], *wvh[2];
  double *floxe, *floxi, *floye, *floyi, *conxe, *conxi, *conye, *conyi, *feex, *feey, *feix, *feiy, *feexy, *feixy, *resee, *resei;
  double *erliz, *erlrc, *eeli, *vsoreec, *vsoree, *pwribkg, *pwrebkg, *pradhyd;
  double *fqp, *fqx, *fqy, *fq2, *fqxb, *fqyb, *fqyn, *fqym, *fqymi, *fqya, *fqydt, *fqydti, *fqyao, *fqyae, *fqyd, *fqygp, *fq2d, *fqpsatlb_, *netap, *resphi, *dphi_iy1;
  double *fniycbo[2], *feeycbo, *feiycbo, *kappal, *kappar, *bcel, *bcer, *bcil, *bcir, *fqpsatlb, *fqpsatrb, *fdiaxlb, *fdiaxrb;
  double *dtuse, *ylodt, *suscal, *sfscal;
  int errc;

  // ---- convsr_vo (convert.m:158-375) ------------------------------------------------------------------------------
  HD int convsr_vo(int ixl, int iyl, const double* yl) {
    int is, ie, js, je;
    if (ixl < 0 || yinc >= 6) { is = 0; ie = nx + 1; } else { is = ixl; ie = ixl; }
    if (iyl < 0 || yinc >= 6) { js = 0; je = ny + 1; } else { js = iyl; je = iyl; }
    if (ixl < 0 && iyl >= 0) { js = mx(0, iyl - yinc); je = mn(ny + 1, iyl + yinc); }
    FOR2(iy, js, je, ix, is, ie) { A(ne, ix, iy) = 0.; A(nit, ix, iy) = 0.; A(nm[0], ix, iy) = 0.; A(nz2, ix, iy) = 0.; }
    int inegni = 0, inegng = 0;
    for (int f = 0; f < nisp; ++f)
      FOR2(iy, js, je, ix, is, ie) {
          const int64_t iv = IDXN(f, ix, iy);
          if (iv >= 0) {
            A(ni[f], ix, iy) = yl[iv] * n0[f];
            if (A(ni[f], ix, iy) < 0) inegni = 1;
          }
          A(ne, ix, iy) = A(ne, ix, iy) + zi[f] * A(ni[f], ix, iy);
          if (isupgon == 1 && zi[f] == 0) A(ng, ix, iy) = A(ni[f], ix, iy);
          else {
            A(nit, ix, iy) = A(nit, ix, iy) + A(ni[f], ix, iy);
            A(nz2, ix, iy) = A(nz2, ix, iy) + A(ni[f], ix, iy) * (zi[f] * zi[f]);
          }
          A(nm[f], ix, iy) = A(ni[f], ix, iy) * mi[f];
        }
    FOR2(iy, js, je, ix, is, ie) {
        double ntemp = A(ne, ix, iy);
        if (isflxvar == 0) ntemp = nnorm;
        int64_t iv = IDXTE(ix, iy);
        if (iv >= 0) { A(te, ix, iy) = yl[iv] * ennorm / (1.5 * ntemp); A(te, ix, iy) = mx(A(te, ix, iy), temin * ev); }
        iv = IDXG(ix, iy);
        if (iv >= 0) { A(ng, ix, iy) = yl[iv] * n0g_[0]; if (A(ng, ix, iy) < 0) inegng = 1; }
        ntemp = A(nit, ix, iy) + cngtgx[0] * A(ng, ix, iy);
        if (isflxvar == 0) ntemp = nnorm;
        iv = IDXTI(ix, iy);
        if (iv >= 0) { A(ti, ix, iy) = yl[iv] * ennorm / (1.5 * ntemp); A(ti, ix, iy) = mx(A(ti, ix, iy), temin * ev); }
        iv = IDXTG(ix, iy);  // convert.m:299-305 (isflxvar = 0: ntemp = n0g)
        if (iv >= 0) { A(tg, ix, iy) = yl[iv] * ennorm / (1.5 * n0g_[0]); A(tg, ix, iy) = mx(A(tg, ix, iy), tgmin * ev); }
        iv = IDXPHI(ix, iy);
        if (iv >= 0) A(phi, ix, iy) = yl[iv] * temp0;
      }
    if (any_flag(inegni, 1)) { errc = 1; return -3; }
    if (any_flag(inegng, 2)) { errc = 2; return -3; }
    for (int f = 0; f < nusp; ++f)
      FOR2(iy, js, je, ix, is, ie)
          if (IDXU(f, ix, iy) >= 0) {
            int ix1 = IXP1(ix, iy), ix2 = mx(0, IXM1(ix, iy));
            double t1 = 0.5 * (A(nm[f], ix2, iy) + A(nm[f], ix, iy)), t2 = 0.5 * (A(nm[f], ix, iy) + A(nm[f], ix1, iy));
            if (isflxvar == 0 || isflxvar == 2) { t1 = mi[f] * n0[f]; t2 = mi[f] * n0[f]; }
            // (a neighbour whose momentum equation is off keeps its stored velocity)
            if (IDXU(f, ix2, iy) >= 0) A(up[f], ix2, iy) = yl[IDXU(f, ix2, iy)] * fnorm[f] / t1;
            A(up[f], ix, iy) = yl[IDXU(f, ix, iy)] * fnorm[f] / t2;
          }
    return 0;
  }

  // ---- convsr_aux (convert.m:379-875) -------------------------------------------------------------------------------
  HD void convsr_aux(int ixl, int iyl) {
    int is, ie, js, je;
    if (ixl < 0 || yinc >= 6) { is = 0; ie = nx + 1; } else { is = ixl; ie = ixl; }
    if (iyl < 0 || yinc >= 6) { js = 0; je = ny + 1; } else { js = iyl; je = iyl; }
    if (ixl < 0 && iyl >= 0) { js = mx(0, iyl - yinc); je = mn(ny + 1, iyl + yinc); }
    FOR2(iy, js, je, ix, is, ie) { A(pr, ix, iy) = 0.; A(zeff, ix, iy) = 0.; }
    for (int f = 0; f < nisp; ++f)
      FOR2(iy, js, je, ix, is, ie) {
          A(pri[f], ix, iy) = A(ni[f], ix, iy) * A(ti, ix, iy);
          if (f == iigsp && istgon == 1) A(pri[f], ix, iy) = A(ni[f], ix, iy) * A(tg, ix, iy);
          if (zi[f] != 0.) {
            A(pr, ix, iy) = A(pr, ix, iy) + A(pri[f], ix, iy);
            A(zeff, ix, iy) = A(zeff, ix, iy) + (zi[f] * zi[f]) * A(ni[f], ix, iy);
          }
        }
    FOR2(iy, js, je, ix, is, ie) {
        A(pre, ix, iy) = A(ne, ix, iy) * A(te, ix, iy);
        A(pr, ix, iy) = A(pr, ix, iy) + A(pre, ix, iy);
        A(zeff, ix, iy) = A(zeff, ix, iy) / A(ne, ix, iy);
        A(znot, ix, iy) = A(ne, ix, iy) * A(zeff, ix, iy) / A(ni[0], ix, iy) - 1;
        if (istgcon[0] > -1.e-20) A(tg, ix, iy) = (1 - istgcon[0]) * rtg2ti[0] * A(ti, ix, iy) + istgcon[0] * tgas[0] * ev;
        A(pg, ix, iy) = A(ng, ix, iy) * A(tg, ix, iy);
      }
    const int xw = (is == ie) ? 3 : ie - is + 3;  // upper bound of the length of a row's x-sequence (two neighbours + the extra one, or the whole row)
    FOR2(iy, js, je, q_, 0, xw - 1) { XR xs = xrange(is, ie, iy, iy); XRQ(ix, xs, q_) A(gprx, ix, iy) = 0.0; }
    const int jlo = mx(js - 1, 0), jhi = mn(ny, je);
    FOR2(iy, jlo, jhi, q_, 0, xw - 1) {
      XR xs = xrange(is, ie, js, js);
      xs.extra = IXP1(ie, iy);
      XRQ(ix, xs, q_) { A(ney0, ix, iy) = 0.; A(ney1, ix, iy) = 0.; A(nity0, ix, iy) = 0.; A(nity1, ix, iy) = 0.; A(gpry, ix, iy) = 0.; }
    }
    for (int f = 0; f < nisp; ++f)
      FOR2(iy, js, je, q_, 0, xw - 1) {
        XR xs = xrange(is, ie, iy, iy);
        XRQ(ix, xs, q_) {
          int ix1 = IXP1(ix, iy);
          A(gpix[f], ix, iy) = (A(pri[f], ix1, iy) - A(pri[f], ix, iy)) * A(gxf, ix, iy);
          if (zi[f] != 0.) A(gprx, ix, iy) = A(gprx, ix, iy) + A(gpix[f], ix, iy);
        }
      }
    auto yface_ion = [&](int f, int ix, int iy) {  // convert.m:631-666
      const St5 s0 = stx(ix, iy, 0), s1 = stx(ix, iy, 1);
      A(niy0[f], ix, iy) = st_log(s0, ni[f]);
      A(niy1[f], ix, iy) = st_log(s1, ni[f]);
      A(nity0, ix, iy) = A(nity0, ix, iy) + A(niy0[f], ix, iy);
      A(nity1, ix, iy) = A(nity1, ix, iy) + A(niy1[f], ix, iy);
      A(ney0, ix, iy) = A(ney0, ix, iy) + zi[f] * A(niy0[f], ix, iy);
      A(ney1, ix, iy) = A(ney1, ix, iy) + zi[f] * A(niy1[f], ix, iy);
      A(priy0[f], ix, iy) = st_log(s0, pri[f]);
      A(priy1[f], ix, iy) = st_log(s1, pri[f]);
      A(gpiy[f], ix, iy) = (A(priy1[f], ix, iy) - A(priy0[f], ix, iy)) / A(dynog, ix, iy);
      if (zi[f] != 0.) A(gpry, ix, iy) = A(gpry, ix, iy) + A(gpiy[f], ix, iy);
    };
    for (int f = 0; f < nisp; ++f)
      FOR2(iy, jlo, jhi, q_, 0, xw - 1) { XR xs = xrange(is, ie, js, js); xs.extra = IXP1(ie, iy); XRQ(ix, xs, q_) yface_ion(f, ix, iy); }
    FOR2(iy, jlo, jhi, q_, 0, xw - 1) {  // convert.m:669-700
      XR xs = xrange(is, ie, js, js); xs.extra = IXP1(ie, iy);
      XRQ(ix, xs, q_) {
        const St5 s0 = stx(ix, iy, 0), s1 = stx(ix, iy, 1);
        A(tey0, ix, iy) = st_lin(s0, te); A(tey1, ix, iy) = st_lin(s1, te);
        A(tiy0, ix, iy) = st_lin(s0, ti); A(tiy1, ix, iy) = st_lin(s1, ti);
        A(phiy0, ix, iy) = st_lin(s0, phi); A(phiy1, ix, iy) = st_lin(s1, phi);
      }
    }
    FOR2(iy, jlo, jhi, q_, 0, xw - 1) {  // convert.m:703-717
      XR xs = xrange(is, ie, js, js); xs.extra = IXP1(ie, iy);
      XRQ(ix, xs, q_) {
        const St5 s0 = stx(ix, iy, 0), s1 = stx(ix, iy, 1);
        A(ngy0, ix, iy) = st_log(s0, ng); A(ngy1, ix, iy) = st_log(s1, ng);
        A(tgy0, ix, iy) = st_lin(s0, tg); A(tgy1, ix, iy) = st_lin(s1, tg);
      }
    }
    if (ineudif == 2)
      FOR2(iy, jlo, jhi, q_, 0, xw - 1) {
        XR xs = xrange(is, ie, js, js); xs.extra = IXP1(ie, iy);
        XRQ(ix, xs, q_) { A(pgy0, ix, iy) = st_log(stx(ix, iy, 0), pg); A(pgy1, ix, iy) = st_log(stx(ix, iy, 1), pg); }
      }
    FOR2(iy, js, je, q_, 0, xw - 1) {  // convert.m:736-765
      XR xs = xrange(is, ie, iy, iy);
      XRQ(ix, xs, q_) {
        int ix1 = IXP1(ix, iy);
        A(gpex, ix, iy) = (A(pre, ix1, iy) - A(pre, ix, iy)) * A(gxf, ix, iy);
        A(gtex, ix, iy) = (A(te, ix1, iy) - A(te, ix, iy)) * A(gxf, ix, iy);
        A(gtix, ix, iy) = (A(ti, ix1, iy) - A(ti, ix, iy)) * A(gxf, ix, iy);
        A(gprx, ix, iy) = A(gprx, ix, iy) + A(gpex, ix, iy);
        if (isphion + isphiofft == 1) A(ex, ix, iy) = (A(phi, ix, iy) - A(phi, ix1, iy)) * A(gxf, ix, iy);
      }
    }
    if (iysptrx < ny) FOR1(iy, js, je) { A(ex, ixlb, iy) = A(ex, ixlb + 1, iy); A(ex, ixrb, iy) = A(ex, ixrb - 1, iy); }
    FOR2(iy, jlo, jhi, q_, 0, xw - 1) {  // convert.m:768-786 (eymask1d = 1)
      XR xs = xrange(is, ie, js, js); xs.extra = IXP1(ie, iy);
      XRQ(ix, xs, q_) {
        A(gpey, ix, iy) = (A(ney1, ix, iy) * A(tey1, ix, iy) - A(ney0, ix, iy) * A(tey0, ix, iy)) / A(dynog, ix, iy);
        A(gtey, ix, iy) = (A(tey1, ix, iy) - A(tey0, ix, iy)) / A(dynog, ix, iy);
        A(gtiy, ix, iy) = (A(tiy1, ix, iy) - A(tiy0, ix, iy)) / A(dynog, ix, iy);
        A(ey, ix, iy) = -1. * (A(phiy1, ix, iy) - A(phiy0, ix, iy)) / A(dynog, ix, iy);
        A(gpry, ix, iy) = A(gpry, ix, iy) + A(gpey, ix, iy);
      }
    }
    // vertex values (convert.m:791-868)
    FOR2(iy, jlo, jhi, q_, 0, xw - 1) {
      XR xs = xrange(is, ie, iy, iy);
      XRQ(ix, xs, q_) {
        int ix1 = IXP1(ix, iy), ix2 = IXP1(ix, iy + 1);
        A(phiv, ix, iy) = 0.25 * (A(phi, ix, iy) + A(phi, ix1, iy) + A(phi, ix, iy + 1) + A(phi, ix2, iy + 1));
        A(tiv, ix, iy) = 0.25 * (A(ti, ix, iy) + A(ti, ix1, iy) + A(ti, ix, iy + 1) + A(ti, ix2, iy + 1));
        A(tev, ix, iy) = 0.25 * (A(te, ix, iy) + A(te, ix1, iy) + A(te, ix, iy + 1) + A(te, ix2, iy + 1));
        A(prev, ix, iy) = 0.25 * (A(pre, ix, iy) + A(pre, ix1, iy) + A(pre, ix, iy + 1) + A(pre, ix2, iy + 1));
        A(prtv, ix, iy) = A(prev, ix, iy);
      }
    }
    for (int f = 0; f < nisp; ++f)
      FOR2(iy, jlo, jhi, q_, 0, xw - 1) {
        XR xs = xrange(is, ie, iy, iy);
        XRQ(ix, xs, q_) {
          int ix1 = IXP1(ix, iy), ix2 = IXP1(ix, iy + 1);
          A(priv[f], ix, iy) = 0.25 * (A(pri[f], ix, iy) + A(pri[f], ix1, iy) + A(pri[f], ix, iy + 1) + A(pri[f], ix2, iy + 1));
          if (zi[f] != 0.) A(prtv, ix, iy) = A(prtv, ix, iy) + A(priv[f], ix, iy);
        }
      }
    SER {  // X-point vertex: 8-cell average (convert.m:831-868); nyomitmx = 0
      const int isx = ixpt1, jsx = iysptrx1, iex = ixpt2;
      if (!(isx < 0 || iex < 0 || iex > nx) && inrow(jsx) && inrow(jsx + 1)) {
        auto av8 = [&](const double* a) {
          return 0.125 * (A(a, isx, jsx) + A(a, isx + 1, jsx) + A(a, isx, jsx + 1) + A(a, isx + 1, jsx + 1) + A(a, iex, jsx) + A(a, iex + 1, jsx) + A(a, iex, jsx + 1) + A(a, iex + 1, jsx + 1));
        };
        A(phiv, isx, jsx) = av8(phi); A(phiv, iex, jsx) = A(phiv, isx, jsx);
        A(tiv, isx, jsx) = av8(ti); A(tiv, iex, jsx) = A(tiv, isx, jsx);
        A(tev, isx, jsx) = av8(te); A(tev, iex, jsx) = A(tev, isx, jsx);
        A(prev, isx, jsx) = av8(pre); A(prev, iex, jsx) = A(prev, isx, jsx);
        A(prtv, isx, jsx) = A(prev, isx, jsx);
        for (int f = 0; f < nisp; ++f) {
          A(priv[f], isx, jsx) = av8(pri[f]); A(priv[f], iex, jsx) = A(priv[f], isx, jsx);
          if (zi[f] != 0.) A(prtv, isx, jsx) = A(prtv, isx, jsx) + A(priv[f], isx, jsx);
        }
        A(prtv, iex, jsx) = A(prtv, isx, jsx);
      }
    }
  }

  // ---- fd2tra (oderhs.m:7-534): every scheme, orthogonal and non-orthogonal ------------------------------------------
  HD void fd2tra(const Win& w, const double* flx, const double* fly, const double* difx, const double* dify, const double* ph, double* trax, double* tray, int pos, int meth) {
    const int posx = pos % 10, posy = pos / 10, methx = iabs(meth % 10), methy = iabs(meth / 10);
    FOR2(iy, w.j4, w.j8, ix, w.i1, w.i5) {
        const int ix1 = IXP1(ix, iy);
        const int ix2 = ix * (1 - posx) + ix1 * posx;
        const double p0 = A(ph, ix, iy), p1 = A(ph, ix1, iy), fl = A(flx, ix2, iy), df = A(difx, ix2, iy);
        double t;
        switch (methx) {
          case 0: t = -df * (p1 - p0); break;
          case 1: t = upwind(fl, p0, p1); break;
          case 2: t = fl * (p1 + p0) / 2. - df * (p1 - p0); break;
          case 4: { double tpv = mx(df - fabs(fl) / 2., 0.); t = upwind(fl, p0, p1) - tpv * (p1 - p0); } break;
          case 5: { double tpv = df * powi(1 - fabs(fl) / mx(mx(10. * df, fabs(fl)), cutlo), 5); t = upwind(fl, p0, p1) - tpv * (p1 - p0); } break;
          default: t = upwind(fl, p0, p1) - df * (p1 - p0); break;  // 3, 6, 7
        }
        A(trax, ix2, iy) = t;
      }
    FOR2(iy, w.j1, w.j5 - posy, ix, w.i4, w.i8) {
        const double fl = A(fly, ix, iy + posy), df = A(dify, ix, iy + posy);
        double py0, py1;
        if (isnonog == 0) { py0 = A(ph, ix, iy); py1 = A(ph, ix, iy + 1); }
        else {
          const St5 s0 = stx(ix, iy, 0), s1 = stx(ix, iy, 1);
          if (methy == 6) { py0 = st_log(s0, ph); py1 = st_log(s1, ph); }
          else if (methy == 7) { py0 = st_inv(s0, ph); py1 = st_inv(s1, ph); }
          else { py0 = st_lin(s0, ph); py1 = st_lin(s1, ph); }  // (scheme 8's velocity stencil is not reachable from pandf)
        }
        double t;
        switch (methy) {
          case 0: t = -df * (py1 - py0); break;
          case 1: t = upwind(fl, py0, py1); break;
          case 2: t = fl * (py1 + py0) / 2. - df * (py1 - py0); break;
          case 4: { double tpv = mx(df - fabs(fl) / 2., 0.); t = upwind(fl, py0, py1) - tpv * (py1 - py0); } break;
          case 5: { double tpv = df * powi(1 - fabs(fl) / mx(mx(10. * df, fabs(fl)), cutlo), 5); t = upwind(fl, py0, py1) - tpv * (py1 - py0); } break;
          default: t = upwind(fl, py0, py1) - df * (py1 - py0); break;
        }
        A(tray, ix, iy + posy) = t;
      }
  }

  // ---- neudif (oderhs.m:5584-6057), ineudif = 1: the older diffusive-neutral model (ng and tg differenced separately);
  //      orthogonal meshes only here (the 2007 Forthon cases ran with it); stretcx = 1
  HD void neudif(const Win& w) {
    const int methgx = methg % 10, methgy = methg / 10;
    const double mg = mg_[0];
    FOR2(iy, w.j4, w.j8, ix, w.i1, w.i5) {
        const int ix2 = IXP1(ix, iy);
        double t0 = mx(A(tg, ix, iy), temin * ev), t1 = mx(A(tg, ix2, iy), temin * ev);
        double vtn = sqrt(t0 / mg), vtnp = sqrt(t1 / mg);
        double nu1 = A(nuix, ix, iy) + vtn / lgmax[0], nu2 = A(nuix, ix2, iy) + vtnp / lgmax[0];
        double qfl = flalfgxa[ix] * A(sx, ix, iy) * (vtn + vtnp) * rt8opi * (A(ng, ix, iy) * A(gx, ix, iy) + A(ng, ix2, iy) * A(gx, ix2, iy)) / (8 * (A(gx, ix, iy) + A(gx, ix2, iy)));
        double csh = (1 - isgasdc) * cdifg[0] * A(sx, ix, iy) * A(gxf, ix, iy) * ave(1. * (vtn * vtn) / nu1, 1. * (vtnp * vtnp) / nu2) + isgasdc * A(sx, ix, iy) * A(gxf, ix, iy) * difcng +
                     (rld2dxg[0] * rld2dxg[0]) * A(sx, ix, iy) * (1 / A(gxf, ix, iy)) * 0.5 * (A(nuiz, ix, iy) + A(nuiz, ix2, iy));
        double qtgf = cngfx_[0] * fgtdx[ix] * A(sx, ix, iy) * ave(1. * A(gx, ix, iy) / nu1, 1. * A(gx, ix2, iy) / nu2) * (vtn * vtn - vtnp * vtnp);
        A(vygtan, ix, iy) = 0.;
        qtgf = qtgf - A(vygtan, ix, iy) * A(sx, ix, iy);
        double nconv = 2.0 * (A(ng, ix, iy) * A(ng, ix2, iy)) / (A(ng, ix, iy) + A(ng, ix2, iy));
        if (methgx != 2) nconv = A(ng, ix, iy) * 0.5 * (1 + sgn(1., qtgf)) + A(ng, ix2, iy) * 0.5 * (1 - sgn(1., qtgf));
        double qsh = csh * (A(ng, ix, iy) - A(ng, ix2, iy)) + qtgf * nconv;
        double qr = fabs(qsh / qfl);
        if (ix == ixlb || ix == ixrb) { qr = gcfacgx * qr; qtgf = gcfacgx * qtgf; }
        A(conxg, ix, iy) = csh / ue_pow(1 + ue_pow(qr, flgamg), 1 / flgamg);
        if (isdifxg_aug == 1) A(conxg, ix, iy) = csh * (1 + qr);
        A(floxg, ix, iy) = qtgf / ue_pow(1 + ue_pow(qr, flgamg), 1 / flgamg);
        A(floxg, ix, iy) = A(floxg, ix, iy) + cngflox[0] * A(sx, ix, iy) * A(uu[0], ix, iy);
      }
    FOR1(iy, w.j4, w.j8) { A(conxg, nx + 1, iy) = 0; }
    FOR2(iy, w.j1, w.j5, ix, w.i4, w.i8) {
        double t0 = mx(A(tg, ix, iy), temin * ev), t1 = mx(A(tg, ix, iy + 1), temin * ev);
        double vtn = sqrt(t0 / mg), vtnp = sqrt(t1 / mg);
        double nu1 = A(nuix, ix, iy) + vtn / lgmax[0], nu2 = A(nuix, ix, iy + 1) + vtnp / lgmax[0];
        double qfl = flalfgya[iy] * A(sy, ix, iy) * (vtn + vtnp) * rt8opi * (A(ngy0, ix, iy) * A(gy, ix, iy) + A(ngy1, ix, iy) * A(gy, ix, iy + 1)) / (8 * (A(gy, ix, iy) + A(gy, ix, iy + 1)));
        double csh = (1 - isgasdc) * cdifg[0] * A(sy, ix, iy) / (A(dynog, ix, iy)) * ave((vtn * vtn) / nu1, (vtnp * vtnp) / nu2) + isgasdc * A(sy, ix, iy) * A(gyf, ix, iy) * difcng +
                     (rld2dyg[0] * rld2dyg[0]) * A(sy, ix, iy) * (1 / A(gyf, ix, iy)) * 0.5 * (A(nuiz, ix, iy) + A(nuiz, ix, iy + 1));
        double qtgf = cngfy_[0] * fgtdy[iy] * A(sy, ix, iy) * ave(A(gy, ix, iy) / nu1, A(gy, ix, iy + 1) / nu2) * (vtn * vtn - vtnp * vtnp);
        double nconv = 2.0 * (A(ngy0, ix, iy) * A(ngy1, ix, iy)) / (A(ngy0, ix, iy) + A(ngy1, ix, iy));
        if (methgy != 2) nconv = A(ngy0, ix, iy) * 0.5 * (1 + sgn(1., qtgf)) + A(ngy1, ix, iy) * 0.5 * (1 - sgn(1., qtgf));
        double qsh = csh * (A(ngy0, ix, iy) - A(ngy1, ix, iy)) + qtgf * nconv;
        double qr = fabs(qsh / qfl);
        if (iy == 0) { qr = gcfacgy * qr; qtgf = gcfacgy * qtgf; }
        if (iy == ny) { qr = gcfacgy * qr; qtgf = gcfacgy * qtgf; }
        A(conyg, ix, iy) = csh / ue_pow(1 + ue_pow(qr, flgamg), 1 / flgamg);
        if (isdifyg_aug == 1) A(conyg, ix, iy) = csh * (1 + qr);
        A(floyg, ix, iy) = qtgf / ue_pow(1 + ue_pow(qr, flgamg), 1 / flgamg);
        A(floyg, ix, iy) = A(floyg, ix, iy) + cngfloy[0] * A(sy, ix, iy) * A(vy[0], ix, iy);
      }
    fd2tra(w, floxg, floyg, conxg, conyg, ng, fngx, fngy, 0, methg);
    FOR2(iy, w.j1, w.j5, ix, w.i1, w.i5) {
        const int ix1 = IXP1(ix, iy);
        A(uug, ix, iy) = A(fngx, ix, iy) / (0.5 * (A(ng, ix, iy) + A(ng, ix1, iy)) * A(sx, ix, iy));
        A(vyg, ix, iy) = A(fngy, ix, iy) / (0.5 * (A(ng, ix, iy) + A(ng, ix, iy + 1)) * A(sy, ix, iy));
      }
    FOR2(iy, w.j2, w.j5, ix, w.i2, w.i5) {
        const int ix1 = IXM1(ix, iy);
        A(resng, ix, iy) = cngsor * (A(psorg, ix, iy) + 0. + A(psorrg, ix, iy)) + 0. - A(fngx, ix, iy) + A(fngx, ix1, iy) - fluxfacy * (A(fngy, ix, iy) - A(fngy, ix, iy - 1)) + 0. * A(vol, ix, iy);
      }
  }

  // ---- neudifpg (oderhs.m:6058-6648), gas species 1 -------------------------------------------------------------------
  HD void neudifpg(const Win& w) {
    const int methgx = methg % 10, methgy = methg / 10;
    const double mg = mg_[0], ngb = ngbackg_[0];
    FOR2(iy, w.j4, w.j8, ix, w.i1, w.i5) {
        const int ix2 = IXP1(ix, iy);
        double ngxface = 0.5 * (A(ng, ix, iy) + A(ng, ix2, iy));
        double t0 = mx(A(tg, ix, iy), temin * ev), t1 = mx(A(tg, ix2, iy), temin * ev);
        double vtn = sqrt(t0 / mg), vtnp = sqrt(t1 / mg);
        double nu1 = A(nuix, ix, iy) + vtn / lgmax[0], nu2 = A(nuix, ix2, iy) + vtnp / lgmax[0];
        double tgf = 0.5 * (A(tg, ix, iy) + A(tg, ix2, iy));
        double flalfgx_adj = flalfgxa[ix] * (1. + powi(cflbg * ngb / ngxface, inflbg));
        double qfl = flalfgx_adj * A(sx, ix, iy) * (vtn + vtnp) * rt8opi * (A(ng, ix, iy) * A(gx, ix, iy) + A(ng, ix2, iy) * A(gx, ix2, iy)) / (8 * (A(gx, ix, iy) + A(gx, ix2, iy)));
        double csh = (1 - isgasdc) * cdifg[0] * A(sx, ix, iy) * A(gxf, ix, iy) * (1 / mg) * ave(1. / nu1, 1. / nu2) + isgasdc * A(sx, ix, iy) * A(gxf, ix, iy) * difcng / tgf +
                     (rld2dxg[0] * rld2dxg[0]) * A(sx, ix, iy) * (1 / A(gxf, ix, iy)) * 0.5 * (A(nuiz, ix, iy) + A(nuiz, ix2, iy)) / tgf;
        double qtgf = alftng * fgtdx[ix] * A(sx, ix, iy) * ave(A(gx, ix, iy) / nu1, A(gx, ix2, iy) / nu2) * (vtn * vtn - vtnp * vtnp);
        if (isupgon == 1) { csh = csh * (1 - A(rrv, ix, iy) * A(rrv, ix, iy)); qtgf = qtgf * (1 - A(rrv, ix, iy) * A(rrv, ix, iy)); }
        A(vygtan, ix, iy) = 0.;
        if (isnonog == 1 && iy <= ny) {
          double grdnv = grdnv_y(tg, ix, iy, 1) / A(dxnog, ix, iy);
          A(vygtan, ix, iy) = ue_exp(0.5 * (ue_log(A(tg, ix2, iy)) + ue_log(A(tg, ix, iy)))) * (alftng / (mg * 0.5 * (nu1 + nu2))) *
                              (grdnv / ue_cos(A(angfx, ix, iy)) - (ue_log(A(tg, ix2, iy)) - ue_log(A(tg, ix, iy))) * A(gxf, ix, iy));
        }
        qtgf = qtgf - A(vygtan, ix, iy) * A(sx, ix, iy);
        if (isupgon == 1) qtgf = qtgf + A(rrv, ix, iy) * A(up[iigsp], ix, iy) * A(sx, ix, iy);
        double nconv = 2.0 * (A(ng, ix, iy) * A(ng, ix2, iy)) / (A(ng, ix, iy) + A(ng, ix2, iy));
        if (methgx != 2) nconv = A(ng, ix, iy) * 0.5 * (1 + sgn(1., qtgf)) + A(ng, ix2, iy) * 0.5 * (1 - sgn(1., qtgf));
        double qsh = csh * (A(pg, ix, iy) - A(pg, ix2, iy)) + qtgf * nconv;
        double qr = fabs(qsh / qfl);
        if (ix == ixlb || ix == ixrb) { qr = gcfacgx * qr; qtgf = gcfacgx * qtgf; }
        A(conxg, ix, iy) = csh / ue_pow(1 + ue_pow(qr, flgamg), 1 / flgamg);
        if (isdifxg_aug == 1) A(conxg, ix, iy) = csh * (1 + qr);
        A(floxg, ix, iy) = (qtgf / tgf) / ue_pow(1 + ue_pow(qr, flgamg), 1 / flgamg);
        A(floxg, ix, iy) = A(floxg, ix, iy) + cngflox[0] * A(sx, ix, iy) * A(uu[0], ix, iy) / tgf;
      }
    FOR1(iy, w.j4, w.j8) { A(conxg, nx + 1, iy) = 0; }
    FOR2(iy, w.j1, w.j5, ix, w.i4, w.i8) {
        double ngyface = 0.5 * (A(ng, ix, iy) + A(ng, ix, iy + 1));
        double t0 = mx(A(tg, ix, iy), tgmin * ev), t1 = mx(A(tg, ix, iy + 1), tgmin * ev);
        double vtn = sqrt(t0 / mg), vtnp = sqrt(t1 / mg);
        double nu1 = A(nuix, ix, iy) + vtn / lgmax[0], nu2 = A(nuix, ix, iy + 1) + vtnp / lgmax[0];
        double tgf = 0.5 * (A(tg, ix, iy) + A(tg, ix, iy + 1));
        double flalfgy_adj = flalfgya[iy] * (1. + powi(cflbg * ngb / ngyface, inflbg));
        double qfl = flalfgy_adj * A(sy, ix, iy) * (vtn + vtnp) * rt8opi * (A(ngy0, ix, iy) * A(gy, ix, iy) + A(ngy1, ix, iy) * A(gy, ix, iy + 1)) / (8 * (A(gy, ix, iy) + A(gy, ix, iy + 1)));
        if (iy == 0) qfl = flalfgy_adj * A(sy, ix, iy) * (vtn + vtnp) * rt8opi * (A(ngy0, ix, iy) + A(ngy1, ix, iy)) / 8.;
        double csh = (1 - isgasdc) * (cdifg[0] * A(sy, ix, iy) / A(dynog, ix, iy)) * (1 / mg) * ave(1. / nu1, 1. / nu2) + isgasdc * A(sy, ix, iy) * difcng / (A(dynog, ix, iy) * tgf) +
                     (rld2dyg[0] * rld2dyg[0]) * A(sy, ix, iy) * A(dynog, ix, iy) * 0.5 * (A(nuiz, ix, iy) + A(nuiz, ix, iy + 1)) / tgf;
        double qtgf = alftng * fgtdy[iy] * A(sy, ix, iy) * ave(A(gy, ix, iy) / nu1, A(gy, ix, iy + 1) / nu2) * (vtn * vtn - vtnp * vtnp);
        if (isnonog == 1 && iy <= ny) {
          const St5 s0 = stx(ix, iy, 0), s1 = stx(ix, iy, 1);
          double ty0 = isintlog == 0 ? st_lin(s0, tg) : st_log(s0, tg), ty1 = isintlog == 0 ? st_lin(s1, tg) : st_log(s1, tg);
          qtgf = alftng * fgtdy[iy] * A(sy, ix, iy) * ave(A(gy, ix, iy) / nu1, A(gy, ix, iy + 1) / nu2) * (ty0 - ty1) / mg;
        }
        double nconv = 2.0 * (A(ngy0, ix, iy) * A(ngy1, ix, iy)) / (A(ngy0, ix, iy) + A(ngy1, ix, iy));
        if (methgy != 2) nconv = A(ngy0, ix, iy) * 0.5 * (1 + sgn(1., qtgf)) + A(ngy1, ix, iy) * 0.5 * (1 - sgn(1., qtgf));
        double qsh = csh * (A(pgy0, ix, iy) - A(pgy1, ix, iy)) + qtgf * nconv;
        double qr = fabs(qsh / qfl);
        if (iy == 0) { qr = gcfacgy * qr; qtgf = gcfacgy * qtgf; }
        if (iy == ny) { qr = gcfacgy * qr; qtgf = gcfacgy * qtgf; }
        A(conyg, ix, iy) = csh / ue_pow(1 + ue_pow(qr, flgamg), 1 / flgamg);
        if (isdifyg_aug == 1) A(conyg, ix, iy) = csh * (1 + qr);
        A(floyg, ix, iy) = (qtgf / tgf) / ue_pow(1 + ue_pow(qr, flgamg), 1 / flgamg);
        A(floyg, ix, iy) = A(floyg, ix, iy) + cngfloy[0] * A(sy, ix, iy) * A(vy[0], ix, iy) / tgf;
      }
    fd2tra(w, floxg, floyg, conxg, conyg, pg, fngx, fngy, 0, methg);
    if (isnonog == 1) {  // oderhs.m:6345-6466
      FOR2(iy, w.j1, mn(w.j6, ny), ix, w.i1, mn(w.i6, nx)) {
          const int ix2 = IXP1(ix, iy);
          double t0 = mx(A(tg, ix, iy), temin * ev), t1 = mx(A(tg, ix2, iy), temin * ev);
          double vtn = sqrt(t0 / mg), vtnp = sqrt(t1 / mg);
          double nu1 = A(nuix, ix, iy) + vtn / lgmax[0], nu2 = A(nuix, ix2, iy) + vtnp / lgmax[0];
          const bool isxyfl = !(ix == ixlb || ix == ixrb);
          const int mode = methgx == 6 ? 1 : (methgx == 7 ? 2 : 0);
          double grdnv = grdnv_y(pg, ix, iy, mode) / A(dxnog, ix, iy);
          double difgx2 = ave(1. / nu1, 1. / nu2) / mg + (rld2dxg[0] * rld2dxg[0]) * (1 / (A(gxf, ix, iy) * A(gxf, ix, iy))) * 0.5 * (A(nuiz, ix, iy) + A(nuiz, ix2, iy));
          if (methgx == 6)
            A(fngxy, ix, iy) = ue_exp(0.5 * (ue_log(A(pg, ix2, iy)) + ue_log(A(pg, ix, iy)))) * difgx2 *
                               (grdnv / ue_cos(A(angfx, ix, iy)) - (ue_log(A(pg, ix2, iy)) - ue_log(A(pg, ix, iy))) * A(gxf, ix, iy)) * A(sx, ix, iy);
          else
            A(fngxy, ix, iy) = difgx2 * (grdnv / ue_cos(A(angfx, ix, iy)) - (A(pg, ix2, iy) - A(pg, ix, iy)) * A(gxf, ix, iy)) * A(sx, ix, iy);
          double ngxface = 0.5 * (A(ng, ix, iy) + A(ng, ix2, iy));
          double flalfgxy_adj = flalfgxya[ix] * (1. + powi(cflbg * ngb / ngxface, inflbg));
          double qfl = flalfgxy_adj * A(sx, ix, iy) * (vtn + vtnp) * rt8opi * (A(ng, ix, iy) * A(gx, ix, iy) + A(ng, ix2, iy) * A(gx, ix2, iy)) / (8 * (A(gx, ix, iy) + A(gx, ix2, iy)));
          if (isxyfl) A(fngxy, ix, iy) = A(fngxy, ix, iy) / sqrt(1 + sq(A(fngxy, ix, iy) / qfl));
        }
      FOR2(iy, w.j4, w.j8, ix, w.i1, w.i5) {
          const int ix2 = IXP1(ix, iy);
          A(fngx, ix, iy) = A(fngx, ix, iy) - A(fngxy, ix, iy);
          double t0 = mx(A(tg, ix, iy), temin * ev), t1 = mx(A(tg, ix2, iy), temin * ev);
          double vtn = sqrt(t0 / mg), vtnp = sqrt(t1 / mg);
          double qfl = flalfgnx * A(sx, ix, iy) * (vtn + vtnp) * rt8opi * (A(ng, ix, iy) + A(ng, ix2, iy)) / 16;
          A(fngx, ix, iy) = A(fngx, ix, iy) / sqrt(1 + sq(A(fngx, ix, iy) / qfl));
          A(fngx, ix, iy) = A(fngx, ix, iy) / (1 - 2 * nlimgx + nlimgx * (A(ng, ix2, iy) / A(ng, ix, iy) + A(ng, ix, iy) / A(ng, ix2, iy)));
        }
      FOR2(iy, w.j1, w.j5, ix, w.i4, w.i8) {
          A(fngy, ix, iy) = A(fngy, ix, iy) / (1 - 2 * nlimgy + nlimgy * (A(ng, ix, iy + 1) / A(ng, ix, iy) + A(ng, ix, iy) / A(ng, ix, iy + 1)));
          double t0 = mx(A(tg, ix, iy), temin * ev), t1 = mx(A(tg, ix, iy + 1), temin * ev);
          double vtn = sqrt(t0 / mg), vtnp = sqrt(t1 / mg);
          double qfl = flalfgny * A(sy, ix, iy) * (vtn + vtnp) * rt8opi * (A(ngy0, ix, iy) + A(ngy1, ix, iy)) / 16;
          A(fngy, ix, iy) = A(fngy, ix, iy) / sqrt(1 + sq(A(fngy, ix, iy) / qfl));
        }
    }
    // neutral flow velocities (oderhs.m:6512-6554)
    FOR2(iy, w.j1, w.j5, ix, w.i1, w.i5) {
        const int ix1 = IXP1(ix, iy);
        if (1. - A(rrv, ix, iy) > 1.e-4 || isupgon == 0) {
          A(uug, ix, iy) = A(fngx, ix, iy) / (0.5 * (A(ng, ix, iy) + A(ng, ix1, iy)) * A(sx, ix, iy));
          A(uuxg, ix, iy) = (A(fngx, ix, iy) + A(fngxy, ix, iy)) / (0.5 * (A(ng, ix, iy) + A(ng, ix1, iy)) * A(sx, ix, iy));
        } else A(uug, ix, iy) = A(up[iigsp], ix, iy);
        A(vyg, ix, iy) = A(fngy, ix, iy) / (0.5 * (A(ng, ix, iy) + A(ng, ix, iy + 1)) * A(sy, ix, iy));
        if (isupgon == 1) A(vy[iigsp], ix, iy) = A(vyg, ix, iy);
      }
    FOR1(iy, w.j1, w.j5) { if (iy <= iysptrx2 && isfixlb == 2) A(uug, ixpt2, iy) = 0; }
    if (isupgon == 1)
      FOR2(iy, w.j4, w.j6, ix, w.i1, w.i6) {
          A(uu[iigsp], ix, iy) = A(uug, ix, iy);
          A(v2[iigsp], ix, iy) = (A(uuxg, ix, iy) - A(up[iigsp], ix, iy) * A(rrv, ix, iy)) / (A(rbfbt, ix, iy) + A(rbfbt, IXP1(ix, iy), iy)) * 2.;
        }
    if (isupgon == 0)
      FOR2(iy, w.j2, w.j5, ix, w.i2, w.i5) {  // oderhs.m:6574-6610 (psorcxg, volpsorg, psgov_use are zero fields)
          const int ix1 = IXM1(ix, iy);
          A(resng, ix, iy) = cngsor * (A(psorg, ix, iy) + 0. + A(psorrg, ix, iy)) + 0. + 0. * A(vol, ix, iy);
          A(resng, ix, iy) = A(resng, ix, iy) - cfneutdiv * cfneutdiv_fng * ((A(fngx, ix, iy) - A(fngx, ix1, iy)) + fluxfacy * (A(fngy, ix, iy) - A(fngy, ix, iy - 1)));
        }
  }


  // ---- engbalg (oderhs.m:7508-7878): the gas energy equation, gas species 1 = the inertial atoms (ngsp = 1, nisp = 2) -------------
  HD void engbalg(const Win& w) {
    const int i1 = w.i1, i2 = w.i2, i4 = w.i4, i5 = w.i5, i6 = w.i6, i8 = w.i8;
    const int j1 = w.j1, j2 = w.j2, j4 = w.j4, j5 = w.j5, j6 = w.j6, j8 = w.j8;
    FOR2(iy, j2, j5, ix, i2, i5) A(segc, ix, iy) = 0.0;
    if (istgon == 1)  // v.grad(pg) work (oderhs.m:7564-7585)
      FOR2(iy, j2, j5, ix, i2, i5) {
          const int ix1 = IXM1(ix, iy), ix2 = IXP1(ix, iy), iy1 = mx(0, iy - 1);
          const double tv = (A(pg, ix2, iy) - A(pg, ix, iy)), t1 = (A(pg, ix, iy) - A(pg, ix1, iy));
          A(segc, ix, iy) = 0.5 * cvgpg * (A(uuxg, ix, iy) * ave(A(gx, ix2, iy), A(gx, ix, iy)) * tv + A(uuxg, ix1, iy) * ave(A(gx, ix, iy), A(gx, ix1, iy)) * t1) * A(vol, ix, iy);
          const double t2 = cvgpg * 0.5 * (A(vyg, ix, iy) * A(dynog, ix, iy) * (A(pgy1, ix, iy) - A(pgy0, ix, iy)) + A(vyg, ix, iy1) * A(dynog, ix, iy1) * (A(pgy1, ix, iy1) - A(pgy0, ix, iy1)));
          A(segc, ix, iy) = A(segc, ix, iy) + cvgpg * t2 * A(vol, ix, iy);
        }
    FOR2(iy, j1, j6, ix, i1, i6) { A(floxge, ix, iy) = 0.0; A(floyge, ix, iy) = 0.0; A(conxge, ix, iy) = 0.0; A(conyge, ix, iy) = 0.0; }
    FOR2(iy, j4, j8, ix, i1, i5) A(conxge, ix, iy) = A(sx, ix, iy) * A(hcxg, ix, iy) * A(gxf, ix, iy);  // conduction (oderhs.m:7605-7636); hcxg is flux-limited already
    FOR1(iy, j4, j8) { A(conxge, nx + 1, iy) = 0; }
    FOR2(iy, j1, j5, ix, i4, i8) A(conyge, ix, iy) = A(sy, ix, iy) * A(hcyg, ix, iy) / A(dynog, ix, iy);
    if (inrow(ny + 1)) FOR1(ix, i1, i6) A(conyge, ix, ny + 1) = 0.0;
    FOR2(iy, j4, j8, ix, i1, i5) A(floxge, ix, iy) = cfcvtg * 2.5 * A(fngx, ix, iy);  // convection (oderhs.m:7643-7705)
    FOR1(iy, j4, j8) { A(floxge, nx + 1, iy) = 0.; }
    FOR1(iy, j4, j8) {  // no inward power from the plates
      if (A(fngx, ixlb, iy) > 0.) A(floxge, ixlb, iy) = A(floxge, ixlb, iy) - (1. - cfloxiplt) * cfcvti * 2.5 * A(fngx, ixlb, iy);
      if (A(fngx, ixrb, iy) < 0.) A(floxge, ixrb, iy) = A(floxge, ixrb, iy) - (1. - cfloxiplt) * cfcvti * 2.5 * A(fngx, ixrb, iy);
      A(floxge, ixrb + 1, iy) = 0.0;
    }
    FOR2(iy, j1, j5, ix, i4, i8) A(floyge, ix, iy) = cfcvtg * 2.5 * A(fngy, ix, iy);
    FOR1(ix, i4, i8) {  // ... nor from the walls
      if (inrow(0) && (ix <= ixpt1 || ix > ixpt2)) { if (A(fngy, ix, 0) > 0.) A(floyge, ix, 0) = A(floyge, ix, 0) - (1. - cfloygwall) * cfcvtg * 2.5 * A(fngy, ix, 0); }
      if (inrow(ny) && A(fngy, ix, ny) < 0.) A(floyge, ix, ny) = A(floyge, ix, ny) - (1. - cfloygwall) * cfcvtg * 2.5 * A(fngy, ix, ny);
      if (inrow(ny + 1)) A(floyge, ix, ny + 1) = 0.0;
    }
    if (istgon == 1) fd2tra(w, floxge, floyge, conxge, conyge, tg, fegx, fegy, 0, methi);  // oderhs.m:7708-7716
    if (isnonog == 1 && istgon == 1)  // y-component of the non-orthogonal diffusive flux (oderhs.m:7720-7782)
      FOR2(iy, j1, mn(j6, ny), ix, i1, i6) {
          const int iy1 = mx(iy - 1, 0);
          const int ix2 = IXP1(ix, iy), ix4 = IXP1(ix, iy1), ix6 = IXP1(ix, iy + 1);
          double t0 = mx(A(tg, ix, iy), tgmin * ev), t1 = mx(A(tg, ix2, iy), tgmin * ev);
          const double vtn = sqrt(t0 / mg_[0]), vtnp = sqrt(t1 / mg_[0]);
          const double nu1 = A(nuix, ix, iy) + vtn / lgmax[0], nu2 = A(nuix, ix2, iy) + vtnp / lgmax[0];
          const double grdnv = ((A(fym[1], ix, iy) * ue_log(A(tg, ix2, iy1)) + A(fy0[1], ix, iy) * ue_log(A(tg, ix2, iy)) + A(fyp[1], ix, iy) * ue_log(A(tg, ix2, iy + 1)) +
                                 A(fymx[1], ix, iy) * ue_log(A(tg, ix, iy1)) + A(fypx[1], ix, iy) * ue_log(A(tg, ix, iy + 1))) -
                                (A(fym[0], ix, iy) * ue_log(A(tg, ix, iy1)) + A(fy0[0], ix, iy) * ue_log(A(tg, ix, iy)) + A(fyp[0], ix, iy) * ue_log(A(tg, ix, iy + 1)) +
                                 A(fymx[0], ix, iy) * ue_log(A(tg, ix4, iy1)) + A(fypx[0], ix, iy) * ue_log(A(tg, ix6, iy + 1)))) / A(dxnog, ix, iy);
          const double difgx2 = ave(A(tg, ix, iy) / nu1, A(tg, ix2, iy) / nu2) / mg_[0] + sq(rld2dxg[0]) * (1 / sq(A(gxf, ix, iy))) * 0.5 * (A(nuiz, ix, iy) + A(nuiz, ix2, iy));
          A(fegxy, ix, iy) = cfegxy * ue_exp(0.5 * (ue_log(A(tg, ix2, iy)) + ue_log(A(tg, ix, iy)))) * difgx2 * ave(A(ng, ix2, iy), A(ng, ix, iy)) *
                             (grdnv / ue_cos(A(angfx, ix, iy)) - (ue_log(A(tg, ix2, iy)) - ue_log(A(tg, ix, iy))) * A(gxf, ix, iy)) * A(sx, ix, iy);
          t0 = mx(A(tg, ix, iy), tgmin * ev); t1 = mx(A(tg, ix2, iy), tgmin * ev);
          const double vttn = t0 * sqrt(t0 / mg_[0]), vttp = t1 * sqrt(t1 / mg_[0]);
          double qfl;
          if (isfegxyqflave == 0) qfl = flalftgxy * 0.25 * A(sx, ix, iy) * (vttn + vttp) * (A(ng, ix, iy) + A(ng, ix2, iy));
          else qfl = flalftgxy * A(sx, ix, iy) * ave(vttn, vttp) * ave(A(ng, ix, iy), A(ng, ix2, iy));
          A(fegxy, ix, iy) = A(fegxy, ix, iy) / sqrt(1. + sq(A(fegxy, ix, iy) / qfl));
          A(fegx, ix, iy) = A(fegx, ix, iy) - A(fegxy, ix, iy);
        }
    FOR2(iy, j2, j5, ix, i2, i5) {  // residual and equipartition with the ions (oderhs.m:7790-7806)
        const int iy1 = mx(0, iy - 1);
        const int ix1 = IXM1(ix, iy);
        A(reseg, ix, iy) = -(A(fegx, ix, iy) - A(fegx, ix1, iy) + A(fegy, ix, iy) - A(fegy, ix, iy1)) + A(segc, ix, iy);
        A(reseg, ix, iy) = A(reseg, ix, iy) + A(vol, ix, iy) * A(eqpg, ix, iy) * (A(ti, ix, iy) - A(tg, ix, iy));
        A(seic, ix, iy) = A(seic, ix, iy) - A(vol, ix, iy) * (1.0 - cftiexclg) * A(eqpg, ix, iy) * (A(ti, ix, iy) - A(tg, ix, iy));
      }
  }

  // ---- pandf (oderhs.m:537-5070) -------------------------------------------------------------------------------------
  HD int pandf(int xc, int yc, const double* yl, double* yldot) {
    const Win w = make_win(xc, yc);
    int rc = convsr_vo(xc, yc, yl);
    if (rc) return rc;
    convsr_aux(xc, yc);
    const int i1 = w.i1, i2 = w.i2, i4 = w.i4, i5 = w.i5, i6 = w.i6, i8 = w.i8;
    const int j1 = w.j1, j2 = w.j2, j4 = w.j4, j5 = w.j5, j6 = w.j6, j8 = w.j8;
    nfsp = nisp;
    // Coulomb logarithm on x-faces (oderhs.m:1138-1155)
    FOR2(iy, j1, j6, ix, i1, i6) {
        int ix1 = IXP1(ix, iy);
        double teev = 0.5 * (A(te, ix, iy) + A(te, ix1, iy)) / ev;
        double nexface = 0.5 * (A(ne, ix, iy) + A(ne, ix1, iy));
        if (islnlamcon == 1) A(loglambda, ix, iy) = lnlam;
        else if (teev < 50.) A(loglambda, ix, iy) = 23.4 - 1.15 * ue_log10(1.e-6 * nexface) + 3.45 * ue_log10(teev);
        else A(loglambda, ix, iy) = 25.3 - 1.15 * ue_log10(1.e-6 * nexface) + 2.33167537087122e+00 * ue_log10(teev);
      }
    // radial and "2" velocities of the ion species: diffusive parts plus the ExB and grad-B / curvature drifts (oderhs.m:1167-1471);
    // with the diamagnetic (cfydd, cf2dd) and resistive (cfrd) parts; the classical ones (cfvycf, cfvycr) are refused in init; bfacx/yrozh = 1
    for (int f = 0; f < nfsp; ++f) {
      if (!(zi[f] > 1.e-10)) continue;
      const double qion = zi[f] * qe;
      FOR2(iy, j1, j5, ix, i1, i6) {
          const int iyp1 = mn(iy + 1, ny + 1);
          const int ix3 = IXM1(ix, iy), ix4 = IXM1(ix, iy + 1);
          const double temp1 = (-4.0) * (A(phiv, ix, iy) - A(phiv, ix3, iy)) * A(gxc, ix, iy);
          const double temp2 = 4.0 * (A(priv[f], ix, iy) - A(priv[f], ix3, iy)) * A(gxc, ix, iy);
          const double lambd_ci = 1e16 * sq(A(ti, ix, iy) / ev) / A(nit, ix, iy), lambd_ce = 2e16 * sq(A(te, ix, iy) / ev) / A(ne, ix, iy);
          const double coll_fi = cfnus_i / (cfnus_i + (lambd_ci / (A(lconi, ix, iy)))), coll_fe = cfnus_e / (cfnus_e + (lambd_ce / (A(lcone, ix, iy))));
          A(vyce[f], ix, iy) = 0.125 * temp1 * (A(rbfbt2, ix, iy) + A(rbfbt2, ix, iy + 1));
          A(vycb[f], ix, iy) = (cfcurv * (0.5 * (A(ti, ix, iy) + A(ti, ix, iyp1)) + mi[f] * sq(0.25 * (A(up[f], ix, iy) + A(up[f], ix, iyp1) + A(up[f], ix3, iy) + A(up[f], ix4, iyp1)))) * A(curvrby, ix, iy) / qion +
                                cfgradb * 0.5 * (A(ti, ix, iy) + A(ti, ix, iyp1)) * A(gradby, ix, iy) / qion) * coll_fi;
          A(veycb, ix, iy) = (-cfcurv * 0.5 * (A(te, ix, iy) + A(te, ix, iyp1)) * A(curvrby, ix, iy) / qe - cfgradb * 0.5 * (A(te, ix, iy) + A(te, ix, iyp1)) * A(gradby, ix, iy) / qe) * coll_fe;
          // (the reference zeroes vycp on the faces iy = 0 and ny from every iteration: written here by the iteration that owns the face)
          A(vycp[f], ix, iy) = (iy == 0 || iy == ny) ? 0. : -0.25 * temp2 * (A(rbfbt2, ix, iy) + A(rbfbt2, ix, iy + 1)) / (qion * (A(niy0[f], ix, iy) + A(niy1[f], ix, iy)));
          A(vydd[f], ix, iy) = vcony[f] + 0. + 0. - (difpr[f] + 0.) * (2 * A(gpry, ix, iy) / (A(pr, ix, iy + 1) + A(pr, ix, iy)) - 3.0 * A(gtey, ix, iy) / (A(tey1, ix, iy) + A(tey0, ix, iy)));
          A(diffusivwrk, ix, iy) = fcdif * difni[f] + 0.;
          if (cfrtaue != 0.)  // classical momentum-transfer and viscosity velocities (oderhs.m:1240-1262)
            A(vycr, ix, iy) = -0.5 * (rtaue_(ix, iy) + rtaue_(ix, iyp1)) * ((A(gpiy[0], ix, iy) + A(gpey, ix, iy)) / (0.5 * (A(niy1[0], ix, iy) + A(niy0[0], ix, iy))) - 1.5 * A(gtey, ix, iy));
          if (cfeta1 != 0. && iy <= ny - 1 && iy > 0) {
            const int iym1 = mx(iy - 1, 0);
            const double geyym = 2 * A(gpiy[0], ix, iym1) / (A(ney1, ix, iym1) + A(ney0, ix, iym1)) - qe * A(ey, ix, iym1);
            const double geyy0 = 2 * A(gpiy[0], ix, iy) / (A(ney1, ix, iy) + A(ney0, ix, iy)) - qe * A(ey, ix, iy);
            const double geyyp = 2 * A(gpiy[0], ix, iyp1) / (A(ney1, ix, iyp1) + A(ney0, ix, iyp1)) - qe * A(ey, ix, iyp1);
            const double dgeyy0 = (geyy0 - geyym) * eta1_(ix, iy) * A(gy, ix, iy), dgeyy1 = (geyyp - geyy0) * eta1_(ix, iyp1) * A(gy, ix, iyp1);
            A(vycf, ix, iy) = 2 * (dgeyy1 - dgeyy0) * A(gy, ix, iy) / ((A(ney1, ix, iy) + A(ney0, ix, iy)) * sq(qe * 0.5 * (A(btot, ix, iy) + A(btot, ix, iym1))));
          }
        }
      FOR2(iy, j1, j5, ix, i1, i6) {
          double difnimix = A(diffusivwrk, ix, iy);
          A(vydd[f], ix, iy) = A(vydd[f], ix, iy) - 1. * difnimix * (2 * (1 - isvylog) * ((A(niy1[f], ix, iy) - A(niy0[f], ix, iy)) / A(dynog, ix, iy)) / (A(niy1[f], ix, iy) + A(niy0[f], ix, iy)) +
                                                                      isvylog * (ue_log(A(niy1[f], ix, iy)) - ue_log(A(niy0[f], ix, iy))) / A(dynog, ix, iy));
          const double vyrd = cfrd == 0. ? 0. : -2. * A(gpry, ix, iy) / (sq(A(btot, ix, iy)) / etaper(ix, iy) + sq(A(btot, ix, iy + 1)) / etaper(ix, iy + 1));  // (not evaluated when switched off)
          A(vy[f], ix, iy) = cfydd * A(vycp[f], ix, iy) + cfrd * vyrd + A(vydd[f], ix, iy) + cfyef * A(vyce[f], ix, iy) + cfybf * A(vycb[f], ix, iy) + cfvycf * A(vycf, ix, iy) + cfvycr * A(vycr, ix, iy);
          A(vygp[f], ix, iy) = (cfydd + cfybf) * A(vycp[f], ix, iy) + cfrd * vyrd + A(vydd[f], ix, iy) + cfyef * A(vyce[f], ix, iy) + cfvycf * A(vycf, ix, iy) + cfvycr * A(vycr, ix, iy);
          if (isybdrywd == 1 && ((iy == 0 && matwalli[ix] > 0) || (iy == ny && matwallo[ix] > 0))) A(vy[f], ix, iy) = A(vydd[f], ix, iy);  // diffusive in wall cells (oderhs.m:1312-1318)
        }
      FOR2(iy, j1, j6, ix, i1, i6) {
          const int ix2 = IXP1(ix, iy), iy1 = mx(0, iy - 1);
          const double temp1 = (-4.) * (A(phiv, ix, iy) - A(phiv, ix, iy1)) * A(gyc, ix, iy);
          const double temp2 = 4. * (A(priv[f], ix, iy) - A(priv[f], ix, iy1)) * A(gyc, ix, iy);
          A(v2ce[f], ix, iy) = -0.5 * temp1 / (A(btot, ix, iy) + A(btot, ix2, iy));
          A(v2cb[f], ix, iy) = (cfcurv * (0.5 * (A(tiv, ix, iy) + A(tiv, ix, iy1)) + mi[f] * sq(A(up[f], ix, iy))) * A(curvrb2, ix, iy) + cfgradb * 0.5 * (A(tiv, ix, iy) + A(tiv, ix, iy1)) * A(gradb2, ix, iy)) / qion;
          A(ve2cb, ix, iy) = -(cfcurv * 0.5 * (A(tev, ix, iy) + A(tev, ix, iy1)) * A(curvrb2, ix, iy) + cfgradb * 0.5 * (A(tev, ix, iy) + A(tev, ix, iy1)) * A(gradb2, ix, iy)) / qe;
          const double v2cd = temp2 / ((A(btot, ix, iy) + A(btot, ix2, iy)) * qion * (A(ni[f], ix, iy) + A(ni[f], ix2, iy)));
          A(this->v2cd[f], ix, iy) = v2cd;
          // plate electron diamagnetic flux for the sheath potential (oderhs.m:1376-1391)
          if (ix == ixlb) {
            const double v2dia = -0.5 * (A(gpey, ixlb + 1, iy) + A(gpey, ixlb + 1, iy1)) / (A(btot, ixlb + 1, iy) * qe * A(ne, ixlb + 1, iy));
            fdiaxlb[iy] = A(ne, ixlb + 1, iy) * A(sx, ixlb, iy) * v2dia * A(rbfbt, ixlb + 1, iy);
          }
          if (ix == ixrb) {
            const double v2dia = -0.5 * (A(gpey, ixrb, iy) + A(gpey, ixrb, iy1)) / (A(btot, ixrb, iy) * qe * A(ne, ixrb, iy));
            fdiaxrb[iy] = A(ne, ixrb, iy) * A(sx, ixrb, iy) * v2dia * A(rbfbt, ixrb, iy);
          }
          A(v2dd[f], ix, iy) = -2. * difpr2[f] * A(gprx, ix, iy) / (A(pr, ix2, iy) / A(rbfbt, ix2, iy) + A(pr, ix, iy) / A(rbfbt, ix, iy)) -
                               2. * (fcdif * difni2[f] + 0.) * (A(ni[f], ix2, iy) - A(ni[f], ix, iy)) /
                                   (A(ni[f], ix2, iy) / (A(rbfbt, ix2, iy) * A(gx, ix2, iy)) + A(ni[f], ix, iy) / (A(rbfbt, ix, iy) * A(gx, ix, iy)));
          const double temp3 = 4. * (A(prev, ix, iy) - A(prev, ix, iy1)) * A(gyc, ix, iy);
          A(ve2cd, ix, iy) = -temp3 / ((A(btot, ix, iy) + A(btot, ix2, iy)) * qe * (A(ni[f], ix, iy) + A(ni[f], ix2, iy)));
          const double v2rd = cfrd == 0. ? 0. : -2. * A(gprx, ix, iy) / (A(btot, ix, iy) / (etaper(ix, iy) * A(rbfbt2, ix, iy)) + A(btot, ix2, iy) / (etaper(ix2, iy) * A(rbfbt2, ix2, iy)));
          A(v2[f], ix, iy) = cf2dd * v2cd + cfrd * v2rd + A(v2dd[f], ix, iy) + cf2ef * A(v2ce[f], ix, iy) + cf2bf * A(v2cb[f], ix, iy);
          A(v2xgp[f], ix, iy) = 0.5 * (A(rbfbt, ix, iy) + A(rbfbt, ix2, iy)) * ((cf2dd + cf2bf) * v2cd + cfrd * v2rd + A(v2dd[f], ix, iy) + cf2ef * A(v2ce[f], ix, iy));
          if (isnonog == 1 && iy <= ny) {  // oderhs.m:1408-1432
            double grdnv = grdnv_y(ni[f], ix, iy, 1) / A(dxnog, ix, iy);
            A(vytan[f], ix, iy) = (fcdif * difni[f] + 0.) * (grdnv / ue_cos(A(angfx, ix, iy)) - (ue_log(A(ni[f], ix2, iy)) - ue_log(A(ni[f], ix, iy))) * A(gxf, ix, iy));
          }
        }
      if (inrow(ny + 1)) FOR1(ix, i1, i6) A(vy[f], ix, ny + 1) = 0.0;
    }
    if (isphion + isphiofft == 1) calc_currents(w);  // oderhs.m:1499
    // thermal force / friction (oderhs.m:1516-1534)
    FOR2(iy, j1, j6, ix, i1, i6) {
        int ix2 = IXP1(ix, iy);
        double nbarx = 0.5 * (A(ne, ix, iy) + A(ne, ix2, iy));
        double ltmax = mn(fabs(A(te, ix, iy) / (A(rrv, ix, iy) * A(gtex, ix, iy) + cutlo)), A(lcone, ix, iy));
        double lmfpe = 2e16 * ((A(te, ix, iy) / ev) * (A(te, ix, iy) / ev)) / A(ne, ix, iy);
        double flxlimf = flalftf * ltmax / (flalftf * ltmax + lmfpe);
        A(frice, ix, iy) = -cthe * flxlimf * nbarx * A(rrv, ix, iy) * A(gtex, ix, iy) + cfnetap * qe * A(netap, ix, iy) * A(fqp, ix, iy) / A(sx, ix, iy);
        A(frici[0], ix, iy) = -A(frice, ix, iy);
      }
    // parallel electric field from the electron momentum balance when the potential is not solved (oderhs.m:1541-1567)
    if (isphion == 0)
      FOR2(iy, w.iys1, w.iyf6, ix, i1, i6) {
          int ix1 = ix;
          if (ix == ixlb) ix1 = ixlb + 1; else if (ix == ixrb) ix1 = ixrb - 1;
          int ix2 = IXP1(ix1, iy);
          double ltmax = mn(fabs(A(te, ix, iy) / (A(rrv, ix, iy) * A(gtex, ix, iy) + cutlo)), A(lcone, ix, iy));
          double lmfpe = 2e16 * ((A(te, ix, iy) / ev) * (A(te, ix, iy) / ev)) / A(ne, ix, iy);
          double flxlimf = flalftf * ltmax / (flalftf * ltmax + lmfpe);
          double nexface = 0.5 * (A(ne, ix2, iy) + A(ne, ix1, iy));
          A(ex, ix, iy) = (1 - isphiofft) * (-(A(gpex, ix1, iy) / nexface + cthe * flxlimf * A(gtex, ix1, iy)) / qe - 0. + 0.) +
                          isphiofft * ((A(phi, ix1, iy) - A(phi, ix2, iy)) * A(gxf, ix1, iy));
        }
    // upi, uup (oderhs.m:1577-1643): every species with a momentum equation takes its own up
    if (xc > 0)
      FOR1(iy, w.iys1, w.iyf6) {
        int ix1 = IXM1(xc, iy);
        for (int f = 0; f < nfsp; ++f) { if (f < nusp) A(upi[f], ix1, iy) = A(up[f], ix1, iy); A(uup[f], ix1, iy) = A(rrv, ix1, iy) * A(upi[f], ix1, iy); }
      }
    FOR2(iy, w.iys1, w.iyf6, ix, w.ixs1, mn(w.ixf6, nx))
        for (int f = 0; f < nfsp; ++f) { if (f < nusp) A(upi[f], ix, iy) = A(up[f], ix, iy); A(uup[f], ix, iy) = A(rrv, ix, iy) * A(upi[f], ix, iy); }
    // poloidal velocities uu (oderhs.m:1648-1682)
    for (int f = 0; f < nfsp; ++f) {
      auto uuf = [&](int ix, int ixe, int iy) {  // face ix, its eastern cell ixe
        return A(uup[f], ix, iy) + 0.5 * (A(rbfbt, ixe, iy) + A(rbfbt, ix, iy)) * A(v2[f], ix, iy) - A(vytan[f], ix, iy) -
               difax[f] * 0.5 * (sq(0.5 * (A(ni[f], ix, iy) / A(ni[f], ixe, iy) + A(ni[f], ixe, iy) / A(ni[f], ix, iy)) - 1)) * (A(ni[f], ixe, iy) - A(ni[f], ix, iy)) * A(gxf, ix, iy) /
                   (A(ni[f], ixe, iy) + A(ni[f], ix, iy));
      };
      if (i1 > 0) FOR1(iy, j1, j6) { int ix1 = IXM1(i1, iy); A(uu[f], ix1, iy) = uuf(ix1, i1, iy); }
      FOR2(iy, j1, j6, ix, i1, i6) A(uu[f], ix, iy) = uuf(ix, IXP1(ix, iy), iy);
    }
    // electron velocities (oderhs.m:1729-1808)
    FOR2(iy, j1, j6, ix, i1, i6) { A(vex, ix, iy) = 0.; A(vey, ix, iy) = 0.; A(upe, ix, iy) = 0.; }
    for (int f = 0; f < nfsp; ++f)
      FOR2(iy, j1, j6, ix, i1, i6) {
          int ix1 = IXP1(ix, iy);
          A(upe, ix, iy) = A(upe, ix, iy) + A(upi[f], ix, iy) * zi[f] * 0.5 * (A(ni[f], ix, iy) + A(ni[f], ix1, iy));
        }
    FOR2(iy, j1, j6, ix, i1, i6) {
        int ix1 = IXP1(ix, iy);
        A(upe, ix, iy) = (A(upe, ix, iy) - 1. * A(fqp, ix, iy) / (A(rrv, ix, iy) * A(sx, ix, iy) * qe)) / (0.5 * (A(ne, ix, iy) + A(ne, ix1, iy)));
      }
    FOR2(iy, j1, j6, ix, i1, i6)
        A(vex, ix, iy) = A(upe, ix, iy) * A(rrv, ix, iy) + (cf2ef * A(v2ce[0], ix, iy) + cf2bf * A(ve2cb, ix, iy) + cf2dd * A(ve2cd, ix, iy)) * 0.5 * (A(rbfbt, ix, iy) + A(rbfbt, IXP1(ix, iy), iy)) - A(vytan[0], ix, iy);
    for (int f = 0; f < nfsp; ++f)
      FOR2(iy, j1, j5, ix, i1, i6) A(vey, ix, iy) = A(vey, ix, iy) + A(vy[f], ix, iy) * zi[f] * 0.5 * (A(niy0[f], ix, iy) + A(niy1[f], ix, iy));
    FOR2(iy, j1, j5, ix, i1, i6) A(vey, ix, iy) = (A(vey, ix, iy) - cfjve * A(fqy, ix, iy) / (A(sy, ix, iy) * qe)) / (0.5 * (A(ney0, ix, iy) + A(ney1, ix, iy)));
    if (isnewpot == 1 && inrow(0))  // fqy(,0) = 0 there (oderhs.m:1794-1800)
      FOR1(ix, i1, i6) A(vey, ix, 0) = cfybf * A(veycb, ix, 0) + A(vydd[0], ix, 0) + cfyef * A(vyce[0], ix, 0);
    if (isybdrywd == 1)  // vey diffusive in wall cells, like vy (oderhs.m:1803-1808)
      FOR1(ix, i1, i6) {
        if (inrow(0) && matwalli[ix] > 0) A(vey, ix, 0) = A(vydd[0], ix, 0);
        if (inrow(ny) && matwallo[ix] > 0) A(vey, ix, ny) = A(vydd[0], ix, ny);
      }

    // zero the source accumulators (oderhs.m:1818-1835)
    FOR2(iy, j2, j5, ix, i2, i5) {
        for (int f = 0; f < nfsp; ++f) { A(snic[f], ix, iy) = 0.; A(sniv[f], ix, iy) = 0.; A(psori[f], ix, iy) = 0.; }
        for (int f = 0; f < nusp; ++f) { A(smoc[f], ix, iy) = 0.; A(smov[f], ix, iy) = 0.; }
        A(seec, ix, iy) = 0.; A(seev, ix, iy) = 0.; A(seic, ix, iy) = 0.; A(seiv, ix, iy) = 0.;
      }
    // ionisation / recombination / charge exchange (oderhs.m:1909-2008), hydrogen ions = species 1, gas species 1
    double nuizold = 0., nurcold = 0.;
    if (xc >= 0 && yc >= 0) { nuizold = A(nuiz, xc, yc); nurcold = A(nurc, xc, yc); }
    FOR2(iy, w.iys1, w.iyf6, ix, w.ixs1, w.ixf6) {
        if (icnuiz == 0) {
          double ne_sgvi = A(ne, ix, iy);
          if (ifxnsgi == 1) ne_sgvi = cne_sgvi;
          A(nuiz, ix, iy) = chioniz * A(ne, ix, iy) * (rsa(A(te, ix, iy), ne_sgvi) + sigvi_floor);
          if (xc >= 0) A(nuiz, ix, iy) = fnnuiz * A(nuiz, ix, iy) + (1 - fnnuiz) * nuizold;
        } else if (icnuiz == 1) A(nuiz, ix, iy) = cnuiz;
        if (isrecmon == 1) {
          A(nurc, ix, iy) = cfrecom * A(ne, ix, iy) * rra(A(te, ix, iy), A(ne, ix, iy));
          if (xc >= 0) A(nurc, ix, iy) = fnnuiz * A(nurc, ix, iy) + (1 - fnnuiz) * nurcold;
        } else A(nurc, ix, iy) = 0.;
        A(psorbgg, ix, iy) = ngbackg_[0] * ((0.9 + 0.1 * powi(ngbackg_[0] / A(ng, ix, iy), ingb))) * A(nuiz, ix, iy) * A(vol, ix, iy);
        A(psorgc, ix, iy) = -A(ng, ix, iy) * A(nuiz, ix, iy) * A(vol, ix, iy) + A(psorbgg, ix, iy);
        A(psorc[0], ix, iy) = -A(psorgc, ix, iy);
        A(psordis, ix, iy) = cfdiss * A(psorc[0], ix, iy);
        A(psorxrc[0], ix, iy) = -A(ni[0], ix, iy) * A(nurc, ix, iy) * A(vol, ix, iy);
        A(psorrgc, ix, iy) = -A(psorxrc[0], ix, iy);
        if (icnucx == 0) {
          double t0 = mx(A(ti, ix, iy), temin * ev);
          double t1 = t0 / (mi[0] / mp);
          A(nucx, ix, iy) = A(ni[0], ix, iy) * rcx(t1);
        } else if (icnucx == 1) A(nucx, ix, iy) = cnucx;
        else {
          double t0 = mx(A(ti, ix, iy), temin * ev);
          A(nucx, ix, iy) = sqrt(t0 / mi[0]) * sigcx * (A(ni[0], ix, iy) + rnn2cx * A(ng, ix, iy));
        }
        A(nuix, ix, iy) = fnuizx * A(nuiz, ix, iy) + fnucxx * A(nucx, ix, iy);
        if (isupgon == 1) { A(psorc[1], ix, iy) = -A(psorc[0], ix, iy); A(psorxrc[1], ix, iy) = -A(psorxrc[0], ix, iy); }
      }
    FOR2(iy, w.iys1, w.iyf6, ix, w.ixs1, w.ixf6) {  // ispsorave = 0 (oderhs.m:2017-2029)
        A(psorg, ix, iy) = A(psorgc, ix, iy); A(psor[0], ix, iy) = A(psorc[0], ix, iy);
        A(psorxr[0], ix, iy) = A(psorxrc[0], ix, iy); A(psorrg, ix, iy) = A(psorrgc, ix, iy);
        if (isupgon == 1) { A(psor[1], ix, iy) = -A(psor[0], ix, iy); A(psorxr[1], ix, iy) = -A(psorxr[0], ix, iy); }
      }

    if (cfqyn > 0.) {  // calc_curr_cx (potencur.m:446-495): current from charge exchange and neoclassical damping, now that nucx is known
      FOR2(iy, mx(w.j1p, 1), mn(w.j5p, ny - 1), ix, i1, i6) {
          const double omgci = qe * A(b_c, ix, iy) / mi[0];
          const int ix3 = IXM1(ix, iy + 1), ix4 = IXM1(ix, iy);
          A(fqyn, ix, iy) = qe * 0.125 * ((A(ngy0, ix, iy) + A(ngy1, ix, iy)) * A(nucx, ix, iy) + (A(niy0[0], ix, iy) + A(niy1[0], ix, iy)) * nuneo) * A(sy, ix, iy) *
                            (A(v2ce[0], ix, iy) + A(v2cd[0], ix, iy) + A(v2ce[0], ix, iy + 1) + A(v2cd[0], ix, iy + 1) + A(v2ce[0], ix4, iy) + A(v2cd[0], ix4, iy) + A(v2ce[0], ix3, iy + 1) + A(v2cd[0], ix3, iy + 1)) / omgci;
        }
      FOR2(iy, w.j1p, w.j5p, ix, i1, i6) A(fqy, ix, iy) = A(fqy, ix, iy) + cfqyn * A(fqyn, ix, iy);
    }
    if (ineudif == 1) neudif(w); else neudifpg(w);  // oderhs.m:2423-2435

    // half-space problem: no flux and no gradients through the cut (oderhs.m:2447-2466)
    if (isfixlb == 2) {
      const int ix = ixpt2;
      if (ix >= i2 && ix <= i5 + 1 && iysptrx1 > 0)
        FOR1(iy, mx(0, rowlo), mn(iysptrx1, rowhi)) {
          A(gpex, ix, iy) = 0.; A(frice, ix, iy) = 0.; A(ex, ix, iy) = 0.; A(upe, ix, iy) = 0.;
          for (int f = 0; f < nfsp; ++f) { A(gpix[f], ix, iy) = 0.; A(frici[f], ix, iy) = 0.; A(uu[f], ix, iy) = 0.; A(upi[f], ix, iy) = 0.; }
        }
    }
    // electron pressure work and momentum source (oderhs.m:2471-2500)
    FOR2(iy, j2, j5, ix, i2, i5) {
        int ix1 = IXM1(ix, iy), ix2 = IXP1(ix, iy);
        double t1old = .5 * cvgp * (A(upe, ix, iy) * A(rrv, ix, iy) * ave(A(gx, ix, iy), A(gx, ix2, iy)) * A(gpex, ix, iy) / A(gxf, ix, iy) +
                                    A(upe, ix1, iy) * A(rrv, ix1, iy) * ave(A(gx, ix, iy), A(gx, ix1, iy)) * A(gpex, ix1, iy) / A(gxf, ix1, iy));
        double t2old = 1.e-20 * 0.25 * (A(fqp, ix, iy) + A(fqp, ix1, iy)) * (A(ex, ix, iy) + A(ex, ix1, iy)) / A(gx, ix, iy);
        int iyp1 = mn(iy + 1, ny + 1), iym1 = mx(iy - 1, 0);
        double t1new = .5 * cvgp * (A(vex, ix, iy) * ave(A(gx, ix, iy), A(gx, ix2, iy)) * A(gpex, ix, iy) / A(gxf, ix, iy) +
                                    A(vex, ix1, iy) * ave(A(gx, ix, iy), A(gx, ix1, iy)) * A(gpex, ix1, iy) / A(gxf, ix1, iy));
        double t2new = .5 * cvgp * (A(vey, ix, iy) * ave(A(gy, ix, iy), A(gy, ix, iyp1)) * A(gpey, ix, iy) / A(gyf, ix, iy) +
                                    A(vey, ix, iy) * ave(A(gy, ix, iy), A(gy, ix, iym1)) * A(gpey, ix, iym1) / A(gyf, ix, iym1));
        A(seec, ix, iy) = A(seec, ix, iy) + (t1old * A(vol, ix, iy) - t2old) * oldseec + ((t1new + t2new) * A(vol, ix, iy)) * (1 - oldseec);
        if (nusp - isupgon == 1) A(smoc[0], ix, iy) = ((-cpgx * A(gpex, ix, iy) - 0.) * A(rrv, ix, iy) + 0.) * A(sx, ix, iy) / A(gxf, ix, iy);
      }
    for (int f = 0; f < nusp; ++f) {  // oderhs.m:2502-2579
      if (!(zi[f] > 1.e-20)) continue;
      FOR2(iy, j2, j5, ix, i2, i5) {
          int ix1 = IXM1(ix, iy), ix2 = IXP1(ix, iy);
          double tv = A(gpix[f], ix, iy) / A(gxf, ix, iy);
          double t1 = A(gpix[f], ix1, iy) / A(gxf, ix1, iy);
          t1 = .5 * cvgp * (A(up[f], ix, iy) * A(rrv, ix, iy) * ave(A(gx, ix2, iy), A(gx, ix, iy)) * tv + A(up[f], ix1, iy) * A(rrv, ix1, iy) * ave(A(gx, ix, iy), A(gx, ix1, iy)) * t1);
          A(seic, ix, iy) = A(seic, ix, iy) + cfvgpx[f] * t1 * A(vol, ix, iy);
          double t0 = -cpiup[f] * (A(gpix[f], ix, iy) * A(rrv, ix, iy) - 0.) * A(sx, ix, iy) / A(gxf, ix, iy);
          if (nusp - isupgon == 1) A(smoc[0], ix, iy) = A(smoc[0], ix, iy) + cpgx * t0;
          else {
            t0 = t0 + (qe * zi[f] * 0.5 * (A(ni[f], ix2, iy) + A(ni[f], ix, iy)) * A(ex, ix, iy) * A(rrv, ix, iy) + A(frici[f], ix, iy)) * A(sx, ix, iy) / A(gxf, ix, iy);
            A(smoc[f], ix, iy) = A(smoc[f], ix, iy) + cpgx * t0;
          }
          tv = 0.25 * (A(frice, ix, iy) + A(frice, ix1, iy)) * (A(upe, ix, iy) + A(upe, ix1, iy) - A(upi[f], ix, iy) - A(upi[f], ix1, iy));
          A(seec, ix, iy) = A(seec, ix, iy) - (zi[f] * zi[f]) * A(ni[f], ix, iy) * tv * A(vol, ix, iy) / A(nz2, ix, iy);
        }
      FOR2(iy, j2, j5, ix, i2, i5) {
          double t1, t2;
          if (isgpye == 0) {
            int ix1 = IXM1(ix, iy), ix2 = IXP1(ix, iy);
            t1 = .5 * cvgp * (A(vygp[f], ix, iy) * A(gpiy[f], ix, iy) + A(vygp[f], ix, iy - 1) * A(gpiy[f], ix, iy - 1) +
                              A(v2xgp[f], ix, iy) * ave(A(gx, ix, iy), A(gx, ix2, iy)) * A(gpix[f], ix, iy) / A(gxf, ix, iy) +
                              A(v2xgp[f], ix1, iy) * ave(A(gx, ix, iy), A(gx, ix1, iy)) * A(gpix[f], ix1, iy) / A(gxf, ix1, iy));
            t2 = t1;
          } else { t1 = -0.5 * (A(vy[f], ix, iy) * A(gpey, ix, iy) + A(vy[f], ix, iy - 1) * A(gpey, ix, iy - 1)); t2 = t1; }
          A(seec, ix, iy) = A(seec, ix, iy) - fluxfacy * t1 * A(vol, ix, iy);
          A(seic, ix, iy) = A(seic, ix, iy) + fluxfacy * cfvgpy[f] * t2 * A(vol, ix, iy);
        }
    }
    if (isupgon == 1 && zi[1] < 1.e-20) {  // oderhs.m:2581-2616: v.grad(p) of the atoms into the ion+atom energy equation
      if (cfvgpx[1] > 0.) {
        FOR2(iy, j2, j5, ix, i2, i5) {
            int ix1 = IXM1(ix, iy), iy1 = mx(0, iy - 1);
            A(seic, ix, iy) = A(seic, ix, iy) + cftiexclg * 0.5 * cfvgpx[1] * (A(uuxg, ix, iy) * A(gpix[1], ix, iy) + A(uuxg, ix1, iy) * A(gpix[1], ix1, iy)) * A(vol, ix, iy);
            A(seic, ix, iy) = A(seic, ix, iy) + cftiexclg * 0.5 * cfvgpy[1] * (A(vyg, ix, iy) * A(gpiy[1], ix, iy) + A(vyg, ix, iy1) * A(gpiy[1], ix, iy1)) * A(vol, ix, iy);
          }
      } else {
        FOR2(iy, j2, j5, ix, i2, i5) {
            int ix1 = IXM1(ix, iy), ix2 = IXP1(ix, iy);
            double tv = A(gpix[1], ix, iy) / A(gxf, ix, iy);
            double t1 = A(gpix[1], ix1, iy) / A(gxf, ix1, iy);
            t1 = .5 * cvgp * (A(up[1], ix, iy) * A(rrv, ix, iy) * ave(A(gx, ix2, iy), A(gx, ix, iy)) * tv + A(up[1], ix1, iy) * A(rrv, ix1, iy) * ave(A(gx, ix, iy), A(gx, ix1, iy)) * t1);
            A(seic, ix, iy) = A(seic, ix, iy) + cftiexclg * t1 * A(vol, ix, iy);
          }
      }
    }
    // viscosities (oderhs.m:2624-2799)
    for (int f = 0; f < nfsp; ++f) {
      if (isupgon == 1 && zi[f] == 0) {  // inertial atoms
        FOR2(iy, j1, j6, ix, i1, i6) {
            const int iyp1 = mn(iy + 1, ny + 1);
            const int ix1 = IXM1(ix, iy);
            double vtn = sqrt(mx(A(tg, ix, iy), tgmin * ev) / mi[f]);
            double qfl = flalfvgxa[ix] * A(nm[f], ix, iy) * vtn * vtn;
            double lmfppar = vtn / (kelhihg * A(ni[0], ix, iy) + kelhghg * A(ni[f], ix, iy));
            double lmfpperp = vtn / (vtn * sigcx * A(ni[0], ix, iy) + kelhihg * A(ni[0], ix, iy) + kelhghg * A(ni[f], ix, iy));
            double rrfac = A(rr, ix, iy) * A(rr, ix, iy);
            double lmfpn = lmfppar * rrfac + lmfpperp * (1 - rrfac);
            double csh = lmfpn * A(nm[f], ix, iy) * vtn * lgvmax / (lgvmax + lmfpn);
            double qsh;
            if (isgxvon == 0) qsh = csh * (A(up[f], ix1, iy) - A(up[f], ix, iy)) * A(gx, ix, iy);
            else qsh = csh * (A(up[f], ix1, iy) - A(up[f], ix, iy)) * 2 * A(gxf, ix, iy) * A(gxf, ix1, iy) / (A(gxf, ix, iy) + A(gxf, ix1, iy));
            A(visx[f], ix, iy) = cfvisxn * csh / ue_pow(1 + ue_pow(fabs(qsh / (qfl + cutlo)), flgamvg), 1. / flgamvg) + 0. * travis[f] * A(nm[f], ix, iy);
            const int ix2 = IXP1(ix, iy), ix3 = IXP1(ix, iyp1);
            double tgupyface = 0.25 * (A(tg, ix, iy) + A(tg, ix, iyp1) + A(tg, ix2, iy) + A(tg, ix3, iyp1));
            vtn = sqrt(mx(tgupyface, tgmin * ev) / mi[f]);
            double nmxface = 0.5 * (A(nm[f], ix, iy) + A(nm[f], ix2, iy));
            double ngupyface = 0.25 * (A(ni[f], ix, iy) + A(ni[f], ix, iyp1) + A(ni[f], ix2, iy) + A(ni[f], ix3, iyp1));
            double n1upyface = 0.25 * (A(ni[0], ix, iy) + A(ni[0], ix, iyp1) + A(ni[0], ix2, iy) + A(ni[0], ix3, iyp1));
            lmfppar = vtn / (kelhihg * n1upyface + kelhghg * ngupyface);
            lmfpperp = vtn / (vtn * sigcx * n1upyface + kelhihg * n1upyface + kelhghg * ngupyface);
            lmfpn = lmfppar * rrfac + lmfpperp * (1 - rrfac);
            csh = lmfpn * ngupyface * mi[f] * vtn * lgvmax / (lgvmax + lmfpn);
            qfl = flalfvgya[iy] * ngupyface * mi[f] * vtn * vtn;
            qsh = csh * (A(up[f], ix, iy) - A(up[f], ix, iyp1)) * A(gyf, ix, iy);
            A(visy[f], ix, iy) = cfvisyn * csh / ue_pow(1 + ue_pow(fabs(qsh / (qfl + cutlo)), flgamvg), 1. / flgamvg) + 0. * travis[f] * nmxface;
          }
      }
      if (zi[f] > 1.e-20) {
        FOR2(iy, j1, j6, ix, i1, i6) A(this->w, ix, iy) = 0.0;
        for (int jf = 0; jf < nisp; ++jf) {
          double tv = (zi[jf] * zi[jf]) / sqrt((mi[f] + mi[jf]) / (2 * mp));
          FOR2(iy, j1, j6, ix, i1, i6) A(this->w, ix, iy) = A(this->w, ix, iy) + tv * A(ni[jf], ix, iy);
        }
        FOR2(iy, j1, j6, ix, i1, i6) {
            double ctaui = 2.1e13 / (A(loglambda, ix, iy) * (zi[f] * zi[f]));
            double tv2 = ctaui / (ev * sqrt(ev));
            double a = (convis == 0) ? mx(A(ti, ix, iy), temin * ev) : afix * ev;
            double visxtmp = tv2 * coef * A(rr, ix, iy) * A(rr, ix, iy) * a * a * sqrt(a) * A(ni[f], ix, iy) / A(this->w, ix, iy);
            A(visx[f], ix, iy) = parvis[f] * visxtmp + 0. * A(nm[f], ix, iy);
            int ix1 = IXM1(ix, iy);
            double t0 = mx(A(ti, ix, iy), temin * ev);
            double mfl = flalfv * A(nm[f], ix, iy) * A(rr, ix, iy) * A(vol, ix, iy) * A(gx, ix, iy) * (t0 / mi[f]);
            double csh;
            if (isgxvon == 0) csh = A(visx[f], ix, iy) * A(vol, ix, iy) * A(gx, ix, iy) * A(gx, ix, iy);
            else csh = A(visx[f], ix, iy) * A(vol, ix, iy) * A(gx, ix, iy) * 2 * A(gxf, ix, iy) * A(gxf, ix1, iy) / (A(gxf, ix, iy) + A(gxf, ix1, iy));
            double msh = fabs(csh * (A(upi[f], ix1, iy) - A(upi[f], ix, iy)));
            A(visx[f], ix, iy) = A(visx[f], ix, iy) / ue_pow(1 + ue_pow(msh / (mfl + 1.e-20 * msh), flgamv), 1 / flgamv);
            A(visy[f], ix, iy) = (fcdif * travis[f] + 0.) * A(nm[f], ix, iy) + 4 * 0.;
          }
      }
    }
    // heat conduction coefficients (oderhs.m:2801-3069)
    FOR2(iy, j1, j6, ix, i1, i6) {
        A(hcxe, ix, iy) = 0.; A(hcxi, ix, iy) = 0.; A(hcye, ix, iy) = 0.; A(hcyi, ix, iy) = 0.;
        for (int f = 0; f < nisp; ++f) { A(hcxij[f], ix, iy) = 0.; A(hcyij[f], ix, iy) = 0.; }
      }
    for (int f = 0; f < nisp; ++f) {
      if (zi[f] == 0.0) continue;
      FOR2(iy, j1, j6, ix, i1, i6) { A(w1, ix, iy) = 0.; A(w2, ix, iy) = 0.; }
      for (int jf = 0; jf < nisp; ++jf) {
        double tv = zi[jf] * zi[jf];
        double a = (zi[jf] * zi[jf]) * sqrt(2 * mi[f] * mi[jf] / (mi[f] + mi[jf]));
        FOR2(iy, j1, j6, ix, i1, i6) {
            int ix1 = IXP1(ix, iy);
            A(w1, ix, iy) = A(w1, ix, iy) + tv * (A(ni[jf], ix, iy) * A(gx, ix, iy) + A(ni[jf], ix1, iy) * A(gx, ix1, iy)) / (A(gx, ix, iy) + A(gx, ix1, iy));
            A(w2, ix, iy) = A(w2, ix, iy) + a * (A(ni[jf], ix, iy) * A(gx, ix, iy) + A(ni[jf], ix1, iy) * A(gx, ix1, iy)) / (A(gx, ix, iy) + A(gx, ix1, iy));
          }
      }
      FOR2(iy, j1, j6, ix, i1, i6) {
          int ix1 = IXP1(ix, iy), iyp1 = mn(ny + 1, iy + 1);
          double ctaue = 3.5e11 * zi[f] / A(loglambda, ix, iy);
          double ctaui = 2.1e13 / (A(loglambda, ix, iy) * (zi[f] * zi[f]));
          double fxe = kxe * ce * ctaue / (me * ev * sqrt(ev));
          double fxi = kxi * ci * ctaui / (ev * sqrt(ev * mp));
          double fxet = fxe, fxit = fxi;
          if ((iy <= iysptrx) && ix > ixpt1 && ix <= ixpt2) {
            fxet = fxe / (1. + (rkxecore - 1.) * powi(yyf[iy] / (yyf[0] + 4.e-50), inkxc));
            fxit = kxicore * fxi;
          }
          double niavex = (A(ni[f], ix, iy) * A(gx, ix, iy) + A(ni[f], ix1, iy) * A(gx, ix1, iy)) / (A(gx, ix, iy) + A(gx, ix1, iy));
          double niavey = (A(niy0[f], ix, iy) * A(gy, ix, iy) + A(niy1[f], ix, iy) * A(gy, ix, iyp1)) / (A(gy, ix, iy) + A(gy, ix, iyp1));
          A(hcxe, ix, iy) = A(hcxe, ix, iy) + fxet * niavex / A(w1, ix, iy);
          double kyemix = fcdif * kye + 0.;
          if (kyet > 1.e-20 && iy > iysptrx) kyemix = (1. - ckyet) * kyemix + ckyet * kyet * A(diffusivwrk, ix, iy);
          A(hcye, ix, iy) = A(hcye, ix, iy) + (kyemix + 2.33 * (dclass_e_(ix, iy) + dclass_e_(ix, iyp1))) * zi[f] * niavey;
          A(hcxij[f], ix, iy) = fxit * niavex / A(w2, ix, iy);
          double kyimix = fcdif * kyi + 0.;
          if (kyit > 1.e-20 && iy > iysptrx) kyimix = (1. - ckyit) * kyimix + ckyit * kyit * A(diffusivwrk, ix, iy);
          A(hcyij[f], ix, iy) = A(hcyij[f], ix, iy) + (kyimix + (dclass_i_(ix, iy) + dclass_i_(ix, iyp1))) * niavey;
        }
    }
    for (int f = 0; f < nisp; ++f) {  // oderhs.m:2906-2965
      if (zi[f] == 0.) continue;
      FOR2(iy, j1, j6, ix, i1, i6) {
          int ix1 = IXP1(ix, iy);
          double a, tiave = 0.;
          if (concap == 0) {
            tiave = (A(ti, ix, iy) * A(gx, ix, iy) + A(ti, ix1, iy) * A(gx, ix1, iy)) / (A(gx, ix, iy) + A(gx, ix1, iy));
            if (ix == ixlb) tiave = A(ti, ixlb + 1, iy);
            if (ix == ixrb) tiave = A(ti, ixrb, iy);
            a = mx(tiave, temin * ev);
          } else a = afix * ev;
          A(hcxij[f], ix, iy) = A(hcxij[f], ix, iy) * A(rrv, ix, iy) * A(rrv, ix, iy) * a * a * sqrt(a);
          double lmfpi = 1.e16 * ((tiave / ev) * (tiave / ev)) / A(ni[0], ix, iy);
          double niavex = (A(ni[f], ix, iy) * A(gx, ix, iy) + A(ni[f], ix1, iy) * A(gx, ix1, iy)) / (A(gx, ix, iy) + A(gx, ix1, iy));
          A(hcxij[f], ix, iy) = A(hcxij[f], ix, iy) / (1. + lmfpi / lmfplim);
          double dti = A(ti, ix, iy) - A(ti, ix1, iy);
          double sti = 0.5 * alfkxi * (A(ti, ix, iy) + A(ti, ix1, iy));
          A(hcxij[f], ix, iy) = A(hcxij[f], ix, iy) * (cutlo + dti * dti) / (cutlo + dti * dti + sti * sti) + 0. * niavex;
          if (isflxldi == 2) {
            niavex = (A(ni[f], ix, iy) * A(gx, ix, iy) + A(ni[f], ix1, iy) * A(gx, ix1, iy)) / (A(gx, ix, iy) + A(gx, ix1, iy));
            double wallfac = 1.;
            if ((ix == ixlb || ix == ixrb) && (isplflxl == 0)) wallfac = flalfipl / flalfi;
            double qflx = wallfac * flalfi * A(rrv, ix, iy) * sqrt(a / mi[f]) * niavex * a;
            double cshx = A(hcxij[f], ix, iy);
            double lxtic = 0.5 * (A(ti, ix, iy) + A(ti, ix1, iy)) / (fabs(A(ti, ix, iy) - A(ti, ix1, iy)) * A(gxf, ix, iy) + 100. * cutlo);
            double qshx = cshx * (A(ti, ix, iy) - A(ti, ix1, iy)) * A(gxf, ix, iy) * (1. + lxtic / lxtimax);
            A(hcxij[f], ix, iy) = cshx / (1 + fabs(qshx / qflx));
          }
          A(hcxi, ix, iy) = A(hcxi, ix, iy) + A(hcxij[f], ix, iy);
          A(hcyi, ix, iy) = A(hcyi, ix, iy) + A(hcyij[f], ix, iy);
        }
    }
    FOR2(iy, j1, j6, ix, i1, i6) {  // oderhs.m:2968-3017
        int ix1 = IXP1(ix, iy), iyp1 = mn(ny + 1, iy + 1);
        double a;
        if (concap == 0) {
          double teave = (A(te, ix, iy) * A(gx, ix, iy) + A(te, ix1, iy) * A(gx, ix1, iy)) / (A(gx, ix, iy) + A(gx, ix1, iy));
          if (ix == ixlb) teave = A(te, ixlb + 1, iy);
          if (ix == ixrb) teave = A(te, ixrb, iy);
          a = mx(teave, temin * ev);
        } else a = afix * ev;
        double zeffave = (A(zeff, ix, iy) * A(gx, ix, iy) + A(zeff, ix1, iy) * A(gx, ix1, iy)) / (A(gx, ix, iy) + A(gx, ix1, iy));
        double zcoef = 0.308 + 0.767 * zeffave - 0.075 * (zeffave * zeffave);
        A(hcxe, ix, iy) = A(hcxe, ix, iy) * A(rrv, ix, iy) * A(rrv, ix, iy) * a * a * sqrt(a) * zcoef;
        double lmfpe = 2e16 * ((A(te, ix, iy) / ev) * (A(te, ix, iy) / ev)) / A(ne, ix, iy);
        double neavex = (A(ne, ix, iy) * A(gx, ix, iy) + A(ne, ix1, iy) * A(gx, ix1, iy)) / (A(gx, ix, iy) + A(gx, ix1, iy));
        double dte = A(te, ix, iy) - A(te, ix1, iy);
        double ste = 0.5 * alfkxe * (A(te, ix, iy) + A(te, ix1, iy));
        A(hcxe, ix, iy) = A(hcxe, ix, iy) * (cutlo + dte * dte) / (cutlo + dte * dte + ste * ste) + 0. * neavex;
        A(hcxe, ix, iy) = A(hcxe, ix, iy) / ((1. + lmfpe / lmfplim) * (1 + A(hcxe, ix, iy) * (A(gx, ix, iy) * A(gx, ix, iy)) * tdiflim / A(ne, ix, iy)));
        if (isupgon == 0) {
          A(hcxn, ix, iy) = 0.; A(hcyn, ix, iy) = 0.;
          A(hcxi, ix, iy) = A(hcxi, ix, iy) + cftiexclg * cfneut * cfneutsor_ei * kxn * (A(ng, ix, iy) * A(ti, ix, iy) + A(ng, ix1, iy) * A(ti, ix1, iy)) / (mi[0] * (A(nucx, ix, iy) + A(nucx, ix1, iy)));
          A(hcyi, ix, iy) = A(hcyi, ix, iy) + cftiexclg * cfneut * cfneutsor_ei * kyn * (A(ngy0, ix, iy) * A(tiy0, ix, iy) + A(ngy1, ix, iy) * A(tiy1, ix, iy)) / (mi[0] * (A(nucx, ix, iy) + A(nucx, ix, iyp1)));
        }
      }
    if (isupgon == 1)  // oderhs.m:3019-3063
      FOR2(iy, j1, j6, ix, i1, i6) {
          const int iy1 = mn(iy, ny);
          const int ix1 = IXP1(ix, iy), g = iigsp;
          double tgavex = mx(0.5 * (A(tg, ix, iy) + A(tg, ix1, iy)), temin * ev);
          double tgavey = mx(0.5 * (A(tgy0, ix, iy) + A(tgy1, ix, iy)), temin * ev);
          double niavex = 0.5 * (A(ni[0], ix, iy) + A(ni[0], ix1, iy));
          double niavey = 0.5 * (A(niy0[0], ix, iy1) + A(niy1[0], ix, iy1));
          double noavex = (A(ni[g], ix, iy) * A(gx, ix, iy) + A(ni[g], ix1, iy) * A(gx, ix1, iy)) / (A(gx, ix, iy) + A(gx, ix1, iy));
          double noavey = 0.5 * (A(niy0[g], ix, iy1) + A(niy1[g], ix, iy1));
          double qflx = flalftgxa[ix] * sqrt(tgavex / mi[g]) * noavex * tgavex;
          double lmfpn = 1. / (sigcx * (niavex + rnn2cx * noavex));
          double cshx = lmfpn * sqrt(tgavex / mi[g]) * noavex * lgtmax[g] / (lmfpn + lgtmax[g]);
          double qshx = cshx * (A(tg, ix, iy) - A(tg, ix1, iy)) * A(gxf, ix, iy);
          A(hcxn, ix, iy) = cshx / ue_pow(1 + ue_pow(fabs(qshx / qflx), flgamtg), 1. / flgamtg);
          A(hcxi, ix, iy) = A(hcxi, ix, iy) + cftiexclg * cfneut * cfneutsor_ei * A(hcxn, ix, iy);
          double qfly = flalftgya[iy] * sqrt(tgavey / mi[g]) * noavey * tgavey;
          lmfpn = 1. / (sigcx * (niavey + rnn2cx * noavey));
          double cshy = lmfpn * sqrt(tgavey / mi[g]) * noavey * lgtmax[g] / (lmfpn + lgtmax[g]);
          double qshy = cshy * (A(tgy0, ix, iy1) - A(tgy1, ix, iy1)) / A(dynog, ix, iy);
          A(hcyn, ix, iy) = cshy / ue_pow(1 + ue_pow(fabs(qshy / qfly), flgamtg), 1. / flgamtg);
          A(hcyi, ix, iy) = A(hcyi, ix, iy) + cftiexclg * cfneut * cfneutsor_ei * A(hcyn, ix, iy);
        }
    // equipartition (oderhs.m:3074-3102)
    FOR2(iy, j1, j6, ix, i1, i6) A(w3, ix, iy) = 0.0;
    for (int f = 0; f < nisp; ++f) {
      double tv = (zi[f] * zi[f]) / mi[f];
      FOR2(iy, j2, j5, ix, i2, i5) A(w3, ix, iy) = A(w3, ix, iy) + tv * A(ni[f], ix, iy);
    }
    FOR2(iy, j2, j5, ix, i2, i5) {
        int ix2 = IXM1(ix, iy);
        double a = mx(A(te, ix, iy), temin * ev);
        double loglmcc = 0.5 * (A(loglambda, ix, iy) + A(loglambda, ix2, iy));
        double coef1 = feqp * 4.8e-15 * loglmcc * sqrt(ev) * ev * mp;
        A(eqp, ix, iy) = coef1 * A(w3, ix, iy) * A(ne, ix, iy) / (a * sqrt(a));
        double d = a - A(ti, ix, iy), s = alfeqp * (a + A(ti, ix, iy));
        A(eqp, ix, iy) = A(eqp, ix, iy) * (d * d) / (cutlo + d * d + s * s);
      }
    if (nisp >= 2) {  // gas conductivities as stored (oderhs.m:3156-3160) and atom/ion equipartition (oderhs.m:3163-3176)
      if (isupgon == 1) FOR1(c_, rowlo * NXS, (rowhi + 1) * NXS - 1) { hcxg[c_] = hcxn[c_]; hcyg[c_] = hcyn[c_]; }
      FOR2(iy, j1, j6, ix, i1, i6) A(eqpg, ix, iy) = cftgeqp * A(ng, ix, iy) * (A(ni[0], ix, iy) + cftiexclg * A(ni[1], ix, iy)) * keligig[0];
      engbalg(w);  // oderhs.m:3180
    }

    // ---- particle fluxes (oderhs.m:3187-3319) ----
    {
      const int methnx = methn % 10, methny = methn / 10;
      for (int f = 0; f < nfsp; ++f) {
        FOR2(iy, j4, j8, ix, i1, i5) {
            if (zi[f] == 0. && ineudif != 0 && 1. - A(rrv, ix, iy) > 1.e-4) { A(fnix[f], ix, iy) = A(fngx, ix, iy); continue; }
            int ix2 = IXP1(ix, iy);
            double t2;
            if (methnx == 2) t2 = (A(ni[f], ix, iy) + A(ni[f], ix2, iy)) / 2;
            else if (methnx == 3) t2 = (A(uu[f], ix, iy) >= 0.) ? A(ni[f], ix, iy) : A(ni[f], ix2, iy);
            else if (methnx == 6) t2 = ue_exp(0.5 * (ue_log(A(ni[f], ix, iy)) + ue_log(A(ni[f], ix2, iy))));
            else {
              double t0 = (A(ni[f], ix, iy) * A(gx, ix, iy) + A(ni[f], ix2, iy) * A(gx, ix2, iy)) / (A(gx, ix, iy) + A(gx, ix2, iy));
              double t1 = (A(gx, ix, iy) + A(gx, ix2, iy)) * A(ni[f], ix, iy) * A(ni[f], ix2, iy) / (cutlo + A(ni[f], ix, iy) * A(gx, ix2, iy) + A(ni[f], ix2, iy) * A(gx, ix, iy));
              t2 = (A(uu[f], ix, iy) * (A(ni[f], ix, iy) - A(ni[f], ix2, iy)) >= 0.) ? t0 : t1;
            }
            A(fnix[f], ix, iy) = cnfx * A(uu[f], ix, iy) * A(sx, ix, iy) * t2;
            double r1 = nlimix[f] * A(ni[f], ix, iy) / A(ni[f], ix2, iy), r2 = nlimix[f] * A(ni[f], ix2, iy) / A(ni[f], ix, iy);
            A(fnix[f], ix, iy) = A(fnix[f], ix, iy) / sqrt(1 + r1 * r1 + r2 * r2);
          }
        FOR2(iy, j1, j5, ix, i4, i8) {
            if (zi[f] == 0.) { A(fniy[f], ix, iy) = A(fngy, ix, iy); continue; }
            double t2;
            if (methny == 2) t2 = (A(niy0[f], ix, iy) + A(niy1[f], ix, iy)) / 2;
            else if (methny == 3) t2 = (A(vy[f], ix, iy) >= 0.) ? A(niy0[f], ix, iy) : A(niy1[f], ix, iy);
            else if (methny == 6) t2 = ue_exp(0.5 * (ue_log(A(niy0[f], ix, iy)) + ue_log(A(niy1[f], ix, iy))));
            else {
              double t0 = (A(niy0[f], ix, iy) * A(gy, ix, iy) + A(niy1[f], ix, iy) * A(gy, ix, iy + 1)) / (A(gy, ix, iy) + A(gy, ix, iy + 1));
              double t1 = (A(gy, ix, iy) + A(gy, ix, iy + 1)) * A(niy0[f], ix, iy) * A(niy1[f], ix, iy) / (cutlo + A(niy0[f], ix, iy) * A(gy, ix, iy + 1) + A(niy1[f], ix, iy) * A(gy, ix, iy));
              t2 = ((A(niy0[f], ix, iy) - A(niy1[f], ix, iy)) * A(vy[f], ix, iy) >= 0.) ? t0 : t1;
            }
            A(fniy[f], ix, iy) = cnfy * A(vy[f], ix, iy) * A(sy, ix, iy) * t2;
            if (A(vy[f], ix, iy) * (A(ni[f], ix, iy) - A(ni[f], ix, iy + 1)) < 0.) {
              double r1 = nlimiy[f] / A(ni[f], ix, iy + 1), r2 = nlimiy[f] / A(ni[f], ix, iy);
              A(fniy[f], ix, iy) = A(fniy[f], ix, iy) / (1 + r1 * r1 + r2 * r2);
            }
          }
        if (inrow(ny + 1)) FOR1(ix, i4, i8) A(fniy[f], ix, ny + 1) = 0.0;
      }
    }
    for (int f = 0; f < nfsp; ++f) {  // oderhs.m:3321-3339 (4th-order radial diffusion)
      if (fabs(dif4order[f]) > 1.e-50)
        FOR2(iy, w.j2p, w.j5m, ix, i4, i8) {
            int iym1 = mx(iy - 1, 0), iyp1 = mn(iy + 1, ny + 1), iyp2 = mn(iy + 2, ny + 1);
            double dndym1 = (A(ni[f], ix, iy) - A(ni[f], ix, iym1)) * A(gyf, ix, iym1);
            double dndy0 = (A(ni[f], ix, iyp1) - A(ni[f], ix, iy)) * A(gyf, ix, iy);
            double dndyp1 = (A(ni[f], ix, iyp2) - A(ni[f], ix, iyp1)) * A(gyf, ix, iyp1);
            double d2ndy20 = (dndy0 - dndym1) * A(gy, ix, iy), d2ndy2p1 = (dndyp1 - dndy0) * A(gy, ix, iyp1);
            double d3ndy3 = (d2ndy2p1 - d2ndy20) * A(gyf, ix, iy);
            A(fniy[f], ix, iy) = A(fniy[f], ix, iy) + dif4order[f] * d3ndy3 * A(sy, ix, iy) / (A(gyf, ix, iy) * A(gyf, ix, iy));
          }
      FOR1(ix, i4, i8)  // oderhs.m:3344-3353 (vycp(,0) = 0; isfniycbozero = 0)
        fniycbo[f][ix] = (A(ni[f], ix, 0) * A(sy, ix, 0)) * ((1 - cfniybbo) * cfybf * A(vycb[f], ix, 0));
    }
    // particle balance (oderhs.m:3407-3456)
    for (int f = 0; f < nfsp; ++f) {
      FOR2(iy, j2, j5, ix, i2, i5)
          if (IDXN(f, ix, iy) >= 0)
            A(resco[f], ix, iy) = A(snic[f], ix, iy) + A(sniv[f], ix, iy) * A(ni[f], ix, iy) + 0. + cfneut * cfneutsor_ni * cnsor * A(psor[f], ix, iy) +
                                  cfneut * cfneutsor_ni * cnsor * A(psorxr[f], ix, iy) + cfneut * cfneutsor_ni * cnsor * A(psori[f], ix, iy) - 0. + 0.;
      FOR2(iy, j2, j5, ix, i2, i5)
          if (IDXN(f, ix, iy) >= 0) {
            int ix1 = IXM1(ix, iy);
            if (zi[f] != 0) A(resco[f], ix, iy) = A(resco[f], ix, iy) - ((A(fnix[f], ix, iy) - A(fnix[f], ix1, iy)) + fluxfacy * (A(fniy[f], ix, iy) - A(fniy[f], ix, iy - 1)));
            else A(resco[f], ix, iy) = A(resco[f], ix, iy) - cfneutdiv * cfneutdiv_fng * ((A(fnix[f], ix, iy) - A(fnix[f], ix1, iy)) + fluxfacy * (A(fniy[f], ix, iy) - A(fniy[f], ix, iy - 1)));
          }
    }
    // ---- parallel momentum (oderhs.m:3463-3911), every species with a momentum equation ----
    for (int f = 0; f < nusp; ++f) {
      if (isupon[f] == 0) continue;
      FOR1(iy, j4, j8) { A(flox, 0, iy) = 0.0; A(conx, 0, iy) = 0.0; }
      FOR2(iy, j4, j8, ix, i2, i6) {
          int ix1 = IXM1(ix, iy);
          double uuv = 0.5 * (A(uu[f], ix1, iy) + A(uu[f], ix, iy));
          A(flox, ix, iy) = cmfx * A(nm[f], ix, iy) * uuv * A(vol, ix, iy) * A(gx, ix, iy);
          if (isgxvon == 0) A(conx, ix, iy) = A(visx[f], ix, iy) * A(vol, ix, iy) * A(gx, ix, iy) * A(gx, ix, iy);
          else A(conx, ix, iy) = A(visx[f], ix, iy) * A(vol, ix, iy) * A(gx, ix, iy) * 2 * A(gxf, ix, iy) * A(gxf, ix1, iy) / (A(gxf, ix, iy) + A(gxf, ix1, iy));
        }
      FOR2(iy, j1, j5, ix, i4, i8) {  // oderhs.m:3506-3575
          int ix2 = IXP1(ix, iy), ix4 = IXP1(ix, iy + 1);
          if (iy == iysptrx1 && (ix == ixpt1 || ix == ixpt2)) {
            A(floy, ix, iy) = (cmfy / 2) * A(syv, ix, iy) * (ave(A(nm[f], ix, iy), A(nm[f], ix, iy + 1))) * A(vy[f], ix, iy);
            if (f == 0) A(floy, ix, iy) = A(floy, ix, iy) + (cmfy / 2) * A(syv, ix, iy) * (ave(A(nm[f], ix, iy), A(nm[f], ix, iy + 1))) * 0.;
          } else if (isugfm1side == 1 && zi[f] == 0.) {
            A(floy, ix, iy) = (cmfy / 4) * A(syv, ix, iy) * (ave(A(nm[f], ix, iy), A(nm[f], ix, iy + 1)) + ave(A(nm[f], ix2, iy), A(nm[f], ix4, iy + 1))) * (A(vy[f], ix, iy) + A(vy[f], ix, iy));
          } else {
            A(floy, ix, iy) = (cmfy / 4) * A(syv, ix, iy) * (ave(A(nm[f], ix, iy), A(nm[f], ix, iy + 1)) + ave(A(nm[f], ix2, iy), A(nm[f], ix4, iy + 1))) * (A(vy[f], ix, iy) + A(vy[f], ix2, iy));
            if (f == 0) A(floy, ix, iy) = A(floy, ix, iy) + (cmfy / 4) * A(syv, ix, iy) * (ave(A(nm[f], ix, iy), A(nm[f], ix, iy + 1)) + ave(A(nm[f], ix2, iy), A(nm[f], ix4, iy + 1))) * (0. + 0.);
          }
          if (ishavisy == 1)
            A(cony, ix, iy) = .5 * A(syv, ix, iy) * (ave(A(visy[f], ix, iy) * A(gy, ix, iy), A(visy[f], ix, iy + 1) * A(gy, ix, iy + 1)) +
                                                     ave(A(visy[f], ix2, iy) * A(gy, ix2, iy), A(visy[f], ix4, iy + 1) * A(gy, ix4, iy + 1)));
          else
            A(cony, ix, iy) = .25 * cfaccony * A(syv, ix, iy) * (A(visy[f], ix, iy) * A(gy, ix, iy) + A(visy[f], ix, iy + 1) * A(gy, ix, iy + 1) +
                                                                 A(visy[f], ix2, iy) * A(gy, ix2, iy) + A(visy[f], ix4, iy + 1) * A(gy, ix4, iy + 1));
        }
      fd2tra(w, flox, floy, conx, cony, up[f], fmix[f], fmiy[f], 1, methu);  // oderhs.m:3579
      if (isnonog == 1)  // y-part of the non-orthogonal diffusive momentum flux (oderhs.m:3583-3633)
        FOR2(iy, j2, j5, ix, i2, i5 + 1) {
            const int iy1 = mx(iy - 1, 0);
            const int c = ix + NXS * iy;
            const int ix1 = IXM1(ix, iy), ix3 = IXM1(ix, iy1), ix5 = IXM1(ix, iy + 1);
            const double* u = up[f];
            double grdnv = (fymv[1][c] * A(u, ix, iy1) + fy0v[1][c] * A(u, ix, iy) + fypv[1][c] * A(u, ix, iy + 1) + fymxv[1][c] * A(u, ix3, iy1) + fypxv[1][c] * A(u, ix5, iy + 1) -
                            fymv[0][c] * A(u, ix3, iy1) - fy0v[0][c] * A(u, ix1, iy) - fypv[0][c] * A(u, ix5, iy + 1) - fymxv[0][c] * A(u, ix, iy1) - fypxv[0][c] * A(u, ix, iy + 1)) * 2 /
                           (A(dxnog, ix, iy) + A(dxnog, ix1, iy));
            double gfac = (isgxvon == 0) ? A(gx, ix, iy) : (2 * A(gxf, ix, iy) * A(gxf, ix1, iy) / (A(gxf, ix, iy) + A(gxf, ix1, iy)));
            A(fmixy[f], ix, iy) = cfvisxy[f] * A(visy[f], ix, iy) * (grdnv / ue_cos(0.5 * (A(angfx, ix1, iy) + A(angfx, ix, iy))) - (A(u, ix, iy) - A(u, ix1, iy)) * gfac) * 0.5 * (A(sx, ix1, iy) + A(sx, ix, iy));
            if (f == 1) {
              double t0 = mx(A(tg, ix, iy), tgmin * ev);
              double vtn = sqrt(t0 / mg_[0]);
              double qfl = flalfvgxya[ix] * 0.5 * (A(sx, ix, iy) + A(sx, ix1, iy)) * vtn * vtn * A(nm[f], ix, iy) + cutlo;
              A(fmixy[f], ix, iy) = A(fmixy[f], ix, iy) / sqrt(1 + sq(A(fmixy[f], ix, iy) / qfl));
            }
          }
      FOR2(iy, j2, j5, ix, i2, i5) {  // sources and pressure gradient (oderhs.m:3746-3836)
          const int ix2 = IXP1(ix, iy);
          if (zi[f] != 0) {
            double dp1 = cngmom[f] * (1 / fac2sp) * (A(ng, ix2, iy) * A(tg, ix2, iy) - A(ng, ix, iy) * A(tg, ix, iy));
            A(resmo[f], ix, iy) = 0.;
            A(resmo[f], ix, iy) = A(smoc[f], ix, iy) + A(smov[f], ix, iy) * A(up[f], ix, iy) - cfneut * cfneutsor_mi * A(sx, ix, iy) * A(rrv, ix, iy) * dp1 -
                                  cfneut * cfneutsor_mi * cmwall[f] * 0.5 * (A(ng, ix, iy) + A(ng, ix2, iy)) * mi[f] * A(up[f], ix, iy) * 0.5 * (A(nucx, ix, iy) + A(nucx, ix2, iy)) * A(volv, ix, iy) +
                                  0. + cfmsor * (0. + 0.) + 0. + 0. + 0.;
          }
          if (isupgon == 1) {
            const int g = iigsp;
            if (f == 0) {
              A(resmo[f], ix, iy) = A(resmo[f], ix, iy) +
                                    cfneut * cfneutsor_mi * cfupcx * 0.25 * A(volv, ix, iy) * (A(nucx, ix, iy) + A(nucx, ix2, iy)) * (A(nm[g], ix, iy) + A(nm[g], ix2, iy)) * (A(up[g], ix, iy) - A(up[0], ix, iy)) +
                                    cfneut * cfneutsor_mi * 0.25 * A(volv, ix, iy) *
                                        ((A(nuiz, ix, iy) + A(nuiz, ix2, iy)) * (A(nm[g], ix, iy) + A(nm[g], ix2, iy)) * A(up[g], ix, iy) -
                                         (A(nurc, ix, iy) + A(nurc, ix2, iy)) * (A(nm[0], ix, iy) + A(nm[0], ix2, iy)) * A(up[0], ix, iy));
            } else if (f == g) {
              A(resmo[g], ix, iy) = -0. - A(sx, ix, iy) * A(rrv, ix, iy) * cpgx * (cftiexclg * (A(ni[g], ix2, iy) * A(ti, ix2, iy) - A(ni[g], ix, iy) * A(ti, ix, iy)) +
                                                                                   (1.0 - cftiexclg) * (A(ni[g], ix2, iy) * A(tg, ix2, iy) - A(ni[g], ix, iy) * A(tg, ix, iy))) -
                                    cfupcx * 0.25 * A(volv, ix, iy) * (A(nucx, ix, iy) + A(nucx, ix2, iy)) * (A(nm[g], ix, iy) + A(nm[g], ix2, iy)) * (A(up[g], ix, iy) - A(up[0], ix, iy)) -
                                    0.25 * A(volv, ix, iy) * ((A(nuiz, ix, iy) + A(nuiz, ix2, iy)) * (A(nm[g], ix, iy) + A(nm[g], ix2, iy)) * A(up[g], ix, iy) -
                                                              (A(nurc, ix, iy) + A(nurc, ix2, iy)) * (A(nm[0], ix, iy) + A(nm[0], ix2, iy)) * A(up[0], ix, iy));
            }
          }
        }
      if (isnonog == 1)
        FOR2(iy, j2, j5, ix, i2, i5) {
            const int ix2 = IXP1(ix, iy);
            if (zi[f] > 1.e-20) A(resmo[f], ix, iy) = A(resmo[f], ix, iy) + (A(fmixy[f], ix2, iy) - A(fmixy[f], ix, iy));
            else A(resmo[f], ix, iy) = A(resmo[f], ix, iy) + cfneutdiv * cfneutdiv_fmg * (A(fmixy[f], ix2, iy) - A(fmixy[f], ix, iy));
          }
      FOR2(iy, j2, j5, ix, i2, i5) {
          const int ix2 = IXP1(ix, iy);
          if (zi[f] > 1.e-20) A(resmo[f], ix, iy) = A(resmo[f], ix, iy) - (A(fmix[f], ix2, iy) - A(fmix[f], ix, iy) + fluxfacy * (A(fmiy[f], ix, iy) - A(fmiy[f], ix, iy - 1)));
          else A(resmo[f], ix, iy) = A(resmo[f], ix, iy) - cfneutdiv * cfneutdiv_fmg * (A(fmix[f], ix2, iy) - A(fmix[f], ix, iy) + fluxfacy * (A(fmiy[f], ix, iy) - A(fmiy[f], ix, iy - 1)));
        }
    }

    // ---- energy equations: convective / conductive coefficients (oderhs.m:3923-4261) ----
    FOR2(iy, j1, j6, ix, i1, i6) {
        A(floxe, ix, iy) = 0.; A(floxi, ix, iy) = 0.; A(floye, ix, iy) = 0.; A(floyi, ix, iy) = 0.;
        feiycbo[ix] = 0.; feeycbo[ix] = 0.; A(w0, ix, iy) = 0.; A(w1, ix, iy) = 0.;
      }
    for (int f = 0; f < nusp; ++f) FOR1(c_, rowlo * NXS, (rowhi + 1) * NXS - 1) wvh[f][c_] = 0.;
    FOR2(iy, j4, j8, ix, i1, i5) {
        int ix2 = IXP1(ix, iy);
        double t0 = mx(A(te, ix, iy), temin * ev), t1 = mx(A(te, ix2, iy), temin * ev);
        double vt0 = sqrt(t0 / me), vt1 = sqrt(t1 / me);
        double wallfac = 1.;
        if ((ix == ixlb || ix == ixrb) && (isplflxl == 0)) wallfac = flalfepl / flalfe;
        double qfl = wallfac * flalfe * A(sx, ix, iy) * A(rrv, ix, iy) * (A(ne, ix, iy) * vt0 * t0 + A(ne, ix2, iy) * vt1 * t1) / 2;
        double csh = A(sx, ix, iy) * A(hcxe, ix, iy) * A(gxf, ix, iy);
        double lxtec = 0.5 * (A(te, ix, iy) + A(te, ix2, iy)) / (fabs(A(te, ix, iy) - A(te, ix2, iy)) * A(gxf, ix, iy) + 100. * cutlo);
        double qsh = csh * (A(te, ix, iy) - A(te, ix2, iy)) * (1. + lxtec / lxtemax);
        double qr = (1 - isflxlde) * fabs(qsh / qfl);
        A(conxe, ix, iy) = (1 - isflxlde) * csh / ((1 + qr) * (1 + qr)) + isflxlde * csh / ue_pow(1 + ue_pow(fabs(qsh / qfl), flgam), 1 / flgam);
        A(floxe, ix, iy) = A(floxe, ix, iy) + (sgn(qr * qr, qsh) / ((1 + qr) * (1 + qr))) * flalfea[ix] * A(sx, ix, iy) * (A(ne, ix, iy) * A(rr, ix, iy) * vt0 + A(ne, ix2, iy) * A(rr, ix2, iy) * vt1) / 2;
        if (isflxldi != 2) {
          t0 = mx(A(ti, ix, iy), temin * ev); t1 = mx(A(ti, ix2, iy), temin * ev);
          vt0 = sqrt(t0 / mi[0]); vt1 = sqrt(t1 / mi[0]);
          wallfac = 1.;
          if ((ix == ixlb || ix == ixrb) && (isplflxl == 0)) wallfac = flalfipl / flalfi;
          qfl = wallfac * flalfia[ix] * A(sx, ix, iy) * A(rrv, ix, iy) * (A(ne, ix, iy) * vt0 * t0 + A(ne, ix2, iy) * vt1 * t1) / 2;
          csh = A(sx, ix, iy) * A(hcxi, ix, iy) * A(gxf, ix, iy);
          double lxtic = 0.5 * (A(ti, ix, iy) + A(ti, ix2, iy)) / (fabs(A(ti, ix, iy) - A(ti, ix2, iy)) * A(gxf, ix, iy) + 100. * cutlo);
          qsh = csh * (A(ti, ix, iy) - A(ti, ix2, iy)) * (1. + lxtic / lxtimax);
          qr = (1 - isflxldi) * fabs(qsh / qfl);
          A(conxi, ix, iy) = (1 - isflxldi) * csh / ((1 + qr) * (1 + qr)) + isflxldi * csh / ue_pow(1 + ue_pow(fabs(qsh / qfl), flgam), 1 / flgam);
          A(floxi, ix, iy) = A(floxi, ix, iy) + (sgn(qr * qr, qsh) / ((1 + qr) * (1 + qr))) * flalfia[ix] * A(sx, ix, iy) * (A(ne, ix, iy) * A(rr, ix, iy) * vt0 + A(ne, ix2, iy) * A(rr, ix2, iy) * vt1) / 2;
        } else A(conxi, ix, iy) = A(sx, ix, iy) * A(hcxi, ix, iy) * A(gxf, ix, iy);
      }
    FOR1(iy, j4, j8) { A(conxe, nx + 1, iy) = 0; A(conxi, nx + 1, iy) = 0; }
    FOR2(iy, j1, j5, ix, i4, i8) {
        A(conye, ix, iy) = A(sy, ix, iy) * A(hcye, ix, iy) / A(dynog, ix, iy);
        A(conyi, ix, iy) = A(sy, ix, iy) * A(hcyi, ix, iy) / A(dynog, ix, iy);
      }
    if (inrow(ny + 1)) FOR1(ix, i1, i6) { A(conye, ix, ny + 1) = 0.0; A(conyi, ix, ny + 1) = 0.0; }
    FOR2(iy, j4, j8, ix, i1, i5) {  // oderhs.m:4024-4036
        int ix1 = IXP1(ix, iy);
        double ltmax = mn(fabs(A(te, ix, iy) / (A(rrv, ix, iy) * A(gtex, ix, iy) + cutlo)), A(lcone, ix, iy));
        double lmfpe = 2e16 * ((A(te, ix, iy) / ev) * (A(te, ix, iy) / ev)) / A(ne, ix, iy);
        double flxlimf = flalftf * ltmax / (flalftf * ltmax + lmfpe);
        A(floxe, ix, iy) = A(floxe, ix, iy) + cfcvte * 1.25 * (A(ne, ix, iy) + A(ne, ix1, iy)) * A(vex, ix, iy) * A(sx, ix, iy) - cthe * flxlimf * cfjhf * A(fqp, ix, iy) / ev;
      }
    FOR1(iy, j4, j8) { A(floxe, nx + 1, iy) = 0.0; }
    for (int f = 0; f < nfsp; ++f) {  // oderhs.m:4038-4072
      if (isupgon == 1 && f == iigsp) {
        FOR2(iy, j4, j8, ix, i1, i5) A(floxi, ix, iy) = A(floxi, ix, iy) + cftiexclg * cfcvti * 2.5 * cfneut * cfneutsor_ei * A(fnix[f], ix, iy);
        FOR1(iy, j4, j8) { if (A(fnix[f], ixlb, iy) > 0.) A(floxi, ixlb, iy) = A(floxi, ixlb, iy) - (1. - cfloxiplt) * cftiexclg * cfcvti * 2.5 * cfneut * cfneutsor_ei * A(fnix[f], ixlb, iy); if (A(fnix[f], ixrb, iy) < 0.) A(floxi, ixrb, iy) = A(floxi, ixrb, iy) - (1. - cfloxiplt) * cftiexclg * cfcvti * 2.5 * cfneut * cfneutsor_ei * A(fnix[f], ixrb, iy); A(floxi, ixrb + 1, iy) = 0.0; }
      } else {
        FOR2(iy, j4, j8, ix, i1, i5) A(floxi, ix, iy) = A(floxi, ix, iy) + cfcvti * 2.5 * A(fnix[f], ix, iy);
        FOR1(iy, j4, j8) { A(floxi, nx + 1, iy) = 0.0; }
      }
    }
    FOR2(iy, j1, j5, ix, i4, i8) {  // oderhs.m:4078-4092; vyte_use, vyte_cft, cfybf = 0
        A(floye, ix, iy) = A(floye, ix, iy) + (cfloye / 2.) * (A(ney0, ix, iy) + A(ney1, ix, iy)) * A(vey, ix, iy) * A(sy, ix, iy) + (0. + 0.) * 0.5 * A(sy, ix, iy) * (A(ney0, ix, iy) + A(ney1, ix, iy));
        if (iy == 0) feeycbo[ix] = cfloye * (A(ne, ix, 0) * A(te, ix, 0) * A(sy, ix, 0)) * ((1 - cfeeybbo) * cfybf * A(veycb, ix, 0));  // veycp(,0) = 0
      }
    for (int f = 0; f < nfsp; ++f) {  // oderhs.m:4093-4128
      if (isupgon == 1 && f == iigsp) {
        FOR2(iy, j1, j5, ix, i4, i8) A(floyi, ix, iy) = A(floyi, ix, iy) + cftiexclg * cfneut * cfneutsor_ei * 2.5 * A(fniy[f], ix, iy);
        FOR1(ix, i4, i8) {
          if (inrow(ny) && matwallo[ix] > 0 && recycwot[ix] > 0.) {
            double fniy_recy = mx(recycwot[ix] * fac2sp * A(fniy[0], ix, ny), 0.);
            A(floyi, ix, ny) = A(floyi, ix, ny) + cftiexclg * cfneut * cfneutsor_ei * 2.5 * (1. - cfloygwall) * fniy_recy;
          }
          if (inrow(0) && matwalli[ix] > 0 && recycwit[ix] > 0.) {
            double fniy_recy = mn(recycwit[ix] * fac2sp * A(fniy[0], ix, 0), 0.);
            A(floyi, ix, 0) = A(floyi, ix, 0) + cftiexclg * cfneut * cfneutsor_ei * 2.5 * (1. - cfloygwall) * fniy_recy;
          }
        }
      } else
        FOR2(iy, j1, j5, ix, i4, i8) {
            A(floyi, ix, iy) = A(floyi, ix, iy) + cfloyi * A(fniy[f], ix, iy) + (0. + 0.) * 0.5 * A(sy, ix, iy) * (A(niy0[f], ix, iy) + A(niy1[f], ix, iy));
            if (iy == 0) feiycbo[ix] = feiycbo[ix] + cfloyi * fniycbo[f][ix] * A(ti, ix, 0);
          }
    }
    if (fabs(cfbgt) > 0) {  // B x grad(T) heat flows (oderhs.m:4131-4229); the plate terms of cfeexdbo / cfeixdbo are refused in init
      for (int f = 0; f < nfsp; ++f) {
      FOR2(iy, j4, j8, ix, i1, i5) {
          const int iy1 = mx(0, iy - 1), ix1 = IXP1(ix, iy);
          if (iy == 0 || iy == ny + 1) continue;
          const double temp1 = 4.0 * (A(tiv, ix, iy) - A(tiv, ix, iy1)) * A(gyc, ix, iy);
          if (zi[f] > 1.e-10) A(floxi, ix, iy) = A(floxi, ix, iy) + cfbgt * ((5 * A(sx, ix, iy) / (32 * qe * zi[f])) * (A(ni[f], ix, iy) + A(ni[f], ix1, iy)) * (A(rbfbt2, ix, iy) + A(rbfbt2, ix1, iy)) * temp1);
        }
      FOR2(iy, j1, j5, ix, i4, i8) {
          const int ix3 = IXM1(ix, iy);
          if (ix == ixlb || ix == ixrb + 1) continue;
          const double temp1 = 4.0 * (A(tiv, ix, iy) - A(tiv, ix3, iy)) * A(gxc, ix, iy);
          if (zi[f] > 1.e-10) A(floyi, ix, iy) = A(floyi, ix, iy) - cfbgt * (5 * A(sy, ix, iy) / (32 * qe * zi[f])) * (A(ni[f], ix, iy) + A(ni[f], ix, iy + 1)) * (A(rbfbt2, ix, iy) + A(rbfbt2, ix, iy + 1)) * temp1;
        }
      }
      FOR2(iy, j4, j8, ix, i1, i5) {
          const int iy1 = mx(0, iy - 1), ix1 = IXP1(ix, iy);
          if (iy == 0 || iy == ny + 1) continue;
          const double temp1 = 4.0 * (A(tev, ix, iy) - A(tev, ix, iy1)) * A(gyc, ix, iy);
          A(floxe, ix, iy) = A(floxe, ix, iy) - cfbgt * ((5 * A(sx, ix, iy) / (32 * qe)) * (A(ne, ix, iy) + A(ne, ix1, iy)) * (A(rbfbt2, ix, iy) + A(rbfbt2, ix1, iy)) * temp1);
        }
      FOR2(iy, j1, j5, ix, i4, i8) {
          const int ix3 = IXM1(ix, iy);
          if (ix == ixlb || ix == ixrb + 1) continue;
          const double temp1 = 4.0 * (A(tev, ix, iy) - A(tev, ix3, iy)) * A(gxc, ix, iy);
          A(floye, ix, iy) = A(floye, ix, iy) + cfbgt * (5 * A(sy, ix, iy) / (32 * qe)) * (A(ne, ix, iy) + A(ne, ix, iy + 1)) * (A(rbfbt2, ix, iy) + A(rbfbt2, ix, iy + 1)) * temp1;
        }
    }
    FOR2(iy, j4, j8, ix, i1, i5) A(floxi, ix, iy) = A(floxi, ix, iy) + cftiexclg * cfneut * cfneutsor_ei * cngtgx[0] * cfcvti * 2.5 * A(fngx, ix, iy);  // oderhs.m:4234-4240
    FOR1(iy, j4, j8) { A(floxi, nx + 1, iy) = 0.0; }
    FOR2(iy, j1, j5, ix, i4, i8) A(floyi, ix, iy) = A(floyi, ix, iy) + cftiexclg * cfneut * cfneutsor_ei * cngtgy[0] * 2.5 * A(fngy, ix, iy);
    if (isteon == 1) fd2tra(w, floxe, floye, conxe, conye, te, feex, feey, 0, methe);  // oderhs.m:4256
    if (istion == 1) fd2tra(w, floxi, floyi, conxi, conyi, ti, feix, feiy, 0, methi);  // oderhs.m:4260
    if (fabs(kye4order) > 1.e-50 || fabs(kyi4order) > 1.e-50)  // oderhs.m:4263-4291
      FOR2(iy, w.j2p, w.j5m, ix, i4, i8) {
          int iym1 = mx(iy - 1, 0), iyp1 = mn(iy + 1, ny + 1), iyp2 = mn(iy + 2, ny + 1);
          auto d3 = [&](const double* t) {
            double dm1 = (A(t, ix, iy) - A(t, ix, iym1)) * A(gyf, ix, iym1), d0 = (A(t, ix, iyp1) - A(t, ix, iy)) * A(gyf, ix, iy), dp1 = (A(t, ix, iyp2) - A(t, ix, iyp1)) * A(gyf, ix, iyp1);
            double d20 = (d0 - dm1) * A(gy, ix, iy), d2p1 = (dp1 - d0) * A(gy, ix, iyp1);
            return (d2p1 - d20) * A(gyf, ix, iy);
          };
          A(feey, ix, iy) = A(feey, ix, iy) + kye4order * d3(te) * A(ney1, ix, iy) * A(sy, ix, iy) / (A(gyf, ix, iy) * A(gyf, ix, iy));
          A(feiy, ix, iy) = A(feiy, ix, iy) + kyi4order * d3(ti) * A(niy1[0], ix, iy) * A(sy, ix, iy) / (A(gyf, ix, iy) * A(gyf, ix, iy));
        }
    FOR2(iy, j2, j5, ix, i2, i5) {  // oderhs.m:4300-4313 (pwrsore/pwrsori/nuvl zero)
        A(resee, ix, iy) = A(seec, ix, iy) + A(seev, ix, iy) * A(te, ix, iy) + 0. + 0. - 0.;
        A(resei, ix, iy) = A(seic, ix, iy) + A(seiv, ix, iy) * A(ti, ix, iy) + 0. + 0. - 0.;
      }
    if (isnonog == 1) {  // y-part of the non-orthogonal diffusive heat fluxes (oderhs.m:4318-4393)
      FOR2(iy, j1, j6, ix, i1, i6) {
          if (iy > ny) continue;
          const int iy1 = mx(iy - 1, 0);
          const int ix2 = IXP1(ix, iy), ix4 = IXP1(ix, iy1);
          double grdnv = grdnv_y(te, ix, iy, 1) / A(dxnog, ix, iy);
          A(feexy, ix, iy) = ue_exp(0.5 * (ue_log(A(te, ix2, iy)) + ue_log(A(te, ix, iy)))) * (fcdif * kye + 0.) * 0.5 * (A(ne, ix2, iy) + A(ne, ix, iy)) *
                             (grdnv / ue_cos(A(angfx, ix, iy)) - (ue_log(A(te, ix2, iy)) - ue_log(A(te, ix, iy))) * A(gxf, ix, iy)) * A(sx, ix, iy);
          grdnv = grdnv_y(ti, ix, iy, 1) / A(dxnog, ix, iy);
          A(feixy, ix, iy) = ue_exp(0.5 * (ue_log(A(ti, ix2, iy)) + ue_log(A(ti, ix, iy)))) *
                             ((fcdif * kyi + 0.) * 0.5 * (A(nit, ix2, iy) + A(nit, ix, iy)) +
                              cftiexclg * cfneut * cfneutsor_ei * 0.25 * (A(hcyn, ix, iy) + A(hcyn, ix, iy1) + A(hcyn, ix2, iy) + A(hcyn, ix4, iy1))) *
                             (grdnv / ue_cos(A(angfx, ix, iy)) - (ue_log(A(ti, ix2, iy)) - ue_log(A(ti, ix, iy))) * A(gxf, ix, iy)) * A(sx, ix, iy);
          double t0 = mx(A(ti, ix, iy), temin * ev), t1 = mx(A(ti, ix2, iy), temin * ev);
          double vttn = t0 * sqrt(t0 / mi[0]), vttp = t1 * sqrt(t1 / mi[0]);
          double qfl = flalftxy * (cftiexclg * 0.125 + (1. - cftiexclg) * 0.25) * A(sx, ix, iy) * (vttn + vttp) *
                       (A(ni[0], ix, iy) + cftiexclg * A(ng, ix, iy) + A(ni[0], ix2, iy) + cftiexclg * A(ng, ix2, iy));
          A(feixy, ix, iy) = A(feixy, ix, iy) / sqrt(1. + sq(A(feixy, ix, iy) / qfl));
        }
      FOR2(iy, j4, j8, ix, i1, i5) { A(feex, ix, iy) = A(feex, ix, iy) - A(feexy, ix, iy); A(feix, ix, iy) = A(feix, ix, iy) - A(feixy, ix, iy); }
    }
    FOR2(iy, j2, j5, ix, i2, i5) {  // oderhs.m:4439-4478
        int ix1 = IXM1(ix, iy);
        A(resee, ix, iy) = A(resee, ix, iy) - (A(feex, ix, iy) - A(feex, ix1, iy) + fluxfacy * (A(feey, ix, iy) - A(feey, ix, iy - 1)));
        A(resei, ix, iy) = A(resei, ix, iy) - (A(feix, ix, iy) - A(feix, ix1, iy) + fluxfacy * (A(feiy, ix, iy) - A(feiy, ix, iy - 1)));
      }
    // hydrogen radiation / ionisation energy sink (oderhs.m:4484-4555)
    FOR2(iy, w.iys1, w.iyf6, ix, w.ixs1, w.ixf6) {
        double ne_sgvi = A(ne, ix, iy);
        if (ifxnsgi == 1) ne_sgvi = cne_sgvi;
        A(erliz, ix, iy) = chradi * erl1(A(te, ix, iy), ne_sgvi) * (A(ng, ix, iy) - ngbackg_[0] * (0.9 + 0.1 * powi(ngbackg_[0] / A(ng, ix, iy), ingb))) * A(vol, ix, iy);
        if (isrecmon != 0) A(erlrc, ix, iy) = chradr * erl2(A(te, ix, iy), ne_sgvi) * fac2sp * A(ni[0], ix, iy) * A(vol, ix, iy);
        if (icnuiz <= 1 && A(psor[0], ix, iy) != 0.) A(eeli, ix, iy) = 13.6 * ev + A(erliz, ix, iy) / (fac2sp * A(psor[0], ix, iy));
        A(pradhyd, ix, iy) = ((A(eeli, ix, iy) - ebind * ev) * A(psor[0], ix, iy) + A(erlrc, ix, iy)) / A(vol, ix, iy);
      }
    FOR2(iy, w.iys1, w.iyf6, ix, w.ixs1, w.ixf6) {
        A(vsoreec, ix, iy) = -cfneut * cfneutsor_ee * cnsor * 13.6 * ev * fac2sp * A(psorc[0], ix, iy) + cfneut * cfneutsor_ee * cnsor * 13.6 * ev * fac2sp * A(psorrgc, ix, iy) -
                             cfneut * cfneutsor_ee * cnsor * A(erliz, ix, iy) - cfneut * cfneutsor_ee * cnsor * A(erlrc, ix, iy) -
                             cfneut * cfneutsor_ee * cnsor * ediss * ev * (0.5 * A(psordis, ix, iy));
        A(vsoree, ix, iy) = A(vsoreec, ix, iy);  // iseesorave = 0
      }
    FOR2(iy, j2, j5, ix, i2, i5) {  // oderhs.m:4589-4640
        int ix1 = IXM1(ix, iy);
        A(w0, ix, iy) = A(vol, ix, iy) * A(eqp, ix, iy) * (A(te, ix, iy) - A(ti, ix, iy));
        A(resee, ix, iy) = A(resee, ix, iy) - A(w0, ix, iy) + A(vsoree, ix, iy);
        if (isupgon == 1) {
          const int g = iigsp;
          double t1 = 0.5 * (A(up[0], ix, iy) + A(up[0], ix1, iy));
          double t2 = 0.5 * (A(up[g], ix, iy) + A(up[g], ix1, iy));
          double temp3 = cfnidhgy * 0.25 * (A(vy[g], ix, iy) + A(vy[g], ix1, iy)) * (A(vy[g], ix, iy) + A(vy[g], ix1, iy));
          double temp4 = cfnidhg2 * 0.25 * (A(v2[g], ix, iy) + A(v2[g], ix1, iy)) * (A(v2[g], ix, iy) + A(v2[g], ix1, iy));
          double tv = cfticx * A(nucx, ix, iy) * A(ng, ix, iy) * A(vol, ix, iy);
          double t0 = 1.5 * (A(tg, ix, iy) * (A(psor[0], ix, iy) + tv) - A(ti, ix, iy) * (A(psorrg, ix, iy) + tv));
          A(resei, ix, iy) = A(resei, ix, iy) + A(w0, ix, iy) +
                             cfneut * cfneutsor_ei * cfnidh * 0.5 * mi[0] * ((t1 - t2) * (t1 - t2) + temp3 + temp4) * (A(psor[0], ix, iy) + cftiexclg * A(psorrg, ix, iy) + tv + cftiexclg * tv) +
                             (1.0 - cftiexclg) * t0 +
                             cftiexclg * cfneut * cfneutsor_ei * cnsor * (eion * ev + cfnidhdis * 0.5 * mg_[0] * (t2 * t2 + temp3 + temp4)) * A(psordis, ix, iy) +
                             cfnidh2 * (-mi[0] * t1 * t2 * (A(psor[0], ix, iy) + tv) + 0.5 * mi[0] * t1 * t1 * (A(psor[0], ix, iy) + A(psorrg, ix, iy) + 2 * tv));
          A(reseg, ix, iy) = A(reseg, ix, iy) - t0 + 0.5 * mg_[0] * ((t1 - t2) * (t1 - t2) + temp3 + temp4) * (A(psorrg, ix, iy) + tv) +
                             (eion * ev + cfnidh * cfnidhdis * 0.5 * mg_[0] * (t2 * t2 + temp3 + temp4)) * A(psordis, ix, iy) +
                             cfnidh2 * (-mg_[0] * t1 * t2 * (A(psorrg, ix, iy) + tv) + 0.5 * mg_[0] * (t2 * t2 + temp3 + temp4) * (A(psor[0], ix, iy) + A(psorrg, ix, iy) + 2 * tv));  // oderhs.m:4618-4627
        } else {
          double us = A(upi[0], ix, iy) + A(upi[0], ix1, iy);
          A(resei, ix, iy) = A(resei, ix, iy) + A(w0, ix, iy) + cfneut * cfneutsor_ei * ctsor * 1.25e-1 * mi[0] * (us * us) * fac2sp * A(psor[0], ix, iy) +
                             cfneut * cfneutsor_ei * ceisor * cnsor * eion * ev * A(psordis, ix, iy) -
                             cfneut * cfneutsor_ei * ccoldsor * A(ng, ix, iy) * A(nucx, ix, iy) * (1.5 * A(ti, ix, iy) - 0.125 * mi[0] * (us * us) - eion * ev) * A(vol, ix, iy);
        }
      }
    // Joule heating (oderhs.m:4832-4875)
    if (jhswitch > 0) {
      const int iy_min = isnewpot == 1 ? 2 : 1, iy_max = isnewpot == 1 ? ny - 1 : ny;
      FOR2(iy, mx(iy_min, j2), mn(iy_max, j5), ix, i2, i5) {
          const int ix1 = IXM1(ix, iy), ix2 = IXP1(ix, iy);
          if (jhswitch == 1) {
            A(wjdote, ix, iy) = -0.5 * (A(fqp, ix, iy) + A(fq2, ix, iy)) * (A(phi, ix2, iy) + A(phi, ix, iy)) + 0.5 * (A(fqp, ix1, iy) + A(fq2, ix1, iy)) * (A(phi, ix, iy) + A(phi, ix1, iy)) -
                                0.5 * A(fqygp, ix, iy) * (A(phi, ix, iy + 1) + A(phi, ix, iy)) + 0.5 * A(fqygp, ix, iy - 1) * (A(phi, ix, iy) + A(phi, ix, iy - 1));
            A(resee, ix, iy) = A(resee, ix, iy) + A(wjdote, ix, iy) / (1. + cfwjdotelim * powi(A(te, ix, iy) / tebg, iteb));
          } else {
            A(wjdote, ix, iy) = 0.5 * (A(ex, ix1, iy) * A(fqx, ix1, iy) + A(ex, ix, iy) * A(fqx, ix, iy)) / A(gx, ix, iy) + 0.5 * (A(ey, ix, iy) * A(fqy, ix, iy) + A(ey, ix, iy - 1) * A(fqy, ix, iy - 1)) / A(gy, ix, iy);
            A(resee, ix, iy) = A(resee, ix, iy) + A(wjdote, ix, iy);
          }
        }
    }
    // viscous heating (oderhs.m:4879-4930)
    FOR2(iy, j2, j5, ix, i2, i5)
        for (int f = 0; f < nusp; ++f) {
          int ix1 = IXM1(ix, iy), ix2 = IXM1(ix, iy + 1), ix3 = IXM1(ix, iy - 1);
          double thetacc = 0.5 * (A(angfx, ix1, iy) + A(angfx, ix, iy));
          double dupdx = A(gx, ix, iy) * (A(upi[f], ix, iy) - A(upi[f], ix1, iy));
          A(wvh[f], ix, iy) = cfvcsx[f] * cfvisx * ue_cos(thetacc) * A(visx[f], ix, iy) * (dupdx * dupdx);
          double dupdy;
          const int isx = (int)isxptyd[ix + NXS * iy];
          const double* u = upi[f];
          if (isx == 0) dupdy = 0.5 * (A(u, ix, iy) + A(u, ix1, iy) - A(u, ix, iy - 1) - A(u, ix3, iy - 1)) * A(gyf, ix, iy - 1);
          else if (isx == -1) dupdy = 0.5 * (A(u, ix, iy + 1) + A(u, ix2, iy + 1) - A(u, ix, iy) - A(u, ix1, iy)) * A(gyf, ix, iy);
          else if (isx == 1 && isvhyha == 1) {
            double upxavep1 = 0.5 * (A(u, ix, iy + 1) + A(u, ix2, iy + 1)), upxave0 = 0.5 * (A(u, ix, iy) + A(u, ix1, iy)), upxavem1 = 0.5 * (A(u, ix, iy - 1) + A(u, ix3, iy - 1));
            double upf0 = 2. * upxavep1 * upxave0 * (upxavep1 + upxave0) / ((upxavep1 + upxave0) * (upxavep1 + upxave0) + upvhflr * upvhflr);
            double upfm1 = 2. * upxave0 * upxavem1 * (upxave0 + upxavem1) / ((upxave0 + upxavem1) * (upxave0 + upxavem1) + upvhflr * upvhflr);
            dupdy = (upf0 - upfm1) * A(gy, ix, iy);
          } else
            dupdy = 0.25 * ((A(u, ix, iy + 1) + A(u, ix2, iy + 1) - A(u, ix, iy) - A(u, ix1, iy)) * A(gyf, ix, iy) + (A(u, ix, iy) + A(u, ix1, iy) - A(u, ix, iy - 1) - A(u, ix3, iy - 1)) * A(gyf, ix, iy - 1));
          A(wvh[f], ix, iy) = A(wvh[f], ix, iy) + cfvcsy[f] * cfvisy * A(visy[f], ix, iy) * (dupdy * dupdy);
          A(wvh[f], ix, iy) = A(wvh[f], ix, iy) - ue_ksin(thetacc) * cfvcsy[f] * cfvisy * A(visy[f], ix, iy) * dupdx * dupdy;
          if (zi[f] == 0.0 && f == iigsp) { A(resei, ix, iy) = A(resei, ix, iy) + cftiexclg * A(wvh[f], ix, iy) * A(vol, ix, iy); A(reseg, ix, iy) = A(reseg, ix, iy) + A(wvh[f], ix, iy) * A(vol, ix, iy); }
          else A(resei, ix, iy) = A(resei, ix, iy) + A(wvh[f], ix, iy) * A(vol, ix, iy);
        }
    FOR2(iy, w.iys, w.iyf, ix, w.ixs, w.ixf) A(pwribkg, ix, iy) = powi(tibg * ev / A(ti, ix, iy), iteb) * pwribkg_c;  // oderhs.m:4936-4947
    FOR2(iy, j2, j5, ix, i2, i5) A(resei, ix, iy) = A(resei, ix, iy) + A(pwribkg, ix, iy) * A(vol, ix, iy);

    // ---- assemble yldot (oderhs.m:4953-4996) ----
    FOR2(iy, j2, j5, ix, i2, i5) {
        int64_t iv;
        for (int f = 0; f < nisp; ++f) { iv = IDXN(f, ix, iy); if (iv >= 0) yldot[iv] = (1 - ALG(iv)) * A(resco[f], ix, iy) / (A(vol, ix, iy) * n0[f]); }
        for (int f = 0; f < nusp; ++f) {
          iv = IDXU(f, ix, iy);
          if (iv >= 0) { yldot[iv] = (1 - ALG(iv)) * A(resmo[f], ix, iy) / (A(volv, ix, iy) * fnorm[f]); if (ix == ixrb) yldot[iv] = A(resmo[f], ix, iy) / (A(volv, ix, iy) * fnorm[f]); }
        }
        iv = IDXTE(ix, iy); if (iv >= 0) yldot[iv] = (1 - ALG(iv)) * A(resee, ix, iy) / (A(vol, ix, iy) * ennorm);
        iv = IDXTI(ix, iy); if (iv >= 0) yldot[iv] = (1 - ALG(iv)) * A(resei, ix, iy) / (A(vol, ix, iy) * ennorm);
        iv = IDXG(ix, iy); if (iv >= 0) yldot[iv] = (1 - ALG(iv)) * A(resng, ix, iy) / (A(vol, ix, iy) * n0g_[0]);
        iv = IDXTG(ix, iy); if (iv >= 0) yldot[iv] = (1 - ALG(iv)) * A(reseg, ix, iy) / (A(vol, ix, iy) * ennorm);
      }
    if (isphion == 1) poteneq(w, yl, yldot);  // oderhs.m:5007
    rc = bouncon(w, yl, yldot);               // oderhs.m:5009
    return rc;
    // (oderhs.m:5012-5049: the partial restore of source terms matters only between the two pandf1 calls of jac_calc,
    //  where none of those fields is read; the second call restores everything.)
  }

// ---- calc_currents (potencur.m:39-445): fqp, fq2, fqy, fqx without cross-field drift currents ------------------------
// isfqpave = 0, isimpon = 0, rnewpot and the time-derivative current (cfqydt) as inputs allow; cfqybf = cfq2bf = cfjp2 = cfjpy = 0.
double sigma1_, frfqpn_, cffqpsat_, exjbdry_, rnewpot_, cfqyae_, cfqyai_, cfgpijr_, sigbar0_, r0slab_, dx0_;
int nfqya0core_, nfqya0pf_, nfqya0ow_;
const double *b_c, *rm_c;

HD void calc_currents(const Win& w) {
  const int i1 = w.i1, i5 = w.i5, i6 = w.i6;
  const int j1p = w.j1p, j5p = w.j5p, j6p = w.j6p;
  FOR2(iy, j1p, j6p, ix, i1, i5) {
      const int ix1 = IXP1(ix, iy);
      double t0 = mx(A(te, ix1, iy), temin * ev), t1 = mx(A(te, ix, iy), temin * ev);
      double zfac0 = 1. / (A(zeff, ix1, iy) * (1.193 - 0.2205 * A(zeff, ix1, iy) + 0.0275 * sq(A(zeff, ix1, iy))));
      double zfac1 = 1. / (A(zeff, ix, iy) * (1.193 - 0.2205 * A(zeff, ix, iy) + 0.0275 * sq(A(zeff, ix, iy))));
      double zfac = (zfac0 * A(gx, ix1, iy) + zfac1 * A(gx, ix, iy)) / (A(gx, ix1, iy) + A(gx, ix, iy));
      double nbarx = (A(ne, ix1, iy) * A(gx, ix1, iy) + A(ne, ix, iy) * A(gx, ix, iy)) / (A(gx, ix1, iy) + A(gx, ix, iy));
      double sigbarx = zfac * cfsigm * sigma1_ * (A(rr, ix1, iy) * ue_pow(t0, 1.5) * A(gx, ix1, iy) + A(rr, ix, iy) * ue_pow(t1, 1.5) * A(gx, ix, iy)) /
                       ((A(gx, ix1, iy) + A(gx, ix, iy)) * ue_pow(ev, 1.5));
      if (isfqpave != 0) {  // simple averages (potencur.m:106-111)
        zfac = 0.5 * (zfac0 + zfac1);
        nbarx = 0.5 * (A(ne, ix1, iy) + A(ne, ix, iy));
        sigbarx = zfac * cfsigm * sigma1_ * A(rrv, ix, iy) * ue_pow(0.5 * (t0 + t1) / ev, 1.5);
      }
      A(netap, ix, iy) = nbarx / sigbarx;
      A(fqp, ix, iy) = (A(rrv, ix, iy) * A(sx, ix, iy) * sigbarx * A(gxf, ix, iy) / qe) *
                       ((A(pre, ix1, iy) - A(pre, ix, iy)) / nbarx - qe * (A(phi, ix1, iy) - A(phi, ix, iy)) + qe * (0. - 0.) + 0. / (A(rrv, ix, iy) * nbarx) + cthe * (A(te, ix1, iy) - A(te, ix, iy)));
      const int ixl = ixlb, ixlp1 = ixlb + 1, ixlp2 = ixlb + 2, ixr = ixrb + 1, ixrm1 = ixrb, ixrm2 = ixrb - 1;
      if (ix == ixl) {
        double fqp_old = A(fqp, ix, iy);
        nbarx = A(ne, ixlp1, iy);
        sigbarx = zfac * cfsigm * sigma1_ * A(rrv, ixlp1, iy) * ue_pow(A(te, ixlp1, iy) / ev, 1.5);
        A(fqp, ix, iy) = (A(rrv, ixlp1, iy) * A(sx, ixlp1, iy) * sigbarx * A(gxf, ixlp1, iy) / qe) *
                         ((A(pre, ixlp2, iy) - A(pre, ixlp1, iy)) / nbarx - qe * (A(phi, ixlp1, iy) - A(phi, ixl, iy)) * A(gxf, ixl, iy) / A(gxf, ixlp1, iy) + cthe * (A(te, ixlp2, iy) - A(te, ixlp1, iy)));
        A(fqp, ix, iy) = (1. - frfqpn_) * fqp_old + frfqpn_ * A(fqp, ix, iy);
        fqpsatlb[iy] = -qe * isfdiax * (A(ne, ixl, iy) * A(v2ce[0], ixl, iy) * A(rbfbt, ixl, iy) * A(sx, ixl, iy) + fdiaxlb[iy]);
        for (int f = 0; f < nusp; ++f) fqpsatlb[iy] = fqpsatlb[iy] - qe * zi[f] * A(ni[f], ixl, iy) * A(up[f], ixl, iy) * A(sx, ixl, iy) * A(rrv, ixl, iy);
        if (A(fqp, ixl, iy) < 0.) {
          double fp1 = A(fqp, ixl, iy), fp2 = cffqpsat_ * fqpsatlb[iy];
          A(fqp, ixl, iy) = -ue_pow(ue_pow(fabs(fp1 * fp2), exjbdry_) / (ue_pow(fabs(fp1), exjbdry_) + ue_pow(fabs(fp2), exjbdry_)), 1 / exjbdry_);
        }
      } else if (ix == ixrm1) {
        double fqp_old = A(fqp, ix, iy);
        nbarx = A(ne, ixrm1, iy);
        sigbarx = zfac * cfsigm * sigma1_ * A(rrv, ixrm2, iy) * ue_pow(A(te, ixrm1, iy) / ev, 1.5);
        A(fqp, ix, iy) = (A(rrv, ixrm2, iy) * A(sx, ixrm2, iy) * sigbarx * A(gxf, ixrm2, iy) / qe) *
                         ((A(pre, ixrm1, iy) - A(pre, ixrm2, iy)) / nbarx - qe * (A(phi, ixr, iy) - A(phi, ixrm1, iy)) * A(gxf, ixrm1, iy) / A(gxf, ixrm2, iy) + cthe * (A(te, ixrm1, iy) - A(te, ixrm2, iy)));
        A(fqp, ix, iy) = (1. - frfqpn_) * fqp_old + frfqpn_ * A(fqp, ix, iy);
        fqpsatrb[iy] = qe * isfdiax * (A(ne, ixr, iy) * A(v2ce[0], ixrm1, iy) * A(rbfbt, ixr, iy) * A(sx, ixrm1, iy) + fdiaxrb[iy]);
        for (int f = 0; f < nusp; ++f) fqpsatrb[iy] = fqpsatrb[iy] + qe * zi[f] * A(ni[f], ixr, iy) * A(up[f], ixrm1, iy) * A(sx, ixrm1, iy) * A(rrv, ixrm1, iy);
        if (A(fqp, ixrm1, iy) > 0.) {
          double fp1 = A(fqp, ixrm1, iy), fp2 = cffqpsat_ * fqpsatrb[iy];
          A(fqp, ixrm1, iy) = ue_pow(ue_pow(fabs(fp1 * fp2), exjbdry_) / (ue_pow(fabs(fp1), exjbdry_) + ue_pow(fabs(fp2), exjbdry_)), 1 / exjbdry_);
        }
      }
    }
  FOR2(iy, j1p, j6p, ix, i1, i5) {  // potencur.m:206-221
      const int iy1 = mx(0, iy - 1), ix1 = IXP1(ix, iy);
      const double temp1 = 4.0 * (A(prtv, ix, iy) - A(prtv, ix, iy1)) * A(gyc, ix, iy);
      A(fq2d, ix, iy) = A(sx, ix, iy) * 0.25 * temp1 * (A(rbfbt, ix1, iy) + A(rbfbt, ix, iy)) / (A(btot, ix, iy) + A(btot, ix1, iy));
      A(fq2, ix, iy) = cfjp2 * A(fq2d, ix, iy);
    }
  FOR2(iy, j1p, j5p, ix, i1, i6) {  // potencur.m:235-290
      double nbary = (A(ne, ix, iy + 1) * A(gy, ix, iy + 1) + A(ne, ix, iy) * A(gy, ix, iy)) / (A(gy, ix, iy + 1) + A(gy, ix, iy));
      double zfac = 1. / (A(zeff, ix, iy) * (1.193 - 0.2205 * A(zeff, ix, iy) + 0.0275 * sq(A(zeff, ix, iy))));
      double sigbary = zfac * rsigpl * sigbar0_;
      if (iy < iysptrx && ix > ixpt1 && ix < ixpt2 + 1) sigbary = sigbary + zfac * rsigplcore * sigbar0_;
      if (iy == 0) sigbary = 0.;
      A(fqyae, ix, iy) = (A(sy, ix, iy) * sigbary / (A(dynog, ix, iy) * qe)) * ((A(ney1, ix, iy) * A(tey1, ix, iy) - A(ney0, ix, iy) * A(tey0, ix, iy)) / nbary - qe * (A(phiy1, ix, iy) - A(phiy0, ix, iy)));
      double fqyai = -(A(sy, ix, iy) * sigbary / (A(dynog, ix, iy) * qe * zi[0])) * ((A(niy1[0], ix, iy) * A(tiy1, ix, iy) - A(niy0[0], ix, iy) * A(tiy0, ix, iy)) / nbary + qe * zi[0] * (A(phiy1, ix, iy) - A(phiy0, ix, iy)));
      A(fqyao, ix, iy) = cfqyao * (cfqyae_ * A(fqyae, ix, iy) + cfqyai_ * fqyai);
      const int ix3 = IXM1(ix, iy);
      const double temp1 = 4.0 * (A(prtv, ix, iy) - A(prtv, ix3, iy)) * A(gxc, ix, iy);
      A(fqyd, ix, iy) = -A(sy, ix, iy) * 0.125 * temp1 * (A(rbfbt2, ix, iy) + A(rbfbt2, ix, iy + 1));
      double nzvibtot = 0.;
      for (int f = 0; f < nisp; ++f) nzvibtot = nzvibtot + 0.5 * zi[f] * (A(niy0[f], ix, iy) + A(niy1[f], ix, iy)) * A(vycb[f], ix, iy);
      A(fqyb, ix, iy) = qe * A(sy, ix, iy) * (nzvibtot - 0.5 * (A(ney0, ix, iy) + A(ney1, ix, iy)) * A(veycb, ix, iy));
    }
  // inertia current (potencur.m:291-376); fmity as a local pair of planes per species
  for (int f = 0; f < nisp; ++f) {
    if (!(zi[f] > 1.e-10)) continue;
    FOR2(iy, mx(j1p, 1), mn(j5p, ny), ix, i1, i6) {
        const int iyp2 = mn(iy + 2, ny + 1);
        auto ut = [&](int jy, int jy1) {  // faces jy (between rows jy and jy1)
          return (4 / sq(A(btot, ix, jy) + A(btot, ix, jy1))) * (A(ey, ix, jy) - 2 * cfgpijr_ * A(gpiy[f], ix, jy) / (qe * zi[f] * (A(niy1[f], ix, jy) + A(niy0[f], ix, jy))));
        };
        double utm = ut(iy - 1, iy), ut0 = ut(iy, iy + 1), utp = (iy < ny) ? ut(iy + 1, iyp2) : 0.;
        A(fmity[f], ix, iy) = -0.25 * mi[f] * (difutm_[f] + 0.) * ((A(niy1[f], ix, iy) + A(niy0[f], ix, iy)) * (2 * r0slab_ + A(rm_c, ix, iy) + A(rm_c, ix, iy + 1)) * ut0 -
                                                                    (A(niy1[f], ix, iy - 1) + A(niy0[f], ix, iy - 1)) * (2 * r0slab_ + A(rm_c, ix, iy - 1) + A(rm_c, ix, iy)) * utm) * A(gy, ix, iy);
        A(fmity[f], ix, iy + 1) = -0.25 * mi[f] * (difutm_[f] + 0.) * ((A(niy1[f], ix, iy + 1) + A(niy0[f], ix, iy + 1)) * (2 * r0slab_ + A(rm_c, ix, iy + 1) + A(rm_c, ix, iyp2)) * utp -
                                                                        (A(niy1[f], ix, iy) + A(niy0[f], ix, iy)) * (2 * r0slab_ + A(rm_c, ix, iy) + A(rm_c, ix, iy + 1)) * ut0) * A(gy, ix, iy + 1);
        double omgci = qe * zi[f] * A(b_c, ix, iy) / mi[f];
        A(fqymi_[f], ix, iy) = qe * 0.5 * (A(niy0[f], ix, iy) + A(niy1[f], ix, iy)) * (A(vyce[f], ix, iy) + A(vycp[f], ix, iy)) *
                               (-0.5 * ((A(btot, ix, iy + 1) + A(btot, ix, iyp2)) * utp - (A(btot, ix, iy - 1) + A(btot, ix, iy)) * utm)) * 0.5 * A(gyf, ix, iy) * A(sy, ix, iy) / omgci;
      }
  }
  FOR2(iy, mx(j1p, 1), mn(j5p, ny), ix, i1, i6) {
      A(fqya, ix, iy) = 0.; A(fqym, ix, iy) = 0.; A(fqydt, ix, iy) = 0.;
      for (int f = 0; f < nisp; ++f)
        if (zi[f] > 1.e-10) {
          A(fqya, ix, iy) = A(fqya, ix, iy) + (2 / (A(rm_c, ix, iy) + A(rm_c, ix, iy + 1))) * ((A(fmity[f], ix, iy + 1) - A(fmity[f], ix, iy)) * A(gyf, ix, iy) * A(sy, ix, iy));
          A(fqym, ix, iy) = A(fqym, ix, iy) + A(fqymi_[f], ix, iy);
        }
    }
  FOR1(ix, i1, i6) {
    if (isixcore[ix] == 1) { for (int iy = 0; iy <= nfqya0core_; ++iy) if (inrow(iy)) A(fqya, ix, iy) = 0.; }
    else { for (int iy = 0; iy <= nfqya0pf_; ++iy) if (inrow(iy)) A(fqya, ix, iy) = 0.; }
    for (int iy = ny; iy >= ny + 1 - nfqya0ow_; --iy) if (inrow(iy)) A(fqya, ix, iy) = 0.;
  }
  FOR2(iy, j1p, j5p, ix, i1, i6) {
      A(fqy, ix, iy) = (1. - rnewpot_) * A(fqyao, ix, iy) + rnewpot_ * A(fqya, ix, iy) + cfqybf * A(fqyb, ix, iy) + cfqym * A(fqym, ix, iy) + cfjpy * A(fqyd, ix, iy);
      A(fqygp, ix, iy) = (1. - rnewpot_) * A(fqyao, ix, iy) + rnewpot_ * A(fqya, ix, iy) + cfqym * A(fqym, ix, iy) + A(fqyd, ix, iy);
      A(fqy, ix, iy) = A(fqy, ix, iy) + cfqydt * A(fqydt, ix, iy);  // nx = nxold, ny = nyold; cfqydt = 0
      if (cfvycf != 0.) A(fqy, ix, iy) = qe * 0.5 * (A(niy1[0], ix, iy) + A(niy0[0], ix, iy)) * A(vycf, ix, iy);  // classical Braginskii model (potencur.m:405-408)
    }
  FOR2(iy, j1p, j6p, ix, i1, i5) {
      const int ix1 = IXP1(ix, iy);
      double nzvibtot = 0.;
      for (int f = 0; f < nisp; ++f) nzvibtot = nzvibtot + 0.5 * zi[f] * (A(ni[f], ix, iy) + A(ni[f], ix1, iy)) * A(v2cb[f], ix, iy);
      A(fqxb, ix, iy) = qe * A(sx, ix, iy) * (nzvibtot - 0.5 * (A(ne, ix, iy) + A(ne, ix1, iy)) * A(ve2cb, ix, iy)) * 0.5 * (A(rbfbt, ix1, iy) + A(rbfbt, ix, iy));
      A(fqx, ix, iy) = A(fqp, ix, iy) + A(fq2, ix, iy) + cfq2bf * A(fqxb, ix, iy);
    }
  if (isexunif == 1)
    FOR1(iy, j1p, j6p) {
      if (i1 <= ixlb + 1 && ixlb + 1 <= i5) A(fqx, ixlb, iy) = A(fqx, ixlb + 1, iy);
      if (i1 <= ixrb && ixrb <= i5) A(fqx, ixrb, iy) = A(fqx, ixrb - 1, iy);
    }
}

// ---- poteneq (potencur.m:497-597) ------------------------------------------------------------------------------------
HD void poteneq(const Win& w, const double* yl, double* yldot) {
  (void)yl;
  FOR2(iy, w.j2p, w.j5p, ix, w.i2, w.i5) {
      const int ix1 = IXM1(ix, iy);
      const bool isgc = (ix == ixlb) || (ix == ixrb + 1);
      if (isgc) A(resphi, ix, iy) = 0.;
      else A(resphi, ix, iy) = (nurlxp * (dx0_ * dx0_) / sigbar0_) * (A(fqx, ix1, iy) - A(fqx, ix, iy) + A(fqy, ix, iy - 1) - A(fqy, ix, iy) + 0.);
    }
  FOR2(iy, w.j2p, w.j5p, ix, w.i2, w.i5) {
      const int64_t iv3 = IDXPHI(ix, iy);
      if (iv3 < 0) continue;
      const bool isgc = (ix == ixlb) || (ix == ixrb + 1), isgc1 = (ix == ixlb + 1) || (ix == ixrb);
      if (isexunif == 0) { if (!isgc) yldot[iv3] = A(resphi, ix, iy) / (A(vol, ix, iy) * temp0); }
      else if (!isgc && !isgc1) yldot[iv3] = A(resphi, ix, iy) / (A(vol, ix, iy) * temp0);
    }
}

// ---- bouncon (boundary.m:4-3700): guard-cell equations ----------------------------------------------------------------
const double *recycmlb, *recycmrb, *lyphiix1, *lyphiix2, *iphibcwoix, *iphibcwiix, *phi0l, *phi0r, *bctype;
double kappamx_, cfkincor_, gamsec_, cgengpl_, cgmompl_, nglfix_, ngrfix_, eedisspr_, eidisspr_, cmntgpr_, phintewi_, phintewo_;
const double *lyup_;
HD int bouncon(const Win& w, const double* yl, double* yldot) {
  (void)yl;
  const double pi = pi_;
  const int ix_fl_bc = mn(ixpt2, nx);
  const double expkmx = ue_exp(-kappamx_);
  const int g = iigsp;
  // ===== iy = 0 boundary (boundary.m:102-983) =====
  if (w.j3 <= 0) {  // isextrnpf = isextrtpf = isextrngc = 0
    for (int f = 0; f < nisp; ++f) {
      FOR1(ix, w.i4, w.i8) {
        const int64_t iv1 = IDXN(f, ix, 0);
        if (iv1 < 0) continue;
        if (isupgon == 1 && zi[f] == 0.0) {  // inertial atoms (boundary.m:131-230)
          if (isixcore[ix] == 1) {
            if (isngcore1 == 0) {
              double t0 = mx(A(tg, ix, 0), tgmin * ev);
              double vyn = sqrt(0.5 * t0 / (pi * mi[f]));
              double nharmave = 2. * (A(ni[f], ix, 0) * A(ni[f], ix, 1)) / (A(ni[f], ix, 0) + A(ni[f], ix, 1));
              double fng_alb = (1 - albedoc[0]) * nharmave * vyn * A(sy, ix, 0);
              yldot[iv1] = -nurlxg * (A(fniy[f], ix, 0) + fng_alb) / (vpnorm * A(sy, ix, 0) * n0[f]);
            } else if (isngcore1 == 1) yldot[iv1] = nurlxn * (ngcore[0] - A(ni[f], ix, 0)) / n0[f];
            else if (isngcore1 == 3) {
              double nbound = A(ng, ix, 1) - A(gyf, ix, 1) * (A(ng, ix, 2) - A(ng, ix, 1)) / A(gyf, ix, 0);
              nbound = 1.2 * nbound / (1 + 0.5 * ue_exp(-2 * (nbound / A(ng, ix, 1) - 1))) + 0.2 * A(ng, ix, 1);
              yldot[iv1] = nurlxn * (nbound - A(ng, ix, 0)) / n0[f];
            } else yldot[iv1] = nurlxn * (A(ni[f], ix, 1) - A(ni[f], ix, 0)) / n0[f];
          } else {
            double t0 = mx(A(tg, ix, 0), tgmin * ev);
            double vyn = sqrt(0.5 * t0 / (pi * mi[0]));
            double fng_chem = 0.;
            double nharmave = 2. * (A(ni[f], ix, 0) * A(ni[f], ix, 1)) / (A(ni[f], ix, 0) + A(ni[f], ix, 1));
            double fng_alb = (1 - albedoi[ix]) * nharmave * vyn * A(sy, ix, 0);
            yldot[iv1] = -nurlxg * (A(fniy[f], ix, 0) + fng_alb - fng_chem) / (vyn * A(sy, ix, 0) * n0[f]);
            if (matwalli[ix] > 0) {
              if (recycwit[ix] > 0.) {
                double fniy_recy = recycwit[ix] * fac2sp * A(fniy[0], ix, 0);
                if (isrefluxclip == 1) fniy_recy = mn(fniy_recy, 0.);
                yldot[iv1] = -nurlxg * (A(fniy[f], ix, 0) + fniy_recy - fngyi_use[ix] - fngysi[ix] + fng_alb - fng_chem) / (vyn * n0[f] * A(sy, ix, 0));
              } else if (recycwit[ix] < -1) yldot[iv1] = nurlxg * (ngbackg_[0] - A(ni[f], ix, 0)) / n0[f];
              else {
                nharmave = 2. * (A(ni[f], ix, 0) * A(ni[f], ix, 1)) / (A(ni[f], ix, 0) + A(ni[f], ix, 1));
                yldot[iv1] = -nurlxg * (A(fniy[f], ix, 0) + (1 + recycwit[ix]) * nharmave * vyn * A(sy, ix, 0)) / (vyn * n0[f] * A(sy, ix, 0));
              }
            }
            if (fngysi[ix] + fngyi_use[ix] != 0. && matwalli[ix] == 0.) yldot[iv1] = -nurlxg * (A(fniy[f], ix, 0) - fngysi[ix] - fngyi_use[ix]) / (vyn * A(sy, ix, 0) * n0[f]);
          }
        } else if (isixcore[ix] == 1) {
          if (isnicore[f] == 1) yldot[iv1] = nurlxn * (ncore[f] - A(ni[f], ix, 0)) / n0[f];
          else if (isnicore[f] == 0) yldot[iv1] = -nurlxn * (qe * (A(fniy[f], ix, 0) - fniycbo[f][ix]) / A(sy, ix, 0) - curcore[f] * A(gyf, ix, 0) / sygytotc) / (qe * vpnorm * n0[f]);
          else { errc = 3; return -4; }
        } else if (isnwconiix[f * NXS + ix] == 0) {
          yldot[iv1] = nurlxn * ((1 - ifluxni) * (A(niy1[f], ix, 0) - A(niy0[f], ix, 0)) - ifluxni * (A(fniy[f], ix, 0) / (A(sy, ix, 0) * vpnorm) - 0.001 * A(ni[f], ix, 1) * A(vy[f], ix, 0) / vpnorm)) / n0[f];
        } else if (isnwconiix[f * NXS + ix] == 1) yldot[iv1] = nurlxn * (nwalli[ix] - A(ni[f], ix, 0)) / n0[f];
        else if (isnwconiix[f * NXS + ix] == 2) {
          double nbound = A(ni[f], ix, 1) - A(gyf, ix, 1) * (A(ni[f], ix, 2) - A(ni[f], ix, 1)) / A(gyf, ix, 0);
          nbound = 1.2 * nbound / (1 + 0.5 * ue_exp(-2 * (nbound / A(ni[f], ix, 1) - 1))) + 0.2 * A(ni[f], ix, 1);
          yldot[iv1] = nurlxn * (nbound - A(ni[f], ix, 0)) / n0[f];
        } else if (isnwconiix[f * NXS + ix] == 3)
          yldot[iv1] = -nurlxn * (A(niy0[f], ix, 0) - A(niy1[f], ix, 0) * (2 * A(gyf, ix, 0) * lynipf[ix] - 1) / (2 * A(gyf, ix, 0) * lynipf[ix] + 1) - nwimin[f]) / n0[f];
      }
      SER {
        if (isfixlb != 2 && IDXN(f, ixlb, 0) >= 0) yldot[IDXN(f, ixlb, 0)] = nurlxn * (ave(A(ni[f], ixlb, 1), A(ni[f], ixlb + 1, 0)) - A(ni[f], ixlb, 0)) / n0[f];
        if (isfixrb != 2 && IDXN(f, ixrb + 1, 0) >= 0) yldot[IDXN(f, ixrb + 1, 0)] = nurlxn * (ave(A(ni[f], ixrb + 1, 1), A(ni[f], ixrb, 0)) - A(ni[f], ixrb + 1, 0)) / n0[f];
      }
    }
    for (int f = 0; f < nusp; ++f)
      FOR1(ix, w.i4, w.i8) {  // parallel velocity, boundary.m:308-383
        const int64_t iv2 = IDXU(f, ix, 0);
        if (iv2 < 0) continue;
        if (isixcore[ix] == 1) {
          if (isupcore[f] == 0) yldot[iv2] = nurlxu * (upcore[f] - A(up[f], ix, 0)) / vpnorm;
          else if (isupcore[f] == 1) yldot[iv2] = nurlxu * (A(up[f], ix, 1) - A(up[f], ix, 0)) / vpnorm;
          else if (isupcore[f] == 2) yldot[iv2] = nurlxu * ((A(up[f], ix, 1) - A(up[f], ix, 0)) * A(gy, ix, 1) - (A(up[f], ix, 2) - A(up[f], ix, 1)) * A(gy, ix, 2)) / (A(gy, ix, 1) * vpnorm);
          else yldot[iv2] = -nurlxu * A(fmiy[f], ix, 0) / (vpnorm * A(sy, ix, 0) * fnorm[f]);
        } else if (isupwiix[f * NXS + ix] == 1) yldot[iv2] = -nurlxu * A(fmiy[f], ix, 0) / (vpnorm * A(sy, ix, 0) * fnorm[f]);
        else if (isupwiix[f * NXS + ix] == 2) yldot[iv2] = nurlxu * A(nm[f], ix, 0) / fnorm[f] * (A(up[f], ix, 1) - A(up[f], ix, 0));
        else if (isupwiix[f * NXS + ix] == 3) yldot[iv2] = -nurlxu * A(nm[f], ix, 0) / fnorm[f] * (A(up[f], ix, 0) - A(up[f], ix, 1) * (2 * A(gyf, ix, 0) * lyup_[0] - 1) / (2 * A(gyf, ix, 0) * lyup_[0] + 1));
        else yldot[iv2] = nurlxu * A(nm[f], ix, 0) / fnorm[f] * (0. - A(up[f], ix, 0));
      }
    FOR1(ix, w.i4, w.i8) {  // Te, Ti, boundary.m:524-628
      const int64_t iv1 = IDXTE(ix, 0), iv2 = IDXTI(ix, 0);
      if (isixcore[ix] == 1) {
        if (iv1 >= 0) yldot[iv1] = nurlxe * (tcoree * ev - A(te, ix, 0)) * 1.5 * A(ne, ix, 0) / ennorm;
        if (iv2 >= 0) yldot[iv2] = nurlxi * (tcorei * ev - A(ti, ix, 0)) * 1.5 * A(ne, ix, 0) / ennorm;
        if (iflcore == 1) {
          if (iv1 >= 0) yldot[iv1] = -nurlxe * (A(te, ix, 0) - A(te, IXP1(ix, 0), 0)) * n0[0] / ennorm;
          if (iv2 >= 0) yldot[iv2] = -nurlxi * (A(ti, ix, 0) - A(ti, IXP1(ix, 0), 0)) * n0[0] / ennorm;
          if (ix == ix_fl_bc) {
            int ii = mx(0, ixpt1 + 1);
            double feeytotc = A(feey, ii, 0) - feeycbo[ii], feiytotc = A(feiy, ii, 0) - feiycbo[ii];
            do { ii = IXP1(ii, 0); feeytotc = feeytotc + A(feey, ii, 0) - feeycbo[ii]; } while (ii != ix_fl_bc);
            ii = mx(0, ixpt1 + 1);
            do { ii = IXP1(ii, 0); feiytotc = feiytotc + A(feiy, ii, 0) - feiycbo[ii]; } while (ii != ix_fl_bc);
            if (iv1 >= 0) yldot[iv1] = -nurlxe * (feeytotc - pcoree) / (vpnorm * ennorm);
            if (iv2 >= 0) yldot[iv2] = -nurlxi * (feiytotc - pcorei) / (vpnorm * ennorm);
          }
        } else if (iflcore == -1) {
          if (iv1 >= 0) yldot[iv1] = -nurlxe * (A(te, ix, 0) - A(te, ix, 1)) * n0[0] / ennorm;
          if (iv2 >= 0) yldot[iv2] = -nurlxi * (A(ti, ix, 0) - A(ti, ix, 1)) * n0[0] / ennorm;
        }
      } else {
        if (iv1 >= 0) {
          if (istepfcix[ix] == 0) yldot[iv1] = -nurlxe * (A(feey, ix, 0) / (n0[0] * vpnorm * A(sy, ix, 0))) / (temp0 * ev);
          else if (istepfcix[ix] == 1) yldot[iv1] = nurlxe * (tewalli[ix] * ev - A(te, ix, 0)) / (temp0 * ev);
          else if (istepfcix[ix] == 2) {
            double tbound = A(te, ix, 1) - A(gyf, ix, 1) * (A(te, ix, 2) - A(te, ix, 1)) / A(gyf, ix, 0);
            tbound = mx(tbound, tbmin * ev);
            yldot[iv1] = nurlxe * (tbound - A(te, ix, 0)) / (temp0 * ev);
          } else yldot[iv1] = nurlxe * ((A(te, ix, 1) - A(te, ix, 0)) - 0.5 * (A(te, ix, 1) + A(te, ix, 0)) / (A(gyf, ix, 0) * lytepf[ix])) / (temp0 * ev);
        }
        if (iv2 >= 0) {
          if (istipfcix[ix] == 0) yldot[iv2] = -nurlxi * (A(feiy, ix, 0) / (n0[0] * vpnorm * A(sy, ix, 0))) / (temp0 * ev);
          else if (istipfcix[ix] == 1) yldot[iv2] = nurlxi * (tiwalli[ix] * ev - A(ti, ix, 0)) / (temp0 * ev);
          else if (istipfcix[ix] == 2) {
            double tbound = A(ti, ix, 1) - A(gyf, ix, 1) * (A(ti, ix, 2) - A(ti, ix, 1)) / A(gyf, ix, 0);
            tbound = mx(tbound, tbmin * ev);
            yldot[iv2] = nurlxi * (tbound - A(ti, ix, 0)) / (temp0 * ev);
          } else yldot[iv2] = nurlxi * ((A(ti, ix, 1) - A(ti, ix, 0)) - 0.5 * (A(ti, ix, 1) + A(ti, ix, 0)) / (A(gyf, ix, 0) * lytipf[ix])) / (temp0 * ev);
        }
      }
    }
    FOR1(ix, w.i4, w.i8) {  // diffusive neutral density, boundary.m:632-767
      const int64_t iv = IDXG(ix, 0);
      if (iv < 0) continue;
      double t0 = mx(cdifg[0] * A(tg, ix, 0), tgmin * ev);
      double vyn = 0.25 * sqrt(8 * t0 / (pi * mg_[0]));
      double nharmave = 2. * (A(ng, ix, 0) * A(ng, ix, 1)) / (A(ng, ix, 0) + A(ng, ix, 1));
      if (isixcore[ix] == 1) {
        if (isngcore1 == 0) { double fng_alb = (1 - albedoc[0]) * nharmave * vyn * A(sy, ix, 0); yldot[iv] = -nurlxg * (A(fngy, ix, 0) + fng_alb) / (vyn * A(sy, ix, 0) * n0g_[0]); }
        else if (isngcore1 == 1) yldot[iv] = nurlxg * (ngcore[0] - A(ng, ix, 0)) / n0g_[0];
        else if (isngcore1 == 2) { double lengg = sqrt(A(tg, ix, 0) / (mg_[0] * (A(nuix, ix, 0) * A(nuiz, ix, 0)))); yldot[iv] = nurlxn * ((A(ng, ix, 1) - A(ng, ix, 0)) - 0.5 * (A(ng, ix, 1) + A(ng, ix, 0)) / (A(gyf, ix, 0) * lengg)) / n0g_[0]; }
        else if (isngcore1 == 3) {
          double nbound = A(ng, ix, 1) - A(gyf, ix, 1) * (A(ng, ix, 2) - A(ng, ix, 1)) / A(gyf, ix, 0);
          nbound = 1.2 * nbound / (1 + 0.5 * ue_exp(-2 * (nbound / A(ng, ix, 1) - 1))) + 0.2 * A(ng, ix, 1);
          yldot[iv] = nurlxn * (nbound - A(ng, ix, 0)) / n0g_[0];
        } else yldot[iv] = nurlxn * (A(ng, ix, 1) - A(ng, ix, 0)) / n0g_[0];
      } else {
        double fng_chem = 0., sputflxpf = 0.;
        double fng_alb = (1 - albedoi[ix]) * nharmave * vyn * A(sy, ix, 0);
        yldot[iv] = -nurlxg * (A(fngy, ix, 0) + fng_alb - fng_chem + sputflxpf) / (vyn * A(sy, ix, 0) * n0g_[0]);
        if (matwalli[ix] > 0) {
          if (recycwit[ix] > 0.) {
            double fniy_recy = fac2sp * A(fniy[0], ix, 0);
            if (isrefluxclip == 1) fniy_recy = mn(fniy_recy, 0.);
            yldot[iv] = -nurlxg * (A(fngy, ix, 0) + fniy_recy * recycwit[ix] - fngyi_use[ix] - fngysi[ix] + fng_alb - fng_chem + sputflxpf) / (vyn * n0g_[0] * A(sy, ix, 0));
          } else if (recycwit[ix] < -1) yldot[iv] = nurlxg * (ngbackg_[0] - A(ng, ix, 0)) / n0g_[0];
          else {
            nharmave = 2. * (A(ng, ix, 0) * A(ng, ix, 1)) / (A(ng, ix, 0) + A(ng, ix, 1));
            yldot[iv] = -nurlxg * (A(fngy, ix, 0) + (1 + recycwit[ix]) * nharmave * vyn * A(sy, ix, 0)) / (vyn * n0g_[0] * A(sy, ix, 0));
          }
        }
      }
    }
    FOR1(ix, w.i4, w.i8) {  // gas temperature at iy = 0 (boundary.m:769-852)
      const int64_t iv = IDXTG(ix, 0);
      if (iv < 0) continue;
      if (isixcore[ix] == 1) {
        if (istgcore == 0) yldot[iv] = nurlxg * (A(ti, ix, 0) * cftgticore - A(tg, ix, 0)) / (temp0 * ev);
        else if (istgcore == 1) yldot[iv] = nurlxg * (tgcore * ev - A(tg, ix, 0)) / (temp0 * ev);
        else if (istgcore == 2) {
          double t0 = mx(A(tg, ix, 0), tgmin * ev);
          double vyn = sqrt(0.5 * t0 / (pi * mg_[0]));
          double nharmave = 2. * (A(ng, ix, 0) * A(ng, ix, 1)) / (A(ng, ix, 0) + A(ng, ix, 1));
          double fng_alb = (1 - albedoc[0]) * nharmave * vyn * A(sy, ix, 0);
          yldot[iv] = -nurlxg * (A(fegy, ix, 0) + cfalbedo * fng_alb * t0) / (vpnorm * ennorm * A(sy, ix, 0));
        } else yldot[iv] = nurlxg * (A(tg, ix, 1) - A(tg, ix, 0)) / (temp0 * ev);
      } else {  // private-flux wall
        if (istgpfc == 0) yldot[iv] = nurlxg * (tgwall * ev - A(tg, ix, 0)) / (temp0 * ev);
        else if (istgpfc == 1) {
          double tbound = A(tg, ix, 1) - A(gyf, ix, 1) * (A(tg, ix, 2) - A(tg, ix, 1)) / A(gyf, ix, 0);
          tbound = mx(tbound, 0.25 * tbmin * ev);
          yldot[iv] = nurlxi * (tbound - A(tg, ix, 0)) / (temp0 * ev);
        } else if (istgpfc == 2) yldot[iv] = nurlxi * ((A(tg, ix, 1) - A(tg, ix, 0)) - 0.5 * (A(tg, ix, 1) + A(tg, ix, 0)) / (A(gyf, ix, 0) * lytg1)) / (temp0 * ev);
        else if (istgpfc == 3) {  // Maxwellian thermal flux to the wall
          double t0 = mx(cdifg[0] * A(tg, ix, 1), temin * ev);
          double vyn = 0.25 * sqrt(8 * t0 / (pi * mg_[0]));
          yldot[iv] = -nurlxg * (A(fegy, ix, 0) + 2 * cgengmw * A(ng, ix, 1) * vyn * t0 * A(sy, ix, 0)) / (A(sy, ix, 0) * vpnorm * ennorm);
        } else if (istgpfc == 4) {
          double t0 = mx(A(tg, ix, 0), tgmin * ev);
          double vyn = sqrt(0.5 * t0 / (pi * mg_[0]));
          double nharmave = 2. * (A(ng, ix, 0) * A(ng, ix, 1)) / (A(ng, ix, 0) + A(ng, ix, 1));
          double fng_alb = (1 - albedoi[ix]) * nharmave * vyn * A(sy, ix, 0), fng_chem = 0.;
          yldot[iv] = -nurlxg * (A(fegy, ix, 0) + cfalbedo * fng_alb * t0 - 2. * fng_chem * t0) / (vpnorm * ennorm * A(sy, ix, 0));
          if (matwalli[ix] > 0 && recycwit[ix] > 0) {
            double fniy_recy = recycwit[ix] * fac2sp * A(fniy[0], ix, 0);
            if (isrefluxclip == 1) fniy_recy = mn(fniy_recy, 0.);
            yldot[iv] = -nurlxg * (A(fegy, ix, 0) + cfalbedo * fng_alb * t0 - 2. * fng_chem * t0 + fniy_recy * (1. - cfdiss) * cfalbedo * recycwe * A(ti, ix, 0)) / (vpnorm * ennorm * A(sy, ix, 0));
          }
        } else yldot[iv] = nurlxg * (A(ti, ix, 0) * cftgtipfc - A(tg, ix, 0)) / (temp0 * ev);  // istgpfc = 5
      }
    }
    FOR1(ix, w.i4, w.i8) {  // potential, isnewpot = 0 (boundary.m:856-863)
      const int64_t iv3 = IDXPHI(ix, 0);
      if (iv3 >= 0) yldot[iv3] = nurlxp * ((A(phi, ix, 1) - A(phi, ix, 0)) - 0.5 * (A(phi, ix, 1) + A(phi, ix, 0)) / (A(gyf, ix, 0) * lyphiix1[ix])) / temp0;
    }
    if (w.xcnearlb || w.openbox) SER {  // corners, boundary.m:897-938
      for (int f = 0; f < nusp; ++f) if (IDXU(f, ixlb, 0) >= 0) yldot[IDXU(f, ixlb, 0)] = -nurlxu * (A(up[f], ixlb, 0) - 0.5 * (A(up[f], ixlb, 1) + A(up[f], ixlb + 1, 0))) / vpnorm;
      if (IDXTE(ixlb, 0) >= 0) yldot[IDXTE(ixlb, 0)] = nurlxe * (0.5 * (A(te, ixlb + 1, 0) + A(te, ixlb, 1)) - A(te, ixlb, 0)) / (temp0 * ev);
      if (IDXTI(ixlb, 0) >= 0) yldot[IDXTI(ixlb, 0)] = nurlxi * (0.5 * (A(ti, ixlb + 1, 0) + A(ti, ixlb, 1)) - A(ti, ixlb, 0)) / (temp0 * ev);
      if (IDXG(ixlb, 0) >= 0) yldot[IDXG(ixlb, 0)] = nurlxg * (A(ng, ixlb + 1, 0) - A(ng, ixlb, 0)) / n0g_[0];
      if (IDXTG(ixlb, 0) >= 0) yldot[IDXTG(ixlb, 0)] = nurlxg * (A(tg, ixlb + 1, 0) - A(tg, ixlb, 0)) / (temp0 * ev);  // boundary.m:930-937
    }
    if (w.xcnearrb || w.openbox) SER {  // boundary.m:939-983
      for (int f = 0; f < nusp; ++f)
        if (IDXU(f, ixrb, 0) >= 0) {
          yldot[IDXU(f, ixrb, 0)] = -nurlxu * (A(up[f], ixrb, 0) - 0.5 * (A(up[f], ixrb - 1, 0) + A(up[f], ixrb, 1))) / vpnorm;
          yldot[IDXU(f, ixrb + 1, 0)] = -nurlxu * (A(up[f], ixrb + 1, 0) - A(up[f], ixrb, 0)) / vpnorm;
        }
      if (IDXTE(ixrb + 1, 0) >= 0) yldot[IDXTE(ixrb + 1, 0)] = nurlxe * (0.5 * (A(te, ixrb + 1, 1) + A(te, ixrb, 0)) - A(te, ixrb + 1, 0)) / (temp0 * ev);
      if (IDXTI(ixrb + 1, 0) >= 0) yldot[IDXTI(ixrb + 1, 0)] = nurlxi * (0.5 * (A(ti, ixrb + 1, 1) + A(ti, ixrb, 0)) - A(ti, ixrb + 1, 0)) / (temp0 * ev);
      if (IDXG(ixrb, 0) >= 0) yldot[IDXG(ixrb + 1, 0)] = nurlxg * (A(ng, ixrb, 0) - A(ng, ixrb + 1, 0)) / n0g_[0];
      if (IDXTG(ixrb, 0) >= 0) yldot[IDXTG(ixrb + 1, 0)] = nurlxg * (A(tg, ixrb, 0) - A(tg, ixrb + 1, 0)) / (temp0 * ev);  // boundary.m:975-982
    }
  }
  // ===== potential with isnewpot = 1: two equations at iy = 0 and 1 (boundary.m:987-1122) =====
  if (isnewpot * isphion == 1 && w.j3 <= 3) {
    const int ixc1 = mx(0, ixpt1 + 1);
    FOR1(ix, mn(w.i4, ixpt1 + 1), mx(w.i8, ixpt2)) {
      const int64_t iv = IDXPHI(ix, 0), iv1 = IDXPHI(ix, 1);
      if (iv < 0 || iv1 < 0) continue;
      if (isixcore[ix] == 1) {  // core boundary (isphicore0 = 0): phi(,0) poloidally constant, phi(,1) by iphibcc = 1, 2, 3
        yldot[iv] = -nurlxp * (A(phi, ix, 0) - A(phi, IXP1(ix, 0), 0)) / temp0;
        if (iphibcc == 1) yldot[iv1] = -nurlxp * ((A(ey, ix, 1) - A(ey, ix, 0)) * A(gy, ix, 1) - (A(ey, ix, 2) - A(ey, ix, 1)) * A(gy, ix, 2)) / (A(gy, ix, 1) * temp0);
        else if (iphibcc == 2) yldot[iv1] = -nurlxp * (A(te, ix, 1) - A(te, IXP1(ix, 1), 1)) / (ev * temp0);
        else yldot[iv1] = -nurlxp * (A(phi, ix, 1) - A(phi, IXP1(ix, 1), 1)) / temp0;
        if (ix == ixmp) {  // midplane column: total radial current through the core boundary = icoreelec (fqyn: cfqyn = 0)
          int ii = ixc1;
          double fqytotc = A(fqya, ii, 1) + cfqyn * A(fqyn, ii, 1) + cfqym * A(fqym, ii, 1) + cfqybbo * A(fqyb, ii, 1) + cfqydbo * A(fqyd, ii, 1);
          do { ii = IXP1(ii, 1); fqytotc = fqytotc + A(fqya, ii, 1) + cfqyn * A(fqyn, ii, 1) + cfqym * A(fqym, ii, 1) + cfqybbo * A(fqyb, ii, 1) + cfqydbo * A(fqyd, ii, 1); } while (ii != ix_fl_bc);
          yldot[iv] = -nurlxp * (fqytotc - icoreelec) / (qe * n0[0] * vpnorm * A(sy, ixc1, 0));
          if (iphibcc == 1) yldot[iv1] = -nurlxp * ((A(ey, ix, 1) - A(ey, ix, 0)) * A(gy, ix, 1) - (A(ey, ix, 2) - A(ey, ix, 1)) * A(gy, ix, 2)) / (A(gy, ix, 1) * temp0);
          else yldot[iv1] = -nurlxp * (A(ey, ix, 0) - eycore) / (A(gyf, ix, 0) * temp0);
        }
        if (cfvycf > 1e-20) {  // boundary.m:1095-1100 (classical Braginskii model)
          const int ix3 = IXM1(ix, 1);
          yldot[iv] = -nurlxp * (A(ey, ix, 0) - A(gpiy[0], ix, 0) / (qe * zi[0] * A(niy0[0], ix, 0))) / (A(btot, ix, 0) * vpnorm);
          yldot[iv1] = nurlxp * (A(fqy, ix, 1) - (A(fqx, ix, 1) - A(fqx, ix3, 1))) / (A(rrv, ix, 0) * A(sy, ix, 0) * vpnorm * ev * n0[0]);
        }
      } else {  // private-flux wall
        const int k = (int)iphibcwiix[ix];
        if (k == 0) yldot[iv] = nurlxp * (A(phi, ix, 1) - A(phi, ix, 0)) / temp0;
        else if (k == 1) yldot[iv] = nurlxp * (phintewi_ * A(te, ix, 0) / ev - A(phi, ix, 0)) / temp0;
        else if (k == 3) yldot[iv] = nurlxp * ((A(phi, ix, 1) - A(phi, ix, 0)) - 0.5 * (A(phi, ix, 1) + A(phi, ix, 0)) / (A(gyf, ix, 0) * lyphiix1[ix])) / temp0;
      }
    }
  }
  // ===== iy = ny+1 boundary (boundary.m:1125-1653) =====
  if (w.j7 >= (ny + 1)) {  // isextrnw = isextrtw = 0
    for (int f = 0; f < nisp; ++f) {
      FOR1(ix, w.i4, w.i8) {
        const int64_t iv1 = IDXN(f, ix, ny + 1);
        if (iv1 < 0) continue;
        if (isupgon == 1 && zi[f] == 0.0) {  // boundary.m:1141-1173
          double t0 = mx(A(tg, ix, ny + 1), tgmin * ev);
          double vyn = sqrt(0.5 * t0 / (pi * mi[f]));
          double fng_chem = 0.;
          double nharmave = 2. * (A(ni[f], ix, ny) * A(ni[f], ix, ny + 1)) / (A(ni[f], ix, ny) + A(ni[f], ix, ny + 1));
          double fng_alb = (1 - albedoo[ix]) * nharmave * vyn * A(sy, ix, ny);
          yldot[iv1] = nurlxg * (A(fniy[f], ix, ny) - fng_alb + fng_chem) / (vyn * A(sy, ix, ny) * n0[f]);
          if (matwallo[ix] > 0) {
            if (recycwot[ix] > 0.) {
              double fniy_recy = recycwot[ix] * fac2sp * A(fniy[0], ix, ny);
              if (isrefluxclip == 1) fniy_recy = mx(fniy_recy, 0.);
              yldot[iv1] = nurlxg * (A(fniy[f], ix, ny) + fniy_recy + fngyo_use[ix] + fngyso[ix] - fng_alb + fng_chem) / (vyn * n0[f] * A(sy, ix, ny));
            } else if (recycwot[ix] < -1) yldot[iv1] = nurlxg * (ngbackg_[0] - A(ni[f], ix, ny + 1)) / n0[f];
            else yldot[iv1] = nurlxg * (A(fniy[f], ix, ny) - (1 + recycwot[ix]) * A(ni[f], ix, ny + 1) * vyn * A(sy, ix, ny)) / (vyn * n0[f] * A(sy, ix, ny));
          }
          if (fngyso[ix] + fngyo_use[ix] != 0. && matwallo[ix] == 0.) yldot[iv1] = nurlxg * (A(fniy[f], ix, ny) + fngyo_use[ix] + fngyso[ix]) / (vyn * A(sy, ix, ny) * n0[f]);
        } else if (isnwconoix[f * NXS + ix] == 0)
          yldot[iv1] = nurlxn * ((1 - ifluxni) * (A(niy0[f], ix, ny) - A(niy1[f], ix, ny)) + ifluxni * (A(fniy[f], ix, ny) / (A(sy, ix, ny) * vpnorm) - 0.001 * A(ni[f], ix, ny) * A(vy[f], ix, ny) / vpnorm)) / n0[f];
        else if (isnwconoix[f * NXS + ix] == 1) yldot[iv1] = nurlxn * (nwallo[ix] - A(ni[f], ix, ny + 1)) / n0[f];
        else if (isnwconoix[f * NXS + ix] == 2) {
          double nbound = A(ni[f], ix, ny) + A(gyf, ix, ny - 1) * (A(ni[f], ix, ny) - A(ni[f], ix, ny - 1)) / A(gyf, ix, ny);
          nbound = 1.2 * nbound / (1 + 0.5 * ue_exp(-2 * (nbound / A(ni[f], ix, ny) - 1))) + 0.2 * A(ni[f], ix, ny);
          yldot[iv1] = nurlxn * (nbound - A(ni[f], ix, ny + 1)) / n0[f];
        } else
          yldot[iv1] = -nurlxn * (A(niy1[f], ix, ny) - A(niy0[f], ix, ny) * (2 * A(gyf, ix, ny) * lyniwc[ix] - 1) / (2 * A(gyf, ix, ny) * lyniwc[ix] + 1) - nwomin[f]) / n0[f];
      }
      SER {
        if (IDXN(f, ixlb, ny + 1) >= 0) yldot[IDXN(f, ixlb, ny + 1)] = nurlxn * (ave(A(ni[f], ixlb, ny), A(ni[f], ixlb + 1, ny + 1)) - A(ni[f], ixlb, ny + 1)) / n0[f];
        if (IDXN(f, ixrb + 1, ny + 1) >= 0) yldot[IDXN(f, ixrb + 1, ny + 1)] = nurlxn * (ave(A(ni[f], ixrb + 1, ny), A(ni[f], ixrb, ny + 1)) - A(ni[f], ixrb + 1, ny + 1)) / n0[f];
      }
    }
    for (int f = 0; f < nusp; ++f)
      FOR1(ix, w.i4, w.i8) {  // boundary.m:1231-1252
        const int64_t iv2 = IDXU(f, ix, ny + 1);
        if (iv2 < 0) continue;
        if (isupwoix[f * NXS + ix] == 1) yldot[iv2] = nurlxu * A(fmiy[f], ix, ny) / (vpnorm * A(sy, ix, ny) * fnorm[f]);
        else if (isupwoix[f * NXS + ix] == 2) yldot[iv2] = nurlxu * A(nm[f], ix, ny) / fnorm[f] * (A(up[f], ix, ny) - A(up[f], ix, ny + 1));
        else if (isupwoix[f * NXS + ix] == 3) yldot[iv2] = -nurlxu * A(nm[f], ix, ny) / fnorm[f] * (A(up[f], ix, ny + 1) - A(up[f], ix, ny) * (2 * A(gyf, ix, ny) * lyup_[1] - 1) / (2 * A(gyf, ix, ny) * lyup_[1] + 1));
        else yldot[iv2] = nurlxu * A(nm[f], ix, ny) / fnorm[f] * (0. - A(up[f], ix, ny + 1));
      }
    FOR1(ix, w.i4, w.i8) {  // boundary.m:1311-1362
      const int64_t iv1 = IDXTE(ix, ny + 1), iv2 = IDXTI(ix, ny + 1);
      if (iv1 >= 0) {
        if (istewcix[ix] == 0) yldot[iv1] = nurlxe * (A(feey, ix, ny) / (n0[0] * vpnorm * A(sy, ix, ny))) / (temp0 * ev);
        else if (istewcix[ix] == 1) yldot[iv1] = nurlxe * (tewallo[ix] * ev - A(te, ix, ny + 1)) / (temp0 * ev);
        else if (istewcix[ix] == 2) {
          double tbound = A(te, ix, ny) + A(gyf, ix, ny - 1) * (A(te, ix, ny) - A(te, ix, ny - 1)) / A(gyf, ix, ny);
          tbound = mx(tbound, tbmin * ev);
          yldot[iv1] = nurlxe * (tbound - A(te, ix, ny + 1)) / (temp0 * ev);
        } else yldot[iv1] = nurlxe * ((A(te, ix, ny) - A(te, ix, ny + 1)) - 0.5 * (A(te, ix, ny) + A(te, ix, ny + 1)) / (A(gyf, ix, ny) * lytewc[ix])) / (temp0 * ev);
      }
      if (iv2 >= 0) {
        if (istiwcix[ix] == 0) yldot[iv2] = nurlxi * (A(feiy, ix, ny) / (n0[0] * vpnorm * A(sy, ix, ny))) / (temp0 * ev);
        else if (istiwcix[ix] == 1) yldot[iv2] = nurlxi * (tiwallo[ix] * ev - A(ti, ix, ny + 1)) / (temp0 * ev);
        else if (istiwcix[ix] == 2) {
          double tbound = A(ti, ix, ny) + A(gyf, ix, ny - 1) * (A(ti, ix, ny) - A(ti, ix, ny - 1)) / A(gyf, ix, ny);
          tbound = mx(tbound, tbmin * ev);
          yldot[iv2] = nurlxi * (tbound - A(ti, ix, ny + 1)) / (temp0 * ev);
        } else yldot[iv2] = nurlxi * ((A(ti, ix, ny) - A(ti, ix, ny + 1)) - 0.5 * (A(ti, ix, ny) + A(ti, ix, ny + 1)) / (A(gyf, ix, ny) * lytiwc[ix])) / (temp0 * ev);
      }
    }
    FOR1(ix, w.i4, w.i8) {  // boundary.m:1366-1462
      const int64_t iv = IDXG(ix, ny + 1);
      if (iv < 0) continue;
      double t0 = mx(cdifg[0] * A(tg, ix, ny + 1), tgmin * ev);
      double vyn = 0.25 * sqrt(8 * t0 / (pi * mg_[0]));
      double fng_chem = 0., sputflxw = 0.;
      double nharmave = 2. * (A(ng, ix, ny) * A(ng, ix, ny + 1)) / (A(ng, ix, ny) + A(ng, ix, ny + 1));
      double fng_alb = (1 - albedoo[ix]) * nharmave * vyn * A(sy, ix, ny);
      yldot[iv] = nurlxg * (A(fngy, ix, ny) - fng_alb + fng_chem + sputflxw) / (vyn * A(sy, ix, ny) * n0g_[0]);
      if (matwallo[ix] > 0) {
        if (recycwot[ix] > 0.) {
          double fniy_recy = fac2sp * A(fniy[0], ix, ny);
          if (isrefluxclip == 1) fniy_recy = mx(fniy_recy, 0.);
          yldot[iv] = nurlxg * (A(fngy, ix, ny) + fniy_recy * recycwot[ix] + fngyso[ix] + fngyo_use[ix] - fng_alb + fng_chem + sputflxw) / (vyn * n0g_[0] * A(sy, ix, ny));
        } else if (recycwot[ix] < -1) yldot[iv] = nurlxg * (ngbackg_[0] - A(ng, ix, ny + 1)) / n0g_[0];
        else {
          nharmave = 2. * (A(ng, ix, ny) * A(ng, ix, ny + 1)) / (A(ng, ix, ny) + A(ng, ix, ny + 1));
          yldot[iv] = nurlxg * (A(fngy, ix, ny) - (1 + recycwot[ix]) * nharmave * vyn * A(sy, ix, ny)) / (vyn * n0g_[0] * A(sy, ix, ny));
        }
      }
    }
    FOR1(ix, w.i4, w.i8) {  // gas temperature at iy = ny+1 (boundary.m:1463-1513)
      const int64_t iv = IDXTG(ix, ny + 1);
      if (iv < 0) continue;
      if (istgwc == 0) yldot[iv] = nurlxg * (tgwall * ev - A(tg, ix, ny + 1)) / (temp0 * ev);
      else if (istgwc == 1) {
        double tbound = A(tg, ix, ny) + A(gyf, ix, ny) * (A(tg, ix, ny) - A(tg, ix, ny - 1)) / A(gyf, ix, ny);
        tbound = mx(tbound, 0.25 * tbmin * ev);
        yldot[iv] = nurlxi * (tbound - A(tg, ix, ny + 1)) / (temp0 * ev);
      } else if (istgwc == 2) yldot[iv] = nurlxi * ((A(tg, ix, ny) - A(tg, ix, ny + 1)) - 0.5 * (A(tg, ix, ny) + A(tg, ix, ny + 1)) / (A(gyf, ix, ny) * lytg2)) / (temp0 * ev);
      else if (istgwc == 3) {
        double t0 = mx(cdifg[0] * A(tg, ix, ny), temin * ev);
        double vyn = 0.25 * sqrt(8 * t0 / (pi * mg_[0]));
        yldot[iv] = nurlxg * (A(fegy, ix, ny) - 2 * cgengmw * A(ng, ix, ny) * vyn * t0 * A(sy, ix, ny)) / (A(sy, ix, ny) * vpnorm * ennorm);
      } else if (istgwc == 4) {
        double t0 = mx(A(tg, ix, ny + 1), tgmin * ev);
        double vyn = sqrt(0.5 * t0 / (pi * mg_[0]));
        double nharmave = 2. * (A(ng, ix, ny) * A(ng, ix, ny + 1)) / (A(ng, ix, ny) + A(ng, ix, ny + 1));
        double fng_alb = (1 - albedoo[ix]) * nharmave * vyn * A(sy, ix, ny), fng_chem = 0.;
        yldot[iv] = nurlxg * (A(fegy, ix, ny) - cfalbedo * fng_alb * t0 + 2. * fng_chem * t0) / (vpnorm * ennorm * A(sy, ix, ny));
        if (matwallo[ix] > 0 && recycwot[ix] > 0.) {
          double fniy_recy = recycwot[ix] * fac2sp * A(fniy[0], ix, ny);
          if (isrefluxclip == 1) fniy_recy = mx(fniy_recy, 0.);
          yldot[iv] = nurlxg * (A(fegy, ix, ny) - cfalbedo * fng_alb * t0 + 2. * fng_chem * t0 + fniy_recy * (1. - cfdiss) * cfalbedo * recycwe * A(ti, ix, ny)) / (vpnorm * ennorm * A(sy, ix, ny));
        }
      } else yldot[iv] = nurlxg * (A(ti, ix, ny + 1) * cftgtiwc - A(tg, ix, ny + 1)) / (temp0 * ev);  // istgwc = 5
    }
    FOR1(ix, w.i4, w.i8) {  // potential (boundary.m:1522-1540)
      const int64_t iv3 = IDXPHI(ix, ny + 1);
      if (iv3 < 0) continue;
      const int k = (int)iphibcwoix[ix];
      if (k == 0) yldot[iv3] = nurlxp * (A(phi, ix, ny) - A(phi, ix, ny + 1)) / temp0;
      else if (k == 1) yldot[iv3] = nurlxp * (phintewo_ * A(te, ix, ny + 1) / ev - A(phi, ix, ny + 1)) / temp0;
      else if (k == 3) yldot[iv3] = nurlxp * ((A(phi, ix, ny) - A(phi, ix, ny + 1)) - 0.5 * (A(phi, ix, ny) + A(phi, ix, ny + 1)) / (A(gyf, ix, ny) * lyphiix2[ix])) / temp0;
    }
    if (w.xcnearlb || w.openbox) SER {  // boundary.m:1543-1583
      for (int f = 0; f < nusp; ++f) if (IDXU(f, ixlb, ny + 1) >= 0) yldot[IDXU(f, ixlb, ny + 1)] = -nurlxu * (A(up[f], ixlb, ny + 1) - 0.5 * (A(up[f], ixlb, ny) + A(up[f], ixlb + 1, ny + 1))) / vpnorm;
      if (IDXTE(ixlb, ny + 1) >= 0) yldot[IDXTE(ixlb, ny + 1)] = nurlxe * (0.5 * (A(te, ixlb + 1, ny + 1) + A(te, ixlb, ny)) - A(te, ixlb, ny + 1)) / (temp0 * ev);
      if (IDXTI(ixlb, ny + 1) >= 0) yldot[IDXTI(ixlb, ny + 1)] = nurlxi * (0.5 * (A(ti, ixlb + 1, ny + 1) + A(ti, ixlb, ny)) - A(ti, ixlb, ny + 1)) / (temp0 * ev);
      if (IDXG(ixlb, ny + 1) >= 0) yldot[IDXG(ixlb, ny + 1)] = nurlxg * (A(ng, ixlb + 1, ny + 1) - A(ng, ixlb, ny + 1)) / n0g_[0];
      if (IDXTG(ixlb, ny + 1) >= 0) yldot[IDXTG(ixlb, ny + 1)] = nurlxg * (0.5 * (A(tg, ixlb + 1, ny + 1) + A(tg, ixlb, ny)) - A(tg, ixlb, ny + 1)) / (temp0 * ev);
    }
    if (w.xcnearrb || w.openbox) SER {  // boundary.m:1585-1630
      for (int f = 0; f < nusp; ++f)
        if (IDXU(f, ixrb, ny + 1) >= 0) {
          yldot[IDXU(f, ixrb, ny + 1)] = -nurlxu * (A(up[f], ixrb, ny + 1) - 0.5 * (A(up[f], ixrb - 1, ny + 1) + A(up[f], ixrb, ny))) / vpnorm;
          yldot[IDXU(f, ixrb + 1, ny + 1)] = -nurlxu * (A(up[f], ixrb + 1, ny + 1) - A(up[f], ixrb, ny + 1)) / vpnorm;
        }
      if (IDXTE(ixrb + 1, ny + 1) >= 0) yldot[IDXTE(ixrb + 1, ny + 1)] = nurlxe * (0.5 * (A(te, ixrb, ny + 1) + A(te, ixrb + 1, ny)) - A(te, ixrb + 1, ny + 1)) / (temp0 * ev);
      if (IDXTI(ixrb + 1, ny + 1) >= 0) yldot[IDXTI(ixrb + 1, ny + 1)] = nurlxi * (0.5 * (A(ti, ixrb, ny + 1) + A(ti, ixrb + 1, ny)) - A(ti, ixrb + 1, ny + 1)) / (temp0 * ev);
      if (IDXG(ixrb + 1, ny + 1) >= 0) yldot[IDXG(ixrb + 1, ny + 1)] = nurlxg * (A(ng, ixrb, ny + 1) - A(ng, ixrb + 1, ny + 1)) / n0g_[0];
      if (IDXTG(ixrb + 1, ny + 1) >= 0) yldot[IDXTG(ixrb + 1, ny + 1)] = nurlxg * (0.5 * (A(tg, ixrb, ny + 1) + A(tg, ixrb + 1, ny)) - A(tg, ixrb + 1, ny + 1)) / (temp0 * ev);
    }
  }
  // ===== ix = 0 as a symmetry plane, isfixlb = 2 (boundary.m:1666-1770; rlimiter beyond the mesh) =====
  if (w.i3 <= 0 && isfixlb == 2)
    FOR1(iy, w.j2, w.j5) {
      for (int f = 0; f < nisp; ++f) if (IDXN(f, 0, iy) >= 0) yldot[IDXN(f, 0, iy)] = nurlxn * (1 / n0[f]) * (A(ni[f], 1, iy) - A(ni[f], 0, iy));
      for (int f = 0; f < nusp; ++f) if (IDXU(f, 0, iy) >= 0) yldot[IDXU(f, 0, iy)] = nurlxu * (0. - A(up[f], 0, iy)) / vpnorm;
      if (IDXTE(0, iy) >= 0) yldot[IDXTE(0, iy)] = nurlxe * A(ne, 0, iy) * (A(te, 1, iy) - A(te, 0, iy)) / ennorm;
      if (IDXTI(0, iy) >= 0) yldot[IDXTI(0, iy)] = nurlxi * A(ne, 0, iy) * (A(ti, 1, iy) - A(ti, 0, iy)) / ennorm;
      if (IDXG(0, iy) >= 0) yldot[IDXG(0, iy)] = nurlxg * (A(ng, 1, iy) - A(ng, 0, iy)) / n0g_[0];
      if (IDXTG(0, iy) >= 0) yldot[IDXTG(0, iy)] = nurlxg * (A(tg, 1, iy) - A(tg, 0, iy)) / (temp0 * ev);  // boundary.m:1747-1756
      if (IDXPHI(0, iy) >= 0) yldot[IDXPHI(0, iy)] = nurlxp * (A(phi, 1, iy) - A(phi, 0, iy)) / temp0;
    }
  if (isfixlb == 2 && w.i2 <= ixpt2 && w.i5 >= ixpt2 && w.j2 <= iysptrx2)  // boundary.m:1772-1785
    for (int f = 0; f < nusp; ++f)
      FOR1(iy, 0, iysptrx2) if (IDXU(f, ixpt2, iy) >= 0) yldot[IDXU(f, ixpt2, iy)] = nurlxu * (0. - A(up[f], ixpt2, iy)) / vpnorm;
  // ===== left plate, ix = ixlb (boundary.m:1787-2318), isfixlb = 0 =====
  if ((w.xcnearlb || w.openbox) && isfixlb == 0) {
    const int ixt = ixlb;
    if (w.i3 <= ixlb + isextrnp)
      for (int f = 0; f < nisp; ++f)
        FOR1(iy, w.j2, w.j5) {
          const int ixt1 = IXP1(ixt, iy);
          const int64_t iv1 = IDXN(f, ixt, iy);
          if (iv1 < 0) continue;
          if (isupgon == 1 && zi[f] == 0.0) {  // boundary.m:1806-1830
            const double recy = recylb[iy];
            if (recy > 0.) {
              double t0 = mx(A(tg, ixt1, iy), tgmin * ev);
              double vxn = 0.25 * sqrt(8 * t0 / (pi * mi[f]));
              double areapl = isoldalbarea * A(sx, ixt, iy) + (1 - isoldalbarea) * A(sxnp, ixt, iy);
              yldot[iv1] = -nurlxg * (A(fnix[f], ixt, iy) + recy * A(fnix[0], ixt, iy) - fngxlb_use[iy] + (1 - alblb[iy]) * A(ni[f], ixt1, iy) * vxn * areapl - fngxslb[iy]) / (vpnorm * n0[f] * A(sx, ixt, iy));
            } else if (recy <= 0. && recy >= -1.) {
              double t0 = mx(A(tg, ixt, iy), tgmin * ev);
              double vyn = sqrt(0.5 * t0 / (pi * mi[0]));
              yldot[iv1] = -nurlxg * (A(fnix[f], ixt, iy) + (1 + recy) * A(ni[f], ixt, iy) * vyn * A(sx, ixt, iy)) / (vpnorm * n0[f] * A(sx, ixt, iy));
            } else if (recy < -1. && recy > -2.) yldot[iv1] = nurlxg * (nglfix_ - A(ni[f], ixt, iy)) / n0[f];
            else yldot[iv1] = nurlxn * (A(ni[f], ixt1, iy) - A(ni[f], ixt, iy)) / n0[f];
          } else yldot[iv1] = nurlxn * (A(ni[f], ixt1, iy) - A(ni[f], ixt, iy)) / n0[f];  // isextrnp = 0
        }
    if (w.i3 <= ixlb)
      FOR1(iy, w.j2, w.j5) {  // boundary.m:1848-2259
        const int ixt1 = IXP1(ixt, iy);
        double kfeix = 0.;
        for (int f = 0; f < nusp; ++f) {
          const int64_t iv2 = IDXU(f, ixt, iy);
          if (iv2 >= 0) {
            double cs = csfaclb[f] * sqrt((A(te, ixt, iy) + csfacti * A(ti, ixt, iy)) / mi[f]);
            if (isupgon == 1 && zi[f] == 0.0) {  // boundary.m:1871-1890
              const double rm = recycmlb[iy];
              if (rm > -9.9) yldot[iv2] = -nurlxu * (rm * A(up[0], ixt, iy) + A(up[f], ixt, iy)) / vpnorm;
              else if (rm <= -9.9 && rm > -10.1) yldot[iv2] = nurlxu * (A(up[f], ixt1, iy) - A(up[f], ixt, iy)) / vpnorm;
              else {
                double t0 = mx(A(tg, ixt1, iy), tgmin * ev);
                double vxn = cgmompl_ * 0.25 * sqrt(8 * t0 / (pi * mi[f]));
                double vparn = A(up[f], ixt, iy);
                yldot[iv2] = -nurlxu * (A(fmix[f], ixt1, iy) + vparn * vxn * 0.5 * (A(nm[f], ixt1, iy) + A(nm[f], ixt, iy)) * A(sx, ixt, iy)) / (vpnorm * fnorm[f] * A(sx, ixt, iy));
              }
            } else {
              double ueb = cfueb * (cf2ef * A(v2ce[f], ixt, iy) * A(rbfbt, ixt, iy) - A(vytan[f], ixt, iy)) / A(rrv, ixt, iy);
              yldot[iv2] = nurlxu * (-cs - ueb - A(up[f], ixt, iy)) / vpnorm;  // isbohmms = 0
              if (isupss[f] == 1 && A(up[f], ixt1, iy) + ueb < -cs) yldot[iv2] = nurlxu * (A(up[f], ixt1, iy) - A(up[f], ixt, iy)) / vpnorm;
              if (isupss[f] == -1) yldot[iv2] = nurlxu * (A(up[f], ixt1, iy) - A(up[f], ixt, iy)) / vpnorm;
            }
          }
          if (zi[f] == 0.0) kfeix = kfeix - cftiexclg * cfvcsx[f] * 0.5 * A(sx, ixt, iy) * A(visx[f], ixt1, iy) * A(gx, ixt1, iy) * (A(up[f], ixt1, iy) * A(up[f], ixt1, iy) - A(up[f], ixt, iy) * A(up[f], ixt, iy));
          else kfeix = kfeix - cfvcsx[f] * 0.5 * A(sx, ixt, iy) * A(visx[f], ixt1, iy) * A(gx, ixt1, iy) * (A(up[f], ixt1, iy) * A(up[f], ixt1, iy) - A(up[f], ixt, iy) * A(up[f], ixt, iy));
        }
        double fqpsate = 0.;
        if (isphion + isphiofft == 1) {  // boundary.m:1926-1969 (ikapmod = 0)
          double lambdae = 2e16 * sq(A(te, ixt, iy) / ev) / A(ne, ixt, iy);
          double kincor = 1. / (1 + cfkincor_ * (lambdae / A(lcone, ixt, iy)) * fabs(ev * A(phi, ixt, iy) / A(te, ixt, iy)));
          fqpsate = qe * A(ne, ixt, iy) * sqrt(A(te, ixt, iy) / (2 * pi * me)) * kincor * A(sx, ixt, iy) * A(rrv, ixt, iy);
          double arglgphi;
          if (fqpsatlb[iy] + (1. - gamsec_) * A(fqp, ixt, iy) > 0) arglgphi = ue_pow(sq((fqpsatlb[iy] + (1. - gamsec_) * A(fqp, ixt, iy)) / fqpsate) + expkmx * expkmx, 0.5);
          else arglgphi = expkmx;
          if (iskaplex == 0) kappal[iy] = -ue_log(arglgphi);
          if (newbcl == 0 && iskaplex == 0) kappal[iy] = 3.0;
          const int64_t iv = IDXPHI(ixt, iy);
          if (iv >= 0) yldot[iv] = -nurlxp * (1. - bctype[iy]) * (A(phi, ixt, iy) - kappal[iy] * A(te, ixt, iy) / ev - phi0l[iy]) / temp0 - nurlxp * bctype[iy] * (1. - gamsec_) * A(fqp, ixt, iy) / (fqpsatlb[iy] + cutlo);
          if (iv >= 0 && isphilbc == 1) yldot[iv] = -nurlxp * (A(phi, ixt, iy) - phi0l[iy]) / temp0;  // boundary.m:1962-1963
        } else { fqpsate = 0.; kappal[iy] = 3.; }
        const int isphion2 = isphion + isphiofft;
        bcel[iy] = (1 - newbcl * isphion2) * bcee + newbcl * isphion2 * (2. + kappal[iy]);
        if (iskaplex == 1) bcel[iy] = (2. + kappal[iy]);
        bcil[iy] = (1 - newbcl * isphion2) * bcei + newbcl * isphion2 * (2.5);
        double t0 = A(te, ixt, iy) / ev;
        double f_cgpld = .5 * (1. - ue_cos(pi * (t0 - temin) / (.3 - temin)));
        if (t0 < temin) f_cgpld = 0.;
        if (t0 > 0.3) f_cgpld = 1.;
        t0 = mx(A(tg, ixt1, iy), tgmin * ev);
        double vxn = f_cgpld * 0.25 * sqrt(8 * t0 / (pi * mg_[0]));
        if (IDXTE(ixt, iy) >= 0) {  // ibctepl == 1, boundary.m:1996-2014
          double faceel = bcel[iy] * (fqpsate / qe) * ue_exp(-kappal[iy]);
          double faceel2 = bcel[iy] * (fqpsate / qe) * ue_exp(-kappamx_ + 2);
          double totfeexl = A(feex, ixt, iy) + 0.;
          double totfnex = A(ne, ixt, iy) * A(vex, ixt, iy) * A(sx, ixt, iy);
          if (isphion + isphiofft == 1)
            yldot[IDXTE(ixt, iy)] = -nurlxe * (totfeexl + faceel * A(te, ixt, iy) + faceel2 * (A(te, ixt, iy) - A(te, ixt1, iy)) - cmneut * A(fnix[0], ixt, iy) * recycp[0] * eedisspl * ev) / (A(sx, ixt, iy) * vpnorm * ennorm);
          else
            yldot[IDXTE(ixt, iy)] = -nurlxe * (totfeexl - totfnex * A(te, ixt, iy) * bcel[iy] + cgpld * A(sx, ixt, iy) * 0.5 * A(ng, ixt1, iy) * vxn * ediss * ev - cmneut * A(fnix[0], ixt, iy) * recycp[0] * eedisspl * ev) / (A(sx, ixt, iy) * vpnorm * ennorm);
        }
        if (IDXTI(ixt, iy) >= 0) {  // ibctipl == 1, boundary.m:2027-2063
          double totfeixl = A(feix, ixt, iy) + ckinfl * kfeix;
          double totfnix = 0.;
          for (int f = 0; f < nfsp; ++f) if (zi[f] > 1e-10) { totfeixl = totfeixl + 0.; totfnix = totfnix + A(fnix[f], ixt, iy); }
          if (isupgon == 1)
            yldot[IDXTI(ixt, iy)] = -nurlxi * (totfeixl - totfnix * A(ti, ixt, iy) * bcil[iy] +
                                               cftiexclg * (-cfneut * A(fnix[g], ixt, iy) * A(tg, ixt, iy) * bcen + (cgengpl_ * 2. * A(tg, ixt, iy) - cgpld * eion * ev) * A(ng, ixt1, iy) * vxn * A(sx, ixt, iy) -
                                                            cmneut * A(fnix[0], ixt, iy) * recycp[0] * cmntgpl * (A(ti, ixt, iy) - eidisspl * ev))) / (vpnorm * ennorm * A(sx, ixt, iy));
          else
            yldot[IDXTI(ixt, iy)] = -nurlxi * (totfeixl - totfnix * bcil[iy] * A(ti, ixt, iy) + cftiexclg * (-cmneut * A(fnix[0], ixt, iy) * recycp[0] * cmntgpl * (A(ti, ixt, iy) - eidisspl * ev))) / (vpnorm * ennorm * A(sx, ixt, iy));
        }
        if (IDXG(ixt, iy) >= 0) {  // diffusive neutral density, boundary.m:2075-2115
          const int64_t iv = IDXG(ixt, iy);
          double recy = recylb[iy];
          if (recy > 0.) {
            double flux_inc = fac2sp * A(fnix[0], ixt, iy);
            double t0g = mx(A(tg, ixt1, iy), tgmin * ev);
            double vxg = 0.25 * sqrt(8 * t0g / (pi * mg_[0]));
            double areapl = isoldalbarea * A(sx, ixt, iy) + (1 - isoldalbarea) * A(sxnp, ixt, iy);
            yldot[iv] = -nurlxg * (A(fngx, ixt, iy) - fngxlb_use[iy] - fngxslb[iy] + recy * flux_inc + (1 - alblb[iy]) * A(ng, ixt1, iy) * vxg * areapl) / (vpnorm * n0g_[0] * A(sx, ixt, iy));
          } else if (recy <= 0. && recy >= -1.) {
            double t0g = mx(A(tg, ixt, iy), tgmin * ev);
            double vxg = 0.25 * sqrt(8 * t0g / (pi * mg_[0]));
            yldot[iv] = -nurlxg * (A(fngx, ixt, iy) + (1 + recy) * A(ng, ixt, iy) * vxg * A(sx, ixt, iy)) / (vxg * A(sx, ixt, iy) * n0g_[0]);
          } else { errc = 4; return -4; }
        }
      }
  }
  if ((w.xcnearlb || w.openbox) && isfixlb == 0 && w.i3 <= ixlb)  // gas temperature at the left plate (boundary.m:2198-2257)
    FOR1(iy, w.j2, w.j5) {
      const int ixt = ixlb, ixt1 = IXP1(ixt, iy), ixt2 = IXP1(ixt1, iy);
      const int64_t iv = IDXTG(ixt, iy);
      if (iv < 0) continue;
      if (istglb == 0) yldot[iv] = nurlxg * (tgwall * ev - A(tg, ixt, iy)) / (temp0 * ev);
      else if (istglb == 1) {
        double tbound = A(tg, ixt1, iy) - A(gyf, ixt1, iy) * (A(tg, ixt2, iy) - A(tg, ixt1, iy)) / A(gxf, ixt, iy);
        tbound = mx(tbound, 0.5 * temin * ev);
        yldot[iv] = nurlxg * (tbound - A(tg, ixt, iy)) / (temp0 * ev);
      } else if (istglb == 3) {
        double t0 = mx(cdifg[0] * A(tg, ixt1, iy), tgmin * ev);
        double vxn = 0.25 * sqrt(8 * t0 / (pi * mg_[0]));
        yldot[iv] = -nurlxg * (A(fegx, ixt, iy) + 2 * cgengmpl * A(ng, ixt1, iy) * vxn * t0 * A(sx, ixt, iy)) / (A(sx, ixt, iy) * vpnorm * ennorm);
      } else if (istglb == 4) {
        const double recy = recylb[iy];
        double t0 = mx(A(tg, ixt1, iy), tgmin * ev);
        if (recy > 0.) {
          double vxn = 0.25 * sqrt(8 * t0 / (pi * mg_[0]));
          double fng_alb = (1 - alblb[iy]) * A(ng, ixt1, iy) * vxn * A(sx, ixt, iy);
          yldot[iv] = -nurlxg * (A(fegx, ixt, iy) + cfalbedo * fng_alb * t0 + recy * (1. - cfdiss) * A(fnix[0], ixt, iy) * recyce * cfalbedo * (kappal[iy] * zi[0] * A(te, ixt, iy) + A(ti, ixt, iy))) / (vpnorm * ennorm * A(sx, ixt, iy));
        } else if (recy >= -1.) {
          double vyn = sqrt(0.5 * t0 / (pi * mg_[0]));
          double fng_alb = (1 + recy) * A(ng, ixt1, iy) * vyn * A(sx, ixt, iy);
          yldot[iv] = -nurlxg * (A(fegx, ixt, iy) + cfalbedo * fng_alb * t0) / (vpnorm * ennorm * A(sx, ixt, iy));
        } else yldot[iv] = -nurlxg * (A(fegx, ixt, iy) + cfalbedo * A(fnix[iigsp], ixt, iy) * t0) / (vpnorm * ennorm * A(sx, ixt, iy));
      } else yldot[iv] = nurlxg * (A(ti, ixt, iy) * cftgtipltl - A(tg, ixt, iy)) / (temp0 * ev);  // istglb = 5
    }
  // ===== right plate, ix = ixrb+1 (boundary.m:2320-3002), isfixrb = 0 =====
  if (w.xcnearrb || w.openbox) {
    const int ixt = ixrb + 1;
    if (w.i6 >= (ixrb + 1 - isextrnp))
      for (int f = 0; f < nisp; ++f)
        FOR1(iy, w.j2, w.j5) {
          const int ixt1 = IXM1(ixt, iy);
          const int64_t iv1 = IDXN(f, ixt, iy);
          if (iv1 < 0) continue;
          if (isupgon == 1 && zi[f] == 0.0) {  // boundary.m:2471-2495
            const double recy = recyrb[iy];
            if (recy > 0.) {
              double t0 = mx(A(tg, ixt1, iy), tgmin * ev);
              double vxn = 0.25 * sqrt(8 * t0 / (pi * mi[f]));
              double areapl = isoldalbarea * A(sx, ixt1, iy) + (1 - isoldalbarea) * A(sxnp, ixt1, iy);
              yldot[iv1] = nurlxg * (A(fnix[f], ixt1, iy) + recy * A(fnix[0], ixt1, iy) + fngxrb_use[iy] - (1 - albrb[iy]) * A(ni[f], ixt1, iy) * vxn * areapl - fngxsrb[iy]) / (vpnorm * n0[f] * A(sx, ixt1, iy));
            } else if (recy <= 0. && recy >= -1.) {
              double t0 = mx(A(tg, ixt1, iy), tgmin * ev);
              double vyn = sqrt(0.5 * t0 / (pi * mi[0]));
              yldot[iv1] = nurlxg * (A(fnix[f], ixt1, iy) - (1 + recy) * A(ni[f], ixt, iy) * vyn * A(sx, ixt1, iy)) / (vpnorm * n0[f] * A(sx, ixt1, iy));
            } else if (recy < -1. && recy > -2.) yldot[iv1] = nurlxg * (ngrfix_ - A(ni[f], ixt, iy)) / n0[f];
            else yldot[iv1] = nurlxn * (A(ni[f], ixt1, iy) - A(ni[f], ixt, iy)) / n0[f];
          } else yldot[iv1] = nurlxn * (A(ni[f], ixt1, iy) - A(ni[f], ixt, iy)) / n0[f];
        }
    if (w.i6 >= ixrb + 1)
      FOR1(iy, w.j2, w.j5) {  // boundary.m:2513-2800
        const int ixt1 = IXM1(ixt, iy), ixt2 = IXM1(ixt1, iy);
        double kfeix = 0.;
        for (int f = 0; f < nfsp; ++f) { A(upi[f], ixt, iy) = A(upi[f], ixt1, iy); A(upi[0], ixt, iy) = A(up[0], ixt, iy); }  // boundary.m:2524-2525
        for (int f = 0; f < nusp; ++f) {
          const int64_t iv2 = IDXU(f, ixt1, iy), iv = IDXU(f, ixt, iy);
          if (iv >= 0) {
            double cs = csfacrb[f] * sqrt((A(te, ixt, iy) + csfacti * A(ti, ixt, iy)) / mi[f]);
            if (isupgon == 1 && zi[f] == 0.0) {  // boundary.m:2540-2567
              const double rm = recycmrb[iy];
              if (rm > -9.9) yldot[iv2] = -nurlxu * (rm * A(up[0], ixt1, iy) + A(up[f], ixt1, iy)) / vpnorm;
              else if (rm <= -9.9 && rm > -10.1) yldot[iv2] = nurlxu * (A(up[f], ixt2, iy) - A(up[f], ixt1, iy)) / vpnorm;
              else {
                double t0 = mx(A(tg, ixt, iy), tgmin * ev);
                double vxn = cgmompl_ * 0.25 * sqrt(8 * t0 / (pi * mi[f]));
                double vparn = A(up[f], ixt, iy);
                yldot[iv2] = -nurlxu * (A(fmix[f], ixt1, iy) - vparn * vxn * 0.5 * (A(nm[f], ixt1, iy) + A(nm[f], ixt, iy)) * A(sx, ixt1, iy)) / (vpnorm * fnorm[f] * A(sx, ixt1, iy));
              }
              yldot[iv] = nurlxu * (A(up[f], ixt1, iy) - A(up[f], ixt, iy)) / vpnorm;
            } else {
              double ueb = cfueb * (cf2ef * A(v2ce[f], ixt1, iy) * A(rbfbt, ixt, iy) - A(vytan[f], ixt1, iy)) / A(rrv, ixt1, iy);
              yldot[iv2] = nurlxu * (cs - ueb - A(up[f], ixt1, iy)) / vpnorm;  // isbohmms = 0
              if (isupss[f] == 1 && A(up[f], ixt2, iy) + ueb > cs) yldot[iv2] = nurlxu * (A(up[f], ixt2, iy) - A(up[f], ixt1, iy)) / vpnorm;
              if (isupss[f] == -1) yldot[iv2] = nurlxu * (A(up[f], ixt2, iy) - A(up[f], ixt1, iy)) / vpnorm;
              yldot[iv] = nurlxu * (A(up[f], ixt1, iy) - A(up[f], ixt, iy)) / vpnorm;
            }
          }
          if (zi[f] == 0.0) kfeix = kfeix - cftiexclg * cfvcsx[f] * 0.5 * A(sx, ixt1, iy) * A(visx[f], ixt1, iy) * A(gx, ixt1, iy) * (A(up[f], ixt1, iy) * A(up[f], ixt1, iy) - A(up[f], ixt2, iy) * A(up[f], ixt2, iy));
          else kfeix = kfeix - cfvcsx[f] * 0.5 * A(sx, ixt1, iy) * A(visx[f], ixt1, iy) * A(gx, ixt1, iy) * (A(up[f], ixt1, iy) * A(up[f], ixt1, iy) - A(up[f], ixt2, iy) * A(up[f], ixt2, iy));
        }
        double fqpsate = 0.;
        if (isphion + isphiofft == 1) {  // boundary.m:2600-2648
          double lambdae = 2e16 * sq(A(te, ixt, iy) / ev) / A(ne, ixt, iy);
          double kincor = 1. / (1 + cfkincor_ * (lambdae / A(lcone, ixt, iy)) * fabs(ev * A(phi, ixt, iy) / A(te, ixt, iy)));
          fqpsate = qe * A(ne, ixt, iy) * sqrt(A(te, ixt, iy) / (2 * pi * me)) * kincor * A(sx, ixt1, iy) * A(rrv, ixt1, iy);
          double arglgphi;
          if (fqpsatrb[iy] - (1. - gamsec_) * A(fqp, ixt1, iy) > 0) arglgphi = ue_pow(sq((fqpsatrb[iy] - (1. - gamsec_) * A(fqp, ixt1, iy)) / fqpsate) + expkmx * expkmx, 0.5);
          else arglgphi = expkmx;
          kappar[iy] = -ue_log(arglgphi);  // iskaprex = 0 (note: NOT reset to 3 when newbcr = 0, unlike the left plate)
          const int64_t iv = IDXPHI(ixt, iy);
          if (iv >= 0) yldot[iv] = -nurlxp * (1. - bctype[iy]) * (A(phi, ixt, iy) - kappar[iy] * A(te, ixt, iy) / ev - phi0r[iy]) / temp0 - nurlxp * bctype[iy] * (1. - gamsec_) * A(fqp, ixt1, iy) / (fqpsatrb[iy] + cutlo);
          if (iv >= 0 && isphirbc == 1) yldot[iv] = -nurlxp * (A(phi, ixt, iy) - phi0r[iy]) / temp0;  // boundary.m:2638-2639
        } else { fqpsate = 0.; kappar[iy] = 3.; }
        const int isphion2 = isphion + isphiofft;
        bcer[iy] = (1 - newbcr * isphion2) * bcee + newbcr * isphion2 * (2. + kappar[iy]);
        bcir[iy] = (1 - newbcr * isphion2) * bcei + newbcr * isphion2 * (2.5);
        double t0 = A(te, ixt, iy) / ev;
        double f_cgpld = .5 * (1. - ue_cos(pi * (t0 - temin) / (.3 - temin)));
        if (t0 < temin) f_cgpld = 0.;
        if (t0 > 0.3) f_cgpld = 1.;
        t0 = mx(A(tg, ixt1, iy), tgmin * ev);
        double vxn = f_cgpld * 0.25 * sqrt(8 * t0 / (pi * mg_[0]));
        if (IDXTE(ixt, iy) >= 0) {  // ibctepr == 1, boundary.m:2673-2696
          double faceel = bcer[iy] * (fqpsate / qe) * ue_exp(-kappar[iy]);
          double faceel2 = bcer[iy] * (fqpsate / qe) * ue_exp(-kappamx_ + 2);
          double totfeexr = A(feex, ixt1, iy) + 0.;
          double totfnex = A(ne, ixt, iy) * A(vex, ixt1, iy) * A(sx, ixt1, iy);
          if (isphion + isphiofft == 1)
            yldot[IDXTE(ixt, iy)] = nurlxe * (totfeexr - faceel * A(te, ixt, iy) - faceel2 * (A(te, ixt, iy) - A(te, ixt1, iy)) - cmneut * A(fnix[0], ixt1, iy) * recycp[0] * eedisspr_ * ev) / (A(sx, ixt1, iy) * vpnorm * ennorm);
          else
            yldot[IDXTE(ixt, iy)] = nurlxe * (totfeexr - totfnex * A(te, ixt, iy) * bcer[iy] - cgpld * A(sx, ixt1, iy) * 0.5 * A(ng, ixt1, iy) * vxn * ediss * ev - cmneut * A(fnix[0], ixt1, iy) * recycp[0] * eedisspr_ * ev) / (A(sx, ixt1, iy) * vpnorm * ennorm);
        }
        if (IDXTI(ixt, iy) >= 0) {  // ibctipr == 1, boundary.m:2709-2747
          double totfeixr = A(feix, ixt1, iy) + ckinfl * kfeix;
          double totfnix = 0.;
          for (int f = 0; f < nfsp; ++f) if (zi[f] > 1e-10) { totfeixr = totfeixr + 0.; totfnix = totfnix + A(fnix[f], ixt1, iy); }
          if (isupgon == 1)
            yldot[IDXTI(ixt, iy)] = nurlxi * (totfeixr - totfnix * bcir[iy] * A(ti, ixt, iy) +
                                              cftiexclg * (-cfneut * A(fnix[g], ixt1, iy) * bcen * A(tg, ixt, iy) - (cgengpl_ * 2. * A(tg, ixt, iy) - cgpld * eion * ev) * A(ng, ixt1, iy) * vxn * A(sx, ixt1, iy) -
                                                           cmneut * A(fnix[0], ixt1, iy) * recycp[0] * cmntgpr_ * (A(ti, ixt, iy) - eidisspr_ * ev))) / (vpnorm * ennorm * A(sx, ixt1, iy));
          else
            yldot[IDXTI(ixt, iy)] = nurlxi * (totfeixr - totfnix * bcir[iy] * A(ti, ixt, iy) + cftiexclg * (-cmneut * A(fnix[0], ixt1, iy) * recycp[0] * cmntgpr_ * (A(ti, ixt, iy) - eidisspr_ * ev))) / (vpnorm * ennorm * A(sx, ixt1, iy));
        }
        if (IDXG(ixt, iy) >= 0) {  // boundary.m:2759-2799
          const int64_t ivg = IDXG(ixt, iy);
          double recy = recyrb[iy];
          if (recy > 0.) {
            double flux_inc = fac2sp * A(fnix[0], ixt1, iy);
            double t0g = mx(A(tg, ixt1, iy), tgmin * ev);
            double vxg = 0.25 * sqrt(8 * t0g / (pi * mg_[0]));
            double areapl = isoldalbarea * A(sx, ixt1, iy) + (1 - isoldalbarea) * A(sxnp, ixt1, iy);
            yldot[ivg] = nurlxg * (A(fngx, ixt1, iy) + fngxrb_use[iy] - fngxsrb[iy] + recy * flux_inc - (1 - albrb[iy]) * A(ng, ixt1, iy) * vxg * areapl) / (vpnorm * n0g_[0] * A(sx, ixt1, iy));
          } else if (recy <= 0. && recy >= -1.) {
            double t0g = mx(A(tg, ixt, iy), tgmin * ev);
            double vxg = 0.25 * sqrt(8 * t0g / (pi * mg_[0]));
            yldot[ivg] = nurlxg * (A(fngx, ixt1, iy) - (1 + recy) * A(ng, ixt, iy) * vxg * A(sx, ixt1, iy)) / (vxg * A(sx, ixt1, iy) * n0g_[0]);
          } else { errc = 5; return -4; }
        }
      }
  }
  if ((w.xcnearrb || w.openbox) && w.i6 >= ixrb + 1)  // gas temperature at the right plate (boundary.m:2880-2937)
    FOR1(iy, w.j2, w.j5) {
      const int ixt = ixrb + 1, ixt1 = IXM1(ixt, iy), ixt2 = IXM1(ixt1, iy);
      const int64_t iv = IDXTG(ixt, iy);
      if (iv < 0) continue;
      if (istgrb == 0) yldot[iv] = nurlxg * (tgwall * ev - A(tg, ixt, iy)) / (temp0 * ev);
      else if (istgrb == 1) {
        double tbound = A(tg, ixt1, iy) + A(gxf, ixt2, iy) * (A(tg, ixt1, iy) - A(tg, ixt2, iy)) / A(gxf, ixt1, iy);
        tbound = mx(tbound, 0.5 * temin * ev);
        yldot[iv] = nurlxg * (tbound - A(tg, ixt, iy)) / (temp0 * ev);
      } else if (istgrb == 3) {
        double t0 = mx(cdifg[0] * A(tg, ixt1, iy), temin * ev);
        double vxn = 0.25 * sqrt(8 * t0 / (pi * mg_[0]));
        yldot[iv] = nurlxg * (A(fegx, ixt1, iy) - 2 * cgengmpl * A(ng, ixt1, iy) * vxn * t0 * A(sx, ixt1, iy)) / (A(sx, ixt1, iy) * vpnorm * ennorm);
      } else if (istgrb == 4) {
        const double recy = recyrb[iy];
        double t0 = mx(A(tg, ixt1, iy), tgmin * ev);
        if (recy > 0.) {
          double vxn = 0.25 * sqrt(8 * t0 / (pi * mg_[0]));
          double fng_alb = (1 - albrb[iy]) * A(ng, ixt1, iy) * vxn * A(sx, ixt1, iy);
          yldot[iv] = nurlxg * (A(fegx, ixt1, iy) - cfalbedo * fng_alb * t0 + recy * (1. - cfdiss) * A(fnix[0], ixt1, iy) * recyce * cfalbedo * (kappar[iy] * zi[0] * A(te, ixt, iy) + A(ti, ixt, iy))) / (vpnorm * ennorm * A(sx, ixt1, iy));
        } else if (recy >= -1.) {
          double vyn = sqrt(0.5 * t0 / (pi * mg_[0]));
          double fng_alb = (1 + recy) * A(ng, ixt1, iy) * vyn * A(sx, ixt1, iy);
          yldot[iv] = nurlxg * (A(fegx, ixt1, iy) - cfalbedo * fng_alb * t0) / (vpnorm * ennorm * A(sx, ixt1, iy));
        } else yldot[iv] = nurlxg * (A(fegx, ixt1, iy) - cfalbedo * A(fnix[iigsp], ixt1, iy) * t0) / (vpnorm * ennorm * A(sx, ixt1, iy));
      } else yldot[iv] = nurlxg * (A(ti, ixt, iy) * cftgtipltr - A(tg, ixt, iy)) / (temp0 * ev);  // istgrb = 5
    }
  return 0;
}

// ---- rscalf (oderhs.m:8059-8213), isflxvar = 0 --------------------------------------------------------------------------
HD void rscalf(const Win& w, const double* yl, double* yldot) {
  FOR2(iy, w.j2, w.j5, ix, w.i2, w.i5) {
      double nbedot = 0., nbidot = 0., nbgdot = 0.;
      for (int f = 0; f < nisp; ++f) {
        const int64_t iv = IDXN(f, ix, iy);
        if (iv < 0) continue;
        if (isupgon == 1 && zi[f] == 0) nbgdot = yldot[iv] * n0[f];
        else nbidot = nbidot + yldot[iv] * n0[f];
        nbedot = nbedot + zi[f] * yldot[iv] * n0[f];
      }
      double nbg2dot = 0.;
      if (IDXG(ix, iy) >= 0) nbg2dot = yldot[IDXG(ix, iy)] * n0g_[0];
      for (int f = 0; f < nusp; ++f) {
        const int64_t iv2 = IDXU(f, ix, iy);
        if (iv2 < 0) continue;
        const int ix1 = IXP1(ix, iy);
        if (ALG(iv2) == 0 && IDXN(f, ix, iy) >= 0) {
          const int64_t iv = IDXN(f, ix, iy), iv1 = IDXN(f, ix1, iy);
          double yldot_np1 = A(resco[f], ix1, iy) / (A(vol, ix1, iy) * n0[f]);
          double nbvdot, nbv;
          if (ALG(iv) == 1) { nbvdot = (isnupdot1sd == 0) ? yldot_np1 * n0[f] : yldot[iv1] * n0[f]; nbv = A(ni[f], ix1, iy); }
          else if (ALG(iv1) == 1) { nbvdot = yldot[iv] * n0[f]; nbv = A(ni[f], ix, iy); }
          else { nbvdot = (isnupdot1sd == 0) ? 0.5 * (yldot[iv] + yldot_np1) * n0[f] : yldot[iv] * n0[f]; nbv = 0.5 * (A(ni[f], ix, iy) + A(ni[f], ix1, iy)); }
          yldot[iv2] = (yldot[iv2] * n0[f] - yl[iv2] * nbvdot) / nbv;
        }
      }
      if (isflxvar == 0) {
        const int64_t ive = IDXTE(ix, iy);
        if (ive >= 0 && ALG(ive) == 0) yldot[ive] = (yldot[ive] * nnorm - yl[ive] * nbedot) / A(ne, ix, iy);
        const int64_t ivi = IDXTI(ix, iy);
        if (ivi >= 0 && ALG(ivi) == 0) {
          if (isupgon == 1) yldot[ivi] = (yldot[ivi] * nnorm - yl[ivi] * (nbidot + cftiexclg * nbgdot)) / (A(nit, ix, iy) + cftiexclg * A(ni[1], ix, iy));
          else yldot[ivi] = (yldot[ivi] * nnorm - yl[ivi] * (nbidot + cngtgx[0] * nbg2dot)) / (A(nit, ix, iy) + cngtgx[0] * A(ng, ix, iy));
        }
        const int64_t ivg = IDXTG(ix, iy);  // oderhs.m:8181-8191 (isupgon = 1)
        if (ivg >= 0 && ALG(ivg) == 0) yldot[ivg] = (yldot[ivg] * n0g_[0] - yl[ivg] * nbgdot) / A(ni[iigsp], ix, iy);
      }
    }
}

// ---- pandf1 (oderhs.m:7883-8056) -------------------------------------------------------------------------------------
HD int pandf1(int xc, int yc, const double* yl, double* yldot) {
  int rc = pandf(xc, yc, yl, yldot);
  if (rc) return rc;
  const Win w = make_win(xc, yc);
  if (isflxvar != 1 && isrscalf == 1) rscalf(w, yl, yldot);
  if (dtreal < 1.e15 && yl[neq] < 0) {  // svrpkg = "nksol", oderhs.m:7963-8053 (fdt*xy = 0)
    int j2l, j5l, i2l, i5l;
    if (isbcwdt == 0) { j2l = 1; j5l = ny; i2l = 1; i5l = nx; } else { j2l = 0; j5l = ny + 1; i2l = 0; i5l = nx + 1; }
    auto step = [&](int64_t iv) { if (iv >= 0) { yldot[iv] = (1. - 0.) * yldot[iv]; yldot[iv] = yldot[iv] - (yl[iv] - ylodt[iv]) / dtuse[iv]; } };
    FOR2(iy, j2l, j5l, ix, i2l, i5l) {
        for (int f = 0; f < nisp; ++f) step(IDXN(f, ix, iy));
        if (ix != nx + 2 * isbcwdt) for (int f = 0; f < nusp; ++f) step(IDXU(f, ix, iy));
        step(IDXTE(ix, iy)); step(IDXTI(ix, iy)); step(IDXG(ix, iy)); step(IDXTG(ix, iy));
        if (isbcwdt == 1) step(IDXPHI(ix, iy));
      }
    if (dtphi < 1e10)
      FOR2(iy, 0, ny + 1, ix, 0, nx + 1) { const int64_t iv = IDXPHI(ix, iy); if (iv >= 0) yldot[iv] = yldot[iv] - (yl[iv] - ylodt[iv]) / dtphi; }
  }
  return 0;
}
  // ---- slab layout -----------------------------------------------------------------------------------------------------
  double *vyce[2], *vycb[2], *vycp[2], *veycb, *v2ce[2], *v2cb[2], *ve2cb, *ve2cd, *v2cd[2], *vycf, *vycr, *wjdote, *fmity[2], *fqymi_[2];
  double *segc, *floxge, *floyge, *conxge, *conyge, *fegx, *fegy, *fegxy, *reseg;  // gas energy equation (engbalg, oderhs.m:7508-7878)  // drift velocities (oderhs.m:1167-1420), Joule heating, inertia-current work planes
  HD static int nplanes() {
    int n = 0;
#define P1(x) n += 1;
#define P2(x) n += 2;
    UE_GEN_PLANES(P1, P2)
#undef P1
#undef P2
    return n;
  }
  // band-rows-only storage: every field plane holds `nrows` rows starting at mesh row r0 (plane pointer offset by -r0 rows, so that the
  // A(a,ix,iy) indexing is unchanged inside the band), the line arrays follow with `line` doubles each
  HD void assign_planes_band(double* slab, int r0, int nrows, int nline_planes, int line) {
    const size_t pl = (size_t)nrows * NXS;
    const int nfield = nplanes() - nline_planes;
    int k = 0;
    const ptrdiff_t off = -(ptrdiff_t)r0 * NXS;
#define P1(x) { x = (k < nfield) ? slab + (size_t)k * pl + off : slab + (size_t)nfield * pl + (size_t)(k - nfield) * line; ++k; }
#define P2(x) { x[0] = (k < nfield) ? slab + (size_t)k * pl + off : slab + (size_t)nfield * pl + (size_t)(k - nfield) * line; ++k; x[1] = (k < nfield) ? slab + (size_t)k * pl + off : slab + (size_t)nfield * pl + (size_t)(k - nfield) * line; ++k; }
    UE_GEN_PLANES(P1, P2)
#undef P1
#undef P2
  }
  HD void assign_planes(double* slab) {  // slab: nplanes() x NC doubles
    size_t k = 0;
#define P1(x) x = slab + (k++) * (size_t)NC;
#define P2(x) x[0] = slab + (k++) * (size_t)NC; x[1] = slab + (k++) * (size_t)NC;
    UE_GEN_PLANES(P1, P2)
#undef P1
#undef P2
  }
};
