/* include/ue_gpu.h — C ABI of the B200 pandf1 / jac_calc drop-in.
 *
 * These entry points are what an ISO_C_BINDING shim compiled into the
 * reference's `bbb` package binds.  They replace, one for
 * one, the bodies of the two reference interface routines
 *
 *   Pandf1rhs_interface(neq,time,yl,yldot)              bbb/oderhs.m:12331-12345
 *   jac_calc_interface(neq,t,yl,yldot00,ml,mu,wk,
 *                      nnzmx,jac,ja,ia)                 bbb/oderhs.m:12297-12329
 *
 * and carry across the ambient module state those routines read (the named
 * inputs listed in include/ue_params.h).
 *
 * Conventions (reference: Forthon builds with 8-byte default INTEGER and REAL):
 *   - all integers are int64_t, all reals are double;
 *   - arrays are caller-owned, Fortran-contiguous; index *content* (ja, ia,
 *     igyl, ixm1, ...) is 1-based / Fortran-valued exactly as in the reference;
 *   - every function returns 0 on success, <0 on error; the message is
 *     available from ue_gpu_last_error() so the shim can `call xerrab(msg)`
 *     (reference error path: com/error.f:1-13);
 *   - there is NO CPU fallback: if no CUDA device is present ue_gpu_init fails.
 */
#ifndef UE_GPU_H
#define UE_GPU_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- static inputs (once per allocate/ueinit; names in ue_params.h) -------- */
int ue_gpu_set_int(const char* name, int64_t value);   /* after ue_gpu_init a CHANGED switch requires a new ue_gpu_init (model_dt excepted) */
int ue_gpu_set_real(const char* name, double value);
int ue_gpu_set_real_array(const char* name, const double* data, int64_t n);   /* arrays are uploaded by ue_gpu_init: call it again after changing one */
int ue_gpu_set_int_array(const char* name, const int64_t* data, int64_t n);
/* Validate that every named input is present and the switch set is one the
 * kernels implement, upload to the device, build the window tables. */
int ue_gpu_init(void);

/* ---- per-solve inputs (change every exmain / psetnk) ----------------------
 * dtuse, ylodt (group Time_dep_nwt, bbb/bbb.v), suscal, sfscal (group Ynorm):
 * length neq each.  Scalars nufak/dtreal/dtphi come through ue_gpu_set_real. */
int ue_gpu_step_params(int64_t neq, const double* dtuse, const double* ylodt,
                       const double* suscal, const double* sfscal);

/* ---- set_dt(neq,yl,f0) of the nksol driver (bbb/oderhs.m:9886-10147, called from bbb/odesolve.m:299): f0 = rhsnk(yl),
 * then dtuse(iv) from ylodt (last ue_gpu_step_params), deldt, dtreal and model_dt (0..3).  dtuse stays on the device
 * for the calls that follow and is also returned (neq doubles) for the Fortran module array. */
int ue_gpu_set_dt(int64_t neq, const double* yl, double* f0, double* dtuse);

/* ---- Timing group of the reference (com/com.v:500-519): seconds spent in ue_gpu_pandf1 (ttotfe) and ue_gpu_jac_calc (ttotjf)
 * since the last reset, host clock around the calls; ttjstor is 0 (storing is fused into the assembly kernels).  The shim adds
 * them to the module variables after each call or once per exmain. */
int ue_gpu_timing(double* ttotfe, double* ttotjf, double* ttjstor, int64_t reset);

/* ---- vnormnk(n, v, s) = sqrt(sum((v(i)*s(i))**2)) of NKSOL (svr/nksol.m:1404-1419: fnrm, unrm, pnrm).  The sum has a fixed
 * shape (1024 strided partial sums + binary tree): deterministic, within a few ulp of the reference's serial sum.
 * ue_gpu_fnrm: the same norm of the residual the last ue_gpu_pandf1 left on the device, with the resident sfscal
 * (nksol.m:1009 fnrm = vnormnk(n, savf, sf)) - nothing crosses the bus but the result. */
int ue_gpu_vnormnk(int64_t neq, const double* v, const double* s, double* out);
int ue_gpu_fnrm(double* out);

/* ---- Pandf1rhs_interface: pandf1(-1,-1,0,neq,time,yl,yldot) ---------------
 * yl has neq+2 entries (yl(neq+1) = Jacobian-mode flag, yl(neq+2) = nufak);
 * yldot receives neq entries. */
int ue_gpu_pandf1(int64_t neq, double time, const double* yl, double* yldot);

/* ---- jac_calc_interface ----------------------------------------------------
 * Finite-difference Jacobian of pandf1 about yl (bbb/oderhs.m:8533-8760) in the
 * reference's CSR layout: jac/ja (nnz, columns ascending within a row, 1-based)
 * and ia (neq+1, 1-based, ia(neq+1)=nnz+1).  yldot00 = pandf1(yl) as computed
 * by the caller (psetnk / sfsetnk), neq+2 long.  yl is left unchanged.
 * Returns -2 with the reference's "More storage needed" message if nnz>nnzmx. */
int ue_gpu_jac_calc(int64_t neq, double t, const double* yl, const double* yldot00,
                    int64_t ml, int64_t mu, int64_t nnzmx,
                    double* jac, int64_t* ja, int64_t* ia, int64_t* nnz_out);

/* ---- sfsetnk (bbb/oderhs.m:9815-9884) with the Jacobian kept on the device --------------------------
 * f0 = pandf1(yl | yl(neq+1)=1); J = jac_calc(yl,f0); J <- J diag(1/su) (amudia); sf(i) = 1/max_k|J_ik| (rnrms,
 * normtype 0); ydt_max0 = max(cutlo, max_i |f0_i sf_i|).  Returns -7 with the reference's "Jacobian row = 0"
 * message if a row vanishes.  Uses dtuse/sfscal of the last ue_gpu_step_params. */
int ue_gpu_sfsetnk(int64_t neq, const double* yl, const double* su, int64_t ml, int64_t mu, double* sf, double* ydt_max0);

/* ---- rhsnk(yl) + jac_calc(yl, yldot00) in one call ------------------------------------------------------------------
 * The pair psetnk and sfsetnk issue back to back (bbb/oderhs.m:9466-9468, 9851-9857): yldot00 (neq) = pandf1(yl) and the
 * Jacobian at yl, with one upload of yl and one synchronisation.  Same outputs and errors as the two separate calls. */
int ue_gpu_rhs_jac(int64_t neq, const double* yl, double* yldot00, int64_t ml, int64_t mu, int64_t nnzmx,
                   double* jac, int64_t* ja, int64_t* ia, int64_t* nnz_out);

/* ---- psetnk's scaling chain (bbb/oderhs.m:9473-9485) on the Jacobian the last ue_gpu_jac_calc left on the device ----
 * J <- J diag(1/su) (amudia, svr/svrut4.m:1104-1148); J <- diag(sf) J (diamua, :1054-1102); if isrnorm==1 the rows are
 * normalised with normtype 0/1/2 = max/1/2-norm (jac_norm_rows -> roscal -> rnrms, oderhs.m:8959-8992, svrut4.m:954-1052)
 * and the factors returned in fnormnw.  The pattern (ja, ia) is the one ue_gpu_jac_calc returned; only the nnz values
 * come back, ready for jac_lu_decomp.  su, sf, fnormnw: neq doubles; jac: nnz doubles. */
int ue_gpu_jac_scale(int64_t neq, const double* su, const double* sf, int64_t isrnorm, int64_t normtype,
                     int64_t nnz, double* jac, double* fnormnw);

/* ---- guard for array inputs the built path does not read --------------------------------------------------------------
 * Volume sources (volpsor, volmsor, pwrsore, pwrsori, voljcsor: bbb/oderhs.m:3407-3456, 4300-4330), user profiles
 * (dif_use, kye_use, kyi_use, tray_use, ...) and wall sources are assumed to vanish.  The shim calls this once per array
 * after ueinit; a non-zero element returns -5 with a message naming the array (-> xerrab) instead of being ignored. */
int ue_gpu_assert_zero(const char* name, const double* a, int64_t n);

/* ---- parity probe: include/ue_math.h evaluated on the device -----------------------------------------------------------
 * op 0 exp, 1 log, 2 log10, 3 pow(x,y), 4 cos, 5 sqrt; host arrays of n doubles.  The CPU checker evaluates the same
 * header; the two must agree bit for bit (that is what makes the value-dependent Jacobian pattern reproducible). */
int ue_gpu_math_probe(int64_t op, int64_t n, const double* x, const double* y, double* out);

/* ---- optional: page-lock caller arrays -------------------------------------------------------------------------------
 * If yl / yldot (NKSOL's work arrays) and jac / ja / ia (group Jacobian, bbb.v:2861-2873) are page-locked - by these
 * calls or by the caller's own cudaHostAlloc / cudaHostRegister - the kernels read yl and write yldot, jac, ja, ia in
 * place over the bus and the entry points above issue no copies.  Pageable arrays work unchanged (staged copies).
 * Unpin before the array is freed. */
int ue_gpu_pin_host_array(void* p, int64_t bytes);
int ue_gpu_unpin_host_array(void* p);

/* ---- device-resident variants (inputs/outputs already in HBM) -------------
 * Same semantics; pointers are device pointers on the current device.  Used by
 * bench.py for the kernel-only figure and by a host that keeps yl on the GPU. */
int ue_gpu_pandf1_dev(int64_t neq, double time, const double* d_yl, double* d_yldot);
/* rhsnk(yl) followed by jac_calc(yl, yldot00) - the pair psetnk/sfsetnk issue (oderhs.m:9466-9468, 9851-9857) - as one
 * stream sequence without a host round trip in between; *ms = CUDA-event time of the sequence on the library stream. */
int ue_gpu_rhs_jac_dev(int64_t neq, const double* d_yl, double* d_yldot00, int64_t ml, int64_t mu, int64_t nnzmx,
                       double* d_jac, int64_t* d_ja, int64_t* d_ia, int64_t* nnz_out, double* ms);
int ue_gpu_jac_calc_dev(int64_t neq, double t, const double* d_yl, const double* d_yldot00,
                        int64_t ml, int64_t mu, int64_t nnzmx,
                        double* d_jac, int64_t* d_ja, int64_t* d_ia, int64_t* nnz_out);

/* Device-pointer callers only: assert (flag=1) that d_yl is unchanged since the last ue_gpu_pandf1_dev, so the
 * next ue_gpu_jac_calc_dev reuses the base fields instead of re-evaluating phases 0-2.  The host-pointer entry
 * points detect this themselves (psetnk/sfsetnk call rhsnk(yl) right before jac_calc, oderhs.m:9466, 9851). */
int ue_gpu_assume_base_current(int64_t flag);

/* ---- column-range split (ppp MPISplitIndex / LocalJacBuilder analogue) ----
 * Restrict the next jac_calc calls to perturbed unknowns iv in [ivmin,ivmax]
 * (1-based, inclusive): the returned CSR holds only those columns.  (1,neq)
 * restores the full Jacobian.  Reference: ppp/parallel.F90:176-381. */
int ue_gpu_set_column_range(int64_t ivmin, int64_t ivmax);

/* ---- multi-GPU: ONE Jacobian assembled by several GPUs (ppp jac_calc_mpi: MPIJacBuilder + MPICollectBroadCastJacobian,
 * ppp/mpi_parallel.F90:2-447) ------------------------------------------------------------------------------------------
 * One host process per GPU (the reference's MPI ranks), every rank holding the full state.  After ue_gpu_comm_init every
 * ue_gpu_jac_calc / ue_gpu_rhs_jac / *_dev call assembles only this rank's contiguous range of columns (MPISplitIndex;
 * ranges balanced by the columns' candidate-list sizes), all-gathers the CSC fragments over NCCL on the library's stream
 * (NVLink / NVSwitch) and transposes the full matrix on every rank: each rank returns the FULL CSR, as after the
 * reference's MPI_BCAST.  All ranks must make the same calls in the same order.
 *   ue_gpu_comm_unique_id : rank 0 obtains the 128-byte NCCL id; the host distributes it (MPI_Bcast in the Fortran host).
 *   ue_gpu_comm_init      : collective; call after ue_gpu_init.  nranks = 1 is allowed (no communication).
 *   ue_gpu_comm_info      : this rank's column range and the NCCL bytes it moved in the last Jacobian.
 *   ue_gpu_comm_finalize  : back to single-GPU operation.
 *   ue_gpu_comm_p2p_handle / ue_gpu_comm_init_p2p : the same split WITHOUT a collective call.  Every rank exports its slot
 *                           arrays (64-byte CUDA IPC handle), the host gathers the handles (MPI_Allgather), every rank maps
 *                           the others' arrays; the assembly kernel then stores each column result into all GPUs over
 *                           NVLink and a device-side flag barrier replaces the all-reduce.  Ranks of ONE node (<= 8);
 *                           synchronise the ranks on the host between ue_gpu_comm_init_p2p and the first Jacobian. */
int ue_gpu_comm_p2p_handle(char* handle64);
int ue_gpu_comm_init_p2p(int64_t nranks, int64_t rank, const char* handles /* nranks x 64 bytes, rank order */);
int ue_gpu_comm_unique_id(char* id128);
int ue_gpu_comm_init(int64_t nranks, int64_t rank, const char* id128);
int ue_gpu_comm_info(int64_t* nranks, int64_t* rank, int64_t* ivmin, int64_t* ivmax, int64_t* bytes_last_jac);
int ue_gpu_comm_finalize(void);

/* CSC copy of the last Jacobian - the arrays rcsc / icsc / jcsc that jac_calc leaves in group Jacobian_csc (bbb/oderhs.m:8620-8752,
 * read by jacmap and the ppp debug dumps): values, 1-based row numbers (ascending within a column), column pointers (neq+1).
 * Not part of the hot path: built on the host from the column fragments on request. */
int ue_gpu_get_csc(int64_t nnzmx, double* rcsc, int64_t* icsc, int64_t* jcsc, int64_t* nnz);

/* Device buffers owned by the library (yl, yldot, yldot00: neq+2; jac/ja: fragment capacity; ia: neq+1),
 * for callers that keep the state resident between calls. */
int ue_gpu_device_buffers(double** yl, double** yldot, double** yldot00, double** jac, int64_t** ja, int64_t** ia);
/* Copy one intermediate field plane (ids: enum Plane in uedge_b200/csrc/ue_device.cuh) to the host. */
int ue_gpu_get_plane(int64_t plane, double* out);

/* ---- diagnostics ----------------------------------------------------------- */
int ue_gpu_kernel_launches(int64_t* n);      /* launches since init (for bench.py) */
int ue_gpu_last_kernel_ms(double* jac_ms, double* res_ms); /* CUDA-event times of the last calls */
const char* ue_gpu_last_error(void);
int ue_gpu_finalize(void);

#ifdef __cplusplus
}
#endif
#endif /* UE_GPU_H */
