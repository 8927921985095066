/* include/ue_gen.h — C ABI of the GENERAL path of libuegpu.so (uedge_b200/csrc/ue_gen.cu, ue_gen_phys.h).
 *
 * Same two compute entry points as include/ue_gpu.h, for the wider switch set of the hydrogen family:
 * inertial atoms as ion species 2 (isupgon=1, nhsp=2: pyexamples/box2, pyexamples/input_example), non-orthogonal
 * meshes (isnonog=1), every fd2tra scheme, the potential equation (isphion=1) and any subset of equations.
 *
 *   ue_gen_pandf1   replaces Pandf1rhs_interface(neq, time, yl, yldot)                      bbb/oderhs.m:8217-8254
 *   ue_gen_jac_calc replaces jac_calc_interface(neq, t, yl, yldot00, ml, mu, wk, nnzmx, jac, ja, ia)
 *                                                                                          bbb/oderhs.m:8533-8760 (+ csrcsc, svr/svrut4.m:1536-1608)
 * Static state crosses once through ONE generic setter under the reference's own variable names (every scalar and
 * array as doubles; species-indexed scalars as arrays; 2-D arrays as [iy][ix] planes, nx+2 fastest), then ue_gen_init
 * validates the switch set (anything outside the built family is refused by name) and uploads it.  The list of names is
 * what uedge_b200/case2.py:Case2.inputs2() produces (the Fortran shim walks the same list: INTEGRATION.md 6).
 * All entry points return 0 on success; ue_gen_last_error() explains a non-zero return.  There is no CPU fallback:
 * ue_gen_init fails without a CUDA device.
 */
#ifndef UE_GEN_H
#define UE_GEN_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
int ue_gen_clear(void);
int ue_gen_set(const char* name, const double* data, int64_t n);
int ue_gen_init(void);
int ue_gen_finalize(void);
/* dtuse, ylodt (time-step term, oderhs.m:7963-8053), suscal, sfscal (perturbation size and clip test of jac_calc) */
int ue_gen_step_params(int64_t neq, const double* dtuse, const double* ylodt, const double* suscal, const double* sfscal);
/* yl: neq + 2 values (yl(neq+1) is the reference's flag: < 0 adds the time-step term, 1 adds nufak to the diagonal) */
int ue_gen_pandf1(int64_t neq, double time, const double* yl, double* yldot);
/* columns ivmin..ivmax only (ppp column split, ppp/omp_parallel.F90:65-117); default 1..neq */
int ue_gen_set_column_range(int64_t ivmin, int64_t ivmax);
/* jac, ja: nnzmx entries; ia: neq + 1; 1-based CSR, columns ascending within a row */
int ue_gen_jac_calc(int64_t neq, double t, const double* yl, const double* yldot00, int64_t ml, int64_t mu, int64_t nnzmx, double* jac, int64_t* ja,
                    int64_t* ia, int64_t* nnz);
/* one intermediate field plane of the last full evaluation by name ("fnix1", "feex", "resphi", ...): (ny+2) x (nx+2) doubles */
int ue_gen_get_plane(const char* name, double* out);
/* ONE Jacobian over the GPUs of a node (ppp jac_calc_mpi, ppp/mpi_parallel.F90:2-447): one process per GPU, every rank with the full
 * state.  Collective, after ue_gen_init on every rank; id128 is the NCCL id rank 0 obtained from ue_gpu_comm_unique_id and the host
 * distributed (MPI_Bcast).  Afterwards ue_gen_jac_calc evaluates only this rank's share of the columns, exchanges the column fragments
 * with the other ranks on the device (grouped in-place ncclBroadcast) and returns the FULL CSR on every rank.  ue_gen_last_comm_ms:
 * CUDA-event time of that exchange in the last call. */
int ue_gen_comm_init(int64_t nranks, int64_t rank, const char* id128);
int ue_gen_comm_finalize(void);
int ue_gen_last_comm_ms(double* comm_ms);
/* CUDA-event times (ms) of the kernels of the last ue_gen_pandf1 / ue_gen_jac_calc: residual, Jacobian columns, CSR transpose */
int ue_gen_last_kernel_ms(double* resid_ms, double* cols_ms, double* csr_ms);
const char* ue_gen_last_error(void);
#ifdef __cplusplus
}
#endif
#endif
