/*
 * sopht_b200.h — C ABI of libsopht_b200.so
 *
 * B200 (sm_100a) implementation of the SophT Eulerian flow time step.  Every
 * entry point replaces one kernel (or kernel family) that the reference builds
 * at run time with pystencils / pyFFTW / numba; the citation after each
 * declaration names the reference interface it stands in for
 * (paths relative to the SophT source tree, sopht/numeric/...).
 *
 * Conventions
 *  - plain C: pointers + sizes only, no torch / C++ types.
 *  - all pointers in sopht_field_t are DEVICE pointers; the caller owns the
 *    memory.  Work is enqueued on `stream` (a cudaStream_t passed as void*)
 *    and the call returns without synchronising.
 *  - fields are strided views: shape/stride are in ELEMENTS of the dtype
 *    (for complex fields: in complex elements), slowest axis first, x last.
 *    Scalar fields are (nz, ny, nx) / (ny, nx); vector fields (dim, nz, ny, nx).
 *  - scalars travel as double and are rounded to the kernel dtype inside.
 *  - return value: 0 on success, negative sopht_status_t otherwise; a text
 *    description of the last failure is available from sopht_last_error().
 *    Nothing throws or aborts across this boundary.
 */
#ifndef SOPHT_B200_H
#define SOPHT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SOPHT_MAX_DIMS 5

typedef enum {
  SOPHT_F32 = 0,
  SOPHT_F64 = 1
} sopht_dtype_t;

typedef enum {
  SOPHT_OK = 0,
  SOPHT_ERR_DTYPE = -1,   /* dtype is neither SOPHT_F32 nor SOPHT_F64            */
  SOPHT_ERR_SHAPE = -2,   /* operand shapes / ndim inconsistent                   */
  SOPHT_ERR_STRIDE = -3,  /* stride pattern not supported by this entry point     */
  SOPHT_ERR_ARG = -4,     /* bad scalar argument (width, order, axis, ...)        */
  SOPHT_ERR_CUDA = -5,    /* a CUDA runtime call or kernel launch failed          */
  SOPHT_ERR_CUFFT = -6,   /* a cuFFT call failed                                  */
  SOPHT_ERR_ALLOC = -7,   /* device allocation failed                             */
  SOPHT_ERR_HANDLE = -8   /* null / stale handle                                  */
} sopht_status_t;

typedef struct {
  void *data;                      /* device pointer to element [0,...,0]          */
  int32_t ndim;                    /* 1..SOPHT_MAX_DIMS                            */
  int64_t shape[SOPHT_MAX_DIMS];   /* slowest axis first                           */
  int64_t stride[SOPHT_MAX_DIMS];  /* element strides (may be any sign-positive)   */
} sopht_field_t;

const char *sopht_last_error(void);
int sopht_version(void);
/* number of kernel launches issued through this library since load (for bench accounting) */
int64_t sopht_launch_count(void);

/* Per-kernel CUDA-event timers around the instrumented launch sites (off by default; used by bench.py
 * for the roofline of the dominant kernel). sopht_profile_report() synchronises the recorded events,
 * clears them, and returns a JSON object {"label": {"launches": n, "ms": total}, ...} owned by the library. */
int sopht_profile_enable(int on);
const char *sopht_profile_report(void);
/* the same timers around work the host layer enqueues itself (NCCL collectives of the slab path):
 * begin returns a token for end; both are no-ops (token -1) while the timers are off */
int sopht_profile_range_begin(const char *label, void *stream);
int sopht_profile_range_end(int token, void *stream);

/* ------------------------------------------------------------------------ */
/* Elementwise operations on strided views (1..4-D, 2D and 3D grids alike)   */
/* ------------------------------------------------------------------------ */

/* field[...] = fixed_val
 * ref: eulerian_grid_ops/stencil_ops_3d/elementwise_ops_3d.py:60-119 (gen_set_fixed_val_pyst_kernel_3d),
 *      stencil_ops_2d/elementwise_ops_2d.py:60-114 */
int sopht_set_fixed_val(int dtype, const sopht_field_t *field, double fixed_val, void *stream);

/* vector_field[c, ...] = fixed_vals[c]  (leading axis = component)
 * ref: elementwise_ops_3d.py:95-115 (vector closure) */
int sopht_set_fixed_vals_vector(int dtype, const sopht_field_t *vector_field,
                                const double *fixed_vals, int n_vals, void *stream);

/* field[...] = rhs_field[...]
 * ref: elementwise_ops_3d.py:122-143 (gen_elementwise_copy_pyst_kernel_3d) */
int sopht_elementwise_copy(int dtype, const sopht_field_t *field, const sopht_field_t *rhs_field,
                           void *stream);

/* sum_field = field_1 + field_2 (aliasing sum_field == field_k allowed)
 * ref: elementwise_ops_3d.py:13-57 (gen_elementwise_sum_pyst_kernel_3d) */
int sopht_elementwise_sum(int dtype, const sopht_field_t *sum_field, const sopht_field_t *field_1,
                          const sopht_field_t *field_2, void *stream);

/* sum_field = field_1_prefac * field_1 + field_2_prefac * field_2
 * ref: elementwise_ops_3d.py:337-387 (gen_elementwise_saxpby_pyst_kernel_3d) */
int sopht_elementwise_saxpby(int dtype, const sopht_field_t *sum_field,
                             const sopht_field_t *field_1, const sopht_field_t *field_2,
                             double field_1_prefac, double field_2_prefac, void *stream);

/* sum_field = field + fixed_val
 * ref: elementwise_ops_3d.py:271-334 (gen_add_fixed_val_pyst_kernel_3d) */
int sopht_add_fixed_val(int dtype, const sopht_field_t *sum_field, const sopht_field_t *field,
                        double fixed_val, void *stream);

/* sum_field[c] = vector_field[c] + fixed_vals[c]
 * ref: elementwise_ops_3d.py:305-329 */
int sopht_add_fixed_vals_vector(int dtype, const sopht_field_t *sum_field,
                                const sopht_field_t *vector_field, const double *fixed_vals,
                                int n_vals, void *stream);

/* product = field_1 * field_2 on complex fields (interleaved re/im; strides in complex elements)
 * ref: elementwise_ops_3d.py:146-197 (gen_elementwise_complex_product_pyst_kernel_3d) */
int sopht_elementwise_complex_product(int dtype, const sopht_field_t *product_field,
                                      const sopht_field_t *field_1, const sopht_field_t *field_2,
                                      void *stream);

/* result = field_1 x field_2, (3, nz, ny, nx) operands
 * ref: elementwise_ops_3d.py:390-449 (gen_elementwise_cross_product_pyst_kernel_3d) */
int sopht_elementwise_cross_product_3d(int dtype, const sopht_field_t *result_field,
                                       const sopht_field_t *field_1, const sopht_field_t *field_2,
                                       void *stream);

/* ring of `width` cells on every grid face <- fixed value(s).  `field` is a scalar grid field
 * (2-D / 3-D) when is_vector == 0, a vector field (dim leading) when is_vector == 1.
 * ref: elementwise_ops_3d.py:200-268, elementwise_ops_2d.py:193-253 */
int sopht_set_fixed_val_at_boundaries(int dtype, const sopht_field_t *field, int width,
                                      const double *fixed_vals, int is_vector, void *stream);

/* penalised = (field + f*chi*penalty) / (1 + f*chi)
 * ref: stencil_ops_3d/brinkmann_penalise_3d.py:13-83, stencil_ops_2d/brinkmann_penalise_2d.py:13-76 */
int sopht_brinkmann_penalise(int dtype, const sopht_field_t *penalised_field,
                             const sopht_field_t *field, const sopht_field_t *char_field,
                             const sopht_field_t *penalty_field, double penalty_factor,
                             void *stream);

/* same with a spatially constant penalty value
 * ref: stencil_ops_2d/brinkmann_penalise_2d.py:79-141 */
int sopht_brinkmann_penalise_vs_fixed_val(int dtype, const sopht_field_t *penalised_field,
                                          const sopht_field_t *field,
                                          const sopht_field_t *char_field, double penalty_val,
                                          double penalty_factor, void *stream);

/* smooth sine Heaviside of a level set
 * ref: stencil_ops_3d/char_func_from_level_set_3d.py:12-51, stencil_ops_2d/char_func_from_level_set_2d.py:12-51 */
int sopht_char_func_from_level_set(int dtype, const sopht_field_t *char_func_field,
                                   const sopht_field_t *level_set_field, double blend_width,
                                   void *stream);

/* velocity_magnitude = sum_c |velocity[c]| (written, as the reference does) and the device scalar
 * *max_out = max over cells (max_out: device pointer to one element of dtype).
 * ref: simulator/flow/passive_transport_flow_simulators.py:139-155 */
int sopht_abs_sum_max(int dtype, const sopht_field_t *velocity_magnitude_field,
                      const sopht_field_t *velocity_field, void *max_out, void *stream);

/* ------------------------------------------------------------------------ */
/* 3D stencils (pystencils ghost-ring rule: only cells whose whole stencil   */
/* is in bounds are written; `reset_ghost_zone` zeroes the ring afterwards)  */
/* ------------------------------------------------------------------------ */

/* flux = prefactor * (sum of 6 neighbours - 6 f); scalar (nz,ny,nx) or vector (3,nz,ny,nx)
 * ref: stencil_ops_3d/diffusion_flux_3d.py:14-115 */
int sopht_diffusion_flux_3d(int dtype, const sopht_field_t *diffusion_flux,
                            const sopht_field_t *field, double prefactor, int reset_ghost_zone,
                            void *stream);

/* curl = prefactor * centred-difference curl (no 1/(2dx))
 * ref: stencil_ops_3d/curl_3d.py:13-132 */
int sopht_curl_3d(int dtype, const sopht_field_t *curl, const sopht_field_t *field,
                  double prefactor, int reset_ghost_zone, void *stream);

/* divergence = 0.5 * inv_dx * centred-difference divergence
 * ref: stencil_ops_3d/divergence_3d.py:13-96 */
int sopht_divergence_3d(int dtype, const sopht_field_t *divergence, const sopht_field_t *field,
                        double inv_dx, int reset_ghost_zone, void *stream);

/* vorticity += prefactor * curl_c(velocity_forcing) on the ring-1 interior
 * ref: stencil_ops_3d/update_vorticity_from_velocity_forcing_3d.py:12-132 */
int sopht_update_vorticity_from_velocity_forcing_3d(int dtype, const sopht_field_t *vorticity_field,
                                                    const sopht_field_t *velocity_forcing_field,
                                                    double prefactor, void *stream);

/* vorticity += prefactor * curl_c(penalised_velocity - velocity)
 * ref: update_vorticity_from_velocity_forcing_3d.py:135-291 */
int sopht_update_vorticity_from_penalised_velocity_3d(int dtype,
                                                      const sopht_field_t *vorticity_field,
                                                      const sopht_field_t *penalised_velocity_field,
                                                      const sopht_field_t *velocity_field,
                                                      double prefactor, void *stream);

/* flux_c = prefactor * (omega . grad_c) u_c, ring <- 0
 * ref: stencil_ops_3d/vorticity_stretching_flux_3d.py:13-107 */
int sopht_vorticity_stretching_flux_3d(int dtype, const sopht_field_t *flux_field,
                                       const sopht_field_t *vorticity_field,
                                       const sopht_field_t *velocity_field, double prefactor,
                                       void *stream);

/* advection_flux += conservative ENO3 flux divergence (six face terms), ring-2 interior
 * ref: stencil_ops_3d/advection_flux_3d.py:12-233 */
int sopht_advection_flux_eno3_3d(int dtype, const sopht_field_t *advection_flux,
                                 const sopht_field_t *field, const sopht_field_t *velocity,
                                 double inv_dx, void *stream);

/* filter_flux = 0.25 * (-f[-1] + 2 f - f[+1]) along `axis` (0 = x, 1 = y, 2 = z), ring-1 interior
 * ref: stencil_ops_3d/laplacian_filter_3d.py:58-77 */
int sopht_laplacian_filter_flux_3d(int dtype, const sopht_field_t *filter_flux,
                                   const sopht_field_t *field, int axis, void *stream);

/* The whole "convolution" Laplacian filter of a scalar (nz, ny, nx) or vector (3, nz, ny, nx) field in place: per
 * component and per direction x, y, z:  f -= M_d^order f,  M_d g = 0.25 (-g[+1] - g[-1] + 2 g) on the ring-1 interior
 * and 0 on the ring - what the reference's closure computes with 5 + 4 order passes per direction, here one read and
 * one write per direction (shared-memory line kernels, csrc/filter3d.cu). scratch: one scalar field of the same grid
 * (the factory's field_buffer), clobbered. Unit x-stride required (SOPHT_ERR_STRIDE otherwise: the caller falls back
 * to the pass-by-pass composition). filter_order in [1, 64].
 * ref: stencil_ops_3d/laplacian_filter_3d.py:129-163 (convolution closure), :58-80 (the 1-D stencils) */
int sopht_laplacian_filter_convolution_3d(int dtype, const sopht_field_t *field, const sopht_field_t *scratch,
                                          int filter_order, void *stream);

/* sine-ramp penalisation of a `width`-cell ring, applied x then y then z.
 * ramp_{x,y,z}: HOST arrays of 2*width factors each (front ramp then back ramp), in the kernel dtype's
 * value range, precomputed by the caller from the grid coordinates exactly as the reference does.
 * ref: stencil_ops_3d/penalise_field_boundary_3d.py:13-240 */
int sopht_penalise_field_boundary_3d(int dtype, const sopht_field_t *field, int width,
                                     const double *ramp_x, const double *ramp_y,
                                     const double *ramp_z, void *stream);

/* The same on a z-slab of a z-decomposed grid: `field` is the slab (nz_local planes), z_faces bit 0 / bit 1
 * say whether the slab's low / high z face is the global boundary (3 = whole grid). A face that is not
 * global is treated as interior. ref: penalise_field_boundary_3d.py:182-208 (the z part applies on the
 * boundary ranks only; the x and y parts on every rank) */
int sopht_penalise_field_boundary_3d_slab(int dtype, const sopht_field_t *field, int width,
                                          const double *ramp_x, const double *ramp_y,
                                          const double *ramp_z, int z_faces, void *stream);

/* ------------------------------------------------------------------------ */
/* 2D stencils                                                               */
/* ------------------------------------------------------------------------ */

/* ref: stencil_ops_2d/diffusion_flux_2d.py:13-72 */
int sopht_diffusion_flux_2d(int dtype, const sopht_field_t *diffusion_flux,
                            const sopht_field_t *field, double prefactor, int reset_ghost_zone,
                            void *stream);
/* ref: stencil_ops_2d/advection_flux_2d.py:12-165 */
int sopht_advection_flux_eno3_2d(int dtype, const sopht_field_t *advection_flux,
                                 const sopht_field_t *field, const sopht_field_t *velocity,
                                 double inv_dx, void *stream);
/* curl (2,ny,nx) of a scalar stream function: u_x = p dpsi/dy, u_y = -p dpsi/dx
 * ref: stencil_ops_2d/outplane_field_curl_2d.py:13-100 */
int sopht_outplane_field_curl_2d(int dtype, const sopht_field_t *curl, const sopht_field_t *field,
                                 double prefactor, int reset_ghost_zone, void *stream);
/* scalar curl of a (2,ny,nx) field, no ring reset
 * ref: stencil_ops_2d/inplane_field_curl_2d.py:10-50 */
int sopht_inplane_field_curl_2d(int dtype, const sopht_field_t *curl, const sopht_field_t *field,
                                double prefactor, void *stream);
/* ref: stencil_ops_2d/update_vorticity_from_velocity_forcing_2d.py:12-71 */
int sopht_update_vorticity_from_velocity_forcing_2d(int dtype, const sopht_field_t *vorticity_field,
                                                    const sopht_field_t *velocity_forcing_field,
                                                    double prefactor, void *stream);
/* ref: stencil_ops_2d/update_vorticity_from_velocity_forcing_2d.py:74-147 */
int sopht_update_vorticity_from_penalised_velocity_2d(int dtype,
                                                      const sopht_field_t *vorticity_field,
                                                      const sopht_field_t *penalised_velocity_field,
                                                      const sopht_field_t *velocity_field,
                                                      double prefactor, void *stream);
/* ref: stencil_ops_2d/penalise_field_boundary_2d.py:12-142 */
int sopht_penalise_field_boundary_2d(int dtype, const sopht_field_t *field, int width,
                                     const double *ramp_x, const double *ramp_y, void *stream);

/* ------------------------------------------------------------------------ */
/* Unbounded Poisson solver (Hockney-Eastwood doubled domain, FFT)           */
/* ------------------------------------------------------------------------ */

typedef struct sopht_poisson *sopht_poisson_t;

#define SOPHT_POISSON_AUTO 0          /* fused power-of-two path when eligible, else generic */
#define SOPHT_POISSON_FORCE_GENERIC 1 /* cuFFT batched 2-D + strided 1-D path (any size, f32/f64) */

/* Builds plans, workspaces and the Green's function spectrum G_hat * dx^dim on the doubled grid.
 * dim = 2: (ny, nx) grids (nz ignored); dim = 3: (nz, ny, nx).
 * mz/my/mx: HOST arrays (2nz / 2ny / 2nx doubles) holding min(x, 2X - x) of the doubled-axis
 * coordinates, and origin_value the regularised G at r = 0, both computed by the caller with the
 * reference's expressions; pass NULL arrays to have them derived from x_range and dx.
 * The library owns everything behind the handle; solve() allocates nothing.
 * ref: poisson_solver_3d/UnboundedPoissonSolverPYFFTW3D.py:12-83 (ctor + Green's function),
 *      poisson_solver_2d/UnboundedPoissonSolverPYFFTW2D.py:11-68,
 *      poisson_solver_3d/FFTPyFFTW3D.py:28-56 (plans and buffers) */
int sopht_poisson_create(sopht_poisson_t *handle, int dtype, int dim, int nz, int ny, int nx,
                         double x_range, double dx, const double *mz, const double *my,
                         const double *mx, double origin_value, int flags, void *stream);

/* Handle for -laplacian(solution) = rhs with homogeneous Neumann conditions on all walls of the cell-centred grid
 * (second-order three-point Laplacian with mirror ghost cells), the mean mode of the solution set to zero: the
 * reference's fast-diagonalisation solver in closed form (DCT-II eigenbasis -> mirror extension + FFT). The handle is
 * used with sopht_poisson_solve / sopht_poisson_path / sopht_poisson_destroy; sopht_poisson_green_hat returns NULL.
 * ref: poisson_solver_3d/FastDiagPoissonSolver3D.py:15-181 (ctor, solve), :183-208 (vector_field_solve),
 *      poisson_solver_2d/FastDiagPoissonSolver2D.py:13-119 */
int sopht_poisson_neumann_create(sopht_poisson_t *handle, int dtype, int dim, int nz, int ny, int nx, double dx,
                                 void *stream);

/* Handle for -laplacian(solution) = rhs with PERIODIC boundaries in every direction, mean mode of the solution zero
 * (BASELINE config 4; an extension - the reference has no periodic solver, so there is no reference interface to
 * cite: parity is against analytic Fourier modes). three_point_symbol = 0: spectral symbol (2 pi m / L)^2;
 * != 0: the symbol of the 7-point Laplacian with wrap-around neighbours (its exact inverse). Used with
 * sopht_poisson_solve / sopht_poisson_path / sopht_poisson_destroy. */
int sopht_poisson_periodic_create(sopht_poisson_t *handle, int dtype, int dim, int nz, int ny, int nx, double dx,
                                  int three_point_symbol, void *stream);

/* -laplacian(solution) = rhs on the unbounded domain. Fields: scalar grid fields, or vector fields
 * with a leading component axis (each component solved independently).
 * ref: UnboundedPoissonSolverPYFFTW3D.py:111-172 (solve, vector_field_solve),
 *      UnboundedPoissonSolverPYFFTW2D.py:95-129 */
int sopht_poisson_solve(sopht_poisson_t handle, const sopht_field_t *solution_field,
                        const sopht_field_t *rhs_field, void *stream);

/* device pointer to Re(G_hat) * dx^dim / (doubled cell count), natural (kz, ky, kx<=nx) order, of the
 * handle's dtype (NULL if the active path does not keep it in that layout).
 * ref: UnboundedPoissonSolverPYFFTW3D.py:47-49 (fourier_greens_function_times_dx_cubed) */
int sopht_poisson_green_hat(sopht_poisson_t handle, const void **device_ptr);

/* "generic" or "pow2" */
const char *sopht_poisson_path(sopht_poisson_t handle);

int sopht_poisson_destroy(sopht_poisson_t handle);

/* ------------------------------------------------------------------------ */
/* The same solve on a z-slab decomposed grid (fp32, power-of-two grids):     */
/* the three LOCAL phases between the two all-to-all transposes, which the    */
/* host layer issues on its own process group (NCCL). Rank r owns planes      */
/* [r nz/P, (r+1) nz/P) of the fields and, for the y/z passes, the kx bins     */
/* [r nx/P, (r+1) nx/P) of the half spectrum. Buffers are caller-owned device */
/* memory of complex64 elements:                                              */
/*   send / recv  (C, P, nz/P, ny, nx/P)  ==  (C, nz, ny, nx/P) after exchange */
/*   work         2 x (C, nz, 2 ny, nx/P): x-major spectrum + the z pass's tile-major output */
/*   nyquist_local (C, nz/P, ny), nyquist_all (C, nz, ny), nyquist_work (C, nz, 2 ny) */
/* Sequence per solve (sopht_b200/parallel/slab_poisson.py):                  */
/*   forward_x -> per component all-to-all(send -> recv), all-gather(nyquist)  */
/*   -> yz -> per component all-to-all(recv -> send), slice nyquist -> inverse_x */
/* ref: UnboundedPoissonSolverPYFFTW3D.py:111-172 (what is computed); the      */
/* reference has no distributed path                            */
/* ------------------------------------------------------------------------ */
typedef struct sopht_poisson_slab *sopht_poisson_slab_t;

int sopht_poisson_slab_create(sopht_poisson_slab_t *handle, int ncomp, int nz, int ny, int nx, int nranks,
                              int rank, double dx, const double *mz, const double *my, const double *mx,
                              double origin_value, void *stream);
/* The same handle for the PERIODIC box (an extension, BASELINE config 4; nearest reference code: none): -lap(psi) = rhs
 * with the spectral ((2 pi m / L)^2) or three-point symbol, mean mode dropped. nx is the REAL grid size; rows carry
 * nx / 2 complex bins + the Nyquist bin, so the exchange buffers are (C, P, nz/P, ny, nx/2/P); y and z passes run in
 * place - work_buffer / nyquist_work of yz() are unused (may be NULL). */
int sopht_poisson_slab_create_periodic(sopht_poisson_slab_t *handle, int ncomp, int nz, int ny, int nx, int nranks,
                                       int rank, double dx, int three_point_symbol, void *stream);
/* rhs_field: this rank's (C, nz/P, ny, nx) slab (strided view allowed) */
int sopht_poisson_slab_forward_x(sopht_poisson_slab_t handle, const sopht_field_t *rhs_field,
                                 void *send_buffer, void *nyquist_local, void *stream);
/* y forward, z forward x G_hat x z inverse, y inverse on the kx-slab, in place in recv_buffer / nyquist_all */
int sopht_poisson_slab_yz(sopht_poisson_slab_t handle, void *recv_buffer, void *nyquist_all,
                          void *work_buffer, void *nyquist_work, void *stream);
int sopht_poisson_slab_inverse_x(sopht_poisson_slab_t handle, const sopht_field_t *solution_field,
                                 void *recv_buffer, void *nyquist_local, void *stream);
/* Transposes fused into the kernels over peer memory (NVLink): the library allocates the two exchange buffers,
 * enable_peer_exchange() returns their CUDA IPC handles (2 x 64 bytes: recv, send), the host layer all-gathers
 * the handles of all ranks (nranks x 128 bytes, rank order) and hands them to open_peers(). Afterwards the phase
 * functions take NULL for send_buffer / recv_buffer: x forward then stores every kx chunk straight into the
 * owning rank's buffer (push), y inverse leaves its kx-slab in this rank's second buffer and x inverse reads
 * chunk q of every spectrum row from rank q's buffer (pull), and no all-to-all is issued - the host layer only
 * separates the phases with a barrier. */
int sopht_poisson_slab_enable_peer_exchange(sopht_poisson_slab_t handle, unsigned char *ipc_handles_out);
int sopht_poisson_slab_open_peers(sopht_poisson_slab_t handle, const unsigned char *all_ipc_handles);
/* Pipelined variant of the peer-exchange solve (needs enable_peer_exchange / open_peers): every call works on ONE
 * component, the transposes are cudaMemcpy2DAsync copies (copy engines, no SM) between the ranks' exchange blocks, so
 * the host layer can run the transposes of one component under the y / z passes of another:
 *   pipe_forward_x(c)  -> pipe_transpose(c, 0, streams)  -> [all ranks: barrier] -> pipe_yz(c)
 *   -> pipe_transpose(c, 1, streams) -> [barrier] -> pipe_inverse_x(c)
 * pipe_transpose enqueues this rank's P copies (kx chunk q of its rows -> rank q; backward: z block p of its kx slab ->
 * rank p) round robin on `streams` (cudaStream_t array); forward it also delivers nyquist_local (C, nz/P, ny) into every
 * rank's copy of the Nyquist plane, which lives inside the exchange block. work_buffer (unbounded solve only):
 * 2 x (nz, 2 ny, nx/P) complex64, nyquist_work (nz, 2 ny). */
int sopht_poisson_slab_pipe_forward_x(sopht_poisson_slab_t handle, const sopht_field_t *rhs_field, int component,
                                      void *nyquist_local, void *stream);
int sopht_poisson_slab_pipe_transpose(sopht_poisson_slab_t handle, int component, int backward, void **streams,
                                      int nstreams, const void *nyquist_local);
int sopht_poisson_slab_pipe_yz(sopht_poisson_slab_t handle, int component, void *work_buffer, void *nyquist_work,
                               void *stream);
int sopht_poisson_slab_pipe_inverse_x(sopht_poisson_slab_t handle, const sopht_field_t *solution_field, int component,
                                      void *nyquist_local, void *stream);
int sopht_poisson_slab_destroy(sopht_poisson_slab_t handle);

/* ------------------------------------------------------------------------ */
/* Real <-> half-spectrum FFT over all axes of a (nz, ny, nx) or (ny, nx)     */
/* grid: the plan objects of FFTPyFFTW{2,3}D and the scipy rfftn/irfftn       */
/* helper. Arrays are contiguous: real (nz, ny, nx), complex (nz, ny, nx/2+1). */
/* ref: poisson_solver_3d/FFTPyFFTW3D.py:7-67, poisson_solver_2d/             */
/*      FFTPyFFTW2D.py:7-65, poisson_solver_3d/scipy_fft_3d.py:7-15           */
/* ------------------------------------------------------------------------ */
typedef struct sopht_fft *sopht_fft_t;
int sopht_fft_create(sopht_fft_t *handle, int dtype, int dim, int nz, int ny, int nx);
/* fourier_field = rfftn(field)                       (fft_plan(input_array=field, output_array=fourier_field)) */
int sopht_fft_forward(sopht_fft_t handle, const sopht_field_t *field, const sopht_field_t *fourier_field,
                      void *stream);
/* field = irfftn(fourier_field), normalised; fourier_field is destroyed like the input of an FFTW c2r plan */
int sopht_fft_inverse(sopht_fft_t handle, const sopht_field_t *fourier_field, const sopht_field_t *field,
                      void *stream);
int sopht_fft_destroy(sopht_fft_t handle);

/* ------------------------------------------------------------------------ */
/* Peer-memory arena of the z-slab decomposition (one process per GPU, one    */
/* box): field storage every rank can address over NVLink (CUDA IPC), the     */
/* halo exchange as one kernel of direct stores into the neighbours' halo     */
/* planes, and a device-side barrier. The reference has no distributed path;  */
/* what travels is the ghost ring its stencil kernels read.    */
/* ------------------------------------------------------------------------ */
typedef struct sopht_peer_arena *sopht_peer_arena_t;

/* Allocates (and zeroes) payload_bytes of device memory plus a small flag header; ipc_handle_out receives the
 * 64-byte CUDA IPC handle the host layer all-gathers (nranks x 64 bytes, rank order) for open(). */
int sopht_peer_arena_create(sopht_peer_arena_t *handle, size_t payload_bytes, int nranks, int rank,
                            unsigned char *ipc_handle_out);
int sopht_peer_arena_open(sopht_peer_arena_t handle, const unsigned char *all_ipc_handles);
/* periodic_z != 0: the ranks form a ring (the low neighbour of rank 0 is rank nranks - 1): halo exchange of a periodic
 * box (BASELINE config 4). Default 0: the global z boundaries have no neighbour. */
int sopht_peer_arena_set_periodic(sopht_peer_arena_t handle, int periodic_z);
/* device pointer to this rank's payload; the host layer lays the same objects out at the same offsets on all ranks */
void *sopht_peer_arena_payload(sopht_peer_arena_t handle);
/* Fills the z halo planes of up to 4 local arrays (ncomp[q], nz_local + 2 halo, ny, nx), given by their byte offset
 * in the payload and the byte stride between components: this rank's first / last owned planes are stored into the
 * low / high neighbour's halo planes after a ready handshake; the kernel retires when both neighbours have done the
 * same for this rank. Must be issued in the same order on every rank. */
int sopht_peer_halo_exchange(sopht_peer_arena_t handle, int nfields, const int64_t *payload_offsets_bytes,
                             const int64_t *comp_stride_bytes, const int *ncomp, int nz_local, int halo,
                             int64_t plane_bytes, void *stream);
/* all-ranks barrier in stream order (everything enqueued before it on every rank is complete and visible) */
int sopht_peer_barrier(sopht_peer_arena_t handle, void *stream);
/* The device-side polls of the two calls above are bounded (SOPHT_PEER_TIMEOUT_S, default 30 s): a rank that died or
 * issued another sequence of exchanges no longer hangs its neighbours' GPUs; the kernel gives up and records the rank
 * it was waiting for. status() synchronises, returns 0 when healthy or a CUDA error code with the stalled rank in
 * *stalled_rank_out (-1 when healthy). */
int sopht_peer_arena_status(sopht_peer_arena_t handle, int *stalled_rank_out);
int sopht_peer_arena_destroy(sopht_peer_arena_t handle);

/* ------------------------------------------------------------------------ */
/* Immersed boundary: Eulerian <-> Lagrangian transfer, 4-point delta kernels */
/* Lagrangian arrays are (dim, N) with N contiguous; nearest_index is int64;  */
/* lag_positions / body velocities are of pos_dtype (SOPHT_F32 / SOPHT_F64),  */
/* all other arrays of `dtype`. dim = 2 or 3.                                 */
/* ------------------------------------------------------------------------ */

/* nearest_index = floor((X - shift)/dx); local_support[d, taps, i] = (idx_d + off_d) dx + shift - X_d
 * ref: immersed_boundary_ops/EulerianLagrangianGridCommunicator3D.py:69-135, ...2D.py:70-134 */
int sopht_ib_local_support(int dtype, int dim, const sopht_field_t *local_support,
                           const sopht_field_t *nearest_index, const sopht_field_t *lag_positions,
                           int pos_dtype, double dx, double eul_grid_coord_shift, void *stream);

/* kernel_type 0: cosine (3D.py:387-412), 1: Peskin 2002 (3D.py:415-518). Mutates local_support as the
 * reference does (/= dx, or |.|/dx). prefactor = (0.25/dx)^dim or (0.125/dx)^dim, evaluated by the caller. */
int sopht_ib_interpolation_weights(int dtype, int dim, int kernel_type,
                                   const sopht_field_t *interp_weights,
                                   const sopht_field_t *local_support, double dx, double prefactor,
                                   void *stream);

/* lag[c, i] = dx^dim * sum_taps eul[c, taps(i)] * w[taps, i]; scalar (N,) or vector (dim, N)
 * ref: EulerianLagrangianGridCommunicator3D.py:180-295, ...2D.py */
int sopht_ib_eulerian_to_lagrangian(int dtype, int dim, const sopht_field_t *lag_grid_field,
                                    const sopht_field_t *eul_grid_field,
                                    const sopht_field_t *interp_weights,
                                    const sopht_field_t *nearest_index, double dx_pow_dim, void *stream);

/* eul[c, taps(i)] += lag[c, i] * w[taps, i] (accumulates; caller zeroes)
 * ref: EulerianLagrangianGridCommunicator3D.py:298-380, ...2D.py */
int sopht_ib_lagrangian_to_eulerian(int dtype, int dim, const sopht_field_t *eul_grid_field,
                                    const sopht_field_t *lag_grid_field,
                                    const sopht_field_t *interp_weights,
                                    const sopht_field_t *nearest_index, void *stream);

/* One launch for the whole virtual-boundary interaction (cosine kernel): support, weights, velocity
 * gather, dv = U - V_body, F = k dX + c dv, spread of F (eul_grid_forcing_field may be NULL: Lagrangian
 * part only). ref: immersed_boundary_ops/VirtualBoundaryForcing.py:187-253 */
/* running count of Lagrangian nodes that overflowed their tile's list in the atomic-free spread (see ib.cu section 6)
 * and were spread with atomics instead; synchronises the device. */
int sopht_ib_spread_stragglers(unsigned long long *count_out);

int sopht_ib_virtual_boundary_forcing(int dtype, int dim, const sopht_field_t *eul_grid_forcing_field,
                                      const sopht_field_t *eul_grid_velocity_field,
                                      const sopht_field_t *lag_positions,
                                      const sopht_field_t *lag_body_velocity, int pos_dtype,
                                      const sopht_field_t *local_support,
                                      const sopht_field_t *interp_weights,
                                      const sopht_field_t *nearest_index,
                                      const sopht_field_t *lag_flow_velocity,
                                      const sopht_field_t *lag_velocity_mismatch,
                                      const sopht_field_t *lag_position_mismatch,
                                      const sopht_field_t *lag_forcing, double dx,
                                      double eul_grid_coord_shift, double weight_prefactor,
                                      double dx_pow_dim, double stiffness, double damping, void *stream);

/* ------------------------------------------------------------------------ */
/* Rigid-body forcing grids. Grid fields are float64 (dim, N) */
/* device arrays; the body state travels as host 3-vectors / a 3x3 matrix.    */
/* ------------------------------------------------------------------------ */

/* r_g = rotation * r_l (local_frame_relative_position_field may be NULL: r_g is used as stored, the sphere's case),
 * position = centre + r_g, velocity = velocity + global_frame_omega x r_g. rotation: row-major 3x3 (director^T).
 * ref: simulator/immersed_body/rigid_body/rigid_body_forcing_grids.py:28-55 (2-D), :128-149 (3-D), :291-300 */
int sopht_rigid_forcing_grid_kinematics(int dim, const sopht_field_t *position_field,
                                        const sopht_field_t *velocity_field,
                                        const sopht_field_t *global_frame_relative_position_field,
                                        const sopht_field_t *local_frame_relative_position_field,
                                        const double *rotation, const double *centre, const double *velocity,
                                        const double *global_frame_omega, void *stream);
/* sums_out (device, 6 doubles) = [sum_i f_i (3), sum_i r_g,i x f_i (3)]; the caller negates and applies the
 * director like the reference (:57-78, :151-169). lag_grid_forcing_field is of forcing_dtype. */
int sopht_rigid_forcing_grid_force_sums(int forcing_dtype, int dim,
                                        const sopht_field_t *global_frame_relative_position_field,
                                        const sopht_field_t *lag_grid_forcing_field, void *sums_out, void *stream);

/* ------------------------------------------------------------------------ */
/* Cosserat-rod forcing grids. grid_kind: 0 nodal, 1 element  */
/* centric, 2 edge (dim 2 only), 3 surface (dim 3 only). rod_state is a DEVICE */
/* block of sopht_rod_state_doubles(n_elems) doubles, rows contiguous:         */
/*   position_collection (3, n+1) | velocity_collection (3, n+1) | mass (n+1)  */
/*   | director_collection (3, 3, n) | omega_collection (3, n) | radius (n)    */
/*   | tangents (3, n)                                                          */
/* ------------------------------------------------------------------------ */

int64_t sopht_rod_state_doubles(int64_t n_elems);
/* position_field / velocity_field: float64 (dim, N_lag) device arrays, N_lag = n+1 | n | 3n | number of surface
 * nodes. moment_arm: float64 (3, n) for the edge grid, (3, N_lag) for the surface grid (written), NULL otherwise.
 * Surface tables (device): node_element int32 (N_lag), local_points float64 (2, N_lag) = (cos, sin) of the node's
 * angle or (0, 0) for a centre node, radius_ratio float64 (N_lag).
 * ref: simulator/immersed_body/cosserat_rod/cosserat_rod_forcing_grids.py:25-33 (nodal), :96-109 (element centric),
 *      :179-237 (edge), :408-467 (surface) */
int sopht_rod_forcing_grid_kinematics(int grid_kind, int dim, int64_t n_elems, const void *rod_state,
                                      const sopht_field_t *position_field, const sopht_field_t *velocity_field,
                                      const sopht_field_t *moment_arm, const void *surface_node_element,
                                      const void *surface_local_points, const void *surface_radius_ratio,
                                      void *stream);
/* forces_torques_out (device, 3 (n+1) + 3 n doubles) = body_flow_forces (3, n+1) | body_flow_torques (3, n), the
 * torques in the elements' material frames; the element-centric grid leaves the torques at 0 (the reference does not
 * touch them). moment_arm: (3, n) written by the nodal grid, read by the edge grid; (3, N_lag) read by the surface
 * grid; NULL for the element-centric grid. surface_element_start: int32 (n+1) windows of the surface nodes.
 * ref: cosserat_rod_forcing_grids.py:35-73, :111-124, :239-284, :469-497 */
int sopht_rod_forcing_grid_transfer(int forcing_dtype, int grid_kind, int dim, int64_t n_elems, const void *rod_state,
                                    const sopht_field_t *lag_grid_forcing_field, const sopht_field_t *moment_arm,
                                    const void *surface_element_start, void *forces_torques_out, void *stream);

/* ------------------------------------------------------------------------ */
/* Fused passes of the 3-D Navier-Stokes step (simulator-level path).         */
/* Vector fields (3, nz, ny, nx) with unit x-stride; outputs must not alias   */
/* inputs (neighbouring cells are read).                                      */
/* ------------------------------------------------------------------------ */

/* out = w + prefactor * curl_c(u x w) on the ring-1 interior, out = w on the ring: the cross product and the
 * vorticity update of the rotational-form advection in one pass.
 * ref: simulator/flow/navier_stokes_flow_simulators.py:454-465
 *      (elementwise_ops_3d.py:390-449 + update_vorticity_from_velocity_forcing_3d.py:12-132) */
int sopht_ns3d_advect_rotational(int dtype, const sopht_field_t *out_vorticity_field,
                                 const sopht_field_t *vorticity_field,
                                 const sopht_field_t *velocity_field, double prefactor, void *stream);

/* out = f + nu_dt_by_dx2 * Lap_7pt(f) on the ring-1 interior, out = f on the ring, all three components;
 * zero_field (may be NULL) is set to 0 in the same pass (the forcing-field reset of the forced step).
 * ref: navier_stokes_flow_simulators.py:466-471, :498; stencil_ops_3d/diffusion_timestep_3d.py:12-80 */
int sopht_ns3d_diffuse(int dtype, const sopht_field_t *out_field, const sopht_field_t *field,
                       double nu_dt_by_dx2, const sopht_field_t *zero_field, void *stream);

/* The same pass followed by the sine penalisation of the boundary ring for the default width 2, where the reference's
 * copy-and-scale reduces to a per-axis factor: ramp_{x,y,z} are DEVICE arrays of nx / ny / nz factors of `dtype`
 * ({0, sin(pi/4), 1, ..., 1, sin(pi/4), 0}), applied in the reference's order x, y, z. Needs 16-byte aligned unit-stride
 * rows (otherwise SOPHT_ERR_STRIDE: call sopht_ns3d_diffuse + sopht_penalise_field_boundary_3d).
 * ref: navier_stokes_flow_simulators.py:466-478; stencil_ops_3d/penalise_field_boundary_3d.py:182-208 */
int sopht_ns3d_diffuse_penalise(int dtype, const sopht_field_t *out_field, const sopht_field_t *field,
                                double nu_dt_by_dx2, const sopht_field_t *zero_field, const void *ramp_x,
                                const void *ramp_y, const void *ramp_z, void *stream);

/* velocity = prefactor * curl_c(psi) (ring <- 0) + free_stream_velocity (HOST array of 3, may be NULL);
 * if max_abs_sum_out (device pointer to one element of dtype) is not NULL it receives max_cells sum_c |u_c|,
 * the reduction the stable-timestep estimate needs.
 * ref: navier_stokes_flow_simulators.py:479-485 (curl_3d.py:13-132, elementwise_ops_3d.py:305-329),
 *      passive_transport_flow_simulators.py:139-155 */
int sopht_ns3d_velocity_from_stream_function(int dtype, const sopht_field_t *velocity_field,
                                             const sopht_field_t *stream_func_field, double prefactor,
                                             const double *free_stream_velocity, void *max_abs_sum_out,
                                             void *stream);

/* The same three passes for a PERIODIC box (BASELINE config 4; an extension - the reference has no periodic case,
 * nearest reference step: navier_stokes_flow_simulators.py:449-485 without the penalisation). x and y neighbours wrap
 * around inside the kernels and every row / cell of a plane is updated; z keeps the ghost-ring rule, i.e. the arrays are
 * (3, nz + 2, ny, nx) with one halo plane per z side that the caller fills before the call: sopht_wrap_z_halos on one
 * GPU, the neighbour rank's planes (sopht_peer_halo_exchange, periodic ring) in a slab decomposition. Rows must be
 * 16-byte multiples (no fallback kernel: anything else is refused). */
int sopht_ns3d_advect_rotational_periodic_xy(int dtype, const sopht_field_t *out_vorticity_field,
                                             const sopht_field_t *vorticity_field,
                                             const sopht_field_t *velocity_field, double prefactor, void *stream);
int sopht_ns3d_diffuse_periodic_xy(int dtype, const sopht_field_t *out_field, const sopht_field_t *field,
                                   double nu_dt_by_dx2, const sopht_field_t *zero_field, void *stream);
int sopht_ns3d_velocity_from_stream_function_periodic_xy(int dtype, const sopht_field_t *velocity_field,
                                                         const sopht_field_t *stream_func_field, double prefactor,
                                                         const double *free_stream_velocity, void *max_abs_sum_out,
                                                         void *stream);
/* field (C, nz + 2, ny, nx) or (nz + 2, ny, nx): plane 0 <- plane nz, plane nz + 1 <- plane 1 (one launch) */
int sopht_wrap_z_halos(int dtype, const sopht_field_t *field, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SOPHT_B200_H */
