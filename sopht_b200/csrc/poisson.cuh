// Internal interface shared by the Poisson solver implementations.
#pragma once
#include "common.cuh"

namespace sopht {

struct PoissonImpl {
  virtual ~PoissonImpl() {}
  // solution / rhs: scalar (grid dims) or vector (leading component axis) fields of the handle's dtype
  virtual int solve(const sopht_field_t* sol, const sopht_field_t* rhs, cudaStream_t st) = 0;
  // device pointer to Re(G_hat)*dx^dim/(doubled cell count) on (n2z, n2y, nx+1), natural order (or null)
  virtual const void* green_hat() const { return nullptr; }
  virtual const char* path_name() const { return "generic"; }
};

PoissonImpl* make_generic_poisson(int dtype, int dim, int nz, int ny, int nx, double dx,
                                  const double* mz, const double* my, const double* mx,
                                  double origin, cudaStream_t st, int* rc);

// Re(G_hat) * dx^dim / (doubled cell count) on (2nz, 2ny, nx+1), natural order (device, caller frees)
template <typename T>
int build_green_hat(T** g_out, int dim, int nz, int ny, int nx, double dx, const double* mz_h,
                    const double* my_h, const double* mx_h, double origin_value, cudaStream_t st);

// fp32, 3-D: the folded spectrum (nz+1, ny+1, nxl) of the kx range [kx0, kx0 + nxl) and the kx = nx plane
// (nz+1, ny+1), both x2 like fold_green_kernel's, built plane batch by plane batch (poisson_generic.cu)
int build_green_folded_slice(float* gm, float* gn, int nz, int ny, int nx, int kx0, int nxl, double dx,
                             const double* mz_h, const double* my_h, const double* mx_h, double origin_value,
                             cudaStream_t st);

// fp32, 3-D, power-of-two grids: pruned + fused shared-memory FFT pipeline (poisson_pow2.cu)
bool pow2_poisson_eligible(int dtype, int dim, int nz, int ny, int nx);
PoissonImpl* make_pow2_poisson(int nz, int ny, int nx, double dx, const double* mz, const double* my,
                               const double* mx, double origin, cudaStream_t st, int* rc);

// Homogeneous Neumann walls on the cell-centred grid (the reference's FastDiagPoissonSolver{2,3}D): mirror
// extension + periodic three-point symbol (poisson_neumann.cu)
PoissonImpl* make_neumann_poisson(int dtype, int dim, int nz, int ny, int nx, double dx, cudaStream_t st, int* rc);
// the same solve through same-length DCT-II transforms (poisson_neumann_dct.cu): 3-D grids with even extents
bool neumann_dct_eligible(int dim, int nz, int ny, int nx);
PoissonImpl* make_neumann_dct_poisson(int dtype, int nz, int ny, int nx, double dx, cudaStream_t st, int* rc);
// Periodic in every direction (an extension, BASELINE config 4): same pipeline without the mirror step
PoissonImpl* make_periodic_poisson(int dtype, int three_point_symbol, int dim, int nz, int ny, int nx, double dx,
                                   cudaStream_t st, int* rc);

// fp32, 3-D, power-of-two grids: the periodic solve on the hand-written FFT pipeline (poisson_pow2.cu)
bool periodic_pow2_eligible(int dtype, int dim, int nz, int ny, int nx);
PoissonImpl* make_periodic_pow2_poisson(int three_point_symbol, int nz, int ny, int nx, double dx, cudaStream_t st,
                                        int* rc);

}  // namespace sopht
