// C entry points of the unbounded Poisson solver handles (2-D and 3-D).
#include <math.h>

#include <vector>

#include "poisson.cuh"

using namespace sopht;

struct sopht_poisson {
  int dtype, dim, nz, ny, nx;
  PoissonImpl* impl;
};

// min(x, 2X - x) on the doubled axis, x_i = linspace(0, 2X - dx, 2n)[i]; used when the caller does not
// supply the arrays (the Python layer always does, computed with the reference's own numpy expressions).
static void default_reflected_axis(std::vector<double>& m, int n, double range, double dx) {
  m.resize(2 * n);
  const double stop = 2 * range - dx;
  for (int i = 0; i < 2 * n; ++i) {
    const double x = 2 * n > 1 ? stop * i / (2 * n - 1) : 0.0;
    m[i] = fmin(x, 2 * range - x);
  }
}

extern "C" {

int sopht_poisson_create(sopht_poisson_t* handle, int dtype, int dim, int nz, int ny, int nx,
                         double x_range, double dx, const double* mz, const double* my,
                         const double* mx, double origin_value, int flags, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (!handle) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null handle pointer", __func__);
  if (dim != 2 && dim != 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dim must be 2 or 3", __func__);
  if (ny <= 0 || nx <= 0 || (dim == 3 && nz <= 0))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid sizes must be positive", __func__);
  if (!(dx > 0)) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dx must be positive", __func__);
  std::vector<double> dz_, dy_, dx_;
  if (!mx || !my || (dim == 3 && !mz)) {
    default_reflected_axis(dx_, nx, x_range, dx);
    default_reflected_axis(dy_, ny, x_range * ((double)ny / nx), dx);
    if (dim == 3) default_reflected_axis(dz_, nz, x_range * ((double)nz / nx), dx);
    mx = dx_.data();
    my = dy_.data();
    mz = dim == 3 ? dz_.data() : nullptr;
    const double pi = 3.14159265358979323846;
    origin_value = dim == 3 ? 1.0 / (4 * pi * dx) : -(2 * log(dx / sqrt(pi)) - 1) / (4 * pi);
  }
  int rc = SOPHT_OK;
  PoissonImpl* impl = nullptr;
  if (flags != SOPHT_POISSON_FORCE_GENERIC && pow2_poisson_eligible(dtype, dim, nz, ny, nx))
    impl = make_pow2_poisson(nz, ny, nx, dx, mz, my, mx, origin_value, as_stream(stream), &rc);
  else
    impl = make_generic_poisson(dtype, dim, nz, ny, nx, dx, mz, my, mx, origin_value,
                                as_stream(stream), &rc);
  if (!impl) return rc;
  sopht_poisson* h = new sopht_poisson{dtype, dim, dim == 3 ? nz : 1, ny, nx, impl};
  *handle = h;
  return SOPHT_OK;
}

int sopht_poisson_neumann_create(sopht_poisson_t* handle, int dtype, int dim, int nz, int ny, int nx, double dx,
                                 void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (!handle) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null handle pointer", __func__);
  if (dim != 2 && dim != 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dim must be 2 or 3", __func__);
  if (ny <= 0 || nx <= 0 || (dim == 3 && nz <= 0))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid sizes must be positive", __func__);
  if (!(dx > 0)) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dx must be positive", __func__);
  int rc = SOPHT_OK;
  PoissonImpl* impl = make_neumann_poisson(dtype, dim, nz, ny, nx, dx, as_stream(stream), &rc);
  if (!impl) return rc;
  *handle = new sopht_poisson{dtype, dim, dim == 3 ? nz : 1, ny, nx, impl};
  return SOPHT_OK;
}

int sopht_poisson_periodic_create(sopht_poisson_t* handle, int dtype, int dim, int nz, int ny, int nx, double dx,
                                  int three_point_symbol, void* stream) {
  SOPHT_CHECK_DTYPE(dtype);
  if (!handle) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null handle pointer", __func__);
  if (dim != 2 && dim != 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dim must be 2 or 3", __func__);
  if (ny <= 0 || nx <= 0 || (dim == 3 && nz <= 0))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: grid sizes must be positive", __func__);
  if (!(dx > 0)) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dx must be positive", __func__);
  int rc = SOPHT_OK;
  PoissonImpl* impl =
      periodic_pow2_eligible(dtype, dim, nz, ny, nx)
          ? make_periodic_pow2_poisson(three_point_symbol, nz, ny, nx, dx, as_stream(stream), &rc)
          : make_periodic_poisson(dtype, three_point_symbol, dim, nz, ny, nx, dx, as_stream(stream), &rc);
  if (!impl) return rc;
  *handle = new sopht_poisson{dtype, dim, dim == 3 ? nz : 1, ny, nx, impl};
  return SOPHT_OK;
}

int sopht_poisson_solve(sopht_poisson_t h, const sopht_field_t* solution_field,
                        const sopht_field_t* rhs_field, void* stream) {
  if (!h || !h->impl) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle", __func__);
  if (!valid_field(solution_field, h->dim, h->dim + 1) || !valid_field(rhs_field, h->dim, h->dim + 1) ||
      !same_shape(solution_field, rhs_field))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: solution and rhs must have the same scalar/vector grid shape",
               __func__);
  const int nd = solution_field->ndim;
  const int64_t* shp = solution_field->shape + (nd - h->dim);
  const bool ok = h->dim == 3 ? (shp[0] == h->nz && shp[1] == h->ny && shp[2] == h->nx)
                              : (shp[0] == h->ny && shp[1] == h->nx);
  if (!ok) SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: field shape does not match the solver's grid", __func__);
  return h->impl->solve(solution_field, rhs_field, as_stream(stream));
}

int sopht_poisson_green_hat(sopht_poisson_t h, const void** device_ptr) {
  if (!h || !h->impl || !device_ptr) SOPHT_FAIL(SOPHT_ERR_HANDLE, "%s: null handle", __func__);
  *device_ptr = h->impl->green_hat();
  return SOPHT_OK;
}

const char* sopht_poisson_path(sopht_poisson_t h) {
  return (h && h->impl) ? h->impl->path_name() : "";
}

int sopht_poisson_destroy(sopht_poisson_t h) {
  if (!h) return SOPHT_OK;
  delete h->impl;
  delete h;
  return SOPHT_OK;
}

}  // extern "C"
