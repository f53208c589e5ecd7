// Rigid-body forcing grids on the device: Lagrangian node positions / velocities from the body
// state and the force / torque sums back to the body, so that a coupled step moves 21 scalars between pyelastica's
// host state and the GPU instead of the (dim, N) position and velocity arrays.
//
//   kinematics : r_g = R r_l (rotate) or r_g as stored (spheres);  X = com + r_g;  V = v_com + omega_g x r_g
//   transfer   : S_f = sum_i f_i,  S_t = sum_i r_g,i x f_i   (the caller applies -1 and the director, as the
//                reference does on its 3-vectors)
// Grid fields are float64 like the reference's (immersed_body_forcing_grid.py:24-25); the Lagrangian forcing is of
// the flow's real_t. R, omega_g, com, v_com are tiny host-side products of the body's collections and travel as
// kernel arguments.
// ref: sopht/simulator/immersed_body/rigid_body/rigid_body_forcing_grids.py:28-78 (2-D cylinder), :128-169 (3-D
//      rigid body), :291-300 (sphere)
#include "common.cuh"

namespace sopht {
namespace {

struct BodyArgs {
  double R[9];      // r_g = R r_l, row major
  double com[3], vel[3], omega[3];
};

__global__ void __launch_bounds__(256)
    rigid_kinematics_kernel(double* pos, int64_t pos_s, double* vel, int64_t vel_s, double* rg, int64_t rg_s,
                            const double* rl, int64_t rl_s, int dim, int64_t n, BodyArgs b) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double g[3] = {0.0, 0.0, 0.0};
    if (rl) {
      double l[3] = {0.0, 0.0, 0.0};
      for (int d = 0; d < dim; ++d) l[d] = rl[d * rl_s + i];
      for (int d = 0; d < dim; ++d) {
        // same accumulation order as np.dot on a (dim, dim) x (dim, N) product
        double acc = 0.0;
        for (int e = 0; e < dim; ++e) acc += b.R[d * 3 + e] * l[e];
        g[d] = acc;
        rg[d * rg_s + i] = acc;
      }
    } else {
      for (int d = 0; d < dim; ++d) g[d] = rg[d * rg_s + i];
    }
    for (int d = 0; d < dim; ++d) pos[d * pos_s + i] = b.com[d] + g[d];
    const double cx = b.omega[1] * g[2] - b.omega[2] * g[1];
    const double cy = b.omega[2] * g[0] - b.omega[0] * g[2];
    const double cz = b.omega[0] * g[1] - b.omega[1] * g[0];
    vel[i] = b.vel[0] + cx;
    vel[vel_s + i] = b.vel[1] + cy;
    if (dim == 3) vel[2 * vel_s + i] = b.vel[2] + cz;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    rigid_force_sum_kernel(double* out6, const double* rg, int64_t rg_s, const T* f, int64_t f_s, int dim,
                           int64_t n) {
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    double r[3] = {0, 0, 0}, q[3] = {0, 0, 0};
    for (int d = 0; d < dim; ++d) r[d] = rg[d * rg_s + i], q[d] = (double)f[d * f_s + i];
    s[0] += q[0], s[1] += q[1], s[2] += q[2];
    s[3] += r[1] * q[2] - r[2] * q[1];
    s[4] += r[2] * q[0] - r[0] * q[2];
    s[5] += r[0] * q[1] - r[1] * q[0];
  }
  __shared__ double part[8][6];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int c = 0; c < 6; ++c) {
    double v = s[c];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if (lane == 0) part[warp][c] = v;
  }
  __syncthreads();
  if (threadIdx.x < 6) {
    double v = 0.0;
    for (int w = 0; w < 8; ++w) v += part[w][threadIdx.x];
    atomicAdd(out6 + threadIdx.x, v);
  }
}

int check_grid_field(const char* fn, const sopht_field_t* f, int dim, int64_t n) {
  if (!f || !f->data || f->ndim != 2 || f->shape[0] != dim || f->shape[1] != n || (n > 1 && f->stride[1] != 1))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected a (%d, N) array with contiguous rows", fn, dim);
  return SOPHT_OK;
}

}  // namespace
}  // namespace sopht

using namespace sopht;

extern "C" {

int sopht_rigid_forcing_grid_kinematics(int dim, const sopht_field_t* position_field,
                                        const sopht_field_t* velocity_field,
                                        const sopht_field_t* global_frame_relative_position_field,
                                        const sopht_field_t* local_frame_relative_position_field,
                                        const double* rotation, const double* centre, const double* velocity,
                                        const double* global_frame_omega, void* stream) {
  if (dim != 2 && dim != 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dim must be 2 or 3", __func__);
  if (!position_field || !rotation || !centre || !velocity || !global_frame_omega)
    SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null argument", __func__);
  const int64_t n = position_field->ndim == 2 ? position_field->shape[1] : -1;
  int rc;
  if ((rc = check_grid_field(__func__, position_field, dim, n))) return rc;
  if ((rc = check_grid_field(__func__, velocity_field, dim, n))) return rc;
  if ((rc = check_grid_field(__func__, global_frame_relative_position_field, dim, n))) return rc;
  if (local_frame_relative_position_field &&
      (rc = check_grid_field(__func__, local_frame_relative_position_field, dim, n)))
    return rc;
  if (n == 0) return SOPHT_OK;
  BodyArgs b;
  for (int q = 0; q < 9; ++q) b.R[q] = rotation[q];
  for (int q = 0; q < 3; ++q) b.com[q] = centre[q], b.vel[q] = velocity[q], b.omega[q] = global_frame_omega[q];
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 4) blocks = 148 * 4;
  cudaStream_t st = as_stream(stream);
  SOPHT_PROF("ib.rigid_grid_kinematics", st);
  rigid_kinematics_kernel<<<blocks, 256, 0, st>>>(
      reinterpret_cast<double*>(position_field->data), position_field->stride[0],
      reinterpret_cast<double*>(velocity_field->data), velocity_field->stride[0],
      reinterpret_cast<double*>(global_frame_relative_position_field->data),
      global_frame_relative_position_field->stride[0],
      local_frame_relative_position_field ? reinterpret_cast<const double*>(local_frame_relative_position_field->data)
                                          : nullptr,
      local_frame_relative_position_field ? local_frame_relative_position_field->stride[0] : 0, dim, n, b);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_rigid_forcing_grid_force_sums(int forcing_dtype, int dim,
                                        const sopht_field_t* global_frame_relative_position_field,
                                        const sopht_field_t* lag_grid_forcing_field, void* sums_out, void* stream) {
  SOPHT_CHECK_DTYPE(forcing_dtype);
  if (dim != 2 && dim != 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dim must be 2 or 3", __func__);
  if (!sums_out || !lag_grid_forcing_field) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null argument", __func__);
  const int64_t n = lag_grid_forcing_field->ndim == 2 ? lag_grid_forcing_field->shape[1] : -1;
  int rc;
  if ((rc = check_grid_field(__func__, global_frame_relative_position_field, dim, n))) return rc;
  if ((rc = check_grid_field(__func__, lag_grid_forcing_field, dim, n))) return rc;
  cudaStream_t st = as_stream(stream);
  SOPHT_PROF("ib.rigid_grid_force_sums", st);
  SOPHT_CUDA(cudaMemsetAsync(sums_out, 0, 6 * sizeof(double), st));
  if (n == 0) return SOPHT_OK;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148) blocks = 148;
  auto* out = reinterpret_cast<double*>(sums_out);
  const auto* rg = reinterpret_cast<const double*>(global_frame_relative_position_field->data);
  const int64_t rg_s = global_frame_relative_position_field->stride[0], f_s = lag_grid_forcing_field->stride[0];
  if (forcing_dtype == SOPHT_F32)
    rigid_force_sum_kernel<float><<<blocks, 256, 0, st>>>(
        out, rg, rg_s, reinterpret_cast<const float*>(lag_grid_forcing_field->data), f_s, dim, n);
  else
    rigid_force_sum_kernel<double><<<blocks, 256, 0, st>>>(
        out, rg, rg_s, reinterpret_cast<const double*>(lag_grid_forcing_field->data), f_s, dim, n);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

}  // extern "C"
