// Cosserat-rod forcing grids on the device: the four Lagrangian grids the reference hangs on a
// pyelastica rod (nodal, element-centric, edge (2-D), surface (3-D)). The rod's state is O(n_elems) doubles and is
// uploaded once per call as one packed block; the (dim, N_lag) position / velocity fields stay on the device and feed
// VirtualBoundaryForcing directly, and the transfer returns the rod's (3, n+1) forces and (3, n) torques in one block.
//
//   packed rod state (doubles, n = n_elems, all rows contiguous):
//     position_collection (3, n+1) | velocity_collection (3, n+1) | mass (n+1) | director_collection (3, 3, n)
//     | omega_collection (3, n) | radius (n) | tangents (3, n)                       -> 7 (n + 1) + 16 n doubles
//
//   kinematics : one thread per Lagrangian node
//   transfer   : one warp per rod node ("slot"); the warp reduces the forcing of the (at most two) elements that touch
//                the node, so the result is deterministic (no atomics) and needs no second pass
// ref: sopht/simulator/immersed_body/cosserat_rod/cosserat_rod_forcing_grids.py:10-79 (nodal), :82-131 (element
//      centric), :134-289 (edge), :292-503 (surface); pyelastica 0.3.x helpers restated from their published
//      definitions: _node_to_element_velocity (mass-weighted mean of the two end nodes), _elements_to_nodes_inplace
//      (half of an element quantity to each end node), _batch_matvec, _batch_cross.
#include "common.cuh"

namespace sopht {
namespace {

enum RodGridMode { ROD_NODAL = 0, ROD_ELEMENT = 1, ROD_EDGE = 2, ROD_SURFACE = 3 };

struct RodView {
  const double *x, *v, *m, *q, *w, *r, *t;
  int64_t n;
  __host__ __device__ explicit RodView(const double* s, int64_t n_) : n(n_) {
    x = s;
    v = x + 3 * (n + 1);
    m = v + 3 * (n + 1);
    q = m + (n + 1);
    w = q + 9 * n;
    r = w + 3 * n;
    t = r + n;
  }
  __device__ double X(int d, int64_t k) const { return x[d * (n + 1) + k]; }
  __device__ double V(int d, int64_t k) const { return v[d * (n + 1) + k]; }
  __device__ double Q(int i, int j, int64_t e) const { return q[(i * 3 + j) * n + e]; }
  __device__ double W(int d, int64_t e) const { return w[d * n + e]; }
  __device__ double T(int d, int64_t e) const { return t[d * n + e]; }
  // 0.5 (x[e+1] + x[e]), cosserat_rod_forcing_grids.py:98-101
  __device__ void element_position(int64_t e, double* out) const {
    for (int d = 0; d < 3; ++d) out[d] = 0.5 * (X(d, e + 1) + X(d, e));
  }
  // pyelastica _node_to_element_velocity: (m[e+1] v[e+1] + m[e] v[e]) / (m[e+1] + m[e])
  __device__ void element_velocity(int64_t e, double* out) const {
    const double m1 = m[e + 1], m0 = m[e];
    for (int d = 0; d < 3; ++d) out[d] = (m1 * V(d, e + 1) + m0 * V(d, e)) / (m1 + m0);
  }
  // Q^T omega, the element's angular velocity in the lab frame (:224-227, :453-456)
  __device__ void global_omega(int64_t e, double* out) const {
    for (int d = 0; d < 3; ++d) {
      double acc = 0.0;
      for (int j = 0; j < 3; ++j) acc += Q(j, d, e) * W(j, e);
      out[d] = acc;
    }
  }
  __device__ void to_local(int64_t e, const double* g, double* out) const {  // Q g
    for (int i = 0; i < 3; ++i) {
      double acc = 0.0;
      for (int j = 0; j < 3; ++j) acc += Q(i, j, e) * g[j];
      out[i] = acc;
    }
  }
};

__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}

struct SurfaceArgs {
  const int32_t* node_element;  // (N)   element of every surface node
  const double* local_points;   // (2, N) cos / sin of the node's angle in the element's d1-d2 plane (0, 0: centre)
  const double* radius_ratio;   // (N)   1 on the lateral surface, < 1 on the cap rings
};

__global__ void __launch_bounds__(128)
    rod_kinematics_kernel(int mode, int dim, const double* state, int64_t n, int64_t n_lag, double* pos, int64_t pos_s,
                          double* vel, int64_t vel_s, double* arm_out, int64_t arm_s, SurfaceArgs sa) {
  const RodView rod(state, n);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_lag;
       i += (int64_t)gridDim.x * blockDim.x) {
    double p[3], u[3];
    if (mode == ROD_NODAL) {
      for (int d = 0; d < 3; ++d) p[d] = rod.X(d, i), u[d] = rod.V(d, i);
    } else {
      int64_t e = i;
      double arm[3] = {0.0, 0.0, 0.0};
      bool has_arm = false;
      if (mode == ROD_EDGE) {
        e = i % n;
        const int side = (int)(i / n);  // 0 centre, 1 left (+ r d), 2 right (- r d)
        if (side) {
          // z x t, the in-plane normal (:192-195); the left edge owns the stored moment arm
          const double nx = -rod.T(1, e), ny = rod.T(0, e), rad = rod.r[e];
          const double ax = nx * rad, ay = ny * rad;
          if (side == 1) arm_out[e] = ax, arm_out[arm_s + e] = ay, arm_out[2 * arm_s + e] = 0.0 * rad;
          arm[0] = side == 1 ? ax : -ax;
          arm[1] = side == 1 ? ay : -ay;
          has_arm = true;
        }
      } else if (mode == ROD_SURFACE) {
        e = sa.node_element[i];
        const double rad = rod.r[e] * sa.radius_ratio[i];
        const double lx = sa.local_points[i], ly = sa.local_points[n_lag + i];
        for (int d = 0; d < 3; ++d) {
          arm[d] = rad * (rod.Q(0, d, e) * lx + rod.Q(1, d, e) * ly);  // r Q^T (lx, ly, 0), :428-431
          arm_out[d * arm_s + i] = arm[d];
        }
        has_arm = true;
      }
      rod.element_position(e, p);
      rod.element_velocity(e, u);
      if (has_arm) {
        double og[3], c[3];
        rod.global_omega(e, og);
        cross3(og, arm, c);
        for (int d = 0; d < 3; ++d) p[d] += arm[d], u[d] += c[d];
      }
    }
    for (int d = 0; d < dim; ++d) pos[d * pos_s + i] = p[d], vel[d * vel_s + i] = u[d];
  }
}

__device__ __forceinline__ double warp_sum(double v) {
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// Sum of the forcing over the surface nodes of element e (lanes stride over the element's window) and, if `tq`, of
// arm x (-f) over the same nodes (:488-489).
template <typename T>
__device__ void surface_element_sums(const T* f, int64_t f_s, const double* arm, int64_t arm_s, const int32_t* start,
                                     int64_t e, int lane, double* fs, double* tq) {
  double s[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t i = start[e] + lane; i < start[e + 1]; i += 32) {
    const double q[3] = {-(double)f[i], -(double)f[f_s + i], -(double)f[2 * f_s + i]};
    s[0] -= q[0], s[1] -= q[1], s[2] -= q[2];
    if (tq) {
      const double a[3] = {arm[i], arm[arm_s + i], arm[2 * arm_s + i]};
      double c[3];
      cross3(a, q, c);
      s[3] += c[0], s[4] += c[1], s[5] += c[2];
    }
  }
  for (int c = 0; c < 3; ++c) fs[c] = warp_sum(s[c]);
  if (tq)
    for (int c = 0; c < 3; ++c) tq[c] = warp_sum(s[3 + c]);
}

template <typename T>
__global__ void __launch_bounds__(128)
    rod_transfer_kernel(int mode, int dim, const double* state, int64_t n, const T* f, int64_t f_s, double* arm,
                        int64_t arm_s, const int32_t* start, double* forces, double* torques) {
  const RodView rod(state, n);
  const int lane = threadIdx.x & 31;
  const int64_t s = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);  // rod node
  if (s > n) return;
  double node_f[3] = {0.0, 0.0, 0.0}, tq_g[3] = {0.0, 0.0, 0.0};
  bool has_tq = false;
  auto F = [&](int d, int64_t k) -> double { return d < dim ? (double)f[d * f_s + k] : 0.0; };

  if (mode == ROD_NODAL) {
    // forces = -f (:42); torque_e = arm_e x (F[e+1] - F[e]) / 2 with the end corrections (:45-67)
    for (int d = 0; d < 3; ++d) node_f[d] = -F(d, s);
    if (s < n) {
      double a[3], df[3], fe[3], c[3];
      for (int d = 0; d < 3; ++d) {
        a[d] = (rod.X(d, s + 1) - rod.X(d, s)) / 2.0;
        df[d] = (-F(d, s + 1) - node_f[d]) / 2.0;
        if (lane == 0) arm[d * arm_s + s] = a[d];
      }
      cross3(a, df, tq_g);
      if (s == n - 1) {
        for (int d = 0; d < 3; ++d) fe[d] = -F(d, n);
        cross3(a, fe, c);
        for (int d = 0; d < 3; ++d) tq_g[d] += c[d] / 2.0;
      }
      if (s == 0) {
        cross3(a, node_f, c);
        for (int d = 0; d < 3; ++d) tq_g[d] -= c[d] / 2.0;
      }
      has_tq = true;
    }
  } else if (mode == ROD_ELEMENT) {
    // half of every element's forcing to each of its nodes (:118-120); torques are left alone (:122-123)
    for (int d = 0; d < 3; ++d) {
      if (s > 0) node_f[d] -= 0.5 * F(d, s - 1);
      if (s < n) node_f[d] -= 0.5 * F(d, s);
    }
  } else if (mode == ROD_EDGE) {
    // centre nodes (:249-254), then left + right edge forces through _elements_to_nodes_inplace (:274-278)
    for (int d = 0; d < 3; ++d) {
      if (s > 0) node_f[d] -= 0.5 * F(d, s - 1);
      if (s < n) node_f[d] -= 0.5 * F(d, s);
    }
    for (int d = 0; d < 3; ++d) {
      if (s > 0) node_f[d] += 0.5 * (-F(d, n + s - 1) + -F(d, 2 * n + s - 1));
      if (s < n) node_f[d] += 0.5 * (-F(d, n + s) + -F(d, 2 * n + s));
    }
    if (s < n) {
      // arm x (-f_left) + (-arm) x (-f_right) (:256-271)
      double a[3], na[3], fl[3], fr[3], c1[3], c2[3];
      for (int d = 0; d < 3; ++d)
        a[d] = arm[d * arm_s + s], na[d] = -a[d], fl[d] = -F(d, n + s), fr[d] = -F(d, 2 * n + s);
      cross3(a, fl, c1);
      cross3(na, fr, c2);
      for (int d = 0; d < 3; ++d) tq_g[d] = c1[d] + c2[d];
      has_tq = true;
    }
  } else {  // ROD_SURFACE (:476-497)
    double fs[3];
    if (s > 0) {
      surface_element_sums<T>(f, f_s, arm, arm_s, start, s - 1, lane, fs, nullptr);
      for (int d = 0; d < 3; ++d) node_f[d] -= 0.5 * fs[d];
    }
    if (s < n) {
      surface_element_sums<T>(f, f_s, arm, arm_s, start, s, lane, fs, tq_g);
      for (int d = 0; d < 3; ++d) node_f[d] -= 0.5 * fs[d];
      has_tq = true;
    }
  }
  if (lane == 0) {
    for (int d = 0; d < 3; ++d) forces[d * (n + 1) + s] = node_f[d];
    if (has_tq) {
      double tl[3];
      rod.to_local(s, tq_g, tl);  // lab frame -> the element's material frame (:70-73, :281-284, :494)
      for (int d = 0; d < 3; ++d) torques[d * n + s] = tl[d];
    }
  }
}

int check_lag_field(const char* fn, const sopht_field_t* f, int rows, int64_t n) {
  if (!f || !f->data || f->ndim != 2 || f->shape[0] != rows || f->shape[1] != n || (n > 1 && f->stride[1] != 1))
    SOPHT_FAIL(SOPHT_ERR_SHAPE, "%s: expected a (%d, %lld) array with contiguous rows", fn, rows, (long long)n);
  return SOPHT_OK;
}

int64_t lag_nodes_of(int mode, int64_t n) {
  return mode == ROD_NODAL ? n + 1 : mode == ROD_ELEMENT ? n : mode == ROD_EDGE ? 3 * n : -1;
}

int check_mode(const char* fn, int mode, int dim) {
  if (mode < ROD_NODAL || mode > ROD_SURFACE) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: unknown rod grid kind %d", fn, mode);
  if (dim != 2 && dim != 3) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: dim must be 2 or 3", fn);
  if (mode == ROD_EDGE && dim != 2) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: the edge grid is defined for dim 2 only", fn);
  if (mode == ROD_SURFACE && dim != 3)
    SOPHT_FAIL(SOPHT_ERR_ARG, "%s: the surface grid is defined for dim 3 only", fn);
  return SOPHT_OK;
}

}  // namespace
}  // namespace sopht

using namespace sopht;

extern "C" {

int64_t sopht_rod_state_doubles(int64_t n_elems) { return n_elems < 1 ? 0 : 7 * (n_elems + 1) + 16 * n_elems; }

int sopht_rod_forcing_grid_kinematics(int grid_kind, int dim, int64_t n_elems, const void* rod_state,
                                      const sopht_field_t* position_field, const sopht_field_t* velocity_field,
                                      const sopht_field_t* moment_arm, const void* surface_node_element,
                                      const void* surface_local_points, const void* surface_radius_ratio,
                                      void* stream) {
  int rc;
  if ((rc = check_mode(__func__, grid_kind, dim))) return rc;
  if (n_elems < 1 || !rod_state || !position_field) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null / empty rod", __func__);
  int64_t n_lag = lag_nodes_of(grid_kind, n_elems);
  if (grid_kind == ROD_SURFACE) {
    if (!surface_node_element || !surface_local_points || !surface_radius_ratio)
      SOPHT_FAIL(SOPHT_ERR_ARG, "%s: the surface grid needs its node tables", __func__);
    n_lag = position_field->ndim == 2 ? position_field->shape[1] : -1;
  }
  if ((rc = check_lag_field(__func__, position_field, dim, n_lag))) return rc;
  if ((rc = check_lag_field(__func__, velocity_field, dim, n_lag))) return rc;
  double* arm = nullptr;
  int64_t arm_s = 0;
  if (grid_kind == ROD_EDGE || grid_kind == ROD_SURFACE) {
    if ((rc = check_lag_field(__func__, moment_arm, 3, grid_kind == ROD_EDGE ? n_elems : n_lag))) return rc;
    arm = reinterpret_cast<double*>(moment_arm->data);
    arm_s = moment_arm->stride[0];
  }
  if (n_lag == 0) return SOPHT_OK;
  int blocks = (int)((n_lag + 127) / 128);
  if (blocks > 148 * 8) blocks = 148 * 8;
  cudaStream_t st = as_stream(stream);
  SOPHT_PROF("ib.rod_grid_kinematics", st);
  SurfaceArgs sa{reinterpret_cast<const int32_t*>(surface_node_element),
                 reinterpret_cast<const double*>(surface_local_points),
                 reinterpret_cast<const double*>(surface_radius_ratio)};
  rod_kinematics_kernel<<<blocks, 128, 0, st>>>(
      grid_kind, dim, reinterpret_cast<const double*>(rod_state), n_elems, n_lag,
      reinterpret_cast<double*>(position_field->data), position_field->stride[0],
      reinterpret_cast<double*>(velocity_field->data), velocity_field->stride[0], arm, arm_s, sa);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

int sopht_rod_forcing_grid_transfer(int forcing_dtype, int grid_kind, int dim, int64_t n_elems, const void* rod_state,
                                    const sopht_field_t* lag_grid_forcing_field, const sopht_field_t* moment_arm,
                                    const void* surface_element_start, void* forces_torques_out, void* stream) {
  SOPHT_CHECK_DTYPE(forcing_dtype);
  int rc;
  if ((rc = check_mode(__func__, grid_kind, dim))) return rc;
  if (n_elems < 1 || !rod_state || !lag_grid_forcing_field || !forces_torques_out)
    SOPHT_FAIL(SOPHT_ERR_ARG, "%s: null / empty rod", __func__);
  int64_t n_lag = lag_nodes_of(grid_kind, n_elems);
  if (grid_kind == ROD_SURFACE) {
    if (!surface_element_start) SOPHT_FAIL(SOPHT_ERR_ARG, "%s: the surface grid needs its element windows", __func__);
    n_lag = lag_grid_forcing_field->ndim == 2 ? lag_grid_forcing_field->shape[1] : -1;
  }
  if ((rc = check_lag_field(__func__, lag_grid_forcing_field, dim, n_lag))) return rc;
  double* arm = nullptr;
  int64_t arm_s = 0;
  if (grid_kind != ROD_ELEMENT) {
    if ((rc = check_lag_field(__func__, moment_arm, 3, grid_kind == ROD_SURFACE ? n_lag : n_elems))) return rc;
    arm = reinterpret_cast<double*>(moment_arm->data);
    arm_s = moment_arm->stride[0];
  }
  cudaStream_t st = as_stream(stream);
  SOPHT_PROF("ib.rod_grid_transfer", st);
  auto* forces = reinterpret_cast<double*>(forces_torques_out);
  double* torques = forces + 3 * (n_elems + 1);
  SOPHT_CUDA(cudaMemsetAsync(forces, 0, (size_t)(6 * n_elems + 3) * sizeof(double), st));
  const int warps = 4;
  const int blocks = (int)((n_elems + 1 + warps - 1) / warps);
  const auto* state = reinterpret_cast<const double*>(rod_state);
  const auto* start = reinterpret_cast<const int32_t*>(surface_element_start);
  if (forcing_dtype == SOPHT_F32)
    rod_transfer_kernel<float><<<blocks, 32 * warps, 0, st>>>(
        grid_kind, dim, state, n_elems, reinterpret_cast<const float*>(lag_grid_forcing_field->data),
        lag_grid_forcing_field->stride[0], arm, arm_s, start, forces, torques);
  else
    rod_transfer_kernel<double><<<blocks, 32 * warps, 0, st>>>(
        grid_kind, dim, state, n_elems, reinterpret_cast<const double*>(lag_grid_forcing_field->data),
        lag_grid_forcing_field->stride[0], arm, arm_s, start, forces, torques);
  SOPHT_CHECK_LAUNCH();
  return SOPHT_OK;
}

}  // extern "C"
