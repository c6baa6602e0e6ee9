// cm_debug.cu -- device-side self-test hook: runs the shared small-matrix routines of cm_math.h ON THE GPU so a test
// can compare them bit for bit with the same header compiled for the host inside the oracle (tests/test_math_gpu.py).  This guards
// the assumption every parity claim rests on: cm_math.h means the same thing on both sides (it once did not --
// see the note in colpiv_qr_solve).
#include "cm_host.h"
#include "cm_math.h"
#include "cm_device.cuh"

namespace cm {

// op 0: colpiv_qr_solve<6,6> (36 + 6 -> 6)   1: colpiv_qr_solve<5,3> (15 + 5 -> 3)   2: eig3_sym (6 -> 3 + 9)
// op 3: eig_sym<6> (36 -> 6 + 36)   4: eig_sym<6> values only (36 -> 6)   5: inverse_lu<6> (36 -> 36)
// op 6: pose_to_matrix + cm_sincosf (6 -> 9 + 3 + 3)   7: cm_atan2f(y, x), cm_atanf(y / x) (2 -> 2)
// op 8: warp_qr_solve6 (cm_device.cuh), the column-parallel 6x6 solve of the Gauss-Newton step, one warp per problem (36 + 6 -> 6): must
//       equal op 0 on the host bit for bit, rank-deficient systems included
__device__ void debug_math_op(int op, const float* in, float* out) {
  if (op == 0) { float A[36], b[6], x[6]; for (int i = 0; i < 36; i++) A[i] = in[i]; for (int i = 0; i < 6; i++) b[i] = in[36 + i]; colpiv_qr_solve<6, 6>(A, b, x); for (int i = 0; i < 6; i++) out[i] = x[i]; }
  else if (op == 1) { float A[15], b[5], x[3]; for (int i = 0; i < 15; i++) A[i] = in[i]; for (int i = 0; i < 5; i++) b[i] = in[15 + i]; colpiv_qr_solve<5, 3>(A, b, x); for (int i = 0; i < 3; i++) out[i] = x[i]; }
  else if (op == 2) { float w[3], V[9]; eig3_sym(in, w, V); for (int i = 0; i < 3; i++) out[i] = w[i]; for (int i = 0; i < 9; i++) out[3 + i] = V[i]; }
  else if (op == 3) { float w[6], V[36]; eig_sym<6>(in, w, V); for (int i = 0; i < 6; i++) out[i] = w[i]; for (int i = 0; i < 36; i++) out[6 + i] = V[i]; }
  else if (op == 4) { float w[6]; eig_sym<6>(in, w, (float*)nullptr); for (int i = 0; i < 6; i++) out[i] = w[i]; }
  else if (op == 5) { float inv[36]; bool ok = inverse_lu<6>(in, inv); for (int i = 0; i < 36; i++) out[i] = ok ? inv[i] : 0.f; }
  else if (op == 7) { out[0] = cm_atan2f(in[0], in[1]); out[1] = cm_atanf(in[0] / in[1]); }
  else if (op == 9) { out[0] = dev_min_eig_above6(in, in[36]) ? 1.f : 0.f; }   // the eigen-solve skip of evaluation 0: 36 + threshold -> 0 / 1
  else if (op == 6) { float R[9]; pose_to_matrix(in, R); for (int i = 0; i < 9; i++) out[i] = R[i]; for (int i = 0; i < 3; i++) cm_sincosf(in[i], &out[9 + i], &out[12 + i]); }
}
__global__ void debug_math_kernel(int op, const float* in, int nin, float* out, int nout, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) debug_math_op(op, in + (size_t)i * nin, out + (size_t)i * nout);
}
__global__ void __launch_bounds__(128) debug_warp_qr_kernel(const float* in, float* out, int n) {
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (i >= n) return;
  const float* p = in + (size_t)i * 42;
  float col[6], X[6];
#pragma unroll
  for (int r = 0; r < 6; r++) col[r] = lane < 6 ? p[r * 6 + lane] : (lane == 6 ? p[36 + r] : 0.f);
  warp_qr_solve6(col, lane, X);
  if (lane == 0) for (int k = 0; k < 6; k++) out[(size_t)i * 6 + k] = X[k];
}
int debug_math_dims(int op, int* nin, int* nout) {
  static const int ni[10] = {42, 20, 6, 36, 36, 36, 6, 2, 42, 37}, no[10] = {6, 3, 12, 42, 6, 36, 15, 2, 6, 1};
  if (op < 0 || op > 9) return -1;
  *nin = ni[op]; *nout = no[op];
  return 0;
}
void launch_debug_math(int op, const float* d_in, int nin, float* d_out, int nout, int n, cudaStream_t stream) {
  if (op == 8) { CM_LAUNCH(debug_warp_qr_kernel, (n * 32 + 127) / 128, 128, 0, stream, d_in, d_out, n); return; }
  CM_LAUNCH(debug_math_kernel, (n + 63) / 64, 64, 0, stream, op, d_in, nin, d_out, nout, n);
}

}  // namespace cm
