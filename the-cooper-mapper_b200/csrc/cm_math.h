// cm_math.h -- small fixed-size float32 linear algebra shared by the CUDA kernels and the CPU oracle.
//
// The reference (L_SLAM) calls Eigen3 for every small dense solve on the hot path; Eigen is NOT under
// /root/reference and is not installed here, so this header restates the published algorithms of the
// Eigen 3.3 routines the reference calls:
//   * SelfAdjointEigenSolver<Matrix3f>           ScanRegistration.cpp:582,628  feature_utils.h:141
//   * SelfAdjointEigenSolver<Matrix<float,6,6>>  ScanMatch.cpp:216
//   * colPivHouseholderQr().solve()              feature_utils.h:182 (5x3)     ScanMatch.cpp:209 (6x6)
//   * Matrix<float,6,6>::inverse()               ScanMatch.cpp:234  (partial-pivot LU for n > 4)
//   * Quaternionf(AngleAxisf) products + toRotationMatrix()   transform_utils.h:288-299
// Every routine uses only + - * / sqrt on float (IEEE-754, no contraction: the CUDA side is compiled with
// -fmad=false, the oracle with -ffp-contract=off), so the GPU and the oracle produce bit-identical results
// for identical inputs.  Agreement with a real Eigen binary is to float tolerance only ("parity unpinned",
// see DESIGN.md).  sin/cos of pose angles use cm_sincosf below on BOTH sides for the same reason (the
// reference calls libm float sin/cos, Angle.h:19-20; cm_sincosf agrees with glibc sinf/cosf to <= 1 ulp).
#pragma once
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define CM_HD __host__ __device__ __forceinline__
#define CM_UNROLL _Pragma("unroll")
#else
#define CM_HD inline
#define CM_UNROLL
#endif

namespace cm {

// ----------------------------------------------------------------------------------------------------------
// sin/cos: Cody-Waite reduction to [-pi/4, pi/4] and Taylor kernels in double, rounded once to float.
// ----------------------------------------------------------------------------------------------------------
CM_HD void cm_sincos_d(double x, double* s, double* c) {
  const double two_over_pi = 6.36619772367581382433e-01;
  const double pio2_hi = 1.57079632673412561417e+00;  // first 33 bits of pi/2
  const double pio2_lo = 6.07710050650619224932e-11;  // pi/2 - pio2_hi
  double kd = rint(x * two_over_pi);
  double r = (x - kd * pio2_hi) - kd * pio2_lo;
  int q = (int)((long long)kd & 3LL);
  double r2 = r * r;
  // Taylor coefficients 1/n! (|r| <= pi/4: truncation error < 5e-17)
  double ps = -1.0 / 1307674368000.0;              // -1/15!
  ps = ps * r2 + 1.0 / 6227020800.0;               // +1/13!
  ps = ps * r2 - 1.0 / 39916800.0;                 // -1/11!
  ps = ps * r2 + 1.0 / 362880.0;                   // +1/9!
  ps = ps * r2 - 1.0 / 5040.0;                     // -1/7!
  ps = ps * r2 + 1.0 / 120.0;                      // +1/5!
  ps = ps * r2 - 1.0 / 6.0;                        // -1/3!
  double sr = r + r * (r2 * ps);
  double pc = 1.0 / 20922789888000.0;              // +1/16!
  pc = pc * r2 - 1.0 / 87178291200.0;              // -1/14!
  pc = pc * r2 + 1.0 / 479001600.0;                // +1/12!
  pc = pc * r2 - 1.0 / 3628800.0;                  // -1/10!
  pc = pc * r2 + 1.0 / 40320.0;                    // +1/8!
  pc = pc * r2 - 1.0 / 720.0;                      // -1/6!
  pc = pc * r2 + 1.0 / 24.0;                       // +1/4!
  pc = pc * r2 - 0.5;                              // -1/2!
  double cr = 1.0 + r2 * pc;
  double ss, cc;
  if (q == 0) { ss = sr; cc = cr; }
  else if (q == 1) { ss = cr; cc = -sr; }
  else if (q == 2) { ss = -sr; cc = -cr; }
  else { ss = -cr; cc = sr; }
  *s = ss; *c = cc;
}
CM_HD void cm_sincosf(float x, float* s, float* c) {
#if defined(CM_LIBM_TRIG) && !defined(__CUDACC__)
  // oracle variant liboracle_libm.so only (tests/test_parity_lapack_cpu.py): the reference's own calls, libm float sin / cos -- used to
  // bound what the canonical definition below changes at the pose level
  *s = sinf(x); *c = cosf(x);
#else
  double sd, cd;
  cm_sincos_d((double)x, &sd, &cd);
  *s = (float)sd; *c = (float)cd;
#endif
}

// ----------------------------------------------------------------------------------------------------------
// atan / atan2 for the raw-sweep front end (MultiScanRegistration.cpp:103-156 calls libm's float atan / atan2): like
// cm_sincosf a canonical definition shared by the GPU and the oracle -- evaluated in double with + - * / only
// (x > 1: pi/2 - atan(1/x); t in [0, 1]: atan(t) = atan(k/8) + atan((t - k/8) / (1 + t k/8)), |argument| <= 1/16, Taylor series to
// u^17) and rounded once to float: error < 1e-16 before the rounding, i.e. the correctly rounded float except within ~1e-8 ulp
// of a rounding boundary; glibc's atanf / atan2f differ from it by 1 ulp on a few percent of the inputs.
// ----------------------------------------------------------------------------------------------------------
CM_HD double cm_atan_tab(int k) {   // atan(k / 8), correctly rounded doubles
  switch (k) {
    case 0: return 0.0;
    case 1: return 1.24354994546761438e-01;
    case 2: return 2.44978663126864143e-01;
    case 3: return 3.58770670270572245e-01;
    case 4: return 4.63647609000806094e-01;
    case 5: return 5.58599315343562441e-01;
    case 6: return 6.43501108793284371e-01;
    case 7: return 7.18829999621624527e-01;
    default: return 7.85398163397448279e-01;
  }
}
CM_HD double cm_atan_pos_d(double x) {   // x >= 0 (or NaN) -> atan(x) in [0, pi/2]
  if (!(x == x)) return x;
  const double pio2_hi = 1.57079632679489656e+00, pio2_lo = 6.12323399573676604e-17;
  const bool inv = x > 1.0;
  const double t = inv ? 1.0 / x : x;           // [0, 1]; 1 / inf = 0
  const int k = (int)(t * 8.0 + 0.5);
  const double c = (double)k * 0.125;
  const double u = (t - c) / (1.0 + t * c);
  const double u2 = u * u;
  double p = 1.0 / 17.0;
  p = p * u2 - 1.0 / 15.0;
  p = p * u2 + 1.0 / 13.0;
  p = p * u2 - 1.0 / 11.0;
  p = p * u2 + 1.0 / 9.0;
  p = p * u2 - 1.0 / 7.0;
  p = p * u2 + 1.0 / 5.0;
  p = p * u2 - 1.0 / 3.0;
  const double a = cm_atan_tab(k) + (u + u * (u2 * p));
  return inv ? (pio2_hi - a) + pio2_lo : a;
}
CM_HD float cm_atanf(float x) {
#if defined(CM_LIBM_TRIG) && !defined(__CUDACC__)
  return atanf(x);   // (oracle variant liboracle_libm.so only, see cm_sincosf)
#endif
  const double a = cm_atan_pos_d(fabs((double)x));
  return (float)copysign(a, (double)x);
}
CM_HD float cm_atan2f(float y, float x) {   // finite arguments or NaN (the front end drops non-finite points before it gets here)
#if defined(CM_LIBM_TRIG) && !defined(__CUDACC__)
  return atan2f(y, x);
#endif
  const double pi_hi = 3.14159265358979312e+00, pi_lo = 1.22464679914735321e-16;
  const double pio2_hi = 1.57079632679489656e+00, pio2_lo = 6.12323399573676604e-17;
  const double ax = fabs((double)x), ay = fabs((double)y);
  double a;
  if (ax == 0.0 && ay == 0.0) a = 0.0;
  else if (ay <= ax) a = cm_atan_pos_d(ay / ax);
  else a = (pio2_hi - cm_atan_pos_d(ax / ay)) + pio2_lo;
  if (x < 0.f || (x == 0.f && signbit(x) && ay == 0.0)) a = (pi_hi - a) + pi_lo;   // second / third quadrant, atan2(+-0, -0) = +-pi
  return (float)((y < 0.f || (y == 0.f && signbit(y))) ? -a : a);
}

// ----------------------------------------------------------------------------------------------------------
// Givens rotation, Eigen JacobiRotation<float>::makeGivens (real case).
// ----------------------------------------------------------------------------------------------------------
CM_HD void make_givens(float p, float q, float* c, float* s) {
  if (q == 0.f) { *c = p < 0.f ? -1.f : 1.f; *s = 0.f; }
  else if (p == 0.f) { *c = 0.f; *s = q < 0.f ? 1.f : -1.f; }
  else if (fabsf(p) > fabsf(q)) {
    float t = q / p;
    float u = sqrtf(1.f + t * t);
    if (p < 0.f) u = -u;
    *c = 1.f / u;
    *s = -t * (*c);
  } else {
    float t = p / q;
    float u = sqrtf(1.f + t * t);
    if (q < 0.f) u = -u;
    *s = -1.f / u;
    *c = -t * (*s);
  }
}

// Eigen numext::hypot (3.3): p * sqrt(1 + (q/p)^2) with p = max(|x|,|y|)
CM_HD float cm_hypot(float x, float y) {
  float ax = fabsf(x), ay = fabsf(y);
  float p = ax > ay ? ax : ay;
  if (p == 0.f) return 0.f;
  float qp = (ax > ay ? ay : ax) / p;
  return p * sqrtf(1.f + qp * qp);
}

// One implicit symmetric QR step with Wilkinson shift on the tridiagonal (diag, subdiag) block [start,end];
// Q (n x n, element (r,c) at Q[r*N + c]) accumulates the rotations on the right.
// Eigen internal::tridiagonal_qr_step.
template <int N>
CM_HD void tridiagonal_qr_step(float* diag, float* subdiag, int start, int end, float* Q) {
  float td = (diag[end - 1] - diag[end]) * 0.5f;
  float e = subdiag[end - 1];
  float mu = diag[end];
  if (td == 0.f) {
    mu -= fabsf(e);
  } else if (e != 0.f) {
    float e2 = e * e;
    float h = cm_hypot(td, e);
    if (e2 == 0.f) mu -= e / ((td + (td > 0.f ? h : -h)) / e);
    else mu -= e2 / (td + (td > 0.f ? h : -h));
  }
  float x = diag[start] - mu;
  float z = subdiag[start];
  for (int k = start; k < end && z != 0.f; ++k) {
    float c, s;
    make_givens(x, z, &c, &s);
    float sdk = s * diag[k] + c * subdiag[k];
    float dkp1 = s * subdiag[k] + c * diag[k + 1];
    diag[k] = c * (c * diag[k] - s * subdiag[k]) - s * (c * subdiag[k] - s * diag[k + 1]);
    diag[k + 1] = s * sdk + c * dkp1;
    subdiag[k] = c * sdk - s * dkp1;
    if (k > start) subdiag[k - 1] = c * subdiag[k - 1] - s * z;
    x = subdiag[k];
    if (k < end - 1) {
      z = -s * subdiag[k + 1];
      subdiag[k + 1] = c * subdiag[k + 1];
    }
    // Q = Q * G : columns k and k+1 (skipped when only eigenvalues are wanted; diag/subdiag do not depend on Q)
    if (Q) {
CM_UNROLL
      for (int r = 0; r < N; ++r) {
        float xi = Q[r * N + k], yi = Q[r * N + k + 1];
        Q[r * N + k] = c * xi - s * yi;
        Q[r * N + k + 1] = s * xi + c * yi;
      }
    }
  }
}

// Iterate QR steps until the tridiagonal matrix is diagonal, then sort ascending (selection sort with
// column swaps).  Eigen internal::computeFromTridiagonal_impl (3.3: precision = 2*epsilon, max 30*n steps).
template <int N>
CM_HD void tridiagonal_to_eigen(float* diag, float* subdiag, float* Q) {
  const float considerAsZero = FLT_MIN;
  const float precision = 2.f * FLT_EPSILON;
  int end = N - 1, start = 0, iter = 0;
  while (end > 0) {
    for (int i = start; i < end; ++i) {
      if (fabsf(subdiag[i]) <= (fabsf(diag[i]) + fabsf(diag[i + 1])) * precision ||
          fabsf(subdiag[i]) <= considerAsZero)
        subdiag[i] = 0.f;
    }
    while (end > 0 && subdiag[end - 1] == 0.f) end--;
    if (end <= 0) break;
    iter++;
    if (iter > 30 * N) break;
    start = end - 1;
    while (start > 0 && subdiag[start - 1] != 0.f) start--;
    tridiagonal_qr_step<N>(diag, subdiag, start, end, Q);
  }
  for (int i = 0; i < N - 1; ++i) {
    int k = 0;
    float m = diag[i];
    for (int j = 1; j < N - i; ++j)
      if (diag[i + j] < m) { m = diag[i + j]; k = j; }
    if (k > 0) {
      float t = diag[i]; diag[i] = diag[k + i]; diag[k + i] = t;
      if (Q) {
        for (int r = 0; r < N; ++r) {
          float u = Q[r * N + i]; Q[r * N + i] = Q[r * N + k + i]; Q[r * N + k + i] = u;
        }
      }
    }
  }
}

// N = 3 specialisation of tridiagonal_qr_step / tridiagonal_to_eigen with compile-time indices only (identical
// arithmetic, operation for operation): diag / subdiag / Q stay in registers on the GPU instead of local memory.
// This is the hot eigen-solve (one per corner query and two per classified scan point).
template <int START, int END>
CM_HD void tridiag3_qr_sweep(float* diag, float* subdiag, float* Q) {
  float td = (diag[END - 1] - diag[END]) * 0.5f;
  float e = subdiag[END - 1];
  float mu = diag[END];
  if (td == 0.f) {
    mu -= fabsf(e);
  } else if (e != 0.f) {
    float e2 = e * e;
    float h = cm_hypot(td, e);
    if (e2 == 0.f) mu -= e / ((td + (td > 0.f ? h : -h)) / e);
    else mu -= e2 / (td + (td > 0.f ? h : -h));
  }
  float x = diag[START] - mu;
  float z = subdiag[START];
  bool alive = true;
CM_UNROLL
  for (int k = START; k < END; ++k) {
    alive = alive && (z != 0.f);
    if (alive) {
      float c, s;
      make_givens(x, z, &c, &s);
      float sdk = s * diag[k] + c * subdiag[k];
      float dkp1 = s * subdiag[k] + c * diag[k + 1];
      diag[k] = c * (c * diag[k] - s * subdiag[k]) - s * (c * subdiag[k] - s * diag[k + 1]);
      diag[k + 1] = s * sdk + c * dkp1;
      subdiag[k] = c * sdk - s * dkp1;
      if (k > START) subdiag[k - 1] = c * subdiag[k - 1] - s * z;
      x = subdiag[k];
      if (k < END - 1) {
        z = -s * subdiag[k + 1];
        subdiag[k + 1] = c * subdiag[k + 1];
      }
CM_UNROLL
      for (int r = 0; r < 3; ++r) {
        float xi = Q[r * 3 + k], yi = Q[r * 3 + k + 1];
        Q[r * 3 + k] = c * xi - s * yi;
        Q[r * 3 + k + 1] = s * xi + c * yi;
      }
    }
  }
}

CM_HD bool tridiag3_negligible(float sub, float da, float db) {
  return fabsf(sub) <= (fabsf(da) + fabsf(db)) * (2.f * FLT_EPSILON) || fabsf(sub) <= FLT_MIN;
}

CM_HD void tridiagonal_to_eigen3(float* diag, float* subdiag, float* Q) {
  int end = 2, start = 0, iter = 0;
  while (end > 0) {
    if (start <= 0 && 0 < end) { if (tridiag3_negligible(subdiag[0], diag[0], diag[1])) subdiag[0] = 0.f; }
    if (start <= 1 && 1 < end) { if (tridiag3_negligible(subdiag[1], diag[1], diag[2])) subdiag[1] = 0.f; }
    if (end == 2 && subdiag[1] == 0.f) end = 1;
    if (end == 1 && subdiag[0] == 0.f) end = 0;
    if (end <= 0) break;
    iter++;
    if (iter > 30 * 3) break;
    if (end == 2) {
      if (subdiag[0] != 0.f) { start = 0; tridiag3_qr_sweep<0, 2>(diag, subdiag, Q); }
      else { start = 1; tridiag3_qr_sweep<1, 2>(diag, subdiag, Q); }
    } else {
      start = 0; tridiag3_qr_sweep<0, 1>(diag, subdiag, Q);
    }
  }
  // ascending selection sort with column swaps (i = 0, then i = 1)
  {
    int k = 0; float m = diag[0];
    if (diag[1] < m) { m = diag[1]; k = 1; }
    if (diag[2] < m) { m = diag[2]; k = 2; }
    if (k == 1) { float t = diag[0]; diag[0] = diag[1]; diag[1] = t;
CM_UNROLL
      for (int r = 0; r < 3; ++r) { float u = Q[r * 3]; Q[r * 3] = Q[r * 3 + 1]; Q[r * 3 + 1] = u; } }
    else if (k == 2) { float t = diag[0]; diag[0] = diag[2]; diag[2] = t;
CM_UNROLL
      for (int r = 0; r < 3; ++r) { float u = Q[r * 3]; Q[r * 3] = Q[r * 3 + 2]; Q[r * 3 + 2] = u; } }
  }
  if (diag[2] < diag[1]) {
    float t = diag[1]; diag[1] = diag[2]; diag[2] = t;
CM_UNROLL
    for (int r = 0; r < 3; ++r) { float u = Q[r * 3 + 1]; Q[r * 3 + 1] = Q[r * 3 + 2]; Q[r * 3 + 2] = u; }
  }
}

// Symmetric 3x3 eigen-decomposition.  Input: lower triangle {a00,a10,a20,a11,a21,a22}; output: eigenvalues
// ascending in w, eigenvectors as COLUMNS of V (V[r*3+c]).
// Eigen SelfAdjointEigenSolver::compute + tridiagonalization_inplace_selector<MatrixType,3,false>.
CM_HD void eig3_sym(const float A[6], float w[3], float V[9]) {
  float m00 = A[0], m10 = A[1], m20 = A[2], m11 = A[3], m21 = A[4], m22 = A[5];
  float scale = fabsf(m00);
  if (fabsf(m10) > scale) scale = fabsf(m10);
  if (fabsf(m20) > scale) scale = fabsf(m20);
  if (fabsf(m11) > scale) scale = fabsf(m11);
  if (fabsf(m21) > scale) scale = fabsf(m21);
  if (fabsf(m22) > scale) scale = fabsf(m22);
  if (scale == 0.f) scale = 1.f;
  m00 /= scale; m10 /= scale; m20 /= scale; m11 /= scale; m21 /= scale; m22 /= scale;
  float diag[3], subdiag[2];
  diag[0] = m00;
  float v1norm2 = m20 * m20;
  if (v1norm2 <= FLT_MIN) {
    diag[1] = m11; diag[2] = m22; subdiag[0] = m10; subdiag[1] = m21;
    V[0] = 1.f; V[1] = 0.f; V[2] = 0.f; V[3] = 0.f; V[4] = 1.f; V[5] = 0.f; V[6] = 0.f; V[7] = 0.f; V[8] = 1.f;
  } else {
    float beta = sqrtf(m10 * m10 + v1norm2);
    float invBeta = 1.f / beta;
    float m01 = m10 * invBeta;
    float m02 = m20 * invBeta;
    float q = 2.f * m01 * m21 + m02 * (m22 - m11);
    diag[1] = m11 + m02 * q;
    diag[2] = m22 - m02 * q;
    subdiag[0] = beta;
    subdiag[1] = m21 - m01 * q;
    V[0] = 1.f; V[1] = 0.f; V[2] = 0.f; V[3] = 0.f; V[4] = m01; V[5] = m02; V[6] = 0.f; V[7] = m02; V[8] = -m01;
  }
  tridiagonal_to_eigen3(diag, subdiag, V);
  w[0] = diag[0] * scale; w[1] = diag[1] * scale; w[2] = diag[2] * scale;
}

// Householder reflector of x[0..n) (stride 1): on return x[1..n) holds the essential part, *tau and *beta
// as in Eigen MatrixBase::makeHouseholder.
CM_HD void make_householder(float* x, int n, float* tau, float* beta) {
  float tailSqNorm = 0.f;
  for (int i = 1; i < n; ++i) tailSqNorm += x[i] * x[i];
  float c0 = x[0];
  if (n == 1 || tailSqNorm <= FLT_MIN) {
    *tau = 0.f; *beta = c0;
    for (int i = 1; i < n; ++i) x[i] = 0.f;
  } else {
    float b = sqrtf(c0 * c0 + tailSqNorm);
    if (c0 >= 0.f) b = -b;
    float d = c0 - b;
    for (int i = 1; i < n; ++i) x[i] = x[i] / d;
    *tau = (b - c0) / b;
    *beta = b;
  }
}

// Symmetric N x N eigen-decomposition (general N; used for the 6x6 degeneracy test, ScanMatch.cpp:216).
// A is full row-major, only the lower triangle is read.  Eigenvalues ascending, eigenvectors = columns of V
// (V may be nullptr: eigenvalues only, bit-identical to the values of a full solve).
// Eigen internal::tridiagonalization_inplace (Householder) + HouseholderSequence::evalTo + QR iterations.
template <int N>
CM_HD void eig_sym(const float* Ain, float* w, float* V) {
  float A[N * N];
  float scale = 0.f;
  for (int r = 0; r < N; ++r)
    for (int c = 0; c <= r; ++c) { float a = fabsf(Ain[r * N + c]); if (a > scale) scale = a; }
  if (scale == 0.f) scale = 1.f;
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) A[r * N + c] = (c <= r) ? Ain[r * N + c] / scale : 0.f;
  float hc[N];  // Householder coefficients
  float col[N], p[N];   // function scope on purpose (see colpiv_qr_solve)
  for (int i = 0; i < N - 1; ++i) {
    int rs = N - i - 1;
    for (int j = 0; j < rs; ++j) col[j] = A[(i + 1 + j) * N + i];
    float h, beta;
    make_householder(col, rs, &h, &beta);
    col[0] = 1.f;
    // p = h * (A22 * u)  using the lower triangle of the trailing block as a self-adjoint view
    for (int r = 0; r < rs; ++r) {
      float acc = 0.f;
      for (int c = 0; c < rs; ++c) {
        float a = (c <= r) ? A[(i + 1 + r) * N + (i + 1 + c)] : A[(i + 1 + c) * N + (i + 1 + r)];
        acc += a * (h * col[c]);
      }
      p[r] = acc;
    }
    float dot = 0.f;
    for (int r = 0; r < rs; ++r) dot += p[r] * col[r];
    float alpha = h * -0.5f * dot;
    for (int r = 0; r < rs; ++r) p[r] += alpha * col[r];
    // rank-2 update of the lower triangle: A22 -= u p^T + p u^T
    for (int r = 0; r < rs; ++r)
      for (int c = 0; c <= r; ++c)
        A[(i + 1 + r) * N + (i + 1 + c)] -= (col[r] * p[c] + p[r] * col[c]);
    A[(i + 1) * N + i] = beta;
    for (int j = 1; j < rs; ++j) A[(i + 1 + j) * N + i] = col[j];
    hc[i] = h;
  }
  float diag[N], subdiag[N - 1];
  for (int i = 0; i < N; ++i) diag[i] = A[i * N + i];
  for (int i = 0; i < N - 1; ++i) subdiag[i] = A[(i + 1) * N + i];
  // Q = H_0 H_1 ... H_{N-2}, H_k = I - hc[k] v_k v_k^T, v_k = e_{k+1} + essential below
  if (V) {
  for (int r = 0; r < N; ++r)
    for (int c = 0; c < N; ++c) V[r * N + c] = (r == c) ? 1.f : 0.f;
  for (int k = N - 2; k >= 0; --k) {
    int r0 = k + 1;  // first row touched
    // apply H_k on the left to V[r0.., r0..]
    for (int c = r0; c < N; ++c) {
      float tmp = V[r0 * N + c];
      for (int r = r0 + 1; r < N; ++r) tmp += A[r * N + k] * V[r * N + c];
      V[r0 * N + c] -= hc[k] * tmp;
      for (int r = r0 + 1; r < N; ++r) V[r * N + c] -= hc[k] * A[r * N + k] * tmp;
    }
  }
  }
  tridiagonal_to_eigen<N>(diag, subdiag, V);
  for (int i = 0; i < N; ++i) w[i] = diag[i] * scale;
}

// Least-squares / square solve A x = b through column-pivoted Householder QR.  A is M x N row-major
// (destroyed), b has M entries (destroyed), x has N entries.
// Eigen ColPivHouseholderQR::computeInPlace + _solve_impl (3.3, LAPACK-style norm downdating).
// Written with compile-time indices only (the data-dependent column pivot is applied through predicated swaps)
// so that on the GPU every array stays in registers.  This is also a deliberate work-around: a first version
// that indexed the scratch arrays with the run-time pivot was mis-compiled by nvcc 12.9 for sm_100a (stack
// slots of two live arrays overlapped; reproduced on a B200, clean under ASan/UBSan on the host).
template <int M, int N>
CM_HD void colpiv_qr_solve(float* A, float* b, float* x) {
  constexpr int size = M < N ? M : N;
  float nu[N], nd[N], hco[size];   // updated / direct column norms, Householder coefficients
  int perm[N];
  float maxnorm = 0.f;
CM_UNROLL
  for (int k = 0; k < N; ++k) {
    float s = 0.f;
CM_UNROLL
    for (int r = 0; r < M; ++r) s += A[r * N + k] * A[r * N + k];
    nd[k] = sqrtf(s);
    nu[k] = nd[k];
    if (nu[k] > maxnorm) maxnorm = nu[k];
    perm[k] = k;
  }
  float te = maxnorm * FLT_EPSILON;
  float threshold_helper = (te * te) / (float)M;
  float norm_downdate_threshold = sqrtf(FLT_EPSILON);
  int nonzero_pivots = size;
CM_UNROLL
  for (int k = 0; k < size; ++k) {
    int big = k;
    float bigv = nu[k];
CM_UNROLL
    for (int j = k + 1; j < N; ++j)
      if (nu[j] > bigv) { bigv = nu[j]; big = j; }
    float big_sq = bigv * bigv;
    if (nonzero_pivots == size && big_sq < threshold_helper * (float)(M - k)) nonzero_pivots = k;
CM_UNROLL
    for (int j = k + 1; j < N; ++j) {
      if (big == j) {
CM_UNROLL
        for (int r = 0; r < M; ++r) { float t = A[r * N + k]; A[r * N + k] = A[r * N + j]; A[r * N + j] = t; }
        float t = nu[k]; nu[k] = nu[j]; nu[j] = t;
        t = nd[k]; nd[k] = nd[j]; nd[j] = t;
        int ti = perm[k]; perm[k] = perm[j]; perm[j] = ti;
      }
    }
    // Householder reflector of column k, rows k..M-1 (MatrixBase::makeHouseholder); essential part stored below the diagonal
    float tailSqNorm = 0.f;
CM_UNROLL
    for (int r = k + 1; r < M; ++r) tailSqNorm += A[r * N + k] * A[r * N + k];
    float c0 = A[k * N + k];
    float tau, beta;
    if (k == M - 1 || tailSqNorm <= FLT_MIN) {
      tau = 0.f; beta = c0;
CM_UNROLL
      for (int r = k + 1; r < M; ++r) A[r * N + k] = 0.f;
    } else {
      float bb = sqrtf(c0 * c0 + tailSqNorm);
      if (c0 >= 0.f) bb = -bb;
      float d = c0 - bb;
CM_UNROLL
      for (int r = k + 1; r < M; ++r) A[r * N + k] = A[r * N + k] / d;
      tau = (bb - c0) / bb;
      beta = bb;
    }
    A[k * N + k] = beta;
    hco[k] = tau;
    // apply H_k to the trailing columns
CM_UNROLL
    for (int c = k + 1; c < N; ++c) {
      float tmp = A[k * N + c];
CM_UNROLL
      for (int r = k + 1; r < M; ++r) tmp += A[r * N + k] * A[r * N + c];
      A[k * N + c] -= tau * tmp;
CM_UNROLL
      for (int r = k + 1; r < M; ++r) A[r * N + c] -= tau * A[r * N + k] * tmp;
    }
CM_UNROLL
    for (int j = k + 1; j < N; ++j) {
      if (nu[j] != 0.f) {
        float temp = fabsf(A[k * N + j]) / nu[j];
        temp = (1.f + temp) * (1.f - temp);
        temp = temp < 0.f ? 0.f : temp;
        float ratio = nu[j] / nd[j];
        float temp2 = temp * (ratio * ratio);
        if (temp2 <= norm_downdate_threshold) {
          float s = 0.f;
CM_UNROLL
          for (int r = k + 1; r < M; ++r) s += A[r * N + j] * A[r * N + j];
          nd[j] = sqrtf(s);
          nu[j] = nd[j];
        } else {
          nu[j] *= sqrtf(temp);
        }
      }
    }
  }
CM_UNROLL
  for (int i = 0; i < N; ++i) x[i] = 0.f;
  if (nonzero_pivots == 0) return;
  // c = Q^T b = H_{p-1} ... H_0 b
CM_UNROLL
  for (int k = 0; k < size; ++k) {
    if (k < nonzero_pivots) {
      float tmp = b[k];
CM_UNROLL
      for (int r = k + 1; r < M; ++r) tmp += A[r * N + k] * b[r];
      b[k] -= hco[k] * tmp;
CM_UNROLL
      for (int r = k + 1; r < M; ++r) b[r] -= hco[k] * A[r * N + k] * tmp;
    }
  }
  // back substitution on the leading nonzero_pivots x nonzero_pivots upper triangle
CM_UNROLL
  for (int i = size - 1; i >= 0; --i) {
    if (i < nonzero_pivots) {
      float s = b[i];
CM_UNROLL
      for (int j = i + 1; j < size; ++j)
        if (j < nonzero_pivots) s -= A[i * N + j] * b[j];
      b[i] = s / A[i * N + i];
    }
  }
CM_UNROLL
  for (int i = 0; i < size; ++i) {
    if (i < nonzero_pivots) {
CM_UNROLL
      for (int c = 0; c < N; ++c)
        if (perm[i] == c) x[c] = b[i];
    }
  }
}

// General N x N inverse by partial-pivot LU (Eigen uses PartialPivLU for sizes > 4).  Returns false if a
// zero pivot is met.
template <int N>
CM_HD bool inverse_lu(const float* Ain, float* inv) {
  float A[N * N];
  int piv[N];
  for (int i = 0; i < N * N; ++i) A[i] = Ain[i];
  for (int i = 0; i < N; ++i) piv[i] = i;
  for (int k = 0; k < N; ++k) {
    int p = k;
    float best = fabsf(A[k * N + k]);
    for (int r = k + 1; r < N; ++r)
      if (fabsf(A[r * N + k]) > best) { best = fabsf(A[r * N + k]); p = r; }
    if (best == 0.f) return false;
    if (p != k) {
      for (int c = 0; c < N; ++c) { float t = A[k * N + c]; A[k * N + c] = A[p * N + c]; A[p * N + c] = t; }
      int t = piv[k]; piv[k] = piv[p]; piv[p] = t;
    }
    for (int r = k + 1; r < N; ++r) {
      A[r * N + k] /= A[k * N + k];
      float l = A[r * N + k];
      for (int c = k + 1; c < N; ++c) A[r * N + c] -= l * A[k * N + c];
    }
  }
  float y[N];
  for (int col = 0; col < N; ++col) {
    for (int r = 0; r < N; ++r) {
      float s = (piv[r] == col) ? 1.f : 0.f;
      for (int c = 0; c < r; ++c) s -= A[r * N + c] * y[c];
      y[r] = s;
    }
    for (int r = N - 1; r >= 0; --r) {
      float s = y[r];
      for (int c = r + 1; c < N; ++c) s -= A[r * N + c] * y[c];
      y[r] = s / A[r * N + r];
    }
    for (int r = 0; r < N; ++r) inv[r * N + col] = y[r];
  }
  return true;
}

// ----------------------------------------------------------------------------------------------------------
// Pose (Twist = rx, ry, rz, tx, ty, tz) -> rotation matrix, the way the reference builds it:
// convertTransform(Twist&, Isometry3f&) -> getTransformationTZYX (transform_utils.h:288-299, 308-311):
// q = AngleAxis(yaw,Z) * AngleAxis(pitch,Y) * AngleAxis(roll,X); R = q.toRotationMatrix().
// ----------------------------------------------------------------------------------------------------------
struct Quat { float w, x, y, z; };
CM_HD Quat quat_mul(const Quat& a, const Quat& b) {
  Quat r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  return r;
}
CM_HD void quat_to_matrix(const Quat& q, float R[9]) {
  float tx = 2.f * q.x, ty = 2.f * q.y, tz = 2.f * q.z;
  float twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  float txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  float tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1.f - (tyy + tzz); R[1] = txy - twz;         R[2] = txz + twy;
  R[3] = txy + twz;         R[4] = 1.f - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy;         R[7] = tyz + twx;         R[8] = 1.f - (txx + tyy);
}
CM_HD void pose_to_matrix(const float pose[6], float R[9]) {
  float sr, cr, sp, cp, sy, cy;
  cm_sincosf(0.5f * pose[0], &sr, &cr);
  cm_sincosf(0.5f * pose[1], &sp, &cp);
  cm_sincosf(0.5f * pose[2], &sy, &cy);
  Quat qz = {cy, 0.f, 0.f, sy}, qy = {cp, 0.f, sp, 0.f}, qx = {cr, sr, 0.f, 0.f};
  Quat q = quat_mul(quat_mul(qz, qy), qx);
  quat_to_matrix(q, R);
}
// p_map = R * p + t   (Isometry3f * Vector3f, pointAssociateToMap transform_utils.h:476-482)
CM_HD void transform_point(const float R[9], const float t[3], float x, float y, float z, float* ox, float* oy,
                           float* oz) {
  *ox = ((R[0] * x + R[1] * y) + R[2] * z) + t[0];
  *oy = ((R[3] * x + R[4] * y) + R[5] * z) + t[1];
  *oz = ((R[6] * x + R[7] * y) + R[8] * z) + t[2];
}

}  // namespace cm
