// Small fixed-size math helpers shared by every phase of the fused step (device + host-emulation builds).
#pragma once
#include <math.h>

#ifdef __CUDACC__
#define BT_DEV __device__ __forceinline__
#define BT_NOINLINE __device__ __noinline__
#define BT_LDG(p) __ldg(p)
#else
#define BT_DEV inline
#define BT_NOINLINE inline
#define BT_LDG(p) (*(p))
#endif

#define BT_MINVAL 1e-15f
#define BT_MINIMP 0.0001f
#define BT_MAXIMP 0.9999f

BT_DEV float bt_dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
BT_DEV void bt_cross(const float* a, const float* b, float* o) {
  float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
BT_DEV void bt_quat_mul(const float* a, const float* b, float* o) {
  float w = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  float x = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  float y = a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1];
  float z = a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
BT_DEV void bt_quat_normalize(float* q) {
  float n = 1.0f / sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] *= n; q[1] *= n; q[2] *= n; q[3] *= n;
}
// v rotated by unit quaternion q (same formula as brax.math.rotate / mjx math.rotate)
BT_DEV void bt_rotate(const float* v, const float* q, float* o) {
  float s = q[0];
  const float* u = q + 1;
  float uv = bt_dot3(u, v), uu = bt_dot3(u, u), c[3];
  bt_cross(u, v, c);
  float k = s * s - uu;
  float r0 = 2 * uv * u[0] + k * v[0] + 2 * s * c[0];
  float r1 = 2 * uv * u[1] + k * v[1] + 2 * s * c[1];
  float r2 = 2 * uv * u[2] + k * v[2] + 2 * s * c[2];
  o[0] = r0; o[1] = r1; o[2] = r2;
}
BT_DEV void bt_quat_to_mat(const float* q, float* m) {
  float w = q[0], x = q[1], y = q[2], z = q[3];
  m[0] = w * w + x * x - y * y - z * z; m[1] = 2 * (x * y - w * z); m[2] = 2 * (x * z + w * y);
  m[3] = 2 * (x * y + w * z); m[4] = w * w - x * x + y * y - z * z; m[5] = 2 * (y * z - w * x);
  m[6] = 2 * (x * z - w * y); m[7] = 2 * (y * z + w * x); m[8] = w * w - x * x - y * y + z * z;
}
BT_DEV void bt_mat_vec(const float* m, const float* v, float* o) {
  float a = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  float b = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  float c = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  o[0] = a; o[1] = b; o[2] = c;
}
BT_DEV void bt_matT_vec(const float* m, const float* v, float* o) {
  float a = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  float b = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  float c = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  o[0] = a; o[1] = b; o[2] = c;
}
// 10-number spatial inertia [Ixx Iyy Izz Ixy Ixz Iyz, m*off(3), m] times motion vector [ang; lin]
BT_DEV void bt_inert_mul(const float* i, const float* v, float* o) {
  float c[3], e[3];
  bt_cross(i + 6, v + 3, c);
  bt_cross(i + 6, v, e);
  float o0 = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] + c[0];
  float o1 = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + c[1];
  float o2 = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] + c[2];
  float o3 = i[9] * v[3] - e[0], o4 = i[9] * v[4] - e[1], o5 = i[9] * v[5] - e[2];
  o[0] = o0; o[1] = o1; o[2] = o2; o[3] = o3; o[4] = o4; o[5] = o5;
}
// motion cross product u x v
BT_DEV void bt_motion_cross(const float* u, const float* v, float* o) {
  float a[3], b[3], c[3];
  bt_cross(u, v, a);
  bt_cross(u + 3, v, b);
  bt_cross(u, v + 3, c);
  o[0] = a[0]; o[1] = a[1]; o[2] = a[2];
  o[3] = b[0] + c[0]; o[4] = b[1] + c[1]; o[5] = b[2] + c[2];
}
// force cross product v x* f
BT_DEV void bt_motion_cross_force(const float* v, const float* f, float* o) {
  float a[3], b[3], c[3];
  bt_cross(v, f, a);
  bt_cross(v + 3, f + 3, b);
  bt_cross(v, f + 3, c);
  o[0] = a[0] + b[0]; o[1] = a[1] + b[1]; o[2] = a[2] + b[2];
  o[3] = c[0]; o[4] = c[1]; o[5] = c[2];
}
// 6-term dot product / axpy of the chain sweeps.  On sm_100a they use the packed fp32 pipe (FMUL2 / FFMA2: two lanes of a
// 64-bit register pair per instruction): 3 packed + 1 scalar instruction instead of 7, 3 instead of 6 -- these sit on the
// critical path of every sweep.  Elsewhere: two independent 3-term chains.
BT_DEV float bt_dot6(const float* a, const float* b) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  float2 t = __fmul2_rn(make_float2(a[0], a[1]), make_float2(b[0], b[1]));
  t = __ffma2_rn(make_float2(a[2], a[3]), make_float2(b[2], b[3]), t);
  t = __ffma2_rn(make_float2(a[4], a[5]), make_float2(b[4], b[5]), t);
  return t.x + t.y;
#else
  const float x = a[0] * b[0] + a[1] * b[1] + a[2] * b[2];
  const float y = a[3] * b[3] + a[4] * b[4] + a[5] * b[5];
  return x + y;
#endif
}
// y += x * s
BT_DEV void bt_axpy6(float* y, const float* x, float s) {
#if defined(__CUDA_ARCH__) && __CUDA_ARCH__ >= 1000
  const float2 ss = make_float2(s, s);
#pragma unroll
  for (int j = 0; j < 6; j += 2) {
    const float2 r = __ffma2_rn(make_float2(x[j], x[j + 1]), ss, make_float2(y[j], y[j + 1]));
    y[j] = r.x; y[j + 1] = r.y;
  }
#else
  for (int j = 0; j < 6; j++) y[j] += x[j] * s;
#endif
}
// reciprocal of a positive, normal number (the pivots D_k): MUFU.RCP + one Newton step (<= 1 ulp), no range check / slow path
BT_DEV float bt_rcp_pos(float x) {
#ifdef __CUDACC__
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return fmaf(r, fmaf(-x, r, 1.0f), r);
#else
  return 1.0f / x;
#endif
}
// a / b where b is known to be a normal, non-zero number (impedances, regularisers, lengths behind a > 0 guard): a times the
// refined reciprocal (<= 1 ulp of 1/b), without the range check and out-of-line slow path of an IEEE division
BT_DEV float bt_div(float a, float b) {
#ifdef __CUDACC__
  return a * bt_rcp_pos(b);
#else
  return a / b;
#endif
}
// 16-byte aligned 12-float record / its first 6 floats (shared memory): 128-bit loads on the device
BT_DEV void bt_ld12(const float* p, float* o) {
#ifdef __CUDACC__
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4), c = *reinterpret_cast<const float4*>(p + 8);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w; o[8] = c.x; o[9] = c.y; o[10] = c.z; o[11] = c.w;
#else
  for (int j = 0; j < 12; j++) o[j] = p[j];
#endif
}
// 16-byte aligned 12-float record store (a then b)
BT_DEV void bt_st12(float* p, const float* a, const float* b) {
#ifdef __CUDACC__
  *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
  *reinterpret_cast<float4*>(p + 4) = make_float4(a[4], a[5], b[0], b[1]);
  *reinterpret_cast<float4*>(p + 8) = make_float4(b[2], b[3], b[4], b[5]);
#else
  for (int j = 0; j < 6; j++) { p[j] = a[j]; p[6 + j] = b[j]; }
#endif
}
// 4 / 2 consecutive table floats through the read-only path (16- / 8-byte aligned)
BT_DEV void bt_ldg4(const float* p, float* o) {
#ifdef __CUDACC__
  const float4 v = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
#else
  for (int j = 0; j < 4; j++) o[j] = p[j];
#endif
}
BT_DEV void bt_ldg2(const float* p, float* o) {
#ifdef __CUDACC__
  const float2 v = __ldg(reinterpret_cast<const float2*>(p));
  o[0] = v.x; o[1] = v.y;
#else
  o[0] = p[0]; o[1] = p[1];
#endif
}
// one 16-byte aligned quaternion (shared memory): a single 128-bit access instead of four stride-4 scalar ones (which are
// 4-way bank conflicts when a warp walks consecutive bodies)
BT_DEV void bt_ld4(const float* p, float* o) {
#ifdef __CUDACC__
  const float4 v = *reinterpret_cast<const float4*>(p);
  o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
#else
  for (int j = 0; j < 4; j++) o[j] = p[j];
#endif
}
BT_DEV void bt_st4(float* p, const float* a) {
#ifdef __CUDACC__
  *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
#else
  for (int j = 0; j < 4; j++) p[j] = a[j];
#endif
}
// 6 floats to a 16-byte aligned slot
BT_DEV void bt_st6(float* p, const float* a) {
#ifdef __CUDACC__
  *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
  *reinterpret_cast<float2*>(p + 4) = make_float2(a[4], a[5]);
#else
  for (int j = 0; j < 6; j++) p[j] = a[j];
#endif
}
BT_DEV void bt_ld6(const float* p, float* o) {
#ifdef __CUDACC__
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float2 b = *reinterpret_cast<const float2*>(p + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y;
#else
  for (int j = 0; j < 6; j++) o[j] = p[j];
#endif
}
BT_DEV float bt_clampf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
// jnp.nan_to_num (fruitfly.py:569-570)
BT_DEV float bt_nan_to_num(float x) {
  if (x != x) return 0.0f;
  if (x > 3.4028234664e38f) return 3.4028234664e38f;
  if (x < -3.4028234664e38f) return -3.4028234664e38f;
  return x;
}

// solimp sigmoid for power != 2 (mj_makeImpedance); out of line: powf is ~500 instructions when inlined
#ifdef __CUDACC__
static __device__ __noinline__ float bt_impedance_pow(float x, float mid, float power) {
#else
static inline float bt_impedance_pow(float x, float mid, float power) {
#endif
  const float ia = (1.f / powf(mid, power - 1.f)) * powf(x, power);
  const float ib = 1.f - (1.f / powf(1.f - mid, power - 1.f)) * powf(1.f - x, power);
  return x < mid ? ia : ib;
}

BT_DEV int bt_clampi(int x, int lo, int hi) { return x < lo ? lo : (x > hi ? hi : x); }

// ---- JAX threefry2x32 ----
BT_DEV unsigned bt_rotl(unsigned x, int d) { return (x << d) | (x >> (32 - d)); }
BT_DEV void bt_threefry2x32(unsigned k0, unsigned k1, unsigned& x0, unsigned& x1) {
  unsigned ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
  const int r0[4] = {13, 15, 26, 6}, r1[4] = {17, 29, 16, 24};
  x0 += ks[0];
  x1 += ks[1];
#pragma unroll
  for (int g = 0; g < 5; g++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      x0 += x1;
      x1 = bt_rotl(x1, (g & 1) ? r1[r] : r0[r]);
      x1 ^= x0;
    }
    x0 += ks[(g + 1) % 3];
    x1 += ks[(g + 2) % 3] + (unsigned)(g + 1);
  }
}
// element idx of threefry_2x32(key, iota(n)) with JAX's split-in-halves convention.  An odd-length counter array is padded
// with a ZERO (jax/_src/prng.py threefry_2x32: concatenate([count, uint32([0])])), so the last element of the first half
// pairs with counter 0, not n (known answer: uniform(PRNGKey(0), ()) = 0.41845703, tests/test_prng.py)
BT_DEV unsigned bt_random_bits(unsigned k0, unsigned k1, int idx, int n) {
  int half = (n + 1) / 2;
  unsigned x0, x1;
  if (idx < half) { x0 = (unsigned)idx; x1 = half + idx < n ? (unsigned)(half + idx) : 0u; }
  else { x0 = (unsigned)(idx - half); x1 = (unsigned)idx; }
  bt_threefry2x32(k0, k1, x0, x1);
  return idx < half ? x0 : x1;
}
BT_DEV float bt_bits_to_uniform(unsigned bits, float lo, float hi) {
  unsigned u = (bits >> 9) | 0x3F800000u;
  float f;
#ifdef __CUDACC__
  f = __uint_as_float(u);
#else
  union { unsigned u; float f; } cv; cv.u = u; f = cv.f;
#endif
  f -= 1.0f;
  // separate multiply and add (no FMA contraction) so the bits match jax.random.uniform on the XLA CPU backend
#ifdef __CUDACC__
  float r = __fadd_rn(__fmul_rn(f, hi - lo), lo);
#else
  volatile float prod = f * (hi - lo);
  float r = prod + lo;
#endif
  return r < lo ? lo : r;
}
