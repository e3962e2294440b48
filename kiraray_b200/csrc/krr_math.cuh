// krr_math.cuh -- small vector / spectrum types for the sm_100a kernels.
// Spectrum = 4 floats (KRR_N_SPECTRUM_SAMPLES, reference src/core/config.in.h:15) held in a float4
// so every queue field moves as one 16-byte load/store.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define KRR_DEV __device__ __forceinline__
#define KRR_HD __host__ __device__ __forceinline__

namespace krr {

constexpr float kPi		= 3.14159265358979323846f;
constexpr float k2Pi	= 6.28318530717958647693f;
constexpr float kInvPi	= 0.318309886183790671538f;
constexpr float kInv2Pi = 0.15915494309189533577f;
constexpr float kInv4Pi = 0.07957747154594766788f;
constexpr float kRayEps = 1e-4f; // KRR_RAY_EPS, src/core/raytracing.h:10
constexpr float kInf	= __builtin_huge_valf();

// ---- "exact" arithmetic: individually rounded, never contracted to FMA.  Used wherever results are
// compared bit-for-bit with the CPU oracle (sampler->float, wavelengths, camera rays, intersection).
#ifdef __CUDA_ARCH__
KRR_DEV float xmul(float a, float b) { return __fmul_rn(a, b); }
KRR_DEV float xadd(float a, float b) { return __fadd_rn(a, b); }
KRR_DEV float xsub(float a, float b) { return __fsub_rn(a, b); }
KRR_DEV float xdiv(float a, float b) { return __fdiv_rn(a, b); }
KRR_DEV float xrcp(float a) { return __frcp_rn(a); } // == __fdiv_rn(1.f, a): both are the correctly rounded reciprocal
KRR_DEV float xsqrt(float a) { return __fsqrt_rn(a); }
#else
inline float xmul(float a, float b) { return a * b; }
inline float xadd(float a, float b) { return a + b; }
inline float xsub(float a, float b) { return a - b; }
inline float xdiv(float a, float b) { return a / b; }
inline float xrcp(float a) { return 1.f / a; }
inline float xsqrt(float a) { return sqrtf(a); }
#endif

struct V3 {
	float x, y, z;
	KRR_HD float operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }
};
KRR_HD V3 mk3(float x, float y, float z) { return V3{x, y, z}; }
KRR_HD V3 mk3(float4 v) { return V3{v.x, v.y, v.z}; }
KRR_HD V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
KRR_HD V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
KRR_HD V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
KRR_HD V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
KRR_HD V3 operator*(float s, V3 a) { return mk3(a.x * s, a.y * s, a.z * s); }
KRR_HD V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
KRR_HD float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
KRR_HD V3 cross(V3 a, V3 b) { return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
KRR_HD float length(V3 a) { return sqrtf(dot(a, a)); }
KRR_HD V3 normalize(V3 a) {
	float z = dot(a, a);
	return z > 0 ? a / sqrtf(z) : a; // Eigen normalized(): returned unchanged when the norm is 0
}
KRR_HD bool anyNonZero(V3 a) { return a.x != 0 || a.y != 0 || a.z != 0; }

// exact variants
KRR_HD V3 xsub3(V3 a, V3 b) { return mk3(xsub(a.x, b.x), xsub(a.y, b.y), xsub(a.z, b.z)); }
KRR_HD float xdot(V3 a, V3 b) { return xadd(xadd(xmul(a.x, b.x), xmul(a.y, b.y)), xmul(a.z, b.z)); }
KRR_HD V3 xcross(V3 a, V3 b) {
	return mk3(xsub(xmul(a.y, b.z), xmul(a.z, b.y)), xsub(xmul(a.z, b.x), xmul(a.x, b.z)),
			   xsub(xmul(a.x, b.y), xmul(a.y, b.x)));
}

// ---- spectrum (float4) ----
typedef float4 Spec;
KRR_HD Spec sp(float c) { return make_float4(c, c, c, c); }
KRR_HD Spec operator+(Spec a, Spec b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
KRR_HD Spec operator-(Spec a, Spec b) { return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w); }
KRR_HD Spec operator*(Spec a, Spec b) { return make_float4(a.x * b.x, a.y * b.y, a.z * b.z, a.w * b.w); }
KRR_HD Spec operator/(Spec a, Spec b) { return make_float4(a.x / b.x, a.y / b.y, a.z / b.z, a.w / b.w); }
KRR_HD Spec operator*(Spec a, float s) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
KRR_HD Spec operator*(float s, Spec a) { return make_float4(a.x * s, a.y * s, a.z * s, a.w * s); }
KRR_HD Spec operator/(Spec a, float s) { return make_float4(a.x / s, a.y / s, a.z / s, a.w / s); }
KRR_HD Spec &operator+=(Spec &a, Spec b) { a = a + b; return a; }
KRR_HD Spec &operator*=(Spec &a, Spec b) { a = a * b; return a; }
KRR_HD Spec &operator*=(Spec &a, float s) { a = a * s; return a; }
KRR_HD Spec &operator/=(Spec &a, float s) { a = a / s; return a; }
KRR_HD bool any(Spec a) { return a.x != 0 || a.y != 0 || a.z != 0 || a.w != 0; }
// Eigen mean(): pairwise redux then / 4
KRR_HD float mean(Spec a) { return ((a.x + a.y) + (a.z + a.w)) / 4; }
KRR_HD float maxCoeff(Spec a) { return fmaxf(fmaxf(a.x, a.y), fmaxf(a.z, a.w)); }
KRR_HD Spec cwiseMax(Spec a, float m) { return make_float4(fmaxf(a.x, m), fmaxf(a.y, m), fmaxf(a.z, m), fmaxf(a.w, m)); }
KRR_HD Spec cwiseMin(Spec a, float m) { return make_float4(fminf(a.x, m), fminf(a.y, m), fminf(a.z, m), fminf(a.w, m)); }
KRR_HD Spec sqrtS(Spec a) { return make_float4(sqrtf(a.x), sqrtf(a.y), sqrtf(a.z), sqrtf(a.w)); }
KRR_HD Spec expS(Spec a) { return make_float4(expf(a.x), expf(a.y), expf(a.z), expf(a.w)); }
KRR_HD float at(const Spec &a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : (i == 2 ? a.z : a.w)); }
KRR_HD bool hasNaN(Spec a) { return a.x != a.x || a.y != a.y || a.z != a.z || a.w != a.w; }
// lerp(x, y, w) = (1 - w) * x + w * y (krrmath/functors.h, scalar form; Eigen form a*(1-t) + b*t)
KRR_HD float lerpf(float x, float y, float w) { return (1.f - w) * x + w * y; }
KRR_HD Spec lerpS(Spec a, Spec b, float t) { return a * (1 - t) + b * t; }
KRR_HD float pow2(float x) { return x * x; }
KRR_HD float pow5(float x) { return x * x * x * x * x; }
KRR_HD float clampf(float v, float lo, float hi) { return fmaxf(fminf(v, hi), lo); }
KRR_HD float safe_sqrt(float v) { return sqrtf(fmaxf(0.f, v)); }

// 3x4 row-major affine
struct Xf { float m[12]; };
KRR_HD V3 xfPoint(const Xf &t, V3 p) {
	return mk3(t.m[0] * p.x + t.m[1] * p.y + t.m[2] * p.z + t.m[3], t.m[4] * p.x + t.m[5] * p.y + t.m[6] * p.z + t.m[7],
			   t.m[8] * p.x + t.m[9] * p.y + t.m[10] * p.z + t.m[11]);
}
KRR_HD V3 xfVector(const Xf &t, V3 p) {
	return mk3(t.m[0] * p.x + t.m[1] * p.y + t.m[2] * p.z, t.m[4] * p.x + t.m[5] * p.y + t.m[6] * p.z,
			   t.m[8] * p.x + t.m[9] * p.y + t.m[10] * p.z);
}
KRR_HD V3 xfNormal(const Xf &inv, V3 n) { // (M^-1)^T n
	return mk3(inv.m[0] * n.x + inv.m[4] * n.y + inv.m[8] * n.z, inv.m[1] * n.x + inv.m[5] * n.y + inv.m[9] * n.z,
			   inv.m[2] * n.x + inv.m[6] * n.y + inv.m[10] * n.z);
}
// exact (unfused, left-to-right) variants used by the intersection spec
KRR_HD V3 xfPointX(const Xf &t, V3 p) {
	return mk3(xadd(xadd(xadd(xmul(t.m[0], p.x), xmul(t.m[1], p.y)), xmul(t.m[2], p.z)), t.m[3]),
			   xadd(xadd(xadd(xmul(t.m[4], p.x), xmul(t.m[5], p.y)), xmul(t.m[6], p.z)), t.m[7]),
			   xadd(xadd(xadd(xmul(t.m[8], p.x), xmul(t.m[9], p.y)), xmul(t.m[10], p.z)), t.m[11]));
}
KRR_HD V3 xfVectorX(const Xf &t, V3 p) {
	return mk3(xadd(xadd(xmul(t.m[0], p.x), xmul(t.m[1], p.y)), xmul(t.m[2], p.z)),
			   xadd(xadd(xmul(t.m[4], p.x), xmul(t.m[5], p.y)), xmul(t.m[6], p.z)),
			   xadd(xadd(xmul(t.m[8], p.x), xmul(t.m[9], p.y)), xmul(t.m[10], p.z)));
}

} // namespace krr
