// bvh.cuh -- compressed 8-wide BVH (two-level: TLAS over instances, one BLAS per mesh) and the
// traversal routines of the Closest / Shadow stages.
//
// The reference delegates this to closed NVIDIA OptiX running on RT cores (optixTrace at
// src/render/wavefront/device.cu:13-22; accel build src/core/device/optix.cpp:143-250, 357-398).
// B200 has no RT cores, so traversal is SM code:
//   * 80-byte nodes: an anchor point + per-axis power-of-two scale, and 8 children whose boxes are
//     quantised to 8 bits per plane (conservatively: lo rounded down, hi rounded up).  One node is
//     five 16-byte loads.
//   * BLAS nodes live in OBJECT space of their mesh; an instance is entered by transforming the ray
//     with the instance's inverse 3x4 (t stays the world-space parameter, the direction is not
//     renormalised), exactly as the intersection spec in oracle/driver.cpp states.
//   * ray/triangle: Moeller-Trumbore with individually rounded operations (xmul/xadd..., never
//     FMA-contracted), accept 0 < t < tmax, ties on t broken by (instance, primitive) so that the
//     result does not depend on traversal order -> first-hit ids are bit-exact against the oracle's
//     brute-force loop.
//   * per-thread traversal with a short stack; hit internal children are pushed far-to-near.
#pragma once
#include "krr_math.cuh"
#include "scene.cuh"

namespace krr {

struct __align__(16) Node8 {
	float ox, oy, oz;		  // anchor (min corner of the node box)
	uint8_t ex, ey, ez;		  // biased exponents: child plane = o + q * 2^(e-127)
	uint8_t imask;			  // bit i: child i is an internal node
	uint32_t childBase;		  // first internal child; child i -> childBase + popc(imask & ((1<<i)-1))
	uint32_t primBase;		  // first primitive of the leaf children (BLAS: triangle pool, TLAS: instance list)
	uint8_t meta[8];		  // leaf child: (count << 5) | offset from primBase; 0 = empty / internal
	uint8_t qlo[3][8], qhi[3][8];
};
static_assert(sizeof(Node8) == 80, "Node8 must be 80 bytes");

struct BvhTri { float4 v0, v1, v2; }; // xyz = object-space vertex, v0.w = primitive id (int bits)

struct BvhDev {
	const Node8 *nodes;	   // node pool: TLAS nodes first, then every mesh's BLAS
	const BvhTri *tris;	   // triangle pool, leaf order per mesh
	const int32_t *tlasInst; // instance ids referenced by TLAS leaves
	int32_t tlasRoot;
	int32_t nInstances;
};

struct Hit {
	int32_t inst, prim;
	float t, u, v;
};

// ---- the intersection spec (keep identical to oracle/driver.cpp triIntersect) ----
KRR_HD bool triIntersect(V3 o, V3 d, V3 v0, V3 v1, V3 v2, float tmax, float &t, float &u, float &v) {
	V3 e1 = xsub3(v1, v0), e2 = xsub3(v2, v0);
	V3 pv = xcross(d, e2);
	float det = xdot(e1, pv);
	if (det == 0.f) return false;
	float inv = xdiv(1.f, det);
	V3 tv = xsub3(o, v0);
	u = xmul(xdot(tv, pv), inv);
	if (!(u >= 0.f && u <= 1.f)) return false;
	V3 qv = xcross(tv, e1);
	v = xmul(xdot(d, qv), inv);
	if (!(v >= 0.f && xadd(u, v) <= 1.f)) return false;
	t = xmul(xdot(e2, qv), inv);
	return t > 0.f && t < tmax;
}
KRR_HD bool betterHit(float t, int inst, int prim, const Hit &h) {
	if (h.inst < 0) return true;
	if (t != h.t) return t < h.t;
	if (inst != h.inst) return inst < h.inst;
	return prim < h.prim;
}

#ifdef __CUDACC__
constexpr int kStackSize = 96;
constexpr uint32_t kInstFlag = 0x80000000u;

// Accept(inst, prim, u, v) -> bool decides whether a candidate counts (any-hit programs: alpha
// kill, null-material skip).  ANY = true: return at the first accepted hit (shadow rays).
template <bool ANY, typename Accept>
KRR_DEV Hit traverse(const BvhDev &bvh, const InstRec *__restrict__ instances, V3 o, V3 d, float tmax,
					 Accept accept, int *overflow) {
	Hit best;
	best.inst = -1, best.prim = -1, best.t = tmax, best.u = best.v = 0;
	uint32_t stack[kStackSize];
	int sp = 0;
	V3 ro = o, rd = d; // current-space ray (world in TLAS, object in BLAS)
	V3 idir = mk3(1.f / rd.x, 1.f / rd.y, 1.f / rd.z);
	int curInst = -1, blasBase = -1;
	uint32_t cur = (uint32_t) bvh.tlasRoot;
	bool have = true;
	while (true) {
		if (!have) {
			if (sp == 0) break;
			if (curInst >= 0 && sp == blasBase) { // BLAS finished: back to world space
				curInst = -1;
				ro = o, rd = d;
				idir = mk3(1.f / rd.x, 1.f / rd.y, 1.f / rd.z);
			}
			cur = stack[--sp];
			if (cur & kInstFlag) {
				curInst = (int) (cur & ~kInstFlag);
				const InstRec &in = instances[curInst];
				ro = xfPointX(in.inv, o), rd = xfVectorX(in.inv, d);
				idir	 = mk3(1.f / rd.x, 1.f / rd.y, 1.f / rd.z);
				blasBase = sp;
				cur		 = (uint32_t) in.blasRoot;
			}
		}
		have = false;
		// ---- fetch node (5 x 16 B) ----
		const float4 *np = reinterpret_cast<const float4 *>(bvh.nodes + cur);
		float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
		uint32_t ew	  = __float_as_uint(n0.w);
		float sx	  = __uint_as_float((ew & 0xff) << 23), sy = __uint_as_float(((ew >> 8) & 0xff) << 23),
			  sz	  = __uint_as_float(((ew >> 16) & 0xff) << 23);
		uint32_t imask = ew >> 24;
		uint32_t childBase = __float_as_uint(n1.x), primBase = __float_as_uint(n1.y);
		uint32_t metaLo = __float_as_uint(n1.z), metaHi = __float_as_uint(n1.w);
		// quantised planes: n2 = qlo[0][0..7], qlo[1][0..7]; n3 = qlo[2], qhi[0]; n4 = qhi[1], qhi[2]
		uint32_t q[12] = {__float_as_uint(n2.x), __float_as_uint(n2.y), __float_as_uint(n2.z), __float_as_uint(n2.w),
						  __float_as_uint(n3.x), __float_as_uint(n3.y), __float_as_uint(n3.z), __float_as_uint(n3.w),
						  __float_as_uint(n4.x), __float_as_uint(n4.y), __float_as_uint(n4.z), __float_as_uint(n4.w)};
		// ray in node-local units: t = (o_node + q*s - o) * idir = q * (s*idir) + (o_node - o)*idir
		float ax = sx * idir.x, ay = sy * idir.y, az = sz * idir.z;
		float bx = (n0.x - ro.x) * idir.x, by = (n0.y - ro.y) * idir.y, bz = (n0.z - ro.z) * idir.z;
		const float lim = best.t;
		uint32_t keys[8];
		int nk = 0;
#pragma unroll
		for (int i = 0; i < 8; i++) {
			uint32_t meta = ((i < 4 ? metaLo : metaHi) >> ((i & 3) * 8)) & 0xff;
			bool internal = (imask >> i) & 1;
			if (!internal && meta == 0) continue;
			auto qb = [&](int row) { return (float) ((q[row * 2 + (i >> 2)] >> ((i & 3) * 8)) & 0xff); };
			float lx = qb(0), ly = qb(1), lz = qb(2), hx = qb(3), hy = qb(4), hz = qb(5);
			float t0x = lx * ax + bx, t1x = hx * ax + bx;
			float t0y = ly * ay + by, t1y = hy * ay + by;
			float t0z = lz * az + bz, t1z = hz * az + bz;
			// fminf/fmaxf drop NaNs (0 * inf), which is the conservative answer for a degenerate slab
			float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.f));
			float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), lim));
			// conservative: boxes only cull, the exact decision is the triangle test
			if (!(tn <= tf * 1.0000010f + 1e-30f)) continue;
			if (internal) {
				// key: distance in the high bits, slot in the low 3 -> one compare orders (tn, slot)
				keys[nk++] = (__float_as_uint(tn) & ~7u) | (uint32_t) i;
			} else {
				uint32_t cnt = meta >> 5, off = meta & 31;
				if (curInst < 0) {
					// TLAS leaf: defer each instance (entered when popped)
					for (uint32_t k = 0; k < cnt; k++) {
						if (sp >= kStackSize) { *overflow = 1; continue; }
						stack[sp++] = kInstFlag | (uint32_t) __ldg(bvh.tlasInst + primBase + off + k);
					}
				} else {
					for (uint32_t k = 0; k < cnt; k++) {
						const float4 *tp = reinterpret_cast<const float4 *>(bvh.tris + primBase + off + k);
						float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
						float t, u, v;
						if (triIntersect(ro, rd, mk3(a), mk3(b), mk3(c), tmax, t, u, v)) {
							int prim = __float_as_int(a.w);
							if (betterHit(t, curInst, prim, best) && accept(curInst, prim, u, v)) {
								best.inst = curInst, best.prim = prim, best.t = t, best.u = u, best.v = v;
								if (ANY) return best;
							}
						}
					}
				}
			}
		}
		if (nk) {
			// sort far-to-near so the nearest child is popped first (insertion sort, nk <= 8)
			for (int i = 1; i < nk; i++) {
				uint32_t k = keys[i];
				int j = i - 1;
				while (j >= 0 && keys[j] < k) { keys[j + 1] = keys[j]; j--; }
				keys[j + 1] = k;
			}
			// continue with the nearest, push the rest
			for (int i = 0; i < nk - 1; i++) {
				uint32_t slot = keys[i] & 7u;
				if (sp >= kStackSize) { *overflow = 1; continue; }
				stack[sp++] = childBase + __popc(imask & ((1u << slot) - 1));
			}
			uint32_t slot = keys[nk - 1] & 7u;
			cur	 = childBase + __popc(imask & ((1u << slot) - 1));
			have = true;
		}
	}
	if (best.inst < 0) best.t = tmax;
	return best;
}
#endif // __CUDACC__

} // namespace krr
