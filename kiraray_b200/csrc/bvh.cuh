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
//   * warp-cooperative traversal (see Traverser below): phase-aligned stepping, shared-memory short
//     stack, hit children sorted by entry distance with a sorting network, persistent warps that
//     refill finished lanes from the queue.
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

// v0.xyz = first vertex, e1 = v1 - v0, e2 = v2 - v0 (the first two operations of the triangle test, done once
// at build time with the same rounding); v0.w = primitive id, e1.w = instance id (merged BLAS only)
struct BvhTri { float4 v0, e1, e2; };

struct BvhDev {
	const Node8 *nodes;	   // node pool: TLAS nodes first, then every mesh's BLAS
	const BvhTri *tris;	   // triangle pool, leaf order per mesh
	const int32_t *tlasInst; // instance ids referenced by TLAS leaves
	int32_t tlasRoot;
	int32_t nInstances;
	const XformNodeRec *xnodes; // motion blur: transform chains + SRT key pool (motion.cuh), null otherwise
	const float *motionKeys;
	// Static instances whose transform is exactly the identity are MERGED into one world-space BLAS
	// (bvh_build.cu): object space == world space for them, so the intersection spec gives the same
	// numbers, and a ray no longer enters each of them separately.  The merged BLAS hangs in the TLAS as
	// pseudo-instance `mergedInst` (= nInstances, identity transform); its triangles carry their real
	// instance id.  mergedOnly: every instance was merged, traversal starts at the merged root.
	int32_t mergedInst; // -1 = nothing merged
	int32_t mergedRoot;
	int32_t mergedOnly;
	// A BLAS with at most `flat_blas_max` triangles is not a tree but a FLAT LIST: its root entry is
	// kFlatFlag | index into flats[] = (first triangle, count).  A warp walks such a list in lock step
	// (no stack, no slab tests, no divergence between lanes), which for a few dozen triangles costs
	// fewer issue slots than the wide-node traversal it replaces.
	const int2 *flats;
	// world box of the merged BLAS (padded): with mergedOnly there is no TLAS above it whose node test would
	// reject rays that miss the scene altogether, so begin() tests it (primary rays beside the Cornell box
	// skip its 36 triangles)
	float rootLo[3], rootHi[3];
};

struct Hit {
	int32_t inst, prim;
	float t, u, v;
};

// ---- the intersection spec (keep identical to oracle/driver.cpp triIntersect) ----
KRR_HD bool triIntersectE(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float tmax, float &t, float &u, float &v);
KRR_HD bool triIntersect(V3 o, V3 d, V3 v0, V3 v1, V3 v2, float tmax, float &t, float &u, float &v) {
	return triIntersectE(o, d, v0, xsub3(v1, v0), xsub3(v2, v0), tmax, t, u, v);
}
// the same test on a triangle stored as (v0, e1, e2)
#ifndef KRR_TRI_RCP
#define KRR_TRI_RCP 1
#endif
#ifndef KRR_LEAF_PAIR
#define KRR_LEAF_PAIR 1
#endif
#ifndef KRR_LEAF_WIDTH
#define KRR_LEAF_WIDTH 2
#endif
#ifndef KRR_LEAF_PAIR_TREE
#define KRR_LEAF_PAIR_TREE 0
#endif
// 1 / det, correctly rounded.  On the device __frcp_rn: IEEE round-to-nearest of the reciprocal, i.e. the
// same float as __fdiv_rn(1.f, det) for every input, in about half the instructions
KRR_HD float triRcp(float det) {
#if defined(__CUDA_ARCH__) && KRR_TRI_RCP
	return __frcp_rn(det);
#else
	return xdiv(1.f, det);
#endif
}
KRR_HD bool triIntersectE(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float tmax, float &t, float &u, float &v) {
	V3 pv = xcross(d, e2);
	float det = xdot(e1, pv);
	if (det == 0.f) return false;
	float inv = triRcp(det);
	V3 tv = xsub3(o, v0);
	u = xmul(xdot(tv, pv), inv);
	if (!(u >= 0.f && u <= 1.f)) return false;
	V3 qv = xcross(tv, e1);
	v = xmul(xdot(d, qv), inv);
	if (!(v >= 0.f && xadd(u, v) <= 1.f)) return false;
	t = xmul(xdot(e2, qv), inv);
	return t > 0.f && t < tmax;
}
#ifdef __CUDACC__
// triIntersectE without early exits: every operation and rounding of the spec above, the three range tests
// combined at the end (a zero determinant makes inv infinite; `det != 0` rejects whatever that produces)
KRR_DEV bool triTestNoBranch(V3 o, V3 d, V3 v0, V3 e1, V3 e2, float tmax, float &t, float &u, float &v) {
	V3 pv = xcross(d, e2);
	float det = xdot(e1, pv);
	float inv = triRcp(det);
	V3 tv = xsub3(o, v0);
	u = xmul(xdot(tv, pv), inv);
	V3 qv = xcross(tv, e1);
	v = xmul(xdot(d, qv), inv);
	t = xmul(xdot(e2, qv), inv);
	return (det != 0.f) & (u >= 0.f) & (u <= 1.f) & (v >= 0.f) & (xadd(u, v) <= 1.f) & (t > 0.f) & (t < tmax);
}
#endif
KRR_HD bool betterHit(float t, int inst, int prim, const Hit &h) {
	if (h.inst < 0) return true;
	if (t != h.t) return t < h.t;
	if (inst != h.inst) return inst < h.inst;
	return prim < h.prim;
}

#ifdef __CUDACC__
// ---- traversal state machine ---------------------------------------------------------------------
// One lane = one ray.  The traversal is written as a STEP function so that the stage kernels can
// run it warp-cooperatively: every trip of the warp's loop executes the three phases
//     [enter instance] -> [wide node: 8 slab tests, sorting network, push] -> [leaf: triangle tests]
// under per-lane predicates, so lanes that are in the same phase execute it together, and the
// kernel refills lanes whose ray has terminated from the queue (persistent warps, one atomicAdd per
// refill) instead of letting them idle until the slowest ray of the warp is done.
//
// Stack: entries are 32-bit tagged words + the entry distance of the box (for culling at pop time
// once a closer hit is known).  The first kShortStack entries of every lane live in SHARED memory
// (slot-major, so a warp's accesses are conflict-free); deeper entries spill to local memory.
//   node     : index into the node pool                                     (bits 31,30 = 00)
//   leaf     : kLeafFlag | (count-1) << 26 | first triangle                 (bits 31,30 = 01)
//   instance : kInstFlag | instance id  (TLAS leaves hold ONE instance)      (bits 31,30 = 10)
//   flat BLAS: kFlatFlag | index into BvhDev::flats (only ever a BLAS root)  (bits 31,30 = 11, != empty)
#ifndef KRR_SHORT_STACK
#define KRR_SHORT_STACK 12
#endif
constexpr int kShortStack  = KRR_SHORT_STACK;
constexpr int kLocalStack  = 128 - KRR_SHORT_STACK; // 8-wide nodes defer up to 7 siblings per level: 128 entries cover TLAS + BLAS depths of ~18 levels
constexpr int kStackSize   = kShortStack + kLocalStack;
constexpr int kTraceBlock  = 128;
constexpr uint32_t kInstFlag = 0x80000000u, kLeafFlag = 0x40000000u, kFlatFlag = 0xc0000000u, kEmptyEntry = 0xffffffffu;
KRR_HD bool isFlatEntry(uint32_t e) { return (e >> 30) == 3u && e != kEmptyEntry; }

struct TraceSmem {
	uint32_t id[kShortStack][kTraceBlock];
	float tn[kShortStack][kTraceBlock];
};
// Spill part of the stack (local memory).  Deliberately NOT a member of Traverser: a dynamically
// indexed array inside the struct keeps the WHOLE struct in local memory (the compiler cannot split an
// aggregate that is indexed with a run-time value), and the ray state would be loaded and stored
// around every phase instead of living in registers.
template <bool ANY> struct LocalStack {
	uint32_t id[kLocalStack];
	float tn[ANY ? 1 : kLocalStack];
};

#define KRR_CSWAP(a, b) { uint32_t lo_ = min(a, b), hi_ = max(a, b); a = lo_; b = hi_; }

// world ray -> object space of a moving instance (kept out of line: static scenes never pay its registers)
static __device__ __noinline__ void movingRay(const BvhDev &bvh, int node, float time, V3 o, V3 d, V3 &ro, V3 &rd) {
	Xf m, inv;
	chainXf(bvh.xnodes, bvh.motionKeys, node, time, m, inv);
	ro = xfPointX(inv, o), rd = xfVectorX(inv, d);
}

// MOTION = false compiles the SRT-chain path out (static scenes keep their register budget)
// PAIR: the scene is one flat triangle list and leaf() walks it in branch-free pairs (see leaf())
template <bool ANY, bool MOTION = true, bool PAIR = false> struct Traverser {
	// ray
	V3 o, d;	  // world space
	V3 ro, rd;	  // current space (world in the TLAS, object space inside a BLAS)
	V3 idir;
	float tmax, time;
	Hit best;
	// control
	uint32_t cur;
	int sp, curInst, blasBase;
	int overflow;
	using LStack = LocalStack<ANY>;

	// reciprocal direction of the slab tests.  A zero component is replaced by +-1e-20: the distances to the two
	// planes of that slab become -+huge (origin inside the slab: the interval covers everything) or huge with
	// one sign (outside: the box is culled), all finite.  With 1/0 = inf they were NaN, the axis was dropped from
	// the test, and a ray parallel to two axes could only be culled along its own direction: it walked every
	// node in front of it (46 ms for one such ray in the 20 M-triangle scene).
	KRR_DEV void setIdir() {
		auto safe = [](float x) { return fabsf(x) >= 1e-20f ? x : copysignf(1e-20f, x); };
		idir = mk3(1.f / safe(rd.x), 1.f / safe(rd.y), 1.f / safe(rd.z));
	}
	KRR_DEV void begin(const BvhDev &bvh, V3 o_, V3 d_, float tmax_, float time_ = 0.f) {
		o = ro = o_, d = rd = d_, tmax = tmax_, time = time_;
		setIdir();
		best.inst = -1, best.prim = -1, best.t = tmax_, best.u = best.v = 0;
		cur = (uint32_t) bvh.tlasRoot, sp = 0, curInst = -1, blasBase = -1, overflow = 0;
		if (bvh.mergedOnly) {
			cur = (uint32_t) bvh.mergedRoot, curInst = bvh.mergedInst, blasBase = 0; // world == object space
			const float t0x = (bvh.rootLo[0] - o_.x) * idir.x, t1x = (bvh.rootHi[0] - o_.x) * idir.x;
			const float t0y = (bvh.rootLo[1] - o_.y) * idir.y, t1y = (bvh.rootHi[1] - o_.y) * idir.y;
			const float t0z = (bvh.rootLo[2] - o_.z) * idir.z, t1z = (bvh.rootHi[2] - o_.z) * idir.z;
			const float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.f));
			const float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tmax_));
			if (!(tn <= tf * 1.00001f + 1e-30f)) cur = kEmptyEntry, curInst = -1; // NaNs (0 * inf) are dropped by min / max
		}
		// A ray with a NaN / infinite component or a zero direction cannot hit anything (every comparison of
		// the triangle test fails, det == 0), but its slab tests cannot cull either: it would walk the WHOLE
		// tree (seconds on a 20 M-triangle scene).  Such rays come out of degenerate BSDF samples; they are
		// misses, as they are for the brute-force loop of the oracle.
#ifdef KRR_DEBUG_RAYS
		{
			const float mo = fmaxf(fabsf(o_.x), fmaxf(fabsf(o_.y), fabsf(o_.z))), md = fmaxf(fabsf(d_.x), fmaxf(fabsf(d_.y), fabsf(d_.z)));
			if (!(mo < 1e3f) || !(md < 1e3f) || !(md > 1e-6f))
				printf("odd ray: o %g %g %g d %g %g %g tmax %g any %d\n", o_.x, o_.y, o_.z, d_.x, d_.y, d_.z, tmax_, (int) ANY);
		}
#endif
		const float chk = ((o_.x + o_.y) + o_.z) + ((d_.x + d_.y) + d_.z);
		if (!(fabsf(chk) < 3.0e38f) || (d_.x == 0.f && d_.y == 0.f && d_.z == 0.f)) cur = kEmptyEntry, curInst = -1;
	}
	KRR_DEV void push(TraceSmem &sm, LStack &ls, uint32_t e, float tn) {
		if (sp < kShortStack) {
			sm.id[sp][threadIdx.x] = e;
			if (!ANY) sm.tn[sp][threadIdx.x] = tn;
		} else if (sp < kStackSize) {
			ls.id[sp - kShortStack] = e;
			if (!ANY) ls.tn[sp - kShortStack] = tn;
		} else { overflow = 1; return; }
		sp++;
	}
	KRR_DEV uint32_t pop(TraceSmem &sm, LStack &ls, float &tn) {
		--sp;
		if (sp < kShortStack) {
			if (!ANY) tn = sm.tn[sp][threadIdx.x];
			return sm.id[sp][threadIdx.x];
		}
		if (!ANY) tn = ls.tn[sp - kShortStack];
		return ls.id[sp - kShortStack];
	}

	// What the lane has to do next: 0 = ray finished (result in `best`), 1 = enter an instance,
	// 2 = wide node, 3 = leaf.  Pops the stack when the current entry is consumed.
	enum { FINISHED = 0, ENTER = 1, NODE = 2, LEAF = 3 };
	KRR_DEV int next(TraceSmem &sm, LStack &ls) {
		while (cur == kEmptyEntry) {
			if (curInst >= 0 && sp == blasBase) { // BLAS finished: back to world space
				curInst = -1;
				ro = o, rd = d;
				setIdir();
			}
			if (sp == 0) return FINISHED;
			float tn = 0.f;
			cur = pop(sm, ls, tn);
			// box entry beyond the closest hit so far (same slack as the slab test: for flat, axis-aligned
			// geometry the rounded entry distance can exceed the exact hit distance by an ulp, and an
			// equal-t candidate with a smaller (instance, primitive) must still be tested)
			if (!ANY && tn > best.t * 1.0000010f + 1e-30f) cur = kEmptyEntry;
		}
		return cur < kLeafFlag ? NODE : ((cur >> 30) == 2u ? ENTER : LEAF);
	}
	// ---- phase 1: enter an instance (object-space ray; t stays the world parameter) ----
	KRR_DEV void enterInstance(const BvhDev &bvh, const InstRec *__restrict__ instances, TraceSmem &sm) {
		curInst = (int) (cur & 0x3fffffffu);
		const InstRec &in = instances[curInst];
		if (MOTION && in.motion >= 0) movingRay(bvh, in.motion, time, o, d, ro, rd); // SRT motion chain at the ray's time
		else ro = xfPointX(in.inv, o), rd = xfVectorX(in.inv, d);
		setIdir();
		blasBase = sp;
		cur		 = (uint32_t) in.blasRoot;
	}
	// ---- phase 2: wide node: 8 slab tests, sort by entry distance, push far-to-near ----
	KRR_DEV void node(const BvhDev &bvh, TraceSmem &sm, LStack &ls) {
		{
			const float4 *np = reinterpret_cast<const float4 *>(bvh.nodes + cur);
			float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
			uint32_t ew = __float_as_uint(n0.w);
			float sx = __uint_as_float((ew & 0xff) << 23), sy = __uint_as_float(((ew >> 8) & 0xff) << 23),
				  sz = __uint_as_float(((ew >> 16) & 0xff) << 23);
			const uint32_t imask = ew >> 24;
			const uint32_t childBase = __float_as_uint(n1.x), primBase = __float_as_uint(n1.y);
			const uint32_t metaLo = __float_as_uint(n1.z), metaHi = __float_as_uint(n1.w);
			// quantised planes: n2 = qlo[0][0..7], qlo[1][0..7]; n3 = qlo[2], qhi[0]; n4 = qhi[1], qhi[2]
			const uint32_t q[12] = {__float_as_uint(n2.x), __float_as_uint(n2.y), __float_as_uint(n2.z), __float_as_uint(n2.w),
									__float_as_uint(n3.x), __float_as_uint(n3.y), __float_as_uint(n3.z), __float_as_uint(n3.w),
									__float_as_uint(n4.x), __float_as_uint(n4.y), __float_as_uint(n4.z), __float_as_uint(n4.w)};
			// t = (o_node + q*s - o) * idir = q * (s*idir) + (o_node - o)*idir
			const float ax = sx * idir.x, ay = sy * idir.y, az = sz * idir.z;
			const float bx = (n0.x - ro.x) * idir.x, by = (n0.y - ro.y) * idir.y, bz = (n0.z - ro.z) * idir.z;
			const float lim = best.t;
			// The byte -> float conversions (48 per node, quarter-rate I2F) are replaced by a byte permute that
			// drops q into the mantissa of 2^23: v = 2^23 + 256 q, and q*a + b = v*(a/256) + (b - 2^15 a) in one
			// FMA.  b - 2^15 a is rounded once: error <= |a| / 512, i.e. 1/512 of a quantisation step, against
			// the full step the builder pads every child plane with (quantize() in bvh_build.cu).
			const float axs = ax * 0.00390625f, ays = ay * 0.00390625f, azs = az * 0.00390625f;
			// Rounding-error bound of the decomposed slab form q*a + b (cancellation between two large terms
			// when the ray grazes an axis-aligned plane), PER AXIS: each slab interval is widened by its own
			// bound, folded into the FMA constant of its near / far plane.  (One scalar bound for all axes let a
			// ray that is nearly parallel to one axis -- huge a, b on that axis -- lose culling on the other two:
			// the pixels of one image column / row took 13x longer than the rest of the frame together.)
			// An axis the ray is exactly parallel to has a = inf and yields NaN distances, which min/max drop.
			auto axisSlack = [](float a, float b) { return fabsf(a) < 3.0e38f ? (fabsf(b) + 255.f * fabsf(a)) * 2.4e-7f + fabsf(a) * 0.00390625f : 0.f; };
			const float slx = axisSlack(ax, bx), sly = axisSlack(ay, by), slz = axisSlack(az, bz);
			const float bx0 = fmaf(-32768.f, ax, bx), by0 = fmaf(-32768.f, ay, by), bz0 = fmaf(-32768.f, az, bz);
			const float bxN = bx0 - slx, bxF = bx0 + slx, byN = by0 - sly, byF = by0 + sly, bzN = bz0 - slz, bzF = bz0 + slz;
			// ray octant: which plane of a slab is entered first only depends on the sign of the direction, so
			// the near / far plane words are selected once per node instead of a min and a max per child
			const bool px = ax >= 0.f, py = ay >= 0.f, pz = az >= 0.f;
			const uint32_t nX[2] = {px ? q[0] : q[6], px ? q[1] : q[7]}, fX[2] = {px ? q[6] : q[0], px ? q[7] : q[1]};
			const uint32_t nY[2] = {py ? q[2] : q[8], py ? q[3] : q[9]}, fY[2] = {py ? q[8] : q[2], py ? q[9] : q[3]};
			const uint32_t nZ[2] = {pz ? q[4] : q[10], pz ? q[5] : q[11]}, fZ[2] = {pz ? q[10] : q[4], pz ? q[11] : q[5]};
			uint32_t k0, k1, k2, k3, k4, k5, k6, k7;
			auto child = [&](int i) -> uint32_t {
				uint32_t meta = ((i < 4 ? metaLo : metaHi) >> ((i & 3) * 8)) & 0xff;
				bool internal = (imask >> i) & 1;
				if (!internal && meta == 0) return kEmptyEntry;
				auto qf = [&](uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4b000000u, 0x7404u | ((i & 3) << 4))); };
				const float tnx = fmaf(qf(nX[i >> 2]), axs, bxN), tfx = fmaf(qf(fX[i >> 2]), axs, bxF);
				const float tny = fmaf(qf(nY[i >> 2]), ays, byN), tfy = fmaf(qf(fY[i >> 2]), ays, byF);
				const float tnz = fmaf(qf(nZ[i >> 2]), azs, bzN), tfz = fmaf(qf(fZ[i >> 2]), azs, bzF);
				// fminf/fmaxf drop NaNs (0 * inf), which is the conservative answer for a degenerate slab
				const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.f));
				const float tf = fminf(fminf(tfx, tfy), fminf(tfz, lim));
				// conservative: boxes only cull, the exact decision is the triangle test
				if (!(tn <= tf * 1.0000010f)) return kEmptyEntry;
				// key: entry distance in the high bits, slot in the low 3 -> one compare orders (tn, slot)
				return (__float_as_uint(tn) & ~7u) | (uint32_t) i;
			};
			k0 = child(0), k1 = child(1), k2 = child(2), k3 = child(3), k4 = child(4), k5 = child(5), k6 = child(6), k7 = child(7);
			// 19-comparator sorting network: ascending, misses (0xffffffff) sink to the end
			KRR_CSWAP(k0, k1) KRR_CSWAP(k2, k3) KRR_CSWAP(k4, k5) KRR_CSWAP(k6, k7)
			KRR_CSWAP(k0, k2) KRR_CSWAP(k1, k3) KRR_CSWAP(k4, k6) KRR_CSWAP(k5, k7)
			KRR_CSWAP(k1, k2) KRR_CSWAP(k5, k6) KRR_CSWAP(k0, k4) KRR_CSWAP(k3, k7)
			KRR_CSWAP(k1, k5) KRR_CSWAP(k2, k6)
			KRR_CSWAP(k1, k4) KRR_CSWAP(k3, k6)
			KRR_CSWAP(k2, k4) KRR_CSWAP(k3, k5)
			KRR_CSWAP(k3, k4)
			const bool tlas = curInst < 0;
			auto entryOf = [&](uint32_t key) -> uint32_t {
				uint32_t slot = key & 7u;
				if ((imask >> slot) & 1) return childBase + __popc(imask & ((1u << slot) - 1));
				uint32_t meta = ((slot < 4 ? metaLo : metaHi) >> ((slot & 3) * 8)) & 0xff;
				uint32_t first = primBase + (meta & 31);
				if (tlas) return kInstFlag | (uint32_t) __ldg(bvh.tlasInst + first); // one instance per TLAS leaf
				return kLeafFlag | (((meta >> 5) - 1) << 26) | first;
			};
			// far-to-near onto the stack, nearest continues
			if (k7 != kEmptyEntry) push(sm, ls, entryOf(k7), __uint_as_float(k7 & ~7u));
			if (k6 != kEmptyEntry) push(sm, ls, entryOf(k6), __uint_as_float(k6 & ~7u));
			if (k5 != kEmptyEntry) push(sm, ls, entryOf(k5), __uint_as_float(k5 & ~7u));
			if (k4 != kEmptyEntry) push(sm, ls, entryOf(k4), __uint_as_float(k4 & ~7u));
			if (k3 != kEmptyEntry) push(sm, ls, entryOf(k3), __uint_as_float(k3 & ~7u));
			if (k2 != kEmptyEntry) push(sm, ls, entryOf(k2), __uint_as_float(k2 & ~7u));
			if (k1 != kEmptyEntry) push(sm, ls, entryOf(k1), __uint_as_float(k1 & ~7u));
			cur = k0 != kEmptyEntry ? entryOf(k0) : kEmptyEntry;
		}
	}
	// ---- phase 3: leaf (1..7 triangles).  Returns true when an any-hit ray terminated. ----
	template <typename Accept> KRR_DEV bool leaf(const BvhDev &bvh, Accept accept) {
		uint32_t first = cur & 0x03ffffffu, cnt = ((cur >> 26) & 7u) + 1;
		if ((cur >> 30) == 3u) { // flat BLAS: the whole triangle list
			const int2 fr = __ldg(bvh.flats + (cur & 0x3fffffffu));
			first = (uint32_t) fr.x, cnt = (uint32_t) fr.y;
		}
		cur = kEmptyEntry;
		uint32_t k = 0;
#if KRR_LEAF_PAIR
		if constexpr (PAIR || KRR_LEAF_PAIR_TREE)
		// KRR_LEAF_WIDTH triangles per trip, evaluated without branches (triTestNoBranch: the same operations
		// and roundings as triIntersectE, combined as predicates).  The early exits of triIntersectE save the
		// WARP little (32 rays per triangle: some lane usually goes on), and their branches serialise the
		// dependent multiply-add chains that independent triangles interleave.
		// Only in the kernels instantiated for flat-list scenes: in a tree leaf the lanes hold different
		// triangles, and the mere presence of this loop cost the tree kernels registers (config 5: -10 %).
		for (; k + KRR_LEAF_WIDTH <= cnt; k += KRR_LEAF_WIDTH) {
			const float4 *tp = reinterpret_cast<const float4 *>(bvh.tris + first + k);
			float4 a[KRR_LEAF_WIDTH], b[KRR_LEAF_WIDTH], c[KRR_LEAF_WIDTH];
			float t[KRR_LEAF_WIDTH], u[KRR_LEAF_WIDTH], v[KRR_LEAF_WIDTH];
			bool hit[KRR_LEAF_WIDTH];
#pragma unroll
			for (int j = 0; j < KRR_LEAF_WIDTH; j++) a[j] = __ldg(tp + 3 * j), b[j] = __ldg(tp + 3 * j + 1), c[j] = __ldg(tp + 3 * j + 2);
#pragma unroll
			for (int j = 0; j < KRR_LEAF_WIDTH; j++) hit[j] = triTestNoBranch(ro, rd, mk3(a[j]), mk3(b[j]), mk3(c[j]), tmax, t[j], u[j], v[j]);
#pragma unroll
			for (int j = 0; j < KRR_LEAF_WIDTH; j++) {
				if (hit[j]) {
					const int prim = __float_as_int(a[j].w);
					const int inst = curInst == bvh.mergedInst ? __float_as_int(b[j].w) : curInst;
					if (betterHit(t[j], inst, prim, best) && accept(inst, prim, u[j], v[j])) {
						best.inst = inst, best.prim = prim, best.t = t[j], best.u = u[j], best.v = v[j];
						if (ANY) return true;
					}
				}
			}
		}
#endif
		for (; k < cnt; k++) {
			const float4 *tp = reinterpret_cast<const float4 *>(bvh.tris + first + k);
			float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
			float t, u, v;
			if (triIntersectE(ro, rd, mk3(a), mk3(b), mk3(c), tmax, t, u, v)) {
				const int prim = __float_as_int(a.w);
				const int inst = curInst == bvh.mergedInst ? __float_as_int(b.w) : curInst;
				if (betterHit(t, inst, prim, best) && accept(inst, prim, u, v)) {
					best.inst = inst, best.prim = prim, best.t = t, best.u = u, best.v = v;
					if (ANY) return true;
				}
			}
		}
		return false;
	}

	// Lane-local traversal to the end (no warp-level primitives): used where one lane runs several
	// dependent traversals interleaved with other work (ratio-tracking shadow rays through media).
	template <typename Accept> KRR_DEV void runToEnd(const BvhDev &bvh, const InstRec *__restrict__ instances, TraceSmem &sm, LStack &ls, Accept accept) {
		while (true) {
			int st = next(sm, ls);
			if (st == FINISHED) return;
			if (st == ENTER) { enterInstance(bvh, instances, sm); st = (cur >> 30) == 3u ? LEAF : NODE; }
			if (st == NODE) {
				node(bvh, sm, ls);
				if (cur != kEmptyEntry && (cur >> 30) == 1u) st = LEAF;
			}
			if (st == LEAF && leaf(bvh, accept)) return;
		}
	}

	// One warp-cooperative trip: lanes vote on the phase to run, so that a phase executes with as many
	// lanes as possible; lanes whose phase lost the vote keep their entry and wait (they would have been
	// masked off anyway).  Returns true for lanes whose ray finished during this trip.
	// VOTE = false runs both phases every trip (a lane may test a node and then its nearest leaf in
	// the same trip): fewer trips per ray, which wins for the short any-hit traversals of shadow rays.
	template <bool VOTE, typename Accept>
	KRR_DEV bool trip(bool active, const BvhDev &bvh, const InstRec *__restrict__ instances, TraceSmem &sm, LStack &ls, Accept accept) {
		const unsigned FULL = 0xffffffffu;
		int st = FINISHED;
		bool fin = false;
		if (active) {
			st	= next(sm, ls);
			fin = st == FINISHED;
		}
		if (__any_sync(FULL, st == ENTER)) {
			if (st == ENTER) { enterInstance(bvh, instances, sm); st = (cur >> 30) == 3u ? LEAF : NODE; }
		}
		if (VOTE) {
			const unsigned mN = __ballot_sync(FULL, st == NODE), mL = __ballot_sync(FULL, st == LEAF);
			if (mN && __popc(mN) >= __popc(mL)) {
				if (st == NODE) node(bvh, sm, ls);
			} else if (mL) {
				if (st == LEAF) fin = leaf(bvh, accept);
			}
		} else {
			if (st == NODE) {
				node(bvh, sm, ls);
				if (cur != kEmptyEntry && (cur >> 30) == 1u) st = LEAF;
			}
			if (st == LEAF) fin = leaf(bvh, accept);
		}
		return fin;
	}
};

#endif // __CUDACC__

} // namespace krr
