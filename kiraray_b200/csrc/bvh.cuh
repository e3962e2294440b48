// bvh.cuh -- compressed 8-wide BVH (two-level: TLAS over instances, one BLAS per mesh) and the
// traversal routines of the Closest / Shadow stages.
//
// The reference delegates this to closed NVIDIA OptiX running on RT cores (optixTrace at
// src/render/wavefront/device.cu:13-22; accel build src/core/device/optix.cpp:143-250, 357-398).
// B200 has no RT cores, so traversal is SM code:
//   * 80-byte nodes: an anchor point + per-axis power-of-two scale, and 8 children whose boxes are
//     quantised to 8 bits per plane (conservatively: lo rounded down, hi rounded up).  One node is
//     five 16-byte loads.
//   * children sit in OCTANT-ORDERED slots (the builder puts the child that lies towards (-x,-y,-z) of
//     the node centre into slot 0, ... towards (+x,+y,+z) into slot 7), so "slot XOR ray octant" is a
//     front-to-back priority: a node test produces ONE 32-bit hit mask (internal children in priority
//     order in the top byte, leaf triangles in the low 24 bits) instead of eight sorted stack entries.
//     A stack entry is a node GROUP (child base + hit mask) or a triangle group (first triangle + hit
//     bits): one 8-byte push per visited node at most (after Ylitie, Karras, Laine: "Efficient
//     incoherent ray traversal on GPUs through compressed wide BVHs", HPG 2017).
//   * BLAS nodes live in OBJECT space of their mesh; an instance is entered by transforming the ray
//     with the instance's inverse 3x4 (t stays the world-space parameter, the direction is not
//     renormalised), exactly as the intersection spec in oracle/driver.cpp states.
//   * ray/triangle: Moeller-Trumbore with individually rounded operations (xmul/xadd..., never
//     FMA-contracted), accept 0 < t < tmax, ties on t broken by (instance, primitive) so that the
//     result does not depend on traversal order -> first-hit ids are bit-exact against the oracle's
//     brute-force loop.
//   * warp-cooperative traversal (see Traverser below): phase-aligned stepping, shared-memory short
//     stack, persistent warps that refill finished lanes from the queue.
#pragma once
#include "krr_math.cuh"
#include "scene.cuh"

namespace krr {

struct __align__(16) Node8 {
	float ox, oy, oz;		  // anchor (min corner of the node box)
	uint8_t ex, ey, ez;		  // biased exponents: child plane = o + q * 2^(e-127)
	uint8_t imask;			  // bit i: the child in slot i is an internal node
	uint32_t childBase;		  // first internal child; slot i -> childBase + popc(imask & ((1<<i)-1))
	uint32_t primBase;		  // first primitive of the leaf children (BLAS: triangle pool; TLAS: 8 instance ids, slot-major)
	// per slot: 0 = empty; internal child: 0x20 | (24 + slot); BLAS leaf: unary triangle count (1, 3, 7) << 5 |
	// offset of its first triangle from primBase (0..21); TLAS leaf (one instance): 0x20 | slot.  The node
	// test turns a hit slot into `(meta >> 5) << (meta & 31)` -- bits 24..31 internal children, bits 0..23
	// triangles / instances -- after XOR-ing the slot number with the ray octant where the bit position is a
	// traversal priority (internal children, TLAS instances).
	uint8_t meta[8];
	uint8_t qlo[3][8], qhi[3][8];
};
static_assert(sizeof(Node8) == 80, "Node8 must be 80 bytes");

// v0.xyz = first vertex, e1 = v1 - v0, e2 = v2 - v0 (the first two operations of the triangle test, done once
// at build time with the same rounding); v0.w = primitive id, e1.w = instance id (merged BLAS only)
struct BvhTri { float4 v0, e1, e2; };

struct BvhDev {
	const Node8 *nodes;	   // node pool: TLAS nodes first, then every mesh's BLAS
	const BvhTri *tris;	   // triangle pool, leaf order per mesh
	const int32_t *tlasInst; // instance ids referenced by TLAS leaves: 8 per TLAS node, indexed by child slot
	// per instance: world-space bounding sphere (centre, radius) over the ray-time window, radius >= 1e30 = none.
	// Tested before an instance is entered: a rotated object's box is much larger than the object, and entering a
	// moving instance costs an SRT-chain evaluation
	const float4 *instSphere;
	int32_t tlasRoot;
	int32_t nInstances;
	const XformNodeRec *xnodes; // motion blur: transform chains + SRT key pool (motion.cuh), null otherwise
	const float *motionKeys;
	const float4 *motionFlat; // per instance: flat motion record (motion.cuh), null = none
	// Static instances whose transform is exactly the identity are MERGED into one world-space BLAS
	// (bvh_build.cu): object space == world space for them, so the intersection spec gives the same
	// numbers, and a ray no longer enters each of them separately.  The merged BLAS hangs in the TLAS as
	// pseudo-instance `mergedInst` (= nInstances, identity transform); its triangles carry their real
	// instance id.  mergedOnly: every instance was merged, traversal starts at the merged root.
	int32_t mergedInst; // -1 = nothing merged
	int32_t mergedRoot;
	int32_t mergedOnly;
	// FLATTENED static instances: a static instance whose mesh no other instance uses is merged as well, whatever
	// its transform.  The merged BLAS then bounds the TRANSFORMED triangles (world space, padded for the rounding
	// of the inverse), but the triangles themselves stay in the object space of their instance and the leaf test
	// transforms the ray with that instance's inverse first -- the very operations of the intersection spec, so
	// hits are bit-identical to entering the instance through the TLAS.  What it buys: no per-instance entry, and
	// a ray that passes between objects walks ONE tree that has carved the empty space out instead of every
	// object box on its way.  mergedXf: some merged instance has a non-identity transform.
	int32_t mergedXf;
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
#ifndef KRR_LEAF_WIDTH
#define KRR_LEAF_WIDTH 2
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

constexpr uint32_t kFlatFlag = 0xc0000000u, kEmptyEntry = 0xffffffffu;
KRR_HD bool isFlatEntry(uint32_t e) { return (e >> 30) == 3u && e != kEmptyEntry; }

#ifdef __CUDACC__
// ---- traversal state machine ---------------------------------------------------------------------
// One lane = one ray.  Per-lane state is two GROUPS (x = base index, y = bit mask):
//   ng: children of the last tested node that the ray hit and that are internal nodes.  y bits 24..31 = hit
//       children in traversal priority (bit 24 + (slot ^ octant)), y bits 0..7 = the node's imask (needed to turn
//       a slot into a child index); x = the node's childBase.
//   tg: hit leaf primitives.  In a BLAS: x = the node's primBase, y bits 0..23 = triangles primBase + bit.  In
//       the TLAS: y bit (slot ^ octant) = the instance tlasInst[primBase + slot].  A flat BLAS (triangle list)
//       is the special group y = 0x80000000 | count, x = first triangle; it is consumed at once, never pushed.
// The traversal is written as a STEP function (trip) so that the stage kernels can run it warp-cooperatively:
// every trip executes the phases
//     [pop] -> [enter instance] -> [wide node: 8 slab tests -> hit mask] -> [triangle tests]
// under per-lane predicates, so lanes that are in the same phase execute it together, and the kernel refills
// lanes whose ray has terminated from the queue (persistent warps, one atomicAdd per refill) instead of
// letting them idle until the slowest ray of the warp is done.
//
// Stack: 8-byte entries (a node group whose top byte is non-zero, or a triangle / instance group whose top
// byte is zero).  The first kShortStack entries of every lane live in SHARED memory (slot-major, so a warp's
// accesses are conflict-free); deeper entries spill to local memory.  A node pushes at most ONE entry (its
// remaining hit children), so the depth is the tree depth, not 7x the tree depth.
#ifndef KRR_SHORT_STACK
#define KRR_SHORT_STACK 8
#endif
constexpr int kShortStack  = KRR_SHORT_STACK;
constexpr int kLocalStack  = 64 - KRR_SHORT_STACK; // one entry per BLAS level, two per TLAS level
constexpr int kStackSize   = kShortStack + kLocalStack;
constexpr int kTraceBlock  = 128;

struct TraceSmem {
	uint2 st[kShortStack][kTraceBlock];
};
// Spill part of the stack (local memory).  Deliberately NOT a member of Traverser: a dynamically
// indexed array inside the struct keeps the WHOLE struct in local memory (the compiler cannot split an
// aggregate that is indexed with a run-time value), and the ray state would be loaded and stored
// around every phase instead of living in registers.
template <bool ANY> struct LocalStack {
	uint2 st[kLocalStack];
};

// world ray -> object space of a moving instance (kept out of line: static scenes never pay its registers)
static __device__ __noinline__ void movingRay(const BvhDev &bvh, int inst, int node, float time, V3 o, V3 d, V3 &ro, V3 &rd) {
	Xf m, inv;
	movingInstanceXf(bvh.xnodes, bvh.motionKeys, bvh.motionFlat, inst, node, time, m, inv);
	ro = xfPointX(inv, o), rd = xfVectorX(inv, d);
}

// triangle tests a lane runs per trip when it holds a triangle group (the rest waits for the next trip)
#ifndef KRR_TRI_PER_TRIP
#define KRR_TRI_PER_TRIP 2
#endif
// triangle phase of a voted trip runs when at least this many lanes hold triangles, or no lane has node work
#ifndef KRR_TRI_VOTE
#define KRR_TRI_VOTE 6
#endif
// Entering a MOVING instance evaluates its SRT chain at the ray's time (~450 instructions against ~170 for a node
// test): with incoherent rays a few lanes want it on almost every trip, and the whole warp pays for it.  The
// phase therefore waits until this many lanes want to enter (the waiting lanes are blocked, so they gather within
// a few trips), or no lane has any other work.  Static instances (a 30-instruction transform) enter at once.
#ifndef KRR_ENTER_VOTE
#define KRR_ENTER_VOTE 8
#endif

// MOTION = false compiles the SRT-chain path out (static scenes keep their register budget)
// PAIR: the scene is one flat triangle list and the whole traversal is a walk over it in branch-free pairs
template <bool ANY, bool MOTION = true, bool PAIR = false> struct Traverser {
	// ray
	V3 o, d;	  // world space
	V3 ro, rd;	  // current space (world in the TLAS, object space inside a BLAS)
	V3 idir;
	float tmax, time;
	Hit best;
	// control
	uint2 ng, tg;
	V3 oo, od;	 // flattened instances (BvhDev::mergedXf): the ray in the object space of instance objInst
	int objInst;
	uint32_t octinv4; // (x >= 0 ? 4 : 0) | (y >= 0 ? 2 : 0) | (z >= 0 ? 1 : 0) of the current-space direction, in every byte
	int sp, curInst, blasBase;
	int overflow;
#ifdef KRR_COUNT_TRIPS
	int nodeSteps, triTests, enters, culled;
#endif
	using LStack = LocalStack<ANY>;

	// reciprocal direction of the slab tests.  A zero component is replaced by +-1e-20: the distances to the two
	// planes of that slab become -+huge (origin inside the slab: the interval covers everything) or huge with
	// one sign (outside: the box is culled), all finite.  With 1/0 = inf they were NaN, the axis was dropped from
	// the test, and a ray parallel to two axes could only be culled along its own direction: it walked every
	// node in front of it (46 ms for one such ray in the 20 M-triangle scene).
	KRR_DEV void setSpace() {
		auto safe = [](float x) { return fabsf(x) >= 1e-20f ? x : copysignf(1e-20f, x); };
		idir	= mk3(1.f / safe(rd.x), 1.f / safe(rd.y), 1.f / safe(rd.z));
		octinv4 = ((idir.x >= 0.f ? 4u : 0u) | (idir.y >= 0.f ? 2u : 0u) | (idir.z >= 0.f ? 1u : 0u)) * 0x01010101u;
	}
	// start at a BLAS root: a tree (node group with the single pseudo-child `root`) or a flat triangle list
	KRR_DEV void enterRoot(const BvhDev &bvh, uint32_t root) {
		if (isFlatEntry(root)) {
			const int2 fr = __ldg(bvh.flats + (root & 0x3fffffffu));
			tg = make_uint2((uint32_t) fr.x, 0x80000000u | (uint32_t) fr.y), ng = make_uint2(0u, 0u);
		} else ng = make_uint2(root, 0x80000000u), tg = make_uint2(0u, 0u); // imask 0: child index = root + 0
	}
	KRR_DEV void begin(const BvhDev &bvh, V3 o_, V3 d_, float tmax_, float time_ = 0.f) {
		o = ro = o_, d = rd = d_, tmax = tmax_, time = time_;
		setSpace();
		best.inst = -1, best.prim = -1, best.t = tmax_, best.u = best.v = 0;
		sp = 0, curInst = -1, blasBase = -1, overflow = 0, objInst = -1;
#ifdef KRR_COUNT_TRIPS
		nodeSteps = triTests = enters = culled = 0;
#endif
		ng = make_uint2((uint32_t) bvh.tlasRoot, 0x80000000u), tg = make_uint2(0u, 0u);
		if (PAIR || bvh.mergedOnly) {
			curInst = bvh.mergedInst, blasBase = 0; // world == object space
			enterRoot(bvh, (uint32_t) bvh.mergedRoot);
			const float t0x = (bvh.rootLo[0] - o_.x) * idir.x, t1x = (bvh.rootHi[0] - o_.x) * idir.x;
			const float t0y = (bvh.rootLo[1] - o_.y) * idir.y, t1y = (bvh.rootHi[1] - o_.y) * idir.y;
			const float t0z = (bvh.rootLo[2] - o_.z) * idir.z, t1z = (bvh.rootHi[2] - o_.z) * idir.z;
			const float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), 0.f));
			const float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tmax_));
			if (!(tn <= tf * 1.00001f + 1e-30f)) ng.y = tg.y = 0u; // NaNs (0 * inf) are dropped by min / max
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
		if (!(fabsf(chk) < 3.0e38f) || (d_.x == 0.f && d_.y == 0.f && d_.z == 0.f)) ng.y = tg.y = 0u;
	}
	KRR_DEV void push(TraceSmem &sm, LStack &ls, uint2 e) {
		if (sp < kShortStack) sm.st[sp][threadIdx.x] = e;
		else if (sp < kStackSize) ls.st[sp - kShortStack] = e;
		else { overflow = 1; return; }
		sp++;
	}
	KRR_DEV uint2 pop(TraceSmem &sm, LStack &ls) {
		--sp;
		return sp < kShortStack ? sm.st[sp][threadIdx.x] : ls.st[sp - kShortStack];
	}
	KRR_DEV bool hasNode() const { return (ng.y & 0xff000000u) != 0u; }

	// The lane has used up both groups: leave the BLAS when its part of the stack is empty, then pop.
	// Returns false when the traversal is finished (result in `best`).
	KRR_DEV bool popNext(TraceSmem &sm, LStack &ls) {
		if (curInst >= 0 && sp == blasBase) { // BLAS finished: back to world space
			curInst = -1;
			ro = o, rd = d;
			setSpace();
		}
		if (sp == 0) return false;
		const uint2 e = pop(sm, ls);
		if (e.y & 0xff000000u) ng = e, tg = make_uint2(0u, 0u);
		else tg = e, ng = make_uint2(0u, 0u);
		return true;
	}
	// ---- phase: enter the nearest hit instance of the TLAS group (object-space ray; t stays the world parameter) ----
	// Two steps.  PROBE: take the nearest instance bit of the group and test the ray against the instance's bounding
	// sphere (a rotated object's box is much larger than the object); culled instances only cost this.  ENTER: push the
	// TLAS-level groups, transform the ray, start at the BLAS root.  Static kernels run both at once (enterInstance).
	// In the MOTION kernels entering a moving instance evaluates its SRT chain at the ray's time (~450 instructions),
	// so the probe PARKS the lane (curInst = -2 - instance) and trip() runs the entry of all parked lanes under a
	// vote: the chain then executes with the lanes that passed the sphere test, not with the ones that merely wanted
	// to look (ncu, 10 000 moving instances: the chain ran with 5 of 32 lanes when the vote counted probing lanes).
	KRR_DEV int probeInstance(const BvhDev &bvh) { // returns the instance to enter, -1 when this bit was culled
		const uint32_t bit = 31u - (uint32_t) __clz(tg.y);
		tg.y ^= 1u << bit;
		const uint32_t slot = bit ^ (octinv4 & 7u);
		const int inst		= __ldg(bvh.tlasInst + tg.x + slot);
		// conservative ray / bounding-sphere test (the bit is consumed either way)
		const float4 sph = __ldg(bvh.instSphere + inst);
		if (sph.w < 1.0e30f) {
			const V3 l	   = mk3(sph.x - o.x, sph.y - o.y, sph.z - o.z);
			const float ts = fminf(fmaxf(dot(l, d) / dot(d, d), 0.f), best.t); // closest approach within [0, best.t]
			const V3 q	   = l - d * ts;
			const float rr = sph.w * 1.0005f + 1e-5f * (fabsf(sph.x) + fabsf(sph.y) + fabsf(sph.z) + 1.f);
			if (dot(q, q) > rr * rr) {
#ifdef KRR_COUNT_TRIPS
				culled++;
#endif
				return -1;
			}
		}
		return inst;
	}
	KRR_DEV void enterProbed(const BvhDev &bvh, const InstRec *__restrict__ instances, TraceSmem &sm, LStack &ls, int inst) {
#ifdef KRR_COUNT_TRIPS
		enters++;
#endif
		// the TLAS-level groups wait on the stack below the BLAS part
		if (hasNode()) push(sm, ls, ng);
		if (tg.y) push(sm, ls, tg);
		curInst = inst;
		const InstRec &in = instances[inst];
		if (MOTION && in.motion >= 0) movingRay(bvh, inst, in.motion, time, o, d, ro, rd); // SRT motion chain at the ray's time
		else ro = xfPointX(in.inv, o), rd = xfVectorX(in.inv, d);
		setSpace();
		blasBase = sp;
		enterRoot(bvh, (uint32_t) in.blasRoot);
	}
	KRR_DEV void enterInstance(const BvhDev &bvh, const InstRec *__restrict__ instances, TraceSmem &sm, LStack &ls) {
		const int inst = probeInstance(bvh);
		if (inst >= 0) enterProbed(bvh, instances, sm, ls, inst);
	}
	// MOTION kernels: probe the group's instances until one passes; a static one is entered at once, a moving one parks
	// the lane for the voted entry
	KRR_DEV bool parked() const { return MOTION && curInst <= -2; }
	KRR_DEV void probeAndPark(const BvhDev &bvh, const InstRec *__restrict__ instances, TraceSmem &sm, LStack &ls) {
		while (tg.y) {
			const int inst = probeInstance(bvh);
			if (inst < 0) continue;
			if (instances[inst].motion < 0) enterProbed(bvh, instances, sm, ls, inst);
			else curInst = -2 - inst;
			return;
		}
	}
	// ---- phase: wide node.  Takes the nearest hit child of the node group, tests its 8 children ----
	KRR_DEV void nodeStep(const BvhDev &bvh, TraceSmem &sm, LStack &ls) {
#ifdef KRR_COUNT_TRIPS
		nodeSteps++;
#endif
		const uint32_t hits = ng.y;
		const uint32_t bit	= 31u - (uint32_t) __clz(hits);
		ng.y ^= 1u << bit;
		if (ng.y & 0xff000000u) push(sm, ls, ng);
		const uint32_t slot = (bit - 24u) ^ (octinv4 & 7u);
		const uint32_t node = ng.x + (uint32_t) __popc(hits & ~(0xffffffffu << slot));
		const float4 *np = reinterpret_cast<const float4 *>(bvh.nodes + node);
		const float4 n0 = __ldg(np), n1 = __ldg(np + 1), n2 = __ldg(np + 2), n3 = __ldg(np + 3), n4 = __ldg(np + 4);
		const uint32_t ew = __float_as_uint(n0.w);
		const float sx = __uint_as_float((ew & 0xff) << 23), sy = __uint_as_float(((ew >> 8) & 0xff) << 23),
					sz = __uint_as_float(((ew >> 16) & 0xff) << 23);
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
		auto axisSlack = [](float a, float b) { return fabsf(a) < 3.0e38f ? (fabsf(b) + 255.f * fabsf(a)) * 2.4e-7f + fabsf(a) * 0.00390625f : 0.f; };
		const float slx = axisSlack(ax, bx), sly = axisSlack(ay, by), slz = axisSlack(az, bz);
		const float bx0 = fmaf(-32768.f, ax, bx), by0 = fmaf(-32768.f, ay, by), bz0 = fmaf(-32768.f, az, bz);
		const float bxN = bx0 - slx, bxF = bx0 + slx, byN = by0 - sly, byF = by0 + sly, bzN = bz0 - slz, bzF = bz0 + slz;
		// ray octant: which plane of a slab is entered first only depends on the sign of the direction, so
		// the near / far plane words are selected once per node instead of a min and a max per child.
		// quantised planes: n2 = qlo[0][0..7], qlo[1][0..7]; n3 = qlo[2], qhi[0]; n4 = qhi[1], qhi[2]
		const bool px = ax >= 0.f, py = ay >= 0.f, pz = az >= 0.f;
		const uint32_t lx0 = __float_as_uint(n2.x), lx1 = __float_as_uint(n2.y), ly0 = __float_as_uint(n2.z), ly1 = __float_as_uint(n2.w);
		const uint32_t lz0 = __float_as_uint(n3.x), lz1 = __float_as_uint(n3.y), hx0 = __float_as_uint(n3.z), hx1 = __float_as_uint(n3.w);
		const uint32_t hy0 = __float_as_uint(n4.x), hy1 = __float_as_uint(n4.y), hz0 = __float_as_uint(n4.z), hz1 = __float_as_uint(n4.w);
		const uint32_t nX[2] = {px ? lx0 : hx0, px ? lx1 : hx1}, fX[2] = {px ? hx0 : lx0, px ? hx1 : lx1};
		const uint32_t nY[2] = {py ? ly0 : hy0, py ? ly1 : hy1}, fY[2] = {py ? hy0 : ly0, py ? hy1 : ly1};
		const uint32_t nZ[2] = {pz ? lz0 : hz0, pz ? lz1 : hz1}, fZ[2] = {pz ? hz0 : lz0, pz ? hz1 : lz1};
		const uint32_t metaW[2] = {__float_as_uint(n1.z), __float_as_uint(n1.w)};
		const bool tlas = curInst < 0;
		uint32_t hitmask = 0u;
#pragma unroll
		for (int h = 0; h < 2; h++) {
			const uint32_t meta4 = metaW[h];
			// internal children (0x38 | slot: bits 3 and 4 both set) get their priority from the ray octant, and
			// so do the one-instance leaves of the TLAS; BLAS leaves keep their triangle offset
			const uint32_t inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
			// (inner4 >> 4) holds 0x01 in every internal byte: times the octant (< 8) = the octant in those bytes
			const uint32_t xor4 = tlas ? octinv4 : (inner4 >> 4) * (octinv4 & 7u);
			const uint32_t bit4 = (meta4 ^ xor4) & 0x1f1f1f1fu;
			const uint32_t cnt4	  = (meta4 >> 5) & 0x07070707u;
#pragma unroll
			for (int j = 0; j < 4; j++) {
				auto qf = [&](uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4b000000u, 0x7404u | (j << 4))); };
				const float tnx = fmaf(qf(nX[h]), axs, bxN), tfx = fmaf(qf(fX[h]), axs, bxF);
				const float tny = fmaf(qf(nY[h]), ays, byN), tfy = fmaf(qf(fY[h]), ays, byF);
				const float tnz = fmaf(qf(nZ[h]), azs, bzN), tfz = fmaf(qf(fZ[h]), azs, bzF);
				// fminf/fmaxf drop NaNs (0 * inf), which is the conservative answer for a degenerate slab
				const float tn = fmaxf(fmaxf(tnx, tny), fmaxf(tnz, 0.f));
				const float tf = fminf(fminf(tfx, tfy), fminf(tfz, lim));
				// conservative: boxes only cull, the exact decision is the triangle test.  The slack lets a
				// candidate at exactly the current best distance through (ties are decided by ids).
				if (tn <= tf * 1.0000010f) hitmask |= ((cnt4 >> (8 * j)) & 0xffu) << ((bit4 >> (8 * j)) & 0xffu);
			}
		}
		ng = make_uint2(__float_as_uint(n1.x), (hitmask & 0xff000000u) | (ew >> 24));
		tg = make_uint2(__float_as_uint(n1.y), hitmask & 0x00ffffffu);
#ifndef KRR_PREFETCH
#define KRR_PREFETCH 0 // measured: -7 % (20 M triangles) and -19 % (10 k moving instances): the extra instructions cost more than the lines save
#endif
#if KRR_PREFETCH
		// A lone ray is a chain of dependent misses (20 M triangles: 250 MB of nodes + 960 MB of triangles against
		// 126 MB of L2).  The hit children are contiguous (childBase + 0..7), so the lines of the SECOND and later
		// hit children and of the hit triangles are requested now and arrive while the first child is walked:
		// DRAM runs at a few per cent of its bandwidth here, latency is what the deep-bounce launches pay for.
		if (hitmask) {
			const uint32_t imask = ew >> 24;
			const uint32_t nInner = (uint32_t) __popc(hitmask >> 24);
			if (nInner > 1u) { // children childBase .. childBase + popc(imask) - 1: 80 B each, at most 5 lines of 128 B
				const char *cp = reinterpret_cast<const char *>(bvh.nodes + ng.x);
				const uint32_t bytes = (uint32_t) __popc(imask) * 80u;
				for (uint32_t off = 0; off < bytes; off += 128u) asm volatile("prefetch.global.L2 [%0];" ::"l"(cp + off));
			}
			if (tg.y && !tlas) {
				const char *tp = reinterpret_cast<const char *>(bvh.tris + tg.x + (uint32_t) (__ffs(tg.y) - 1));
				asm volatile("prefetch.global.L2 [%0];" ::"l"(tp));
			}
		}
#endif
	}
	template <typename Accept> KRR_DEV bool tryHit(const BvhDev &bvh, const float4 &a, const float4 &b, float t, float u, float v, Accept accept) {
		const int prim = __float_as_int(a.w);
		const int inst = curInst == bvh.mergedInst ? __float_as_int(b.w) : curInst;
		if (betterHit(t, inst, prim, best) && accept(inst, prim, u, v)) {
			best.inst = inst, best.prim = prim, best.t = t, best.u = u, best.v = v;
			return true;
		}
		return false;
	}
	// flat BLAS: the whole triangle list.  Returns true when an any-hit ray terminated.
	template <typename Accept> KRR_DEV bool walkFlat(const BvhDev &bvh, const InstRec *__restrict__ instances, Accept accept) {
		const uint32_t first = tg.x, cnt = tg.y & 0x00ffffffu;
		tg.y	   = 0u;
		uint32_t k = 0;
		if constexpr (PAIR)
		// KRR_LEAF_WIDTH triangles per iteration, evaluated without branches (triTestNoBranch: the same operations
		// and roundings as triIntersectE, combined as predicates).  The early exits of triIntersectE save the
		// WARP little (32 rays per triangle: some lane usually goes on), and their branches serialise the
		// dependent multiply-add chains that independent triangles interleave.
		// Only in the kernels instantiated for flat-list scenes: where the lanes hold different lists the mere
		// presence of this loop cost the tree kernels registers.
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
			for (int j = 0; j < KRR_LEAF_WIDTH; j++)
				if (hit[j] && tryHit(bvh, a[j], b[j], t[j], u[j], v[j], accept) && ANY) return true;
		}
		const bool flattened = !PAIR && bvh.mergedXf && curInst == bvh.mergedInst;
		for (; k < cnt; k++) {
			const float4 *tp = reinterpret_cast<const float4 *>(bvh.tris + first + k);
			const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
			V3 to = ro, td = rd;
			if (flattened) objectRay(instances, __float_as_int(b.w), to, td);
			float t, u, v;
			if (triIntersectE(to, td, mk3(a), mk3(b), mk3(c), tmax, t, u, v) && tryHit(bvh, a, b, t, u, v, accept) && ANY) return true;
		}
		return false;
	}
	// flattened instances: the ray in the object space of instance `ti` (cached while consecutive triangles belong to it)
	KRR_DEV void objectRay(const InstRec *__restrict__ instances, int ti, V3 &to, V3 &td) {
		if (ti != objInst) {
			const InstRec &in = instances[ti];
			oo = xfPointX(in.inv, o), od = xfVectorX(in.inv, d);
			objInst = ti;
		}
		to = oo, td = od;
	}
	// ---- phase: triangles of the current group (up to `budget` of them).  Returns true when an any-hit ray terminated. ----
	template <typename Accept> KRR_DEV bool triStep(const BvhDev &bvh, const InstRec *__restrict__ instances, Accept accept, int budget) {
		if (tg.y & 0x80000000u) return walkFlat(bvh, instances, accept);
		const bool flattened = bvh.mergedXf && curInst == bvh.mergedInst;
		for (int it = 0; it < budget && tg.y; it++) {
			const uint32_t bit = 31u - (uint32_t) __clz(tg.y);
			tg.y ^= 1u << bit;
			const float4 *tp = reinterpret_cast<const float4 *>(bvh.tris + tg.x + bit);
			const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
#ifdef KRR_COUNT_TRIPS
			triTests++;
#endif
			V3 to = ro, td = rd;
			if (flattened) objectRay(instances, __float_as_int(b.w), to, td); // the triangle lives in the object space of its own instance
			float t, u, v;
			if (triIntersectE(to, td, mk3(a), mk3(b), mk3(c), tmax, t, u, v) && tryHit(bvh, a, b, t, u, v, accept) && ANY) return true;
		}
		return false;
	}

	// Lane-local traversal to the end (no warp-level primitives): used where one lane runs several
	// dependent traversals interleaved with other work (ratio-tracking shadow rays through media).
	template <typename Accept> KRR_DEV void runToEnd(const BvhDev &bvh, const InstRec *__restrict__ instances, TraceSmem &sm, LStack &ls, Accept accept) {
		while (true) {
			if (!hasNode() && !tg.y && !popNext(sm, ls)) return;
			if (!tg.y) nodeStep(bvh, sm, ls);
			while (tg.y) {
				if (curInst < 0) enterInstance(bvh, instances, sm, ls); // leaves a flat list in tg, or a root in ng
				else if (triStep(bvh, instances, accept, 24)) return;
			}
		}
	}

	// One warp-cooperative trip: every phase runs under a warp vote, so that it executes with as many lanes as
	// possible.  Returns true for lanes whose ray finished during this trip.
	// VOTE: the triangle phase waits until KRR_TRI_VOTE lanes hold triangles (or no lane has node work left);
	// VOTE = false runs it every trip: fewer trips per ray, which wins for the short any-hit traversals of
	// shadow rays.
	template <bool VOTE, typename Accept>
	KRR_DEV bool trip(bool active, const BvhDev &bvh, const InstRec *__restrict__ instances, TraceSmem &sm, LStack &ls, Accept accept) {
		if constexpr (PAIR) { // the whole scene is one triangle list: one trip per ray
			if (active && tg.y) walkFlat(bvh, instances, accept);
			return active;
		} else {
			const unsigned FULL = 0xffffffffu;
			bool fin = false;
			if (active && !parked() && !hasNode() && !tg.y) fin = !popNext(sm, ls);
			const bool live = active && !fin;
			if constexpr (MOTION) {
				if (live && tg.y && curInst == -1) probeAndPark(bvh, instances, sm, ls);
				const bool wantEnter = live && parked();
				const unsigned mE	 = __ballot_sync(FULL, wantEnter);
				bool runE = mE != 0u;
				if (runE && __popc(mE) < KRR_ENTER_VOTE) runE = !__any_sync(FULL, live && !wantEnter); // every live lane waits to enter
				if (runE && wantEnter) enterProbed(bvh, instances, sm, ls, -2 - curInst);
			} else {
				if (live && tg.y && curInst < 0) enterInstance(bvh, instances, sm, ls);
			}
			const bool wantNode = live && !parked() && !tg.y && hasNode();
			if (__any_sync(FULL, wantNode)) {
				if (wantNode) nodeStep(bvh, sm, ls);
			}
			const bool wantTri = live && tg.y && curInst >= 0;
			const unsigned mT  = __ballot_sync(FULL, wantTri);
			bool run = mT != 0u;
			if (VOTE && run && __popc(mT) < KRR_TRI_VOTE) run = !__any_sync(FULL, live && !parked() && !tg.y && hasNode());
			if (run && wantTri) fin = triStep(bvh, instances, accept, KRR_TRI_PER_TRIP);
			return fin;
		}
	}
};

#endif // __CUDACC__

} // namespace krr
