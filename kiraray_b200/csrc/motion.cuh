// motion.cuh -- transform chains with SRT motion keys (motion blur over a multi-level scene graph).
//
// The reference builds an IAS per scene-graph node and wraps animated nodes in an
// OptixSRTMotionTransform (src/core/device/optix.cpp:400-563); OptiX then evaluates the transform
// list of a hit at the ray's time.  Here the graph above every mesh instance is a flat chain of
// XformNodeRec (leaf node first, parent links towards the root); an instance that has at least one
// motion node in its chain is "moving" and its object<->world transforms are evaluated per ray from
// the chain at the ray's time, everything else keeps the precomputed InstRec::xf / inv.
//
// THE SPEC (identical, operation by operation, in oracle/driver.cpp -- every product and sum is
// individually rounded, so object-space rays agree bit-for-bit):
//   key interpolation (OptiX SRT semantics): time clamped to [t0, t1]; u = (time-t0)/(t1-t0)*(n-1);
//     k = min(int(u), n-2); f = u-k; every component a + f*(b-a); quaternion then divided by its norm
//   node matrix       M = T * R(q) * S                 node inverse  M^-1 = S^-1 * R(q)^T * T^-1
//   chain             M = M_root * ... * M_leaf        M^-1 = M_leaf^-1 * ... * M_root^-1
#pragma once
#include "krr_math.cuh"

namespace krr {

struct XformNodeRec { // 128 B
	int32_t parent;	  // next node towards the root, -1 = none
	int32_t keyOff;	  // first key (10 floats each: s[3], q[4] xyzw, t[3]) in the key pool
	int32_t nKeys;	  // < 2: static node (local / localInv below)
	int32_t pad0;
	float t0, t1;
	float pad1[2];
	Xf local, localInv;
};

KRR_HD Xf xfMulX(const Xf &a, const Xf &b) { // a * b, unfused, left to right
	Xf c;
#pragma unroll
	for (int r = 0; r < 3; r++) {
#pragma unroll
		for (int k = 0; k < 4; k++) {
			float s = xadd(xadd(xmul(a.m[r * 4], b.m[k]), xmul(a.m[r * 4 + 1], b.m[4 + k])), xmul(a.m[r * 4 + 2], b.m[8 + k]));
			c.m[r * 4 + k] = k == 3 ? xadd(s, a.m[r * 4 + 3]) : s;
		}
	}
	return c;
}

// interpolated SRT values v = (s[3], q[4] xyzw, t[3]) -> node matrix and inverse
KRR_HD void srtMatrices(const float *v, Xf &m, Xf &inv) {
	float len = xsqrt(xadd(xadd(xmul(v[3], v[3]), xmul(v[4], v[4])), xadd(xmul(v[5], v[5]), xmul(v[6], v[6]))));
	// one correctly rounded reciprocal and four products instead of four divisions, three reciprocals for the nine
	// entries of the inverse: 14 divisions per node were 15 % of the instructions of the motion-blur trace kernel
	const float rl = xrcp(len);
	float x = xmul(v[3], rl), y = xmul(v[4], rl), z = xmul(v[5], rl), w = xmul(v[6], rl);
	const float rs[3] = {xrcp(v[0]), xrcp(v[1]), xrcp(v[2])};
	float xx = xmul(x, x), yy = xmul(y, y), zz = xmul(z, z), xy = xmul(x, y), xz = xmul(x, z), yz = xmul(y, z);
	float wx = xmul(w, x), wy = xmul(w, y), wz = xmul(w, z);
	float R[9] = {xsub(1.f, xmul(2.f, xadd(yy, zz))), xmul(2.f, xsub(xy, wz)), xmul(2.f, xadd(xz, wy)),
				  xmul(2.f, xadd(xy, wz)), xsub(1.f, xmul(2.f, xadd(xx, zz))), xmul(2.f, xsub(yz, wx)),
				  xmul(2.f, xsub(xz, wy)), xmul(2.f, xadd(yz, wx)), xsub(1.f, xmul(2.f, xadd(xx, yy)))};
#pragma unroll
	for (int r = 0; r < 3; r++) {
#pragma unroll
		for (int c = 0; c < 3; c++) {
			m.m[r * 4 + c]	 = xmul(R[r * 3 + c], v[c]);
			inv.m[r * 4 + c] = xmul(R[c * 3 + r], rs[r]);
		}
		m.m[r * 4 + 3] = v[7 + r];
	}
#pragma unroll
	for (int r = 0; r < 3; r++)
		inv.m[r * 4 + 3] = -xadd(xadd(xmul(inv.m[r * 4], v[7]), xmul(inv.m[r * 4 + 1], v[8])), xmul(inv.m[r * 4 + 2], v[9]));
}

// SRT keys -> node matrix and inverse at `time`
KRR_HD void srtNodeXf(const float *__restrict__ keys, int n, float t0, float t1, float time, Xf &m, Xf &inv) {
	int k	= 0;
	float f = 0.f;
	if (time >= t1) k = n - 2, f = 1.f;
	else if (time > t0) {
		float u = xmul(xdiv(xsub(time, t0), xsub(t1, t0)), (float) (n - 1));
		k		= (int) u;
		if (k > n - 2) k = n - 2;
		f = xsub(u, (float) k);
	}
	const float *a = keys + 10 * k, *b = a + 10;
	float v[10];
#pragma unroll
	for (int i = 0; i < 10; i++) v[i] = xadd(a[i], xmul(f, xsub(b[i], a[i])));
	srtMatrices(v, m, inv);
}

KRR_HD void nodeXf(const XformNodeRec &nd, const float *__restrict__ keyPool, float time, Xf &m, Xf &inv) {
	if (nd.nKeys >= 2) srtNodeXf(keyPool + 10 * (size_t) nd.keyOff, nd.nKeys, nd.t0, nd.t1, time, m, inv);
	else m = nd.local, inv = nd.localInv;
}

// object->world and world->object of the chain that starts at `node`, at `time`
KRR_HD void chainXf(const XformNodeRec *__restrict__ nodes, const float *__restrict__ keyPool, int node, float time, Xf &m, Xf &inv) {
	nodeXf(nodes[node], keyPool, time, m, inv);
	for (int p = nodes[node].parent; p >= 0; p = nodes[p].parent) {
		Xf pm, pinv;
		nodeXf(nodes[p], keyPool, time, pm, pinv);
		m	= xfMulX(pm, m);
		inv = xfMulX(inv, pinv);
	}
}

// ---- flat motion records ---------------------------------------------------------------------------------------
// chainXf walks instance -> transform node -> key pool -> parent node -> key pool: four or five DEPENDENT loads
// before the first multiplication, with the warp waiting on each (the trace kernels of the 10 000-instance scene are
// latency-bound).  The common chain -- at most two levels, each an SRT node with exactly two keys -- is therefore
// also stored per INSTANCE as one contiguous record (krr_wfpt_set_scene): every load depends on the instance id
// alone.  Record = kMotionFlatStride float4:
//   [0] (levels (int bits): 0 = no record, use the chain; 1; 2,  -, -, -)   [1] (t0, t1 of level 0, t0, t1 of level 1)
//   [2..6] level 0 (the instance's own node): key a[10], then b[i] - a[i] (rounded once, as xsub(b[i], a[i]) is)
//   [7..11] level 1 (its parent)
// The arithmetic is THE SPEC's, operation by operation: with two keys u = (time-t0)/(t1-t0) * 1 = the quotient itself,
// k = 0 (u <= 1; int(u) = 1 is clamped to n - 2 = 0) and f = u - 0 = u, so the interpolated values, and everything
// computed from them by srtMatrices / xfMulX, are the floats chainXf produces (tests/test_gpu_motion.py compares the
// tap that runs this path with the oracle bit for bit).
constexpr int kMotionFlatStride = 12;
#ifdef __CUDACC__
KRR_DEV void flatLevelXf(const float4 *__restrict__ p, float t0, float t1, float time, Xf &m, Xf &inv) {
	float f = 0.f;
	if (time >= t1) f = 1.f;
	else if (time > t0) f = xdiv(xsub(time, t0), xsub(t1, t0));
	const float4 q0 = __ldg(p), q1 = __ldg(p + 1), q2 = __ldg(p + 2), q3 = __ldg(p + 3), q4 = __ldg(p + 4);
	const float a[10]	 = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y};
	const float diff[10] = {q2.z, q2.w, q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, q4.z, q4.w};
	float v[10];
#pragma unroll
	for (int i = 0; i < 10; i++) v[i] = xadd(a[i], xmul(f, diff[i]));
	srtMatrices(v, m, inv);
}
// object->world and world->object of a moving instance from its flat record; false = no record (use chainXf).
// KRR_FLAT_LOOP=1 runs the levels as a loop (one copy of the level code: a smaller instruction footprint, but the
// matrices then live in local memory: measured slower)
#ifndef KRR_FLAT_LOOP
#define KRR_FLAT_LOOP 0
#endif
KRR_DEV bool flatChainXf(const float4 *__restrict__ rec, float time, Xf &m, Xf &inv) {
	const int levels = __float_as_int(__ldg(rec).x);
	if (levels == 0) return false;
	const float4 tt = __ldg(rec + 1);
	if (levels == 2) asm volatile("prefetch.global.L1 [%0];" ::"l"(rec + 7)); // the parent's keys arrive while level 0 is evaluated
#if KRR_FLAT_LOOP
#pragma unroll 1
	for (int l = 0; l < levels; l++) {
		Xf lm, linv;
		flatLevelXf(rec + 2 + 5 * l, l ? tt.z : tt.x, l ? tt.w : tt.y, time, lm, linv);
		if (l == 0) m = lm, inv = linv;
		else m = xfMulX(lm, m), inv = xfMulX(inv, linv);
	}
#else
	flatLevelXf(rec + 2, tt.x, tt.y, time, m, inv);
	if (levels == 2) {
		Xf pm, pinv;
		flatLevelXf(rec + 7, tt.z, tt.w, time, pm, pinv);
		m	= xfMulX(pm, m);
		inv = xfMulX(inv, pinv);
	}
#endif
	return true;
}
// the general chain, kept out of line (and out of the hot loop's instruction footprint)
static __device__ __noinline__ void chainXfCold(const XformNodeRec *__restrict__ nodes, const float *__restrict__ keyPool, int node, float time, Xf *m, Xf *inv) {
	chainXf(nodes, keyPool, node, time, *m, *inv);
}
// transforms of moving instance `inst` (chain starting at `node`) at `time`: the flat record when there is one
KRR_DEV void movingInstanceXf(const XformNodeRec *__restrict__ nodes, const float *__restrict__ keyPool, const float4 *__restrict__ flat, int inst, int node,
							  float time, Xf &m, Xf &inv) {
	if (flat && flatChainXf(flat + (size_t) inst * kMotionFlatStride, time, m, inv)) return;
#ifndef KRR_CHAIN_COLD
#define KRR_CHAIN_COLD 0 // measured: the call makes movingRay a non-leaf function, 695 -> 669 Mrays/s
#endif
#if KRR_CHAIN_COLD
	chainXfCold(nodes, keyPool, node, time, &m, &inv);
#else
	chainXf(nodes, keyPool, node, time, m, inv);
#endif
}
#endif

} // namespace krr
