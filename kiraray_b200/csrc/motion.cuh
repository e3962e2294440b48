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

} // namespace krr
