// bvh_build.cu -- device-side acceleration-structure build and refit.
//
// Replaces optixAccelBuild / optixAccelCompact / OPTIX_BUILD_OPERATION_UPDATE (reference
// src/core/device/optix.cpp:143-250 GAS per mesh, :357-398 single-level IAS, :618-669 refit).
// Pipeline (all on the GPU, one stream):
//   1. primitive AABBs + centroid bounds                       (k_prim_bounds_*)
//   2. 63-bit Morton codes of the centroids, radix sort        (k_morton + cub::DeviceRadixSort)
//   3. Karras-2012 binary radix tree, bottom-up AABB fit        (k_radix_tree, k_fit)
//   4. top-down collapse to 8-wide nodes: a wide node opens the child with the largest surface
//      area until it has 8 children (greedy SAH), leaves hold up to `maxLeaf` primitives; child
//      boxes are quantised to 8 bits against a padded power-of-two frame      (k_collapse)
// TLAS refit (per-frame instance transform updates): instance boxes are recomputed and the wide
// nodes re-fitted level by level, deepest first, reusing the topology        (k_refit_level).
#include "bvh_build.h"

#include <cub/cub.cuh>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

namespace krr {

namespace {

struct Aabb { float lo[3], hi[3]; };
struct WorkItem { int32_t bin, out; }; // collapse work list: binary node -> output slot of its wide node

KRR_DEV void atomicMinF(float *addr, float v) {
	// ordered-int trick; valid for all finite floats
	if (v >= 0) atomicMin((int *) addr, __float_as_int(v));
	else atomicMax((unsigned int *) addr, __float_as_uint(v));
}
KRR_DEV void atomicMaxF(float *addr, float v) {
	if (v >= 0) atomicMax((int *) addr, __float_as_int(v));
	else atomicMin((unsigned int *) addr, __float_as_uint(v));
}

__global__ void k_init_bounds(float *cb) {
	if (threadIdx.x < 3) cb[threadIdx.x] = 3.0e38f;
	else if (threadIdx.x < 6) cb[threadIdx.x] = -3.0e38f;
}

__global__ void k_prim_bounds_tris(const float *__restrict__ pos, const int32_t *__restrict__ idx, int n, Aabb *boxes, float *cb) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	Aabb b;
	for (int k = 0; k < 3; k++) b.lo[k] = 3.0e38f, b.hi[k] = -3.0e38f;
	for (int c = 0; c < 3; c++) {
		int v = idx[3 * i + c];
		for (int k = 0; k < 3; k++) {
			float p = pos[3 * v + k];
			b.lo[k] = fminf(b.lo[k], p), b.hi[k] = fmaxf(b.hi[k], p);
		}
	}
	boxes[i] = b;
	for (int k = 0; k < 3; k++) {
		float c = 0.5f * (b.lo[k] + b.hi[k]);
		atomicMinF(cb + k, c), atomicMaxF(cb + 3 + k, c);
	}
}

// world box of an instance = transformed 8 corners of its mesh's object-space box.  A MOVING instance
// (SRT motion chain, motion.cuh) is bounded over the time window [w0, w1] the rays of the frame can
// carry (the camera's shutter interval): its corners are evaluated at kMotionSamples + 1 times, all
// positions are united, and the box is padded by the largest second difference of a corner's
// trajectory -- 8x the deviation of a smooth curve from the chords between consecutive samples.
constexpr int kMotionSamples = 8;
__global__ void k_prim_bounds_insts(const InstRec *__restrict__ inst, const Aabb *__restrict__ meshBoxes, const int32_t *__restrict__ ids, int n,
									Aabb *boxesByInst, Aabb *boxesCompact, float *cb, const XformNodeRec *__restrict__ xnodes,
									const float *__restrict__ keyPool, float w0, float w1) {
	// ids: the instances that are TLAS primitives (merged instances are not; the merged BLAS is pseudo-
	// instance nInstances).  Boxes are stored by instance id (refit) and, for the build, by TLAS primitive.
	const int slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= n) return;
	const int i = ids[slot];
	const Aabb mb = meshBoxes[inst[i].mesh];
	Aabb b;
	for (int k = 0; k < 3; k++) b.lo[k] = 3.0e38f, b.hi[k] = -3.0e38f;
	auto grow = [&](V3 w) {
		b.lo[0] = fminf(b.lo[0], w.x), b.lo[1] = fminf(b.lo[1], w.y), b.lo[2] = fminf(b.lo[2], w.z);
		b.hi[0] = fmaxf(b.hi[0], w.x), b.hi[1] = fmaxf(b.hi[1], w.y), b.hi[2] = fmaxf(b.hi[2], w.z);
	};
	auto corner = [&](int c) { return mk3(c & 1 ? mb.hi[0] : mb.lo[0], c & 2 ? mb.hi[1] : mb.lo[1], c & 4 ? mb.hi[2] : mb.lo[2]); };
	float pad = 0.f;
	if (inst[i].motion >= 0 && xnodes) {
		const int steps = w1 > w0 ? kMotionSamples : 0;
		V3 prev[8], prev2[8];
		for (int j = 0; j <= steps; j++) {
			float t = steps ? w0 + (w1 - w0) * ((float) j / (float) steps) : w0;
			if (j == steps) t = w1;
			Xf m, inv;
			chainXf(xnodes, keyPool, inst[i].motion, t, m, inv);
			for (int c = 0; c < 8; c++) {
				V3 w = xfPoint(m, corner(c));
				grow(w);
				if (j >= 2) {
					V3 dd = prev2[c] - 2.f * prev[c] + w;
					pad	  = fmaxf(pad, fmaxf(fabsf(dd.x), fmaxf(fabsf(dd.y), fabsf(dd.z))));
				}
				prev2[c] = prev[c], prev[c] = w;
			}
		}
	} else {
		for (int c = 0; c < 8; c++) grow(xfPoint(inst[i].xf, corner(c)));
	}
	// rays are intersected in object space with the rounded inverse transform: pad the world box so
	// that culling stays conservative w.r.t. that round trip
	for (int k = 0; k < 3; k++) {
		float e = 1e-5f * fmaxf(1.f, fmaxf(fabsf(b.lo[k]), fabsf(b.hi[k]))) + 1e-6f * (b.hi[k] - b.lo[k]) + pad;
		b.lo[k] -= e, b.hi[k] += e;
	}
	boxesByInst[i] = b;
	if (boxesCompact) boxesCompact[slot] = b;
	if (cb)
		for (int k = 0; k < 3; k++) {
			float c = 0.5f * (b.lo[k] + b.hi[k]);
			atomicMinF(cb + k, c), atomicMaxF(cb + 3 + k, c);
		}
}

// merged BLAS: triangle boxes of every merged instance (identity transform: object space == world
// space), concatenated; pairs[t] = (group slot, local primitive).  blockIdx.y strides over the merged
// instances, blockIdx.x over 256-triangle chunks.
struct MergedSrc { int32_t posOff, idxOff, nTri, inst, outOff; };
__global__ void k_prim_bounds_merged(const float *__restrict__ positions, const int32_t *__restrict__ indices, const MergedSrc *__restrict__ src,
									 int nSrc, Aabb *boxes, int2 *pairs, float *cb) {
	for (int j = blockIdx.y; j < nSrc; j += gridDim.y) {
		const MergedSrc ms = src[j];
		const float *pos   = positions + 3 * (size_t) ms.posOff;
		const int32_t *idx = indices + 3 * (size_t) ms.idxOff;
		for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ms.nTri; i += gridDim.x * blockDim.x) {
			Aabb b;
			for (int k = 0; k < 3; k++) b.lo[k] = 3.0e38f, b.hi[k] = -3.0e38f;
			for (int c = 0; c < 3; c++) {
				int v = idx[3 * i + c];
				for (int k = 0; k < 3; k++) {
					float p = pos[3 * v + k];
					b.lo[k] = fminf(b.lo[k], p), b.hi[k] = fmaxf(b.hi[k], p);
				}
			}
			boxes[ms.outOff + i] = b;
			pairs[ms.outOff + i] = make_int2(j, i);
			for (int k = 0; k < 3; k++) {
				float c = 0.5f * (b.lo[k] + b.hi[k]);
				atomicMinF(cb + k, c), atomicMaxF(cb + 3 + k, c);
			}
		}
	}
}

KRR_DEV uint64_t expand21(uint32_t v) { // spread 21 bits to every third bit
	uint64_t x = v & 0x1fffff;
	x = (x | x << 32) & 0x1f00000000ffffULL;
	x = (x | x << 16) & 0x1f0000ff0000ffULL;
	x = (x | x << 8) & 0x100f00f00f00f00fULL;
	x = (x | x << 4) & 0x10c30c30c30c30c3ULL;
	x = (x | x << 2) & 0x1249249249249249ULL;
	return x;
}

__global__ void k_morton(const Aabb *__restrict__ boxes, int n, const float *__restrict__ cb, uint64_t *keys, uint32_t *vals) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint32_t q[3];
	for (int k = 0; k < 3; k++) {
		float ext = cb[3 + k] - cb[k];
		float c	  = 0.5f * (boxes[i].lo[k] + boxes[i].hi[k]);
		float u	  = ext > 0 ? (c - cb[k]) / ext : 0.f;
		q[k] = (uint32_t) fminf(fmaxf(u * 2097152.f, 0.f), 2097151.f);
	}
	keys[i] = (expand21(q[0]) << 2) | (expand21(q[1]) << 1) | expand21(q[2]);
	vals[i] = (uint32_t) i;
}

// binary radix tree over sorted keys (Karras 2012). internal nodes 0..n-2, leaf i = node (n-1)+i
struct BinTree {
	int32_t *left, *right, *parent; // per internal node / per node
	int32_t *first, *last;			// sorted-primitive range covered by each node
	Aabb *bounds;					// per node (2n-1)
	int32_t *flags;					// per internal node, for the bottom-up fit
};

KRR_DEV int delta(const uint64_t *keys, int n, int i, int j) {
	if (j < 0 || j >= n) return -1;
	uint64_t a = keys[i], b = keys[j];
	if (a == b) return 64 + __clz(i ^ j);
	return __clzll(a ^ b);
}

__global__ void k_radix_tree(const uint64_t *__restrict__ keys, int n, BinTree t) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n - 1) return;
	int d	 = (delta(keys, n, i, i + 1) - delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
	int dmin = delta(keys, n, i, i - d);
	int lmax = 2;
	while (delta(keys, n, i, i + lmax * d) > dmin) lmax *= 2;
	int l = 0;
	for (int tt = lmax / 2; tt >= 1; tt /= 2)
		if (delta(keys, n, i, i + (l + tt) * d) > dmin) l += tt;
	int j	  = i + l * d;
	int dnode = delta(keys, n, i, j);
	int s	  = 0;
	for (int div = 2, tt = (l + div - 1) / div; ; div *= 2, tt = (l + div - 1) / div) {
		if (delta(keys, n, i, i + (s + tt) * d) > dnode) s += tt;
		if (tt <= 1) break;
	}
	int gamma = i + s * d + min(d, 0);
	int lo = min(i, j), hi = max(i, j);
	int lc = (lo == gamma) ? (n - 1) + gamma : gamma;
	int rc = (hi == gamma + 1) ? (n - 1) + gamma + 1 : gamma + 1;
	t.left[i] = lc, t.right[i] = rc;
	t.parent[lc] = i, t.parent[rc] = i;
	t.first[i] = lo, t.last[i] = hi;
	if (i == 0) t.parent[0] = -1;
}

__global__ void k_fit(const Aabb *__restrict__ primBoxes, const uint32_t *__restrict__ sortedIdx, int n, BinTree t) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int node = (n - 1) + i;
	t.bounds[node] = primBoxes[sortedIdx[i]];
	t.first[node] = t.last[node] = i;
	int p = t.parent[node];
	while (p >= 0) {
		__threadfence();
		if (atomicAdd(&t.flags[p], 1) == 0) return; // first child to arrive stops; second continues
		// children boxes were written by other SMs: read them through L2 (L1 is not coherent)
		Aabb a, b, r;
		const float *pa = (const float *) &t.bounds[t.left[p]], *pb = (const float *) &t.bounds[t.right[p]];
		for (int k = 0; k < 3; k++) a.lo[k] = __ldcg(pa + k), a.hi[k] = __ldcg(pa + 3 + k), b.lo[k] = __ldcg(pb + k), b.hi[k] = __ldcg(pb + 3 + k);
		for (int k = 0; k < 3; k++) r.lo[k] = fminf(a.lo[k], b.lo[k]), r.hi[k] = fmaxf(a.hi[k], b.hi[k]);
		t.bounds[p] = r;
		p = t.parent[p];
	}
}

KRR_DEV float halfArea(const Aabb &b) {
	float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
	return dx * dy + dy * dz + dz * dx;
}

// quantisation frame of a node: padded so that every child plane keeps >= 1 quantum of slack
KRR_DEV void makeFrame(const Aabb &nb, float o[3], uint32_t e[3]) {
	for (int k = 0; k < 3; k++) {
		float ext = fmaxf(nb.hi[k] - nb.lo[k], 1e-20f);
		float pad = 0.02f * ext + 1e-7f * fmaxf(fabsf(nb.lo[k]), fabsf(nb.hi[k]));
		o[k]	  = nb.lo[k] - pad;
		float span = (nb.hi[k] + pad) - o[k];
		int ex;
		frexpf(span / 255.f, &ex); // span/255 = m * 2^ex, m in [0.5,1) -> 2^ex >= span/255
		int be = min(max(ex + 127, 1), 254);
		e[k]   = (uint32_t) be;
	}
}
KRR_DEV void quantize(const Aabb &cb, const float o[3], const uint32_t e[3], uint8_t qlo[3], uint8_t qhi[3]) {
	for (int k = 0; k < 3; k++) {
		float s	 = __uint_as_float(e[k] << 23);
		float fl = floorf((cb.lo[k] - o[k]) / s) - 1.f, fh = ceilf((cb.hi[k] - o[k]) / s) + 1.f;
		qlo[k] = (uint8_t) fminf(fmaxf(fl, 0.f), 255.f);
		qhi[k] = (uint8_t) fminf(fmaxf(fh, 0.f), 255.f);
	}
}

struct CollapseOut {
	Node8 *nodes;		// output pool
	Aabb *nodeBounds;	// full-precision box per wide node (refit / TLAS)
	int32_t *counters;	// [0] next free node, [1] next free primitive slot
	uint32_t nodeBase;	// offset of this tree's nodes in the global pool
	uint32_t primBase;	// offset of this tree's primitives in the global pool
};


// leaf payload writers
struct TriWriter {
	const float *pos;
	const int32_t *idx;
	BvhTri *tris;
	KRR_DEV void operator()(uint32_t slot, uint32_t prim) const {
		BvhTri t;
		int a = idx[3 * prim], b = idx[3 * prim + 1], c = idx[3 * prim + 2];
		const V3 v0 = mk3(pos[3 * a], pos[3 * a + 1], pos[3 * a + 2]);
		const V3 e1 = xsub3(mk3(pos[3 * b], pos[3 * b + 1], pos[3 * b + 2]), v0), e2 = xsub3(mk3(pos[3 * c], pos[3 * c + 1], pos[3 * c + 2]), v0);
		t.v0 = make_float4(v0.x, v0.y, v0.z, __int_as_float((int) prim));
		t.e1 = make_float4(e1.x, e1.y, e1.z, 0.f);
		t.e2 = make_float4(e2.x, e2.y, e2.z, 0.f);
		tris[slot] = t;
	}
};
struct MergedTriWriter { // triangles of the merged BLAS carry (primitive, instance)
	const float *positions;
	const int32_t *indices;
	const MergedSrc *src;
	const int2 *pairs;
	BvhTri *tris;
	KRR_DEV void operator()(uint32_t slot, uint32_t prim) const {
		const int2 pr	   = pairs[prim];
		const MergedSrc ms = src[pr.x];
		const float *pos   = positions + 3 * (size_t) ms.posOff;
		const int32_t *idx = indices + 3 * ((size_t) ms.idxOff + pr.y);
		int a = idx[0], b = idx[1], c = idx[2];
		BvhTri t;
		const V3 v0 = mk3(pos[3 * a], pos[3 * a + 1], pos[3 * a + 2]);
		const V3 e1 = xsub3(mk3(pos[3 * b], pos[3 * b + 1], pos[3 * b + 2]), v0), e2 = xsub3(mk3(pos[3 * c], pos[3 * c + 1], pos[3 * c + 2]), v0);
		t.v0 = make_float4(v0.x, v0.y, v0.z, __int_as_float(pr.y));
		t.e1 = make_float4(e1.x, e1.y, e1.z, __int_as_float(ms.inst));
		t.e2 = make_float4(e2.x, e2.y, e2.z, 0.f);
		tris[slot] = t;
	}
};
struct InstWriter {
	int32_t *tlasInst;
	const int32_t *ids; // TLAS primitive -> instance id
	KRR_DEV void operator()(uint32_t slot, uint32_t prim) const { tlasInst[slot] = ids[prim]; }
};

// flat BLAS (<= flatMax triangles): the leaf payload in primitive order, no nodes
template <typename Writer> __global__ void k_write_flat(int n, uint32_t base, Writer writer) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) writer(base + i, (uint32_t) i);
}

template <typename Writer>
__global__ void k_collapse(const WorkItem *__restrict__ in, int nIn, WorkItem *out, int32_t *nOut, BinTree t, int n,
						   const uint32_t *__restrict__ sortedIdx, int maxLeaf, CollapseOut co, Writer writer) {
	int w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= nIn) return;
	WorkItem item = in[w];
	int c[8], nc = 0;
	auto count = [&](int node) { return t.last[node] - t.first[node] + 1; };
	auto isLeafUnit = [&](int node) { return node >= n - 1 || count(node) <= maxLeaf; };
	if (item.bin >= n - 1 || n == 1) c[nc++] = item.bin;
	else { c[nc++] = t.left[item.bin]; c[nc++] = t.right[item.bin]; }
	while (nc < 8) {
		int best = -1;
		float bestA = -1;
		for (int j = 0; j < nc; j++)
			if (!isLeafUnit(c[j])) {
				float a = halfArea(t.bounds[c[j]]);
				if (a > bestA) bestA = a, best = j;
			}
		if (best < 0) break;
		int node = c[best];
		c[best]	 = t.left[node];
		c[nc++]	 = t.right[node];
	}
	const Aabb nb = t.bounds[item.bin];
	float o[3];
	uint32_t e[3];
	makeFrame(nb, o, e);
	int nInternal = 0, nPrims = 0;
	for (int j = 0; j < nc; j++) {
		if (isLeafUnit(c[j])) nPrims += count(c[j]);
		else nInternal++;
	}
	uint32_t childBase = nInternal ? (uint32_t) atomicAdd(&co.counters[0], nInternal) : 0u;
	uint32_t primBase  = nPrims ? (uint32_t) atomicAdd(&co.counters[1], nPrims) : 0u;
	int outBase		   = nInternal ? atomicAdd(nOut, nInternal) : 0;
	Node8 node;
	node.ox = o[0], node.oy = o[1], node.oz = o[2];
	node.ex = (uint8_t) e[0], node.ey = (uint8_t) e[1], node.ez = (uint8_t) e[2];
	node.imask	   = 0;
	node.childBase = co.nodeBase + childBase;
	node.primBase  = co.primBase + primBase;
	int ii = 0, po = 0;
	for (int j = 0; j < 8; j++) {
		node.meta[j] = 0;
		for (int k = 0; k < 3; k++) node.qlo[k][j] = 255, node.qhi[k][j] = 0;
	}
	for (int j = 0; j < nc; j++) {
		uint8_t ql[3], qh[3];
		quantize(t.bounds[c[j]], o, e, ql, qh);
		for (int k = 0; k < 3; k++) node.qlo[k][j] = ql[k], node.qhi[k][j] = qh[k];
		if (isLeafUnit(c[j])) {
			int cnt		 = count(c[j]);
			node.meta[j] = (uint8_t) ((cnt << 5) | po);
			for (int k = 0; k < cnt; k++) writer(co.primBase + primBase + po + k, sortedIdx[t.first[c[j]] + k]);
			po += cnt;
		} else {
			node.imask |= (uint8_t) (1u << j);
			out[outBase + ii] = WorkItem{c[j], (int32_t) (childBase + ii)};
			ii++;
		}
	}
	co.nodes[co.nodeBase + item.out]	  = node;
	co.nodeBounds[co.nodeBase + item.out] = nb;
}

// ---- refit of the TLAS: topology is kept, boxes and quantisation are recomputed ----
__global__ void k_refit_level(Node8 *nodes, Aabb *nodeBounds, int first, int count, const Aabb *__restrict__ primBoxes,
							  const int32_t *__restrict__ tlasInst) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count) return;
	Node8 node = nodes[first + i];
	Aabb cb[8];
	bool used[8];
	Aabb nb;
	for (int k = 0; k < 3; k++) nb.lo[k] = 3.0e38f, nb.hi[k] = -3.0e38f;
	for (int j = 0; j < 8; j++) {
		used[j] = false;
		bool internal = (node.imask >> j) & 1;
		if (internal) {
			cb[j]	= nodeBounds[node.childBase + __popc(node.imask & ((1u << j) - 1))];
			used[j] = true;
		} else if (node.meta[j]) {
			int cnt = node.meta[j] >> 5, off = node.meta[j] & 31;
			for (int k = 0; k < 3; k++) cb[j].lo[k] = 3.0e38f, cb[j].hi[k] = -3.0e38f;
			for (int q = 0; q < cnt; q++) {
				const Aabb pb = primBoxes[tlasInst[node.primBase + off + q]];
				for (int k = 0; k < 3; k++) cb[j].lo[k] = fminf(cb[j].lo[k], pb.lo[k]), cb[j].hi[k] = fmaxf(cb[j].hi[k], pb.hi[k]);
			}
			used[j] = true;
		}
		if (used[j])
			for (int k = 0; k < 3; k++) nb.lo[k] = fminf(nb.lo[k], cb[j].lo[k]), nb.hi[k] = fmaxf(nb.hi[k], cb[j].hi[k]);
	}
	float o[3];
	uint32_t e[3];
	makeFrame(nb, o, e);
	node.ox = o[0], node.oy = o[1], node.oz = o[2];
	node.ex = (uint8_t) e[0], node.ey = (uint8_t) e[1], node.ez = (uint8_t) e[2];
	for (int j = 0; j < 8; j++)
		if (used[j]) {
			uint8_t ql[3], qh[3];
			quantize(cb[j], o, e, ql, qh);
			for (int k = 0; k < 3; k++) node.qlo[k][j] = ql[k], node.qhi[k][j] = qh[k];
		}
	nodes[first + i]	  = node;
	nodeBounds[first + i] = nb;
}

__global__ void k_mesh_box(const Aabb *__restrict__ boxes, int n, Aabb *out) {
	// single block reduction of primitive boxes -> object-space mesh box
	__shared__ float lo[3][256], hi[3][256];
	float l[3] = {3.0e38f, 3.0e38f, 3.0e38f}, h[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
	for (int i = threadIdx.x; i < n; i += blockDim.x)
		for (int k = 0; k < 3; k++) l[k] = fminf(l[k], boxes[i].lo[k]), h[k] = fmaxf(h[k], boxes[i].hi[k]);
	for (int k = 0; k < 3; k++) lo[k][threadIdx.x] = l[k], hi[k][threadIdx.x] = h[k];
	__syncthreads();
	for (int s = blockDim.x / 2; s > 0; s >>= 1) {
		if (threadIdx.x < s)
			for (int k = 0; k < 3; k++) {
				lo[k][threadIdx.x] = fminf(lo[k][threadIdx.x], lo[k][threadIdx.x + s]);
				hi[k][threadIdx.x] = fmaxf(hi[k][threadIdx.x], hi[k][threadIdx.x + s]);
			}
		__syncthreads();
	}
	if (threadIdx.x == 0)
		for (int k = 0; k < 3; k++) out->lo[k] = lo[k][0], out->hi[k] = hi[k][0];
}

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { snprintf(err, 256, "%s: %s", #x, cudaGetErrorString(e_)); return false; } } while (0)

template <typename T> struct DevBuf {
	T *p = nullptr;
	size_t n = 0;
	bool alloc(size_t count) {
		free();
		n = count;
		return cudaMalloc((void **) &p, std::max<size_t>(count, 1) * sizeof(T)) == cudaSuccess;
	}
	// grow-only: scratch that is reused across the trees of one build (cudaMalloc / cudaFree synchronise the
	// device and cost milliseconds each once gigabytes are resident -- 30 of them per mesh made the build of a
	// 200-mesh scene take 15-20 s)
	bool ensure(size_t count) { return count <= n && p ? true : alloc(count + count / 4); }
	void free() { if (p) cudaFree(p); p = nullptr; n = 0; }
	~DevBuf() { free(); }
};

struct TreeScratch {
	DevBuf<uint64_t> keys, keysSorted;
	DevBuf<uint32_t> vals, valsSorted;
	DevBuf<int32_t> left, right, parent, first, last, flags;
	DevBuf<Aabb> bounds;
	DevBuf<WorkItem> q0, q1;
	DevBuf<int32_t> qCount;
	DevBuf<unsigned char> tmp;
};

} // namespace

struct BvhBuilder::Impl {
	DevBuf<Node8> nodes;
	DevBuf<Aabb> nodeBounds;
	DevBuf<BvhTri> tris;
	DevBuf<int32_t> tlasInst;
	DevBuf<Aabb> meshBoxes, instBoxes;
	DevBuf<int32_t> counters, tlasIds;
	DevBuf<int2> flats;
	int nTlasPrims = 0, mergedRoot = -1, mergedInst = -1, nMergedTris = 0, nFlat = 0;
	Aabb mergedBox{{0, 0, 0}, {0, 0, 0}};
	std::vector<int> tlasLevelStart; // node index (relative to the pool) where each TLAS level begins
	int tlasNodeCount = 0, totalNodes = 0, totalTris = 0, nInstances = 0, nMeshes = 0;
	std::vector<int32_t> blasRoots, triBases;
	MotionWindow motion{};
};

BvhBuilder::BvhBuilder() : m(new Impl) {}
BvhBuilder::~BvhBuilder() { delete m; }

namespace {
// Builds one wide tree over `n` primitives whose boxes are in `boxes`; nodes are appended to the
// pool at *nodeCursor, primitives at *primCursor.  Returns the root index; levelStart (optional)
// receives the pool index of the first node of each level.
template <typename Writer>
bool buildTree(const Aabb *boxes, int n, int maxLeaf, Node8 *nodePool, Aabb *boundsPool, int32_t *counters, int &nodeCursor,
			   int &primCursor, Writer writer, cudaStream_t stream, float *cb, std::vector<int> *levelStart, int *root, char *err,
			   TreeScratch &sc) {
	const int T = 256;
	DevBuf<uint64_t> &keys = sc.keys, &keysSorted = sc.keysSorted;
	DevBuf<uint32_t> &vals = sc.vals, &valsSorted = sc.valsSorted;
	DevBuf<int32_t> &left = sc.left, &right = sc.right, &parent = sc.parent, &first = sc.first, &last = sc.last, &flags = sc.flags;
	DevBuf<Aabb> &bounds = sc.bounds;
	DevBuf<WorkItem> &q0 = sc.q0, &q1 = sc.q1;
	DevBuf<int32_t> &qCount = sc.qCount;
	DevBuf<unsigned char> &tmp = sc.tmp;
	if (!keys.ensure(n) || !keysSorted.ensure(n) || !vals.ensure(n) || !valsSorted.ensure(n) || !left.ensure(n) || !right.ensure(n) ||
		!parent.ensure(2 * n) || !first.ensure(2 * n) || !last.ensure(2 * n) || !flags.ensure(n) || !bounds.ensure(2 * n) ||
		!q0.ensure(n + 1) || !q1.ensure(n + 1) || !qCount.ensure(1)) {
		snprintf(err, 256, "bvh build: out of device memory for %d primitives", n);
		return false;
	}
	k_morton<<<(n + T - 1) / T, T, 0, stream>>>(boxes, n, cb, keys.p, vals.p);
	size_t tmpBytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, keys.p, keysSorted.p, vals.p, valsSorted.p, n, 0, 63, stream);
	if (!tmp.ensure(tmpBytes)) { snprintf(err, 256, "bvh build: sort scratch alloc failed"); return false; }
	CK(cub::DeviceRadixSort::SortPairs(tmp.p, tmpBytes, keys.p, keysSorted.p, vals.p, valsSorted.p, n, 0, 63, stream));
	BinTree t{left.p, right.p, parent.p, first.p, last.p, bounds.p, flags.p};
	CK(cudaMemsetAsync(flags.p, 0, sizeof(int32_t) * std::max(n, 1), stream));
	CK(cudaMemsetAsync(parent.p, 0xff, sizeof(int32_t) * 2 * n, stream));
	if (n > 1) k_radix_tree<<<(n - 1 + T - 1) / T, T, 0, stream>>>(keysSorted.p, n, t);
	k_fit<<<(n + T - 1) / T, T, 0, stream>>>(boxes, valsSorted.p, n, t);
	// collapse, level by level
	int32_t cnt[2] = {1, 0}; // node 0 of this tree is the root
	CK(cudaMemcpyAsync(counters, cnt, 8, cudaMemcpyHostToDevice, stream));
	WorkItem rootItem{n == 1 ? 0 : 0, 0}; // n == 1: the only node is leaf node (n-1)+0 = 0
	CK(cudaMemcpyAsync(q0.p, &rootItem, sizeof rootItem, cudaMemcpyHostToDevice, stream));
	CollapseOut co{nodePool, boundsPool, counters, (uint32_t) nodeCursor, (uint32_t) primCursor};
	int nIn = 1, levelFirst = 0;
	WorkItem *qin = q0.p, *qout = q1.p;
	*root = nodeCursor;
	while (nIn > 0) {
		if (levelStart) levelStart->push_back(nodeCursor + levelFirst);
		CK(cudaMemsetAsync(qCount.p, 0, 4, stream));
		k_collapse<<<(nIn + T - 1) / T, T, 0, stream>>>(qin, nIn, qout, qCount.p, t, n, valsSorted.p, maxLeaf, co, writer);
		int nOut = 0;
		CK(cudaMemcpyAsync(&nOut, qCount.p, 4, cudaMemcpyDeviceToHost, stream));
		CK(cudaStreamSynchronize(stream));
		levelFirst += nIn;
		nIn = nOut;
		std::swap(qin, qout);
	}
	CK(cudaMemcpyAsync(cnt, counters, 8, cudaMemcpyDeviceToHost, stream));
	CK(cudaStreamSynchronize(stream));
	if (levelStart) levelStart->push_back(nodeCursor + cnt[0]);
	nodeCursor += cnt[0];
	primCursor += cnt[1];
	CK(cudaGetLastError());
	return true;
}
} // namespace

bool BvhBuilder::build(const float *dPositions, const int32_t *dIndices, const MeshRec *hMeshes, int nMeshes,
					   const InstRec *dInstances, const InstRec *hInstances, int nInstances, const uint8_t *hMerge, int flatMax,
					   const MotionWindow &motion, cudaStream_t stream, char *err) {
	Impl &b = *m;
	// triangles per leaf child: the node format addresses 8 leaf children x maxLeaf <= 32 primitives
	const int maxLeaf = std::min(std::max(getenv("KRR_BVH_MAX_LEAF") ? atoi(getenv("KRR_BVH_MAX_LEAF")) : 3, 1), 4);
	std::vector<int2> flats;
	b.nMeshes = nMeshes, b.nInstances = nInstances;
	// which instances go into the merged world-space BLAS, which meshes still need a BLAS of their own
	std::vector<MergedSrc> msrc;
	std::vector<char> meshNeedsBlas(nMeshes, 0);
	std::vector<int32_t> tlasIds;
	size_t mergedTris = 0;
	for (int i = 0; i < nInstances; i++) {
		const MeshRec &mr = hMeshes[hInstances[i].mesh];
		if (hMerge && hMerge[i]) {
			msrc.push_back(MergedSrc{mr.posOff, mr.idxOff, mr.nTri, i, (int32_t) mergedTris});
			mergedTris += mr.nTri;
		} else {
			meshNeedsBlas[hInstances[i].mesh] = 1;
			tlasIds.push_back(i);
		}
	}
	if (mergedTris > 0x03ffffffu) { snprintf(err, 256, "bvh build: merged BLAS too large (%zu triangles)", mergedTris); return false; }
	const bool haveMerged = !msrc.empty();
	if (haveMerged) tlasIds.push_back(nInstances); // the pseudo-instance (api.cu appends its InstRec)
	b.mergedInst = haveMerged ? nInstances : -1;
	b.mergedRoot = -1;
	b.nMergedTris = (int) mergedTris;
	b.nTlasPrims  = (int) tlasIds.size();
	size_t totalTris = mergedTris;
	int maxTris = 1, nBlas = haveMerged ? 1 : 0;
	for (int i = 0; i < nMeshes; i++)
		if (meshNeedsBlas[i]) totalTris += hMeshes[i].nTri, maxTris = std::max(maxTris, hMeshes[i].nTri), nBlas++;
	const int tlasReserve = nInstances + 2;
	const size_t nodeCap  = totalTris + (size_t) nBlas + (size_t) tlasReserve + 8;
	if (!b.nodes.alloc(nodeCap) || !b.nodeBounds.alloc(nodeCap) || !b.tris.alloc(totalTris) || !b.tlasInst.alloc(nInstances + 1) ||
		!b.meshBoxes.alloc(nMeshes + 1) || !b.instBoxes.alloc(nInstances + 1) || !b.counters.alloc(2) || !b.tlasIds.alloc(tlasIds.size())) {
		snprintf(err, 256, "bvh build: out of device memory (%zu triangles)", totalTris);
		return false;
	}
	CK(cudaMemcpyAsync(b.tlasIds.p, tlasIds.data(), tlasIds.size() * sizeof(int32_t), cudaMemcpyHostToDevice, stream));
	TreeScratch scratch;
	DevBuf<Aabb> primBoxes;
	DevBuf<float> cb;
	if (!primBoxes.alloc(std::max<size_t>(std::max<size_t>(maxTris, mergedTris), tlasIds.size())) || !cb.alloc(6)) { snprintf(err, 256, "bvh build: alloc failed"); return false; }
	const int T = 256;
	// TLAS occupies the front of the pool: reserve its nodes first (upper bound: one per primitive + 1)
	int nodeCursor = tlasReserve, primCursor = 0;
	b.blasRoots.assign(nMeshes, -1), b.triBases.assign(nMeshes, -1);
	for (int i = 0; i < nMeshes; i++) {
		const MeshRec &mr = hMeshes[i];
		if (mr.nTri <= 0) { snprintf(err, 256, "bvh build: mesh %d has no triangles", i); return false; }
		if (!meshNeedsBlas[i]) continue; // only instanced through the merged BLAS
		k_init_bounds<<<1, 32, 0, stream>>>(cb.p);
		k_prim_bounds_tris<<<(mr.nTri + T - 1) / T, T, 0, stream>>>(dPositions + 3 * (size_t) mr.posOff, dIndices + 3 * (size_t) mr.idxOff,
																  mr.nTri, primBoxes.p, cb.p);
		k_mesh_box<<<1, 256, 0, stream>>>(primBoxes.p, mr.nTri, b.meshBoxes.p + i);
		TriWriter wr{dPositions + 3 * (size_t) mr.posOff, dIndices + 3 * (size_t) mr.idxOff, b.tris.p};
		b.triBases[i] = primCursor;
		if (mr.nTri <= flatMax) { // flat list instead of a tree
			k_write_flat<<<(mr.nTri + T - 1) / T, T, 0, stream>>>(mr.nTri, (uint32_t) primCursor, wr);
			b.blasRoots[i] = (int32_t) (kFlatFlag | (uint32_t) flats.size());
			flats.push_back(make_int2(primCursor, mr.nTri));
			primCursor += mr.nTri;
			continue;
		}
		int root = 0;
		if (!buildTree(primBoxes.p, mr.nTri, maxLeaf, b.nodes.p, b.nodeBounds.p, b.counters.p, nodeCursor, primCursor, wr, stream, cb.p, nullptr, &root, err, scratch))
			return false;
		b.blasRoots[i] = root;
	}
	if (haveMerged) {
		DevBuf<MergedSrc> dsrc;
		DevBuf<int2> pairs;
		if (!dsrc.alloc(msrc.size()) || !pairs.alloc(mergedTris)) { snprintf(err, 256, "bvh build: alloc failed (merged BLAS)"); return false; }
		CK(cudaMemcpyAsync(dsrc.p, msrc.data(), msrc.size() * sizeof(MergedSrc), cudaMemcpyHostToDevice, stream));
		int maxSrcTris = 1;
		for (const MergedSrc &s : msrc) maxSrcTris = std::max(maxSrcTris, s.nTri);
		dim3 grid((unsigned) std::min((maxSrcTris + T - 1) / T, 4096), (unsigned) std::min<size_t>(msrc.size(), 16384));
		k_init_bounds<<<1, 32, 0, stream>>>(cb.p);
		k_prim_bounds_merged<<<grid, T, 0, stream>>>(dPositions, dIndices, dsrc.p, (int) msrc.size(), primBoxes.p, pairs.p, cb.p);
		k_mesh_box<<<1, 256, 0, stream>>>(primBoxes.p, (int) mergedTris, b.meshBoxes.p + nMeshes);
		MergedTriWriter wr{dPositions, dIndices, dsrc.p, pairs.p, b.tris.p};
		if ((int) mergedTris <= flatMax) {
			k_write_flat<<<((int) mergedTris + T - 1) / T, T, 0, stream>>>((int) mergedTris, (uint32_t) primCursor, wr);
			b.mergedRoot = (int32_t) (kFlatFlag | (uint32_t) flats.size());
			flats.push_back(make_int2(primCursor, (int) mergedTris));
			primCursor += (int) mergedTris;
			CK(cudaStreamSynchronize(stream)); // dsrc / pairs die with this scope
		} else {
			int root = 0;
			if (!buildTree(primBoxes.p, (int) mergedTris, maxLeaf, b.nodes.p, b.nodeBounds.p, b.counters.p, nodeCursor, primCursor, wr, stream, cb.p, nullptr, &root, err, scratch))
				return false;
			b.mergedRoot = root;
		}
	}
	if (haveMerged) CK(cudaMemcpyAsync(&b.mergedBox, b.meshBoxes.p + nMeshes, sizeof(Aabb), cudaMemcpyDeviceToHost, stream));
	b.nFlat = (int) flats.size();
	if (!b.flats.alloc(flats.size())) { snprintf(err, 256, "bvh build: alloc failed (flat table)"); return false; }
	if (!flats.empty()) CK(cudaMemcpyAsync(b.flats.p, flats.data(), flats.size() * sizeof(int2), cudaMemcpyHostToDevice, stream));
	b.totalTris = primCursor;
	b.totalNodes = nodeCursor;
	// TLAS over the world boxes of its primitives (one instance per leaf child)
	b.motion = motion;
	b.tlasLevelStart.clear();
	b.tlasNodeCount = 0;
	if (b.nTlasPrims > 0) {
		k_init_bounds<<<1, 32, 0, stream>>>(cb.p);
		k_prim_bounds_insts<<<(b.nTlasPrims + T - 1) / T, T, 0, stream>>>(dInstances, b.meshBoxes.p, b.tlasIds.p, b.nTlasPrims, b.instBoxes.p, primBoxes.p,
																		 cb.p, motion.xnodes, motion.keys, motion.w0, motion.w1);
		int tlasCursor = 0, tlasPrims = 0, root = 0;
		InstWriter iw{b.tlasInst.p, b.tlasIds.p};
		if (!buildTree(primBoxes.p, b.nTlasPrims, 1, b.nodes.p, b.nodeBounds.p, b.counters.p, tlasCursor, tlasPrims, iw, stream, cb.p, &b.tlasLevelStart, &root, err, scratch))
			return false;
		if (tlasCursor > tlasReserve) { snprintf(err, 256, "bvh build: TLAS node reservation exceeded"); return false; }
		b.tlasNodeCount = tlasCursor;
	}
	CK(cudaStreamSynchronize(stream));
	return true;
}

bool BvhBuilder::refitTlas(const InstRec *dInstances, cudaStream_t stream, char *err, const MotionWindow *window) {
	Impl &b = *m;
	const int T = 128;
	if (window) b.motion = *window;
	if (b.nTlasPrims <= 0) return true;
	k_prim_bounds_insts<<<(b.nTlasPrims + T - 1) / T, T, 0, stream>>>(dInstances, b.meshBoxes.p, b.tlasIds.p, b.nTlasPrims, b.instBoxes.p, nullptr, nullptr,
																	 b.motion.xnodes, b.motion.keys, b.motion.w0, b.motion.w1);
	for (int l = (int) b.tlasLevelStart.size() - 2; l >= 0; l--) {
		int first = b.tlasLevelStart[l], count = b.tlasLevelStart[l + 1] - first;
		if (count <= 0) continue;
		k_refit_level<<<(count + T - 1) / T, T, 0, stream>>>(b.nodes.p, b.nodeBounds.p, first, count, b.instBoxes.p, b.tlasInst.p);
	}
	CK(cudaGetLastError());
	return true;
}

int BvhBuilder::refitLaunches() const { return 1 + std::max(0, (int) m->tlasLevelStart.size() - 1); }

BvhDev BvhBuilder::device() const {
	BvhDev d;
	d.nodes = m->nodes.p, d.tris = m->tris.p, d.tlasInst = m->tlasInst.p, d.tlasRoot = 0, d.nInstances = m->nInstances;
	d.xnodes = m->motion.xnodes, d.motionKeys = m->motion.keys;
	d.mergedInst = m->mergedInst, d.mergedRoot = m->mergedRoot;
	d.flats = m->flats.p;
	for (int k = 0; k < 3; k++) { // padded: the box only culls, the exact decision is the triangle test
		const float lo = m->mergedBox.lo[k], hi = m->mergedBox.hi[k], pad = 1e-4f * (hi - lo) + 1e-5f * std::max(std::fabs(lo), std::fabs(hi)) + 1e-30f;
		d.rootLo[k] = lo - pad, d.rootHi[k] = hi + pad;
	}
	d.mergedOnly = m->mergedInst >= 0 && m->nTlasPrims == 1; // the pseudo-instance is the only TLAS primitive
	return d;
}
int BvhBuilder::blasRoot(int mesh) const { return m->blasRoots[mesh]; }
int BvhBuilder::mergedRoot() const { return m->mergedRoot; }
int BvhBuilder::mergedTriCount() const { return m->nMergedTris; }
int BvhBuilder::flatCount() const { return m->nFlat; }
int BvhBuilder::triBase(int mesh) const { return m->triBases[mesh]; }
int BvhBuilder::nodeCount() const { return m->totalNodes - (m->nInstances + 2) + m->tlasNodeCount; }
int BvhBuilder::tlasNodeCount() const { return m->tlasNodeCount; }
int BvhBuilder::triCount() const { return m->totalTris; }

} // namespace krr
