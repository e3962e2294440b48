// bvh_build.cu -- device-side acceleration-structure build and refit.
//
// Replaces optixAccelBuild / optixAccelCompact / OPTIX_BUILD_OPERATION_UPDATE (reference
// src/core/device/optix.cpp:143-250 GAS per mesh, :357-398 single-level IAS, :618-669 refit).
// Pipeline (all on the GPU, one stream):
//   1. primitive AABBs + centroid bounds                       (k_prim_bounds_*)
//   2. 63-bit Morton codes of the centroids, radix sort        (k_morton + cub::DeviceRadixSort)
//   3. binary tree by PLOC (parallel locally-ordered clustering, Meister & Bittner 2018): the Morton
//      order is only the SEARCH order -- every cluster looks kPlocRadius neighbours to either side for
//      the partner with the smallest merged surface area, mutual nearest neighbours merge, the cluster
//      array is compacted, repeat until one cluster is left.  (The Karras-2012 radix tree of round 1
//      splits by Morton bits alone: ~60 node + leaf visits per ray on the 20 M-triangle scene.)
//                                                                (k_ploc_nn, k_ploc_merge, k_ploc_compact)
//   4. top-down collapse to 8-wide nodes: a wide node opens the child with the largest surface
//      area until it has 8 children (greedy SAH) and spends slots that are still free on splitting its
//      leaves; leaves hold up to `maxLeaf` <= 3 primitives; children are assigned to OCTANT-ORDERED
//      slots (bvh.cuh); child boxes are quantised to 8 bits against a padded power-of-two frame
//                                                                (k_collapse)
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

// largest singular value of the linear part of a 3x4 transform (|M v| <= sigma |v|): sqrt of the largest eigenvalue of
// A^T A, closed form for a symmetric 3x3 matrix, in double, with a relative margin
KRR_DEV float sigmaMax(const Xf &m) {
	double a[3][3];
	for (int r = 0; r < 3; r++)
		for (int c = 0; c < 3; c++) {
			double v = 0;
			for (int k = 0; k < 3; k++) v += (double) m.m[4 * k + r] * (double) m.m[4 * k + c];
			a[r][c] = v;
		}
	const double p1 = a[0][1] * a[0][1] + a[0][2] * a[0][2] + a[1][2] * a[1][2];
	const double q	= (a[0][0] + a[1][1] + a[2][2]) / 3.0;
	double lmax;
	if (p1 <= 1e-30 * (q * q + 1e-300)) lmax = fmax(a[0][0], fmax(a[1][1], a[2][2]));
	else {
		const double p2 = (a[0][0] - q) * (a[0][0] - q) + (a[1][1] - q) * (a[1][1] - q) + (a[2][2] - q) * (a[2][2] - q) + 2.0 * p1;
		const double p	= sqrt(p2 / 6.0);
		double bm[3][3];
		for (int r = 0; r < 3; r++)
			for (int c = 0; c < 3; c++) bm[r][c] = (a[r][c] - (r == c ? q : 0.0)) / p;
		double det = bm[0][0] * (bm[1][1] * bm[2][2] - bm[1][2] * bm[2][1]) - bm[0][1] * (bm[1][0] * bm[2][2] - bm[1][2] * bm[2][0]) +
					 bm[0][2] * (bm[1][0] * bm[2][1] - bm[1][1] * bm[2][0]);
		const double rr = fmin(1.0, fmax(-1.0, det / 2.0));
		lmax = q + 2.0 * p * cos(acos(rr) / 3.0);
	}
	return (float) (sqrt(fmax(lmax, 0.0)) * 1.0001);
}

// World box of an instance.  Two conservative bounds, intersected: (1) the transformed 8 corners of its mesh's
// object-space box, (2) the box of its mesh's bounding SPHERE (object-space centre c, radius r around it):
// |M v - M c| <= sigma_max(M) r.  The corner box of a rotated object is up to sqrt(3) wider per axis than the
// object, the sphere box does not grow under rotation; compact, roundish meshes get boxes ~1.5x smaller per axis.
// The world sphere itself is kept for a ray / sphere test before an instance is entered (bvh.cuh).
// A MOVING instance (SRT motion chain, motion.cuh) is bounded over the time window [w0, w1] the rays of the frame
// can carry (the camera's shutter interval): corners and sphere centre are evaluated at kMotionSamples + 1 times, all
// positions are united, and the box is padded by a PROVEN bound of how far a point can be from its nearest sample:
// V h / 2, h = the sample spacing, V = an upper bound of the point's speed from the keys (chainMotionBound below).
// (Round 1 padded by the largest second difference of the sampled trajectory: a heuristic that a fast rotation
// between two samples escapes.)
// Speed bound, level by level from the instance's own node to the root.  A level maps v to R(q(f)) (s(f) o v) + T(f)
// with s, T linear in the key parameter f and q = the normalised linear blend of two key quaternions; f is piecewise
// linear in time with slope (n - 1) / (t1 - t0).  With |v| <= B and |dv/dt| <= V below the level:
//   |d/dt (level v)| <= f' (omega sigma B + ds B + dT) + sigma V,      |level v| <= sigma B + Tmax
// sigma = largest |scale component| over the keys, ds = largest |difference of a scale component| and dT = largest
// |difference of the translations| over the key segments, omega = 2 |qb - qa| / min_f |qa + f (qb - qa)| = the
// largest angular speed per unit f (the derivative of a normalised vector is at most |q'| / |q|, and a rotation
// turns by twice the angle its quaternion moves).  A static node contributes its largest singular value and its
// translation.  Everything is evaluated over ALL key segments of a node (a superset of those the window meets).
struct MotionBound { float speed, scaleMax, scaleMin; };
KRR_DEV MotionBound chainMotionBound(const XformNodeRec *__restrict__ nodes, const float *__restrict__ keyPool, int node, float pointNorm) {
	double B = pointNorm, V = 0.0, sMax = 1.0, sMin = 1.0;
	for (int p = node; p >= 0; p = nodes[p].parent) {
		const XformNodeRec &nd = nodes[p];
		if (nd.nKeys >= 2) {
			const float *keys = keyPool + 10 * (size_t) nd.keyOff;
			double sigma = 0.0, sigmaLo = 3.0e38, tMax = 0.0, A = 0.0, C = 0.0;
			for (int k = 0; k < nd.nKeys; k++) {
				const float *a = keys + 10 * k;
				for (int c = 0; c < 3; c++) sigma = fmax(sigma, fabs((double) a[c])), sigmaLo = fmin(sigmaLo, fabs((double) a[c]));
				tMax = fmax(tMax, sqrt((double) a[7] * a[7] + (double) a[8] * a[8] + (double) a[9] * a[9]));
			}
			for (int k = 0; k + 1 < nd.nKeys; k++) {
				const float *a = keys + 10 * k, *b = a + 10;
				double ds = 0.0, dT = 0.0, dd = 0.0, aa = 0.0, ad = 0.0, bb = 0.0;
				for (int c = 0; c < 3; c++) ds = fmax(ds, fabs((double) b[c] - a[c])), dT += ((double) b[7 + c] - a[7 + c]) * ((double) b[7 + c] - a[7 + c]);
				for (int c = 3; c < 7; c++) {
					const double d = (double) b[c] - a[c];
					dd += d * d, aa += (double) a[c] * a[c], ad += (double) a[c] * d, bb += (double) b[c] * b[c];
				}
				// min over f in [0, 1] of |a + f d|^2
				double q2 = fmin(aa, bb);
				if (dd > 0.0) {
					const double f = -ad / dd;
					if (f > 0.0 && f < 1.0) q2 = fmin(q2, fmax(aa - ad * ad / dd, 0.0));
				}
				const double omega = dd > 0.0 ? 2.0 * sqrt(dd) / fmax(sqrt(q2), 1e-30) : 0.0;
				A = fmax(A, omega * sigma + ds), C = fmax(C, sqrt(dT));
			}
			const double fp = (double) (nd.nKeys - 1) / fmax((double) nd.t1 - (double) nd.t0, 1e-30);
			V = fp * (A * B + C) + sigma * V;
			B = sigma * B + tMax;
			sMax *= sigma, sMin *= sigmaLo;
		} else {
			const double sg = sigmaMax(nd.local), sgInv = sigmaMax(nd.localInv);
			V = sg * V;
			B = sg * B + sqrt((double) nd.local.m[3] * nd.local.m[3] + (double) nd.local.m[7] * nd.local.m[7] + (double) nd.local.m[11] * nd.local.m[11]);
			sMax *= sg, sMin *= sgInv > 0.0 ? 1.0 / sgInv : 0.0;
		}
	}
	MotionBound r;
	r.speed = (float) fmin(V * 1.0001, 3.0e38), r.scaleMax = (float) fmin(sMax * 1.0001, 3.0e38), r.scaleMin = (float) (sMin * 0.9999);
	return r;
}
constexpr int kMotionSamples = 8;
__global__ void k_prim_bounds_insts(const InstRec *__restrict__ inst, const Aabb *__restrict__ meshBoxes, const float4 *__restrict__ meshSpheres,
									const int32_t *__restrict__ ids, int n, Aabb *boxesByInst, Aabb *boxesCompact, float4 *spheresByInst, float *cb,
									const XformNodeRec *__restrict__ xnodes, const float *__restrict__ keyPool, float w0, float w1) {
	// ids: the instances that are TLAS primitives (merged instances are not; the merged BLAS is pseudo-
	// instance nInstances).  Boxes are stored by instance id (refit) and, for the build, by TLAS primitive.
	const int slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot >= n) return;
	const int i = ids[slot];
	const Aabb mb	= meshBoxes[inst[i].mesh];
	const float4 ms = meshSpheres[inst[i].mesh];
	const bool useSphere = ms.w < 1.0e30f;
	Aabb b, sb; // corner box, sphere box
	for (int k = 0; k < 3; k++) b.lo[k] = sb.lo[k] = 3.0e38f, b.hi[k] = sb.hi[k] = -3.0e38f;
	auto grow = [&](V3 w) {
		b.lo[0] = fminf(b.lo[0], w.x), b.lo[1] = fminf(b.lo[1], w.y), b.lo[2] = fminf(b.lo[2], w.z);
		b.hi[0] = fmaxf(b.hi[0], w.x), b.hi[1] = fmaxf(b.hi[1], w.y), b.hi[2] = fmaxf(b.hi[2], w.z);
	};
	auto growSphere = [&](V3 c, float r) {
		sb.lo[0] = fminf(sb.lo[0], c.x - r), sb.lo[1] = fminf(sb.lo[1], c.y - r), sb.lo[2] = fminf(sb.lo[2], c.z - r);
		sb.hi[0] = fmaxf(sb.hi[0], c.x + r), sb.hi[1] = fmaxf(sb.hi[1], c.y + r), sb.hi[2] = fmaxf(sb.hi[2], c.z + r);
	};
	auto corner = [&](int c) { return mk3(c & 1 ? mb.hi[0] : mb.lo[0], c & 2 ? mb.hi[1] : mb.lo[1], c & 4 ? mb.hi[2] : mb.lo[2]); };
	const V3 oc = mk3(ms.x, ms.y, ms.z);
	float pad = 0.f, cpad = 0.f, rmin = 3.0e38f;
	if (inst[i].motion >= 0 && xnodes) {
		const int steps = w1 > w0 ? kMotionSamples : 0;
		float cornerNorm = 0.f;
		for (int c = 0; c < 8; c++) cornerNorm = fmaxf(cornerNorm, length(corner(c)));
		const float half = steps ? 0.5f * (w1 - w0) / (float) steps * 1.0001f : 0.f; // farthest a time of the window is from a sample
		const MotionBound mbCorner = chainMotionBound(xnodes, keyPool, inst[i].motion, cornerNorm);
		// (key quaternions that are exactly opposite make the blend pass through zero: the bound is then astronomically large;
		// keep the box finite so that its centre stays a number -- such a box is never culled, which is the right answer)
		pad = fminf(mbCorner.speed * half, 1.0e15f);
		if (useSphere) cpad = fminf(chainMotionBound(xnodes, keyPool, inst[i].motion, length(oc)).speed * half, 1.0e15f);
		for (int j = 0; j <= steps; j++) {
			float t = steps ? w0 + (w1 - w0) * ((float) j / (float) steps) : w0;
			if (j == steps) t = w1;
			Xf m, inv;
			chainXf(xnodes, keyPool, inst[i].motion, t, m, inv);
			for (int c = 0; c < 8; c++) grow(xfPoint(m, corner(c)));
			// the sphere's radius over the WHOLE window from the key scales (its largest / smallest value, not a sampled one)
			if (useSphere) growSphere(xfPoint(m, oc), mbCorner.scaleMax * ms.w);
		}
		if (useSphere) rmin = mbCorner.scaleMin * ms.w;
	} else {
		for (int c = 0; c < 8; c++) grow(xfPoint(inst[i].xf, corner(c)));
		if (useSphere) {
			rmin = sigmaMax(inst[i].xf) * ms.w;
			growSphere(xfPoint(inst[i].xf, oc), rmin);
		}
	}
	// rays are intersected in object space with the rounded inverse transform: pad the world box so
	// that culling stays conservative w.r.t. that round trip
	for (int k = 0; k < 3; k++) {
		float e = 1e-5f * fmaxf(1.f, fmaxf(fabsf(b.lo[k]), fabsf(b.hi[k]))) + 1e-6f * (b.hi[k] - b.lo[k]);
		b.lo[k] -= e + pad, b.hi[k] += e + pad;
		if (useSphere) {
			sb.lo[k] -= e + cpad, sb.hi[k] += e + cpad;
			b.lo[k] = fmaxf(b.lo[k], sb.lo[k]), b.hi[k] = fminf(b.hi[k], sb.hi[k]); // intersection of the two bounds
		}
	}
	// world sphere over the window: around the centre of the sphere box, reaching its farthest face-centre distance
	// (= the largest half extent: the swept sphere is inside the sphere box, and every point of the sphere at a
	// sampled time is within r of a centre that lies in the centres' box)
	float4 ws = make_float4(0.f, 0.f, 0.f, 3.0e38f);
	if (useSphere) {
		const float hx = 0.5f * (sb.hi[0] - sb.lo[0]), hy = 0.5f * (sb.hi[1] - sb.lo[1]), hz = 0.5f * (sb.hi[2] - sb.lo[2]);
		// a sphere of radius r_j has its centre within (h - r_j) of the box centre per axis, so it lies inside the ball
		// of radius f(r_j) = |h - r_j| + r_j around the box centre; f decreases with r, so the SMALLEST radius of the
		// window gives the bound for all of them
		const float rs = fminf(rmin, fminf(hx, fminf(hy, hz)));
		const float ex = hx - rs, ey = hy - rs, ez = hz - rs;
		ws = make_float4(0.5f * (sb.hi[0] + sb.lo[0]), 0.5f * (sb.hi[1] + sb.lo[1]), 0.5f * (sb.hi[2] + sb.lo[2]), sqrtf(ex * ex + ey * ey + ez * ez) + rs);
	}
	boxesByInst[i] = b;
	spheresByInst[i] = ws;
	if (boxesCompact) boxesCompact[slot] = b;
	if (cb)
		for (int k = 0; k < 3; k++) {
			float c = 0.5f * (b.lo[k] + b.hi[k]);
			atomicMinF(cb + k, c), atomicMaxF(cb + 3 + k, c);
		}
}

// bounding sphere of a mesh around the centre of its box: radius = the farthest triangle vertex (one block)
__global__ void k_mesh_sphere(const float *__restrict__ pos, const int32_t *__restrict__ idx, int nTri, const Aabb *__restrict__ box, float4 *out) {
	__shared__ float red[256];
	const float cx = 0.5f * (box->lo[0] + box->hi[0]), cy = 0.5f * (box->lo[1] + box->hi[1]), cz = 0.5f * (box->lo[2] + box->hi[2]);
	float r2 = 0.f;
	for (int i = threadIdx.x; i < 3 * nTri; i += blockDim.x) {
		const int v = idx[i];
		const float dx = pos[3 * v] - cx, dy = pos[3 * v + 1] - cy, dz = pos[3 * v + 2] - cz;
		r2 = fmaxf(r2, dx * dx + dy * dy + dz * dz);
	}
	red[threadIdx.x] = r2;
	__syncthreads();
	for (int s = blockDim.x / 2; s > 0; s >>= 1) {
		if (threadIdx.x < s) red[threadIdx.x] = fmaxf(red[threadIdx.x], red[threadIdx.x + s]);
		__syncthreads();
	}
	if (threadIdx.x == 0) *out = make_float4(cx, cy, cz, sqrtf(red[0]) * 1.00001f + 1e-30f);
}

// merged BLAS: triangle boxes of every merged instance (identity transform: object space == world
// space), concatenated; pairs[t] = (group slot, local primitive).  blockIdx.y strides over the merged
// instances, blockIdx.x over 256-triangle chunks.
struct MergedSrc {
	int32_t posOff, idxOff, nTri, inst, outOff;
	int32_t identity; // 0: flattened instance, the boxes bound the transformed triangles
	float xf[12];
};
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
				float p[3] = {pos[3 * v], pos[3 * v + 1], pos[3 * v + 2]};
				if (!ms.identity) {
					const float x = p[0], y = p[1], z = p[2];
					for (int k = 0; k < 3; k++) p[k] = ms.xf[4 * k] * x + ms.xf[4 * k + 1] * y + ms.xf[4 * k + 2] * z + ms.xf[4 * k + 3];
				}
				for (int k = 0; k < 3; k++) b.lo[k] = fminf(b.lo[k], p[k]), b.hi[k] = fmaxf(b.hi[k], p[k]);
			}
			if (!ms.identity) // the leaf test runs in object space through the ROUNDED inverse: pad the world box for that round trip
				for (int k = 0; k < 3; k++) {
					const float e = 1e-5f * fmaxf(1.f, fmaxf(fabsf(b.lo[k]), fabsf(b.hi[k]))) + 1e-6f * (b.hi[k] - b.lo[k]);
					b.lo[k] -= e, b.hi[k] += e;
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

// binary tree over the sorted primitives: leaves are nodes 0..n-1 (leaf i = sorted primitive i), internal nodes
// n..2n-2 in creation order (the root is created last)
struct BinTree {
	int32_t *left, *right; // per node id (internal ids only are written)
	int32_t *count;		   // primitives below each node
	Aabb *bounds;		   // per node (2n-1)
};

KRR_DEV float halfArea(const Aabb &b) {
	float dx = b.hi[0] - b.lo[0], dy = b.hi[1] - b.lo[1], dz = b.hi[2] - b.lo[2];
	return dx * dy + dy * dz + dz * dx;
}

// ---- PLOC ----
constexpr int kPlocRadius = 16, kPlocBlock = 256;

__global__ void k_ploc_init(const Aabb *__restrict__ primBoxes, const uint32_t *__restrict__ sortedIdx, int n, BinTree t, int32_t *cid, Aabb *cbox) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const Aabb b = primBoxes[sortedIdx[i]];
	t.bounds[i] = b, t.count[i] = 1;
	cid[i] = i, cbox[i] = b;
}

// nearest neighbour (smallest surface area of the union) of every cluster within kPlocRadius positions; ties go
// to the smaller index, which makes the lexicographically smallest minimal pair mutual: every pass merges
__global__ void __launch_bounds__(kPlocBlock) k_ploc_nn(const Aabb *__restrict__ cbox, int n, int32_t *nn) {
	constexpr int W = kPlocBlock + 2 * kPlocRadius;
	__shared__ float lo[3][W], hi[3][W];
	const int base = blockIdx.x * kPlocBlock - kPlocRadius;
	for (int s = threadIdx.x; s < W; s += kPlocBlock) {
		const int g = base + s;
		if (g >= 0 && g < n) {
			const Aabb b = cbox[g];
			for (int k = 0; k < 3; k++) lo[k][s] = b.lo[k], hi[k][s] = b.hi[k];
		}
	}
	__syncthreads();
	const int i = blockIdx.x * kPlocBlock + threadIdx.x;
	if (i >= n) return;
	const int me = threadIdx.x + kPlocRadius;
	const float l0 = lo[0][me], l1 = lo[1][me], l2 = lo[2][me], h0 = hi[0][me], h1 = hi[1][me], h2 = hi[2][me];
	float bestA = 3.0e38f;
	int best	= -1;
	for (int r = -kPlocRadius; r <= kPlocRadius; r++) {
		const int g = i + r;
		if (r == 0 || g < 0 || g >= n) continue;
		const int s	   = me + r;
		const float dx = fmaxf(h0, hi[0][s]) - fminf(l0, lo[0][s]), dy = fmaxf(h1, hi[1][s]) - fminf(l1, lo[1][s]),
					dz = fmaxf(h2, hi[2][s]) - fminf(l2, lo[2][s]);
		const float a  = dx * dy + dy * dz + dz * dx;
		if (a < bestA) bestA = a, best = g; // ascending g: the first minimum is the smallest index
	}
	nn[i] = best;
}

// mutual nearest neighbours merge into a new node that takes the place of the lower one
__global__ void k_ploc_merge(const int32_t *__restrict__ cidIn, const Aabb *__restrict__ cboxIn, const int32_t *__restrict__ nn, int n,
							 int32_t *nodeCounter, BinTree t, int32_t *cidTmp, Aabb *cboxTmp, int32_t *valid) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int j		  = nn[i];
	const bool mutual = j >= 0 && nn[j] == i;
	if (mutual && i > j) { valid[i] = 0; return; }
	int id = cidIn[i];
	Aabb b = cboxIn[i];
	if (mutual) {
		const Aabb o = cboxIn[j];
		for (int k = 0; k < 3; k++) b.lo[k] = fminf(b.lo[k], o.lo[k]), b.hi[k] = fmaxf(b.hi[k], o.hi[k]);
		const int l = id, r = cidIn[j];
		id			= atomicAdd(nodeCounter, 1);
		t.left[id] = l, t.right[id] = r;
		t.count[id]	 = t.count[l] + t.count[r];
		t.bounds[id] = b;
	}
	cidTmp[i] = id, cboxTmp[i] = b, valid[i] = 1;
}

__global__ void k_ploc_compact(const int32_t *__restrict__ cidTmp, const Aabb *__restrict__ cboxTmp, const int32_t *__restrict__ valid,
							   const int32_t *__restrict__ pos, int n, int32_t *cidOut, Aabb *cboxOut, int32_t *nOut) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	if (valid[i]) cidOut[pos[i]] = cidTmp[i], cboxOut[pos[i]] = cboxTmp[i];
	if (i == n - 1) *nOut = pos[i] + valid[i];
}

// The last kPlocTail clusters are finished by ONE block (the passes of a small array are launch-bound: ~25
// passes x 4 launches + a host round trip each, per BLAS)
constexpr int kPlocTail = 1024;
__global__ void __launch_bounds__(kPlocTail) k_ploc_tail(int32_t *cid, Aabb *cbox, int n, int32_t *nodeCounter, BinTree t) {
	__shared__ float lo[3][kPlocTail], hi[3][kPlocTail];
	__shared__ int32_t id[kPlocTail], nnS[kPlocTail], posS[kPlocTail];
	__shared__ int32_t warpSum[32];
	__shared__ int32_t nNow;
	const int i = threadIdx.x;
	if (i < n) {
		const Aabb b = cbox[i];
		for (int k = 0; k < 3; k++) lo[k][i] = b.lo[k], hi[k][i] = b.hi[k];
		id[i] = cid[i];
	}
	if (i == 0) nNow = n;
	__syncthreads();
	while (true) {
		const int m = nNow;
		if (m <= 1) break;
		int best = -1;
		if (i < m) {
			float bestA = 3.0e38f;
			for (int r = -kPlocRadius; r <= kPlocRadius; r++) {
				const int g = i + r;
				if (r == 0 || g < 0 || g >= m) continue;
				const float dx = fmaxf(hi[0][i], hi[0][g]) - fminf(lo[0][i], lo[0][g]), dy = fmaxf(hi[1][i], hi[1][g]) - fminf(lo[1][i], lo[1][g]),
							dz = fmaxf(hi[2][i], hi[2][g]) - fminf(lo[2][i], lo[2][g]);
				const float a  = dx * dy + dy * dz + dz * dx;
				if (a < bestA) bestA = a, best = g;
			}
			nnS[i] = best;
		}
		__syncthreads();
		int myId = 0, v = 0;
		float b[6] = {0, 0, 0, 0, 0, 0};
		if (i < m) {
			const int j		  = best;
			const bool mutual = j >= 0 && nnS[j] == i;
			v				  = !(mutual && i > j);
			myId			  = id[i];
			for (int k = 0; k < 3; k++) b[k] = lo[k][i], b[3 + k] = hi[k][i];
			if (mutual && i < j) {
				for (int k = 0; k < 3; k++) b[k] = fminf(b[k], lo[k][j]), b[3 + k] = fmaxf(b[3 + k], hi[k][j]);
				const int l = myId, r = id[j];
				myId		= atomicAdd(nodeCounter, 1);
				t.left[myId] = l, t.right[myId] = r;
				t.count[myId] = t.count[l] + t.count[r];
				Aabb nb;
				for (int k = 0; k < 3; k++) nb.lo[k] = b[k], nb.hi[k] = b[3 + k];
				t.bounds[myId] = nb;
			}
		}
		// block-wide exclusive scan of v
		const unsigned bal = __ballot_sync(0xffffffffu, v);
		const int lane = i & 31, warp = i >> 5;
		if (lane == 0) warpSum[warp] = __popc(bal);
		__syncthreads();
		if (warp == 0) {
			int s = warpSum[lane], incl = s;
			for (int d = 1; d < 32; d <<= 1) {
				const int o = __shfl_up_sync(0xffffffffu, incl, d);
				if (lane >= d) incl += o;
			}
			warpSum[lane] = incl - s;
			if (lane == 31) nNow = incl;
		}
		__syncthreads();
		posS[i] = warpSum[warp] + __popc(bal & ((1u << lane) - 1u));
		__syncthreads(); // everyone has read lo / hi / id of the old array
		if (v && i < m) {
			const int p = posS[i];
			for (int k = 0; k < 3; k++) lo[k][p] = b[k], hi[k][p] = b[3 + k];
			id[p] = myId;
		}
		__syncthreads();
	}
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
	KRR_DEV void clear(uint32_t, bool) const {}
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
	KRR_DEV void clear(uint32_t, bool) const {}
};
struct InstWriter {
	int32_t *tlasInst;
	const int32_t *ids; // TLAS primitive -> instance id
	KRR_DEV void operator()(uint32_t slot, uint32_t prim) const { tlasInst[slot] = ids[prim]; }
	KRR_DEV void clear(uint32_t slot, bool any) const { if (any) tlasInst[slot] = 0; } // slots without an instance
};

// flat BLAS (<= flatMax triangles): the leaf payload in primitive order, no nodes
template <typename Writer> __global__ void k_write_flat(int n, uint32_t base, Writer writer) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) writer(base + i, (uint32_t) i);
}

// One thread = one wide node.  `item.bin` is the binary node whose subtree the wide node covers.
// TLAS = true: one instance per leaf child, instance ids stored slot-major (8 per node).
template <bool TLAS, typename Writer>
__global__ void k_collapse(const WorkItem *__restrict__ in, int nIn, WorkItem *out, int32_t *nOut, BinTree t, int n,
						   const uint32_t *__restrict__ sortedIdx, int maxLeaf, CollapseOut co, Writer writer) {
	int w = blockIdx.x * blockDim.x + threadIdx.x;
	if (w >= nIn) return;
	WorkItem item = in[w];
	int c[8], nc = 0;
	if (item.bin < n) c[nc++] = item.bin; // the whole tree is one primitive
	else { c[nc++] = t.left[item.bin]; c[nc++] = t.right[item.bin]; }
	// greedy SAH: open the child with the largest surface area -- first among the subtrees that are too big
	// to be a leaf, then (slots are free: an empty slot costs the same 80-byte node) among the leaves that
	// still hold more than one primitive, which gives every triangle the tightest box the node can afford
	for (int pass = 0; pass < 2; pass++) {
		const int limit = pass == 0 ? maxLeaf : 1;
		while (nc < 8) {
			int best = -1;
			float bestA = -1;
			for (int j = 0; j < nc; j++)
				if (t.count[c[j]] > limit) {
					float a = halfArea(t.bounds[c[j]]);
					if (a > bestA) bestA = a, best = j;
				}
			if (best < 0) break;
			int node = c[best];
			c[best]	 = t.left[node];
			c[nc++]	 = t.right[node];
		}
	}
	const Aabb nb = t.bounds[item.bin];
	float o[3];
	uint32_t e[3];
	makeFrame(nb, o, e);
	// octant-ordered slots: slot s "lies" in direction ((s&4 ? + : -), (s&2 ? + : -), (s&1 ? + : -)) of the node
	// centre; children are assigned greedily by the largest projection of (child centre - node centre) on the
	// slot direction, so that slot ^ (ray octant) orders the children front to back (bvh.cuh)
	int slotOf[8], childAt[8];
	{
		float cx[8], cy[8], cz[8];
		const float mx = 0.5f * (nb.lo[0] + nb.hi[0]), my = 0.5f * (nb.lo[1] + nb.hi[1]), mz = 0.5f * (nb.lo[2] + nb.hi[2]);
		for (int j = 0; j < nc; j++) {
			const Aabb cb = t.bounds[c[j]];
			cx[j] = 0.5f * (cb.lo[0] + cb.hi[0]) - mx, cy[j] = 0.5f * (cb.lo[1] + cb.hi[1]) - my, cz[j] = 0.5f * (cb.lo[2] + cb.hi[2]) - mz;
		}
		for (int s = 0; s < 8; s++) childAt[s] = -1;
		for (int j = 0; j < nc; j++) slotOf[j] = -1;
		for (int r = 0; r < nc; r++) {
			float bestC = -3.0e38f;
			int bj = -1, bs = -1;
			for (int j = 0; j < nc; j++) {
				if (slotOf[j] >= 0) continue;
				for (int s = 0; s < 8; s++) {
					if (childAt[s] >= 0) continue;
					const float cost = (s & 4 ? cx[j] : -cx[j]) + (s & 2 ? cy[j] : -cy[j]) + (s & 1 ? cz[j] : -cz[j]);
					if (cost > bestC) bestC = cost, bj = j, bs = s;
				}
			}
			slotOf[bj] = bs, childAt[bs] = bj;
		}
	}
	int nInternal = 0, nPrims = 0;
	for (int j = 0; j < nc; j++) {
		if (t.count[c[j]] <= maxLeaf) nPrims += t.count[c[j]];
		else nInternal++;
	}
	if (TLAS && nPrims) nPrims = 8; // slot-major instance ids
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
	for (int s = 0; s < 8; s++) { // internal children and leaf triangles are numbered in slot order
		node.meta[s] = 0;
		for (int k = 0; k < 3; k++) node.qlo[k][s] = 255, node.qhi[k][s] = 0;
		if (TLAS) writer.clear(co.primBase + primBase + s, nPrims != 0);
		const int j = childAt[s];
		if (j < 0) continue;
		uint8_t ql[3], qh[3];
		quantize(t.bounds[c[j]], o, e, ql, qh);
		for (int k = 0; k < 3; k++) node.qlo[k][s] = ql[k], node.qhi[k][s] = qh[k];
		const int cnt = t.count[c[j]];
		if (cnt <= maxLeaf) {
			// the primitives below c[j] (at most 3): depth-first walk of its little subtree
			int stk[4], sp = 0, k = 0;
			stk[sp++] = c[j];
			while (sp) {
				const int x = stk[--sp];
				if (x < n) { writer(co.primBase + primBase + (TLAS ? s : po + k), sortedIdx[x]); k++; }
				else stk[sp++] = t.right[x], stk[sp++] = t.left[x];
			}
			node.meta[s] = TLAS ? (uint8_t) (0x20 | s) : (uint8_t) ((((1u << cnt) - 1u) << 5) | (uint32_t) po);
			po += cnt;
		} else {
			node.imask |= (uint8_t) (1u << s);
			node.meta[s] = (uint8_t) (0x20 | (24 + s));
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
		used[j] = node.meta[j] != 0;
		if (!used[j]) continue;
		if ((node.imask >> j) & 1) cb[j] = nodeBounds[node.childBase + __popc(node.imask & ((1u << j) - 1))];
		else cb[j] = primBoxes[tlasInst[node.primBase + j]]; // one instance per TLAS leaf, slot-major
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
	DevBuf<int32_t> left, right, count;
	DevBuf<Aabb> bounds;
	DevBuf<int32_t> cid[2], cidTmp, nn, valid, pos, plocCounters; // PLOC cluster arrays (ping-pong), scratch
	DevBuf<Aabb> cbox[2], cboxTmp;
	DevBuf<WorkItem> q0, q1;
	DevBuf<int32_t> qCount;
	DevBuf<unsigned char> tmp, tmpScan;
};

} // namespace

struct BvhBuilder::Impl {
	DevBuf<Node8> nodes;
	DevBuf<Aabb> nodeBounds;
	DevBuf<BvhTri> tris;
	DevBuf<int32_t> tlasInst;
	DevBuf<Aabb> meshBoxes, instBoxes;
	DevBuf<float4> meshSpheres, instSpheres;
	DevBuf<int32_t> counters, tlasIds;
	DevBuf<int2> flats;
	int nTlasPrims = 0, mergedRoot = -1, mergedInst = -1, nMergedTris = 0, nFlat = 0, mergedXf = 0;
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
template <bool TLAS, typename Writer>
bool buildTree(const Aabb *boxes, int n, int maxLeaf, Node8 *nodePool, Aabb *boundsPool, int32_t *counters, int &nodeCursor,
			   int &primCursor, Writer writer, cudaStream_t stream, float *cb, std::vector<int> *levelStart, int *root, char *err,
			   TreeScratch &sc) {
	const int T = 256;
	if (!sc.keys.ensure(n) || !sc.keysSorted.ensure(n) || !sc.vals.ensure(n) || !sc.valsSorted.ensure(n) || !sc.left.ensure(2 * n) ||
		!sc.right.ensure(2 * n) || !sc.count.ensure(2 * n) || !sc.bounds.ensure(2 * n) || !sc.cid[0].ensure(n) || !sc.cid[1].ensure(n) ||
		!sc.cidTmp.ensure(n) || !sc.nn.ensure(n) || !sc.valid.ensure(n) || !sc.pos.ensure(n) || !sc.plocCounters.ensure(2) ||
		!sc.cbox[0].ensure(n) || !sc.cbox[1].ensure(n) || !sc.cboxTmp.ensure(n) || !sc.q0.ensure(n + 1) || !sc.q1.ensure(n + 1) ||
		!sc.qCount.ensure(1)) {
		snprintf(err, 256, "bvh build: out of device memory for %d primitives", n);
		return false;
	}
	k_morton<<<(n + T - 1) / T, T, 0, stream>>>(boxes, n, cb, sc.keys.p, sc.vals.p);
	size_t tmpBytes = 0;
	cub::DeviceRadixSort::SortPairs(nullptr, tmpBytes, sc.keys.p, sc.keysSorted.p, sc.vals.p, sc.valsSorted.p, n, 0, 63, stream);
	if (!sc.tmp.ensure(tmpBytes)) { snprintf(err, 256, "bvh build: sort scratch alloc failed"); return false; }
	CK(cub::DeviceRadixSort::SortPairs(sc.tmp.p, tmpBytes, sc.keys.p, sc.keysSorted.p, sc.vals.p, sc.valsSorted.p, n, 0, 63, stream));
	// ---- binary tree: PLOC over the Morton order ----
	BinTree t{sc.left.p, sc.right.p, sc.count.p, sc.bounds.p};
	int32_t *nodeCounter = sc.plocCounters.p, *nOutDev = sc.plocCounters.p + 1;
	{
		const int32_t init[2] = {n, n}; // next free node id (leaves are 0..n-1), cluster count
		CK(cudaMemcpyAsync(sc.plocCounters.p, init, 8, cudaMemcpyHostToDevice, stream));
	}
	k_ploc_init<<<(n + T - 1) / T, T, 0, stream>>>(boxes, sc.valsSorted.p, n, t, sc.cid[0].p, sc.cbox[0].p);
	int m = n, cur = 0;
	size_t scanBytes = 0;
	cub::DeviceScan::ExclusiveSum(nullptr, scanBytes, sc.valid.p, sc.pos.p, n, stream);
	if (!sc.tmpScan.ensure(scanBytes)) { snprintf(err, 256, "bvh build: scan scratch alloc failed"); return false; }
	while (m > kPlocTail) {
		const int g = (m + kPlocBlock - 1) / kPlocBlock;
		k_ploc_nn<<<g, kPlocBlock, 0, stream>>>(sc.cbox[cur].p, m, sc.nn.p);
		k_ploc_merge<<<g, kPlocBlock, 0, stream>>>(sc.cid[cur].p, sc.cbox[cur].p, sc.nn.p, m, nodeCounter, t, sc.cidTmp.p, sc.cboxTmp.p, sc.valid.p);
		CK(cub::DeviceScan::ExclusiveSum(sc.tmpScan.p, scanBytes, sc.valid.p, sc.pos.p, m, stream));
		k_ploc_compact<<<g, kPlocBlock, 0, stream>>>(sc.cidTmp.p, sc.cboxTmp.p, sc.valid.p, sc.pos.p, m, sc.cid[cur ^ 1].p, sc.cbox[cur ^ 1].p, nOutDev);
		int mNew = 0;
		CK(cudaMemcpyAsync(&mNew, nOutDev, 4, cudaMemcpyDeviceToHost, stream));
		CK(cudaStreamSynchronize(stream));
		if (mNew >= m || mNew < 1) { snprintf(err, 256, "bvh build: PLOC made no progress (%d -> %d clusters)", m, mNew); return false; }
		m = mNew, cur ^= 1;
	}
	if (m > 1) k_ploc_tail<<<1, kPlocTail, 0, stream>>>(sc.cid[cur].p, sc.cbox[cur].p, m, nodeCounter, t);
	const int binRoot = n == 1 ? 0 : 2 * n - 2; // the last merge creates the root
	// ---- collapse, level by level ----
	int32_t cnt[2] = {1, 0}; // node 0 of this tree is the root
	CK(cudaMemcpyAsync(counters, cnt, 8, cudaMemcpyHostToDevice, stream));
	WorkItem rootItem{binRoot, 0};
	CK(cudaMemcpyAsync(sc.q0.p, &rootItem, sizeof rootItem, cudaMemcpyHostToDevice, stream));
	CollapseOut co{nodePool, boundsPool, counters, (uint32_t) nodeCursor, (uint32_t) primCursor};
	int nIn = 1, levelFirst = 0;
	WorkItem *qin = sc.q0.p, *qout = sc.q1.p;
	*root = nodeCursor;
	while (nIn > 0) {
		if (levelStart) levelStart->push_back(nodeCursor + levelFirst);
		CK(cudaMemsetAsync(sc.qCount.p, 0, 4, stream));
		k_collapse<TLAS><<<(nIn + 127) / 128, 128, 0, stream>>>(qin, nIn, qout, sc.qCount.p, t, n, sc.valsSorted.p, maxLeaf, co, writer);
		int nOut = 0;
		CK(cudaMemcpyAsync(&nOut, sc.qCount.p, 4, cudaMemcpyDeviceToHost, stream));
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
	// triangles per leaf child: the hit mask of a node has 24 triangle bits = 8 leaf children x 3 triangles
	const int maxLeaf = std::min(std::max(getenv("KRR_BVH_MAX_LEAF") ? atoi(getenv("KRR_BVH_MAX_LEAF")) : 3, 1), 3);
	std::vector<int2> flats;
	b.nMeshes = nMeshes, b.nInstances = nInstances, b.mergedXf = 0;
	// which instances go into the merged world-space BLAS, which meshes still need a BLAS of their own
	std::vector<MergedSrc> msrc;
	std::vector<char> meshNeedsBlas(nMeshes, 0);
	std::vector<int32_t> tlasIds;
	size_t mergedTris = 0;
	for (int i = 0; i < nInstances; i++) {
		const MeshRec &mr = hMeshes[hInstances[i].mesh];
		if (hMerge && hMerge[i]) {
			MergedSrc ms{mr.posOff, mr.idxOff, mr.nTri, i, (int32_t) mergedTris, 1, {}};
			static const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
			for (int k = 0; k < 12; k++) {
				ms.xf[k] = hInstances[i].xf.m[k];
				if (!(hInstances[i].xf.m[k] == I[k]) || !(hInstances[i].inv.m[k] == I[k])) ms.identity = 0;
			}
			if (!ms.identity) b.mergedXf = 1;
			msrc.push_back(ms);
			mergedTris += mr.nTri;
		} else {
			meshNeedsBlas[hInstances[i].mesh] = 1;
			tlasIds.push_back(i);
		}
	}
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
	// triangle and node indices are 32-bit (Node8::primBase / childBase, int cursors): refuse pools that do not fit
	// instead of wrapping (2^31 triangles = 96 GB of leaf triangles: within reach of a 180 GB device)
	if (totalTris + (size_t) nInstances + 16 > 0x7fffffffull) {
		snprintf(err, 256, "bvh build: %zu pooled triangles exceed the 31-bit triangle / node index range", totalTris);
		return false;
	}
	const int tlasReserve = nInstances + 2;
	const size_t nodeCap  = totalTris + (size_t) nBlas + (size_t) tlasReserve + 8;
	if (!b.nodes.alloc(nodeCap) || !b.nodeBounds.alloc(nodeCap) || !b.tris.alloc(totalTris) || !b.tlasInst.alloc(8 * (size_t) (nInstances + 2) + 8) ||
		!b.meshBoxes.alloc(nMeshes + 1) || !b.instBoxes.alloc(nInstances + 1) || !b.meshSpheres.alloc(nMeshes + 1) || !b.instSpheres.alloc(nInstances + 1) || !b.counters.alloc(2) || !b.tlasIds.alloc(tlasIds.size())) {
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
		k_mesh_sphere<<<1, 256, 0, stream>>>(dPositions + 3 * (size_t) mr.posOff, dIndices + 3 * (size_t) mr.idxOff, mr.nTri, b.meshBoxes.p + i, b.meshSpheres.p + i);
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
		if (!buildTree<false>(primBoxes.p, mr.nTri, maxLeaf, b.nodes.p, b.nodeBounds.p, b.counters.p, nodeCursor, primCursor, wr, stream, cb.p, nullptr, &root, err, scratch))
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
		{ // the merged BLAS is bounded by its box alone
			const float4 none = make_float4(0.f, 0.f, 0.f, 3.0e38f);
			CK(cudaMemcpyAsync(b.meshSpheres.p + nMeshes, &none, sizeof none, cudaMemcpyHostToDevice, stream));
		}
		MergedTriWriter wr{dPositions, dIndices, dsrc.p, pairs.p, b.tris.p};
		if ((int) mergedTris <= flatMax) {
			k_write_flat<<<((int) mergedTris + T - 1) / T, T, 0, stream>>>((int) mergedTris, (uint32_t) primCursor, wr);
			b.mergedRoot = (int32_t) (kFlatFlag | (uint32_t) flats.size());
			flats.push_back(make_int2(primCursor, (int) mergedTris));
			primCursor += (int) mergedTris;
			CK(cudaStreamSynchronize(stream)); // dsrc / pairs die with this scope
		} else {
			int root = 0;
			if (!buildTree<false>(primBoxes.p, (int) mergedTris, maxLeaf, b.nodes.p, b.nodeBounds.p, b.counters.p, nodeCursor, primCursor, wr, stream, cb.p, nullptr, &root, err, scratch))
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
		k_prim_bounds_insts<<<(b.nTlasPrims + T - 1) / T, T, 0, stream>>>(dInstances, b.meshBoxes.p, b.meshSpheres.p, b.tlasIds.p, b.nTlasPrims, b.instBoxes.p, primBoxes.p, b.instSpheres.p,
																		 cb.p, motion.xnodes, motion.keys, motion.w0, motion.w1);
		int tlasCursor = 0, tlasPrims = 0, root = 0;
		InstWriter iw{b.tlasInst.p, b.tlasIds.p};
		if (!buildTree<true>(primBoxes.p, b.nTlasPrims, 1, b.nodes.p, b.nodeBounds.p, b.counters.p, tlasCursor, tlasPrims, iw, stream, cb.p, &b.tlasLevelStart, &root, err, scratch))
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
	k_prim_bounds_insts<<<(b.nTlasPrims + T - 1) / T, T, 0, stream>>>(dInstances, b.meshBoxes.p, b.meshSpheres.p, b.tlasIds.p, b.nTlasPrims, b.instBoxes.p, nullptr, b.instSpheres.p, nullptr,
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
	d.xnodes = m->motion.xnodes, d.motionKeys = m->motion.keys, d.motionFlat = nullptr; // (set by the pass: api.cu makeWavefront)
	d.mergedInst = m->mergedInst, d.mergedRoot = m->mergedRoot, d.mergedXf = m->mergedXf;
	d.flats = m->flats.p;
	d.instSphere = m->instSpheres.p;
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
