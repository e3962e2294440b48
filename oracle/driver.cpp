/* driver.cpp -- TEST INFRASTRUCTURE ONLY (checker; never shipped, never on the product path).
 *
 * CPU restatement of the reference's WavefrontPathTracer stage bodies, run depth-first per pixel
 * (each pixel's RNG stream and accumulator are private, so depth-first execution yields the same
 * per-pixel results as the wavefront order -- SURVEY.md section 8d):
 *   beginFrame            src/render/wavefront/integrator.cpp:205-221
 *   generateCameraRays    integrator.cpp:166-179
 *   Closest CH/AH/MS      src/render/wavefront/device.cu:43-81
 *   prepareSurfaceInter.  src/render/shading.h:115-226   (getHitInfo :78-87)
 *   handleHit / Miss      integrator.cpp:78-108
 *   generateScatterRays   integrator.cpp:110-164
 *   Shadow RG/AH/MS       device.cu:83-100
 *   resolve + film write  integrator.cpp:257-266, src/core/device/cuda.h:33-45
 *   scene upload          src/core/device/scene.cpp:28-173, src/core/mesh.cpp:16-63
 * All leaf maths (sampler, camera, spectra, BSDFs, lights) comes through oracle_leaf.h, i.e. from
 * the reference's own classes (backend_ref) or from the plain-C++ port (backend_port).
 *
 * Ray/triangle intersection and traversal are NOT in the reference (closed NVIDIA OptiX,
 * device.cu:15); this build specifies its own routine (see DESIGN.md "Intersection spec") and the
 * CUDA kernels implement the same arithmetic, so first-hit ids are comparable bit-for-bit.
 */
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <limits>
#include <vector>
#include <atomic>
#include <thread>

#include "driver.h"
#include "oracle_leaf.h"

namespace {

struct V3 {
	float x, y, z;
	float operator[](int i) const { return (&x)[i]; }
	float &operator[](int i) { return (&x)[i]; }
};
inline V3 mk(float x, float y, float z) { return V3{x, y, z}; }
inline V3 mk(const float *p) { return V3{p[0], p[1], p[2]}; }
inline V3 operator+(V3 a, V3 b) { return mk(a.x + b.x, a.y + b.y, a.z + b.z); }
inline V3 operator-(V3 a, V3 b) { return mk(a.x - b.x, a.y - b.y, a.z - b.z); }
inline V3 operator-(V3 a) { return mk(-a.x, -a.y, -a.z); }
inline V3 operator*(V3 a, float s) { return mk(a.x * s, a.y * s, a.z * s); }
inline V3 operator*(float s, V3 a) { return mk(a.x * s, a.y * s, a.z * s); }
inline V3 operator/(V3 a, float s) { return mk(a.x / s, a.y / s, a.z / s); }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) {
	return mk(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
inline float length(V3 a) { return std::sqrt(dot(a, a)); }
inline V3 normalize(V3 a) {
	float z = dot(a, a);
	return z > 0 ? a / std::sqrt(z) : a; /* Eigen normalized(): unchanged when the norm is 0 */
}
inline void st(float *o, V3 v) { o[0] = v.x, o[1] = v.y, o[2] = v.z; }

struct Spec {
	float v[4];
	float operator[](int i) const { return v[i]; }
	float &operator[](int i) { return v[i]; }
};
inline Spec sconst(float c) { return Spec{{c, c, c, c}}; }
inline Spec operator*(Spec a, Spec b) { Spec r; for (int i = 0; i < 4; i++) r.v[i] = a.v[i] * b.v[i]; return r; }
inline Spec operator*(Spec a, float s) { Spec r; for (int i = 0; i < 4; i++) r.v[i] = a.v[i] * s; return r; }
inline Spec operator/(Spec a, float s) { Spec r; for (int i = 0; i < 4; i++) r.v[i] = a.v[i] / s; return r; }
inline Spec operator+(Spec a, Spec b) { Spec r; for (int i = 0; i < 4; i++) r.v[i] = a.v[i] + b.v[i]; return r; }
inline bool any(Spec a) { return a.v[0] != 0 || a.v[1] != 0 || a.v[2] != 0 || a.v[3] != 0; }
/* Eigen mean() of 4 floats: redux order (a0+a1)+(a2+a3) (packet/halving reduction), then /4 */
inline float mean(Spec a) { return ((a.v[0] + a.v[1]) + (a.v[2] + a.v[3])) / 4; }

/* ---- 3x4 affine helpers ---- */
struct Xf { float m[12]; };
inline V3 xfPoint(const Xf &t, V3 p) {
	return mk(t.m[0] * p.x + t.m[1] * p.y + t.m[2] * p.z + t.m[3],
			  t.m[4] * p.x + t.m[5] * p.y + t.m[6] * p.z + t.m[7],
			  t.m[8] * p.x + t.m[9] * p.y + t.m[10] * p.z + t.m[11]);
}
inline V3 xfVector(const Xf &t, V3 p) {
	return mk(t.m[0] * p.x + t.m[1] * p.y + t.m[2] * p.z, t.m[4] * p.x + t.m[5] * p.y + t.m[6] * p.z,
			  t.m[8] * p.x + t.m[9] * p.y + t.m[10] * p.z);
}
/* (M^-1)^T * n, given inv = M^-1 (Transformation::transposedInverse, raytracing.h:91-93) */
inline V3 xfNormal(const Xf &inv, V3 n) {
	return mk(inv.m[0] * n.x + inv.m[4] * n.y + inv.m[8] * n.z, inv.m[1] * n.x + inv.m[5] * n.y + inv.m[9] * n.z,
			  inv.m[2] * n.x + inv.m[6] * n.y + inv.m[10] * n.z);
}
/* Affine inverse, cofactors in double, rounded once to float.  DESIGN.md "Intersection spec":
 * the product's scene upload uses the identical routine so object-space rays match bit-for-bit. */
Xf xfInverse(const Xf &t) {
	double a = t.m[0], b = t.m[1], c = t.m[2], d = t.m[4], e = t.m[5], f = t.m[6], g = t.m[8], h = t.m[9], i = t.m[10];
	double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
	double det = a * A + b * B + c * C;
	double id  = 1.0 / det;
	double r[9] = {A * id, -(b * i - c * h) * id, (b * f - c * e) * id,
				   B * id, (a * i - c * g) * id,  -(a * f - c * d) * id,
				   C * id, -(a * h - b * g) * id, (a * e - b * d) * id};
	double tx = t.m[3], ty = t.m[7], tz = t.m[11];
	Xf o;
	for (int k = 0; k < 3; k++) {
		o.m[k * 4 + 0] = (float) r[k * 3 + 0];
		o.m[k * 4 + 1] = (float) r[k * 3 + 1];
		o.m[k * 4 + 2] = (float) r[k * 3 + 2];
		o.m[k * 4 + 3] = (float) -(r[k * 3 + 0] * tx + r[k * 3 + 1] * ty + r[k * 3 + 2] * tz);
	}
	return o;
}

/* ---- transform chains with SRT motion keys (motion blur over a multi-level scene graph) --------
 * The reference wraps animated nodes in OptixSRTMotionTransform objects (optix.cpp:400-563) and OptiX
 * evaluates the transform list at the ray's time; OptiX is closed, so this build states the
 * evaluation itself (DESIGN.md "Motion spec"; the kernels in kiraray_b200/csrc/motion.cuh perform the
 * same operations in the same order):
 *   time clamped to [t0,t1]; u = (time-t0)/(t1-t0)*(n-1); k = min(int(u), n-2); f = u-k;
 *   every SRT component a + f*(b-a); quaternion divided by its norm;
 *   node M = T*R*S, M^-1 = S^-1 * R^T * T^-1; chain M = M_root*...*M_leaf, M^-1 in reverse order. */
struct XNode {
	int parent = -1, nKeys = 0;
	std::vector<float> keys; /* 10 floats per key: s[3], q[4] (x,y,z,w), t[3] */
	float t0 = 0, t1 = 1;
	Xf local, localInv;
};
Xf xfMul(const Xf &a, const Xf &b) {
	Xf c;
	for (int r = 0; r < 3; r++)
		for (int k = 0; k < 4; k++) {
			float v = a.m[r * 4] * b.m[k];
			v = v + a.m[r * 4 + 1] * b.m[4 + k];
			v = v + a.m[r * 4 + 2] * b.m[8 + k];
			c.m[r * 4 + k] = k == 3 ? v + a.m[r * 4 + 3] : v;
		}
	return c;
}
void nodeXf(const XNode &nd, float time, Xf &m, Xf &inv) {
	if (nd.nKeys < 2) { m = nd.local, inv = nd.localInv; return; }
	int n = nd.nKeys, k = 0;
	float f = 0.f;
	if (time >= nd.t1) k = n - 2, f = 1.f;
	else if (time > nd.t0) {
		float u = ((time - nd.t0) / (nd.t1 - nd.t0)) * (float) (n - 1);
		k = (int) u;
		if (k > n - 2) k = n - 2;
		f = u - (float) k;
	}
	const float *a = &nd.keys[10 * k], *b = a + 10;
	float v[10];
	for (int i = 0; i < 10; i++) { float d = b[i] - a[i]; d = f * d; v[i] = a[i] + d; }
	float l0 = v[3] * v[3], l1 = v[4] * v[4], l2 = v[5] * v[5], l3 = v[6] * v[6];
	float len = std::sqrt((l0 + l1) + (l2 + l3));
	/* one correctly rounded reciprocal and four products instead of four divisions (and three reciprocals for
	 * the nine entries of the inverse below): the GPU evaluates this per ray and instance entry */
	float rl = 1.f / len;
	float x = v[3] * rl, y = v[4] * rl, z = v[5] * rl, w = v[6] * rl;
	float rs[3] = {1.f / v[0], 1.f / v[1], 1.f / v[2]};
	float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
	float R[9] = {1.f - 2.f * (yy + zz), 2.f * (xy - wz), 2.f * (xz + wy),
				  2.f * (xy + wz), 1.f - 2.f * (xx + zz), 2.f * (yz - wx),
				  2.f * (xz - wy), 2.f * (yz + wx), 1.f - 2.f * (xx + yy)};
	for (int r = 0; r < 3; r++) {
		for (int c = 0; c < 3; c++) m.m[r * 4 + c] = R[r * 3 + c] * v[c], inv.m[r * 4 + c] = R[c * 3 + r] * rs[r];
		m.m[r * 4 + 3] = v[7 + r];
	}
	for (int r = 0; r < 3; r++) {
		float t = inv.m[r * 4] * v[7];
		t = t + inv.m[r * 4 + 1] * v[8];
		t = t + inv.m[r * 4 + 2] * v[9];
		inv.m[r * 4 + 3] = -t;
	}
}
/* The ORIGINAL node evaluation: four divisions for the quaternion, a division per entry of the inverse (the
 * formula this build used before the reciprocal form above replaced it for speed).  Kept as an independent second
 * statement: tests/test_oracle_motion.py bounds the difference between the two in ulp, so that the pair
 * (oracle, kernel) cannot drift together unnoticed. */
void nodeXfDiv(const XNode &nd, float time, Xf &m, Xf &inv) {
	if (nd.nKeys < 2) { m = nd.local, inv = nd.localInv; return; }
	int n = nd.nKeys, k = 0;
	float f = 0.f;
	if (time >= nd.t1) k = n - 2, f = 1.f;
	else if (time > nd.t0) {
		float u = ((time - nd.t0) / (nd.t1 - nd.t0)) * (float) (n - 1);
		k = (int) u;
		if (k > n - 2) k = n - 2;
		f = u - (float) k;
	}
	const float *a = &nd.keys[10 * k], *b = a + 10;
	float v[10];
	for (int i = 0; i < 10; i++) { float d = b[i] - a[i]; d = f * d; v[i] = a[i] + d; }
	float l0 = v[3] * v[3], l1 = v[4] * v[4], l2 = v[5] * v[5], l3 = v[6] * v[6];
	float len = std::sqrt((l0 + l1) + (l2 + l3));
	float x = v[3] / len, y = v[4] / len, z = v[5] / len, w = v[6] / len;
	float xx = x * x, yy = y * y, zz = z * z, xy = x * y, xz = x * z, yz = y * z, wx = w * x, wy = w * y, wz = w * z;
	float R[9] = {1.f - 2.f * (yy + zz), 2.f * (xy - wz), 2.f * (xz + wy),
				  2.f * (xy + wz), 1.f - 2.f * (xx + zz), 2.f * (yz - wx),
				  2.f * (xz - wy), 2.f * (yz + wx), 1.f - 2.f * (xx + yy)};
	for (int r = 0; r < 3; r++) {
		for (int c = 0; c < 3; c++) m.m[r * 4 + c] = R[r * 3 + c] * v[c], inv.m[r * 4 + c] = R[c * 3 + r] / v[r];
		m.m[r * 4 + 3] = v[7 + r];
	}
	for (int r = 0; r < 3; r++) {
		float t = inv.m[r * 4] * v[7];
		t = t + inv.m[r * 4 + 1] * v[8];
		t = t + inv.m[r * 4 + 2] * v[9];
		inv.m[r * 4 + 3] = -t;
	}
}
void chainXfDiv(const std::vector<XNode> &nodes, int node, float time, Xf &m, Xf &inv) {
	nodeXfDiv(nodes[node], time, m, inv);
	for (int p = nodes[node].parent; p >= 0; p = nodes[p].parent) {
		Xf pm, pinv;
		nodeXfDiv(nodes[p], time, pm, pinv);
		m = xfMul(pm, m), inv = xfMul(inv, pinv);
	}
}
void chainXf(const std::vector<XNode> &nodes, int node, float time, Xf &m, Xf &inv) {
	nodeXf(nodes[node], time, m, inv);
	for (int p = nodes[node].parent; p >= 0; p = nodes[p].parent) {
		Xf pm, pinv;
		nodeXf(nodes[p], time, pm, pinv);
		m = xfMul(pm, m), inv = xfMul(inv, pinv);
	}
}

/* ---- scene ---- */
struct BvhNode { float lo[3], hi[3]; int left, right, first, count; };
struct Mesh {
	std::vector<V3> P, N, T;
	std::vector<float> UV;
	std::vector<int32_t> I;
	int material, mediumIn, mediumOut;
	V3 Le;
	int ntri() const { return (int) I.size() / 3; }
	/* object-space median-split BVH over the triangles (use_bvh runs of scenes with MOVING instances: the ray is
	 * moved to the instance's object space at its time, exactly as testPrim does, and walks this tree instead of
	 * the whole triangle list; the boxes only cull, hits and tie-breaking are those of the brute-force loop) */
	std::vector<BvhNode> bvh;
	std::vector<int> bvhTris;
};
struct Instance {
	int mesh;
	Xf xf, inv;
	int lightBase; /* index of this instance's first triangle light in Scene::lights, -1 = none */
	int motion = -1; /* moving instance: first node of its transform chain in OrcScene::xnodes */
};
struct LightRef {
	int type;	 /* KRR_LIGHT_* */
	int inst, prim; /* diffuse area */
	int analytic;	/* index into Scene::analytic */
};
/* HomogeneousMedium / dense-grid stand-in for NanoVDBMedium<float> (src/render/media.h:108-227) */
struct MediumData {
	int type;
	float sigma_t[3], albedo[3], Le[3], g;
	Xf xf, inv;
	float bmin[3], bmax[3];
	int res[3];
	std::vector<float> density;
	std::vector<float> albedoGrid; /* optional: 3 floats per voxel on the density lattice (NanoVDBMedium::albedoGrid) */
	float scale;
	std::vector<float> majorant; /* 64^3 max-density grid, media.cpp:18-75 */
};

} // namespace

struct OrcScene {
	std::vector<Mesh> meshes;
	std::vector<Instance> instances;
	std::vector<KrrMaterialDesc> materials;
	std::vector<LightRef> lights;
	std::vector<OlLight> analytic;
	std::vector<int> infinite; /* indices into analytic */
	std::vector<KrrTextureDesc> analyticTex; /* per analytic light: lat-long image of an infinite light (image == NULL: none) */
	std::vector<MediumData> media;
	std::vector<XNode> xnodes;
	/* object<->world of an instance for a ray that carries `time` (getInstanceTransform, shading.h:70-76) */
	void instanceXf(int inst, float time, Xf &m, Xf &inv) const {
		const Instance &in = instances[inst];
		if (in.motion >= 0) chainXf(xnodes, in.motion, time, m, inv);
		else m = in.xf, inv = in.inv;
	}
	std::vector<int> movingInstances; /* not in the BVH: always tested */
	/* use_bvh == 2 only: a BVH over the moving instances' boxes for the ray-time window of the current render (see
	 * buildMovingTlas); null = every moving instance is visited.  Set for the duration of one render call. */
	mutable std::vector<BvhNode> movingTlas;
	mutable std::vector<int> movingTlasIds;
	mutable bool useMovingTlas = false;
	/* flattened world-independent primitive list for the BVH: (instance, prim) in object space */
	struct Prim { int inst, prim; };
	std::vector<Prim> prims;
	std::vector<BvhNode> bvh;
	std::vector<Prim> bvhPrims;
};

namespace {

/* ------------------------------------------------------------------------------------------------
 * Intersection spec (shared with kiraray_b200/csrc/intersect.cuh -- keep the arithmetic identical):
 *  - the world ray is moved to object space with the instance's inverse 3x4: o' = Minv*o (point),
 *    d' = Minv*d (vector); t is therefore the world-space parameter (d is not renormalised)
 *  - Moeller-Trumbore, every product and sum individually rounded (no FMA), in the order below
 *  - accept 0 < t < tmax; among equal t the smaller (instance, primitive) wins, so the result is
 *    independent of traversal order
 * ---------------------------------------------------------------------------------------------- */
inline bool triIntersect(V3 o, V3 d, V3 v0, V3 v1, V3 v2, float tmax, float &t, float &u, float &v) {
	V3 e1 = v1 - v0, e2 = v2 - v0;
	V3 pv = cross(d, e2);
	float det = dot(e1, pv);
	if (det == 0.f) return false;
	float inv = 1.f / det;
	V3 tv = o - v0;
	u = dot(tv, pv) * inv;
	if (!(u >= 0.f && u <= 1.f)) return false;
	V3 qv = cross(tv, e1);
	v = dot(d, qv) * inv;
	if (!(v >= 0.f && u + v <= 1.f)) return false;
	t = dot(e2, qv) * inv;
	return t > 0.f && t < tmax;
}

struct Hit { int inst = -1, prim = -1; float t = 0, u = 0, v = 0; };

inline bool better(float t, int inst, int prim, const Hit &h) {
	if (h.inst < 0) return true;
	if (t != h.t) return t < h.t;
	if (inst != h.inst) return inst < h.inst;
	return prim < h.prim;
}

void testPrim(const OrcScene &s, int ii, int pi, V3 o, V3 d, float tmax, float time, Hit &best) {
	const Instance &in = s.instances[ii];
	const Mesh &m	   = s.meshes[in.mesh];
	V3 oo, dd;
	if (in.motion >= 0) {
		Xf xf, inv;
		chainXf(s.xnodes, in.motion, time, xf, inv);
		oo = xfPoint(inv, o), dd = xfVector(inv, d);
	} else oo = xfPoint(in.inv, o), dd = xfVector(in.inv, d);
	const int32_t *idx = &m.I[3 * pi];
	float t, u, v;
	if (triIntersect(oo, dd, m.P[idx[0]], m.P[idx[1]], m.P[idx[2]], tmax, t, u, v))
		if (better(t, ii, pi, best)) best = Hit{ii, pi, t, u, v};
}

void primBounds(const OrcScene &s, const OrcScene::Prim &p, float lo[3], float hi[3]) {
	const Instance &in = s.instances[p.inst];
	const Mesh &m	   = s.meshes[in.mesh];
	for (int k = 0; k < 3; k++) lo[k] = 1e30f, hi[k] = -1e30f;
	for (int c = 0; c < 3; c++) {
		V3 w = xfPoint(in.xf, m.P[m.I[3 * p.prim + c]]);
		for (int k = 0; k < 3; k++) lo[k] = std::min(lo[k], w[k]), hi[k] = std::max(hi[k], w[k]);
	}
	/* pad: world-space boxes are only a culling aid for object-space tests; keep them conservative */
	for (int k = 0; k < 3; k++) {
		float e = 1e-4f * std::max(1.f, std::max(std::fabs(lo[k]), std::fabs(hi[k])));
		lo[k] -= e, hi[k] += e;
	}
}

int buildBvh(OrcScene &s, int first, int count) {
	BvhNode n;
	for (int k = 0; k < 3; k++) n.lo[k] = 1e30f, n.hi[k] = -1e30f;
	float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
	for (int i = first; i < first + count; i++) {
		float lo[3], hi[3];
		primBounds(s, s.bvhPrims[i], lo, hi);
		for (int k = 0; k < 3; k++) {
			n.lo[k] = std::min(n.lo[k], lo[k]), n.hi[k] = std::max(n.hi[k], hi[k]);
			float c = 0.5f * (lo[k] + hi[k]);
			clo[k] = std::min(clo[k], c), chi[k] = std::max(chi[k], c);
		}
	}
	n.left = n.right = -1, n.first = first, n.count = count;
	int id = (int) s.bvh.size();
	s.bvh.push_back(n);
	if (count <= 4) return id;
	int axis = 0;
	for (int k = 1; k < 3; k++) if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
	if (chi[axis] - clo[axis] <= 0) return id;
	int mid = first + count / 2;
	std::nth_element(s.bvhPrims.begin() + first, s.bvhPrims.begin() + mid, s.bvhPrims.begin() + first + count,
		[&](const OrcScene::Prim &a, const OrcScene::Prim &b) {
			float lo[3], hi[3], lo2[3], hi2[3];
			primBounds(s, a, lo, hi); primBounds(s, b, lo2, hi2);
			return lo[axis] + hi[axis] < lo2[axis] + hi2[axis];
		});
	int l = buildBvh(s, first, mid - first);
	int r = buildBvh(s, mid, first + count - mid);
	s.bvh[id].left = l, s.bvh[id].right = r, s.bvh[id].count = 0;
	return id;
}

inline bool boxHit(const BvhNode &n, V3 o, V3 invd, float tmax) {
	float t0 = 0, t1 = tmax;
	for (int k = 0; k < 3; k++) {
		float a = (n.lo[k] - o[k]) * invd[k], b = (n.hi[k] - o[k]) * invd[k];
		if (a > b) std::swap(a, b);
		if (a != a || b != b) continue; /* NaN (0*inf): slab gives no information */
		t0 = a > t0 ? a : t0;
		t1 = b < t1 ? b : t1;
	}
	return t0 <= t1 * 1.0000004f + 1e-30f;
}

int buildMeshBvh(Mesh &m, int first, int count) {
	BvhNode n;
	for (int k = 0; k < 3; k++) n.lo[k] = 1e30f, n.hi[k] = -1e30f;
	float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
	auto centre = [&](int t, int k) { return (m.P[m.I[3 * t]][k] + m.P[m.I[3 * t + 1]][k] + m.P[m.I[3 * t + 2]][k]); };
	for (int i = first; i < first + count; i++) {
		const int t = m.bvhTris[i];
		for (int c = 0; c < 3; c++)
			for (int k = 0; k < 3; k++) {
				float v = m.P[m.I[3 * t + c]][k];
				n.lo[k] = std::min(n.lo[k], v), n.hi[k] = std::max(n.hi[k], v);
			}
		for (int k = 0; k < 3; k++) clo[k] = std::min(clo[k], centre(t, k)), chi[k] = std::max(chi[k], centre(t, k));
	}
	for (int k = 0; k < 3; k++) { /* conservative: the boxes only cull */
		float e = 1e-5f * std::max(1.f, std::max(std::fabs(n.lo[k]), std::fabs(n.hi[k])));
		n.lo[k] -= e, n.hi[k] += e;
	}
	n.left = n.right = -1, n.first = first, n.count = count;
	int id = (int) m.bvh.size();
	m.bvh.push_back(n);
	if (count <= 4) return id;
	int axis = 0;
	for (int k = 1; k < 3; k++) if (chi[k] - clo[k] > chi[axis] - clo[axis]) axis = k;
	if (chi[axis] - clo[axis] <= 0) return id;
	int mid = first + count / 2;
	std::nth_element(m.bvhTris.begin() + first, m.bvhTris.begin() + mid, m.bvhTris.begin() + first + count,
					 [&](int a, int b) { return centre(a, axis) < centre(b, axis); });
	int l = buildMeshBvh(m, first, mid - first);
	int r = buildMeshBvh(m, mid, first + count - mid);
	m.bvh[id].left = l, m.bvh[id].right = r, m.bvh[id].count = 0;
	return id;
}

/* one MOVING instance through its mesh's object-space BVH: same object-space ray and triangle test as testPrim */
template <typename Accept>
void traceMovingInstance(const OrcScene &s, int ii, V3 o, V3 d, float tmax, float time, Accept accept, Hit &best) {
	const Instance &in = s.instances[ii];
	const Mesh &m	   = s.meshes[in.mesh];
	Xf xf, inv;
	chainXf(s.xnodes, in.motion, time, xf, inv);
	const V3 oo = xfPoint(inv, o), dd = xfVector(inv, d);
	V3 invd = mk(1.f / dd.x, 1.f / dd.y, 1.f / dd.z);
	int stack[128], sp = 0;
	stack[sp++] = 0;
	while (sp) {
		const BvhNode &n = m.bvh[stack[--sp]];
		float lim = best.inst < 0 ? tmax : best.t;
		if (!boxHit(n, oo, invd, lim)) continue;
		if (n.left < 0) {
			for (int i = n.first; i < n.first + n.count; i++) {
				const int pi	   = m.bvhTris[i];
				const int32_t *idx = &m.I[3 * pi];
				float t, u, v;
				if (triIntersect(oo, dd, m.P[idx[0]], m.P[idx[1]], m.P[idx[2]], best.inst < 0 ? tmax : std::nextafter(best.t, 1e30f), t, u, v)) {
					Hit h{ii, pi, t, u, v};
					if (accept(h) && better(h.t, h.inst, h.prim, best)) best = h;
				}
			}
		} else {
			stack[sp++] = n.left;
			stack[sp++] = n.right;
		}
	}
}

/* ---- optional TLAS over the MOVING instances (use_bvh == 2) -------------------------------------------------------
 * The default use_bvh == 1 run visits every moving instance for every ray: that is what the kernels' motion boxes are
 * verified against, and it is slow by design.  This mode bounds each moving instance over the render's ray-time window
 * [w0, w1] the way the pass does (kiraray_b200/csrc/bvh_build.cu): the 8 corners of the mesh's object-space box at 9
 * sample times, padded by V h / 2 with V the speed bound of chainMotionBound -- restated here in double -- and walks a
 * median-split BVH over those boxes.  tests/test_oracle_motion.py checks that both modes return the same hits (so the
 * bound is checked on the CPU too); bench.py's cpu_baseline of the 10 000-instance workload uses it. */
double sigmaMaxHost(const Xf &m) { /* largest singular value of the 3x3 part, by power iteration on A^T A, with a margin */
	double a[3][3];
	for (int r = 0; r < 3; r++)
		for (int c = 0; c < 3; c++) {
			double v = 0;
			for (int k = 0; k < 3; k++) v += (double) m.m[4 * k + r] * (double) m.m[4 * k + c];
			a[r][c] = v;
		}
	double x[3] = {0.577, 0.577, 0.577}, lam = 0;
	for (int it = 0; it < 64; it++) {
		double y[3] = {0, 0, 0};
		for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) y[r] += a[r][c] * x[c];
		double n = std::sqrt(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]);
		if (n <= 0) break;
		lam = n;
		for (int r = 0; r < 3; r++) x[r] = y[r] / n;
	}
	/* power iteration converges from below; the Frobenius norm is an upper bound: take a padded iterate, capped by it */
	double fro = std::sqrt(a[0][0] + a[1][1] + a[2][2]);
	return std::min(fro, std::sqrt(lam) * 1.01 + 1e-12);
}
double chainSpeedBoundHost(const std::vector<XNode> &nodes, int node, double pointNorm) {
	double B = pointNorm, V = 0.0;
	for (int p = node; p >= 0; p = nodes[p].parent) {
		const XNode &nd = nodes[p];
		if (nd.nKeys >= 2) {
			const float *keys = nd.keys.data();
			double sigma = 0, tMax = 0, A = 0, C = 0;
			for (int k = 0; k < nd.nKeys; k++) {
				const float *a = keys + 10 * k;
				for (int c = 0; c < 3; c++) sigma = std::max(sigma, std::fabs((double) a[c]));
				tMax = std::max(tMax, std::sqrt((double) a[7] * a[7] + (double) a[8] * a[8] + (double) a[9] * a[9]));
			}
			for (int k = 0; k + 1 < nd.nKeys; k++) {
				const float *a = keys + 10 * k, *b = a + 10;
				double ds = 0, dT = 0, dd = 0, aa = 0, ad = 0, bb = 0;
				for (int c = 0; c < 3; c++) ds = std::max(ds, std::fabs((double) b[c] - a[c])), dT += ((double) b[7 + c] - a[7 + c]) * ((double) b[7 + c] - a[7 + c]);
				for (int c = 3; c < 7; c++) {
					const double d = (double) b[c] - a[c];
					dd += d * d, aa += (double) a[c] * a[c], ad += (double) a[c] * d, bb += (double) b[c] * b[c];
				}
				double q2 = std::min(aa, bb);
				if (dd > 0) {
					const double f = -ad / dd;
					if (f > 0 && f < 1) q2 = std::min(q2, std::max(aa - ad * ad / dd, 0.0));
				}
				const double omega = dd > 0 ? 2.0 * std::sqrt(dd) / std::max(std::sqrt(q2), 1e-30) : 0.0;
				A = std::max(A, omega * sigma + ds), C = std::max(C, std::sqrt(dT));
			}
			const double fp = (double) (nd.nKeys - 1) / std::max((double) nd.t1 - (double) nd.t0, 1e-30);
			V = fp * (A * B + C) + sigma * V;
			B = sigma * B + tMax;
		} else {
			const double sg = sigmaMaxHost(nd.local);
			V = sg * V;
			B = sg * B + std::sqrt((double) nd.local.m[3] * nd.local.m[3] + (double) nd.local.m[7] * nd.local.m[7] + (double) nd.local.m[11] * nd.local.m[11]);
		}
	}
	return V * 1.0001;
}
int buildMovingTlasNode(const OrcScene &s, const std::vector<BvhNode> &boxes, int first, int count) {
	BvhNode n;
	for (int k = 0; k < 3; k++) n.lo[k] = 1e30f, n.hi[k] = -1e30f;
	for (int i = first; i < first + count; i++) {
		const BvhNode &b = boxes[s.movingTlasIds[i]];
		for (int k = 0; k < 3; k++) n.lo[k] = std::min(n.lo[k], b.lo[k]), n.hi[k] = std::max(n.hi[k], b.hi[k]);
	}
	n.left = n.right = -1, n.first = first, n.count = count;
	const int id = (int) s.movingTlas.size();
	s.movingTlas.push_back(n);
	if (count <= 2) return id;
	int axis = 0;
	for (int k = 1; k < 3; k++) if (n.hi[k] - n.lo[k] > n.hi[axis] - n.lo[axis]) axis = k;
	const int mid = first + count / 2;
	std::nth_element(s.movingTlasIds.begin() + first, s.movingTlasIds.begin() + mid, s.movingTlasIds.begin() + first + count,
					 [&](int a, int b) { return boxes[a].lo[axis] + boxes[a].hi[axis] < boxes[b].lo[axis] + boxes[b].hi[axis]; });
	const int l = buildMovingTlasNode(s, boxes, first, mid - first);
	const int r = buildMovingTlasNode(s, boxes, mid, first + count - mid);
	s.movingTlas[id].left = l, s.movingTlas[id].right = r, s.movingTlas[id].count = 0;
	return id;
}
void buildMovingTlas(const OrcScene &s, float w0, float w1) {
	s.movingTlas.clear(), s.movingTlasIds.clear();
	s.useMovingTlas = false;
	if (s.movingInstances.empty()) return;
	if (w1 < w0) std::swap(w0, w1);
	const int steps = w1 > w0 ? 8 : 0;
	std::vector<BvhNode> boxes(s.instances.size());
	std::vector<BvhNode> meshBox(s.meshes.size()); /* object-space box per mesh, computed when first needed */
	std::vector<char> haveMeshBox(s.meshes.size(), 0);
	for (int ii : s.movingInstances) {
		const Instance &in = s.instances[ii];
		const Mesh &m	   = s.meshes[in.mesh];
		if (!haveMeshBox[in.mesh]) {
			BvhNode &mb = meshBox[in.mesh];
			for (int k = 0; k < 3; k++) mb.lo[k] = 1e30f, mb.hi[k] = -1e30f;
			for (const V3 &p : m.P)
				for (int k = 0; k < 3; k++) mb.lo[k] = std::min(mb.lo[k], p[k]), mb.hi[k] = std::max(mb.hi[k], p[k]);
			haveMeshBox[in.mesh] = 1;
		}
		const float *lo = meshBox[in.mesh].lo, *hi = meshBox[in.mesh].hi;
		double cornerNorm = 0;
		V3 corners[8];
		for (int c = 0; c < 8; c++) {
			corners[c] = mk(c & 1 ? hi[0] : lo[0], c & 2 ? hi[1] : lo[1], c & 4 ? hi[2] : lo[2]);
			cornerNorm = std::max(cornerNorm, std::sqrt((double) corners[c].x * corners[c].x + (double) corners[c].y * corners[c].y + (double) corners[c].z * corners[c].z));
		}
		BvhNode b;
		for (int k = 0; k < 3; k++) b.lo[k] = 1e30f, b.hi[k] = -1e30f;
		for (int j = 0; j <= steps; j++) {
			const float t = steps ? (j == steps ? w1 : w0 + (w1 - w0) * ((float) j / (float) steps)) : w0;
			Xf xf, inv;
			chainXf(s.xnodes, in.motion, t, xf, inv);
			for (int c = 0; c < 8; c++) {
				const V3 w = xfPoint(xf, corners[c]);
				for (int k = 0; k < 3; k++) b.lo[k] = std::min(b.lo[k], w[k]), b.hi[k] = std::max(b.hi[k], w[k]);
			}
		}
		const double pad = steps ? chainSpeedBoundHost(s.xnodes, in.motion, cornerNorm) * 0.5 * ((double) w1 - (double) w0) / steps * 1.0001 : 0.0;
		for (int k = 0; k < 3; k++) { /* + the rounding of the object-space round trip, as the world boxes of static instances */
			const float e = 1e-5f * std::max(1.f, std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k]))) + 1e-6f * (b.hi[k] - b.lo[k]);
			b.lo[k] -= e + (float) pad, b.hi[k] += e + (float) pad;
		}
		boxes[ii] = b;
		s.movingTlasIds.push_back(ii);
	}
	buildMovingTlasNode(s, boxes, 0, (int) s.movingTlasIds.size());
	s.useMovingTlas = true;
}

/* closest hit over the whole scene; `skipNull`/anyhit variants below */
template <typename Accept>
Hit traceClosest(const OrcScene &s, bool useBvh, V3 o, V3 d, float tmax, float time, Accept accept) {
	Hit best;
	if (!useBvh || (s.bvh.empty() && !s.useMovingTlas)) {
		for (int ii = 0; ii < (int) s.instances.size(); ii++) {
			int nt = s.meshes[s.instances[ii].mesh].ntri();
			for (int pi = 0; pi < nt; pi++) {
				Hit h;
				testPrim(s, ii, pi, o, d, best.inst < 0 ? tmax : std::nextafter(best.t, 1e30f), time, h);
				if (h.inst >= 0 && accept(h) && better(h.t, h.inst, h.prim, best)) best = h;
			}
		}
		return best;
	}
	/* moving instances are kept out of the (static, world-space) BVH: EVERY one of them is visited (no motion
	 * bounds in the oracle: the kernels' conservative motion boxes are verified against this), each through the
	 * object-space BVH of its mesh */
	V3 invd = mk(1.f / d.x, 1.f / d.y, 1.f / d.z);
	if (s.useMovingTlas) { /* use_bvh == 2: only the moving instances whose window box the ray meets */
		int mstack[128], msp = 0;
		mstack[msp++] = 0;
		while (msp) {
			const BvhNode &n = s.movingTlas[mstack[--msp]];
			if (!boxHit(n, o, invd, best.inst < 0 ? tmax : best.t)) continue;
			if (n.left < 0) {
				for (int i = n.first; i < n.first + n.count; i++) traceMovingInstance(s, s.movingTlasIds[i], o, d, tmax, time, accept, best);
			} else mstack[msp++] = n.left, mstack[msp++] = n.right;
		}
	} else
		for (int ii : s.movingInstances) traceMovingInstance(s, ii, o, d, tmax, time, accept, best);
	int stack[128], sp = 0;
	if (!s.bvh.empty()) stack[sp++] = 0;
	while (sp) {
		const BvhNode &n = s.bvh[stack[--sp]];
		float lim = best.inst < 0 ? tmax : best.t;
		if (!boxHit(n, o, invd, lim)) continue;
		if (n.left < 0) {
			for (int i = n.first; i < n.first + n.count; i++) {
				Hit h;
				testPrim(s, s.bvhPrims[i].inst, s.bvhPrims[i].prim, o, d,
						 best.inst < 0 ? tmax : std::nextafter(best.t, 1e30f), time, h);
				if (h.inst >= 0 && accept(h) && better(h.t, h.inst, h.prim, best)) best = h;
			}
		} else {
			stack[sp++] = n.left;
			stack[sp++] = n.right;
		}
	}
	return best;
}

/* ---- texture sampling: sampleTexture, shading.h:35-44 (constant, or bilinear wrap image) ---- */
void sampleTex(const KrrTextureDesc &t, float u, float v, const float fallback[4], float out[4]) {
	if (!t.valid) { memcpy(out, fallback, 16); return; }
	if (!t.image) { memcpy(out, t.value, 16); return; }
	/* cudaFilterModeLinear, normalized coords, wrap: x = u*W - 0.5 */
	float x = u * t.width - 0.5f, y = v * t.height - 0.5f;
	float fx = std::floor(x), fy = std::floor(y);
	float ax = x - fx, ay = y - fy;
	auto wrap = [](int i, int n) { i %= n; return i < 0 ? i + n : i; };
	int x0 = wrap((int) fx, t.width), x1 = wrap((int) fx + 1, t.width);
	int y0 = wrap((int) fy, t.height), y1 = wrap((int) fy + 1, t.height);
	for (int c = 0; c < 4; c++) {
		float a = t.image[(y0 * t.width + x0) * 4 + c], b = t.image[(y0 * t.width + x1) * 4 + c];
		float cc = t.image[(y1 * t.width + x0) * 4 + c], dd = t.image[(y1 * t.width + x1) * 4 + c];
		out[c] = (1 - ay) * ((1 - ax) * a + ax * b) + ay * ((1 - ax) * cc + ax * dd);
	}
}

inline float luminanceRGB(const float c[3]) { return c[0] * 0.299f + c[1] * 0.587f + c[2] * 0.114f; }

/* InfiniteLight::Li with a lat-long IMAGE (light.h:242-246): L = tint (= 1) * image.evaluate(uv), uv =
 * worldToLatLong(rotation^T wi) (math_utils.h:133-139).  On the host the reference's image.evaluate() only
 * knows the constant (tex2D is device code), so the texel is fetched here (sampleTex: the bilinear / wrap
 * filter of the CUDA texture object, texture.cpp:229-246) and handed to the reference's InfiniteLight as its
 * tint, which then runs the reference's own fromRGB(RGBIlluminant). */
void infiniteLightLi(const OlLight &l, const KrrTextureDesc &tex, const float wi[3], const float lambda[4], float L[4]) {
	if (!tex.image) { ol_inflight_Li(&l, wi, lambda, L); return; }
	float d[3];
	for (int c = 0; c < 3; c++) d[c] = l.rotation[0 * 3 + c] * wi[0] + l.rotation[1 * 3 + c] * wi[1] + l.rotation[2 * 3 + c] * wi[2];
	float n = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
	if (n > 0) d[0] /= n, d[1] /= n, d[2] /= n;
	float u = std::atan2(d[0], -d[2]) * 0.15915494309189533577f + 0.5f, v = std::acos(d[1]) * 0.318309886183790671538f;
	const float one[4] = {1, 1, 1, 1};
	float texel[4];
	sampleTex(tex, u, v, one, texel);
	OlLight tinted = l;
	memcpy(tinted.color, texel, 12);
	ol_inflight_Li(&tinted, wi, lambda, L);
}

/* getPerpendicular, src/util/math_utils.h:122-133 */
V3 getPerpendicular(V3 u) {
	V3 a = mk(std::fabs(u.x), std::fabs(u.y), std::fabs(u.z));
	uint32_t uyx = (a.x - a.y) < 0 ? 1 : 0, uzx = (a.x - a.z) < 0 ? 1 : 0, uzy = (a.y - a.z) < 0 ? 1 : 0;
	uint32_t xm = uyx & uzx, ym = (1 ^ xm) & uzy, zm = 1 ^ (xm | ym);
	return normalize(cross(u, mk((float) xm, (float) ym, (float) zm)));
}

struct SurfIntr {
	V3 p, wo, n, tangent, bitangent;
	float uv[2];
	float time;
	int material; /* -1 = null */
	int light;	  /* index into Scene::lights or -1 */
	int inst, prim;
	int medium; /* intr.medium = ray.medium, shading.h:126 */
	int mesh;
	OlShading sd;
};

/* offsetRayOrigin / spawnRayTowards / spawnRayTo: src/core/raytracing.h:106-160 */
inline V3 offsetRayOrigin(V3 p, V3 n, V3 w) {
	V3 off = n * 1e-4f;
	if (dot(n, w) < 0.f) off = -off;
	return p + off;
}

/* prepareSurfaceInteraction, shading.h:115-226 */
void prepareInteraction(const OrcScene &s, const Hit &h, V3 rayDir, float rayTime, const float lambda[4],
						const float pdf[4], SurfIntr &it) {
	const Instance &in = s.instances[h.inst];
	const Mesh &m	   = s.meshes[in.mesh];
	float b[3]		   = {1 - h.u - h.v, h.u, h.v}; /* getHitInfo, shading.h:85 */
	const int32_t *v   = &m.I[3 * h.prim];
	it.inst = h.inst, it.prim = h.prim, it.mesh = in.mesh;
	it.medium = -1;
	it.time = rayTime;
	it.wo	= normalize(-normalize(rayDir)); /* hitInfo.wo = -normalize(dir); intr.wo = normalize(hitInfo.wo) */
	V3 p0 = m.P[v[0]], p1 = m.P[v[1]], p2 = m.P[v[2]];
	it.p = b[0] * p0 + b[1] * p1 + b[2] * p2;
	if (!m.N.empty()) it.n = normalize(b[0] * m.N[v[0]] + b[1] * m.N[v[1]] + b[2] * m.N[v[2]]);
	else it.n = normalize(cross(p1 - p0, p2 - p0));
	if (!m.T.empty()) {
		it.tangent = normalize(b[0] * m.T[v[0]] + b[1] * m.T[v[1]] + b[2] * m.T[v[2]]);
		it.tangent = normalize(it.tangent - it.n * dot(it.n, it.tangent));
	} else it.tangent = getPerpendicular(it.n);
	it.bitangent = normalize(cross(it.n, it.tangent));
	it.uv[0] = it.uv[1] = 0;
	if (!m.UV.empty())
		for (int k = 0; k < 2; k++)
			it.uv[k] = b[0] * m.UV[2 * v[0] + k] + b[1] * m.UV[2 * v[1] + k] + b[2] * m.UV[2 * v[2] + k];
	it.light	= in.lightBase >= 0 ? in.lightBase + h.prim : -1;
	it.material = m.material;
	/* object -> world, with the transform list evaluated at the ray's time */
	Xf ixf, iinv;
	s.instanceXf(h.inst, rayTime, ixf, iinv);
	it.p		 = xfPoint(ixf, it.p);
	it.n		 = normalize(xfNormal(iinv, it.n));
	it.tangent	 = normalize(xfNormal(iinv, it.tangent));
	it.bitangent = normalize(xfNormal(iinv, it.bitangent));
	if (it.material < 0) return;

	const KrrMaterialDesc &mat = s.materials[it.material];
	memset(&it.sd, 0, sizeof(it.sd));
	it.sd.bsdfType			   = mat.bsdf_type;
	it.sd.specularTransmission = mat.specular_transmission;
	it.sd.IoR				   = mat.ior;
	if (mat.spectral_eta.kind == KRR_SPEC_CONSTANT) {
		it.sd.IoR = mat.spectral_eta.a[0];
		it.sd.etaKind = 1, it.sd.etaValue[0] = mat.spectral_eta.a[0];
	}
	if (mat.spectral_k.kind == KRR_SPEC_CONSTANT) it.sd.kKind = 1, it.sd.kValue[0] = mat.spectral_k.a[0];
	float diff[4], spec[4];
	sampleTex(mat.textures[KRR_TEX_DIFFUSE], it.uv[0], it.uv[1], mat.diffuse, diff);
	sampleTex(mat.textures[KRR_TEX_SPECULAR], it.uv[0], it.uv[1], mat.specular, spec);
	const KrrTextureDesc &nt = mat.textures[KRR_TEX_NORMAL];
	if (nt.valid && !m.UV.empty()) {
		float fb[4] = {0, 0, 1, 0}, nv[4];
		sampleTex(nt, it.uv[0], it.uv[1], fb, nv);
		V3 nm = mk(2 * nv[0] - 1, 2 * nv[1] - 1, 2 * nv[2] - 1);
		it.n  = normalize(it.tangent * nm.x + it.bitangent * nm.y + it.n * nm.z);
		it.tangent	 = normalize(it.tangent - it.n * dot(it.tangent, it.n));
		it.bitangent = normalize(cross(it.n, it.tangent));
	}
	float diffuse[3], specular[3];
	if (mat.shading_model == KRR_SHADING_METALLIC_ROUGHNESS) {
		for (int k = 0; k < 3; k++) {
			/* lerp(a,b,t) = (1-t)*a + t*b  (krrmath/functors.h) */
			diffuse[k]	= (1 - spec[2]) * diff[k] + spec[2] * 0.f;
			specular[k] = (1 - spec[2]) * 0.f + spec[2] * diff[k];
		}
		it.sd.metallic	= spec[2];
		it.sd.roughness = spec[1];
	} else {
		for (int k = 0; k < 3; k++) diffuse[k] = diff[k], specular[k] = spec[k];
		it.sd.roughness = 1.f - spec[3];
		it.sd.metallic	= ol_get_metallic(diffuse, specular);
	}
	ol_from_rgb(diffuse, 0, lambda, it.sd.diffuse);
	ol_from_rgb(specular, 0, lambda, it.sd.specular);
	it.sd.anisotropic = mat.anisotropic;
	memcpy(it.sd.lambda, lambda, 16);
	memcpy(it.sd.pdf, pdf, 16);
	st(it.sd.woWorld, it.wo);
}

inline V3 toLocal(const SurfIntr &it, V3 v) { return mk(dot(it.tangent, v), dot(it.bitangent, v), dot(it.n, v)); }
inline V3 toWorld(const SurfIntr &it, V3 v) { return it.tangent * v.x + it.bitangent * v.y + it.n * v.z; }

/* alphaKilled, shading.h:89-113: only image/constant Transmission textures matter; HashFloat-based
 * stochastic kill is restated with the same MurmurHash64A (src/util/hash.h:46-90, 117-133). */
uint64_t murmur64A(const unsigned char *key, size_t len, uint64_t seed) {
	const uint64_t m = 0xc6a4a7935bd1e995ull;
	const int r = 47;
	uint64_t h = seed ^ (len * m);
	const unsigned char *end = key + 8 * (len / 8);
	while (key != end) {
		uint64_t k; memcpy(&k, key, 8); key += 8;
		k *= m; k ^= k >> r; k *= m; h ^= k; h *= m;
	}
	switch (len & 7) {
		case 7: h ^= uint64_t(key[6]) << 48; case 6: h ^= uint64_t(key[5]) << 40;
		case 5: h ^= uint64_t(key[4]) << 32; case 4: h ^= uint64_t(key[3]) << 24;
		case 3: h ^= uint64_t(key[2]) << 16; case 2: h ^= uint64_t(key[1]) << 8;
		case 1: h ^= uint64_t(key[0]); h *= m;
	}
	h ^= h >> r; h *= m; h ^= h >> r;
	return h;
}
bool alphaKilled(const OrcScene &s, const Hit &h, V3 o, V3 d) {
	const Mesh &m = s.meshes[s.instances[h.inst].mesh];
	if (m.material < 0) return false;
	const KrrTextureDesc &t = s.materials[m.material].textures[KRR_TEX_TRANSMISSION];
	if (!t.valid) return false;
	float b[3] = {1 - h.u - h.v, h.u, h.v}, uv[2] = {0, 0};
	const int32_t *v = &m.I[3 * h.prim];
	if (!m.UV.empty())
		for (int k = 0; k < 2; k++) uv[k] = b[0] * m.UV[2 * v[0] + k] + b[1] * m.UV[2 * v[1] + k] + b[2] * m.UV[2 * v[2] + k];
	float fb[4] = {1, 1, 1, 1}, op[4];
	sampleTex(t, uv[0], uv[1], fb, op);
	float alpha = 1 - luminanceRGB(op);
	if (alpha >= 1) return false;
	if (alpha <= 0) return true;
	float buf[6] = {o.x, o.y, o.z, d.x, d.y, d.z};
	float u = uint32_t(murmur64A((const unsigned char *) buf, 24, 0)) * 0x1p-32f;
	return u > alpha;
}

/* ---- light dispatch ---- */
void fillTri(const OrcScene &s, const LightRef &lr, OlTriLight &tl) {
	const Instance &in = s.instances[lr.inst];
	const Mesh &m	   = s.meshes[in.mesh];
	const int32_t *v   = &m.I[3 * lr.prim];
	for (int c = 0; c < 3; c++) {
		st(tl.p[c], m.P[v[c]]);
		st(tl.n[c], m.N.empty() ? mk(0, 0, 0) : m.N[v[c]]);
	}
	memcpy(tl.xform, in.xf.m, 48);
	/* mesh.cpp:39-59: Le = emissive constant (or Mesh::Le); scale = max; Le /= scale; one-sided.
	 * DiffuseAreaLight::L (light.h:183-188) evaluates the (valid, constant) emissive TEXTURE, i.e.
	 * the un-normalised colour, and still multiplies by scale -- kept as is. */
	float Le[3];
	bool hasTex = m.material >= 0 && s.materials[m.material].textures[KRR_TEX_EMISSIVE].valid;
	if (hasTex) memcpy(Le, s.materials[m.material].textures[KRR_TEX_EMISSIVE].value, 12);
	else st(Le, m.Le);
	float scale = std::max(Le[0], std::max(Le[1], Le[2]));
	if (!hasTex) for (int k = 0; k < 3; k++) Le[k] /= scale;
	memcpy(tl.Le, Le, 12);
	tl.scale	= scale;
	tl.twoSided = 0;
}

/* ------------------------------------------------------------------------------------------------
 * Participating media.  HomogeneousMedium::samplePoint, MajorantIterator (DDA) and the HG phase
 * function run through the leaf API (= the reference's own classes in the reference backend); the
 * dense density grid (stand-in for NanoVDBMedium<float>: this build's scenes carry a dense float array
 * instead of a .vdb file), its 64^3 majorant grid and the tracking loops are restated here.
 * ---------------------------------------------------------------------------------------------- */
constexpr int kMajRes = 64; /* majorantGridRes, media.h:223 */

/* NanoVDBGrid<float>::getValue (util/volume.h:83-87): worldToIndexF + SampleFromVoxels<Tree, 1, false>
 * = trilinear interpolation of the voxel lattice, background value 0 outside the grid */
float gridValue(const MediumData &m, const std::vector<float> &grid, int nc, int comp, V3 p) {
	float idx[3], w[3];
	int i0[3];
	for (int k = 0; k < 3; k++) {
		idx[k] = (p[k] - m.bmin[k]) / (m.bmax[k] - m.bmin[k]) * m.res[k];
		float f = std::floor(idx[k]);
		i0[k] = (int) f, w[k] = idx[k] - f;
	}
	auto at = [&](int x, int y, int z) -> float {
		if (x < 0 || y < 0 || z < 0 || x >= m.res[0] || y >= m.res[1] || z >= m.res[2]) return 0.f;
		return grid[nc * (x + (size_t) m.res[0] * (y + (size_t) m.res[1] * z)) + comp];
	};
	auto lerp = [](float a, float b, float t) { return a + t * (b - a); }; /* nanovdb TrilinearSampler */
	float c00 = lerp(at(i0[0], i0[1], i0[2]), at(i0[0] + 1, i0[1], i0[2]), w[0]);
	float c10 = lerp(at(i0[0], i0[1] + 1, i0[2]), at(i0[0] + 1, i0[1] + 1, i0[2]), w[0]);
	float c01 = lerp(at(i0[0], i0[1], i0[2] + 1), at(i0[0] + 1, i0[1], i0[2] + 1), w[0]);
	float c11 = lerp(at(i0[0], i0[1] + 1, i0[2] + 1), at(i0[0] + 1, i0[1] + 1, i0[2] + 1), w[0]);
	return lerp(lerp(c00, c10, w[1]), lerp(c01, c11, w[1]), w[2]);
}
float gridDensity(const MediumData &m, V3 p) { return gridValue(m, m.density, 1, 0, p); }

/* initializeMajorantGrid, media.cpp:18-75 (index bbox of a dense grid = [0, res-1]) */
void buildMajorant(MediumData &m) {
	m.majorant.assign((size_t) kMajRes * kMajRes * kMajRes, 0.f);
	for (int z = 0; z < kMajRes; z++)
		for (int y = 0; y < kMajRes; y++)
			for (int x = 0; x < kMajRes; x++) {
				int c[3] = {x, y, z}, n0[3], n1[3];
				for (int k = 0; k < 3; k++) {
					float ext = m.bmax[k] - m.bmin[k];
					float w0 = m.bmin[k] + ext * (float(c[k]) / kMajRes), w1 = m.bmin[k] + ext * (float(c[k] + 1) / kMajRes);
					float i0 = (w0 - m.bmin[k]) / ext * m.res[k], i1 = (w1 - m.bmin[k]) / ext * m.res[k];
					n0[k] = std::max(int(i0 - 1.f), 0), n1[k] = std::min(int(i1 + 1.f), m.res[k] - 1);
				}
				float mx = 0;
				for (int nz = n0[2]; nz <= n1[2]; nz++)
					for (int ny = n0[1]; ny <= n1[1]; ny++)
						for (int nx = n0[0]; nx <= n1[0]; nx++)
							mx = std::max(mx, m.density[nx + (size_t) m.res[0] * (ny + (size_t) m.res[1] * nz)]);
				m.majorant[x + kMajRes * (y + kMajRes * z)] = mx;
			}
}

struct MediumPoint { Spec sigma_a, sigma_s, Le; };
Spec fromRGB(const float rgb[3], int type, const float lambda[4]) {
	Spec r;
	ol_from_rgb(rgb, type, lambda, r.v);
	return r;
}
/* Medium::samplePoint: HomogeneousMedium media.h:121-126, NanoVDBMedium<float> media.h:161-173 */
MediumPoint mediumSamplePoint(const MediumData &m, V3 p, const float lambda[4]) {
	MediumPoint mp;
	if (m.type == KRR_MEDIUM_HOMOGENEOUS) {
		ol_homogeneous_sample_point(m.sigma_t, m.albedo, m.Le, lambda, mp.sigma_a.v, mp.sigma_s.v, mp.Le.v);
		return mp;
	}
	V3 pm = xfPoint(m.inv, p);
	Spec sigma_t = fromRGB(m.sigma_t, 1, lambda) * (gridDensity(m, pm) * m.scale); /* density * scale * fromRGB(sigma_t) */
	/* albedoGrid ? RGB(albedoGrid.getValue(p)) : albedo, as RGBBounded (media.h:168-170) */
	float albedo[3] = {m.albedo[0], m.albedo[1], m.albedo[2]};
	if (!m.albedoGrid.empty())
		for (int c = 0; c < 3; c++) albedo[c] = gridValue(m, m.albedoGrid, 3, c, pm);
	Spec sigma_s = sigma_t * fromRGB(albedo, 0, lambda);
	for (int i = 0; i < 4; i++) mp.sigma_a.v[i] = sigma_t.v[i] - sigma_s.v[i];
	mp.sigma_s = sigma_s, mp.Le = sconst(0); /* no temperature grid */
	return mp;
}

struct Segment { float tMin, tMax; Spec sigma_maj; };
/* Medium::sampleRay: media.h:128-132 (homogeneous), 175-187 (grid; AABB::intersect krrmath/aabb.h:58-76) */
int mediumSampleRay(const MediumData &m, V3 o, V3 d, float raytMax, const float lambda[4], Segment *segs, int cap) {
	float out[6 * 256];
	int n;
	if (m.type == KRR_MEDIUM_HOMOGENEOUS) {
		Spec st = fromRGB(m.sigma_t, 1, lambda);
		float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
		n = ol_majorant_segments(m.bmin, m.bmax, m.res, nullptr, oo, dd, 0.f, raytMax, st.v, out, 256);
	} else {
		V3 lo = xfPoint(m.inv, o), ld = xfVector(m.inv, d);
		float t0 = 0, t1 = raytMax;
		for (int i = 0; i < 3; i++) {
			float inv = 1 / ld[i];
			float tn = (m.bmin[i] - lo[i]) * inv, tf = (m.bmax[i] - lo[i]) * inv;
			if (tn > tf) std::swap(tn, tf);
			t0 = tn > t0 ? tn : t0;
			t1 = tf < t1 ? tf : t1;
			if (t0 > t1) return 0;
		}
		Spec st = fromRGB(m.sigma_t, 1, lambda) * m.scale;
		int res[3] = {kMajRes, kMajRes, kMajRes};
		float oo[3] = {lo.x, lo.y, lo.z}, dd[3] = {ld.x, ld.y, ld.z};
		n = ol_majorant_segments(m.bmin, m.bmax, res, m.majorant.data(), oo, dd, t0, t1, st.v, out, 256);
	}
	n = std::min(n, std::min(cap, 256));
	for (int i = 0; i < n; i++) {
		segs[i].tMin = out[6 * i], segs[i].tMax = out[6 * i + 1];
		memcpy(segs[i].sigma_maj.v, out + 6 * i + 2, 16);
	}
	return n;
}

inline Spec expNeg(Spec s, float dt) { Spec r; for (int i = 0; i < 4; i++) r.v[i] = std::exp(-dt * s.v[i]); return r; }
inline Spec operator-(Spec a, Spec b) { Spec r; for (int i = 0; i < 4; i++) r.v[i] = a.v[i] - b.v[i]; return r; }
inline Spec cwiseMax0(Spec a) { Spec r; for (int i = 0; i < 4; i++) r.v[i] = std::max(a.v[i], 0.f); return r; }
inline Spec operator/(Spec a, Spec b) { Spec r; for (int i = 0; i < 4; i++) r.v[i] = a.v[i] / b.v[i]; return r; }
inline float maxCoeff(Spec a) { return std::max(std::max(a.v[0], a.v[1]), std::max(a.v[2], a.v[3])); }
inline bool hasNaN(Spec a) { return a.v[0] != a.v[0] || a.v[1] != a.v[1] || a.v[2] != a.v[2] || a.v[3] != a.v[3]; }

/* sampleT_maj, wavefront.h:34-78.  callback(p, mp, sigma_maj, T_maj) -> continue? */
template <typename F>
Spec sampleT_maj(const MediumData &m, V3 o, V3 d, float tMax, OlSampler *smp, const float lambda[4], F callback) {
	tMax *= length(d);
	d = normalize(d);
	Spec T_maj = sconst(1);
	Segment segs[256];
	int nseg = mediumSampleRay(m, o, d, tMax, lambda, segs, 256);
	const int channel = 0; /* lambda.mainIndex() in the spectral build, spectrum.h:76-81 */
	for (int si = 0; si < nseg; si++) {
		const Segment &seg = segs[si];
		if (seg.sigma_maj.v[channel] == 0) {
			float dt = seg.tMax - seg.tMin;
			if (std::isinf(dt)) dt = std::numeric_limits<float>::max();
			T_maj = T_maj * expNeg(seg.sigma_maj, dt);
			continue;
		}
		float tMin = seg.tMin;
		while (true) {
			float t = tMin + (-std::log(1 - ol_pcg_get1d(smp)) / seg.sigma_maj.v[channel]); /* sampleExponential */
			if (t < seg.tMax) {
				T_maj = T_maj * expNeg(seg.sigma_maj, t - tMin);
				V3 p = o + d * t;
				MediumPoint mp = mediumSamplePoint(m, p, lambda);
				if (!callback(p, mp, seg.sigma_maj, T_maj)) return sconst(1);
				T_maj = sconst(1);
				tMin  = t;
			} else {
				float dt = seg.tMax - tMin;
				if (std::isinf(dt)) dt = std::numeric_limits<float>::max();
				T_maj = T_maj * expNeg(seg.sigma_maj, dt);
				break;
			}
		}
	}
	return T_maj;
}

/* sampleDiscrete({a, b, c}, u), render/sampling.h:90-105 */
int sampleDiscrete3(float w0, float w1, float w2, float u) {
	float weights[3] = {w0, w1, w2};
	float sumWeights = 0;
	for (float w : weights) sumWeights += w;
	float up = u * sumWeights;
	if (up == sumWeights) up = std::nextafter(up, -std::numeric_limits<float>::infinity()); /* nextFloatDown */
	int offset = 0;
	float sum  = 0;
	while (offset < 2 && sum + weights[offset] <= up) sum += weights[offset++]; /* (the reference's loop is unbounded) */
	return offset;
}

struct PathStats {
	uint64_t camera = 0, closest = 0, shadow = 0, scatter = 0, hitLight = 0, miss = 0, mediumSample = 0, mediumScatter = 0;
	uint64_t closestByDepth[KRR_MAX_DEPTH_STATS] = {0}, shadowByDepth[KRR_MAX_DEPTH_STATS] = {0};
};

struct Capture {
	int sample, depth;
	int32_t **items;
	int32_t *counts;
	void push(int q, int pixelId, int depth_, int bsdfType, int aux) {
		int slot = __atomic_fetch_add(&counts[q], 1, __ATOMIC_RELAXED);
		int32_t *o = items[q] + 4 * slot;
		o[0] = pixelId, o[1] = depth_, o[2] = bsdfType, o[3] = aux;
	}
};

} // namespace

extern "C" int orc_intersect_triangle(const float o[3], const float d[3], const float v0[3], const float v1[3],
									  const float v2[3], float tmax, float *t, float *u, float *v) {
	return triIntersect(mk(o), mk(d), mk(v0), mk(v1), mk(v2), tmax, *t, *u, *v) ? 1 : 0;
}

extern "C" void orc_instance_xf(const OrcScene *s, int32_t inst, float time, float m[12], float inv[12]) {
	Xf a, b;
	s->instanceXf(inst, time, a, b);
	memcpy(m, a.m, 48), memcpy(inv, b.m, 48);
}

/* the same with the division-based node evaluation (nodeXfDiv); static instances: the uploaded matrices */
extern "C" void orc_instance_xf_div(const OrcScene *s, int32_t inst, float time, float m[12], float inv[12]) {
	Xf a, b;
	const Instance &in = s->instances[inst];
	if (in.motion >= 0) chainXfDiv(s->xnodes, in.motion, time, a, b);
	else a = in.xf, b = in.inv;
	memcpy(m, a.m, 48), memcpy(inv, b.m, 48);
}

extern "C" OrcScene *orc_scene_create(const KrrSceneDesc *d) {
	ol_init();
	OrcScene *s = new OrcScene();
	for (int i = 0; i < d->n_meshes; i++) {
		const KrrMeshDesc &md = d->meshes[i];
		Mesh m;
		m.P.resize(md.n_vertices);
		for (int k = 0; k < md.n_vertices; k++) m.P[k] = mk(md.positions + 3 * k);
		if (md.normals) { m.N.resize(md.n_vertices); for (int k = 0; k < md.n_vertices; k++) m.N[k] = mk(md.normals + 3 * k); }
		if (md.tangents) { m.T.resize(md.n_vertices); for (int k = 0; k < md.n_vertices; k++) m.T[k] = mk(md.tangents + 3 * k); }
		if (md.texcoords) m.UV.assign(md.texcoords, md.texcoords + 2 * md.n_vertices);
		m.I.assign(md.indices, md.indices + 3 * md.n_triangles);
		m.material = md.material, m.mediumIn = md.medium_inside, m.mediumOut = md.medium_outside;
		m.Le = mk(md.Le);
		s->meshes.push_back(std::move(m));
	}
	s->materials.assign(d->materials, d->materials + d->n_materials);
	for (int i = 0; i < d->n_instances; i++) {
		Instance in;
		in.mesh = d->instances[i].mesh;
		memcpy(in.xf.m, d->instances[i].transform, 48);
		in.inv		 = xfInverse(in.xf);
		in.lightBase = -1;
		s->instances.push_back(in);
	}
	/* transform chains: read only when motion blur is on (optix.cpp:402) */
	if (d->options.motionblur) {
		int nGraph = d->transform_nodes ? std::max(d->n_transform_nodes, 0) : 0;
		for (int i = 0; i < nGraph; i++) {
			const KrrTransformNodeDesc &nd = d->transform_nodes[i];
			XNode x;
			x.parent = nd.parent;
			memcpy(x.local.m, nd.transform, 48);
			x.localInv = xfInverse(x.local);
			if (nd.n_motion_keys >= 2) {
				x.nKeys = nd.n_motion_keys, x.t0 = nd.time_begin, x.t1 = nd.time_end;
				x.keys.assign((const float *) nd.motion_keys, (const float *) nd.motion_keys + 10 * (size_t) nd.n_motion_keys);
			}
			s->xnodes.push_back(std::move(x));
		}
		for (int i = 0; i < d->n_instances; i++) {
			const KrrInstanceDesc &id = d->instances[i];
			if (nGraph > 0 && id.transform_node >= 0) {
				bool moves = false;
				for (int p = id.transform_node; p >= 0; p = s->xnodes[p].parent) moves |= s->xnodes[p].nKeys >= 2;
				if (moves) s->instances[i].motion = id.transform_node;
			} else if (id.n_motion_keys >= 2 && id.motion_keys) {
				XNode x;
				x.nKeys = id.n_motion_keys, x.t0 = d->options.starttime, x.t1 = d->options.endtime;
				x.keys.assign((const float *) id.motion_keys, (const float *) id.motion_keys + 10 * (size_t) id.n_motion_keys);
				x.local = s->instances[i].xf, x.localInv = s->instances[i].inv;
				s->instances[i].motion = (int) s->xnodes.size();
				s->xnodes.push_back(std::move(x));
			}
			if (s->instances[i].motion >= 0) s->movingInstances.push_back(i);
		}
	}
	/* uploadSceneLightData, device/scene.cpp:91-145: mesh lights first (instance order), then scene lights */
	for (int i = 0; i < (int) s->instances.size(); i++) {
		const Mesh &m = s->meshes[s->instances[i].mesh];
		bool emissive = (m.material >= 0 && s->materials[m.material].textures[KRR_TEX_EMISSIVE].valid) ||
						(m.Le.x != 0 || m.Le.y != 0 || m.Le.z != 0);
		if (!emissive) continue;
		s->instances[i].lightBase = (int) s->lights.size();
		for (int t = 0; t < m.ntri(); t++) s->lights.push_back(LightRef{KRR_LIGHT_DIFFUSE_AREA, i, t, -1});
	}
	for (int i = 0; i < d->n_lights; i++) {
		const KrrLightDesc &ld = d->lights[i];
		OlLight l;
		memset(&l, 0, sizeof(l));
		l.type = ld.type;
		memcpy(l.color, ld.color, 12);
		l.scale = ld.scale;
		l.position[0] = ld.transform[3], l.position[1] = ld.transform[7], l.position[2] = ld.transform[11];
		for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) l.rotation[r * 3 + c] = ld.transform[r * 4 + c];
		l.sceneRadius = ld.scene_radius;
		l.cosInner = std::cos(ld.inner_cone_deg * (float) M_PI / 180.f);
		l.cosOuter = std::cos(ld.outer_cone_deg * (float) M_PI / 180.f);
		memcpy(l.xform, ld.transform, 48);
		Xf x; memcpy(x.m, ld.transform, 48);
		Xf xi = xfInverse(x);
		memcpy(l.xformInv, xi.m, 48);
		if (ld.type == KRR_LIGHT_INFINITE) {
			/* InfiniteLight::getObjectData (src/core/light.cpp:28-41) uploads through the TEXTURE constructor
			 * (light.h:213-216): tint = 1 and the colour comes from the texture; the light's own colour is
			 * not transferred.  Constant textures only here (L = tint * image.evaluate(uv), light.h:239-241). */
			l.color[0] = l.color[1] = l.color[2] = 1.f;
			if (ld.texture.valid) memcpy(l.color, ld.texture.value, 12);
			s->infinite.push_back((int) s->analytic.size());
		}
		s->lights.push_back(LightRef{ld.type, -1, -1, (int) s->analytic.size()});
		s->analytic.push_back(l);
		KrrTextureDesc tex;
		memset(&tex, 0, sizeof tex);
		if (ld.type == KRR_LIGHT_INFINITE && ld.texture.valid && ld.texture.image && ld.texture.width > 0 && ld.texture.height > 0) tex = ld.texture;
		s->analyticTex.push_back(tex);
	}
	for (int i = 0; i < d->n_media; i++) {
		const KrrMediumDesc &md = d->media[i];
		MediumData m;
		m.type = md.type, m.g = md.g, m.scale = md.scale;
		memcpy(m.sigma_t, md.sigma_t, 12), memcpy(m.albedo, md.albedo, 12), memcpy(m.Le, md.Le, 12);
		memcpy(m.xf.m, md.transform, 48);
		m.inv = xfInverse(m.xf);
		memcpy(m.bmin, md.bounds_min, 12), memcpy(m.bmax, md.bounds_max, 12), memcpy(m.res, md.res, 12);
		if (md.type == KRR_MEDIUM_GRID) {
			m.density.assign(md.density, md.density + (size_t) md.res[0] * md.res[1] * md.res[2]);
			if (md.albedo_grid) m.albedoGrid.assign(md.albedo_grid, md.albedo_grid + 3 * (size_t) md.res[0] * md.res[1] * md.res[2]);
			buildMajorant(m);
		}
		s->media.push_back(std::move(m));
	}
	for (int i = 0; i < (int) s->instances.size(); i++) {
		if (s->instances[i].motion >= 0) continue;
		for (int t = 0; t < s->meshes[s->instances[i].mesh].ntri(); t++) s->bvhPrims.push_back({i, t});
	}
	if (!s->bvhPrims.empty()) buildBvh(*s, 0, (int) s->bvhPrims.size());
	for (int i : s->movingInstances) {
		Mesh &m = s->meshes[s->instances[i].mesh];
		if (!m.bvh.empty()) continue;
		m.bvhTris.resize(m.ntri());
		for (int t = 0; t < m.ntri(); t++) m.bvhTris[t] = t;
		buildMeshBvh(m, 0, m.ntri());
	}
	return s;
}

extern "C" void orc_scene_destroy(OrcScene *s) { delete s; }
extern "C" int32_t orc_scene_num_lights(const OrcScene *s) { return (int32_t) s->lights.size(); }

extern "C" double orc_render(const OrcScene *sp, const OrcParams *p, const KrrCameraData *cam, int32_t W, int32_t H,
							 uint64_t frameIndex, float *film, int32_t *firstHits, uint64_t *samplerState,
							 float *lambdaOut, float *cameraSample, KrrStats *stats, int32_t capSample,
							 int32_t capDepth, int32_t *capItems[6], int32_t capCounts[6]) {
	const OrcScene &s = *sp;
	ol_init();
	OlCamera oc;
	memcpy(oc.filmSize, cam->film_size, 8);
	oc.focalLength = cam->focal_length, oc.focalDistance = cam->focal_distance, oc.lensRadius = cam->lens_radius;
	oc.aspectRatio = cam->aspect_ratio, oc.shutterOpen = cam->shutter_open, oc.shutterTime = cam->shutter_time;
	memcpy(oc.transform, cam->transform, 48);
	const int spp	   = p->spp > 0 ? p->spp : 1;
	const int row0 = p->row_end > 0 ? p->row_begin : 0, row1 = p->row_end > 0 ? p->row_end : H;
	const int nLights  = (int) s.lights.size();
	const bool useBvh  = p->use_bvh != 0;
	auto t0 = std::chrono::steady_clock::now(); /* (the per-render TLAS build over the moving instances is part of the timed work) */
	/* use_bvh == 2: moving instances through a BVH over their boxes for this render's ray-time window */
	struct MovingTlasScope {
		const OrcScene &sc;
		MovingTlasScope(const OrcScene &sc_, bool on, float w0, float w1) : sc(sc_) { if (on) buildMovingTlas(sc, w0, w1); else sc.useMovingTlas = false; }
		~MovingTlasScope() { sc.useMovingTlas = false; }
	} movingTlasScope(s, p->use_bvh == 2, cam->shutter_open, cam->shutter_open + cam->shutter_time);
	const bool enableMedium = p->enable_medium && !s.media.empty(); /* integrator.cpp:200 */
	Capture cap{capSample, capDepth, capItems, capCounts};
	if (capSample >= 0) for (int q = 0; q < 6; q++) capCounts[q] = 0;
	int nthreads = p->threads > 0 ? p->threads : (int) std::max(1u, std::thread::hardware_concurrency());
	std::vector<PathStats> tstats(nthreads);
	/* pixels are independent (private RNG stream + accumulator): dynamic chunks over host threads */
	std::atomic<int> nextChunk{row0 * W};
	const int chunk = 256, pixelEnd = row1 * W;
	auto worker = [&](int tid) {
	PathStats &st_ = tstats[tid];
	for (;;) {
	const int c0 = nextChunk.fetch_add(chunk);
	if (c0 >= pixelEnd) break;
	for (int pixelId = c0; pixelId < std::min(c0 + chunk, pixelEnd); pixelId++) {
		const int px = pixelId % W, py = pixelId / W;
		/* beginFrame, integrator.cpp:213-220 */
		Spec L = sconst(0);
		float pixel[3] = {0, 0, 0};
		OlSampler smp;
		ol_pcg_set_pixel_sample(&smp, px, py, (uint32_t) (frameIndex * spp));
		ol_pcg_advance(&smp, (int64_t) (256 * pixelId)); /* int arithmetic as the reference (256 * pixelId) */
		float lambda[4], lpdf[4];
		ol_sample_wavelengths(ol_pcg_get1d(&smp), lambda, lpdf);
		if (samplerState) samplerState[2 * pixelId] = smp.state, samplerState[2 * pixelId + 1] = smp.inc;
		if (lambdaOut) memcpy(lambdaOut + 4 * pixelId, lambda, 16);
		int fhInst = -1, fhPrim = -1;

		for (int sampleId = 0; sampleId < spp; sampleId++) {
			const bool capS = capSample == sampleId;
			/* generateCameraRays, integrator.cpp:166-179 */
			float cs[5];
			for (int k = 0; k < 5; k++) cs[k] = ol_pcg_get1d(&smp);
			if (cameraSample) memcpy(cameraSample + 5 * pixelId, cs, 20);
			float ro[3], rd[3], rtime;
			ol_camera_ray(&oc, px, py, W, H, cs, ro, rd, &rtime);
			st_.camera++;
			/* RayWorkItem */
			V3 rayO = mk(ro), rayD = mk(rd);
			int rayMedium = enableMedium ? cam->medium : -1;
			V3 ctxP = mk(0, 0, 0), ctxN = mk(0, 0, 0);
			Spec thp = sconst(1), pu = sconst(1), pl = sconst(1);
			int bsdfType = 0, depth = 0;
			bool alive = true;
			/* Interaction::getMedium(w), raytracing.h:162-166 */
			auto getMedium = [&](const SurfIntr &it, V3 w) -> int {
				const Mesh &m = s.meshes[it.mesh];
				if (m.mediumIn != m.mediumOut) return dot(w, it.n) > 0 ? m.mediumOut : m.mediumIn;
				return it.medium;
			};
			auto alphaAccept = [&](V3 o_, V3 d_) { return [&s, o_, d_](const Hit &c) { return !alphaKilled(s, c, o_, d_); }; };
			/* one shadow item: [2.5] traceShadow -> "Shadow" (device.cu:83-100) or "ShadowTr" (device.cu:102-128) */
			auto traceShadowItem = [&](V3 so, V3 sdir, int sMedium, Spec Ld, Spec spu, Spec spl, int loopDepth, bool capD, int auxLight) {
				st_.shadow++;
				if (loopDepth < KRR_MAX_DEPTH_STATS) st_.shadowByDepth[loopDepth]++;
				if (capD) cap.push(4, pixelId, depth, 0, auxLight);
				if (!enableMedium) {
					Hit sh = traceClosest(s, useBvh, so, sdir, 1.f, rtime, [&](const Hit &c) {
						/* __anyhit__Shadow: ignore null-material and alpha-killed hits */
						if (s.meshes[s.instances[c.inst].mesh].material < 0) return false;
						return !alphaKilled(s, c, so, sdir);
					});
					if (sh.inst < 0) L = (Ld / mean(spl + spu)) + L;
					return;
				}
				/* traceTransmittance, wavefront.h:80-139 (ratio tracking + null-interface stepping) */
				const V3 srO = so, srD = sdir; /* the work item's ray: CH(ShadowTr) prepares intr from IT (device.cu:102-109) */
				V3 ro_ = so, rd_ = sdir;
				int rmed = sMedium;
				const float tMax = 1.f;
				const V3 pLight = so + sdir * tMax;
				Spec T_ray = sconst(1), tpu = sconst(1), tpl = sconst(1);
				SurfIntr sit;
				sit.material = -1; /* SurfaceInteraction intr = {} */
				auto isZeroV = [](V3 v, float prec) { return std::fabs(v.x) <= prec && std::fabs(v.y) <= prec && std::fabs(v.z) <= prec; };
				while (!isZeroV(rd_, 1e-4f * 2)) {
					Hit sh = traceClosest(s, useBvh, ro_, rd_, tMax, rtime, alphaAccept(ro_, rd_));
					bool visible = sh.inst < 0;
					if (!visible) {
						prepareInteraction(s, sh, srD, rtime, lambda, lpdf, sit);
						sit.medium = sMedium;
					}
					if (!visible && sit.material >= 0) { T_ray = sconst(0); break; }
					if (rmed >= 0) {
						float tEnd = visible ? tMax : length(sit.p - ro_) / length(rd_);
						Spec T_maj = sampleT_maj(s.media[rmed], ro_, rd_, tEnd, &smp, lambda,
							[&](V3, MediumPoint mp, Spec sigma_maj, Spec T_maj_) -> bool {
								Spec sigma_n = cwiseMax0(sigma_maj - mp.sigma_a - mp.sigma_s);
								float pr = T_maj_.v[0] * sigma_maj.v[0];
								T_ray = T_ray * T_maj_ * sigma_n / pr;
								tpl	  = tpl * T_maj_ * sigma_maj / pr;
								tpu	  = tpu * T_maj_ * sigma_n / pr;
								Spec Tr = T_ray / mean(tpu + tpl);
								if (maxCoeff(Tr) < 0.05f) {
									if (ol_pcg_get1d(&smp) < 0.75f) T_ray = sconst(0);
									else T_ray = T_ray / 0.25f;
								}
								return any(T_ray);
							});
						T_ray = T_ray * T_maj / T_maj.v[0];
						tpu	  = tpu * T_maj / T_maj.v[0];
						tpl	  = tpl * T_maj / T_maj.v[0];
					}
					if (visible || !any(T_ray)) break;
					/* ray = intr.spawnRayTo(pLight), raytracing.h:151-155 */
					V3 p_o = offsetRayOrigin(sit.p, sit.n, pLight - sit.p);
					rd_	   = pLight - p_o;
					ro_	   = p_o;
					rmed   = getMedium(sit, rd_);
				}
				if (any(T_ray)) L = (Ld * T_ray / mean(spu * tpu + spl * tpl)) + L;
			};

			for (int loopDepth = 0; alive; loopDepth++) {
				const bool capD = capS && capDepth == loopDepth;
				if (capD) cap.push(0, pixelId, depth, bsdfType, -1);
				/* [2.1] traceClosest: device.cu:43-81 */
				st_.closest++;
				if (loopDepth < KRR_MAX_DEPTH_STATS) st_.closestByDepth[loopDepth]++;
				Hit h = traceClosest(s, useBvh, rayO, rayD, std::numeric_limits<float>::infinity(), rtime, alphaAccept(rayO, rayD));
				if (loopDepth == 0) fhInst = h.inst, fhPrim = h.prim;
				alive = false; /* the ray item is consumed */
				bool haveScatter = false, haveHitLight = false, haveMiss = false, haveMediumScatter = false;
				SurfIntr it;
				if (h.inst >= 0) {
					prepareInteraction(s, h, rayD, rtime, lambda, lpdf, it);
					it.medium = rayMedium;
				}
				/* MediumScatterWorkItem */
				V3 msP = mk(0, 0, 0), msWo = mk(0, 0, 0);
				Spec msThp = sconst(0), msPu = sconst(0);
				int msMedium = -1;
				auto passNull = [&]() {
					/* null material: same-depth re-push of intr.spawnRayTowards(ray.dir), device.cu:54-58 / medium.cpp:88-93 */
					rayO	  = offsetRayOrigin(it.p, it.n, rayD);
					rayMedium = enableMedium ? getMedium(it, rayD) : -1;
					alive	  = true;
					if (capD) cap.push(5, pixelId, depth, bsdfType, -1);
				};
				if (enableMedium && rayMedium >= 0) {
					/* [2.2] sampleMediumInteraction, medium.cpp:13-103 */
					st_.mediumSample++;
					const MediumData &med = s.media[rayMedium];
					const float tMaxM = h.inst >= 0 ? h.t : std::numeric_limits<float>::infinity();
					Spec Lm = sconst(0);
					bool scattered = false;
					Spec T_maj = sampleT_maj(med, rayO, rayD, tMaxM, &smp, lambda,
						[&](V3 pm, MediumPoint mp, Spec sigma_maj, Spec T_maj_) -> bool {
							if (depth < p->max_depth && any(mp.Le)) {
								float pr = sigma_maj.v[0] * T_maj_.v[0];
								Spec pe	 = pu * sigma_maj * T_maj_ / pr;
								if (any(pe)) Lm = Lm + thp * mp.sigma_a * T_maj_ * mp.Le / (pr * mean(pe));
							}
							float pAbsorb  = mp.sigma_a.v[0] / sigma_maj.v[0];
							float pScatter = mp.sigma_s.v[0] / sigma_maj.v[0];
							float pNull	   = std::max(0.f, 1.f - pAbsorb - pScatter);
							int mode	   = sampleDiscrete3(pAbsorb, pScatter, pNull, ol_pcg_get1d(&smp));
							if (mode == 0) {
								thp = sconst(0);
								return false;
							} else if (mode == 1) {
								float pr = T_maj_.v[0] * mp.sigma_s.v[0];
								thp		 = thp * T_maj_ * mp.sigma_s / pr;
								pu		 = pu * T_maj_ * mp.sigma_s / pr;
								if (any(thp) && any(pu)) {
									haveMediumScatter = true;
									msP = pm, msThp = thp, msPu = pu, msWo = -rayD, msMedium = rayMedium;
								}
								scattered = true;
								return false;
							} else {
								Spec sigma_n = cwiseMax0(sigma_maj - mp.sigma_a - mp.sigma_s);
								float pr	 = T_maj_.v[0] * sigma_n.v[0];
								thp			 = thp * T_maj_ * sigma_n / pr;
								if (pr == 0) thp = sconst(0);
								pu = pu * T_maj_ * sigma_n / pr;
								pl = pl * T_maj_ * sigma_maj / pr;
								return any(thp) && any(pu);
							}
						});
					if (any(Lm)) L = Lm + L;
					if (!scattered && any(thp)) {
						thp = thp * T_maj / T_maj.v[0];
						pu	= pu * T_maj / T_maj.v[0];
						pl	= pl * T_maj / T_maj.v[0];
					}
					if (haveMediumScatter) st_.mediumScatter++;
					if (scattered || !any(thp) || !any(pu) || depth == p->max_depth) {
						/* nothing else leaves this item */
					} else if (h.inst < 0) {
						haveMiss = true;
					} else if (it.material < 0) {
						passNull();
					} else {
						if (it.light >= 0) haveHitLight = true;
						haveScatter = true;
						st_.scatter++;
						if (capD) cap.push(3, pixelId, depth, ol_bsdf_type(&it.sd), it.sd.bsdfType);
					}
				} else if (h.inst < 0) {
					haveMiss = true;
				} else if (it.material < 0) {
					passNull();
				} else {
					if (it.light >= 0) haveHitLight = true;
					if (any(thp)) {
						haveScatter = true;
						st_.scatter++;
						if (capD) cap.push(3, pixelId, depth, ol_bsdf_type(&it.sd), it.sd.bsdfType);
					}
				}
				/* [2.3] handleHit, integrator.cpp:78-90 */
				if (haveHitLight) {
					st_.hitLight++;
					if (capD) cap.push(2, pixelId, depth, bsdfType, it.light);
					OlTriLight tl;
					fillTri(s, s.lights[it.light], tl);
					float pp[3], nn[3], ww[3], Le4[4];
					st(pp, it.p), st(nn, it.n), st(ww, it.wo);
					ol_arealight_L(&tl, pp, nn, ww, lambda, Le4);
					Spec Le = Spec{{Le4[0], Le4[1], Le4[2], Le4[3]}} * thp;
					if (p->nee && depth && !(bsdfType & (32 | 1))) { /* BSDF_DELTA = SPECULAR|NULL */
						float cp[3], cn[3];
						st(cp, ctxP), st(cn, ctxN);
						float lightPdf = ol_arealight_pdf_li(&tl, pp, nn, cp, cn) * (1.f / nLights);
						Le = Le / mean(pl * lightPdf + pu);
					} else Le = Le / mean(pu);
					L = Le + L; /* addRadiance: L[pixel] = L_val + L[pixel], workqueue.h:23-25 */
				}
				/* handleMiss, integrator.cpp:92-108 */
				if (haveMiss) {
					st_.miss++;
					if (capD) cap.push(1, pixelId, depth, bsdfType, -1);
					Spec Lm = sconst(0);
					for (int li : s.infinite) {
						float w[3], Li4[4];
						st(w, rayD);
						infiniteLightLi(s.analytic[li], s.analyticTex[li], w, lambda, Li4);
						Spec Li = Spec{{Li4[0], Li4[1], Li4[2], Li4[3]}};
						if (p->nee && depth && !(bsdfType & (32 | 1))) {
							float lightPdf = 0.07957747154594767f /* M_INV_4PI */ * (1.f / nLights);
							Lm = Lm + Li / mean(pu + pl * lightPdf);
						} else Lm = Lm + Li / mean(pu);
					}
					L = (thp * Lm) + L;
				}
				if (loopDepth == p->max_depth) break;
				/* light sampling shared by the two scattering stages: lightSampler.sample + light.sampleLi */
				auto sampleLight = [&](V3 cpv, V3 cnv, float lp[3], float ln[3], float Ll[4], float &lightPdf, bool &delta, int &lightId) {
					float u1 = ol_pcg_get1d(&smp);
					lightId	 = (int) (uint32_t) (u1 * nLights);
					const LightRef &lr = s.lights[lightId];
					float u2[2];
					u2[0] = ol_pcg_get1d(&smp), u2[1] = ol_pcg_get1d(&smp);
					float cp[3], cn[3], lpdfv;
					st(cp, cpv), st(cn, cnv);
					ln[0] = ln[1] = ln[2] = 0;
					delta = lr.type != KRR_LIGHT_DIFFUSE_AREA && lr.type != KRR_LIGHT_INFINITE;
					if (lr.type == KRR_LIGHT_DIFFUSE_AREA) {
						OlTriLight tl;
						fillTri(s, lr, tl);
						ol_arealight_sample_li(&tl, u2, cp, cn, lambda, lp, ln, Ll, &lpdfv);
					} else {
						ol_light_sample_li(&s.analytic[lr.analytic], u2, cp, lambda, lp, Ll, &lpdfv);
						if (s.analyticTex[lr.analytic].image) {
							/* sampleLi (light.h:222-231): p = ctx.p + wi * 2 * sceneRadius with wi uniform on the sphere; the
							 * image is looked up in that direction */
							float wi3[3] = {lp[0] - cp[0], lp[1] - cp[1], lp[2] - cp[2]};
							infiniteLightLi(s.analytic[lr.analytic], s.analyticTex[lr.analytic], wi3, lambda, Ll);
						}
					}
					lightPdf = (1.f / nLights) * lpdfv;
				};
				/* [2.4a] sampleMediumScattering, medium.cpp:105-153 */
				if (haveMediumScatter) {
					const float g = s.media[msMedium].g;
					float wo3[3];
					st(wo3, msWo);
					if (p->nee) {
						float lp[3], ln[3], Ll[4], lightPdf;
						bool delta;
						int lightId;
						sampleLight(msP, mk(0, 0, 0), lp, ln, Ll, lightPdf, delta, lightId);
						/* Interaction(p, time, medium).spawnRayTo(ls.intr): n = 0, so the origin is not offset */
						V3 lP = mk(lp), lN = mk(ln);
						V3 to  = offsetRayOrigin(lP, lN, msP - lP);
						V3 sd_ = to - msP;
						V3 wi  = normalize(sd_);
						float wi3[3];
						st(wi3, wi);
						float ph	   = ol_hg_p(g, wo3, wi3);
						Spec thpL	   = msThp * ph;
						float phasePdf = delta ? 0 : ph; /* HGPhaseFunction::pdf == p */
						Spec Ld		   = thpL * Spec{{Ll[0], Ll[1], Ll[2], Ll[3]}};
						if (any(Ld) && lightPdf > 0) traceShadowItem(msP, sd_, msMedium, Ld, msPu * phasePdf, msPu * lightPdf, loopDepth, capD, lightId);
					}
					float u2[2], wi3[3], php, phpdf;
					u2[0] = ol_pcg_get1d(&smp), u2[1] = ol_pcg_get1d(&smp);
					ol_hg_sample(g, wo3, u2, wi3, &php, &phpdf);
					Spec nthp	 = msThp * php / phpdf;
					float rrProb = maxCoeff(nthp / mean(msPu));
					bool killed	 = false;
					if (depth >= 1 && rrProb < 1) {
						if (ol_pcg_get1d(&smp) >= rrProb) killed = true;
						else nthp = nthp / rrProb;
					}
					if (!killed && any(nthp) && !hasNaN(nthp)) {
						rayO = msP, rayD = mk(wi3), rayMedium = msMedium;
						ctxP = msP, ctxN = mk(0, 0, 0);
						thp = nthp, pu = msPu, pl = msPu / phpdf;
						depth	 = depth + 1;
						bsdfType = 8 | 16; /* BSDF_SMOOTH */
						alive	 = true;
						if (capD) cap.push(5, pixelId, depth, bsdfType, -1);
					}
				}
				/* [2.4] generateScatterRays, integrator.cpp:110-164 */
				if (haveScatter) {
					if (ol_pcg_get1d(&smp) >= p->rr) continue; /* alive == false: path ends */
					thp = thp / p->rr;
					V3 woLocal	  = toLocal(it, it.wo);
					int bt		  = ol_bsdf_type(&it.sd);
					float wo3[3];
					st(wo3, woLocal);
					if (p->nee && (bt & (8 | 16))) { /* BSDF_SMOOTH = DIFFUSE | GLOSSY */
						float lp[3], ln[3], Ll[4], lightPdf;
						bool delta;
						int lightId;
						sampleLight(it.p, it.n, lp, ln, Ll, lightPdf, delta, lightId);
						/* spawnRayTo(ls.intr): raytracing.h:148-157 */
						V3 lP = mk(lp), lN = mk(ln);
						V3 to  = offsetRayOrigin(lP, lN, it.p - lP);
						V3 p_o = offsetRayOrigin(it.p, it.n, to - it.p);
						V3 sd_ = to - p_o;
						V3 wiWorld = normalize(sd_);
						V3 wiLocal = toLocal(it, wiWorld);
						float wi3[3], f4[4], bpdf;
						st(wi3, wiLocal);
						ol_bsdf_f_pdf(&it.sd, wo3, wi3, f4, &bpdf);
						Spec bsdfVal = Spec{{f4[0], f4[1], f4[2], f4[3]}};
						float bsdfPdf = delta ? 0 : bpdf;
						if (lightPdf > 0 && any(bsdfVal)) {
							Spec Ld = Spec{{Ll[0], Ll[1], Ll[2], Ll[3]}} * thp * bsdfVal * std::fabs(wiLocal.z);
							Spec spu = pu * bsdfPdf, spl = pu * lightPdf;
							if (any(Ld)) traceShadowItem(p_o, sd_, enableMedium ? getMedium(it, sd_) : -1, Ld, spu, spl, loopDepth, capD, lightId);
						}
					}
					/* sample BSDF */
					float f4[4], wi3[3], spdf;
					int flags;
					ol_bsdf_sample(&it.sd, wo3, &smp, f4, wi3, &spdf, &flags);
					Spec sf = Spec{{f4[0], f4[1], f4[2], f4[3]}};
					if (spdf != 0 && any(sf)) {
						V3 wiWorld = toWorld(it, mk(wi3));
						Spec nthp  = thp * sf * std::fabs(wi3[2]) / spdf;
						if (any(nthp)) {
							bsdfType = flags;
							pl		 = pu / spdf;
							/* pu unchanged */
							rayO	  = offsetRayOrigin(it.p, it.n, wiWorld);
							rayD	  = wiWorld;
							rayMedium = enableMedium ? getMedium(it, wiWorld) : -1;
							ctxP = it.p, ctxN = it.n;
							depth	 = depth + 1;
							thp		 = nthp;
							alive	 = true;
							if (capD) cap.push(5, pixelId, depth, bsdfType, -1);
						}
					}
				}
			}
			/* per-sample resolve, integrator.cpp:257-260: pixel += L.toRGB (L is NOT reset between
			 * samples of one frame -- it keeps accumulating, so sample k adds the running sum) */
			float rgb[3];
			ol_to_rgb(L.v, lambda, lpdf, rgb);
			for (int k = 0; k < 3; k++) pixel[k] += rgb[k];
		}
		if (firstHits) firstHits[2 * pixelId] = fhInst, firstHits[2 * pixelId + 1] = fhPrim;
		/* film write, integrator.cpp:262-266 + y flip of CudaRenderTarget::write (cuda.h:33-36) */
		if (film) {
			float o[4] = {pixel[0] / float(spp), pixel[1] / float(spp), pixel[2] / float(spp), 1.f};
			if (p->enable_clamp) for (int k = 0; k < 3; k++) o[k] = std::min(std::max(o[k], 0.f), p->clamp_max);
			int x = pixelId % W, y = pixelId / W;
			memcpy(film + 4 * ((size_t) (H - 1 - y) * W + x), o, 16);
		}
	}
	}
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < nthreads; t++) pool.emplace_back(worker, t);
	worker(0);
	for (auto &t : pool) t.join();
	double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
	if (stats) {
		memset(stats, 0, sizeof(*stats));
		for (auto &t : tstats) {
			stats->camera_rays += t.camera, stats->closest_rays += t.closest, stats->shadow_rays += t.shadow;
			stats->scatter_items += t.scatter, stats->hit_light_items += t.hitLight, stats->miss_items += t.miss;
			stats->medium_sample_items += t.mediumSample, stats->medium_scatter_items += t.mediumScatter;
			for (int k = 0; k < KRR_MAX_DEPTH_STATS; k++)
				stats->closest_by_depth[k] += t.closestByDepth[k], stats->shadow_by_depth[k] += t.shadowByDepth[k];
		}
		stats->bvh_triangles = s.bvhPrims.size(), stats->bvh_nodes = s.bvh.size();
	}
	return secs;
}

/* =================================================================================================
 * MegakernelPathTracer, restated from the reference's OptiX programs (src/render/megakernel/device.cu):
 * __raygen__Pathtracer :148-195, handleHit :50-64, handleMiss :66-79, generateShadowRay / evalDirect :81-113,
 * generateScatterRay :115-127; power-heuristic MIS (render/sampling.h:20-28) on (bsdf pdf, light pdf) instead of the
 * wavefront pass's spectral path pdfs.  It is the independent estimator the GPU megakernel is checked against
 * (tests/test_gpu_megakernel_oracle.py).  The reference draws from a cuRAND sampler there (device code); this
 * restatement and the GPU kernel both use the PCG sampler with the reference's seeding of that pass,
 * setPixelSample(pixel, frameID * 512) (device.cu:159), so their streams are identical.
 * A hit without a material (medium interface) ends the path: the megakernel has no medium handling. */
namespace {
inline float evalMIS(float n0, float p0, float n1, float p1) { /* sampling.h:20-28 */
	float q0 = (n0 * p0) * (n0 * p0), q1 = (n1 * p1) * (n1 * p1);
	return q0 / (q0 + q1);
}
} // namespace

extern "C" double orc_render_megakernel(const OrcScene *sp, const OrcParams *p, const KrrCameraData *cam, int32_t W, int32_t H,
										uint64_t frameId, float *film) {
	const OrcScene &s = *sp;
	OlCamera oc;
	memcpy(oc.filmSize, cam->film_size, 8);
	oc.focalLength = cam->focal_length, oc.focalDistance = cam->focal_distance, oc.lensRadius = cam->lens_radius;
	oc.aspectRatio = cam->aspect_ratio, oc.shutterOpen = cam->shutter_open, oc.shutterTime = cam->shutter_time;
	memcpy(oc.transform, cam->transform, 48);
	const int spp = p->spp > 0 ? p->spp : 1, nLights = (int) s.lights.size();
	const bool useBvh = p->use_bvh != 0;
	auto t0 = std::chrono::steady_clock::now(); /* (the per-render TLAS build over the moving instances is part of the timed work) */
	/* use_bvh == 2: moving instances through a BVH over their boxes for this render's ray-time window */
	struct MovingTlasScope {
		const OrcScene &sc;
		MovingTlasScope(const OrcScene &sc_, bool on, float w0, float w1) : sc(sc_) { if (on) buildMovingTlas(sc, w0, w1); else sc.useMovingTlas = false; }
		~MovingTlasScope() { sc.useMovingTlas = false; }
	} movingTlasScope(s, p->use_bvh == 2, cam->shutter_open, cam->shutter_open + cam->shutter_time);
	const float lightSelPdf = nLights > 0 ? 1.f / nLights : 0.f;
	const int BSDF_SPECULAR_ = 32, BSDF_SMOOTH_ = 8 | 16;
	int nthreads = p->threads > 0 ? p->threads : (int) std::max(1u, std::thread::hardware_concurrency());
	std::atomic<int> nextChunk{0};
	const int chunk = 256, pixelEnd = W * H;
	auto worker = [&]() {
		for (;;) {
			const int c0 = nextChunk.fetch_add(chunk);
			if (c0 >= pixelEnd) break;
			for (int pixelId = c0; pixelId < std::min(c0 + chunk, pixelEnd); pixelId++) {
				const int px = pixelId % W, py = pixelId / W;
				OlSampler smp;
				ol_pcg_set_pixel_sample(&smp, px, py, (uint32_t) frameId * 512u); /* device.cu:159 */
				float color[3] = {0, 0, 0};
				for (int i = 0; i < spp; i++) { /* device.cu:166-191 */
					Spec thp = sconst(1), L = sconst(0);
					float cs[5];
					for (int k = 0; k < 5; k++) cs[k] = ol_pcg_get1d(&smp);
					float ro[3], rd[3], rtime;
					ol_camera_ray(&oc, px, py, W, H, cs, ro, rd, &rtime);
					float lambda[4], lpdf[4];
					ol_sample_wavelengths(ol_pcg_get1d(&smp), lambda, lpdf);
					V3 rayO = mk(ro), rayD = mk(rd), ctxP = mk(0, 0, 0), ctxN = mk(0, 0, 0);
					float pdfPrev = 0;
					int typePrev  = 0;
					for (int depth = 0;; depth++) {
						Hit h = traceClosest(s, useBvh, rayO, rayD, std::numeric_limits<float>::infinity(), rtime,
											 [&](const Hit &c) { return !alphaKilled(s, c, rayO, rayD); }); /* __anyhit__Radiance */
						if (h.inst < 0) { /* handleMiss */
							for (int li : s.infinite) {
								float weight = 1;
								if (p->nee && depth > 0 && !(typePrev & BSDF_SPECULAR_)) {
									weight = evalMIS(1, pdfPrev, 1, 0.07957747154594767f * lightSelPdf);
									if (std::isnan(weight) || std::isinf(weight)) weight = 1;
								}
								float w[3], Li4[4];
								st(w, rayD);
								infiniteLightLi(s.analytic[li], s.analyticTex[li], w, lambda, Li4);
								L = L + thp * weight * Spec{{Li4[0], Li4[1], Li4[2], Li4[3]}};
							}
							break;
						}
						SurfIntr it;
						prepareInteraction(s, h, rayD, rtime, lambda, lpdf, it);
						if (it.material < 0) break;
						if (it.light >= 0) { /* handleHit */
							OlTriLight tl;
							fillTri(s, s.lights[it.light], tl);
							float pp[3], nn[3], ww[3], Le4[4];
							st(pp, it.p), st(nn, it.n), st(ww, it.wo);
							ol_arealight_L(&tl, pp, nn, ww, lambda, Le4);
							float weight = 1;
							if (p->nee && depth > 0) {
								float cp[3], cn[3];
								st(cp, ctxP), st(cn, ctxN);
								float lightPdf = ol_arealight_pdf_li(&tl, pp, nn, cp, cn) * lightSelPdf;
								if (!(typePrev & BSDF_SPECULAR_)) weight = evalMIS(1, pdfPrev, 1, lightPdf);
								if (std::isnan(weight) || std::isinf(weight)) weight = 1;
							}
							L = L + Spec{{Le4[0], Le4[1], Le4[2], Le4[3]}} * weight * thp;
						}
						if (depth == p->max_depth || (p->rr < 1.f && ol_pcg_get1d(&smp) > p->rr)) break;
						thp = thp / p->rr;
						V3 woLocal = toLocal(it, it.wo);
						float wo3[3];
						st(wo3, woLocal);
						if (p->nee && (ol_bsdf_type(&it.sd) & BSDF_SMOOTH_) && nLights > 0) { /* evalDirect -> generateShadowRay */
							float u1	= ol_pcg_get1d(&smp);
							int lightId = std::min((int) (uint32_t) (u1 * nLights), nLights - 1);
							const LightRef &lr = s.lights[lightId];
							float u2[2], cp[3], cn[3], lp[3], ln[3] = {0, 0, 0}, Ll[4], lpdfv;
							u2[0] = ol_pcg_get1d(&smp), u2[1] = ol_pcg_get1d(&smp);
							st(cp, it.p), st(cn, it.n);
							bool delta = lr.type != KRR_LIGHT_DIFFUSE_AREA && lr.type != KRR_LIGHT_INFINITE;
							if (lr.type == KRR_LIGHT_DIFFUSE_AREA) {
								OlTriLight tl;
								fillTri(s, lr, tl);
								ol_arealight_sample_li(&tl, u2, cp, cn, lambda, lp, ln, Ll, &lpdfv);
							} else {
								ol_light_sample_li(&s.analytic[lr.analytic], u2, cp, lambda, lp, Ll, &lpdfv);
								if (s.analyticTex[lr.analytic].image) {
									float wi3[3] = {lp[0] - cp[0], lp[1] - cp[1], lp[2] - cp[2]};
									infiniteLightLi(s.analytic[lr.analytic], s.analyticTex[lr.analytic], wi3, lambda, Ll);
								}
							}
							V3 lP = mk(lp), lN = mk(ln);
							V3 wiLocal = toLocal(it, normalize(lP - it.p));
							float lightPdf = lightSelPdf * lpdfv;
							if (lightPdf != 0) {
								float wi3[3], f4[4], bpdf;
								st(wi3, wiLocal);
								ol_bsdf_f_pdf(&it.sd, wo3, wi3, f4, &bpdf);
								float bsdfPdf = delta ? 0.f : bpdf;
								Spec bsdfVal  = Spec{{f4[0], f4[1], f4[2], f4[3]}} * std::fabs(wiLocal.z);
								float mis	  = evalMIS(1, lightPdf, 1, bsdfPdf);
								if (!(std::isnan(mis) || std::isinf(mis)) && any(bsdfVal)) {
									V3 to  = offsetRayOrigin(lP, lN, it.p - lP); /* spawnRayTo(ls.intr) */
									V3 p_o = offsetRayOrigin(it.p, it.n, to - it.p);
									V3 sd_ = to - p_o;
									Hit sh = traceClosest(s, useBvh, p_o, sd_, 1.f, rtime, [&](const Hit &c) { return !alphaKilled(s, c, p_o, sd_); });
									if (sh.inst < 0) L = L + thp * bsdfVal * mis / (1 * lightPdf) * Spec{{Ll[0], Ll[1], Ll[2], Ll[3]}};
								}
							}
						}
						/* generateScatterRay */
						float f4[4], wi3[3], spdf;
						int flags;
						ol_bsdf_sample(&it.sd, wo3, &smp, f4, wi3, &spdf, &flags);
						Spec sf = Spec{{f4[0], f4[1], f4[2], f4[3]}};
						if (spdf == 0 || !any(sf)) break;
						V3 wiWorld = toWorld(it, mk(wi3));
						typePrev = flags, pdfPrev = spdf;
						rayO = offsetRayOrigin(it.p, it.n, wiWorld), rayD = wiWorld;
						ctxP = it.p, ctxN = it.n;
						thp	 = thp * sf * std::fabs(wi3[2]) / spdf;
						if (!any(thp)) break;
					}
					float rgb[3];
					ol_to_rgb(L.v, lambda, lpdf, rgb);
					for (int k = 0; k < 3; k++) color[k] += rgb[k];
				}
				/* colorBuffer.write(RGBA(color, 1), fbIndex): the SUM over the samples (device.cu:194), row H-1-y */
				float o[4] = {color[0], color[1], color[2], 1.f};
				memcpy(film + 4 * ((size_t) (H - 1 - py) * W + px), o, 16);
			}
		}
	};
	std::vector<std::thread> pool;
	for (int t = 1; t < nthreads; t++) pool.emplace_back(worker);
	worker();
	for (auto &t : pool) t.join();
	return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
