// wavefront_kernels.cuh -- the stage kernels of the B200-native WavefrontPathTracer.
//
// Stage map (reference src/render/wavefront/integrator.cpp:223-267 loop; device.cu programs):
//   k_begin_frame          beginFrame                 integrator.cpp:205-221
//   k_generate_camera_rays generateCameraRays         integrator.cpp:166-179
//   k_trace_closest        __raygen__/CH/AH/MS Closest device.cu:43-81
//   k_handle_hit_miss      handleHit + handleMiss     integrator.cpp:78-108
//   k_scatter<MT>          generateScatterRays        integrator.cpp:110-164 (+ prepareSurfaceInteraction,
//                                                     shading.h:115-226, moved here from the CH program)
//   k_trace_shadow         __raygen__/AH/MS Shadow    device.cu:83-100
//   k_resolve / k_film     per-sample resolve, film   integrator.cpp:257-266, device/cuda.h:33-45
//
// Data layout (HBM), all arrays 16-byte vectorised SoA:
//   * ray queues (ping-pong): 7 x float4 per item = 112 B (reference RayWorkItem: 120 B SoA)
//   * the trace stage writes a 16-byte hit record next to the ray and pushes the SLOT INDEX (4 B)
//     into the miss / hit-light / per-material scatter index queues; the 232-byte ScatterRayWorkItem
//     of the reference is never materialised -- the scatter stage rebuilds the interaction from
//     (instance, primitive, barycentrics), which are L2-resident mesh reads
//   * scatter indices are binned by material type at push time (one counter per type), so the
//     scatter kernel of each type runs a single BSDF: this is the "material sort", done for free
//   * shadow queue: 3 x float4 = 48 B (ray, pixel, and the MIS-weighted contribution
//     Ld/(pl+pu).mean() pre-divided), reference 92 B
//   * per pixel: L (16 B) + rgb (16 B) + PCG state (8 B, inc is per-frame uniform) + lambda0 (4 B)
//   * queue sizes live in a per-depth counter block on the device; kernels are launched with a fixed
//     persistent grid (multiple of the SM count) and grid-stride over the live count, so there is
//     no host synchronisation and no full-width launch over a nearly empty queue
//   * pushes are warp-aggregated: one atomicAdd per warp per queue
#pragma once
#include <cub/block/block_radix_sort.cuh>

#include "bsdf.cuh"
#include "bvh.cuh"
#include "lights.cuh"
#include "medium.cuh"
#include "sampler.cuh"
#include "scene.cuh"
#include "spectrum.cuh"

namespace krr {

constexpr int kMaxDepthSlots = 66;

struct DepthCounters {	// one block per loop depth; 16 ints = 64 B
	int32_t nRay;		// items in the ray queue consumed at this depth
	int32_t nMiss, nHitLight;
	int32_t nScatter[MAT_COUNT];
	int32_t nShadow;
	int32_t nMediumSample, nMediumScatter;
	int32_t cursorRay, cursorShadow; // work cursors of the persistent trace kernels (next unclaimed item)
	int32_t nScatterKilled;			 // scatter items terminated by Russian roulette in the closest stage (rrInTrace)
	int32_t pad[2];
};
static_assert(sizeof(DepthCounters) == 64, "DepthCounters");

struct RayQueue { // SoA, 7 x 16 B
	float4 *o_time;	  // origin.xyz, ray time
	float4 *d_medium; // dir.xyz, medium index (int bits, -1 = none)
	float4 *thp, *pu, *pl;
	float4 *ctxP_pix; // LightSampleContext::p, pixelId (int bits)
	float4 *ctxN_dep; // LightSampleContext::n, depth | bsdfType << 8 (int bits)
};

struct ShadowQueue { // 3 x 16 B (+2 for the transmittance variant)
	float4 *o_tmax;
	float4 *d_pix;
	float4 *contrib; // Ld / (pl + pu).mean()   (Ld itself when media are enabled)
	float4 *pu, *pl; // media only
	int2 *aux;		 // media only: medium index of the ray, ray time (float bits)
};

struct MediumScatterQueue { // MediumScatterWorkItem (workitem.h:85-95), media only: 4 x 16 B + 8 B
	float4 *p_time;	   // scattering point, ray time
	float4 *wo_medium; // wo = -ray.dir, medium index (int bits)
	float4 *thp, *pu;
	int2 *pix_depth;
};

struct PixelState {
	float4 *L;
	float4 *pixel;
	uint64_t *rng;
	float *lambda;		 // lambda[0], sign bit = secondary wavelengths terminated
	float *cameraSample; // 5 floats per pixel, debug builds of the state only (may be null)
};

struct Params {
	int32_t width, height;
	int32_t pixelBegin, partPixels; // this handle's pixel range (row partition of the frame)
	// the band of the partition these queues serve: rows rowPhase, rowPhase + rowStride, ... of the
	// partition, pixelCount pixels in all, indexed densely (local index i = bandRow * width + x)
	int32_t rowStride, rowPhase, pixelCount;
	// FRAME BATCH ("frame_batch": F): the launches of one render() carry F independent frames -- frame indices
	// frameIndex .. frameIndex + F - 1, each with its own PCG sequence (sampleIndex = frameIndex * spp, integrator.cpp:
	// 217) -- as F layers of pixel state: local index i = layer * layerPixels + (pixel of the band), pixelCount =
	// F * layerPixels.  A stage launch lasts as long as its slowest ray, whatever the queue holds; with F frames in
	// the same launches that latency is paid once per F frames.  The film is the mean of the F frames' films.
	int32_t layers, layerPixels;
	uint32_t seedIndex; // frameIndex * spp of layer 0
	int32_t spp, maxDepth, nee, enableMedium, enableClamp;
	// Russian roulette of generateScatterRays (integrator.cpp:118-119) evaluated by the closest stage
	// when it routes a hit to the scatter queue: it is the NEXT draw of the pixel's stream either way,
	// so films are identical; terminated paths never occupy a lane of the scatter kernel
	int32_t rrInTrace;
	float probRR, clampMax;
	uint64_t rngInc; // PCG increment of this frame
	uint32_t sampleIndex;
	int32_t refill; // kRefill or kRefillFlat
	// The camera kernel only writes origin and direction: every other field of a depth-0 ray item is a
	// constant (thp = pu = pl = 1, ctx = 0, depth 0) or the slot itself (pixel = slot: the queue is dense), so
	// the stages of loop depth 0 substitute them instead of reading 80 B per ray that were never worth
	// writing (-166 MB of stores per 1080p sample).  Off with participating media (the medium stage updates
	// thp / pu / pl in place) and while a debug capture is armed.
	int32_t implicitDepth0;
	int32_t sortKey; // k_sort_rays: 0 = direction octant above the origin's Morton code, 1 = Morton code above the octant
};

struct KrrCameraDev {
	float filmSize[2];
	float focalLength, focalDistance, lensRadius, aspectRatio, shutterOpen, shutterTime;
	Xf transform;
	int32_t medium;
	float tanFov; // tan(atan2(filmSize[1] * 0.5, focalLength)) evaluated ONCE on the host (glibc), see api
};

struct Wavefront {
	Params p;
	KrrCameraDev cam;
	SceneDev scene;
	BvhDev bvh;
	PixelState px;
	RayQueue rays[2];
	int4 *hits;			// per ray slot of the CURRENT queue: inst, prim, u bits, v bits
	float *hitT;		// media only: hit distance per ray slot (tMax of the medium-sample item)
	int32_t *missIdx, *hitLightIdx, *mediumSampleIdx;
	MediumScatterQueue mscatter;
	int32_t *scatterIdx[MAT_COUNT];
	ShadowQueue shadow;
	DepthCounters *counters; // [kMaxDepthSlots]
	int4 *firstHits;		 // per pixel depth-0 hit (debug tap), may be null
	int32_t *errorFlags;	 // [0] traversal stack overflow
	// per instance: bit0 null material, bit1 has alpha (transmission) texture, bit2 emissive, bits 4-6 BSDF
	// type of its material -- everything the closest stage needs to route a hit, in one load
	const uint8_t *instFlags;
	// RAY REORDERING ("sort_rays"): the trace stage of depth >= 1 takes its rays through these permutations of the
	// queue slots (null = queue order).  perm[k] = slot of the k-th ray in traversal order; see k_sort_rays
	const int32_t *permClosest, *permShadow;
	int32_t *tripHist; // KRR_COUNT_TRIPS builds: histogram of node visits per closest ray (64 bins of 8, 64-bit counts)
};

// Programmatic dependent launch: every stage kernel lets its successor in the stream be scheduled at
// once (its CTAs become resident as ours retire and block in griddepcontrol.wait) and then waits itself
// for the complete, flushed results of its predecessor -- the launch latency and the CTA ramp-up of a
// stage overlap the tail of the previous one.  Without the launch attribute both are no-ops.
#define KRR_PDL_ENTRY() asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")

// ---- warp-aggregated push: one atomicAdd per warp ----
KRR_DEV int warpPush(int32_t *counter, bool pred) {
	unsigned mask = __ballot_sync(__activemask(), pred);
	if (!pred) return -1;
	int lane   = threadIdx.x & 31;
	int leader = __ffs(mask) - 1;
	int base   = 0;
	if (lane == leader) base = atomicAdd(counter, __popc(mask));
	base = __shfl_sync(mask, base, leader);
	return base + __popc(mask & ((1u << lane) - 1));
}

// same, for code that runs with the whole warp converged (persistent trace kernels)
KRR_DEV int warpPushFull(int32_t *counter, bool pred) {
	unsigned mask = __ballot_sync(0xffffffffu, pred);
	if (!mask) return -1;
	int lane   = threadIdx.x & 31;
	int leader = __ffs(mask) - 1;
	int base   = 0;
	if (lane == leader) base = atomicAdd(counter, __popc(mask));
	base = __shfl_sync(0xffffffffu, base, leader);
	return pred ? base + __popc(mask & ((1u << lane) - 1)) : -1;
}

// local pixel index of a band -> pixel id in the frame (the reference's pixelId: RNG seed, film position)
KRR_DEV int layerOf(const Params &p, int i) { return p.layers == 1 ? 0 : i / p.layerPixels; }
// PCG increment of the frame that local index i belongs to (PCGSampler::setSeed, sampler.h:22-28: inc = sequence << 1 | 1)
KRR_DEV uint64_t rngIncOf(const Params &p, int i) {
	return p.layers == 1 ? p.rngInc : (((uint64_t) (p.seedIndex + (uint32_t) (layerOf(p, i) * p.spp)) << 1u) | 1u);
}
KRR_DEV int framePixel(const Params &p, int i) {
	if (p.layers != 1) i -= layerOf(p, i) * p.layerPixels;
	if (p.rowStride == 1) return p.pixelBegin + i;
	int row = i / p.width;
	return p.pixelBegin + (row * p.rowStride + p.rowPhase) * p.width + (i - row * p.width);
}

KRR_DEV float4 ldg4(const float4 *p) { return __ldg(p); }
// streaming (read-once / write-once) queue traffic: keep it out of L1
KRR_DEV float4 ldcs4(const float4 *p) { return __ldcs(p); }
KRR_DEV void stcs4(float4 *p, float4 v) { __stcs(p, v); }

// =================================================================================================
__global__ void k_begin_frame(const __grid_constant__ Wavefront wf, uint32_t seedIndex) {
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < wf.p.pixelCount; i += gridDim.x * blockDim.x) {
		int pixelId = framePixel(wf.p, i);
		int px = pixelId % wf.p.width, py = pixelId / wf.p.width;
		wf.px.L[i]	   = make_float4(0, 0, 0, 0);
		wf.px.pixel[i] = make_float4(0, 0, 0, 0);
		Pcg rng;
		rng.setPixelSample((uint32_t) px, (uint32_t) py, seedIndex + (uint32_t) (layerOf(wf.p, i) * wf.p.spp));
		rng.advance((int64_t) (256 * pixelId));
		wf.px.lambda[i] = sampleLambda0(rng.get1D());
		wf.px.rng[i]	= rng.state;
	}
}

// CameraData::generateSample + getRay, src/core/camera.h:32-58.  Bit-exact against the oracle: every
// operation individually rounded, in Eigen's evaluation order.
KRR_DEV void cameraRay(const KrrCameraDev &c, int px, int py, int W, int H, const float cs[5], V3 &o, V3 &d, float &time) {
	float pxf = xadd(xadd((float) px, 0.5f), cs[0]), pyf = xadd(xadd((float) py, 0.5f), cs[1]);
	float ndcx = xadd(xdiv(xmul(2.f, pxf), (float) W), -1.f), ndcy = xadd(xdiv(xmul(2.f, pyf), (float) H), -1.f);
	time = xadd(c.shutterOpen, xmul(c.shutterTime, cs[4]));
	V3 fd = mk3(xmul(xmul(c.tanFov, c.aspectRatio), ndcx), xmul(c.tanFov, ndcy), -1.f);
	// Eigen normalized(): v / sqrt(squaredNorm); squaredNorm of a 3-vector reduces as x*x + (y*y + z*z)
	float n2 = xadd(xmul(fd.x, fd.x), xadd(xmul(fd.y, fd.y), xmul(fd.z, fd.z)));
	float nn = xsqrt(n2);
	fd = mk3(xdiv(fd.x, nn), xdiv(fd.y, nn), xdiv(fd.z, nn));
	V3 lo = mk3(0, 0, 0), ld = fd;
	if (c.lensRadius > 1e-5f) {
		float ax, ay;
		uniformSampleDisk(cs[2], cs[3], ax, ay);
		lo = mk3(c.lensRadius * ax, c.lensRadius * ay, 0.f);
		ld = normalize(fd * c.focalDistance - lo);
	}
	// Transformation::operator()(ray): m * origin, linear(m) * dir (raytracing.h:95-97)
	const float *m = c.transform.m;
	auto row = [&](int r, V3 v) { return xadd(xmul(m[r * 4], v.x), xadd(xmul(m[r * 4 + 1], v.y), xmul(m[r * 4 + 2], v.z))); };
	o = mk3(xadd(row(0, lo), m[3]), xadd(row(1, lo), m[7]), xadd(row(2, lo), m[11]));
	d = mk3(row(0, ld), row(1, ld), row(2, ld));
}

__global__ void k_generate_camera_rays(const __grid_constant__ Wavefront wf) {
	KRR_PDL_ENTRY();
	RayQueue q = wf.rays[0];
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < wf.p.pixelCount; i += gridDim.x * blockDim.x) {
		int pixelId = framePixel(wf.p, i);
		Pcg rng{wf.px.rng[i], rngIncOf(wf.p, i)};
		float cs[5];
#pragma unroll
		for (int k = 0; k < 5; k++) cs[k] = rng.get1D();
		wf.px.rng[i] = rng.state;
		if (wf.px.cameraSample)
			for (int k = 0; k < 5; k++) wf.px.cameraSample[5 * (size_t) i + k] = cs[k];
		V3 o, d;
		float time;
		cameraRay(wf.cam, pixelId % wf.p.width, pixelId / wf.p.width, wf.p.width, wf.p.height, cs, o, d, time);
		// pushCameraRay, workqueue.h:180-189 (slot = pixel: every pixel spawns exactly one ray, so the
		// queue is dense and the writes are coalesced without any atomic)
		stcs4(q.o_time + i, make_float4(o.x, o.y, o.z, time));
		stcs4(q.d_medium + i, make_float4(d.x, d.y, d.z, __int_as_float(wf.cam.medium)));
		if (!wf.p.implicitDepth0) {
			stcs4(q.thp + i, sp(1));
			stcs4(q.pu + i, sp(1));
			stcs4(q.pl + i, sp(1));
			stcs4(q.ctxP_pix + i, make_float4(0, 0, 0, __int_as_float(i)));
			stcs4(q.ctxN_dep + i, make_float4(0, 0, 0, __int_as_float(0)));
		}
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) wf.counters[0].nRay = wf.p.pixelCount;
}

// alphaKilled, shading.h:89-113 + HashFloat (util/hash.h:117-133)
KRR_DEV uint64_t murmur64A(const unsigned char *key, int len, uint64_t seed) {
	const uint64_t m = 0xc6a4a7935bd1e995ull;
	const int r = 47;
	uint64_t h = seed ^ (len * m);
	for (int i = 0; i < len / 8; i++) {
		uint64_t k;
		memcpy(&k, key + 8 * i, 8);
		k *= m; k ^= k >> r; k *= m; h ^= k; h *= m;
	}
	h ^= h >> r; h *= m; h ^= h >> r; // len is a multiple of 8 here (24 bytes)
	return h;
}
__device__ __noinline__ bool alphaKilled(const Wavefront &wf, int inst, int prim, float u, float v, V3 o, V3 d) {
	const MeshRec &mesh = wf.scene.meshes[wf.scene.instances[inst].mesh];
	if (mesh.material < 0) return false;
	const TexRec &t = wf.scene.materials[mesh.material].tex[4];
	if (!t.valid) return false;
	float uvx = 0, uvy = 0;
	if (mesh.uvOff >= 0) {
		const int32_t *idx = wf.scene.indices + 3 * ((size_t) mesh.idxOff + prim);
		float b0 = 1 - u - v;
		const float *uv = wf.scene.texcoords + 2 * (size_t) mesh.uvOff;
		uvx = b0 * uv[2 * idx[0]] + u * uv[2 * idx[1]] + v * uv[2 * idx[2]];
		uvy = b0 * uv[2 * idx[0] + 1] + u * uv[2 * idx[1] + 1] + v * uv[2 * idx[2] + 1];
	}
	float4 op	= sampleTex(t, wf.scene.texels, uvx, uvy, make_float4(1, 1, 1, 1));
	float alpha = 1 - luminanceRGB(op.x, op.y, op.z);
	if (alpha >= 1) return false;
	if (alpha <= 0) return true;
	float buf[6] = {o.x, o.y, o.z, d.x, d.y, d.z};
	float h = (uint32_t) murmur64A((const unsigned char *) buf, 24, 0) * 0x1p-32f;
	return h > alpha;
}

// =================================================================================================
// Closest stage: traverse, record the hit, route the slot index (device.cu:43-81).
// Persistent warps: a warp claims rays from the queue with one atomicAdd, steps all its lanes
// through the phase-aligned traversal (bvh.cuh), finalises the lanes whose ray terminated and
// refills them as soon as kRefill lanes are idle -- SIMT lanes do not wait for the slowest ray.
// lanes that must be idle / finished before a warp refills / finalises (Params::refill): 8 for tree
// traversals; 16 when the whole scene is one flat triangle list, where every ray takes the same single trip
// (measured: +2 % on the Cornell box, -4.5 % on the instanced scene with 16)
constexpr int kRefill = 8, kRefillFlat = 16;
#ifndef KRR_CLAIM
#define KRR_CLAIM 32 // measured 8 / 16 / 32 / 64: 32 is best on all three scene kinds (finer balance at the end of a queue)
#endif
constexpr int kClaim  = KRR_CLAIM; // rays a warp claims from the queue per atomicAdd once its static share is done

// Work distribution of the persistent trace kernels.  Warp w starts with the static slice
// [32 w, 32 w + 32) of the queue -- no atomic, and warps beyond a short queue exit at once -- and
// afterwards claims kClaim items at a time from the shared cursor (which counts from 32 * #warps).
struct WarpWork {
	int next, end;	// private range still to hand out
	int n, first;	// queue size, first dynamically claimed index
	int32_t *cursor;
	bool exhausted; // warp-uniform
	int claim;
	// per: items a warp starts with (one per lane); claim_: items per later atomic claim
	// warp / nWarps: this warp's index among the warps that work on the queue
	KRR_DEV void init(int n_, int32_t *cursor_, int warp, int nWarps, int per = 32, int claim_ = kClaim) {
		n = n_, cursor = cursor_, first = nWarps * per, claim = claim_;
		next = warp * per, end = min(next + per, n_);
		exhausted = false;
	}
	// Same, for the tree traversals: a SHORT queue is spread evenly over all warps (ceil(n / nWarps) rays each,
	// CTAs are dealt round-robin to the SMs) instead of 32 rays to each of the first few warps.  A launch lasts as
	// long as its slowest warp, and a warp as long as the slowest of its rays at the pace of its divergence: with
	// few rays per warp the deep-bounce launches of a large scene come down from the latency of "the slowest of
	// 32 divergent rays" towards that of one ray (KRR_SPREAD=0 restores the packed slices for A/B runs).
	KRR_DEV void initSpread(int n_, int32_t *cursor_, int block, int nBlocks, int warpInBlock, int warpsPerBlock) {
		const int nWarps = nBlocks * warpsPerBlock;
#ifndef KRR_SPREAD
#define KRR_SPREAD 1
#endif
#ifndef KRR_SPREAD_MIN
#define KRR_SPREAD_MIN 8
#endif
		const int per = KRR_SPREAD ? min(32, max(KRR_SPREAD_MIN, (n_ + nWarps - 1) / nWarps)) : 32;
		init(n_, cursor_, block * warpsPerBlock + warpInBlock, nWarps, per); // block-major: neighbouring rays stay on one SM (L1)
	}
	// hands one item to every lane of `idle` (lane order); -1 when the queue is exhausted
	KRR_DEV int take(unsigned idle, int lane) {
		const unsigned FULL = 0xffffffffu;
		const int want = __popc(idle);
		if (next >= end && !exhausted) { // private range used up: claim the next batch
			int base = n;
			if (first < n) { // (short queues are covered by the static slices alone: no atomic at all)
				if (lane == 0) base = first + atomicAdd(cursor, claim);
				base = __shfl_sync(FULL, base, 0);
			}
			next = base, end = min(base + claim, n);
			if (base >= n) exhausted = true, next = end = 0;
		}
		const int rank = __popc(idle & ((1u << lane) - 1));
		int r = -1;
		if ((idle >> lane) & 1) { r = next + rank; if (r >= end) r = -1; }
		next = min(next + want, end);
		return r;
	}
};

// object<->world of an instance at ray time `time`: the transform list OptiX reports for a hit
// (getInstanceTransform, shading.h:70-76) -- static instances use the uploaded matrices, moving ones
// evaluate their SRT chain (motion.cuh)
static __device__ __noinline__ void movingXf(const SceneDev &sc, int inst, int node, float time, Xf *xf, Xf *inv) {
	movingInstanceXf(sc.xnodes, sc.motionKeys, sc.motionFlat, inst, node, time, *xf, *inv);
}

// null-material hit: the ray continues behind the surface (device.cu:54-58): new origin / medium in o4 / d4
template <bool MOTION = true>
KRR_DEV void continueThroughNull(const Wavefront &wf, Hit h, float4 &o4, float4 &d4) {
	const InstRec &in	= wf.scene.instances[h.inst];
	const MeshRec &mesh = wf.scene.meshes[in.mesh];
	const int32_t *idx	= wf.scene.indices + 3 * ((size_t) mesh.idxOff + h.prim);
	const float *P		= wf.scene.positions + 3 * (size_t) mesh.posOff;
	V3 p0 = ld3(P + 3 * idx[0]), p1 = ld3(P + 3 * idx[1]), p2 = ld3(P + 3 * idx[2]);
	float b0 = 1 - h.u - h.v;
	const Xf *xf = &in.xf, *inv = &in.inv;
	Xf mxf, minv;
	if (MOTION && in.motion >= 0) {
		movingXf(wf.scene, h.inst, in.motion, o4.w, &mxf, &minv);
		xf = &mxf, inv = &minv;
	}
	V3 p = xfPoint(*xf, b0 * p0 + h.u * p1 + h.v * p2);
	V3 n;
	if (mesh.nrmOff >= 0) {
		const float *N = wf.scene.normals + 3 * (size_t) mesh.nrmOff;
		n = normalize(b0 * ld3(N + 3 * idx[0]) + h.u * ld3(N + 3 * idx[1]) + h.v * ld3(N + 3 * idx[2]));
	} else n = normalize(cross(p1 - p0, p2 - p0));
	n = normalize(xfNormal(*inv, n));
	V3 d = mk3(d4);
	V3 off = n * kRayEps;
	if (dot(n, d) < 0.f) off = -off;
	V3 no = p + off;
	o4 = make_float4(no.x, no.y, no.z, o4.w);
	if (wf.p.enableMedium && mesh.mediumIn != mesh.mediumOut) // Interaction::getMedium(dir), raytracing.h:162-166
		d4.w = __int_as_float(dot(d, n) > 0 ? mesh.mediumOut : mesh.mediumIn);
}
// ... re-queued at the same item depth
template <bool MOTION = true>
__device__ __noinline__ void requeueThroughNull(const Wavefront &wf, const RayQueue &q, const RayQueue &nq, int i, int s, Hit h, float4 o4, float4 d4,
												bool implicit = false) {
	continueThroughNull<MOTION>(wf, h, o4, d4);
	stcs4(nq.o_time + s, o4);
	stcs4(nq.d_medium + s, d4);
	stcs4(nq.thp + s, implicit ? sp(1) : ldcs4(q.thp + i));
	stcs4(nq.pu + s, implicit ? sp(1) : ldcs4(q.pu + i));
	stcs4(nq.pl + s, implicit ? sp(1) : ldcs4(q.pl + i));
	stcs4(nq.ctxP_pix + s, implicit ? make_float4(0, 0, 0, __int_as_float(i)) : ldcs4(q.ctxP_pix + i));
	stcs4(nq.ctxN_dep + s, implicit ? make_float4(0, 0, 0, __int_as_float(0)) : ldcs4(q.ctxN_dep + i));
}

// block / nBlocks: this CTA's index among the CTAs that run the closest stage (the fused trace kernel
// splits its grid between the shadow rays of one depth and the closest rays of the next)
// MODE (trace kernels): kTraceStatic = static scene, kTraceMotion = instances with motion transforms,
// kTraceFlat = static scene that is ONE flat triangle list (the leaf trips are branch-free triangle pairs,
// bvh.cuh Traverser PAIR; a separate instantiation because the pair code costs the tree scenes registers:
// -10 % on the 10 000-instance scene when it lived in the same kernel)
constexpr int kTraceStatic = 0, kTraceMotion = 1, kTraceFlat = 2;
template <int MODE>
KRR_DEV void traceClosestBody(const Wavefront &wf, int depth, TraceSmem &sm, int block, int nBlocks) {
	constexpr bool MOTION = MODE == kTraceMotion;
	const RayQueue q	= wf.rays[depth & 1];
	const RayQueue nq	= wf.rays[(depth & 1) ^ 1];
	DepthCounters *dc	= wf.counters + depth;
	const int n			= dc->nRay;
	const unsigned FULL = 0xffffffffu;
	const int lane		= threadIdx.x & 31;
	const bool implicit = wf.p.implicitDepth0 && depth == 0; // depth-0 items: only origin / direction are stored
	Traverser<false, MOTION, MODE == kTraceFlat> tr;
	LocalStack<false> ls;
	int ray	  = -1;	   // queue slot this lane holds, -1 = idle
	int pix	  = 0;	   // its pixel (fetched with the ray: the finalisation needs it first)
	bool done = false; // traversal finished, result in tr.best, not yet finalised
	WarpWork work;
	if (MODE == kTraceFlat) work.init(n, &dc->cursorRay, (block * kTraceBlock + (int) threadIdx.x) >> 5, (nBlocks * kTraceBlock) >> 5);
	else work.initSpread(n, &dc->cursorRay, block, nBlocks, (int) threadIdx.x >> 5, kTraceBlock >> 5);
	if (work.next >= n) return; // short queue: this warp has no static share and nothing to claim
	const int32_t *__restrict__ perm = depth > 0 ? wf.permClosest : nullptr; // reordered queue (k_sort_rays)
	int medium = -1; // medium the ray travels in (d_medium.w)
	while (true) {
		// ---- finalise finished rays in BATCHES: the finalisation is a chain of dependent long-latency
		// operations (pixel RNG read-modify-write, queue-counter atomics), so it runs once kRefill lanes
		// have finished (they would idle until the refill anyway), or when nothing is left to trace ----
		const unsigned doneMask = __ballot_sync(FULL, done);
		if (doneMask && (__popc(doneMask) >= wf.p.refill || !__ballot_sync(FULL, ray >= 0 && !done))) {
			// queue id of the lane: 0 miss, 1 + mt scatter, MAT_COUNT + 1 null pass-through, MAT_COUNT + 2 medium
			// sample, -1 none (lane not finished, or path ended by Russian roulette)
			int qid		= -1;
			bool light	= false;
			const Hit h = tr.best;
			const int i = ray;
			if (done) {
				if (tr.overflow) atomicExch(&wf.errorFlags[0], 1);
#ifdef KRR_COUNT_TRIPS
				atomicAdd(&wf.errorFlags[1], tr.nodeSteps), atomicMax(&wf.errorFlags[2], tr.nodeSteps), atomicAdd(&wf.errorFlags[3], tr.triTests);
				atomicAdd((unsigned long long *) &wf.tripHist[min(tr.nodeSteps >> 3, 63) * 2], 1ull);
				atomicAdd(&wf.tripHist[128], tr.enters), atomicAdd(&wf.tripHist[129], tr.culled);
#endif
				const int4 rec = make_int4(h.inst, h.prim, __float_as_int(h.u), __float_as_int(h.v));
				wf.hits[i]	   = rec;
				if (depth == 0 && wf.firstHits) wf.firstHits[pix] = rec;
				if (wf.p.enableMedium && medium >= 0) {
					// ray inside a medium: hit or miss, the item goes to the medium stage (device.cu:50-53, 69-72)
					qid		   = MAT_COUNT + 2;
					wf.hitT[i] = h.inst < 0 ? kInf : h.t;
				} else if (h.inst < 0) qid = 0;
				else {
					const uint32_t f = wf.instFlags[h.inst];
					if (f & 1) qid = MAT_COUNT + 1;
					else {
						qid	  = 1 + (int) (f >> 4);
						light = (f & 4) != 0;
						if (wf.p.rrInTrace && depth < wf.p.maxDepth) { // no scatter stage (hence no draw) at the last depth
							Pcg rng{wf.px.rng[pix], rngIncOf(wf.p, pix)};
							const bool alive = rng.get1D() < wf.p.probRR;
							wf.px.rng[pix]	 = rng.state;
							if (!alive) qid = -2;
						}
					}
				}
				ray = -1, done = false;
			}
			// one atomicAdd per QUEUE per warp, all queues in the same instruction: lanes are grouped by
			// queue id, the first lane of each group reserves the group's slots
			const unsigned grp = __match_any_sync(FULL, qid);
			const int leader   = __ffs(grp) - 1;
			int32_t *counter   = qid == 0 ? &dc->nMiss : qid <= MAT_COUNT ? &dc->nScatter[max(qid - 1, 0)] : qid == MAT_COUNT + 1 ? &dc[1].nRay : &dc->nMediumSample;
			int base = 0;
			if (qid >= 0 && lane == leader) base = atomicAdd(counter, __popc(grp));
			const unsigned lightMask = __ballot_sync(FULL, light);
			int lbase = 0;
			if (lightMask && lane == __ffs(lightMask) - 1) lbase = atomicAdd(&dc->nHitLight, __popc(lightMask));
			const unsigned killed = __ballot_sync(FULL, qid == -2);
			if (killed && lane == 0) atomicAdd(&dc->nScatterKilled, __popc(killed));
			base = __shfl_sync(FULL, base, leader);
			if (lightMask) lbase = __shfl_sync(FULL, lbase, __ffs(lightMask) - 1);
			const unsigned below = (1u << lane) - 1;
			if (light) wf.hitLightIdx[lbase + __popc(lightMask & below)] = i;
			if (qid >= 0) {
				const int s = base + __popc(grp & below);
				if (qid == 0) wf.missIdx[s] = i;
				else if (qid <= MAT_COUNT) wf.scatterIdx[qid - 1][s] = i;
				else if (qid == MAT_COUNT + 1)
					requeueThroughNull<MOTION>(wf, q, nq, i, s, h, make_float4(tr.o.x, tr.o.y, tr.o.z, tr.time), make_float4(tr.d.x, tr.d.y, tr.d.z, __int_as_float(medium)),
											   implicit);
				else wf.mediumSampleIdx[s] = i;
			}
		}
		unsigned idle = __ballot_sync(FULL, ray < 0);
		if (!work.exhausted && __popc(idle) >= wf.p.refill) {
			int r = work.take(idle, lane);
			if (r >= 0) {
				if (perm) r = __ldg(perm + r);
				ray = r;
				const float4 o4 = ldcs4(q.o_time + r), d4 = ldcs4(q.d_medium + r);
				pix	   = implicit ? r : __float_as_int(__ldcs(reinterpret_cast<const float *>(q.ctxP_pix + r) + 3));
				medium = __float_as_int(d4.w);
				tr.begin(wf.bvh, mk3(o4), mk3(d4), kInf, o4.w);
			}
			idle = __ballot_sync(FULL, ray < 0);
		}
		if (idle == FULL) {
			if (work.exhausted) break;
			continue; // private range ran dry mid-refill: claim again
		}
		#ifndef KRR_CLOSEST_VOTE
#define KRR_CLOSEST_VOTE true
#endif
		const bool fin = tr.trip<KRR_CLOSEST_VOTE>(ray >= 0 && !done, wf.bvh, wf.scene.instances, sm, ls, [&](int inst, int prim, float u, float v) {
			if (!(wf.instFlags[inst] & 2)) return true;
			return !alphaKilled(wf, inst, prim, u, v, tr.o, tr.d);
		});
		done |= fin;
	}
}

// Resident CTAs per SM the stand-alone closest kernel (depth 0) is compiled for.  The moving-instance kernel runs 7 CTAs
// at 72 registers like the fused kernel (4K primary rays of the 10 000-instance scene: 23.5 -> 21.0 ms); the static
// kernels get no occupancy request: 128 registers, 4 CTAs (a request of 1 lets the compiler take 168+ registers: -25 %)
#ifndef KRR_CLOSEST_MINB
#define KRR_CLOSEST_MINB 7
#endif
template <int MODE> struct ClosestBounds { static constexpr int minBlocks = MODE == kTraceMotion ? KRR_CLOSEST_MINB : 4; };
#define KRR_CLOSEST_BOUNDS __launch_bounds__(kTraceBlock, ClosestBounds<MODE>::minBlocks)
template <int MODE>
__global__ void KRR_CLOSEST_BOUNDS k_trace_closest(const __grid_constant__ Wavefront wf, int depth) {
	KRR_PDL_ENTRY();
	__shared__ TraceSmem sm;
	traceClosestBody<MODE>(wf, depth, sm, blockIdx.x, gridDim.x);
}

// =================================================================================================
// shared by the hit-light and scatter stages: rebuild the interaction geometry of a hit
struct SurfaceGeom {
	V3 p, n, tangent, bitangent, wo;
	float uvx, uvy;
	int inst, prim, mesh, material, light;
};

template <bool MOTION = true>
KRR_DEV void rebuildGeometry(const Wavefront &wf, int4 hit, V3 rayDir, float time, SurfaceGeom &g) {
	// getHitInfo (shading.h:78-87) + prepareSurfaceInteraction geometry part (shading.h:121-170)
	const SceneDev &sc	= wf.scene;
	g.inst = hit.x, g.prim = hit.y;
	float u = __int_as_float(hit.z), v = __int_as_float(hit.w);
	float b0 = 1 - u - v;
	const InstRec &in	= sc.instances[g.inst];
	const MeshRec &mesh = sc.meshes[in.mesh];
	g.mesh = in.mesh, g.material = mesh.material;
	const int32_t *idx = sc.indices + 3 * ((size_t) mesh.idxOff + g.prim);
	int i0 = __ldg(idx), i1 = __ldg(idx + 1), i2 = __ldg(idx + 2);
	const float *P = sc.positions + 3 * (size_t) mesh.posOff;
	V3 p0 = ld3(P + 3 * i0), p1 = ld3(P + 3 * i1), p2 = ld3(P + 3 * i2);
	g.wo = normalize(-normalize(rayDir));
	g.p	 = b0 * p0 + u * p1 + v * p2;
	if (mesh.nrmOff >= 0) {
		const float *N = sc.normals + 3 * (size_t) mesh.nrmOff;
		g.n = normalize(b0 * ld3(N + 3 * i0) + u * ld3(N + 3 * i1) + v * ld3(N + 3 * i2));
	} else g.n = normalize(cross(p1 - p0, p2 - p0));
	if (mesh.tanOff >= 0) {
		const float *T = sc.tangents + 3 * (size_t) mesh.tanOff;
		g.tangent = normalize(b0 * ld3(T + 3 * i0) + u * ld3(T + 3 * i1) + v * ld3(T + 3 * i2));
		g.tangent = normalize(g.tangent - g.n * dot(g.n, g.tangent));
	} else { // getPerpendicular, util/math_utils.h:122-133
		V3 a = mk3(fabsf(g.n.x), fabsf(g.n.y), fabsf(g.n.z));
		uint32_t uyx = (a.x - a.y) < 0 ? 1 : 0, uzx = (a.x - a.z) < 0 ? 1 : 0, uzy = (a.y - a.z) < 0 ? 1 : 0;
		uint32_t xm = uyx & uzx, ym = (1 ^ xm) & uzy, zm = 1 ^ (xm | ym);
		g.tangent = normalize(cross(g.n, mk3((float) xm, (float) ym, (float) zm)));
	}
	g.bitangent = normalize(cross(g.n, g.tangent));
	g.uvx = g.uvy = 0;
	if (mesh.uvOff >= 0) {
		const float *UV = sc.texcoords + 2 * (size_t) mesh.uvOff;
		g.uvx = b0 * UV[2 * i0] + u * UV[2 * i1] + v * UV[2 * i2];
		g.uvy = b0 * UV[2 * i0 + 1] + u * UV[2 * i1 + 1] + v * UV[2 * i2 + 1];
	}
	g.light = in.lightBase >= 0 ? in.lightBase + g.prim : -1;
	const Xf *xf = &in.xf, *inv = &in.inv;
	Xf mxf, minv;
	if (MOTION && in.motion >= 0) {
		movingXf(sc, g.inst, in.motion, time, &mxf, &minv);
		xf = &mxf, inv = &minv;
	}
	g.p			= xfPoint(*xf, g.p);
	g.n			= normalize(xfNormal(*inv, g.n));
	g.tangent	= normalize(xfNormal(*inv, g.tangent));
	g.bitangent = normalize(xfNormal(*inv, g.bitangent));
}

// material part of prepareSurfaceInteraction, shading.h:172-225
KRR_DEV void evalMaterial(const Wavefront &wf, SurfaceGeom &g, const Wavelengths &wl, ShadingData &sd, bool &terminateSecondary) {
	const SceneDev &sc = wf.scene;
	const MatRec &mat  = sc.materials[g.material];
	sd.bsdfType = mat.bsdfType;
	sd.specularTransmission = mat.specularTransmission;
	sd.IoR = mat.ior;
	sd.hasEta = sd.hasK = 0;
	terminateSecondary = false;
	if (mat.eta.kind != 0) {
		auto evalSpec = [&](const SpectrumRec &s, float lambda) -> float {
			switch (s.kind) {
				case 1: return s.a[0];
				case 2: return s.a[0] + s.b[0] / pow2(lambda / 1000.f);						 // CauchyIoRSpectrum, spectrum.h:113
				case 3: {																		 // SellmeierIoRSpectrum, spectrum.h:130-132
					float l2 = pow2(lambda), sum = 0;
					for (int k = 0; k < 3; k++) { float den = l2 - s.b[k]; sum += den == 0 ? 0.f : (l2 * s.a[k]) / den; }
					return sqrtf(1 + sum);
				}
				default: { // piecewise linear table, spectrum.h:251-259
					const float *L = sc.spectrumTables + s.tabOff, *V = L + s.n;
					if (s.n == 0 || lambda < L[0] || lambda > L[s.n - 1]) return 0.f;
					int o = 0;
					while (o + 2 < s.n && L[o + 1] <= lambda) o++;
					float t = (lambda - L[o]) / (L[o + 1] - L[o]);
					return lerpf(V[o], V[o + 1], t);
				}
			}
		};
		sd.IoR = evalSpec(mat.eta, wl.lambda[0]);
		if (mat.eta.kind != 1) terminateSecondary = true;
		sd.hasEta  = 1;
		sd.etaSpec = make_float4(evalSpec(mat.eta, wl.lambda[0]), evalSpec(mat.eta, wl.lambda[1]), evalSpec(mat.eta, wl.lambda[2]), evalSpec(mat.eta, wl.lambda[3]));
	}
	float diffuse[3], specular[3], spec3;
	const MeshRec &mesh = sc.meshes[g.mesh];
	if (mat.constColours) {
		for (int k = 0; k < 3; k++) diffuse[k] = mat.constDiffuse[k], specular[k] = mat.constSpecular[k];
		spec3 = mat.constSpecular[3];
	} else {
		float4 diff = sampleTex(mat.tex[0], sc.texels, g.uvx, g.uvy, make_float4(mat.diffuse[0], mat.diffuse[1], mat.diffuse[2], mat.diffuse[3]));
		float4 spec = sampleTex(mat.tex[1], sc.texels, g.uvx, g.uvy, make_float4(mat.specular[0], mat.specular[1], mat.specular[2], mat.specular[3]));
		diffuse[0] = diff.x, diffuse[1] = diff.y, diffuse[2] = diff.z;
		specular[0] = spec.x, specular[1] = spec.y, specular[2] = spec.z, spec3 = spec.w;
	}
	if (mat.tex[3].valid && mesh.uvOff >= 0) { // normal map
		float4 nv = sampleTex(mat.tex[3], sc.texels, g.uvx, g.uvy, make_float4(0, 0, 1, 0));
		V3 nm = mk3(2 * nv.x - 1, 2 * nv.y - 1, 2 * nv.z - 1);
		g.n	  = normalize(g.tangent * nm.x + g.bitangent * nm.y + g.n * nm.z);
		g.tangent	= normalize(g.tangent - g.n * dot(g.tangent, g.n));
		g.bitangent = normalize(cross(g.n, g.tangent));
	}
	float dr[3], srgb[3];
	if (mat.shadingModel == 0) { // MetallicRoughness: [SPECULAR] G roughness, B metallic
		for (int k = 0; k < 3; k++) {
			dr[k]	= diffuse[k] * (1 - specular[2]) + 0.f * specular[2];
			srgb[k] = 0.f * (1 - specular[2]) + diffuse[k] * specular[2];
		}
		sd.metallic	 = specular[2];
		sd.roughness = specular[1];
	} else { // SpecularGlossiness
		for (int k = 0; k < 3; k++) dr[k] = diffuse[k], srgb[k] = specular[k];
		sd.roughness = 1.f - spec3;
		// getMetallic, shading.h:17-30
		float d = luminanceRGB(dr[0], dr[1], dr[2]), s = luminanceRGB(srgb[0], srgb[1], srgb[2]);
		if (s == 0) sd.metallic = 0;
		else {
			float b = s + d - 0.08f, c = 0.04f - s;
			float root = sqrtf(b * b - 0.16f * c);
			sd.metallic = fmaxf(0.f, (root - b) * 12.5f);
		}
	}
	if (mat.constColours) {
		sd.diffuse	= sampleBounded(mat.diffuseSpec, wl);
		sd.specular = sampleBounded(mat.specularSpec, wl);
	} else {
		sd.diffuse	= sampleBounded(makeBounded(sc.cs.zNodes, sc.cs.coeffs, dr[0], dr[1], dr[2]), wl);
		sd.specular = sampleBounded(makeBounded(sc.cs.zNodes, sc.cs.coeffs, srgb[0], srgb[1], srgb[2]), wl);
	}
	sd.anisotropic = mat.anisotropic;
}

// =================================================================================================
// handleHit + handleMiss (integrator.cpp:78-108)
// Out of line: the scatter stage of the same depth runs it as a prologue (one launch less per depth;
// the queues of this stage are short: a few light hits, and misses only matter with an environment light)
// One hit-light item (handleHit, integrator.cpp:78-90): emitted radiance x spectral-MIS weight.  cp = (ctx.p, pixel),
// packed = depth | bsdfType << 8 of the ray item.  Shared by the stage kernels and the tail kernel.
template <bool MOTION>
KRR_DEV Spec hitLightItem(const Wavefront &wf, int4 hit, float4 d4, float time, float4 cp, int packed, Spec thp, Spec pu, Spec pl, float lightSelPdf) {
	SurfaceGeom g;
	rebuildGeometry<MOTION>(wf, hit, mk3(d4), time, g);
	const int pix = __float_as_int(cp.w);
	const int itemDepth = packed & 0xff, bsdfType = packed >> 8;
	// normal map changes intr.n before the light is evaluated (the CH program prepares the full
	// interaction before pushing): apply it when present
	const MatRec &mat = wf.scene.materials[g.material];
	if (mat.tex[3].valid && wf.scene.meshes[g.mesh].uvOff >= 0) {
		float4 nv = sampleTex(mat.tex[3], wf.scene.texels, g.uvx, g.uvy, make_float4(0, 0, 1, 0));
		g.n = normalize(g.tangent * (2 * nv.x - 1) + g.bitangent * (2 * nv.y - 1) + g.n * (2 * nv.z - 1));
	}
	Wavelengths wl = expandWavelengths(wf.px.lambda[pix]);
	const LightRec lr	  = wf.scene.lights[g.light];
	const TriLightRec &tl = wf.scene.triLights[lr.index];
	Spec Le = areaLightL(tl, g.n, g.wo, wl, wf.scene.cs) * thp;
	if (wf.p.nee && itemDepth && !(bsdfType & BSDF_DELTA)) {
		float lightPdf = areaLightPdfLi(tl, wf.scene.instances[tl.inst], g.p, g.n, mk3(cp)) * lightSelPdf;
		Le = Le / mean(pl * lightPdf + pu);
	} else Le = Le / mean(pu);
	return Le;
}
// One miss item (handleMiss, integrator.cpp:92-108); returns thp * sum of the infinite lights' weighted radiance
KRR_DEV Spec missItem(const Wavefront &wf, float4 d4, int pix, int packed, Spec thp, Spec pu, Spec pl, float lightSelPdf) {
	const int itemDepth = packed & 0xff, bsdfType = packed >> 8;
	Wavelengths wl = expandWavelengths(wf.px.lambda[pix]);
	Spec L = sp(0);
	for (int li = 0; li < wf.scene.nInfinite; li++) {
		const AnalyticLightRec &light = wf.scene.analytic[wf.scene.infiniteLights[li]];
		Spec Li = infiniteLi(light, mk3(d4), wl, wf.scene);
		if (wf.p.nee && itemDepth && !(bsdfType & BSDF_DELTA)) {
			float lightPdf = kInv4Pi * lightSelPdf;
			L += Li / mean(pu + pl * lightPdf);
		} else L += Li / mean(pu);
	}
	return thp * L;
}

template <bool MOTION>
__device__ __noinline__ void handleHitMissBody(const Wavefront &wf, int depth) {
	const RayQueue q  = wf.rays[depth & 1];
	DepthCounters *dc = wf.counters + depth;
	const int stride  = gridDim.x * blockDim.x;
	const float lightSelPdf = wf.scene.nLights > 0 ? 1.f / wf.scene.nLights : 0.f;
	const bool implicit = wf.p.implicitDepth0 && depth == 0;
	const int nHit = dc->nHitLight;
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nHit; k += stride) {
		int i	  = wf.hitLightIdx[k];
		int4 hit  = wf.hits[i];
		float4 d4 = ldg4(q.d_medium + i);
		float4 cp = implicit ? make_float4(0, 0, 0, __int_as_float(i)) : ldg4(q.ctxP_pix + i), cn = implicit ? make_float4(0, 0, 0, 0) : ldg4(q.ctxN_dep + i);
		const int pix = __float_as_int(cp.w), packed = __float_as_int(cn.w);
		Spec thp = implicit ? sp(1) : ldg4(q.thp + i), pu = implicit ? sp(1) : ldg4(q.pu + i);
		Spec pl	 = (wf.p.nee && (packed & 0xff) && !((packed >> 8) & BSDF_DELTA)) ? ldg4(q.pl + i) : sp(1); // (never a depth-0 item)
		Spec Le	 = hitLightItem<MOTION>(wf, hit, d4, ldg4(q.o_time + i).w, cp, packed, thp, pu, pl, lightSelPdf);
		wf.px.L[pix] = Le + wf.px.L[pix]; // addRadiance (<= 1 item per pixel per stage: plain RMW)
	}
	const int nMiss = dc->nMiss;
	if (wf.scene.nInfinite == 0) return;
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nMiss; k += stride) {
		int i	  = wf.missIdx[k];
		float4 d4 = ldg4(q.d_medium + i);
		float4 cp = implicit ? make_float4(0, 0, 0, __int_as_float(i)) : ldg4(q.ctxP_pix + i), cn = implicit ? make_float4(0, 0, 0, 0) : ldg4(q.ctxN_dep + i);
		int pix = __float_as_int(cp.w), packed = __float_as_int(cn.w);
		Spec thp = implicit ? sp(1) : ldg4(q.thp + i), pu = implicit ? sp(1) : ldg4(q.pu + i), pl = implicit ? sp(1) : ldg4(q.pl + i);
		wf.px.L[pix] = missItem(wf, d4, pix, packed, thp, pu, pl, lightSelPdf) + wf.px.L[pix];
	}
}

template <bool MOTION>
__global__ void __launch_bounds__(128) k_handle_hit_miss(const __grid_constant__ Wavefront wf, int depth) {
	KRR_PDL_ENTRY();
	handleHitMissBody<MOTION>(wf, depth);
}

// =================================================================================================
// generateScatterRays (integrator.cpp:110-164), one launch per material type
// The Disney instantiation is ~11k SASS instructions (180 KB) against a 32 KB instruction cache, and
// warps at unrelated program counters evict each other's lines: ncu attributed 54 % of the stall samples
// to instruction fetch (stall_no_inst).  Two counter-measures: (1) the warps of a CTA are re-aligned with
// block barriers at phase boundaries (KRR_SCATTER_LOCKSTEP), so that a line fetched for one warp serves
// the others: -17 % stage time at every block size tried; one 640-thread CTA per SM loses that again to
// barrier waits, 4 x 128 and 2 x 256 measured the same; (2) the Disney path evaluates f() and pdf() for
// the light direction and for the cosine sample through one copy of the code (Bsdf::eval).
#ifndef KRR_SCATTER_BLOCK
#define KRR_SCATTER_BLOCK 128
#endif
// CTAs per SM: the stage is bound by dependent-issue latency, so warps in flight are worth more than the
// spills they cost.  Disney instantiation, bench workload (2 bands): 3 CTAs (165 registers) 3921, 4 (128, no
// spills) 4120-4141, 5 (96) 4126, 6 (80, 548 B of spill stores) 4188, 7 (72) 4172, 8 (64) 4094 Mrays/s
#ifndef KRR_SCATTER_MINB
#define KRR_SCATTER_MINB 6
#endif
#ifndef KRR_SCATTER_LOCKSTEP
#define KRR_SCATTER_LOCKSTEP 1
#endif
constexpr int kScatterBlock = KRR_SCATTER_BLOCK;

// One scatter item = one surface vertex of a path that survived Russian roulette (generateScatterRays,
// integrator.cpp:120-163; prepareSurfaceInteraction, shading.h:115-226).  Shared, operation for operation, by the
// stage kernel k_scatter and the tail kernel, so that both produce the same bits.
struct ScatterIn {
	int4 hit;		 // inst, prim, u, v
	float4 d4;		 // ray direction, medium
	float time;
	int pix, medium;
	Spec thp, pu;	 // thp BEFORE the division by probRR
};
struct ScatterOut {
	bool pushShadow, pushNext;
	V3 so, sdv;				 // shadow ray (tMax = 1)
	Spec sContrib, sPu, sPl; // Ld / (pl + pu).mean()  (Ld itself with media)
	int sMedium;
	V3 no, nd, ctxP, ctxN;	 // next ray + its LightSampleContext
	Spec nthp, npu, npl;
	int nflags, medium;
};
// LOCK: block-uniform -- every thread of the CTA is in here with an item, so the phase barriers are legal
template <int MT, bool MOTION>
KRR_DEV void scatterVertex(const Wavefront &wf, const ScatterIn &in, Pcg &rng, const bool lock, ScatterOut &out) {
#define KRR_PHASE() do { if (lock) __syncthreads(); } while (0)
	const float lightSelPdf = wf.scene.nLights > 0 ? 1.f / wf.scene.nLights : 0.f;
	const int pix = in.pix;
	out.pushShadow = out.pushNext = false;
	out.medium = in.medium, out.sMedium = -1, out.nflags = 0;
	Spec thp = in.thp / wf.p.probRR, pu = in.pu;
	SurfaceGeom g;
	KRR_PHASE();
	rebuildGeometry<MOTION>(wf, in.hit, mk3(in.d4), in.time, g);
	KRR_PHASE();
	float lam = wf.px.lambda[pix];
	Wavelengths wl = expandWavelengths(lam);
	ShadingData sd;
	bool term;
	evalMaterial(wf, g, wl, sd, term);
	if (term && lam > 0) { // lambda.terminateSecondary() persists in the pixel state (shading.h:184-185)
		wf.px.lambda[pix] = -lam;
		wl = expandWavelengths(-lam);
	}
	KRR_PHASE();
	auto toLocal = [&](V3 v) { return mk3(dot(g.tangent, v), dot(g.bitangent, v), dot(g.n, v)); };
	V3 woLocal = toLocal(g.wo);
	int bsdfType = getBsdfType(sd);
	Bsdf<MT> bsdf;
	BsdfSetupCtx ctx{g.wo, &wl, &wf.scene.cs};
	bsdf.setup(sd, ctx);
	KRR_PHASE();
	// [A] light sample of the next-event estimation (draws: light, u0, u1)
	bool needL = false, delta = false;
	LightSample ls;
	V3 p_o = mk3(0, 0, 0), dd = p_o, wiL = p_o;
	float lightPdf = 0;
	if (wf.p.nee && wf.scene.nLights > 0 && (bsdfType & BSDF_SMOOTH)) {
		float ul = rng.get1D();
		uint32_t lightId = (uint32_t) (ul * wf.scene.nLights);
		const LightRec lr = wf.scene.lights[lightId];
		float u0 = rng.get1D(), u1 = rng.get1D();
		if (lr.type == LIGHT_DIFFUSE_AREA) {
			const TriLightRec &tl = wf.scene.triLights[lr.index];
			ls = areaLightSampleLi(tl, wf.scene.instances[tl.inst], u0, u1, g.p, wl, wf.scene.cs);
		} else {
			const AnalyticLightRec &al = wf.scene.analytic[lr.index];
			ls	  = analyticSampleLi(al, u0, u1, g.p, wl, wf.scene);
			delta = al.type != LIGHT_INFINITE;
		}
		// spawnRayTo(ls.intr), raytracing.h:148-157
		auto offs = [](V3 p, V3 n, V3 w) { V3 off = n * kRayEps; if (dot(n, w) < 0.f) off = -off; return p + off; };
		V3 to = offs(ls.p, ls.n, g.p - ls.p);
		p_o	  = offs(g.p, g.n, to - g.p);
		dd	  = to - p_o;
		wiL	  = toLocal(normalize(dd));
		lightPdf = lightSelPdf * ls.pdf;
		needL	 = true;
	}
	// [B] BSDF value / pdf towards the light, and the BSDF sample (draws: lobe, u0, u1)
	Spec bsdfVal = sp(0);
	float bsdfPdf = 0;
	BSDFSample bs;
	if constexpr (MT == MAT_DISNEY) {
		// the diffuse lobe's sample is a cosine-distributed direction followed by f() and pdf(), the
		// same two functions the light direction needs: both go through ONE copy of the code
		const int comp = bsdf.pickLobe(rng);
		V3 wiS = mk3(0, 0, 1);
		if (comp == 0) {
			float u0 = rng.get1D(), u1 = rng.get1D();
			wiS = cosineSampleHemisphere(u0, u1);
			if (woLocal.z < 0) wiS.z *= -1;
		}
		Spec fS = sp(0);
		float pS = 0;
#pragma unroll 1
		for (int c = 0; c < 2; c++) {
			KRR_PHASE();
			const bool need = c == 0 ? needL : comp == 0;
			const V3 wi		= c == 0 ? wiL : wiS;
			Spec fv	 = sp(0);
			float pv = 0;
			if (need) bsdf.eval(woLocal, wi, fv, pv);
			if (c == 0) bsdfVal = fv, bsdfPdf = pv;
			else fS = fv, pS = pv;
		}
		KRR_PHASE();
		if (comp == 0) bs = BSDFSample{fS, wiS, pS, BSDF_DIFFUSE_REFLECTION};
		else bs = bsdf.sampleSpecular(comp, woLocal, rng);
	} else {
		if (needL) bsdfVal = bsdf.f(woLocal, wiL), bsdfPdf = bsdf.pdf(woLocal, wiL);
		KRR_PHASE();
		bs = bsdf.sample(woLocal, rng);
	}
	KRR_PHASE();
	if (delta) bsdfPdf = 0.f;
	if (needL && lightPdf > 0 && any(bsdfVal)) {
		Spec Ld = ls.L * thp * bsdfVal * fabsf(wiL.z);
		if (any(Ld)) {
			out.pushShadow = true;
			out.so = p_o, out.sdv = dd;
			const MeshRec &smesh = wf.scene.meshes[g.mesh]; // spawnRayTo: medium = getMedium(d)
			out.sMedium = smesh.mediumIn != smesh.mediumOut ? (dot(dd, g.n) > 0 ? smesh.mediumOut : smesh.mediumIn) : in.medium;
			out.sPu = pu * bsdfPdf, out.sPl = pu * lightPdf;
			out.sContrib = wf.p.enableMedium ? Ld : Ld / mean(out.sPl + out.sPu);
		}
	}
	if (bs.pdf != 0 && any(bs.f)) {
		V3 wiWorld = g.tangent * bs.wi.x + g.bitangent * bs.wi.y + g.n * bs.wi.z;
		out.nthp = thp * bs.f * fabsf(bs.wi.z) / bs.pdf;
		if (any(out.nthp)) {
			out.pushNext = true;
			out.nflags = bs.flags;
			out.npu = pu, out.npl = pu / bs.pdf;
			V3 off = g.n * kRayEps;
			if (dot(g.n, wiWorld) < 0.f) off = -off;
			out.no = g.p + off, out.nd = wiWorld;
			out.ctxP = g.p, out.ctxN = g.n;
			// Interaction::getMedium(dir), raytracing.h:163-167
			const MeshRec &mesh = wf.scene.meshes[g.mesh];
			if (mesh.mediumIn != mesh.mediumOut) out.medium = dot(wiWorld, g.n) > 0 ? mesh.mediumOut : mesh.mediumIn;
		}
	}
#undef KRR_PHASE
}

template <int MT, bool MOTION>
__global__ void __launch_bounds__(kScatterBlock, KRR_SCATTER_MINB) k_scatter(const __grid_constant__ Wavefront wf, int depth, int withHitMiss) {
	KRR_PDL_ENTRY();
	// handleHit / handleMiss of this depth (they touch L only, this stage does not)
	if (withHitMiss) handleHitMissBody<MOTION>(wf, depth);
	const RayQueue q  = wf.rays[depth & 1];
	const RayQueue nq = wf.rays[(depth & 1) ^ 1];
	DepthCounters *dc = wf.counters + depth;
	const int n		  = dc->nScatter[MT];
	const int stride  = gridDim.x * blockDim.x;
	const int nIter	  = (n + stride - 1) / stride;
	const bool implicit = wf.p.implicitDepth0 && depth == 0;
	for (int it = 0; it < nIter; it++) {
		int k		= it * stride + blockIdx.x * blockDim.x + threadIdx.x;
		bool active = k < n;
		// block-uniform: every thread of the CTA has an item and takes the same top-level path (with
		// rrInTrace no path ends inside this stage before the phases below), so barriers are legal
		const bool lock = KRR_SCATTER_LOCKSTEP && wf.p.rrInTrace && it * stride + (blockIdx.x + 1) * blockDim.x <= n;
		ScatterOut out;
		out.pushShadow = out.pushNext = false;
		int pix = 0, itemDepth = 0;
		float time = 0;
		if (active) {
			int i	  = wf.scatterIdx[MT][k];
			ScatterIn in;
			in.hit	  = wf.hits[i];
			float4 o4 = ldcs4(q.o_time + i);
			in.d4	  = ldcs4(q.d_medium + i);
			float4 cp = implicit ? make_float4(0, 0, 0, __int_as_float(i)) : ldcs4(q.ctxP_pix + i), cn = implicit ? make_float4(0, 0, 0, 0) : ldcs4(q.ctxN_dep + i);
			pix = __float_as_int(cp.w), itemDepth = __float_as_int(cn.w) & 0xff;
			time = o4.w;
			in.time = time, in.pix = pix, in.medium = __float_as_int(in.d4.w);
			Pcg rng{wf.px.rng[pix], rngIncOf(wf.p, pix)};
			// Russian roulette at every depth, including 0 (integrator.cpp:118-119)
			bool alive = wf.p.rrInTrace ? true : rng.get1D() < wf.p.probRR;
			if (alive) {
				in.thp = implicit ? sp(1) : ldcs4(q.thp + i), in.pu = implicit ? sp(1) : ldcs4(q.pu + i);
				scatterVertex<MT, MOTION>(wf, in, rng, lock, out);
			}
			wf.px.rng[pix] = rng.state;
		}
		int s = warpPush(&dc->nShadow, out.pushShadow);
		if (s >= 0) {
			stcs4(wf.shadow.o_tmax + s, make_float4(out.so.x, out.so.y, out.so.z, 1.f));
			stcs4(wf.shadow.d_pix + s, make_float4(out.sdv.x, out.sdv.y, out.sdv.z, __int_as_float(pix)));
			stcs4(wf.shadow.contrib + s, out.sContrib);
			if (wf.p.enableMedium) stcs4(wf.shadow.pu + s, out.sPu), stcs4(wf.shadow.pl + s, out.sPl);
			if (wf.p.enableMedium || MOTION) wf.shadow.aux[s] = make_int2(out.sMedium, __float_as_int(time));
		}
		s = warpPush(&dc[1].nRay, out.pushNext);
		if (s >= 0) {
			stcs4(nq.o_time + s, make_float4(out.no.x, out.no.y, out.no.z, time));
			stcs4(nq.d_medium + s, make_float4(out.nd.x, out.nd.y, out.nd.z, __int_as_float(out.medium)));
			stcs4(nq.thp + s, out.nthp);
			stcs4(nq.pu + s, out.npu);
			stcs4(nq.pl + s, out.npl);
			stcs4(nq.ctxP_pix + s, make_float4(out.ctxP.x, out.ctxP.y, out.ctxP.z, __int_as_float(pix)));
			stcs4(nq.ctxN_dep + s, make_float4(out.ctxN.x, out.ctxN.y, out.ctxN.z, __int_as_float((itemDepth + 1) | (out.nflags << 8))));
		}
	}
}

// =================================================================================================
// Shadow stage (device.cu:83-100): any-hit visibility, L += Ld / (pl + pu).mean().  Same persistent
// warp scheme as the closest stage; the ray terminates at the first accepted hit.
template <int MODE>
KRR_DEV void traceShadowBody(const Wavefront &wf, int depth, TraceSmem &sm, int block, int nBlocks) {
	constexpr bool MOTION = MODE == kTraceMotion;
	DepthCounters *dc	= wf.counters + depth;
	const int n			= dc->nShadow;
	const unsigned FULL = 0xffffffffu;
	const int lane		= threadIdx.x & 31;
	Traverser<true, MOTION, MODE == kTraceFlat> tr;
	LocalStack<true> ls;
	int ray = -1, pix = 0;
	bool done = false; // traversal finished, radiance not yet added
	WarpWork work;
	if (MODE == kTraceFlat) work.init(n, &dc->cursorShadow, (block * kTraceBlock + (int) threadIdx.x) >> 5, (nBlocks * kTraceBlock) >> 5);
	else work.initSpread(n, &dc->cursorShadow, block, nBlocks, (int) threadIdx.x >> 5, kTraceBlock >> 5);
	if (work.next >= n) return;
	const int32_t *__restrict__ perm = wf.permShadow;
	while (true) {
		// finished rays add their contribution in batches (same reasoning as in the closest stage: the
		// read-modify-write of L is a long-latency chain the whole warp would wait for on every trip)
		const unsigned doneMask = __ballot_sync(FULL, done);
		if (doneMask && (__popc(doneMask) >= wf.p.refill || !__ballot_sync(FULL, ray >= 0 && !done))) {
			if (done) {
				if (tr.overflow) atomicExch(&wf.errorFlags[0], 1);
				if (tr.best.inst < 0) wf.px.L[pix] = ldcs4(wf.shadow.contrib + ray) + wf.px.L[pix];
				ray = -1, done = false;
			}
		}
		unsigned idle = __ballot_sync(FULL, ray < 0);
		if (!work.exhausted && __popc(idle) >= wf.p.refill) {
			int r = work.take(idle, lane);
			if (r >= 0) {
				if (perm) r = __ldg(perm + r);
				ray = r;
				float4 o4 = ldcs4(wf.shadow.o_tmax + r), d4 = ldcs4(wf.shadow.d_pix + r);
				pix = __float_as_int(d4.w);
				tr.begin(wf.bvh, mk3(o4), mk3(d4), o4.w, MOTION ? __int_as_float(wf.shadow.aux[r].y) : 0.f);
			}
			idle = __ballot_sync(FULL, ray < 0);
		}
		if (idle == FULL) {
			if (work.exhausted) break;
			continue;
		}
#ifndef KRR_SHADOW_VOTE
#define KRR_SHADOW_VOTE false
#endif
		const bool fin = tr.trip<KRR_SHADOW_VOTE>(ray >= 0 && !done, wf.bvh, wf.scene.instances, sm, ls, [&](int inst, int prim, float u, float v) {
			uint8_t f = wf.instFlags[inst];
			if (f & 1) return false; // __anyhit__Shadow ignores null-material surfaces
			if (f & 2) return !alphaKilled(wf, inst, prim, u, v, tr.o, tr.d);
			return true;
		});
		done |= fin;
	}
}

template <int MODE>
__global__ void __launch_bounds__(kTraceBlock) k_trace_shadow(const __grid_constant__ Wavefront wf, int depth) {
	KRR_PDL_ENTRY();
	__shared__ TraceSmem sm;
	traceShadowBody<MODE>(wf, depth, sm, blockIdx.x, gridDim.x);
}

// =================================================================================================
// Ray reordering.  The scatter stage emits the rays of the next depth in the order its items were pushed: origins
// scattered over the scene, directions over the sphere.  A trace warp that takes 32 consecutive queue slots then
// holds 32 unrelated rays -- every lane in another instance / subtree, in another traversal phase (SIMT efficiency
// 13 of 32 on the 10 000-instance scene), every node fetch a different line.  k_sort_rays re-orders the queue for
// the trace stage WITHOUT moving the items: every CTA takes a tile of kSortTile consecutive slots, gives each ray
// the key (direction octant, Morton code of its origin in the frame of the BVH root node), sorts the tile's
// (key, slot) pairs in shared memory (cub::BlockRadixSort) and writes the sorted slots to `perm`; the trace
// warps read their slots through it.  Sorting inside tiles needs no global pass and no host-side size (the queue
// length lives on the device); rays of one tile that share an octant and a region end up in the same warps.
// Pixels are independent (own RNG stream, own accumulator; one ray per pixel and depth), so the order in which a
// stage processes its queue does not change a single draw or addition: films are bit-identical
// (tests/test_gpu_sort_rays.py).
constexpr int kSortThreads = 256, kSortItems = 16, kSortTile = kSortThreads * kSortItems, kSortMortonBits = 6;
constexpr int kSortKeyBits = 3 + 3 * kSortMortonBits;

KRR_DEV uint32_t spreadBits3(uint32_t v) { // bit i of a 10-bit value -> bit 3 i
	v = (v | (v << 16)) & 0x030000ffu;
	v = (v | (v << 8)) & 0x0300f00fu;
	v = (v | (v << 4)) & 0x030c30c3u;
	v = (v | (v << 2)) & 0x09249249u;
	return v;
}

// frame of the Morton code: the quantisation frame of the BVH root node (anchor, 256 steps of 2^(e-127) per axis)
struct SortFrame { float ox, oy, oz, ix, iy, iz; };
KRR_DEV SortFrame sortFrame(const BvhDev &bvh) {
	const uint32_t root = bvh.mergedOnly ? (uint32_t) bvh.mergedRoot : (uint32_t) bvh.tlasRoot;
	SortFrame f{0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
	if (isFlatEntry(root)) return f;
	const float4 n0	  = __ldg(reinterpret_cast<const float4 *>(bvh.nodes + root));
	const uint32_t ew = __float_as_uint(n0.w);
	const float k	  = (float) (1 << kSortMortonBits) / 256.f;
	f.ox = n0.x, f.oy = n0.y, f.oz = n0.z;
	f.ix = k / __uint_as_float((ew & 0xff) << 23), f.iy = k / __uint_as_float(((ew >> 8) & 0xff) << 23), f.iz = k / __uint_as_float(((ew >> 16) & 0xff) << 23);
	return f;
}
KRR_DEV uint32_t rayKey(const SortFrame &f, float4 o, float4 d, int mode) {
	const float top	   = (float) ((1 << kSortMortonBits) - 1);
	const uint32_t qx  = (uint32_t) fminf(fmaxf((o.x - f.ox) * f.ix, 0.f), top); // NaN -> 0
	const uint32_t qy  = (uint32_t) fminf(fmaxf((o.y - f.oy) * f.iy, 0.f), top);
	const uint32_t qz  = (uint32_t) fminf(fmaxf((o.z - f.oz) * f.iz, 0.f), top);
	const uint32_t mor = (spreadBits3(qx) << 2) | (spreadBits3(qy) << 1) | spreadBits3(qz);
	const uint32_t oct = (d.x >= 0.f ? 4u : 0u) | (d.y >= 0.f ? 2u : 0u) | (d.z >= 0.f ? 1u : 0u);
	return mode == 0 ? (oct << (3 * kSortMortonBits)) | mor : (mor << 3) | oct;
}

// closest rays of loop depth `depth + 1` (perm = wf.permClosest) and shadow rays of `depth` (wf.permShadow): the two
// queues the scatter stage of `depth` filled, i.e. what the fused trace launch of `depth` consumes
__global__ void __launch_bounds__(kSortThreads) k_sort_rays(const __grid_constant__ Wavefront wf, int depth, int32_t *permClosest, int32_t *permShadow) {
	using Sort = cub::BlockRadixSort<uint32_t, kSortThreads, kSortItems, uint32_t>;
	__shared__ typename Sort::TempStorage tmp;
	const int nC = permClosest ? wf.counters[depth + 1].nRay : 0, nS = permShadow ? wf.counters[depth].nShadow : 0;
	const int tilesC = (nC + kSortTile - 1) / kSortTile, tilesS = (nS + kSortTile - 1) / kSortTile;
	const RayQueue q  = wf.rays[(depth + 1) & 1];
	const SortFrame f = sortFrame(wf.bvh);
	for (int tile = blockIdx.x; tile < tilesC + tilesS; tile += gridDim.x) {
		const bool shadow	= tile >= tilesC;
		const int base		= (shadow ? tile - tilesC : tile) * kSortTile;
		const int n			= shadow ? nS : nC;
		const float4 *org	= shadow ? wf.shadow.o_tmax : q.o_time;
		const float4 *dir	= shadow ? wf.shadow.d_pix : q.d_medium;
		int32_t *perm		= shadow ? permShadow : permClosest;
		uint32_t keys[kSortItems], vals[kSortItems];
#pragma unroll
		for (int k = 0; k < kSortItems; k++) { // striped reads: coalesced
			const int local = k * kSortThreads + (int) threadIdx.x, i = base + local;
			vals[k] = (uint32_t) local;
			keys[k] = i < n ? rayKey(f, __ldg(org + i), __ldg(dir + i), wf.p.sortKey) : (1u << kSortKeyBits); // past the end: sorts last
		}
		Sort(tmp).SortBlockedToStriped(keys, vals, 0, kSortKeyBits + 1);
#pragma unroll
		for (int k = 0; k < kSortItems; k++) {
			const int i = base + k * kSortThreads + (int) threadIdx.x;
			if (i < n) perm[i] = base + (int) vals[k];
		}
		__syncthreads(); // tmp is reused by the next tile
	}
}

// Global variant ("sort_rays" bit 2): keys of a whole queue for a device-wide radix sort (cub::DeviceRadixSort in
// krr_wfpt_render).  The queue length lives on the device, the sort's item count on the host: the sort covers the
// queue's CAPACITY and slots past the end get a key above every real one, so that perm[0 .. n) are the live slots.
__global__ void k_ray_keys(const __grid_constant__ Wavefront wf, int depth, int shadow, uint32_t *keys, int32_t *vals, int capacity) {
	const int n		  = shadow ? wf.counters[depth].nShadow : wf.counters[depth + 1].nRay;
	const RayQueue q  = wf.rays[(depth + 1) & 1];
	const float4 *org = shadow ? wf.shadow.o_tmax : q.o_time;
	const float4 *dir = shadow ? wf.shadow.d_pix : q.d_medium;
	const SortFrame f = sortFrame(wf.bvh);
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < capacity; i += gridDim.x * blockDim.x) {
		keys[i] = i < n ? rayKey(f, __ldg(org + i), __ldg(dir + i), wf.p.sortKey) : (1u << kSortKeyBits);
		vals[i] = i;
	}
}

// Fused trace stage: the shadow rays of `depth` and the closest rays of `depth + 1` were both produced
// by the scatter stage of `depth` and do not depend on each other (the shadow stage only adds to L, the
// closest stage only draws the Russian-roulette sample; handleHit/Miss of depth + 1 runs afterwards, so
// the order of the additions to a pixel's L is unchanged).  One launch traces both: every warp first
// works on the closest queue and, when that is exhausted, moves on to the shadow queue, so the tail of
// one overlaps the head of the other and a launch per depth is saved.  (A static split of the grid
// between the two queues measured slower than two launches: the cost ratio of the two ray kinds moves
// with depth.)  Not used with participating media (the transmittance shadow rays draw random numbers).
// 7 CTAs/SM = the occupancy of the two stand-alone kernels (72 registers); the flat-list instantiation runs
// 6 CTAs/SM (80 registers: its branch-free triangle pairs spill at 72; A/B on the bench workload: 6 CTAs
// 4158, 7 CTAs 4062, 5 CTAs 4078 Mrays/s)
#ifndef KRR_FUSED_MINB
#define KRR_FUSED_MINB 6
#endif
#ifndef KRR_TREE_MINB
#define KRR_TREE_MINB 7
#endif
// static tree scenes: 6 CTAs/SM at 80 registers (20 M triangles: 1 195 -> 1 256 Mrays/s against 7 CTAs at 72; the
// moving-instance kernel is the other way round: 7 CTAs 695, 6 CTAs 646, 8 CTAs 665 Mrays/s)
#ifndef KRR_STATIC_MINB
#define KRR_STATIC_MINB 6
#endif
template <int MODE>
__global__ void __launch_bounds__(kTraceBlock, MODE == kTraceFlat ? KRR_FUSED_MINB : MODE == kTraceStatic ? KRR_STATIC_MINB : KRR_TREE_MINB) k_trace_fused(
const __grid_constant__ Wavefront wf, int depth) {
	KRR_PDL_ENTRY();
	__shared__ TraceSmem sm;
	traceClosestBody<MODE>(wf, depth + 1, sm, blockIdx.x, gridDim.x);
	traceShadowBody<MODE>(wf, depth, sm, blockIdx.x, gridDim.x);
}

// =================================================================================================
// Tail kernel.  Deep bounces hold few rays, but a stage launch lasts as long as its slowest ray (20 M triangles:
// ~200-450 us per depth whatever the queue holds; Cornell box: 25-35 us), and the depth loop pays that once per
// stage per depth.  From loop depth `depth` on, ONE launch finishes the paths: every lane takes a ray item of that
// depth and runs its path to the end -- closest hit, emitted / environment radiance, Russian roulette, BSDF
// sampling with next-event estimation, shadow ray, next bounce -- with the routines of the stage kernels
// (scatterVertex, hitLightItem, missItem, the same traversal).  Per pixel the order of the random draws and of
// the additions to L is the order of the depth loop (integrator.cpp:232-256), so the film is bit-identical
// (tests/test_gpu_tail.py); the per-depth counters are kept with atomics (the tail holds few rays by design).
// Surface-only scenes with NEE (the fused schedule); the shadow rays of depth - 1 are traced by a k_trace_shadow
// launch before it, so that their contribution reaches L before this kernel's.
template <int MT, bool MOTION>
__device__ __noinline__ void tailVertex(const Wavefront &wf, const ScatterIn &in, Pcg &rng, ScatterOut &out) {
	scatterVertex<MT, MOTION>(wf, in, rng, false, out);
}

#ifndef KRR_TAIL_MINB
#define KRR_TAIL_MINB 3
#endif
// lanes that must be waiting for the shading phase before a warp runs it (or: no lane of the warp is tracing)
#ifndef KRR_TAIL_BATCH
#define KRR_TAIL_BATCH 8
#endif
template <int MODE>
__global__ void __launch_bounds__(kTraceBlock, KRR_TAIL_MINB) k_tail(const __grid_constant__ Wavefront wf, int depth) {
	constexpr bool MOTION = MODE == kTraceMotion;
	__shared__ TraceSmem sm;
	const unsigned FULL = 0xffffffffu;
	const int lane		= threadIdx.x & 31;
	const RayQueue q	= wf.rays[depth & 1];
	const int n			= wf.counters[depth].nRay;
	const float lightSelPdf = wf.scene.nLights > 0 ? 1.f / wf.scene.nLights : 0.f;
	// A warp is a small wavefront of its own, WITHOUT a barrier between its stages: every lane holds one path and
	// is in one of three states -- tracing (the shadow ray of its last vertex, then the ray to its next vertex),
	// ready (next vertex found, waiting to be shaded), idle.  The warp runs traversal trips for its tracing lanes and,
	// as soon as KRR_TAIL_BATCH lanes are ready (or none is tracing), ONE pass of the shading phase -- emitted /
	// environment radiance, Russian roulette, BSDF sampling + NEE -- for all ready lanes, which sends them tracing
	// again.  A slow ray delays nobody: the other lanes of its warp go on through their bounces around it, and a
	// lane whose path has ended takes the next ray of the queue.  Short queues are spread over all warps.
	enum { T_IDLE = 0, T_CLOSEST = 1, T_SHADOW = 2, T_READY = 3 };
	WarpWork work;
	work.initSpread(n, &wf.counters[depth].cursorRay, blockIdx.x, gridDim.x, (int) threadIdx.x >> 5, kTraceBlock >> 5);
	if (work.next >= n) return;
	int state = T_IDLE;
	float4 o4 = make_float4(0, 0, 0, 0), d4 = o4, cp = o4;
	Spec thp = sp(0), pu = sp(0), pl = sp(0), L = sp(0), sContrib = sp(0);
	int pix = 0, packed = 0, loopDepth = depth; // packed = item depth | bsdfType << 8
	bool hasNext = false;
	Pcg rng{0, wf.p.rngInc};
	Traverser<false, MOTION, MODE == kTraceFlat> tr;
	LocalStack<false> ls;
	auto endPath = [&]() {
		wf.px.L[pix]   = L;
		wf.px.rng[pix] = rng.state;
		state		   = T_IDLE;
	};
	auto beginClosest = [&]() { // the ray in (o4, d4) becomes the closest ray of the next loop depth
		loopDepth++;
		atomicAdd(&wf.counters[loopDepth].nRay, 1);
		tr.begin(wf.bvh, mk3(o4), mk3(d4), kInf, o4.w);
		state = T_CLOSEST;
	};
	while (true) {
		const unsigned idle = __ballot_sync(FULL, state == T_IDLE);
		if (idle && !work.exhausted) {
			const int r = work.take(idle, lane);
			if (r >= 0) {
				o4 = ldcs4(q.o_time + r), d4 = ldcs4(q.d_medium + r);
				thp = ldcs4(q.thp + r), pu = ldcs4(q.pu + r), pl = ldcs4(q.pl + r);
				cp = ldcs4(q.ctxP_pix + r);
				packed = __float_as_int(__ldcs(reinterpret_cast<const float *>(q.ctxN_dep + r) + 3));
				pix = __float_as_int(cp.w);
				L = wf.px.L[pix];
				rng.state = wf.px.rng[pix], rng.inc = rngIncOf(wf.p, pix);
				loopDepth = depth; // (the queue of `depth` itself was counted by its producer)
				tr.begin(wf.bvh, mk3(o4), mk3(d4), kInf, o4.w);
				state = T_CLOSEST;
			}
		}
		const unsigned mTrace = __ballot_sync(FULL, state == T_CLOSEST || state == T_SHADOW), mReady = __ballot_sync(FULL, state == T_READY);
		if (!mTrace && !mReady) {
			if (work.exhausted) break;
			continue;
		}
		if (mReady && (__popc(mReady) >= KRR_TAIL_BATCH || !mTrace)) {
			// ================= shading phase for the ready lanes =================
			const bool ready  = state == T_READY;
			DepthCounters *dc = wf.counters + loopDepth;
			bool ended = false, wantScatter = false;
			int mt = 0;
			ScatterIn in;
			// ---- [hit / miss / Russian roulette] integrator.cpp:78-108, 118-119 ----
			if (ready) {
				const Hit h = tr.best;
				if (tr.overflow) atomicExch(&wf.errorFlags[0], 1);
				if (h.inst < 0) {
					atomicAdd(&dc->nMiss, 1);
					if (wf.scene.nInfinite) L = missItem(wf, d4, pix, packed, thp, pu, pl, lightSelPdf) + L;
					ended = true;
				} else {
					const uint32_t f = wf.instFlags[h.inst];
					if (f & 1) { // null material: same item depth, next loop depth (never traced behind the last one)
						if (loopDepth == wf.p.maxDepth) ended = true;
						else {
							continueThroughNull<MOTION>(wf, h, o4, d4);
							beginClosest();
						}
					} else {
						in.hit = make_int4(h.inst, h.prim, __float_as_int(h.u), __float_as_int(h.v));
						if (f & 4) {
							atomicAdd(&dc->nHitLight, 1);
							L = hitLightItem<MOTION>(wf, in.hit, d4, o4.w, cp, packed, thp, pu, pl, lightSelPdf) + L;
						}
						mt = (int) (f >> 4);
						if (loopDepth == wf.p.maxDepth) { // pushed, but the depth loop ends before the scatter stage
							atomicAdd(&dc->nScatter[mt], 1);
							ended = true;
						} else if (!(rng.get1D() < wf.p.probRR)) { // the next draw of the pixel's stream
							atomicAdd(wf.p.rrInTrace ? &dc->nScatterKilled : &dc->nScatter[mt], 1);
							ended = true;
						} else {
							atomicAdd(&dc->nScatter[mt], 1);
							in.d4 = d4, in.time = o4.w, in.pix = pix, in.medium = __float_as_int(d4.w), in.thp = thp, in.pu = pu;
							wantScatter = true;
						}
					}
				}
			}
			// ---- [scatter] integrator.cpp:120-163: one material type at a time, each for all its lanes ----
			ScatterOut out;
			out.pushShadow = out.pushNext = false;
			const unsigned types = __reduce_or_sync(FULL, wantScatter ? 1u << mt : 0u);
			if ((types >> MAT_DISNEY) & 1u) { if (wantScatter && mt == MAT_DISNEY) tailVertex<MAT_DISNEY, MOTION>(wf, in, rng, out); }
			if ((types >> MAT_DIFFUSE) & 1u) { if (wantScatter && mt == MAT_DIFFUSE) tailVertex<MAT_DIFFUSE, MOTION>(wf, in, rng, out); }
			if ((types >> MAT_DIELECTRIC) & 1u) { if (wantScatter && mt == MAT_DIELECTRIC) tailVertex<MAT_DIELECTRIC, MOTION>(wf, in, rng, out); }
			if ((types >> MAT_CONDUCTOR) & 1u) { if (wantScatter && mt == MAT_CONDUCTOR) tailVertex<MAT_CONDUCTOR, MOTION>(wf, in, rng, out); }
			if ((types >> MAT_NULL) & 1u) { if (wantScatter && mt == MAT_NULL) tailVertex<MAT_NULL, MOTION>(wf, in, rng, out); }
			if (wantScatter) {
				hasNext = out.pushNext;
				const float time = o4.w;
				if (out.pushNext) { // the path's next ray waits in the path state while the shadow ray is traced
					o4 = make_float4(out.no.x, out.no.y, out.no.z, time);
					d4 = make_float4(out.nd.x, out.nd.y, out.nd.z, __int_as_float(out.medium));
					thp = out.nthp, pu = out.npu, pl = out.npl;
					cp = make_float4(out.ctxP.x, out.ctxP.y, out.ctxP.z, cp.w);
					packed = ((packed & 0xff) + 1) | (out.nflags << 8);
				}
				if (out.pushShadow) { // [shadow] device.cu:83-100, before the next vertex can add to L
					atomicAdd(&dc->nShadow, 1);
					sContrib = out.sContrib;
					tr.begin(wf.bvh, out.so, out.sdv, 1.f, time);
					state = T_SHADOW;
				} else if (hasNext) beginClosest();
				else ended = true;
			}
			if (ended) endPath();
			continue;
		}
		// ================= one traversal trip for the tracing lanes =================
		const bool tracing = state == T_CLOSEST || state == T_SHADOW, isShadow = state == T_SHADOW;
		const bool fin = tr.template trip<true>(tracing, wf.bvh, wf.scene.instances, sm, ls, [&](int inst, int prim, float u, float v) {
			const uint8_t fl = wf.instFlags[inst];
			if (isShadow && (fl & 1)) return false; // __anyhit__Shadow ignores null-material surfaces
			if (fl & 2) return !alphaKilled(wf, inst, prim, u, v, tr.o, tr.d);
			return true;
		});
		if (isShadow) {
			if (fin || tr.best.inst >= 0) { // any accepted hit occludes
				if (tr.overflow) atomicExch(&wf.errorFlags[0], 1);
				if (tr.best.inst < 0) L = sContrib + L;
				if (hasNext) beginClosest();
				else endPath();
			}
		} else if (tracing && fin) state = T_READY;
	}
}

// =================================================================================================
// Participating media (BASELINE config 4).  Stage map: k_medium_sample = sampleMediumInteraction
// (medium.cpp:13-103), k_medium_scatter = sampleMediumScattering (medium.cpp:105-153),
// k_trace_shadow_tr = __raygen__ShadowTr + traceTransmittance (device.cu:102-128, wavefront.h:80-139).
// A medium-sample item is the ray slot itself (+ hit record + hit distance): the stage updates the
// slot's thp / pu / pl in place and then routes the slot to the miss / hit-light / scatter index
// queues, exactly where the surface-only path would have put it.

// Interaction::getMedium(w) for a surface hit (raytracing.h:162-166)
KRR_DEV int mediumAcross(const MeshRec &mesh, V3 n, V3 w, int rayMedium) {
	return mesh.mediumIn != mesh.mediumOut ? (dot(w, n) > 0 ? mesh.mediumOut : mesh.mediumIn) : rayMedium;
}

__global__ void __launch_bounds__(128) k_medium_sample(const __grid_constant__ Wavefront wf, int depth) {
	const RayQueue q  = wf.rays[depth & 1];
	const RayQueue nq = wf.rays[(depth & 1) ^ 1];
	DepthCounters *dc = wf.counters + depth;
	const int n		  = dc->nMediumSample;
	const int stride  = gridDim.x * blockDim.x;
	const int nIter	  = (n + stride - 1) / stride;
	for (int iter = 0; iter < nIter; iter++) {
		const int k = iter * stride + blockIdx.x * blockDim.x + threadIdx.x;
		int route = -1; // 0 miss, 1 surface (scatter + light), 2 null pass-through
		bool pushMS = false, light = false;
		int matType = 0, i = 0, pix = 0, itemDepth = 0, medium = -1;
		float4 o4 = make_float4(0, 0, 0, 0), d4 = o4;
		V3 msP = mk3(0, 0, 0);
		Spec thp = sp(0), pu = sp(0), pl = sp(0);
		Hit h;
		h.inst = -1, h.prim = -1, h.t = 0, h.u = h.v = 0;
		if (k < n) {
			i  = wf.mediumSampleIdx[k];
			o4 = ldcs4(q.o_time + i), d4 = ldcs4(q.d_medium + i);
			float4 cp = ldg4(q.ctxP_pix + i), cn = ldg4(q.ctxN_dep + i);
			pix = __float_as_int(cp.w), itemDepth = __float_as_int(cn.w) & 0xff, medium = __float_as_int(d4.w);
			thp = ldcs4(q.thp + i), pu = ldcs4(q.pu + i), pl = ldcs4(q.pl + i);
			int4 hit = wf.hits[i];
			h.inst = hit.x, h.prim = hit.y, h.u = __int_as_float(hit.z), h.v = __int_as_float(hit.w);
			const float tMax = wf.hitT[i];
			Pcg rng{wf.px.rng[pix], rngIncOf(wf.p, pix)};
			Wavelengths wl = expandWavelengths(wf.px.lambda[pix]);
			const MediumRec &med = wf.scene.media[medium];
			V3 o = mk3(o4), d = mk3(d4);
			Spec L = sp(0);
			bool scattered = false;
			Spec T_maj = sampleT_maj(med, wf.scene, o, d, tMax, rng, wl, [&](V3 p, const MediumPoint &mp, Spec sigma_maj, Spec Tm) -> bool {
				if (itemDepth < wf.p.maxDepth && any(mp.Le)) {
					float pr = sigma_maj.x * Tm.x;
					Spec pe	 = pu * sigma_maj * Tm / pr;
					if (any(pe)) L += thp * mp.sigma_a * Tm * mp.Le / (pr * mean(pe));
				}
				float pAbsorb = mp.sigma_a.x / sigma_maj.x, pScatter = mp.sigma_s.x / sigma_maj.x;
				float pNull	  = fmaxf(0.f, 1.f - pAbsorb - pScatter);
				int mode	  = sampleDiscrete3(pAbsorb, pScatter, pNull, rng.get1D());
				if (mode == 0) { // absorbed
					thp = sp(0);
					return false;
				} else if (mode == 1) { // real scattering
					float pr = Tm.x * mp.sigma_s.x;
					thp *= Tm * mp.sigma_s / pr;
					pu *= Tm * mp.sigma_s / pr;
					if (any(thp) && any(pu)) pushMS = true, msP = p;
					scattered = true;
					return false;
				} else { // null collision
					Spec sigma_n = cwiseMax(sigma_maj - mp.sigma_a - mp.sigma_s, 0.f);
					float pr	 = Tm.x * sigma_n.x;
					thp *= Tm * sigma_n / pr;
					if (pr == 0) thp = sp(0);
					pu *= Tm * sigma_n / pr;
					pl *= Tm * sigma_maj / pr;
					return any(thp) && any(pu);
				}
			});
			wf.px.rng[pix] = rng.state;
			if (any(L)) wf.px.L[pix] = L + wf.px.L[pix];
			if (!scattered && any(thp)) {
				thp *= T_maj / T_maj.x;
				pu *= T_maj / T_maj.x;
				pl *= T_maj / T_maj.x;
			}
			if (!(scattered || !any(thp) || !any(pu) || itemDepth == wf.p.maxDepth)) {
				// the ray survived to the end of its segment: update the slot, route it like a surface-only ray
				q.thp[i] = thp, q.pu[i] = pu, q.pl[i] = pl;
				if (h.inst < 0) route = 0;
				else {
					const InstRec &in	= wf.scene.instances[h.inst];
					const MeshRec &mesh = wf.scene.meshes[in.mesh];
					if (mesh.material < 0) route = 2;
					else {
						route	= 1;
						matType = wf.scene.materials[mesh.material].bsdfType;
						light	= in.lightBase >= 0;
					}
				}
			}
		}
		int s = warpPush(&dc->nMediumScatter, pushMS);
		if (s >= 0) {
			V3 d = mk3(d4);
			stcs4(wf.mscatter.p_time + s, make_float4(msP.x, msP.y, msP.z, o4.w));
			stcs4(wf.mscatter.wo_medium + s, make_float4(-d.x, -d.y, -d.z, d4.w));
			stcs4(wf.mscatter.thp + s, thp);
			stcs4(wf.mscatter.pu + s, pu);
			wf.mscatter.pix_depth[s] = make_int2(pix, itemDepth);
		}
		s = warpPush(&dc->nMiss, route == 0);
		if (s >= 0) wf.missIdx[s] = i;
		s = warpPush(&dc->nHitLight, route == 1 && light);
		if (s >= 0) wf.hitLightIdx[s] = i;
#pragma unroll
		for (int mt = 0; mt < MAT_COUNT; mt++) {
			s = warpPush(&dc->nScatter[mt], route == 1 && matType == mt);
			if (s >= 0) wf.scatterIdx[mt][s] = i;
		}
		s = warpPush(&dc[1].nRay, route == 2);
		if (s >= 0) requeueThroughNull(wf, q, nq, i, s, h, o4, d4);
	}
}

__global__ void __launch_bounds__(128) k_medium_scatter(const __grid_constant__ Wavefront wf, int depth) {
	const RayQueue nq = wf.rays[(depth & 1) ^ 1];
	DepthCounters *dc = wf.counters + depth;
	const int n		  = dc->nMediumScatter;
	const int stride  = gridDim.x * blockDim.x;
	const int nIter	  = (n + stride - 1) / stride;
	const float lightSelPdf = wf.scene.nLights > 0 ? 1.f / wf.scene.nLights : 0.f;
	for (int iter = 0; iter < nIter; iter++) {
		const int k = iter * stride + blockIdx.x * blockDim.x + threadIdx.x;
		bool pushShadow = false, pushNext = false;
		V3 p = mk3(0, 0, 0), sdv = p, nd = p;
		Spec Ld = sp(0), sPu = Ld, sPl = Ld, nthp = Ld, wpu = Ld, npl = Ld;
		int pix = 0, itemDepth = 0, medium = -1;
		float time = 0;
		if (k < n) {
			float4 pt = ldcs4(wf.mscatter.p_time + k), wm = ldcs4(wf.mscatter.wo_medium + k);
			p = mk3(pt), time = pt.w, medium = __float_as_int(wm.w);
			V3 wo = mk3(wm);
			Spec wthp = ldcs4(wf.mscatter.thp + k);
			wpu		  = ldcs4(wf.mscatter.pu + k);
			int2 pd	  = wf.mscatter.pix_depth[k];
			pix = pd.x, itemDepth = pd.y;
			const float g = wf.scene.media[medium].g;
			Pcg rng{wf.px.rng[pix], rngIncOf(wf.p, pix)};
			Wavelengths wl = expandWavelengths(wf.px.lambda[pix]);
			if (wf.p.nee && wf.scene.nLights > 0) { // [PART-A] direct lighting through ShadowTr (no light, no draws: as in the surface stage)
				float ul = rng.get1D();
				uint32_t lightId  = (uint32_t) (ul * wf.scene.nLights);
				const LightRec lr = wf.scene.lights[lightId];
				float u0 = rng.get1D(), u1 = rng.get1D();
				LightSample ls;
				bool delta = false;
				if (lr.type == LIGHT_DIFFUSE_AREA) {
					const TriLightRec &tl = wf.scene.triLights[lr.index];
					ls = areaLightSampleLi(tl, wf.scene.instances[tl.inst], u0, u1, p, wl, wf.scene.cs);
				} else {
					const AnalyticLightRec &al = wf.scene.analytic[lr.index];
					ls	  = analyticSampleLi(al, u0, u1, p, wl, wf.scene);
					delta = al.type != LIGHT_INFINITE;
				}
				// Interaction(p, time, medium).spawnRayTo(ls.intr): n = 0, the origin is not offset
				V3 off = ls.n * kRayEps;
				if (dot(ls.n, p - ls.p) < 0.f) off = -off;
				V3 dd = (ls.p + off) - p;
				V3 wi = normalize(dd);
				float ph	   = hgP(g, wo, wi);
				float lightPdf = lightSelPdf * ls.pdf;
				float phasePdf = delta ? 0.f : ph;
				Spec l = wthp * ph * ls.L;
				if (any(l) && lightPdf > 0) pushShadow = true, sdv = dd, Ld = l, sPl = wpu * lightPdf, sPu = wpu * phasePdf;
			}
			// [PART-B] phase-function sampling + Russian roulette
			float u0 = rng.get1D(), u1 = rng.get1D();
			float php, phpdf;
			hgSample(g, wo, u0, u1, nd, php, phpdf);
			nthp		 = wthp * php / phpdf;
			float rrProb = maxCoeff(nthp / mean(wpu));
			bool killed	 = false;
			if (itemDepth >= 1 && rrProb < 1) {
				if (rng.get1D() >= rrProb) killed = true;
				else nthp = nthp / rrProb;
			}
			wf.px.rng[pix] = rng.state;
			if (!killed && any(nthp) && !hasNaN(nthp)) pushNext = true, npl = wpu / phpdf;
		}
		int s = warpPush(&dc->nShadow, pushShadow);
		if (s >= 0) {
			stcs4(wf.shadow.o_tmax + s, make_float4(p.x, p.y, p.z, 1.f));
			stcs4(wf.shadow.d_pix + s, make_float4(sdv.x, sdv.y, sdv.z, __int_as_float(pix)));
			stcs4(wf.shadow.contrib + s, Ld);
			stcs4(wf.shadow.pu + s, sPu), stcs4(wf.shadow.pl + s, sPl);
			wf.shadow.aux[s] = make_int2(medium, __float_as_int(time));
		}
		s = warpPush(&dc[1].nRay, pushNext);
		if (s >= 0) {
			stcs4(nq.o_time + s, make_float4(p.x, p.y, p.z, time));
			stcs4(nq.d_medium + s, make_float4(nd.x, nd.y, nd.z, __int_as_float(medium)));
			stcs4(nq.thp + s, nthp);
			stcs4(nq.pu + s, wpu);
			stcs4(nq.pl + s, npl);
			stcs4(nq.ctxP_pix + s, make_float4(p.x, p.y, p.z, __int_as_float(pix)));
			stcs4(nq.ctxN_dep + s, make_float4(0, 0, 0, __int_as_float((itemDepth + 1) | (BSDF_SMOOTH << 8))));
		}
	}
}

// ShadowTr: transmittance along the shadow ray by ratio tracking, stepping through null-material
// interfaces; an opaque surface ends it.  One lane runs the whole chain of one shadow ray.
// MODE: kTraceMotion = any scene (motion transforms evaluated where an instance has them), kTraceFlat = static
// scene that is one flat triangle list (branch-free triangle pairs).
template <int MODE>
__global__ void __launch_bounds__(kTraceBlock) k_trace_shadow_tr(const __grid_constant__ Wavefront wf, int depth) {
	__shared__ TraceSmem sm;
	DepthCounters *dc = wf.counters + depth;
	const int n		  = dc->nShadow;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
		float4 o4 = ldcs4(wf.shadow.o_tmax + i), d4 = ldcs4(wf.shadow.d_pix + i);
		const int pix = __float_as_int(d4.w);
		int2 aux	  = wf.shadow.aux[i];
		V3 ro = mk3(o4), rd = mk3(d4);
		int medium		 = aux.x;
		const float tMax = o4.w;
		const V3 pLight	 = ro + rd * tMax;
		Spec T_ray = sp(1), pu = sp(1), pl = sp(1);
		Pcg rng{wf.px.rng[pix], rngIncOf(wf.p, pix)};
		Wavelengths wl = expandWavelengths(wf.px.lambda[pix]);
		bool usedRng   = false;
		// SurfaceInteraction intr = {}: material == nullptr until a closest-hit program fills it
		bool intrOpaque = false;
		V3 ip = mk3(0, 0, 0), in_ = ip;
		int imesh = -1;
		while (!(fabsf(rd.x) <= 2 * kRayEps && fabsf(rd.y) <= 2 * kRayEps && fabsf(rd.z) <= 2 * kRayEps)) {
			Traverser<false, MODE == kTraceMotion, MODE == kTraceFlat> tr;
			LocalStack<false> ls;
			tr.begin(wf.bvh, ro, rd, tMax, __int_as_float(aux.y));
			tr.runToEnd(wf.bvh, wf.scene.instances, sm, ls, [&](int inst, int prim, float u, float v) {
				if (!(wf.instFlags[inst] & 2)) return true;
				return !alphaKilled(wf, inst, prim, u, v, ro, rd);
			});
			if (tr.overflow) atomicExch(&wf.errorFlags[0], 1);
			const bool visible = tr.best.inst < 0;
			if (!visible) {
				SurfaceGeom g;
				rebuildGeometry(wf, make_int4(tr.best.inst, tr.best.prim, __float_as_int(tr.best.u), __float_as_int(tr.best.v)), mk3(d4), __int_as_float(aux.y), g);
				ip = g.p, in_ = g.n, imesh = g.mesh, intrOpaque = g.material >= 0;
			}
			if (!visible && intrOpaque) { T_ray = sp(0); break; }
			if (medium >= 0) {
				float tEnd = visible ? tMax : length(ip - ro) / length(rd);
				usedRng	   = true;
				Spec T_maj = sampleT_maj(wf.scene.media[medium], wf.scene, ro, rd, tEnd, rng, wl, [&](V3, const MediumPoint &mp, Spec sigma_maj, Spec Tm) -> bool {
					Spec sigma_n = cwiseMax(sigma_maj - mp.sigma_a - mp.sigma_s, 0.f);
					float pr	 = Tm.x * sigma_maj.x;
					T_ray *= Tm * sigma_n / pr;
					pl *= Tm * sigma_maj / pr;
					pu *= Tm * sigma_n / pr;
					Spec Tr = T_ray / mean(pu + pl);
					if (maxCoeff(Tr) < 0.05f) {
						if (rng.get1D() < 0.75f) T_ray = sp(0);
						else T_ray = T_ray / 0.25f;
					}
					return any(T_ray);
				});
				T_ray *= T_maj / T_maj.x;
				pu *= T_maj / T_maj.x;
				pl *= T_maj / T_maj.x;
			}
			if (visible || !any(T_ray)) break;
			// ray = intr.spawnRayTo(pLight)
			V3 off = in_ * kRayEps;
			if (dot(in_, pLight - ip) < 0.f) off = -off;
			ro	   = ip + off;
			rd	   = pLight - ro;
			medium = mediumAcross(wf.scene.meshes[imesh], in_, rd, aux.x);
		}
		if (usedRng) wf.px.rng[pix] = rng.state;
		if (any(T_ray)) {
			Spec Ld = ldcs4(wf.shadow.contrib + i), spu = ldcs4(wf.shadow.pu + i), spl = ldcs4(wf.shadow.pl + i);
			wf.px.L[pix] = Ld * T_ray / mean(spu * pu + spl * pl) + wf.px.L[pix];
		}
	}
}

// 64^3 max-density grid of a dense density grid (initializeMajorantGrid, media.cpp:18-75)
__global__ void k_build_majorant(const float *__restrict__ density, int rx, int ry, int rz, float3 bmin, float3 bmax, float *majorant) {
	int index = blockIdx.x * blockDim.x + threadIdx.x;
	if (index >= kMajRes * kMajRes * kMajRes) return;
	int c[3] = {index % kMajRes, (index / kMajRes) % kMajRes, index / (kMajRes * kMajRes)}, res[3] = {rx, ry, rz}, n0[3], n1[3];
	const float lo[3] = {bmin.x, bmin.y, bmin.z}, hi[3] = {bmax.x, bmax.y, bmax.z};
	for (int k = 0; k < 3; k++) {
		// medium-space bounds of the cell -> index space; individually rounded so that the integer ranges
		// (and with them the majorants, hence the number of tracking steps) equal the CPU oracle's
		float ext = xsub(hi[k], lo[k]);
		float w0 = xadd(lo[k], xmul(ext, xdiv((float) c[k], (float) kMajRes))), w1 = xadd(lo[k], xmul(ext, xdiv((float) (c[k] + 1), (float) kMajRes)));
		float i0 = xmul(xdiv(xsub(w0, lo[k]), ext), (float) res[k]), i1 = xmul(xdiv(xsub(w1, lo[k]), ext), (float) res[k]);
		n0[k] = max(int(xsub(i0, 1.f)), 0), n1[k] = min(int(xadd(i1, 1.f)), res[k] - 1);
	}
	float mx = 0;
	for (int z = n0[2]; z <= n1[2]; z++)
		for (int y = n0[1]; y <= n1[1]; y++)
			for (int x = n0[0]; x <= n1[0]; x++) mx = fmaxf(mx, density[x + (size_t) rx * (y + (size_t) ry * z)]);
	majorant[index] = mx;
}

// per-sample resolve (integrator.cpp:257-260).  Note the reference does NOT reset L between the
// samples of one frame, so sample k adds the running sum; kept as is.
__global__ void k_resolve(const __grid_constant__ Wavefront wf) {
	KRR_PDL_ENTRY();
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < wf.p.pixelCount; i += gridDim.x * blockDim.x) {
		Wavelengths wl = expandWavelengths(wf.px.lambda[i]);
		float rgb[3];
		toRGB(wf.px.L[i], wl, wf.scene.cs, rgb);
		float4 p = wf.px.pixel[i];
		wf.px.pixel[i] = make_float4(p.x + rgb[0], p.y + rgb[1], p.z + rgb[2], 0.f);
	}
}

// film write (integrator.cpp:262-266): /spp, optional clamp, alpha 1, row H-1-y (cuda.h:33-36).  With a frame batch:
// the mean of the F frames' films, summed in frame order (what averaging F separately rendered films gives)
__global__ void k_film(const __grid_constant__ Wavefront wf, float4 *film, int zeroOutside) {
	KRR_PDL_ENTRY();
	const int N = wf.p.width * wf.p.height;
	for (int pixelId = blockIdx.x * blockDim.x + threadIdx.x; pixelId < N; pixelId += gridDim.x * blockDim.x) {
		int i = pixelId - wf.p.pixelBegin;
		int x = pixelId % wf.p.width, y = pixelId / wf.p.width;
		float4 *dst = film + (size_t) (wf.p.height - 1 - y) * wf.p.width + x;
		if (i < 0 || i >= wf.p.partPixels) { // outside this handle's partition
			if (zeroOutside) *dst = make_float4(0, 0, 0, 0);
			continue;
		}
		if (wf.p.rowStride > 1) { // rows of the partition that belong to another band are that band's to write
			int row = i / wf.p.width;
			if (row % wf.p.rowStride != wf.p.rowPhase) continue;
			i = (row / wf.p.rowStride) * wf.p.width + x;
		}
		const float spp = (float) wf.p.spp;
		float r = 0, g = 0, b = 0;
		for (int l = 0; l < wf.p.layers; l++) {
			const float4 p = wf.px.pixel[(size_t) l * wf.p.layerPixels + i];
			float fr = p.x / spp, fg = p.y / spp, fb = p.z / spp;
			if (wf.p.enableClamp) {
				fr = fminf(fmaxf(fr, 0.f), wf.p.clampMax), fg = fminf(fmaxf(fg, 0.f), wf.p.clampMax), fb = fminf(fmaxf(fb, 0.f), wf.p.clampMax);
			}
			r = l ? r + fr : fr, g = l ? g + fg : fg, b = l ? b + fb : fb;
		}
		if (wf.p.layers > 1) {
			const float F = (float) wf.p.layers;
			r = __fdiv_rn(r, F), g = __fdiv_rn(g, F), b = __fdiv_rn(b, F); // correctly rounded: the mean a host computes from F films
		}
		*dst = make_float4(r, g, b, 1.f);
	}
}

// end-of-sample bookkeeping: fold the per-depth counters into 64-bit totals and clear them
struct StatTotals {
	unsigned long long camera, closest, shadow, scatter, hitLight, miss, mediumSample, mediumScatter;
	unsigned long long closestByDepth[64], shadowByDepth[64];
};
__global__ void k_fold_counters(DepthCounters *c, StatTotals *t, int nDepth, int cameraRays) {
	KRR_PDL_ENTRY();
	int d = threadIdx.x;
	if (d == 0) atomicAdd(&t->camera, (unsigned long long) cameraRays);
	if (d >= nDepth) return;
	DepthCounters dc = c[d];
	int sc = 0;
	for (int k = 0; k < MAT_COUNT; k++) sc += dc.nScatter[k];
	sc += dc.nScatterKilled;
	if (d == nDepth - 1) dc.nRay = 0; // the queue behind the last traced depth (null-interface re-pushes) is never traced
	atomicAdd(&t->closest, (unsigned long long) dc.nRay);
	atomicAdd(&t->shadow, (unsigned long long) dc.nShadow);
	atomicAdd(&t->scatter, (unsigned long long) sc);
	atomicAdd(&t->hitLight, (unsigned long long) dc.nHitLight);
	atomicAdd(&t->miss, (unsigned long long) dc.nMiss);
	atomicAdd(&t->mediumSample, (unsigned long long) dc.nMediumSample);
	atomicAdd(&t->mediumScatter, (unsigned long long) dc.nMediumScatter);
	if (d < 64) { t->closestByDepth[d] += dc.nRay; t->shadowByDepth[d] += dc.nShadow; }
	DepthCounters z = {};
	c[d] = z;
}

// debug tap: integer fields of the queues at one (sample, depth)
__global__ void k_capture(const __grid_constant__ Wavefront wf, int depth, int queue, int4 *out, int32_t *outCount) {
	const RayQueue q  = wf.rays[depth & 1];
	const RayQueue nq = wf.rays[(depth & 1) ^ 1];
	DepthCounters *dc = wf.counters + depth;
	int n = 0;
	switch (queue) {
		case 0: n = dc->nRay; break;
		case 1: n = dc->nMiss; break;
		case 2: n = dc->nHitLight; break;
		case 3: for (int k = 0; k < MAT_COUNT; k++) n += dc->nScatter[k]; break;
		case 4: n = dc->nShadow; break;
		case 5: n = dc[1].nRay; break;
	}
	if (blockIdx.x == 0 && threadIdx.x == 0) *outCount = n;
	for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
		int4 r = make_int4(0, 0, 0, -1);
		auto fromRay = [&](const RayQueue &rq, int i) {
			int packed = __float_as_int(rq.ctxN_dep[i].w);
			r.x = framePixel(wf.p, __float_as_int(rq.ctxP_pix[i].w)), r.y = packed & 0xff, r.z = packed >> 8;
		};
		if (queue == 0) fromRay(q, k);
		else if (queue == 5) fromRay(nq, k);
		else if (queue == 1) fromRay(q, wf.missIdx[k]);
		else if (queue == 2) {
			int i = wf.hitLightIdx[k];
			fromRay(q, i);
			int4 hit = wf.hits[i];
			r.w = wf.scene.instances[hit.x].lightBase + hit.y;
		} else if (queue == 3) {
			int kk = k, mt = 0;
			while (kk >= dc->nScatter[mt]) kk -= dc->nScatter[mt], mt++;
			int i = wf.scatterIdx[mt][kk];
			fromRay(q, i);
			// ScatterRayWorkItem carries the prepared interaction: report ITS BSDF type (shared.h:46-73)
			SurfaceGeom g;
			rebuildGeometry(wf, wf.hits[i], mk3(q.d_medium[i]), q.o_time[i].w, g);
			Wavelengths wl = expandWavelengths(wf.px.lambda[__float_as_int(q.ctxP_pix[i].w)]);
			ShadingData sd;
			bool term;
			evalMaterial(wf, g, wl, sd, term);
			r.z = getBsdfType(sd);
			r.w = mt;
		} else if (queue == 4) {
			r.x = framePixel(wf.p, __float_as_int(wf.shadow.d_pix[k].w)), r.y = 0, r.z = 0;
		}
		out[k] = r;
	}
}

// AccumulatePass kernel (src/render/passes/accumulate/accumulate.cu:30-52)
__global__ void k_accumulate(float4 *accum, float4 *film, long long n, unsigned long long accumCount, unsigned long long maxAccum, int movingAverage) {
	for (long long i = blockIdx.x * (long long) blockDim.x + threadIdx.x; i < n; i += (long long) gridDim.x * blockDim.x) {
		float w	 = 1.f / (accumCount + 1);
		float4 c = film[i], a;
		if (accumCount > 0) {
			a = accum[i];
			if (movingAverage) a = make_float4(a.x * (1 - w) + c.x * w, a.y * (1 - w) + c.y * w, a.z * (1 - w) + c.z * w, a.w * (1 - w) + c.w * w);
			else if (!maxAccum || accumCount < maxAccum) a = make_float4(a.x + c.x, a.y + c.y, a.z + c.z, a.w + c.w);
		} else a = c;
		accum[i] = a;
		film[i]	 = movingAverage ? a : make_float4(a.x * w, a.y * w, a.z * w, a.w * w);
	}
}

} // namespace krr
