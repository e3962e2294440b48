// api.cu -- handle, scene upload, stage driver and the C ABI (include/krr_wfpt.h).
// Host code is C++17 compiled by nvcc's host compiler; everything the caller sees is extern "C".
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h> // types and prototypes only: the library is dlopen-ed at the first collective call

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "../host/json.h"
#include "bvh_build.h"
#include "krr_wfpt.h"
#include "wavefront_kernels.cuh"
#include "leaf_debug.cuh"
#include "megakernel.cuh"

using namespace krr;

namespace {

thread_local char gErr[512] = "";
int fail(int code, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(gErr, sizeof gErr, fmt, ap);
	va_end(ap);
	return code;
}
#define CUDA_OK(x)                                                                                         \
	do {                                                                                                   \
		cudaError_t e_ = (x);                                                                              \
		if (e_ != cudaSuccess) return fail(KRR_E_CUDA, "%s failed: %s", #x, cudaGetErrorString(e_));       \
	} while (0)

template <typename T> struct Buf {
	T *p	 = nullptr;
	size_t n = 0;
	int alloc(size_t count) {
		if (count == n && p) return KRR_OK;
		release();
		n = count;
		if (cudaMalloc((void **) &p, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) {
			p = nullptr, n = 0;
			cudaGetLastError();
			return fail(KRR_E_CUDA, "cudaMalloc of %zu bytes failed", count * sizeof(T));
		}
		return KRR_OK;
	}
	int upload(const std::vector<T> &h) {
		int rc = alloc(h.size());
		if (rc) return rc;
		if (!h.empty()) CUDA_OK(cudaMemcpy(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
		return KRR_OK;
	}
	void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
	~Buf() { release(); }
};

// Affine inverse: cofactors in double, rounded once (the intersection spec; identical to
// oracle/driver.cpp xfInverse so object-space rays agree bit-for-bit)
Xf xfInverse(const Xf &t) {
	double a = t.m[0], b = t.m[1], c = t.m[2], d = t.m[4], e = t.m[5], f = t.m[6], g = t.m[8], h = t.m[9], i = t.m[10];
	double A = e * i - f * h, B = -(d * i - f * g), C = d * h - e * g;
	double det = a * A + b * B + c * C;
	double id  = 1.0 / det;
	double r[9] = {A * id, -(b * i - c * h) * id, (b * f - c * e) * id, B * id, (a * i - c * g) * id, -(a * f - c * d) * id,
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

} // namespace

// Queues, pixel state and counters of one BAND of the frame.  A frame is rendered as `bands` interleaved
// row sets (band b owns rows rowBegin + b, rowBegin + b + bands, ...), each with its own queues on its own
// stream: pixels are independent (private RNG stream and accumulator), so the film is
// bit-identical, and the launches of one band fill the SMs that the other band's launch leaves idle
// while it drains (the persistent CTAs of a stage exit one by one) or walks a short deep-bounce queue
// (a launch costs 20-35 us however few rays it has).  The parity taps, debug captures and the per-stage
// event timing run the frame as ONE band (band 0 is always allocated for the whole partition).
constexpr int kMaxBands = 4;
constexpr int kSortAuto = 0; // automatic "sort_rays": off (measured: no gain on either tree workload)
constexpr int kTailAutoFlat = 0, kTailAutoTree = 0; // automatic "tail_depth" per scene kind (0 = off until measured)
struct WaveState {
	Buf<float4> L, pixel, rayBuf[2][7], shadowBuf[5];
	Buf<uint64_t> rng;
	Buf<float> lambda, cameraSample;
	Buf<int4> hits, firstHits;
	Buf<int32_t> missIdx, hitLightIdx, scatterIdx[MAT_COUNT], errorFlags, mediumSampleIdx;
	Buf<float> hitT;
	Buf<float4> msBuf[4];
	Buf<int2> msPixDepth, shadowAux;
	Buf<DepthCounters> counters;
	Buf<StatTotals> totals;
	Buf<int32_t> permClosest, permShadow; // "sort_rays": traversal order of the next-depth ray queue / the shadow queue (k_sort_rays)
	Buf<uint32_t> sortKeys[2];			  // global variant: keys in / out, slot numbers in, cub's scratch
	Buf<int32_t> sortVals;
	Buf<uint8_t> sortTemp;
};

struct KrrWfpt : WaveState {
	int device = 0;
	// params (integrator.h:78-88)
	int spp = 1, maxDepth = 10;
	float probRR = 0.8f, clampMax = 1e3f;
	bool nee = true, enableMedium = true, enableClamp = false;
	bool rrInTrace = true; // "rr_in_trace": internal scheduling switch (not a reference parameter), see Params::rrInTrace
	int flatBlasMax = 48;	 // "flat_blas_max": a BLAS with at most this many triangles is a flat list (takes effect at set_scene)
	bool pdl = false;		 // "pdl": programmatic dependent launch between the stage kernels (measured: -1.6 % on the bench workload, so off)
	bool usePdl() const { return pdl && !profile; }
	bool implicitDepth0 = true; // "implicit_depth0": depth-0 ray items store origin + direction only
	// "tail_depth": from this loop depth on the paths are finished by ONE launch of the tail kernel (k_tail) instead of
	// 2 launches per depth; 0 = never, -1 = automatic (kTailAuto*).  Fused schedule only.
	int tailDepth = -1;
	// "frame_batch": F >= 1 frames (frame indices frameIndex .. frameIndex + F - 1) rendered by the same launches;
	// the film is their mean (Params::layers).  Takes effect at the next resize / set_scene / set_partition.
	int frameBatch = 1;
	int layers() const { return debugState ? 1 : frameBatch; } // the parity taps address the pixel state of ONE frame
	bool fuseStages = true;	 // "fuse_stages": 2 launches per depth (hit/miss in the scatter launch, shadow + next closest in one trace launch)
	bool flattenStatic = true; // "flatten_static": static single-use instances join the merged BLAS whatever their transform (takes effect at set_scene)
	bool mergeStatic = true; // "merge_static": identity-transform static instances share one world-space BLAS (takes effect at set_scene)
	int width = 0, height = 0, rowBegin = 0, rowEnd = 0;
	bool haveScene = false, haveColorSpace = false, frameBegun = false;
	uint64_t frameIndex = 0;
	int numSMs = 148;

	// colour space
	Buf<float> csData;
	ColorSpaceDev cs{};
	std::vector<float> csHost; // host copy of zNodes + coeffs for upload-time conversions
	// scene
	Buf<float> positions, normals, texcoords, tangents, densityPool, spectrumTables, motionKeys;
	Buf<int32_t> indices, infiniteLights;
	Buf<MeshRec> meshes;
	Buf<InstRec> instances;
	Buf<MatRec> materials;
	Buf<LightRec> lights;
	Buf<TriLightRec> triLights;
	Buf<AnalyticLightRec> analytic;
	Buf<MediumRec> media;
	Buf<float4> texels;
	Buf<XformNodeRec> xnodes;
	Buf<float4> motionFlat; // per instance: flat motion record of short two-key chains (motion.cuh)
	float motionW0 = 0.f, motionW1 = 0.f; // ray-time window the TLAS boxes of moving instances currently cover
	Buf<uint8_t> instFlags;
	std::vector<InstRec> hInstances;
	std::vector<MeshRec> hMeshes;
	std::vector<uint8_t> mergedInst, dynamicInst; // per instance: lives in the merged BLAS / has been moved by update_instances
	bool anyMotion = false;
	SceneDev scene{};
	BvhBuilder bvh;
	bool matTypePresent[MAT_COUNT] = {false, false, false, false, false};
	bool sceneHasMedia = false;
	// wavefront state of band 0 (the whole partition when the frame runs as one band): WaveState base
	WaveState extra[kMaxBands - 1]; // bands 1..: allocated the first time a frame runs with more than one band
	cudaStream_t bandStream[kMaxBands - 1] = {};
	cudaEvent_t evFork = nullptr, evJoin[kMaxBands - 1] = {};
	int bands = 0;		 // "bands": see WaveState; 0 = automatic (2 for a scene that is one flat triangle list, else 1)
	int activeBands = 1; // decided by begin_frame
	int allocLayers = 1; // frames in flight the wavefront state was allocated for
	// pipelined film read-back (krr_wfpt_render_to_host_async): two device films, a copy stream
	Buf<float4> asyncFilm[2], syncFilm;
	cudaStream_t copyStream = nullptr;
	cudaEvent_t evRendered[2] = {}, evCopied[2] = {};
	int asyncSlot = 0;
	// multi-GPU: this handle's rank in the film-reduction communicator (krr_wfpt_comm_init_rank / _init_all)
	ncclComm_t comm = nullptr;
	int commRank = 0, commWorld = 1;
	// "sort_rays" (experiment, off by default): the rays of depth >= 1 are traced in (direction octant, origin Morton code)
	// order -- bit 0 = closest rays, bit 1 = shadow rays, bit 2 = device-wide sort over the queue capacity
	// (cub::DeviceRadixSort) instead of a sort inside tiles of 4096 queue slots (k_sort_rays); -1 = automatic (= off).
	// "sort_key": 0 = octant-major, 1 = origin-major.  Films are bit-identical either way.  Measured on the 20 M-triangle
	// and 10 000-instance workloads: the tile sort changes nothing (+-0.3 %), the global sort costs its own time (-6 % /
	// -18 %): the lanes of a trace warp diverge by traversal PHASE (node / triangle / instance entry), not by where
	// their rays are.  Takes effect at the next resize / set_scene / set_partition.
	int sortRays = -1, sortKey = 0;
	// "l2_persist_mb" (experiment, off): megabytes of the BVH node pool (from its start: TLAS + top BLAS levels) kept as persisting
	// L2 lines during render() (cudaLaunchAttributeAccessPolicyWindow).  Measured: 20 M triangles 32 MB 1 215 vs 1 222 Mrays/s,
	// 10 000 instances 8 MB 695 vs 695: the queues already stream past L2 (ld.cs / st.cs), the nodes were not being evicted
	int l2PersistMb = 0;
	int refill = 0;		 // "refill": idle lanes of a trace warp that trigger finalisation + refill; 0 = automatic (kRefill / kRefillFlat)
	WaveState &band(int b) { return b == 0 ? *this : extra[b - 1]; }
	int bandRows(int b, int nb) const { return (rowEnd - rowBegin - b + nb - 1) / nb; }
	KrrCameraDev cam{};
	// "debug_taps": keep the camera samples and depth-0 hits of every pixel for the parity taps
	// (krr_wfpt_debug_first_hits / _pixel_state): 36 B per pixel per sample of extra stores.  Off by default;
	// the test binding switches it on.  Takes effect at the next resize / set_scene.
	bool debugState = false;
	// debug capture
	int capSample = -1, capDepth = -1;
	Buf<int4> capItems[6];
	Buf<int32_t> capCounts;
	uint64_t launches = 0;
	cudaStream_t lastStream = nullptr;
	// optional per-stage timing (CUDA events on the launching stream)
	bool profile = false;
	std::vector<cudaEvent_t> evPool;
	struct EvRec { int stage; cudaEvent_t a, b; };
	std::vector<EvRec> evRecs;
	size_t evUsed = 0;
	cudaEvent_t nextEvent() {
		if (evUsed == evPool.size()) { cudaEvent_t e; cudaEventCreate(&e); evPool.push_back(e); }
		return evPool[evUsed++];
	}

	int pixelCount() const { return (rowEnd - rowBegin) * width; }
	// launch grids (occupancy x SM count of THIS handle's device), one entry per kernel instantiation.  Kept in the
	// handle: handles of one process may sit on different devices and be driven by different host threads.
	std::unordered_map<const void *, int> gridCache;
};

namespace {

int parseParams(KrrWfpt *h, const char *text) {
	if (!text || !*text) return KRR_OK;
	try {
		Json j = Json::parse(text);
		if (!j.isObject()) return fail(KRR_E_INVALID, "params must be a JSON object");
		h->nee			= j.value("nee", h->nee);
		h->enableMedium = j.value("enable_medium", h->enableMedium);
		h->maxDepth		= j.value("max_depth", h->maxDepth);
		h->probRR		= j.value("rr", h->probRR);
		h->enableClamp	= j.value("enable_clamp", h->enableClamp);
		h->clampMax		= j.value("clamp_max", h->clampMax);
		h->spp			= j.value("spp", h->spp);
		h->rrInTrace	= j.value("rr_in_trace", h->rrInTrace);
		h->mergeStatic	= j.value("merge_static", h->mergeStatic);
		h->flattenStatic = j.value("flatten_static", h->flattenStatic);
		h->fuseStages	= j.value("fuse_stages", h->fuseStages);
		h->tailDepth	= j.value("tail_depth", h->tailDepth);
		h->frameBatch	= j.value("frame_batch", h->frameBatch);
		h->implicitDepth0 = j.value("implicit_depth0", h->implicitDepth0);
		h->debugState	= j.value("debug_taps", h->debugState);
		h->pdl			= j.value("pdl", h->pdl);
		h->flatBlasMax	= j.value("flat_blas_max", h->flatBlasMax);
		h->bands		= j.value("bands", h->bands);
		h->refill		= j.value("refill", h->refill);
		h->sortRays		= j.value("sort_rays", h->sortRays);
		h->sortKey		= j.value("sort_key", h->sortKey);
		h->l2PersistMb	= j.value("l2_persist_mb", h->l2PersistMb);
	} catch (const std::exception &e) { return fail(KRR_E_INVALID, "bad params JSON: %s", e.what()); }
	if (h->maxDepth < 0 || h->maxDepth > kMaxDepthSlots - 2) return fail(KRR_E_INVALID, "max_depth must be in [0, %d]", kMaxDepthSlots - 2);
	if (h->spp < 1) return fail(KRR_E_INVALID, "spp must be >= 1");
	if (h->bands < 0 || h->bands > kMaxBands) return fail(KRR_E_INVALID, "bands must be in [0, %d]", kMaxBands);
	if (h->frameBatch < 1 || h->frameBatch > 64) return fail(KRR_E_INVALID, "frame_batch must be in [1, 64]");
	if (!(h->probRR > 0.f && h->probRR <= 1.f)) return fail(KRR_E_INVALID, "rr must be in (0, 1]");
	return KRR_OK;
}

int allocBand(KrrWfpt *h, WaveState &w, size_t n) {
	int rc = 0;
	rc |= w.L.alloc(n) | w.pixel.alloc(n) | w.rng.alloc(n) | w.lambda.alloc(n) | w.hits.alloc(n);
	if (h->debugState) rc |= w.cameraSample.alloc(5 * n) | w.firstHits.alloc(n);
	for (int q = 0; q < 2; q++) for (int a = 0; a < 7; a++) rc |= w.rayBuf[q][a].alloc(n);
	for (int a = 0; a < 5; a++) rc |= w.shadowBuf[a].alloc(n);
	rc |= w.missIdx.alloc(n) | w.hitLightIdx.alloc(n);
	for (int m = 0; m < MAT_COUNT; m++) rc |= w.scatterIdx[m].alloc(n);
	if (h->sceneHasMedia || h->scene.hasMotion) rc |= w.shadowAux.alloc(n); // shadow rays carry (medium, time)
	if (h->sceneHasMedia) { // media queues (mediumSampleQueue / mediumScatterQueue, integrator.cpp:40-44)
		rc |= w.mediumSampleIdx.alloc(n) | w.hitT.alloc(n) | w.msPixDepth.alloc(n);
		for (int a = 0; a < 4; a++) rc |= w.msBuf[a].alloc(n);
	}
	if (h->sortRays > 0) rc |= w.permClosest.alloc(n) | w.permShadow.alloc(n);
	if (h->sortRays > 0 && (h->sortRays & 4)) {
		size_t bytes = 0;
		cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *) nullptr, (uint32_t *) nullptr, (const int32_t *) nullptr, (int32_t *) nullptr, (int) n, 0, kSortKeyBits + 1);
		rc |= w.sortKeys[0].alloc(n) | w.sortKeys[1].alloc(n) | w.sortVals.alloc(n) | w.sortTemp.alloc(bytes);
	}
	const bool fresh = !w.counters.p;
	rc |= w.counters.alloc(kMaxDepthSlots) | w.totals.alloc(1) | w.errorFlags.alloc(4 + 132);
	if (rc) return KRR_E_CUDA;
	if (fresh) {
		CUDA_OK(cudaMemset(w.counters.p, 0, sizeof(DepthCounters) * kMaxDepthSlots));
		CUDA_OK(cudaMemset(w.totals.p, 0, sizeof(StatTotals)));
		CUDA_OK(cudaMemset(w.errorFlags.p, 0, 4 * (4 + 132)));
	}
	return KRR_OK;
}

// band 0 always covers the whole partition (frames that run as one band, the megakernel)
int allocState(KrrWfpt *h) {
	h->activeBands = 1;
	h->counters.release(); // fresh counters for a new frame size / scene
	h->allocLayers = h->layers();
	return allocBand(h, *h, (size_t) h->pixelCount() * h->allocLayers);
}

// bands 1.. of a frame that runs as `nb` bands (no-op when they already have the right size)
int allocExtraBands(KrrWfpt *h, int nb) {
	for (int b = 1; b < nb; b++) {
		int rc = allocBand(h, h->band(b), (size_t) h->bandRows(b, nb) * h->width * h->allocLayers);
		if (rc) return rc;
		if (!h->bandStream[b - 1]) {
			CUDA_OK(cudaStreamCreateWithFlags(&h->bandStream[b - 1], cudaStreamNonBlocking));
			CUDA_OK(cudaEventCreateWithFlags(&h->evJoin[b - 1], cudaEventDisableTiming));
		}
	}
	if (!h->evFork) CUDA_OK(cudaEventCreateWithFlags(&h->evFork, cudaEventDisableTiming));
	return KRR_OK;
}

TexRec makeTex(const KrrTextureDesc &t, std::vector<float4> &texels) {
	TexRec r{};
	memcpy(r.value, t.value, 16);
	r.valid	 = t.valid;
	r.texOff = -1;
	if (t.valid && t.image && t.width > 0 && t.height > 0) {
		r.texOff = (int32_t) texels.size();
		r.width = t.width, r.height = t.height;
		for (int i = 0; i < t.width * t.height; i++) texels.push_back(make_float4(t.image[4 * i], t.image[4 * i + 1], t.image[4 * i + 2], t.image[4 * i + 3]));
	}
	return r;
}

SpectrumRec makeSpectrum(const KrrSpectrumDesc &s, std::vector<float> &tables) {
	SpectrumRec r{};
	r.kind = s.kind;
	memcpy(r.a, s.a, 12), memcpy(r.b, s.b, 12);
	if (s.kind == KRR_SPEC_TABULATED && s.n > 0) {
		r.tabOff = (int32_t) tables.size(), r.n = s.n;
		tables.insert(tables.end(), s.lambdas, s.lambdas + s.n);
		tables.insert(tables.end(), s.values, s.values + s.n);
	}
	return r;
}

Wavefront makeWavefront(KrrWfpt *h, int sampleId, int bandId = 0, int nb = 1) {
	Wavefront wf{};
	WaveState &w = h->band(bandId);
	wf.p.width = h->width, wf.p.height = h->height;
	wf.p.pixelBegin = h->rowBegin * h->width, wf.p.partPixels = h->pixelCount();
	wf.p.layers = h->allocLayers, wf.p.layerPixels = h->bandRows(bandId, nb) * h->width;
	wf.p.rowStride = nb, wf.p.rowPhase = bandId, wf.p.pixelCount = wf.p.layers * wf.p.layerPixels;
	wf.p.spp = h->spp, wf.p.maxDepth = h->maxDepth, wf.p.nee = h->nee;
	wf.p.enableMedium = h->enableMedium && h->sceneHasMedia; // integrator.cpp:200
	wf.p.enableClamp = h->enableClamp, wf.p.probRR = h->probRR, wf.p.clampMax = h->clampMax;
	wf.p.rrInTrace	 = h->rrInTrace && !wf.p.enableMedium; // with media the medium stage draws before the scatter stage
	uint32_t seedIndex = (uint32_t) (h->frameIndex * (uint64_t) h->spp);
	wf.p.rngInc		 = ((uint64_t) seedIndex << 1u) | 1u;
	wf.p.seedIndex	 = seedIndex;
	wf.p.sampleIndex = (uint32_t) sampleId;
	{
		const BvhDev bd = h->bvh.device();
		wf.p.refill = bd.mergedOnly && !bd.mergedXf && isFlatEntry((uint32_t) bd.mergedRoot) ? kRefillFlat : kRefill;
		if (h->refill >= 1 && h->refill <= 32) wf.p.refill = h->refill;
	}
	wf.p.implicitDepth0 = h->implicitDepth0 && !(h->enableMedium && h->sceneHasMedia) && h->capSample < 0;
	wf.cam	 = h->cam;
	wf.scene = h->scene;
	wf.bvh	 = h->bvh.device();
	wf.bvh.motionFlat = h->scene.motionFlat;
	wf.px.L = w.L.p, wf.px.pixel = w.pixel.p, wf.px.rng = w.rng.p, wf.px.lambda = w.lambda.p;
	wf.px.cameraSample = h->debugState ? w.cameraSample.p : nullptr;
	for (int q = 0; q < 2; q++) {
		RayQueue &r = wf.rays[q];
		r.o_time = w.rayBuf[q][0].p, r.d_medium = w.rayBuf[q][1].p, r.thp = w.rayBuf[q][2].p, r.pu = w.rayBuf[q][3].p;
		r.pl = w.rayBuf[q][4].p, r.ctxP_pix = w.rayBuf[q][5].p, r.ctxN_dep = w.rayBuf[q][6].p;
	}
	wf.hits = w.hits.p;
	wf.missIdx = w.missIdx.p, wf.hitLightIdx = w.hitLightIdx.p;
	for (int m = 0; m < MAT_COUNT; m++) wf.scatterIdx[m] = w.scatterIdx[m].p;
	wf.shadow.o_tmax = w.shadowBuf[0].p, wf.shadow.d_pix = w.shadowBuf[1].p, wf.shadow.contrib = w.shadowBuf[2].p;
	wf.shadow.pu = w.shadowBuf[3].p, wf.shadow.pl = w.shadowBuf[4].p, wf.shadow.aux = w.shadowAux.p;
	wf.hitT = w.hitT.p, wf.mediumSampleIdx = w.mediumSampleIdx.p;
	wf.mscatter.p_time = w.msBuf[0].p, wf.mscatter.wo_medium = w.msBuf[1].p, wf.mscatter.thp = w.msBuf[2].p, wf.mscatter.pu = w.msBuf[3].p;
	wf.mscatter.pix_depth = w.msPixDepth.p;
	wf.counters	  = w.counters.p;
	wf.firstHits  = h->debugState ? w.firstHits.p : nullptr;
	wf.errorFlags = w.errorFlags.p;
	wf.instFlags  = h->instFlags.p;
	wf.tripHist	  = w.errorFlags.p + 4;
	wf.p.sortKey  = h->sortKey;
	wf.permClosest = wf.permShadow = nullptr; // set per launch by the fused schedule (krr_wfpt_render)
	return wf;
}

KrrCameraDev makeCamera(const KrrCameraData *c) {
	KrrCameraDev cam{};
	memcpy(cam.filmSize, c->film_size, 8);
	cam.focalLength = c->focal_length, cam.focalDistance = c->focal_distance, cam.lensRadius = c->lens_radius;
	cam.aspectRatio = c->aspect_ratio, cam.shutterOpen = c->shutter_open, cam.shutterTime = c->shutter_time;
	memcpy(cam.transform.m, c->transform, 48);
	cam.medium = c->medium;
	// fov = atan2(filmSize[1] * 0.5f, focalLength); tan(fov)  (camera.h:41,46): per-frame constants,
	// evaluated once here with the host libm so that every pixel sees the oracle's exact value
	float fov  = atan2f(c->film_size[1] * 0.5f, c->focal_length);
	cam.tanFov = tanf(fov);
	return cam;
}

template <typename K> int gridFor(KrrWfpt *h, K kernel, int block) {
	auto it = h->gridCache.find((const void *) kernel);
	if (it != h->gridCache.end()) return it->second;
	int occ = 1;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, block, 0);
	if (occ < 1) occ = 1;
	const int grid = h->numSMs * occ; // a whole number of waves: every SM gets the same number of resident CTAs
	h->gridCache[(const void *) kernel] = grid;
	return grid;
}

} // namespace

// =================================================================================================
extern "C" const char *krr_wfpt_last_error(void) { return gErr; }
namespace krr {
void setLastError(const char *fmt, ...) { // for the other translation units of the library (post_passes.cu)
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(gErr, sizeof gErr, fmt, ap);
	va_end(ap);
}
} // namespace krr
extern "C" int krr_wfpt_abi_version(void) { return KRR_WFPT_ABI_VERSION; }

extern "C" int krr_wfpt_create(const char *params_json, KrrWfpt **out) {
	if (!out) return fail(KRR_E_INVALID, "out is null");
	*out = nullptr;
	KrrWfpt *h = new KrrWfpt();
	int rc = parseParams(h, params_json); // argument errors are reported before any CUDA call
	if (rc) { delete h; return rc; }
	int dev = 0;
	cudaError_t ce = cudaGetDevice(&dev);
	if (ce != cudaSuccess) { delete h; return fail(KRR_E_CUDA, "cudaGetDevice failed: %s (the pass has no CPU path)", cudaGetErrorString(ce)); }
	h->device = dev;
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, dev) == cudaSuccess) h->numSMs = prop.multiProcessorCount;
	*out = h;
	return KRR_OK;
}

extern "C" void krr_wfpt_destroy(KrrWfpt *h) {
	if (!h) return;
	cudaDeviceSynchronize();
	for (int b = 0; b < kMaxBands - 1; b++) {
		if (h->bandStream[b]) cudaStreamDestroy(h->bandStream[b]);
		if (h->evJoin[b]) cudaEventDestroy(h->evJoin[b]);
	}
	if (h->evFork) cudaEventDestroy(h->evFork);
	krr_wfpt_comm_destroy(h);
	if (h->copyStream) cudaStreamDestroy(h->copyStream);
	for (int i = 0; i < 2; i++) {
		if (h->evRendered[i]) cudaEventDestroy(h->evRendered[i]);
		if (h->evCopied[i]) cudaEventDestroy(h->evCopied[i]);
	}
	delete h;
}

extern "C" int krr_wfpt_set_params(KrrWfpt *h, const char *params_json) {
	if (!h) return fail(KRR_E_INVALID, "null handle");
	return parseParams(h, params_json);
}

extern "C" int krr_wfpt_set_color_space(KrrWfpt *h, const KrrColorSpaceData *c) {
	if (!h || !c || !c->cie_x || !c->cie_y || !c->cie_z || !c->illuminant || !c->z_nodes || !c->coeffs)
		return fail(KRR_E_INVALID, "null colour-space data");
	const size_t nCo = (size_t) 3 * 64 * 64 * 64 * 3;
	std::vector<float> blob(4 * 471 + 64 + nCo);
	memcpy(&blob[0], c->cie_x, 471 * 4), memcpy(&blob[471], c->cie_y, 471 * 4), memcpy(&blob[942], c->cie_z, 471 * 4);
	memcpy(&blob[1413], c->illuminant, 471 * 4);
	memcpy(&blob[1884], c->z_nodes, 64 * 4), memcpy(&blob[1948], c->coeffs, nCo * 4);
	int rc = h->csData.upload(blob);
	if (rc) return rc;
	h->cs.cieX = h->csData.p, h->cs.cieY = h->csData.p + 471, h->cs.cieZ = h->csData.p + 942, h->cs.illum = h->csData.p + 1413;
	h->cs.zNodes = h->csData.p + 1884, h->cs.coeffs = h->csData.p + 1948;
	memcpy(h->cs.rgbFromXyz, c->rgb_from_xyz, 36);
	h->csHost.assign(blob.begin() + 1884, blob.end());
	h->haveColorSpace = true;
	h->scene.cs		  = h->cs;
	return KRR_OK;
}

namespace {
bool isIdentity(const Xf &t) {
	static const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
	for (int k = 0; k < 12; k++) if (!(t.m[k] == I[k])) return false;
	return true;
}

// (Re)builds BLASes + TLAS from the host copies of the mesh / instance records and uploads both.
// Static instances whose transform is exactly the identity are merged into one world-space BLAS
// (bvh_build.h); instances that krr_wfpt_update_instances has moved are kept out of it (dynamicInst).
int buildAccel(KrrWfpt *h) {
	const int nInst = (int) h->hInstances.size(), nMesh = (int) h->hMeshes.size();
	std::vector<uint8_t> merge(nInst, 0);
	bool any = false;
	if (h->mergeStatic) {
		// static instances with an identity transform, and (flatten_static) static instances that are the only
		// user of their mesh, whatever their transform (bvh.cuh BvhDev::mergedXf)
		std::vector<int> users(nMesh, 0);
		for (int i = 0; i < nInst; i++) users[h->hInstances[i].mesh]++;
		for (int i = 0; i < nInst; i++) {
			const InstRec &r = h->hInstances[i];
			const bool stat = r.motion < 0 && !h->dynamicInst[i];
			merge[i] = stat && ((isIdentity(r.xf) && isIdentity(r.inv)) || (h->flattenStatic && users[r.mesh] == 1));
			any |= merge[i] != 0;
		}
	}
	h->mergedInst = merge;
	std::vector<InstRec> up = h->hInstances;
	if (any) { // pseudo-instance of the merged BLAS
		InstRec r{};
		r.xf.m[0] = r.xf.m[5] = r.xf.m[10] = 1.f;
		r.inv = r.xf;
		r.mesh = nMesh, r.lightBase = -1, r.motion = -1, r.blasRoot = -1;
		up.push_back(r);
	}
	if (h->instances.upload(up)) return KRR_E_CUDA;
	MotionWindow mw;
	if (h->anyMotion) mw.xnodes = h->xnodes.p, mw.keys = h->motionKeys.p;
	mw.w0 = h->motionW0, mw.w1 = h->motionW1;
	char err[256] = "";
	if (!h->bvh.build(h->positions.p, h->indices.p, h->hMeshes.data(), nMesh, h->instances.p, h->hInstances.data(), nInst, any ? merge.data() : nullptr,
					  h->flatBlasMax, mw, nullptr, err))
		return fail(KRR_E_CUDA, "%s", err);
	for (int i = 0; i < nMesh; i++) h->hMeshes[i].blasRoot = h->bvh.blasRoot(i), h->hMeshes[i].triBase = h->bvh.triBase(i);
	for (int i = 0; i < nInst; i++) h->hInstances[i].blasRoot = up[i].blasRoot = h->hMeshes[h->hInstances[i].mesh].blasRoot;
	if (any) up[nInst].blasRoot = h->bvh.mergedRoot();
	if (h->instances.upload(up) | h->meshes.upload(h->hMeshes)) return KRR_E_CUDA;
	return KRR_OK;
}
} // namespace

extern "C" int krr_wfpt_set_scene(KrrWfpt *h, const KrrSceneDesc *d) {
	if (!h || !d) return fail(KRR_E_INVALID, "null argument");
	if (!h->haveColorSpace) return fail(KRR_E_STATE, "krr_wfpt_set_color_space must be called before set_scene");
	if (d->n_meshes <= 0 || d->n_instances <= 0) return fail(KRR_E_INVALID, "scene has no geometry");
	CUDA_OK(cudaSetDevice(h->device));
	const float *zn = h->csHost.data(), *co = h->csHost.data() + 64;
	std::vector<float> P, N, UV, T, density, specTables, motionKeys;
	std::vector<int32_t> I;
	std::vector<float4> texels;
	std::vector<MeshRec> meshes(d->n_meshes);
	for (int i = 0; i < d->n_meshes; i++) {
		const KrrMeshDesc &m = d->meshes[i];
		if (!m.positions || !m.indices || m.n_vertices <= 0 || m.n_triangles <= 0) return fail(KRR_E_INVALID, "mesh %d is empty", i);
		if (m.material >= d->n_materials) return fail(KRR_E_INVALID, "mesh %d: material index out of range", i);
		for (int k = 0; k < 3 * m.n_triangles; k++)
			if (m.indices[k] < 0 || m.indices[k] >= m.n_vertices) return fail(KRR_E_INVALID, "mesh %d: vertex index out of range", i);
		MeshRec r{};
		r.posOff = (int32_t) (P.size() / 3), r.idxOff = (int32_t) (I.size() / 3);
		r.nrmOff = m.normals ? (int32_t) (N.size() / 3) : -1;
		r.uvOff	 = m.texcoords ? (int32_t) (UV.size() / 2) : -1;
		r.tanOff = m.tangents ? (int32_t) (T.size() / 3) : -1;
		r.nTri = m.n_triangles, r.material = m.material, r.mediumIn = m.medium_inside, r.mediumOut = m.medium_outside;
		P.insert(P.end(), m.positions, m.positions + 3 * (size_t) m.n_vertices);
		if (m.normals) N.insert(N.end(), m.normals, m.normals + 3 * (size_t) m.n_vertices);
		if (m.texcoords) UV.insert(UV.end(), m.texcoords, m.texcoords + 2 * (size_t) m.n_vertices);
		if (m.tangents) T.insert(T.end(), m.tangents, m.tangents + 3 * (size_t) m.n_vertices);
		I.insert(I.end(), m.indices, m.indices + 3 * (size_t) m.n_triangles);
		meshes[i] = r;
	}
	// materials (MaterialData::getObjectData, texture.cpp:254-268)
	std::vector<MatRec> mats(std::max(d->n_materials, 0));
	for (int m = 0; m < MAT_COUNT; m++) h->matTypePresent[m] = false;
	for (int i = 0; i < d->n_materials; i++) {
		const KrrMaterialDesc &m = d->materials[i];
		if (m.color_space != 0) return fail(KRR_E_UNSUPPORTED, "material %d: only the sRGB colour space is supported", i);
		if (m.bsdf_type < 0 || m.bsdf_type >= MAT_COUNT) return fail(KRR_E_INVALID, "material %d: bad bsdf_type", i);
		MatRec r{};
		memcpy(r.diffuse, m.diffuse, 16), memcpy(r.specular, m.specular, 16);
		r.specularTransmission = m.specular_transmission, r.anisotropic = m.anisotropic, r.ior = m.ior;
		r.bsdfType = m.bsdf_type, r.shadingModel = m.shading_model;
		for (int t = 0; t < 5; t++) r.tex[t] = makeTex(m.textures[t], texels);
		r.eta = makeSpectrum(m.spectral_eta, specTables), r.k = makeSpectrum(m.spectral_k, specTables);
		// constant diffuse/specular -> final RGBs and their sigmoid coefficients at upload
		bool constD = !r.tex[0].valid || r.tex[0].texOff < 0, constS = !r.tex[1].valid || r.tex[1].texOff < 0;
		r.constColours = constD && constS;
		if (r.constColours) {
			const float *dv = r.tex[0].valid ? r.tex[0].value : r.diffuse, *sv = r.tex[1].valid ? r.tex[1].value : r.specular;
			memcpy(r.constDiffuse, dv, 12), memcpy(r.constSpecular, sv, 16);
			float dr[3], sr[3];
			if (r.shadingModel == KRR_SHADING_METALLIC_ROUGHNESS)
				for (int k = 0; k < 3; k++) dr[k] = dv[k] * (1 - sv[2]) + 0.f * sv[2], sr[k] = 0.f * (1 - sv[2]) + dv[k] * sv[2];
			else
				for (int k = 0; k < 3; k++) dr[k] = dv[k], sr[k] = sv[k];
			r.diffuseSpec  = makeBounded(zn, co, dr[0], dr[1], dr[2]);
			r.specularSpec = makeBounded(zn, co, sr[0], sr[1], sr[2]);
		}
		h->matTypePresent[r.bsdfType] = true;
		mats[i] = r;
	}
	// media
	std::vector<MediumRec> media(std::max(d->n_media, 0));
	for (int i = 0; i < d->n_media; i++) {
		const KrrMediumDesc &m = d->media[i];
		MediumRec r{};
		r.type = m.type;
		memcpy(r.sigma_t, m.sigma_t, 12), memcpy(r.albedo, m.albedo, 12), memcpy(r.Le, m.Le, 12);
		r.g = m.g;
		r.sigmaTSpec  = makeUnbounded(zn, co, m.sigma_t[0], m.sigma_t[1], m.sigma_t[2]);
		r.albedoUSpec = makeUnbounded(zn, co, m.albedo[0], m.albedo[1], m.albedo[2]);
		r.albedoBSpec = makeBounded(zn, co, m.albedo[0], m.albedo[1], m.albedo[2]);
		r.LeSpec	  = makeUnbounded(zn, co, m.Le[0], m.Le[1], m.Le[2]);
		memcpy(r.xf.m, m.transform, 48);
		r.inv = xfInverse(r.xf);
		memcpy(r.boundsMin, m.bounds_min, 12), memcpy(r.boundsMax, m.bounds_max, 12), memcpy(r.res, m.res, 12);
		r.scale = m.scale;
		r.densityOff = r.majorantOff = r.albedoOff = -1;
		if (m.type == KRR_MEDIUM_GRID) {
			if (!m.density || m.res[0] <= 0 || m.res[1] <= 0 || m.res[2] <= 0) return fail(KRR_E_INVALID, "medium %d: grid medium without density data", i);
			const size_t voxels = (size_t) m.res[0] * m.res[1] * m.res[2];
			r.densityOff = (int32_t) density.size();
			density.insert(density.end(), m.density, m.density + voxels);
			if (m.albedo_grid) { // RGB albedo per voxel on the density lattice (NanoVDBMedium::albedoGrid)
				r.albedoOff = (int32_t) density.size();
				density.insert(density.end(), m.albedo_grid, m.albedo_grid + 3 * voxels);
			}
		}
		media[i] = r;
	}
	for (MediumRec &r : media)
		if (r.type == KRR_MEDIUM_GRID) { // room for the 64^3 majorant grid, filled on the device below
			r.majorantOff = (int32_t) density.size();
			density.resize(density.size() + (size_t) kMajRes * kMajRes * kMajRes, 0.f);
		}
	h->sceneHasMedia = d->n_media > 0;
	// instances (InstanceData::getObjectData, mesh.cpp:16-63) + mesh lights (scene.cpp:91-112)
	std::vector<InstRec> insts(d->n_instances);
	std::vector<LightRec> lights;
	std::vector<TriLightRec> triLights;
	std::vector<uint8_t> flags(d->n_instances, 0);
	// transform chains (multi-level graph with SRT motion transforms, optix.cpp:400-563); read only when
	// motion blur is on, exactly like OptixSceneMultiLevel::getMotionKeyframes (optix.cpp:402)
	std::vector<XformNodeRec> xnodes;
	std::vector<char> chainMoves;
	const bool useMotion = d->options.motionblur != 0;
	const int nGraphNodes = useMotion && d->transform_nodes ? std::max(d->n_transform_nodes, 0) : 0;
	for (int i = 0; i < nGraphNodes; i++) {
		const KrrTransformNodeDesc &nd = d->transform_nodes[i];
		if (nd.parent < -1 || nd.parent >= nGraphNodes || nd.parent == i) return fail(KRR_E_INVALID, "transform node %d: bad parent", i);
		XformNodeRec r{};
		r.parent = nd.parent, r.keyOff = 0, r.nKeys = 0;
		memcpy(r.local.m, nd.transform, 48);
		r.localInv = xfInverse(r.local);
		if (nd.n_motion_keys >= 2) {
			if (!nd.motion_keys) return fail(KRR_E_INVALID, "transform node %d: motion keys missing", i);
			if (!(nd.time_end > nd.time_begin)) return fail(KRR_E_INVALID, "transform node %d: time_end must be > time_begin", i);
			r.keyOff = (int32_t) (motionKeys.size() / 10), r.nKeys = nd.n_motion_keys, r.t0 = nd.time_begin, r.t1 = nd.time_end;
			for (int k = 0; k < nd.n_motion_keys; k++) {
				const KrrSRT &s = nd.motion_keys[k];
				motionKeys.insert(motionKeys.end(), s.s, s.s + 3);
				motionKeys.insert(motionKeys.end(), s.q, s.q + 4);
				motionKeys.insert(motionKeys.end(), s.t, s.t + 3);
			}
		}
		xnodes.push_back(r);
	}
	chainMoves.assign(nGraphNodes, 0);
	for (int i = 0; i < nGraphNodes; i++) {
		int depth = 0;
		for (int p = i; p >= 0; p = xnodes[p].parent) {
			if (++depth > 64) return fail(KRR_E_INVALID, "transform node %d: chain deeper than 64 (cycle?)", i);
			if (xnodes[p].nKeys >= 2) chainMoves[i] = 1;
		}
	}
	bool anyMotion = false;
	for (int i = 0; i < d->n_instances; i++) {
		const KrrInstanceDesc &in = d->instances[i];
		if (in.mesh < 0 || in.mesh >= d->n_meshes) return fail(KRR_E_INVALID, "instance %d: mesh index out of range", i);
		InstRec r{};
		memcpy(r.xf.m, in.transform, 48);
		r.inv = xfInverse(r.xf);
		r.mesh = in.mesh, r.lightBase = -1, r.motion = -1;
		if (nGraphNodes > 0 && in.transform_node >= nGraphNodes) return fail(KRR_E_INVALID, "instance %d: transform_node out of range", i);
		if (nGraphNodes > 0 && in.transform_node >= 0) {
			if (chainMoves[in.transform_node]) r.motion = in.transform_node, anyMotion = true;
		} else if (useMotion && in.n_motion_keys >= 2 && in.motion_keys) {
			if (!(d->options.endtime > d->options.starttime)) return fail(KRR_E_INVALID, "motion blur needs options.endtime > options.starttime");
			r.motion = (int32_t) xnodes.size();
			XformNodeRec nr{};
			nr.parent = -1, nr.keyOff = (int32_t) (motionKeys.size() / 10), nr.nKeys = in.n_motion_keys;
			nr.t0 = d->options.starttime, nr.t1 = d->options.endtime;
			nr.local = r.xf, nr.localInv = r.inv;
			xnodes.push_back(nr);
			for (int k = 0; k < in.n_motion_keys; k++) {
				const KrrSRT &s = in.motion_keys[k];
				motionKeys.insert(motionKeys.end(), s.s, s.s + 3);
				motionKeys.insert(motionKeys.end(), s.q, s.q + 4);
				motionKeys.insert(motionKeys.end(), s.t, s.t + 3);
			}
			anyMotion = true;
		}
		const KrrMeshDesc &m = d->meshes[in.mesh];
		if (m.material < 0) flags[i] |= 1;
		else {
			if (d->materials[m.material].textures[KRR_TEX_TRANSMISSION].valid) flags[i] |= 2;
			flags[i] |= (uint8_t) (mats[m.material].bsdfType << 4); // routing info of the closest stage in one byte
		}
		bool emissiveTex = m.material >= 0 && d->materials[m.material].textures[KRR_TEX_EMISSIVE].valid;
		bool emissive	 = emissiveTex || m.Le[0] != 0 || m.Le[1] != 0 || m.Le[2] != 0;
		if (emissiveTex && d->materials[m.material].textures[KRR_TEX_EMISSIVE].image)
			// DiffuseAreaLight::L looks the emissive texture up at the hit's uv (light.h:186); this pass evaluates a constant
			// per emissive triangle, so an IMAGE-backed emissive texture would render with the wrong radiance: refuse it
			return fail(KRR_E_INVALID, "instance %d: image-backed emissive textures are not supported (constant emissive colours are)", i);
		if (emissive) {
			flags[i] |= 4;
			r.lightBase = (int32_t) lights.size();
			float Le[3];
			if (emissiveTex) memcpy(Le, d->materials[m.material].textures[KRR_TEX_EMISSIVE].value, 12);
			else memcpy(Le, m.Le, 12);
			float scale = std::max(Le[0], std::max(Le[1], Le[2]));
			// DiffuseAreaLight::L evaluates the emissive texture when it is valid (un-normalised
			// colour), otherwise the normalised Le; both are multiplied by scale (light.h:183-188)
			if (!emissiveTex) for (float &c : Le) c /= scale;
			RgbSpectrum LeSpec = makeUnbounded(zn, co, Le[0], Le[1], Le[2]);
			for (int t = 0; t < m.n_triangles; t++) {
				TriLightRec tl{};
				for (int c = 0; c < 3; c++) {
					int v = m.indices[3 * t + c];
					memcpy(tl.p[c], m.positions + 3 * v, 12);
					if (m.normals) memcpy(tl.n[c], m.normals + 3 * v, 12);
				}
				tl.inst = i, tl.scale = scale, tl.LeSpec = LeSpec, tl.twoSided = 0, tl.hasNormals = m.normals ? 1 : 0;
				memcpy(tl.Le, Le, 12);
				lights.push_back(LightRec{LIGHT_DIFFUSE_AREA, (int32_t) triLights.size()});
				triLights.push_back(tl);
			}
		}
		insts[i] = r;
	}
	// analytic lights, after the mesh lights (scene.cpp:114-132)
	std::vector<AnalyticLightRec> analytic;
	std::vector<int32_t> infinite;
	for (int i = 0; i < d->n_lights; i++) {
		const KrrLightDesc &l = d->lights[i];
		if (l.type == KRR_LIGHT_DIFFUSE_AREA || l.type < 0 || l.type > KRR_LIGHT_INFINITE) return fail(KRR_E_INVALID, "light %d: bad type", i);
		AnalyticLightRec r{};
		r.type = l.type;
		memcpy(r.color, l.color, 12);
		r.scale = l.scale;
		r.position[0] = l.transform[3], r.position[1] = l.transform[7], r.position[2] = l.transform[11];
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) r.rotation[a * 3 + b] = l.transform[a * 4 + b];
		r.sceneRadius = l.scene_radius;
		r.cosInner = std::cos(l.inner_cone_deg * 3.14159265358979323846f / 180.f);
		r.cosOuter = std::cos(l.outer_cone_deg * 3.14159265358979323846f / 180.f);
		memcpy(r.xf.m, l.transform, 48);
		r.inv = xfInverse(r.xf);
		r.image = makeTex(l.texture, texels);
		float tint[3] = {l.color[0], l.color[1], l.color[2]};
		if (l.type == KRR_LIGHT_INFINITE) {
			// uploaded through the texture constructor: tint = 1, colour comes from the (constant)
			// texture if there is one (light.cpp:28-41, light.h:213-216)
			if (r.image.valid && r.image.texOff < 0) { tint[0] = r.image.value[0], tint[1] = r.image.value[1], tint[2] = r.image.value[2]; r.image.valid = 0; }
			else if (!r.image.valid) tint[0] = tint[1] = tint[2] = 1.f;
			else { r.color[0] = r.color[1] = r.color[2] = 1.f; }
			infinite.push_back((int32_t) analytic.size());
		}
		r.colorSpec = makeUnbounded(zn, co, tint[0], tint[1], tint[2]);
		lights.push_back(LightRec{l.type, (int32_t) analytic.size()});
		analytic.push_back(r);
	}
	// flat motion records (motion.cuh): chains of at most two levels whose nodes are all two-key SRT nodes
	std::vector<float4> motionFlat;
	if (anyMotion) {
		motionFlat.assign((size_t) (d->n_instances + 1) * kMotionFlatStride, make_float4(0.f, 0.f, 0.f, 0.f));
		for (int i = 0; i < d->n_instances; i++) {
			const int n0 = insts[i].motion;
			if (n0 < 0) continue;
			const int n1 = xnodes[n0].parent;
			if (xnodes[n0].nKeys != 2 || (n1 >= 0 && (xnodes[n1].nKeys != 2 || xnodes[n1].parent >= 0))) continue;
			float4 *rec = motionFlat.data() + (size_t) i * kMotionFlatStride;
			const int levels = n1 >= 0 ? 2 : 1;
			rec[0] = make_float4(0.f, 0.f, 0.f, 0.f);
			memcpy(&rec[0].x, &levels, 4); // int bits
			rec[1] = make_float4(xnodes[n0].t0, xnodes[n0].t1, n1 >= 0 ? xnodes[n1].t0 : 0.f, n1 >= 0 ? xnodes[n1].t1 : 0.f);
			for (int l = 0; l < levels; l++) {
				const float *a = motionKeys.data() + 10 * (size_t) xnodes[l ? n1 : n0].keyOff, *b = a + 10;
				float v[20];
				for (int k = 0; k < 10; k++) v[k] = a[k], v[10 + k] = b[k] - a[k]; // one rounding: what xsub(b, a) gives on the device
				memcpy(rec + 2 + 5 * l, v, sizeof v);
			}
		}
	}
	int rc = 0;
	rc |= h->motionFlat.upload(motionFlat);
	rc |= h->positions.upload(P) | h->normals.upload(N) | h->texcoords.upload(UV) | h->tangents.upload(T) | h->indices.upload(I);
	rc |= h->densityPool.upload(density) | h->spectrumTables.upload(specTables) | h->motionKeys.upload(motionKeys) | h->xnodes.upload(xnodes);
	rc |= h->texels.upload(texels) | h->materials.upload(mats) | h->media.upload(media);
	rc |= h->lights.upload(lights) | h->triLights.upload(triLights) | h->analytic.upload(analytic) | h->infiniteLights.upload(infinite);
	rc |= h->instFlags.upload(flags);
	if (rc) return KRR_E_CUDA;
	// acceleration structures
	h->hInstances = insts, h->hMeshes = meshes;
	h->dynamicInst.assign(d->n_instances, 0);
	// until the first begin_frame supplies the camera's shutter interval, moving instances are bounded
	// over the whole animation range
	h->anyMotion = anyMotion;
	h->motionW0 = d->options.starttime, h->motionW1 = anyMotion ? std::max(d->options.endtime, d->options.starttime) : d->options.starttime;
	if (!anyMotion) h->motionW0 = h->motionW1 = 0.f;
	rc = buildAccel(h);
	if (rc) return rc;
	SceneDev &s = h->scene;
	s.positions = h->positions.p, s.normals = h->normals.p, s.texcoords = h->texcoords.p, s.tangents = h->tangents.p, s.indices = h->indices.p;
	s.meshes = h->meshes.p, s.instances = h->instances.p, s.materials = h->materials.p, s.lights = h->lights.p;
	s.triLights = h->triLights.p, s.analytic = h->analytic.p, s.infiniteLights = h->infiniteLights.p, s.media = h->media.p;
	s.densityPool = h->densityPool.p, s.texels = h->texels.p, s.spectrumTables = h->spectrumTables.p;
	s.motionKeys = h->motionKeys.p, s.xnodes = h->xnodes.p, s.motionFlat = anyMotion ? h->motionFlat.p : nullptr;
	s.nMeshes = d->n_meshes, s.nInstances = d->n_instances, s.nMaterials = d->n_materials, s.nLights = (int32_t) lights.size();
	s.nInfinite = (int32_t) infinite.size(), s.nMedia = d->n_media;
	s.motionStart = d->options.starttime, s.motionEnd = d->options.endtime, s.hasMotion = anyMotion;
	s.cs = h->cs;
	for (const MediumRec &r : media)
		if (r.type == KRR_MEDIUM_GRID)
			k_build_majorant<<<(kMajRes * kMajRes * kMajRes + 255) / 256, 256>>>(h->densityPool.p + r.densityOff, r.res[0], r.res[1], r.res[2],
				make_float3(r.boundsMin[0], r.boundsMin[1], r.boundsMin[2]), make_float3(r.boundsMax[0], r.boundsMax[1], r.boundsMax[2]),
				h->densityPool.p + r.majorantOff);
	CUDA_OK(cudaDeviceSynchronize());
	if (h->width > 0) { rc = allocState(h); if (rc) return rc; } // media queues depend on the scene
	h->haveScene = true;
	return KRR_OK;
}

extern "C" int krr_wfpt_resize(KrrWfpt *h, int32_t w, int32_t hgt) {
	if (!h || w <= 0 || hgt <= 0) return fail(KRR_E_INVALID, "bad size");
	if ((int64_t) w * hgt * 256 > 0x7fffffffLL) return fail(KRR_E_INVALID, "frame too large (256 * pixelId must fit in int32 as in the reference)");
	CUDA_OK(cudaSetDevice(h->device));
	CUDA_OK(cudaDeviceSynchronize()); // as the reference does before resizing queues (integrator.cpp:26)
	h->width = w, h->height = hgt;
	h->rowBegin = 0, h->rowEnd = hgt;
	h->frameBegun = false;
	return allocState(h);
}

extern "C" int krr_wfpt_set_partition(KrrWfpt *h, int32_t rb, int32_t re) {
	if (!h || h->width <= 0) return fail(KRR_E_STATE, "resize first");
	if (rb < 0 || re > h->height || rb >= re) return fail(KRR_E_INVALID, "bad row range");
	CUDA_OK(cudaDeviceSynchronize());
	h->rowBegin = rb, h->rowEnd = re;
	h->frameBegun = false;
	return allocState(h);
}

extern "C" int krr_wfpt_update_instances(KrrWfpt *h, const int32_t *ids, const float *xf, int32_t n, void *stream) {
	if (!h || !h->haveScene) return fail(KRR_E_STATE, "no scene");
	if (n <= 0) return KRR_OK;
	if (!ids || !xf) return fail(KRR_E_INVALID, "null argument");
	cudaStream_t st = (cudaStream_t) stream;
	bool rebuild = false;
	for (int i = 0; i < n; i++) {
		if (ids[i] < 0 || ids[i] >= (int) h->hInstances.size()) return fail(KRR_E_INVALID, "instance id out of range");
		InstRec &r = h->hInstances[ids[i]];
		memcpy(r.xf.m, xf + 12 * i, 48);
		r.inv = xfInverse(r.xf);
		h->dynamicInst[ids[i]] = 1;
		// an instance that was merged into the static world-space BLAS starts to move: take it out
		// (one rebuild; from then on it is an ordinary TLAS instance and updates are refits)
		if (h->mergedInst[ids[i]]) rebuild = true;
		// updateAccelStructure memcpy's the changed transforms only (optix.cpp:618-643)
		else CUDA_OK(cudaMemcpyAsync(h->instances.p + ids[i], &r, sizeof(InstRec), cudaMemcpyHostToDevice, st));
	}
	if (rebuild) {
		CUDA_OK(cudaStreamSynchronize(st));
		CUDA_OK(cudaDeviceSynchronize());
		int rc = buildAccel(h);
		if (rc) return rc;
		h->scene.instances = h->instances.p, h->scene.meshes = h->meshes.p;
		return KRR_OK;
	}
	char err[256] = "";
	if (!h->bvh.refitTlas(h->instances.p, st, err)) return fail(KRR_E_CUDA, "%s", err);
	h->launches += h->bvh.refitLaunches();
	return KRR_OK;
}

extern "C" int krr_wfpt_begin_frame(KrrWfpt *h, uint64_t frameIndex, const KrrCameraData *c, void *stream) {
	if (!h || !c) return fail(KRR_E_INVALID, "null argument");
	if (!h->haveScene) return fail(KRR_E_STATE, "set_scene first");
	if (h->width <= 0) return fail(KRR_E_STATE, "resize first");
	CUDA_OK(cudaSetDevice(h->device));
	cudaStream_t st = (cudaStream_t) stream;
	if (h->layers() != h->allocLayers) { // "frame_batch" / "debug_taps" changed through set_params
		CUDA_OK(cudaDeviceSynchronize());
		int rc = allocState(h);
		if (rc) return rc;
	}
	h->frameIndex = frameIndex;
	h->cam = makeCamera(c);
	h->launches = 0;
	if (h->scene.hasMotion) {
		// rays of this frame carry times in [shutterOpen, shutterOpen + shutterTime] (camera.h:44): re-fit
		// the TLAS boxes of the moving instances to that interval when it changed
		float w0 = c->shutter_open, w1 = c->shutter_open + c->shutter_time;
		if (w1 < w0) std::swap(w0, w1);
		if (w0 != h->motionW0 || w1 != h->motionW1) {
			MotionWindow mw;
			mw.xnodes = h->xnodes.p, mw.keys = h->motionKeys.p, mw.w0 = w0, mw.w1 = w1;
			char err[256] = "";
			if (!h->bvh.refitTlas(h->instances.p, st, err, &mw)) return fail(KRR_E_CUDA, "%s", err);
			h->launches += h->bvh.refitLaunches();
			h->motionW0 = w0, h->motionW1 = w1;
		}
	}
	// the frame runs as `bands` interleaved row sets unless something needs the frame's state in one piece
	// (parity taps, debug capture) or serial launches (per-stage event timing)
	int nb = h->bands;
	if (nb == 0) {
		// automatic: the short, latency-bound launches of a flat-list scene (Cornell box: +11 %, with the grid
		// medium +7 %) gain from a second band; the long traversal launches of a tree scene lose to it (20 M
		// triangles -5 %, 10 000 moving instances -11 %: two resident kernels share L1 and the stack space)
		const BvhDev bd = h->bvh.device();
		nb = bd.mergedOnly && !bd.mergedXf && isFlatEntry((uint32_t) bd.mergedRoot) ? 2 : 1;
	}
	nb = std::min(nb, h->rowEnd - h->rowBegin);
	if (h->debugState || h->capSample >= 0 || h->profile) nb = 1;
	if (nb > 1) { int rc = allocExtraBands(h, nb); if (rc) return rc; }
	h->activeBands = nb;
	uint32_t seedIndex = (uint32_t) (frameIndex * (uint64_t) h->spp);
	int grid = gridFor(h, k_begin_frame, 256);
	for (int b = 0; b < nb; b++) {
		CUDA_OK(cudaMemsetAsync(h->band(b).totals.p, 0, sizeof(StatTotals), st));
		Wavefront wf = makeWavefront(h, 0, b, nb);
		k_begin_frame<<<grid, 256, 0, st>>>(wf, seedIndex);
		h->launches++;
	}
	CUDA_OK(cudaGetLastError());
	h->frameBegun = true;
	h->lastStream = st;
	return KRR_OK;
}

namespace {
// launch with the programmatic-stream-serialization attribute (see KRR_PDL_ENTRY in wavefront_kernels.cuh)
// "l2_persist_mb": L2 access-policy window of the launches of the current render (the first bytes of the BVH node pool = the
// TLAS and the top levels of the BLASes, emitted breadth first): hits are kept as persisting lines, everything else streams
thread_local cudaAccessPolicyWindow gL2Window = {};
template <typename... KArgs, typename... Args>
void launchK(bool pdl, void (*kernel)(KArgs...), int grid, int block, cudaStream_t st, Args &&...args) {
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((unsigned) grid), cfg.blockDim = dim3((unsigned) block), cfg.dynamicSmemBytes = 0, cfg.stream = st;
	cudaLaunchAttribute at[2];
	int n = 0;
	if (pdl) {
		at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
		at[n].val.programmaticStreamSerializationAllowed = 1;
		n++;
	}
	if (gL2Window.num_bytes) {
		at[n].id = cudaLaunchAttributeAccessPolicyWindow;
		at[n].val.accessPolicyWindow = gL2Window;
		n++;
	}
	cfg.attrs = at, cfg.numAttrs = n;
	cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}
struct StageTimer { // RAII: brackets one launch with events when profiling is on
	KrrWfpt *h; cudaStream_t st; KrrWfpt::EvRec rec; bool on;
	StageTimer(KrrWfpt *h_, int stage, cudaStream_t st_) : h(h_), st(st_), on(h_->profile) {
		if (on) { rec.stage = stage; rec.a = h->nextEvent(); rec.b = h->nextEvent(); cudaEventRecord(rec.a, st); }
	}
	~StageTimer() { if (on) { cudaEventRecord(rec.b, st); h->evRecs.push_back(rec); } }
};
template <int MT> void launchScatter(KrrWfpt *h, const Wavefront &wf, int depth, cudaStream_t st, int &withHitMiss) {
	const int grid = gridFor(h, k_scatter<MT, false>, kScatterBlock), gridM = wf.scene.hasMotion ? gridFor(h, k_scatter<MT, true>, kScatterBlock) : 0;
	StageTimer t(h, KRR_STAGE_SCATTER, st);
	if (wf.scene.hasMotion) launchK(h->usePdl(), k_scatter<MT, true>, gridM, kScatterBlock, st, wf, depth, withHitMiss);
	else launchK(h->usePdl(), k_scatter<MT, false>, grid, kScatterBlock, st, wf, depth, withHitMiss);
	withHitMiss = 0; // only the first scatter launch of a depth carries the hit / miss prologue
	h->launches++;
}

int capture(KrrWfpt *h, const Wavefront &wf, int depth, int queue, cudaStream_t st) {
	if (h->capItems[queue].alloc((size_t) h->pixelCount())) return KRR_E_CUDA;
	k_capture<<<h->numSMs, 256, 0, st>>>(wf, depth, queue, h->capItems[queue].p, h->capCounts.p + queue);
	return KRR_OK;
}
} // namespace

extern "C" int krr_wfpt_render(KrrWfpt *h, float *film, void *stream) {
	if (!h || !film) return fail(KRR_E_INVALID, "null argument");
	if (!h->frameBegun) return fail(KRR_E_STATE, "begin_frame must precede render");
	CUDA_OK(cudaSetDevice(h->device));
	cudaStream_t st = (cudaStream_t) stream;
	const bool motion = h->scene.hasMotion != 0; // scenes with moving instances run the variants that evaluate SRT chains at the ray's time
	const bool media  = h->enableMedium && h->sceneHasMedia;
	// a scene that is one flat triangle list runs the kTraceFlat instantiations (branch-free triangle pairs)
	bool flatScene = false;
	{
		const BvhDev bd = h->bvh.device();
		flatScene = !motion && bd.mergedOnly && !bd.mergedXf && isFlatEntry((uint32_t) bd.mergedRoot);
	}
	const int gridCam = gridFor(h, k_generate_camera_rays, 256), gridResolve = gridFor(h, k_resolve, 256);
	const int gridTrace = gridFor(h, k_trace_closest<kTraceStatic>, 128), gridHit = gridFor(h, k_handle_hit_miss<false>, 128);
	const int gridShadow = gridFor(h, k_trace_shadow<kTraceStatic>, 128), gridFused = gridFor(h, k_trace_fused<kTraceStatic>, 128);
	const int gridTraceM = motion ? gridFor(h, k_trace_closest<kTraceMotion>, 128) : 0, gridHitM = motion ? gridFor(h, k_handle_hit_miss<true>, 128) : 0;
	const int gridShadowM = motion ? gridFor(h, k_trace_shadow<kTraceMotion>, 128) : 0, gridFusedM = motion ? gridFor(h, k_trace_fused<kTraceMotion>, 128) : 0;
	const int gridMSample = media ? gridFor(h, k_medium_sample, 128) : 0, gridMScatter = media ? gridFor(h, k_medium_scatter, 128) : 0;
	const int gridShadowTr = media ? gridFor(h, k_trace_shadow_tr<kTraceMotion>, kTraceBlock) : 0;
	const int gridTraceF = flatScene ? gridFor(h, k_trace_closest<kTraceFlat>, 128) : 0, gridFusedF = flatScene ? gridFor(h, k_trace_fused<kTraceFlat>, 128) : 0;
	const int gridSort = gridFor(h, k_sort_rays, kSortThreads);
	gL2Window = {};
	if (h->l2PersistMb > 0 && !flatScene) {
		const BvhDev bd = h->bvh.device();
		static thread_local int limitSetFor = -1; // device whose persisting-L2 carve-out has been sized
		static thread_local size_t maxWindow = 0, maxPersist = 0;
		if (limitSetFor != h->device) {
			cudaDeviceProp prop;
			if (cudaGetDeviceProperties(&prop, h->device) == cudaSuccess) maxWindow = (size_t) prop.accessPolicyMaxWindowSize, maxPersist = (size_t) prop.persistingL2CacheMaxSize;
			cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, std::min((size_t) h->l2PersistMb << 20, maxPersist));
			limitSetFor = h->device;
		}
		// (a window larger than the device's limits is an invalid launch attribute)
		const size_t want = std::min(std::min((size_t) h->l2PersistMb << 20, (size_t) h->bvh.nodeCount() * sizeof(Node8)), std::min(maxWindow, maxPersist));
		gL2Window.base_ptr = (void *) bd.nodes, gL2Window.num_bytes = want, gL2Window.hitRatio = 1.f;
		gL2Window.hitProp = cudaAccessPropertyPersisting, gL2Window.missProp = cudaAccessPropertyStreaming;
	}
	const int gridShadowTrF = flatScene && media ? gridFor(h, k_trace_shadow_tr<kTraceFlat>, kTraceBlock) : 0;
	if (h->capSample >= 0 && h->capCounts.alloc(8)) return KRR_E_CUDA;
	const bool pdl = h->usePdl();
	const int nDepthSlots = h->maxDepth + 2;
	// bands: band 0 runs on the caller's stream, bands 1.. on their own streams, forked from and joined to it
	const int nb = h->activeBands;
	const cudaStream_t callerStream = st;
	if (nb > 1) {
		if (h->capSample >= 0 || h->profile) return fail(KRR_E_STATE, "begin_frame must follow set_profiling / debug_capture");
		CUDA_OK(cudaEventRecord(h->evFork, callerStream));
		for (int b = 1; b < nb; b++) CUDA_OK(cudaStreamWaitEvent(h->bandStream[b - 1], h->evFork, 0));
	}
	for (int sampleId = 0; sampleId < h->spp; sampleId++)
	for (int bandId = 0; bandId < nb; bandId++) {
		st = bandId == 0 ? callerStream : h->bandStream[bandId - 1];
		Wavefront wf = makeWavefront(h, sampleId, bandId, nb);
		// [1] primary rays.  Queue counters were cleared by k_fold_counters of the previous sample
		{ StageTimer t(h, KRR_STAGE_CAMERA, st); launchK(pdl, k_generate_camera_rays, gridCam, 256, st, wf); }
		h->launches++;
		// Fused schedule (default; surface-only scenes with NEE): handleHit/Miss rides as the prologue of
		// the scatter launch of its depth, and ONE trace launch covers the shadow rays of depth d and the
		// closest rays of depth d + 1 -> 2 launches per depth instead of 4.  Debug captures and scenes with
		// participating media keep the reference's stage order below.
		const bool anyScatter = h->matTypePresent[MAT_DISNEY] || h->matTypePresent[MAT_DIFFUSE] || h->matTypePresent[MAT_DIELECTRIC] ||
								h->matTypePresent[MAT_CONDUCTOR] || h->matTypePresent[MAT_NULL];
		const bool fused = h->fuseStages && !media && h->nee && anyScatter && h->capSample != sampleId;
		auto launchAllScatter = [&](int depth, int withHitMiss) {
			if (h->matTypePresent[MAT_DISNEY]) launchScatter<MAT_DISNEY>(h, wf, depth, st, withHitMiss);
			if (h->matTypePresent[MAT_DIFFUSE]) launchScatter<MAT_DIFFUSE>(h, wf, depth, st, withHitMiss);
			if (h->matTypePresent[MAT_DIELECTRIC]) launchScatter<MAT_DIELECTRIC>(h, wf, depth, st, withHitMiss);
			if (h->matTypePresent[MAT_CONDUCTOR]) launchScatter<MAT_CONDUCTOR>(h, wf, depth, st, withHitMiss);
			if (h->matTypePresent[MAT_NULL]) launchScatter<MAT_NULL>(h, wf, depth, st, withHitMiss);
		};
		auto launchHitMiss = [&](int depth) {
			StageTimer t(h, KRR_STAGE_HIT_MISS, st);
			if (motion) launchK(pdl, k_handle_hit_miss<true>, gridHitM, 128, st, wf, depth);
			else launchK(pdl, k_handle_hit_miss<false>, gridHit, 128, st, wf, depth);
			h->launches++;
		};
		auto launchClosest = [&](int depth) {
			StageTimer t(h, KRR_STAGE_CLOSEST, st);
			if (motion) launchK(pdl, k_trace_closest<true>, gridTraceM, 128, st, wf, depth);
			else if (flatScene) launchK(pdl, k_trace_closest<kTraceFlat>, gridTraceF, 128, st, wf, depth);
			else launchK(pdl, k_trace_closest<false>, gridTrace, 128, st, wf, depth);
			h->launches++;
		};
		// ray reordering: tree scenes only (a flat triangle list is walked in lock step whatever the rays are)
		const int sortMask = flatScene || !h->band(bandId).permClosest.p ? 0 : (h->sortRays < 0 ? kSortAuto : h->sortRays & 3);
		if (fused) {
			launchClosest(0);
			// automatic tail depth: where the queues have shrunk to a few per cent of the frame (rr 0.8 per bounce and
			// the paths that left the scene); measured per scene kind, see DESIGN.md
			int tail = h->tailDepth < 0 ? (flatScene ? kTailAutoFlat : kTailAutoTree) : h->tailDepth;
			if (tail < 1 || tail > h->maxDepth) tail = 0;
			bool tailed = false;
			for (int depth = 0; depth < h->maxDepth && !tailed; depth++) {
				launchAllScatter(depth, 1);
				if (tail && depth + 1 == tail) {
					// shadow rays of this depth on their own, then the rest of every path in one launch
					wf.permClosest = wf.permShadow = nullptr; // (queue order: the permutations belong to the previous depth)
					{
						StageTimer t(h, KRR_STAGE_SHADOW, st);
						if (motion) launchK(pdl, k_trace_shadow<kTraceMotion>, gridShadowM, 128, st, wf, depth);
						else if (flatScene) launchK(pdl, k_trace_shadow<kTraceFlat>, gridFor(h, k_trace_shadow<kTraceFlat>, 128), 128, st, wf, depth);
						else launchK(pdl, k_trace_shadow<kTraceStatic>, gridShadow, 128, st, wf, depth);
					}
					{
						StageTimer t(h, KRR_STAGE_TAIL, st);
						if (motion) launchK(pdl, k_tail<kTraceMotion>, gridFor(h, k_tail<kTraceMotion>, kTraceBlock), kTraceBlock, st, wf, tail);
						else if (flatScene) launchK(pdl, k_tail<kTraceFlat>, gridFor(h, k_tail<kTraceFlat>, kTraceBlock), kTraceBlock, st, wf, tail);
						else launchK(pdl, k_tail<kTraceStatic>, gridFor(h, k_tail<kTraceStatic>, kTraceBlock), kTraceBlock, st, wf, tail);
					}
					h->launches += 2;
					tailed = true;
					break;
				}
				StageTimer t(h, KRR_STAGE_TRACE, st);
				if (sortMask) { // traversal order of the two queues this launch consumes
					WaveState &w = h->band(bandId);
					wf.permClosest = (sortMask & 1) ? w.permClosest.p : nullptr, wf.permShadow = (sortMask & 2) ? w.permShadow.p : nullptr;
					if (h->sortRays > 0 && (h->sortRays & 4) && w.sortTemp.p) { // device-wide sort over the queue capacity
						const int cap = (int) w.permClosest.n;
						for (int sh = 0; sh < 2; sh++) {
							if (!(sortMask & (1 << sh))) continue;
							launchK(pdl, k_ray_keys, h->numSMs * 8, 256, st, wf, depth, sh, w.sortKeys[0].p, w.sortVals.p, cap);
							size_t bytes = w.sortTemp.n;
							cub::DeviceRadixSort::SortPairs(w.sortTemp.p, bytes, (const uint32_t *) w.sortKeys[0].p, w.sortKeys[1].p, (const int32_t *) w.sortVals.p,
															sh ? w.permShadow.p : w.permClosest.p, cap, 0, kSortKeyBits + 1, st);
							h->launches += 2;
						}
					} else {
						launchK(pdl, k_sort_rays, gridSort, kSortThreads, st, wf, depth, (sortMask & 1) ? w.permClosest.p : (int32_t *) nullptr,
								(sortMask & 2) ? w.permShadow.p : (int32_t *) nullptr);
						h->launches++;
					}
				}
				if (motion) launchK(pdl, k_trace_fused<true>, gridFusedM, 128, st, wf, depth);
				else if (flatScene) launchK(pdl, k_trace_fused<kTraceFlat>, gridFusedF, 128, st, wf, depth);
				else launchK(pdl, k_trace_fused<false>, gridFused, 128, st, wf, depth);
				h->launches++;
			}
			if (!tailed) launchHitMiss(h->maxDepth);
		}
		for (int depth = 0; !fused; depth++) {
			const bool cap = h->capSample == sampleId && h->capDepth == depth;
			if (cap && capture(h, wf, depth, 0, st)) return KRR_E_CUDA;
			// [2.1] closest hits
			launchClosest(depth);
			// [2.2] medium interactions along the rays that travel inside a medium
			if (media) {
				{ StageTimer t(h, KRR_STAGE_MEDIUM, st); k_medium_sample<<<gridMSample, 128, 0, st>>>(wf, depth); }
				h->launches++;
			}
			if (cap) for (int q = 1; q <= 3; q++) if (capture(h, wf, depth, q, st)) return KRR_E_CUDA;
			// [2.3] emitted / environment radiance
			launchHitMiss(depth);
			if (depth == h->maxDepth) break;
			if (media) {
				{ StageTimer t(h, KRR_STAGE_MEDIUM, st); k_medium_scatter<<<gridMScatter, 128, 0, st>>>(wf, depth); }
				h->launches++;
			}
			// [2.4] BSDF sampling + NEE, one launch per material type present in the scene
			launchAllScatter(depth, 0);
			if (cap) for (int q = 4; q <= 5; q++) if (capture(h, wf, depth, q, st)) return KRR_E_CUDA;
			// [2.5] shadow rays
			if (h->nee) {
				StageTimer t(h, KRR_STAGE_SHADOW, st);
				if (media && flatScene) k_trace_shadow_tr<kTraceFlat><<<gridShadowTrF, kTraceBlock, 0, st>>>(wf, depth);
				else if (media) k_trace_shadow_tr<kTraceMotion><<<gridShadowTr, kTraceBlock, 0, st>>>(wf, depth);
				else if (motion) launchK(pdl, k_trace_shadow<true>, gridShadowM, 128, st, wf, depth);
				else launchK(pdl, k_trace_shadow<false>, gridShadow, 128, st, wf, depth);
				h->launches++;
			}
		}
		{ StageTimer t(h, KRR_STAGE_RESOLVE, st); launchK(pdl, k_resolve, gridResolve, 256, st, wf); }
		{ StageTimer t(h, KRR_STAGE_RESOLVE, st); launchK(pdl, k_fold_counters, 1, 128, st, h->band(bandId).counters.p, h->band(bandId).totals.p, nDepthSlots, wf.p.pixelCount); }
		h->launches += 2;
	}
	for (int bandId = 0; bandId < nb; bandId++) { // each band writes its rows of the film; band 0 also clears the rows outside the partition
		st = bandId == 0 ? callerStream : h->bandStream[bandId - 1];
		Wavefront wf = makeWavefront(h, 0, bandId, nb);
		{ StageTimer t(h, KRR_STAGE_RESOLVE, st); launchK(pdl, k_film, gridResolve, 256, st, wf, (float4 *) film, bandId == 0 ? 1 : 0); }
		h->launches++;
		if (bandId > 0) {
			CUDA_OK(cudaEventRecord(h->evJoin[bandId - 1], st));
			CUDA_OK(cudaStreamWaitEvent(callerStream, h->evJoin[bandId - 1], 0));
		}
	}
	st = callerStream;
	gL2Window = {};
	CUDA_OK(cudaGetLastError());
	h->lastStream = st;
	return KRR_OK;
}

// MegakernelPathTracer::render (src/render/megakernel/pathtracer.cpp): one launch, one lane per pixel
extern "C" int krr_wfpt_render_megakernel(KrrWfpt *h, uint64_t frameIndex, const KrrCameraData *c, float *film, void *stream) {
	if (!h || !c || !film) return fail(KRR_E_INVALID, "null argument");
	if (!h->haveScene) return fail(KRR_E_STATE, "set_scene first");
	if (h->width <= 0) return fail(KRR_E_STATE, "resize first");
	if (h->allocLayers != 1) return fail(KRR_E_STATE, "the megakernel estimator renders one frame at a time: set \"frame_batch\": 1");
	CUDA_OK(cudaSetDevice(h->device));
	cudaStream_t st = (cudaStream_t) stream;
	h->cam = makeCamera(c);
	if (h->scene.hasMotion) {
		float w0 = c->shutter_open, w1 = c->shutter_open + c->shutter_time;
		if (w1 < w0) std::swap(w0, w1);
		if (w0 != h->motionW0 || w1 != h->motionW1) {
			MotionWindow mw;
			mw.xnodes = h->xnodes.p, mw.keys = h->motionKeys.p, mw.w0 = w0, mw.w1 = w1;
			char err[256] = "";
			if (!h->bvh.refitTlas(h->instances.p, st, err, &mw)) return fail(KRR_E_CUDA, "%s", err);
			h->motionW0 = w0, h->motionW1 = w1;
		}
	}
	Wavefront wf = makeWavefront(h, 0);
	MegaParams mp{h->width, h->height, h->spp, h->maxDepth, h->nee ? 1 : 0, h->probRR, (uint32_t) frameIndex};
	if (h->scene.hasMotion) k_megakernel<true><<<gridFor(h, k_megakernel<true>, kTraceBlock), kTraceBlock, 0, st>>>(wf, mp, (float4 *) film);
	else k_megakernel<false><<<gridFor(h, k_megakernel<false>, kTraceBlock), kTraceBlock, 0, st>>>(wf, mp, (float4 *) film);
	CUDA_OK(cudaGetLastError());
	h->lastStream = st;
	h->launches	  = 1;
	return KRR_OK;
}

extern "C" int krr_wfpt_render_to_host(KrrWfpt *h, float *film_host, void *stream) {
	if (!h || !film_host) return fail(KRR_E_INVALID, "null argument");
	Buf<float4> &staging = h->syncFilm; // per handle (a handle belongs to one device)
	size_t n = (size_t) h->width * h->height;
	if (staging.alloc(n)) return KRR_E_CUDA;
	int rc = krr_wfpt_render(h, (float *) staging.p, stream);
	if (rc) return rc;
	CUDA_OK(cudaMemcpyAsync(film_host, staging.p, n * 16, cudaMemcpyDeviceToHost, (cudaStream_t) stream));
	CUDA_OK(cudaStreamSynchronize((cudaStream_t) stream));
	return KRR_OK;
}

// Pipelined variant: the film of this frame travels to the host on a copy stream while the caller's stream
// goes on with the next frame.  Two device films alternate; rendering into one waits (on the device) for
// the copy that last read it.
extern "C" int krr_wfpt_render_to_host_async(KrrWfpt *h, float *film_host, void *stream) {
	if (!h || !film_host) return fail(KRR_E_INVALID, "null argument");
	CUDA_OK(cudaSetDevice(h->device));
	const size_t n = (size_t) h->width * h->height;
	if (!h->copyStream) {
		CUDA_OK(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
		for (int i = 0; i < 2; i++) {
			CUDA_OK(cudaEventCreateWithFlags(&h->evRendered[i], cudaEventDisableTiming));
			CUDA_OK(cudaEventCreateWithFlags(&h->evCopied[i], cudaEventDisableTiming));
		}
	}
	const int slot = h->asyncSlot;
	h->asyncSlot ^= 1;
	if (h->asyncFilm[slot].n != n) {
		CUDA_OK(cudaStreamSynchronize(h->copyStream)); // a resize between frames: no copy may still read the old buffer
		if (h->asyncFilm[slot].alloc(n)) return KRR_E_CUDA;
	}
	cudaStream_t st = (cudaStream_t) stream;
	CUDA_OK(cudaStreamWaitEvent(st, h->evCopied[slot], 0)); // no-op until the slot has been copied once
	int rc = krr_wfpt_render(h, (float *) h->asyncFilm[slot].p, stream);
	if (rc) return rc;
	CUDA_OK(cudaEventRecord(h->evRendered[slot], st));
	CUDA_OK(cudaStreamWaitEvent(h->copyStream, h->evRendered[slot], 0));
	CUDA_OK(cudaMemcpyAsync(film_host, h->asyncFilm[slot].p, n * 16, cudaMemcpyDeviceToHost, h->copyStream));
	CUDA_OK(cudaEventRecord(h->evCopied[slot], h->copyStream));
	return KRR_OK;
}

// =================================================================================================
// Multi-GPU: the ONE exchange step of the path -- the films of the ranks (image tiles add up, spp
// slices average) are summed onto the root rank with ncclReduce over NVLink.  The reference has no counterpart
// (single device, src/core/device/context.cpp:37-40).  NCCL is resolved at run time (dlopen "libnccl.so.2": the
// copy a host process already carries, e.g. torch's, or the system one), so a single-GPU build has no NCCL
// dependency; a missing library is an error return, never a fallback.
namespace {
struct NcclApi {
	void *lib = nullptr;
	decltype(&ncclGetUniqueId) getUniqueId = nullptr;
	decltype(&ncclCommInitRank) commInitRank = nullptr;
	decltype(&ncclCommInitAll) commInitAll = nullptr;
	decltype(&ncclCommDestroy) commDestroy = nullptr;
	decltype(&ncclReduce) reduce = nullptr;
	decltype(&ncclGetErrorString) errorString = nullptr;
	bool ok = false;
};
NcclApi &nccl() {
	static NcclApi api = [] {
		NcclApi a;
		a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
		if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
		if (!a.lib) return a;
		a.getUniqueId  = (decltype(a.getUniqueId)) dlsym(a.lib, "ncclGetUniqueId");
		a.commInitRank = (decltype(a.commInitRank)) dlsym(a.lib, "ncclCommInitRank");
		a.commInitAll  = (decltype(a.commInitAll)) dlsym(a.lib, "ncclCommInitAll");
		a.commDestroy  = (decltype(a.commDestroy)) dlsym(a.lib, "ncclCommDestroy");
		a.reduce	   = (decltype(a.reduce)) dlsym(a.lib, "ncclReduce");
		a.errorString  = (decltype(a.errorString)) dlsym(a.lib, "ncclGetErrorString");
		a.ok = a.getUniqueId && a.commInitRank && a.commInitAll && a.commDestroy && a.reduce && a.errorString;
		return a;
	}();
	return api;
}
#define NCCL_OK(x)                                                                                         \
	do {                                                                                                   \
		ncclResult_t r_ = (x);                                                                             \
		if (r_ != ncclSuccess) return fail(KRR_E_CUDA, "%s failed: %s", #x, nccl().errorString(r_));       \
	} while (0)
int needNccl() { return nccl().ok ? KRR_OK : fail(KRR_E_UNSUPPORTED, "libnccl.so.2 not found (multi-GPU film reduction needs NCCL): %s", dlerror() ? dlerror() : "missing symbols"); }

__global__ void k_scale_film(float4 *film, size_t n, float s) {
	for (size_t i = blockIdx.x * (size_t) blockDim.x + threadIdx.x; i < n; i += (size_t) gridDim.x * blockDim.x) {
		float4 v = film[i];
		film[i]	 = make_float4(v.x * s, v.y * s, v.z * s, v.w * s);
	}
}
} // namespace

extern "C" int krr_wfpt_comm_unique_id(uint8_t *out128) {
	if (!out128) return fail(KRR_E_INVALID, "null argument");
	if (int rc = needNccl()) return rc;
	static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
	ncclUniqueId id;
	NCCL_OK(nccl().getUniqueId(&id));
	memcpy(out128, &id, 128);
	return KRR_OK;
}

extern "C" int krr_wfpt_comm_init_rank(KrrWfpt *h, const uint8_t *id128, int32_t world, int32_t rank) {
	if (!h || !id128 || world < 1 || rank < 0 || rank >= world) return fail(KRR_E_INVALID, "bad argument");
	if (int rc = needNccl()) return rc;
	krr_wfpt_comm_destroy(h);
	CUDA_OK(cudaSetDevice(h->device));
	ncclUniqueId id;
	memcpy(&id, id128, 128);
	NCCL_OK(nccl().commInitRank(&h->comm, world, id, rank));
	h->commRank = rank, h->commWorld = world;
	return KRR_OK;
}

// one process, one handle per device (each handle was created with its device current): rank i = handles[i]
extern "C" int krr_wfpt_comm_init_all(KrrWfpt **handles, int32_t n) {
	if (!handles || n < 1 || n > 64) return fail(KRR_E_INVALID, "bad argument");
	if (int rc = needNccl()) return rc;
	int devs[64];
	ncclComm_t comms[64];
	for (int i = 0; i < n; i++) {
		if (!handles[i]) return fail(KRR_E_INVALID, "null handle");
		devs[i] = handles[i]->device;
		for (int j = 0; j < i; j++)
			if (devs[j] == devs[i]) return fail(KRR_E_INVALID, "handles %d and %d are on the same device %d: NCCL needs one device per rank", j, i, devs[i]);
		krr_wfpt_comm_destroy(handles[i]);
	}
	NCCL_OK(nccl().commInitAll(comms, n, devs));
	for (int i = 0; i < n; i++) handles[i]->comm = comms[i], handles[i]->commRank = i, handles[i]->commWorld = n;
	return KRR_OK;
}

extern "C" int krr_wfpt_comm_destroy(KrrWfpt *h) {
	if (!h) return fail(KRR_E_INVALID, "null handle");
	if (h->comm) {
		cudaSetDevice(h->device);
		nccl().commDestroy(h->comm);
		h->comm = nullptr, h->commRank = 0, h->commWorld = 1;
	}
	return KRR_OK;
}

// film (device, W x H RGBA32F, on every rank) is summed onto `root` IN PLACE and multiplied by `scale` there
// (1 / number of spp slices); the other ranks' films are left unchanged.  Ordered on `stream`.
extern "C" int krr_wfpt_reduce_film(KrrWfpt *h, float *film, int32_t root, float scale, void *stream) {
	if (!h || !film) return fail(KRR_E_INVALID, "null argument");
	if (h->width <= 0) return fail(KRR_E_STATE, "resize first");
	const size_t n = (size_t) h->width * h->height;
	cudaStream_t st = (cudaStream_t) stream;
	if (h->commWorld > 1) {
		if (!h->comm) return fail(KRR_E_STATE, "krr_wfpt_comm_init_rank / _init_all first");
		if (root < 0 || root >= h->commWorld) return fail(KRR_E_INVALID, "bad root");
		CUDA_OK(cudaSetDevice(h->device));
		NCCL_OK(nccl().reduce(film, film, n * 4, ncclFloat32, ncclSum, root, h->comm, st));
	}
	if (h->commRank == root && scale != 1.f) {
		k_scale_film<<<h->numSMs * 4, 256, 0, st>>>((float4 *) film, n, scale);
		CUDA_OK(cudaGetLastError());
	}
	return KRR_OK;
}

// render + reduce + (root) pipelined read-back: as krr_wfpt_render_to_host_async, with the film reduction
// between the render and the copy.  film_host is only used on the root rank.
extern "C" int krr_wfpt_render_reduce_to_host_async(KrrWfpt *h, float *film_host, int32_t root, float scale, void *stream) {
	if (!h) return fail(KRR_E_INVALID, "null handle");
	const bool isRoot = h->commRank == root;
	if (isRoot && !film_host) return fail(KRR_E_INVALID, "the root rank needs a host film");
	CUDA_OK(cudaSetDevice(h->device));
	const size_t n = (size_t) h->width * h->height;
	if (!h->copyStream) {
		CUDA_OK(cudaStreamCreateWithFlags(&h->copyStream, cudaStreamNonBlocking));
		for (int i = 0; i < 2; i++) {
			CUDA_OK(cudaEventCreateWithFlags(&h->evRendered[i], cudaEventDisableTiming));
			CUDA_OK(cudaEventCreateWithFlags(&h->evCopied[i], cudaEventDisableTiming));
		}
	}
	const int slot = h->asyncSlot;
	h->asyncSlot ^= 1;
	if (h->asyncFilm[slot].n != n) {
		CUDA_OK(cudaStreamSynchronize(h->copyStream));
		if (h->asyncFilm[slot].alloc(n)) return KRR_E_CUDA;
	}
	cudaStream_t st = (cudaStream_t) stream;
	CUDA_OK(cudaStreamWaitEvent(st, h->evCopied[slot], 0));
	int rc = krr_wfpt_render(h, (float *) h->asyncFilm[slot].p, stream);
	if (rc) return rc;
	rc = krr_wfpt_reduce_film(h, (float *) h->asyncFilm[slot].p, root, scale, stream);
	if (rc) return rc;
	if (isRoot) {
		CUDA_OK(cudaEventRecord(h->evRendered[slot], st));
		CUDA_OK(cudaStreamWaitEvent(h->copyStream, h->evRendered[slot], 0));
		CUDA_OK(cudaMemcpyAsync(film_host, h->asyncFilm[slot].p, n * 16, cudaMemcpyDeviceToHost, h->copyStream));
		CUDA_OK(cudaEventRecord(h->evCopied[slot], h->copyStream));
	}
	return KRR_OK;
}

extern "C" int krr_wfpt_wait_host(KrrWfpt *h) {
	if (!h) return fail(KRR_E_INVALID, "null handle");
	if (h->copyStream) CUDA_OK(cudaStreamSynchronize(h->copyStream));
	return KRR_OK;
}

extern "C" int krr_wfpt_get_stats(KrrWfpt *h, KrrStats *out) {
	if (!h || !out) return fail(KRR_E_INVALID, "null argument");
	memset(out, 0, sizeof *out);
	if (!h->totals.p) return KRR_OK;
	CUDA_OK(cudaStreamSynchronize(h->lastStream));
	StatTotals t;
	CUDA_OK(cudaMemcpy(&t, h->totals.p, sizeof t, cudaMemcpyDeviceToHost));
	int32_t flags[4];
	CUDA_OK(cudaMemcpy(flags, h->errorFlags.p, 16, cudaMemcpyDeviceToHost));
	for (int b = 1; b < h->activeBands; b++) { // bands joined the caller's stream at the end of render
		StatTotals tb;
		int32_t fb[4];
		CUDA_OK(cudaMemcpy(&tb, h->band(b).totals.p, sizeof tb, cudaMemcpyDeviceToHost));
		CUDA_OK(cudaMemcpy(fb, h->band(b).errorFlags.p, 16, cudaMemcpyDeviceToHost));
		t.camera += tb.camera, t.closest += tb.closest, t.shadow += tb.shadow, t.scatter += tb.scatter, t.hitLight += tb.hitLight;
		t.miss += tb.miss, t.mediumSample += tb.mediumSample, t.mediumScatter += tb.mediumScatter;
		for (int i = 0; i < 64; i++) t.closestByDepth[i] += tb.closestByDepth[i], t.shadowByDepth[i] += tb.shadowByDepth[i];
		for (int i = 0; i < 4; i++) flags[i] |= fb[i];
	}
	if (flags[0]) return fail(KRR_E_CUDA, "BVH traversal stack overflow (scene deeper than %d entries)", kStackSize);
#ifdef KRR_COUNT_TRIPS
	{ // debug build: node visits / triangle tests of the closest rays since the last call
		int32_t all[4 + 132];
		CUDA_OK(cudaMemcpy(all, h->errorFlags.p, sizeof all, cudaMemcpyDeviceToHost));
		CUDA_OK(cudaMemset(h->errorFlags.p, 0, sizeof all));
		const double n = (double) t.closest;
		fprintf(stderr, "[trips] closest rays %.0f: node visits/ray %.2f (max %d), triangle tests/ray %.2f, instances entered/ray %.2f, sphere-culled/ray %.2f; histogram (bins of 8 visits):", n, all[1] / n, all[2], all[3] / n, all[4 + 128] / n, all[4 + 129] / n);
		for (int i = 0; i < 64; i++) fprintf(stderr, " %llu", ((unsigned long long *) (all + 4))[i]);
		fprintf(stderr, "\n");
	}
#endif
	out->camera_rays = t.camera, out->closest_rays = t.closest, out->shadow_rays = t.shadow, out->scatter_items = t.scatter;
	out->hit_light_items = t.hitLight, out->miss_items = t.miss, out->medium_sample_items = t.mediumSample, out->medium_scatter_items = t.mediumScatter;
	for (int i = 0; i < 64; i++) out->closest_by_depth[i] = t.closestByDepth[i], out->shadow_by_depth[i] = t.shadowByDepth[i];
	out->kernel_launches = h->launches;
	out->bvh_nodes = h->bvh.nodeCount(), out->bvh_triangles = h->bvh.triCount(), out->tlas_nodes = h->bvh.tlasNodeCount();
	return KRR_OK;
}

extern "C" int krr_wfpt_set_profiling(KrrWfpt *h, int32_t enable) {
	if (!h) return fail(KRR_E_INVALID, "null handle");
	h->profile = enable != 0;
	h->evRecs.clear();
	h->evUsed = 0;
	return KRR_OK;
}

extern "C" int krr_wfpt_get_stage_times(KrrWfpt *h, double *ms, int32_t *launches, int32_t reset) {
	if (!h || !ms || !launches) return fail(KRR_E_INVALID, "null argument");
	for (int i = 0; i < KRR_STAGE_COUNT; i++) ms[i] = 0, launches[i] = 0;
	if (!h->evRecs.empty()) CUDA_OK(cudaEventSynchronize(h->evRecs.back().b));
	for (auto &r : h->evRecs) {
		float t = 0;
		CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
		ms[r.stage] += t, launches[r.stage]++;
	}
	if (reset) { h->evRecs.clear(); h->evUsed = 0; }
	return KRR_OK;
}

extern "C" int krr_wfpt_get_launch_times(KrrWfpt *h, int32_t *stage, float *ms, int32_t capacity) {
	if (!h || !stage || !ms) return fail(KRR_E_INVALID, "null argument");
	if (!h->evRecs.empty()) CUDA_OK(cudaEventSynchronize(h->evRecs.back().b));
	int n = 0;
	for (auto &r : h->evRecs) {
		if (n < capacity) {
			float t = 0;
			CUDA_OK(cudaEventElapsedTime(&t, r.a, r.b));
			stage[n] = r.stage, ms[n] = t;
		}
		n++;
	}
	return n;
}

extern "C" int krr_wfpt_debug_first_hits(KrrWfpt *h, int32_t *inst, int32_t *prim) {
	if (!h || !h->firstHits.p) return fail(KRR_E_STATE, "no depth-0 hits recorded: create the pass with \"debug_taps\": true");
	CUDA_OK(cudaDeviceSynchronize());
	std::vector<int4> tmp(h->pixelCount());
	CUDA_OK(cudaMemcpy(tmp.data(), h->firstHits.p, tmp.size() * 16, cudaMemcpyDeviceToHost));
	for (size_t i = 0; i < tmp.size(); i++) {
		if (inst) inst[i] = tmp[i].x;
		if (prim) prim[i] = tmp[i].y;
	}
	return KRR_OK;
}

extern "C" int krr_wfpt_debug_pixel_state(KrrWfpt *h, uint64_t *sampler, float *lambda, float *cameraSample) {
	if (!h || !h->rng.p) return fail(KRR_E_STATE, "no state");
	if (h->activeBands > 1) return fail(KRR_E_STATE, "the last frame ran as %d bands: create the pass with \"debug_taps\": true (or \"bands\": 1) to read pixel state", h->activeBands);
	CUDA_OK(cudaDeviceSynchronize());
	const size_t n = h->pixelCount();
	if (sampler) {
		std::vector<uint64_t> st(n);
		CUDA_OK(cudaMemcpy(st.data(), h->rng.p, n * 8, cudaMemcpyDeviceToHost));
		uint32_t seedIndex = (uint32_t) (h->frameIndex * (uint64_t) h->spp);
		for (size_t i = 0; i < n; i++) sampler[2 * i] = st[i], sampler[2 * i + 1] = ((uint64_t) seedIndex << 1u) | 1u;
	}
	if (lambda) {
		std::vector<float> l0(n);
		CUDA_OK(cudaMemcpy(l0.data(), h->lambda.p, n * 4, cudaMemcpyDeviceToHost));
		for (size_t i = 0; i < n; i++) {
			Wavelengths w = expandWavelengths(l0[i]);
			memcpy(lambda + 4 * i, w.lambda, 16);
		}
	}
	if (cameraSample) {
		if (!h->cameraSample.p) return fail(KRR_E_STATE, "no camera samples recorded: create the pass with \"debug_taps\": true");
		CUDA_OK(cudaMemcpy(cameraSample, h->cameraSample.p, n * 20, cudaMemcpyDeviceToHost));
	}
	return KRR_OK;
}

extern "C" int krr_wfpt_debug_capture(KrrWfpt *h, int32_t sampleId, int32_t depth) {
	if (!h) return fail(KRR_E_INVALID, "null handle");
	h->capSample = sampleId, h->capDepth = depth;
	return KRR_OK;
}

extern "C" int krr_wfpt_debug_queue(KrrWfpt *h, int32_t queue, int32_t *items, int32_t capacity) {
	if (!h || queue < 0 || queue > 5) return fail(KRR_E_INVALID, "bad queue");
	if (!h->capCounts.p || !h->capItems[queue].p) return fail(KRR_E_STATE, "nothing captured");
	CUDA_OK(cudaDeviceSynchronize());
	int32_t n = 0;
	CUDA_OK(cudaMemcpy(&n, h->capCounts.p + queue, 4, cudaMemcpyDeviceToHost));
	if (items) {
		if (n > capacity) return fail(KRR_E_INVALID, "capacity %d < %d items", capacity, n);
		CUDA_OK(cudaMemcpy(items, h->capItems[queue].p, (size_t) n * 16, cudaMemcpyDeviceToHost));
	}
	return n;
}

extern "C" int krr_accumulate_f32(float *accum, float *film, int64_t n, uint64_t accumCount, uint64_t maxAccum, int32_t movingAverage, void *stream) {
	if (!accum || !film || n <= 0) return fail(KRR_E_INVALID, "bad argument");
	k_accumulate<<<148 * 4, 256, 0, (cudaStream_t) stream>>>((float4 *) accum, (float4 *) film, n, accumCount, maxAccum, movingAverage);
	CUDA_OK(cudaGetLastError());
	return KRR_OK;
}

// ---- leaf-function taps (parity tests only) ----
namespace {
template <typename Q, typename R, typename Launch>
int leafRun(const Q *q, int32_t n, size_t qElems, R *out, size_t rElems, Launch launch) {
	if (!q || !out || n <= 0) return fail(KRR_E_INVALID, "bad argument");
	Buf<Q> dq;
	Buf<R> dr;
	if (dq.alloc(qElems * n) || dr.alloc(rElems * n)) return KRR_E_CUDA;
	CUDA_OK(cudaMemcpy(dq.p, q, sizeof(Q) * qElems * n, cudaMemcpyHostToDevice));
	launch(dq.p, dr.p);
	CUDA_OK(cudaGetLastError());
	CUDA_OK(cudaDeviceSynchronize());
	CUDA_OK(cudaMemcpy(out, dr.p, sizeof(R) * rElems * n, cudaMemcpyDeviceToHost));
	return KRR_OK;
}
} // namespace

extern "C" int krr_wfpt_debug_eval_bsdf(KrrWfpt *h, const KrrLeafBsdfQuery *q, int32_t n, KrrLeafBsdfResult *out) {
	if (!h || !h->haveColorSpace) return fail(KRR_E_STATE, "set_color_space first");
	return leafRun(q, n, 1, out, 1, [&](const KrrLeafBsdfQuery *dq, KrrLeafBsdfResult *dr) { k_leaf_bsdf<<<(n + 63) / 64, 64>>>(dq, n, dr, h->cs); });
}

extern "C" int krr_wfpt_debug_eval_light(KrrWfpt *h, const KrrLeafLightQuery *q, int32_t n, KrrLeafLightResult *out) {
	if (!h || !h->haveColorSpace) return fail(KRR_E_STATE, "set_color_space first");
	if (!q || n <= 0) return fail(KRR_E_INVALID, "bad argument");
	std::vector<Xf> inv(n);
	for (int i = 0; i < n; i++) {
		Xf t;
		memcpy(t.m, q[i].transform, 48);
		inv[i] = xfInverse(t);
	}
	Buf<Xf> dinv;
	if (dinv.upload(inv)) return KRR_E_CUDA;
	SceneDev sc{};
	sc.cs = h->cs;
	return leafRun(q, n, 1, out, 1, [&](const KrrLeafLightQuery *dq, KrrLeafLightResult *dr) { k_leaf_light<<<(n + 63) / 64, 64>>>(dq, dinv.p, n, dr, sc); });
}

extern "C" int krr_wfpt_debug_instance_xf(KrrWfpt *h, const int32_t *ids, const float *times, int32_t n, float *out) {
	if (!h || !h->haveScene) return fail(KRR_E_STATE, "set_scene first");
	if (!ids || !times || !out || n <= 0) return fail(KRR_E_INVALID, "bad argument");
	for (int i = 0; i < n; i++)
		if (ids[i] < 0 || ids[i] >= h->scene.nInstances) return fail(KRR_E_INVALID, "instance id out of range");
	Buf<int32_t> dIds;
	if (dIds.upload(std::vector<int32_t>(ids, ids + n))) return KRR_E_CUDA;
	return leafRun(times, n, 1, out, 24, [&](const float *dt, float *dr) { k_leaf_instance_xf<<<(n + 63) / 64, 64>>>(h->scene, dIds.p, dt, n, dr); });
}

extern "C" int krr_wfpt_debug_eval_color(KrrWfpt *h, const float *in, int32_t n, float *out) {
	if (!h || !h->haveColorSpace) return fail(KRR_E_STATE, "set_color_space first");
	return leafRun(in, n, 8, out, 20, [&](const float *dq, float *dr) { k_leaf_color<<<(n + 63) / 64, 64>>>(dq, n, dr, h->cs); });
}

extern "C" int krr_wfpt_debug_camera_rays(KrrWfpt *h, const KrrCameraData *c, int32_t W, int32_t H, const float *in, int32_t n, float *out) {
	if (!h || !c || W <= 0 || H <= 0) return fail(KRR_E_INVALID, "bad argument");
	KrrCameraDev cam = makeCamera(c);
	return leafRun(in, n, 7, out, 7, [&](const float *dq, float *dr) { k_leaf_camera<<<(n + 63) / 64, 64>>>(cam, W, H, dq, n, dr); });
}
