// megakernel.cuh -- the reference's MegakernelPathTracer as ONE kernel.
//
// Reference: src/render/megakernel/device.cu:50-195 (__raygen__Pathtracer and its helpers), path.h:27-58,
// pathtracer.h:33-45 (JSON: nee, max_depth, rr).  It is an independent estimator of the same integral as the
// wavefront pass -- power-heuristic MIS on (bsdf pdf, light pdf) instead of the spectral path pdfs pu / pl,
// one lane per pixel from camera ray to termination -- built from the same device routines (traversal,
// interaction rebuild, BSDFs, lights).  Its purpose here is cross-validation: tests/test_gpu_passes.py checks
// that both estimators converge to the same image.  It is not a fast path (every lane runs its own
// traversal to the end, all five BSDFs are resident in one kernel).
#pragma once
#include "wavefront_kernels.cuh"

namespace krr {

struct MegaParams {
	int32_t width, height, spp, maxDepth, nee;
	float probRR;
	uint32_t frameId;
};

KRR_DEV float evalMIS(float n0, float p0, float n1, float p1) { // render/sampling.h:20-28 (power heuristic)
	float q0 = (n0 * p0) * (n0 * p0), q1 = (n1 * p1) * (n1 * p1);
	return q0 / (q0 + q1);
}

template <bool MOTION>
__device__ __noinline__ Hit megaTraceClosest(const Wavefront &wf, TraceSmem &sm, V3 o, V3 d, float time) {
	Traverser<false, MOTION> tr;
	LocalStack<false> ls;
	tr.begin(wf.bvh, o, d, kInf, time);
	tr.runToEnd(wf.bvh, wf.scene.instances, sm, ls, [&](int inst, int prim, float u, float v) {
		if (!(wf.instFlags[inst] & 2)) return true;
		return !alphaKilled(wf, inst, prim, u, v, o, d); // __anyhit__Radiance
	});
	if (tr.overflow) atomicExch(&wf.errorFlags[0], 1);
	return tr.best;
}
template <bool MOTION>
__device__ __noinline__ bool megaVisible(const Wavefront &wf, TraceSmem &sm, V3 o, V3 d, float time) {
	Traverser<true, MOTION> tr;
	LocalStack<true> ls;
	tr.begin(wf.bvh, o, d, 1.f, time);
	tr.runToEnd(wf.bvh, wf.scene.instances, sm, ls, [&](int inst, int prim, float u, float v) {
		if (!(wf.instFlags[inst] & 2)) return true;
		return !alphaKilled(wf, inst, prim, u, v, o, d); // __anyhit__ShadowRay
	});
	if (tr.overflow) atomicExch(&wf.errorFlags[0], 1);
	return tr.best.inst < 0;
}

// one surface vertex with BSDF type MT: evalDirect + generateScatterRay (device.cu:81-127).
// Returns false when the path ends here.
template <int MT, bool MOTION>
__device__ __noinline__ bool megaVertex(const Wavefront &wf, TraceSmem &sm, const MegaParams &mp, const SurfaceGeom &g, const ShadingData &sd,
										const Wavelengths &wl, Pcg &rng, float time, Spec &thp, Spec &L, V3 &ro, V3 &rd, float &pdfPrev, int &typePrev) {
	auto toLocal = [&](V3 v) { return mk3(dot(g.tangent, v), dot(g.bitangent, v), dot(g.n, v)); };
	const V3 woLocal = toLocal(g.wo);
	Bsdf<MT> bsdf;
	BsdfSetupCtx ctx{g.wo, &wl, &wf.scene.cs};
	bsdf.setup(sd, ctx);
	const float lightSelPdf = wf.scene.nLights > 0 ? 1.f / wf.scene.nLights : 0.f;
	if (mp.nee && (getBsdfType(sd) & BSDF_SMOOTH) && wf.scene.nLights > 0) { // evalDirect, lightSamples = 1
		const float ul = rng.get1D();
		const uint32_t lightId = min((uint32_t) (ul * wf.scene.nLights), (uint32_t) wf.scene.nLights - 1);
		const LightRec lr = wf.scene.lights[lightId];
		const float u0 = rng.get1D(), u1 = rng.get1D();
		LightSample ls;
		bool delta = false;
		if (lr.type == LIGHT_DIFFUSE_AREA) {
			const TriLightRec &tl = wf.scene.triLights[lr.index];
			ls = areaLightSampleLi(tl, wf.scene.instances[tl.inst], u0, u1, g.p, wl, wf.scene.cs);
		} else {
			const AnalyticLightRec &al = wf.scene.analytic[lr.index];
			ls	  = analyticSampleLi(al, u0, u1, g.p, wl, wf.scene);
			delta = al.type != LIGHT_INFINITE;
		}
		const V3 wiLocal	 = toLocal(normalize(ls.p - g.p));
		const float lightPdf = lightSelPdf * ls.pdf;
		if (lightPdf != 0) {
			const float bsdfPdf = delta ? 0.f : bsdf.pdf(woLocal, wiLocal);
			const Spec bsdfVal	= bsdf.f(woLocal, wiLocal) * fabsf(wiLocal.z);
			const float mis		= evalMIS(1, lightPdf, 1, bsdfPdf);
			if (!(isnan(mis) || isinf(mis)) && any(bsdfVal)) {
				auto offs = [](V3 p, V3 n, V3 w) { V3 off = n * kRayEps; if (dot(n, w) < 0.f) off = -off; return p + off; };
				const V3 to = offs(ls.p, ls.n, g.p - ls.p), po = offs(g.p, g.n, to - g.p); // spawnRayTo(ls.intr)
				if (megaVisible<MOTION>(wf, sm, po, to - po, time)) L += thp * bsdfVal * mis / (1 * lightPdf) * ls.L;
			}
		}
	}
	const BSDFSample bs = bsdf.sample(woLocal, rng); // generateScatterRay
	if (bs.pdf == 0 || !any(bs.f)) return false;
	const V3 wiWorld = g.tangent * bs.wi.x + g.bitangent * bs.wi.y + g.n * bs.wi.z;
	typePrev = bs.flags, pdfPrev = bs.pdf;
	V3 off = g.n * kRayEps;
	if (dot(g.n, wiWorld) < 0.f) off = -off;
	ro = g.p + off, rd = wiWorld; // spawnRayTowards
	thp = thp * bs.f * fabsf(bs.wi.z) / bs.pdf;
	return any(thp);
}

template <bool MOTION>
__global__ void __launch_bounds__(kTraceBlock) k_megakernel(const __grid_constant__ Wavefront wf, MegaParams mp, float4 *film) {
	__shared__ TraceSmem sm;
	const int nPix = mp.width * mp.height;
	const float lightSelPdf = wf.scene.nLights > 0 ? 1.f / wf.scene.nLights : 0.f;
	for (int pixelId = blockIdx.x * blockDim.x + threadIdx.x; pixelId < nPix; pixelId += gridDim.x * blockDim.x) {
		const int px = pixelId % mp.width, py = pixelId / mp.width;
		Pcg rng;
		rng.setPixelSample((uint32_t) px, (uint32_t) py, mp.frameId * 512u); // device.cu:159
		float color[3] = {0, 0, 0};
		for (int s = 0; s < mp.spp; s++) {
			Spec thp = sp(1), L = sp(0);
			float cs[5];
			for (int k = 0; k < 5; k++) cs[k] = rng.get1D();
			V3 ro, rd;
			float time;
			cameraRay(wf.cam, px, py, mp.width, mp.height, cs, ro, rd, time);
			float lam		= sampleLambda0(rng.get1D());
			Wavelengths wl	= expandWavelengths(lam);
			float pdfPrev	= 0;
			int typePrev	= 0;
			V3 ctxP = mk3(0, 0, 0);
			// (the reference does not reset path.depth between the samples of a pixel; with its default spp = 1
			// that never shows, and it is not reproduced)
			for (int depth = 0; true; depth++) {
				const Hit h = megaTraceClosest<MOTION>(wf, sm, ro, rd, time);
				if (h.inst < 0) { // handleMiss, device.cu:66-79
					for (int li = 0; li < wf.scene.nInfinite; li++) {
						const AnalyticLightRec &light = wf.scene.analytic[wf.scene.infiniteLights[li]];
						float weight = 1;
						if (mp.nee && depth > 0 && !(typePrev & BSDF_SPECULAR)) {
							weight = evalMIS(1, pdfPrev, 1, kInv4Pi * lightSelPdf);
							if (isnan(weight) || isinf(weight)) weight = 1;
						}
						L += thp * weight * infiniteLi(light, rd, wl, wf.scene);
					}
					break;
				}
				SurfaceGeom g;
				rebuildGeometry<MOTION>(wf, make_int4(h.inst, h.prim, __float_as_int(h.u), __float_as_int(h.v)), rd, time, g);
				if (g.material < 0) break; // medium interfaces have no material: the megakernel does not handle media
				ShadingData sd;
				bool term;
				evalMaterial(wf, g, wl, sd, term);
				if (term && lam > 0) { lam = -lam; wl = expandWavelengths(lam); }
				if (g.light >= 0) { // handleHit, device.cu:50-64
					const LightRec lr	  = wf.scene.lights[g.light];
					const TriLightRec &tl = wf.scene.triLights[lr.index];
					const Spec Le = areaLightL(tl, g.n, g.wo, wl, wf.scene.cs);
					float weight  = 1;
					if (mp.nee && depth > 0) {
						const float lightPdf = areaLightPdfLi(tl, wf.scene.instances[tl.inst], g.p, g.n, ctxP) * lightSelPdf;
						if (!(typePrev & BSDF_SPECULAR)) weight = evalMIS(1, pdfPrev, 1, lightPdf);
						if (isnan(weight) || isinf(weight)) weight = 1;
					}
					L += Le * weight * thp;
				}
				if (depth == mp.maxDepth || (mp.probRR < 1.f && rng.get1D() > mp.probRR)) break;
				thp = thp / mp.probRR;
				ctxP = g.p;
				bool go;
				switch (sd.bsdfType) {
					case MAT_DIFFUSE: go = megaVertex<MAT_DIFFUSE, MOTION>(wf, sm, mp, g, sd, wl, rng, time, thp, L, ro, rd, pdfPrev, typePrev); break;
					case MAT_DIELECTRIC: go = megaVertex<MAT_DIELECTRIC, MOTION>(wf, sm, mp, g, sd, wl, rng, time, thp, L, ro, rd, pdfPrev, typePrev); break;
					case MAT_CONDUCTOR: go = megaVertex<MAT_CONDUCTOR, MOTION>(wf, sm, mp, g, sd, wl, rng, time, thp, L, ro, rd, pdfPrev, typePrev); break;
					case MAT_DISNEY: go = megaVertex<MAT_DISNEY, MOTION>(wf, sm, mp, g, sd, wl, rng, time, thp, L, ro, rd, pdfPrev, typePrev); break;
					default: go = megaVertex<MAT_NULL, MOTION>(wf, sm, mp, g, sd, wl, rng, time, thp, L, ro, rd, pdfPrev, typePrev); break;
				}
				if (!go) break;
			}
			float rgb[3];
			toRGB(L, wl, wf.scene.cs, rgb);
			for (int k = 0; k < 3; k++) color[k] += rgb[k];
		}
		// colorBuffer.write(RGBA(color, 1), fbIndex): the SUM over the samples (device.cu:194), row H-1-y
		film[(size_t) (mp.height - 1 - py) * mp.width + px] = make_float4(color[0], color[1], color[2], 1.f);
	}
}

} // namespace krr
