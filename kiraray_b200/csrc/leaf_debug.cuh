// leaf_debug.cuh -- parity taps: one thread per query runs the device leaf functions (BSDFs, lights,
// colour, camera) on caller-supplied inputs.  Never on the render path; see krr_wfpt.h.
#pragma once
#include "krr_wfpt.h"
#include "wavefront_kernels.cuh"

namespace krr {

// object<->world of instance ids[i] at times[i] (24 floats per query: 3x4 transform, 3x4 inverse)
__global__ void k_leaf_instance_xf(SceneDev sc, const int32_t *ids, const float *times, int n, float *out24) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const InstRec &in = sc.instances[ids[i]];
	Xf m = in.xf, inv = in.inv;
	if (in.motion >= 0) movingInstanceXf(sc.xnodes, sc.motionKeys, sc.motionFlat, ids[i], in.motion, times[i], m, inv); // the path the kernels take
	for (int k = 0; k < 12; k++) out24[24 * i + k] = m.m[k], out24[24 * i + 12 + k] = inv.m[k];
}

template <int MT> KRR_DEV void leafBsdf(const ShadingData &sd, const BsdfSetupCtx &ctx, V3 wo, V3 wi, Pcg &rng, KrrLeafBsdfResult &r) {
	Bsdf<MT> bsdf;
	bsdf.setup(sd, ctx);
	Spec f = bsdf.f(wo, wi);
	r.f[0] = f.x, r.f[1] = f.y, r.f[2] = f.z, r.f[3] = f.w;
	r.pdf = bsdf.pdf(wo, wi);
	BSDFSample s = bsdf.sample(wo, rng);
	r.s_f[0] = s.f.x, r.s_f[1] = s.f.y, r.s_f[2] = s.f.z, r.s_f[3] = s.f.w;
	r.s_wi[0] = s.wi.x, r.s_wi[1] = s.wi.y, r.s_wi[2] = s.wi.z;
	r.s_pdf = s.pdf, r.s_flags = s.flags;
}

__global__ void k_leaf_bsdf(const KrrLeafBsdfQuery *q, int n, KrrLeafBsdfResult *out, ColorSpaceDev cs) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const KrrLeafBsdfQuery &k = q[i];
	ShadingData sd{};
	sd.IoR = k.ior;
	sd.diffuse	= make_float4(k.diffuse[0], k.diffuse[1], k.diffuse[2], k.diffuse[3]);
	sd.specular = make_float4(k.specular[0], k.specular[1], k.specular[2], k.specular[3]);
	sd.specularTransmission = k.specular_transmission, sd.roughness = k.roughness, sd.metallic = k.metallic, sd.anisotropic = k.anisotropic;
	sd.bsdfType = k.bsdf_type;
	sd.hasEta = k.eta_kind == 1, sd.etaSpec = sp(k.eta), sd.hasK = 0, sd.kSpec = sp(0);
	Wavelengths wl = expandWavelengths(sampleLambda0(k.wavelength_u));
	V3 wo = mk3(k.wo[0], k.wo[1], k.wo[2]), wi = mk3(k.wi[0], k.wi[1], k.wi[2]);
	BsdfSetupCtx ctx{wo, &wl, &cs};
	Pcg rng;
	rng.setPixelSample(k.seed_px, k.seed_py, k.seed_index);
	KrrLeafBsdfResult r{};
	r.type_flags = getBsdfType(sd);
	switch (k.bsdf_type) {
		case MAT_NULL: leafBsdf<MAT_NULL>(sd, ctx, wo, wi, rng, r); break;
		case MAT_DIFFUSE: leafBsdf<MAT_DIFFUSE>(sd, ctx, wo, wi, rng, r); break;
		case MAT_DIELECTRIC: leafBsdf<MAT_DIELECTRIC>(sd, ctx, wo, wi, rng, r); break;
		case MAT_CONDUCTOR: leafBsdf<MAT_CONDUCTOR>(sd, ctx, wo, wi, rng, r); break;
		default: leafBsdf<MAT_DISNEY>(sd, ctx, wo, wi, rng, r); break;
	}
	out[i] = r;
}

__global__ void k_leaf_light(const KrrLeafLightQuery *q, const Xf *inverses, int n, KrrLeafLightResult *out, SceneDev sc) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const KrrLeafLightQuery &k = q[i];
	Wavelengths wl = expandWavelengths(sampleLambda0(k.wavelength_u));
	V3 ctxP = mk3(k.ctx_p[0], k.ctx_p[1], k.ctx_p[2]);
	KrrLeafLightResult r{};
	auto st4 = [](float *o, Spec s) { o[0] = s.x, o[1] = s.y, o[2] = s.z, o[3] = s.w; };
	if (k.type == LIGHT_DIFFUSE_AREA) {
		TriLightRec t{};
		for (int c = 0; c < 3; c++) for (int a = 0; a < 3; a++) t.p[c][a] = k.p[c][a], t.n[c][a] = k.n[c][a];
		t.scale = k.scale, t.twoSided = k.two_sided, t.hasNormals = 1;
		t.LeSpec = makeUnbounded(sc.cs.zNodes, sc.cs.coeffs, k.color[0], k.color[1], k.color[2]);
		InstRec in{};
		for (int a = 0; a < 12; a++) in.xf.m[a] = k.transform[a];
		in.inv = inverses[i];
		LightSample ls = areaLightSampleLi(t, in, k.u[0], k.u[1], ctxP, wl, sc.cs);
		r.p[0] = ls.p.x, r.p[1] = ls.p.y, r.p[2] = ls.p.z, r.n[0] = ls.n.x, r.n[1] = ls.n.y, r.n[2] = ls.n.z;
		st4(r.L, ls.L), r.pdf = ls.pdf;
		V3 w = normalize(ctxP - ls.p);
		st4(r.L_eval, areaLightL(t, ls.n, w, wl, sc.cs));
		r.pdf_li = areaLightPdfLi(t, in, ls.p, ls.n, ctxP);
	} else {
		AnalyticLightRec l{};
		l.type = k.type, l.scale = k.scale, l.sceneRadius = k.scene_radius, l.cosInner = k.cos_inner, l.cosOuter = k.cos_outer;
		for (int a = 0; a < 3; a++) l.color[a] = k.color[a];
		for (int a = 0; a < 12; a++) l.xf.m[a] = k.transform[a];
		l.inv = inverses[i];
		l.position[0] = k.transform[3], l.position[1] = k.transform[7], l.position[2] = k.transform[11];
		for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) l.rotation[a * 3 + b] = k.transform[a * 4 + b];
		l.colorSpec = makeUnbounded(sc.cs.zNodes, sc.cs.coeffs, k.color[0], k.color[1], k.color[2]);
		l.image.valid = 0;
		LightSample ls = analyticSampleLi(l, k.u[0], k.u[1], ctxP, wl, sc);
		r.p[0] = ls.p.x, r.p[1] = ls.p.y, r.p[2] = ls.p.z;
		st4(r.L, ls.L), r.pdf = ls.pdf;
		if (k.type == LIGHT_INFINITE) st4(r.L_eval, infiniteLi(l, mk3(k.wi[0], k.wi[1], k.wi[2]), wl, sc));
	}
	out[i] = r;
}

__global__ void k_leaf_color(const float *in, int n, float *out, ColorSpaceDev cs) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float *q = in + 8 * i;
	float *o = out + 20 * i;
	Wavelengths wl = expandWavelengths(sampleLambda0(q[3]));
	Spec b	= sampleBounded(makeBounded(cs.zNodes, cs.coeffs, q[0], q[1], q[2]), wl);
	RgbSpectrum us = makeUnbounded(cs.zNodes, cs.coeffs, q[0], q[1], q[2]);
	Spec u	= sampleUnbounded(us, wl), il = sampleIlluminant(us, wl, cs);
	o[0] = b.x, o[1] = b.y, o[2] = b.z, o[3] = b.w;
	o[4] = u.x, o[5] = u.y, o[6] = u.z, o[7] = u.w;
	o[8] = il.x, o[9] = il.y, o[10] = il.z, o[11] = il.w;
	Spec s = make_float4(q[4], q[5], q[6], q[7]);
	toRGB(s, wl, cs, o + 12);
	o[15] = lum(s, wl, cs);
	for (int k = 0; k < 4; k++) o[16 + k] = wl.lambda[k];
}

__global__ void k_leaf_camera(KrrCameraDev cam, int W, int H, const float *in, int n, float *out) {
	int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float *q = in + 7 * i;
	V3 o, d;
	float time;
	cameraRay(cam, (int) q[0], (int) q[1], W, H, q + 2, o, d, time);
	float *r = out + 7 * i;
	r[0] = o.x, r[1] = o.y, r[2] = o.z, r[3] = d.x, r[4] = d.y, r[5] = d.z, r[6] = time;
}

} // namespace krr
