// lights.cuh -- light sampling / evaluation on the device.
// Follows reference src/core/light.h:30-259 (Point/Directional/Spot/DiffuseArea/Infinite),
// src/core/shape.h:23-132 (Triangle area/sample/pdf), src/render/lightsampler.h:18-40 (uniform).
#pragma once
#include "bsdf.cuh"
#include "scene.cuh"

namespace krr {

enum : int { LIGHT_POINT = 0, LIGHT_DIRECTIONAL = 1, LIGHT_SPOT = 2, LIGHT_DIFFUSE_AREA = 3, LIGHT_INFINITE = 4 };

struct LightSample {
	V3 p, n;   // sampled point (n = 0 for analytic lights)
	Spec L;
	float pdf;
};

KRR_DEV V3 ld3(const float *p) { return mk3(p[0], p[1], p[2]); }

// Triangle::area, shape.h:30-41 (vertices transformed to world: instances may be scaled)
KRR_DEV float triArea(const TriLightRec &t, const InstRec &in) {
	V3 p0 = xfPoint(in.xf, ld3(t.p[0])), p1 = xfPoint(in.xf, ld3(t.p[1])), p2 = xfPoint(in.xf, ld3(t.p[2]));
	return 0.5f * length(cross(p1 - p0, p2 - p0));
}

// DiffuseAreaLight::L, light.h:182-188.  The light was built with the material's (valid, constant)
// emissive TEXTURE (mesh.cpp:44-57), so L() evaluates that un-normalised colour and still applies
// scale = max(Le): TriLightRec::Le / LeSpec hold exactly the colour the reference evaluates.
KRR_DEV Spec areaLightL(const TriLightRec &t, V3 n, V3 w, const Wavelengths &wl, const ColorSpaceDev &cs) {
	if (!t.twoSided && dot(n, w) < 0.f) return sp(0);
	return t.scale * sampleIlluminant(t.LeSpec, wl, cs);
}

// Triangle::sample(u) + sample(u, ctx), shape.h:54-113; DiffuseAreaLight::sampleLi, light.h:167-180
KRR_DEV LightSample areaLightSampleLi(const TriLightRec &t, const InstRec &in, float u0, float u1, V3 ctxP,
									  const Wavelengths &wl, const ColorSpaceDev &cs) {
	LightSample ls;
	V3 p0 = ld3(t.p[0]), p1 = ld3(t.p[1]), p2 = ld3(t.p[2]);
	V3 b  = uniformSampleTriangle(u0, u1);
	V3 p  = b.x * p0 + b.y * p1 + b.z * p2;
	V3 n  = normalize(cross(p1 - p0, p2 - p0));
	if (t.hasNormals) {
		V3 ns = normalize(b.x * ld3(t.n[0]) + b.y * ld3(t.n[1]) + b.z * ld3(t.n[2]));
		if (dot(n, ns) < 0) n = -n;
	}
	p = xfPoint(in.xf, p);
	n = normalize(xfNormal(in.inv, n));
	float pdf = 1 / triArea(t, in);
	V3 wi	  = normalize(p - ctxP);
	V3 dcp	  = ctxP - p;
	pdf /= fabsf(dot(n, wi)) / dot(dcp, dcp);
	if (length(wi) == 0 || isinf(pdf)) pdf = 0;
	ls.p = p, ls.n = n, ls.pdf = pdf;
	V3 wo = normalize(ctxP - p);
	ls.L  = areaLightL(t, n, wo, wl, cs);
	return ls;
}

// DiffuseAreaLight::pdfLi -> Triangle::pdf(sample, ctx), shape.h:119-128
KRR_DEV float areaLightPdfLi(const TriLightRec &t, const InstRec &in, V3 p, V3 n, V3 ctxP) {
	V3 wi	  = normalize(p - ctxP);
	V3 dcp	  = ctxP - p;
	float pdf = (1 / triArea(t, in)) / (fabsf(dot(n, -wi)) / dot(dcp, dcp));
	if (length(wi) == 0 || isinf(pdf)) pdf = 0;
	return pdf;
}

// texture fetch: constant, or bilinear + wrap over an RGBA32F image (cudaFilterModeLinear,
// normalized coordinates, cudaAddressModeWrap -- src/core/texture.cpp:229-246)
KRR_DEV float4 sampleTex(const TexRec &t, const float4 *__restrict__ texels, float u, float v, float4 fallback) {
	if (!t.valid) return fallback;
	if (t.texOff < 0) return make_float4(t.value[0], t.value[1], t.value[2], t.value[3]);
	float x = u * t.width - 0.5f, y = v * t.height - 0.5f;
	float fx = floorf(x), fy = floorf(y);
	float ax = x - fx, ay = y - fy;
	int x0 = ((int) fx % t.width + t.width) % t.width, x1 = (x0 + 1) % t.width;
	int y0 = ((int) fy % t.height + t.height) % t.height, y1 = (y0 + 1) % t.height;
	const float4 *img = texels + t.texOff;
	float4 a = __ldg(img + y0 * t.width + x0), b = __ldg(img + y0 * t.width + x1);
	float4 c = __ldg(img + y1 * t.width + x0), d = __ldg(img + y1 * t.width + x1);
	auto mix = [&](float p, float q, float r, float s) { return (1 - ay) * ((1 - ax) * p + ax * q) + ay * ((1 - ax) * r + ax * s); };
	return make_float4(mix(a.x, b.x, c.x, d.x), mix(a.y, b.y, c.y, d.y), mix(a.z, b.z, c.z, d.z), mix(a.w, b.w, c.w, d.w));
}

// InfiniteLight::Li, light.h:238-242; worldToLatLong, util/math_utils.h:136-142
KRR_DEV Spec infiniteLi(const AnalyticLightRec &l, V3 wi, const Wavelengths &wl, const SceneDev &sc) {
	// rotation^T * wi
	V3 d = mk3(l.rotation[0] * wi.x + l.rotation[3] * wi.y + l.rotation[6] * wi.z, l.rotation[1] * wi.x + l.rotation[4] * wi.y + l.rotation[7] * wi.z,
			   l.rotation[2] * wi.x + l.rotation[5] * wi.y + l.rotation[8] * wi.z);
	if (l.image.valid) {
		V3 p = normalize(d);
		float u = atan2f(p.x, -p.z) * kInv2Pi + 0.5f, v = acosf(p.y) * kInvPi;
		float4 c = sampleTex(l.image, sc.texels, u, v, make_float4(1, 1, 1, 1));
		float r = l.color[0] * c.x, g = l.color[1] * c.y, b = l.color[2] * c.z;
		RgbSpectrum s = makeUnbounded(sc.cs.zNodes, sc.cs.coeffs, r, g, b);
		return l.scale * sampleIlluminant(s, wl, sc.cs);
	}
	return l.scale * sampleIlluminant(l.colorSpec, wl, sc.cs);
}

KRR_DEV float smoothStep(float x, float a, float b) {
	if (a == b) return (x < a) ? 0.f : 1.f;
	float t = clampf((x - a) / (b - a), 0.f, 1.f);
	return t * t * (3 - 2 * t);
}

// Point / Directional / Spot / Infinite ::sampleLi, light.h:40-46, 72-81, 108-116, 220-231
KRR_DEV LightSample analyticSampleLi(const AnalyticLightRec &l, float u0, float u1, V3 ctxP, const Wavelengths &wl, const SceneDev &sc) {
	LightSample ls;
	ls.n = mk3(0, 0, 0);
	ls.pdf = 1;
	V3 pos = ld3(l.position);
	switch (l.type) {
		case LIGHT_POINT: {
			V3 d = pos - ctxP;
			ls.L = sampleIlluminant(l.colorSpec, wl, sc.cs) * (l.scale / dot(d, d));
			ls.p = pos;
			break;
		}
		case LIGHT_DIRECTIONAL: {
			V3 wi = mk3(l.rotation[2], l.rotation[5], l.rotation[8]); // rotation * UnitZ
			ls.p  = ctxP + wi * 2 * l.sceneRadius;
			ls.L  = l.scale * sampleIlluminant(l.colorSpec, wl, sc.cs);
			break;
		}
		case LIGHT_SPOT: {
			V3 wLight = normalize(xfPoint(l.inv, ctxP));
			V3 d	  = pos - ctxP;
			ls.L = sampleIlluminant(l.colorSpec, wl, sc.cs) * l.scale * smoothStep(fabsf(wLight.z), l.cosOuter, l.cosInner) / dot(d, d);
			ls.p = pos;
			break;
		}
		default: { // infinite
			V3 wi  = uniformSampleSphere(u0, u1);
			ls.p   = ctxP + wi * 2 * l.sceneRadius;
			ls.L   = infiniteLi(l, wi, wl, sc);
			ls.pdf = kInv4Pi;
		}
	}
	return ls;
}

} // namespace krr
