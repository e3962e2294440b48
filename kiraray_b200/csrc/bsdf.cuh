// bsdf.cuh -- device BSDFs of the wavefront pass, restated from the reference's KRR_CALLABLE
// material headers (each block cites the lines it follows).  All functions work in the local
// shading frame (z = shading normal).  Static (compile-time) dispatch on MaterialType replaces the
// reference's VariantClass (src/render/bsdf.h:19-54): the scatter stage is launched once per
// material bin, so every warp runs one BSDF.
#pragma once
#include "krr_math.cuh"
#include "sampler.cuh"
#include "spectrum.cuh"

namespace krr {

// BSDFType, src/render/materials/bxdf.h:44-66
enum : int {
	BSDF_UNSET = 0, BSDF_NULL = 1, BSDF_REFLECTION = 2, BSDF_TRANSMISSION = 4, BSDF_DIFFUSE = 8, BSDF_GLOSSY = 16, BSDF_SPECULAR = 32,
	BSDF_DIFFUSE_REFLECTION = BSDF_DIFFUSE | BSDF_REFLECTION, BSDF_SPECULAR_REFLECTION = BSDF_SPECULAR | BSDF_REFLECTION,
	BSDF_GLOSSY_REFLECTION = BSDF_GLOSSY | BSDF_REFLECTION, BSDF_SPECULAR_TRANSMISSION = BSDF_SPECULAR | BSDF_TRANSMISSION,
	BSDF_GLOSSY_TRANSMISSION = BSDF_GLOSSY | BSDF_TRANSMISSION,
	BSDF_SMOOTH = BSDF_DIFFUSE | BSDF_GLOSSY, BSDF_DELTA = BSDF_SPECULAR | BSDF_NULL,
};
enum : int { MAT_NULL = 0, MAT_DIFFUSE = 1, MAT_DIELECTRIC = 2, MAT_CONDUCTOR = 3, MAT_DISNEY = 4, MAT_COUNT = 5 };

// BSDFData, src/render/shared.h:34-44
struct ShadingData {
	float IoR;
	Spec diffuse, specular;
	float specularTransmission, roughness, metallic, anisotropic;
	int bsdfType;
	// conductor only: sampled spectral eta / k (flags say whether the material provides them)
	Spec etaSpec, kSpec;
	int hasEta, hasK;
};

// BSDFData::getBsdfType, shared.h:46-73
KRR_HD int getBsdfType(const ShadingData &sd) {
	int type = BSDF_UNSET;
	switch (sd.bsdfType) {
		case MAT_NULL: type = BSDF_NULL; break;
		case MAT_DIFFUSE: type = BSDF_DIFFUSE_REFLECTION; break;
		case MAT_DIELECTRIC: type = (sd.roughness <= 1e-3f ? BSDF_SPECULAR : BSDF_GLOSSY) | BSDF_REFLECTION | BSDF_TRANSMISSION; break;
		case MAT_CONDUCTOR: type = BSDF_REFLECTION | (sd.roughness <= 1e-3f ? BSDF_SPECULAR : BSDF_GLOSSY); break;
		case MAT_DISNEY:
			type = sd.roughness <= 1e-3f ? BSDF_SPECULAR_REFLECTION : BSDF_GLOSSY_REFLECTION;
			if (any(sd.diffuse) && sd.specularTransmission < 1 && sd.metallic < 1) type |= BSDF_DIFFUSE_REFLECTION;
			if (sd.specularTransmission > 0) type |= BSDF_TRANSMISSION;
			break;
	}
	return type;
}

struct BSDFSample {
	Spec f;
	V3 wi;
	float pdf;
	int flags;
};
KRR_HD BSDFSample emptySample() { return BSDFSample{sp(0), mk3(0, 0, 0), 0.f, 0}; }

// ---- render/sampling.h ----
KRR_HD void uniformSampleDisk(float u0, float u1, float &dx, float &dy) { // sampling.h:52-66
	float ox = 2.f * u0 - 1, oy = 2.f * u1 - 1;
	if (ox == 0 && oy == 0) { dx = dy = 0; return; }
	float theta, r;
	if (fabsf(ox) > fabsf(oy)) { r = ox; theta = kPi / 4 * (oy / ox); }
	else { r = oy; theta = kPi / 2 - kPi / 4 * (ox / oy); }
	dx = r * cosf(theta), dy = r * sinf(theta);
}
KRR_HD V3 cosineSampleHemisphere(float u0, float u1) { // sampling.h:75-79
	float dx, dy;
	uniformSampleDisk(u0, u1, dx, dy);
	float z = sqrtf(fmaxf(0.f, 1 - dx * dx - dy * dy));
	return mk3(dx, dy, z);
}
KRR_HD V3 uniformSampleSphere(float u0, float u1) { // sampling.h:31-36
	float z = 1.0f - 2.0f * u0, r = sqrtf(fmaxf(0.0f, 1.0f - z * z)), phi = k2Pi * u1;
	return mk3(r * cosf(phi), r * sinf(phi), z);
}
KRR_HD V3 uniformSampleTriangle(float u0, float u1) { // sampling.h:81-91
	float b0, b1;
	if (u0 < u1) { b0 = u0 / 2; b1 = u1 - b0; }
	else { b1 = u1 / 2; b0 = u0 - b1; }
	return mk3(b0, b1, 1 - b0 - b1);
}
KRR_HD float nextFloatDown(float v) { // math_utils.h:62-71
	if (isinf(v) && v < 0.f) return v;
	if (v == 0.f) v = -0.f;
#ifdef __CUDA_ARCH__
	uint32_t ui = __float_as_uint(v);
	if (v > 0) --ui; else ++ui;
	return __uint_as_float(ui);
#else
	uint32_t ui; memcpy(&ui, &v, 4);
	if (v > 0) --ui; else ++ui;
	memcpy(&v, &ui, 4); return v;
#endif
}
KRR_HD int sampleDiscrete3(float w0, float w1, float w2, float u) { // sampling.h:98-113
	float w[3] = {w0, w1, w2};
	float sumWeights = 0;
	sumWeights += w0; sumWeights += w1; sumWeights += w2;
	float up = u * sumWeights;
	if (up == sumWeights) up = nextFloatDown(up);
	int offset = 0;
	float sum  = 0;
	while (offset < 2 && sum + w[offset] <= up) sum += w[offset++];
	return offset;
}
KRR_HD float sampleExponential(float u, float a) { return -logf(1 - u) / a; }

// ---- materials/matutils.h ----
KRR_HD bool SameHemisphere(V3 w, V3 wp) { return w.z * wp.z > 0; }
KRR_HD float CosTheta(V3 w) { return w.z; }
KRR_HD float Cos2Theta(V3 w) { return w.z * w.z; }
KRR_HD float AbsCosTheta(V3 w) { return fabsf(w.z); }
KRR_HD float Sin2Theta(V3 w) { return fmaxf(0.f, 1.f - Cos2Theta(w)); }
KRR_HD float SinTheta(V3 w) { return sqrtf(Sin2Theta(w)); }
KRR_HD float TanTheta(V3 w) { return SinTheta(w) / CosTheta(w); }
KRR_HD float Tan2Theta(V3 w) { return Sin2Theta(w) / Cos2Theta(w); }
KRR_HD float CosPhi(V3 w) { float s = SinTheta(w); return (s == 0) ? 1 : clampf(w.x / s, -1.f, 1.f); }
KRR_HD float SinPhi(V3 w) { float s = SinTheta(w); return (s == 0) ? 0 : clampf(w.y / s, -1.f, 1.f); }
KRR_HD float Cos2Phi(V3 w) { return CosPhi(w) * CosPhi(w); }
KRR_HD float Sin2Phi(V3 w) { return SinPhi(w) * SinPhi(w); }
KRR_HD float AbsDot(V3 a, V3 b) { return fabsf(dot(a, b)); }
KRR_HD V3 Reflect(V3 wo, V3 n) { return -wo + 2 * dot(wo, n) * n; }
KRR_HD V3 FaceForward(V3 w, V3 wp) { return dot(w, wp) > 0 ? w : -w; }
// Refract (absolute eta, optional etap), matutils.h:84-98
KRR_HD bool Refract(V3 wi, V3 n, float eta, float *etap, V3 *wt) {
	float cosThetaI = dot(n, wi);
	if (wi.z < 0) eta = 1 / eta;
	float sin2ThetaI = fmaxf(0.f, 1 - pow2(cosThetaI));
	float sin2ThetaT = sin2ThetaI / pow2(eta);
	if (sin2ThetaT >= 1) return false;
	float cosThetaT = sqrtf(1 - sin2ThetaT);
	*wt = -wi / eta + (cosThetaI / eta - cosThetaT) * n;
	if (etap) *etap = eta;
	return true;
}

// ---- materials/fresnel.h ----
KRR_HD float FrDielectric(float cosTheta_i, float eta) { // fresnel.h:22-41
	cosTheta_i = clampf(cosTheta_i, -1.f, 1.f);
	if (cosTheta_i < 0) { eta = 1 / eta; cosTheta_i = -cosTheta_i; }
	float sin2Theta_i = 1 - pow2(cosTheta_i);
	float sin2Theta_t = sin2Theta_i / pow2(eta);
	if (sin2Theta_t >= 1) return 1.f;
	float cosTheta_t = sqrtf(fmaxf(1 - sin2Theta_t, 0.f));
	float r_parl = (eta * cosTheta_i - cosTheta_t) / (eta * cosTheta_i + cosTheta_t);
	float r_perp = (cosTheta_i - eta * cosTheta_t) / (cosTheta_i + eta * cosTheta_t);
	return (pow2(r_parl) + pow2(r_perp)) / 2;
}
struct Cpx { float re, im; };
KRR_HD Cpx cmul(Cpx a, Cpx b) { return Cpx{a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re}; }
KRR_HD Cpx cdiv(Cpx a, Cpx z) { float s = 1 / (z.re * z.re + z.im * z.im); return Cpx{s * (a.re * z.re + a.im * z.im), s * (a.im * z.re - a.re * z.im)}; }
KRR_HD Cpx csub(Cpx a, Cpx b) { return Cpx{a.re - b.re, a.im - b.im}; }
KRR_HD Cpx cadd(Cpx a, Cpx b) { return Cpx{a.re + b.re, a.im + b.im}; }
KRR_HD float cnorm(Cpx a) { return a.re * a.re + a.im * a.im; }
KRR_HD Cpx csqrt(Cpx a) { // krrmath/complex.h:56-61
	float n = sqrtf(cnorm(a)), t1 = sqrtf(.5f * (n + fabsf(a.re))), t2 = .5f * a.im / t1;
	if (n == 0) return Cpx{0, 0};
	if (a.re >= 0) return Cpx{t1, t2};
	return Cpx{fabsf(t2), copysignf(t1, a.im)};
}
KRR_HD float FrComplex(float cosTheta_i, Cpx eta) { // fresnel.h:43-55
	cosTheta_i = clampf(cosTheta_i, 0.f, 1.f);
	float sin2Theta_i = 1 - pow2(cosTheta_i);
	Cpx sin2Theta_t = cdiv(Cpx{sin2Theta_i, 0}, cmul(eta, eta));
	Cpx cosTheta_t	= csqrt(csub(Cpx{1, 0}, sin2Theta_t));
	Cpx ci{cosTheta_i, 0};
	Cpx r_parl = cdiv(csub(cmul(eta, ci), cosTheta_t), cadd(cmul(eta, ci), cosTheta_t));
	Cpx r_perp = cdiv(csub(ci, cmul(eta, cosTheta_t)), cadd(ci, cmul(eta, cosTheta_t)));
	return (cnorm(r_parl) + cnorm(r_perp)) / 2;
}
KRR_HD Spec FrComplexS(float c, Spec eta, Spec k) {
	return make_float4(FrComplex(c, Cpx{eta.x, k.x}), FrComplex(c, Cpx{eta.y, k.y}), FrComplex(c, Cpx{eta.z, k.z}), FrComplex(c, Cpx{eta.w, k.w}));
}
KRR_HD Spec FrSchlickS(Spec f0, Spec f90, float cosTheta) { return f0 + (f90 - f0) * pow5(clampf(1 - fabsf(cosTheta), 0.f, 1.f)); }
// DisneyFresnel, fresnel.h:74-76 (Eigen lerp: a * (1 - t) + b * t)
KRR_HD Spec DisneyFresnel(Spec R0, float metallic, float eta, float cosI) {
	return lerpS(sp(FrDielectric(cosI, eta)), FrSchlickS(R0, sp(1), cosI), metallic);
}

// ---- materials/microfacet.h: GGX / Disney distributions ----
template <bool SEPARABLE_G> struct GGX {
	float alphax, alphay;
	KRR_HD void set(float ax, float ay) { alphax = fmaxf(1e-3f, ax); alphay = fmaxf(1e-3f, ay); }
	KRR_HD bool isDelta() const { return fmaxf(alphax, alphay) <= 1e-3f; }
	KRR_HD float Lambda(V3 w) const { // microfacet.h:59-67
		float absTanTheta = fabsf(TanTheta(w));
		if (isinf(absTanTheta)) return 0.;
		float alpha = sqrtf(Cos2Phi(w) * alphax * alphax + Sin2Phi(w) * alphay * alphay);
		float a2t2	= pow2(alpha * absTanTheta);
		return (-1 + sqrtf(1.f + a2t2)) / 2;
	}
	KRR_HD float G1(V3 w) const { return 1 / (1 + Lambda(w)); }
	KRR_HD float G(V3 wo, V3 wi) const { return SEPARABLE_G ? G1(wo) * G1(wi) : 1 / (1 + Lambda(wo) + Lambda(wi)); }
	KRR_HD float D(V3 wh) const { // microfacet.h:43-49
		float tan2Theta = Tan2Theta(wh);
		if (isinf(tan2Theta)) return 0.;
		const float cos4Theta = pow2(Cos2Theta(wh));
		float e = (pow2(CosPhi(wh) / alphax) + pow2(SinPhi(wh) / alphay)) * tan2Theta;
		return 1 / (kPi * alphax * alphay * cos4Theta * pow2(1 + e));
	}
	KRR_HD float Pdf(V3 wo, V3 wh) const { return D(wh) * G1(wo) * fabsf(dot(wo, wh)) / AbsCosTheta(wo); }
	// legacy sampling path (KRR_GGX_SAMPLE_LEGACY), microfacet.h:86-158
	KRR_HD static void Sample11(float cosTheta, float U1, float U2, float *slope_x, float *slope_y) {
		if (cosTheta > .9999f) {
			float r = sqrtf(U1 / (1 - U1)), phi = k2Pi * U2;
			*slope_x = r * cosf(phi), *slope_y = r * sinf(phi);
			return;
		}
		float sinTheta = sqrtf(fmaxf(0.f, 1.f - cosTheta * cosTheta));
		float tanTheta = sinTheta / cosTheta;
		float a	 = 1 / tanTheta;
		float G1 = 2 / (1 + sqrtf(1.f + 1.f / (a * a)));
		float A	 = 2 * U1 / G1 - 1;
		float tmp = 1.f / (A * A - 1.f);
		if (tmp > 1e10f) tmp = 1e10f;
		float B = tanTheta;
		float D = sqrtf(fmaxf(B * B * tmp * tmp - (A * A - B * B) * tmp, 0.f));
		float s1 = B * tmp - D, s2 = B * tmp + D;
		*slope_x = (A < 0 || s2 > 1.f / tanTheta) ? s1 : s2;
		float S;
		if (U2 > 0.5f) { S = 1.f; U2 = 2.f * (U2 - .5f); }
		else { S = -1.f; U2 = 2.f * (.5f - U2); }
		float z = (U2 * (U2 * (U2 * 0.27385f - 0.73369f) + 0.46341f)) / (U2 * (U2 * (U2 * 0.093073f + 0.309420f) - 1.000000f) + 0.597999f);
		*slope_y = S * z * sqrtf(1.f + *slope_x * *slope_x);
	}
	KRR_HD V3 Sample(V3 wo, float u0, float u1) const { // microfacet.h:160-187
		bool flip = wo.z < 0;
		V3 wi	  = flip ? -wo : wo;
		V3 ws	  = normalize(mk3(alphax * wi.x, alphay * wi.y, wi.z));
		float sx, sy;
		Sample11(CosTheta(ws), u0, u1, &sx, &sy);
		float tmp = CosPhi(ws) * sx - SinPhi(ws) * sy;
		sy = SinPhi(ws) * sx + CosPhi(ws) * sy;
		sx = tmp;
		sx = alphax * sx, sy = alphay * sy;
		V3 wh = normalize(mk3(-sx, -sy, 1.f));
		return flip ? -wh : wh;
	}
};

// MicrofacetBrdf with KRR_USE_DISNEY, microfacet.h:189-282
struct MicrofacetBrdf {
	Spec R, disneyR0;
	float eta, metallic;
	GGX<true> dist;
	KRR_HD Spec Fr(V3 wo, V3 wh) const { return DisneyFresnel(disneyR0, metallic, eta, dot(wo, wh)) * R; }
	KRR_HD Spec f(V3 wo, V3 wi) const {
		if (dist.isDelta()) return sp(0);
		if (!SameHemisphere(wi, wo)) return sp(0);
		float cosThetaO = AbsCosTheta(wo), cosThetaI = AbsCosTheta(wi);
		V3 wh = wi + wo;
		if (cosThetaI == 0 || cosThetaO == 0) return sp(0);
		if (!anyNonZero(wh)) return sp(0);
		wh = normalize(wh);
		Spec F = Fr(wo, wh);
		return dist.D(wh) * dist.G(wo, wi) * F / (4 * cosThetaI * cosThetaO);
	}
	KRR_HD float pdf(V3 wo, V3 wi) const {
		if (dist.isDelta()) return 0;
		if (!SameHemisphere(wo, wi)) return 0;
		V3 wh = normalize(wo + wi);
		return dist.Pdf(wo, wh) / (4 * dot(wo, wh));
	}
	KRR_HD BSDFSample sample(V3 wo, Pcg &sg) const {
		BSDFSample s = emptySample();
		float u0 = sg.get1D(), u1 = sg.get1D(); // drawn before the early-outs (microfacet.h:222)
		if (wo.z == 0) return s;
		if (dist.isDelta()) {
			s.f = Fr(wo, mk3(0, 0, 1)) / AbsCosTheta(wo);
			s.wi = mk3(-wo.x, -wo.y, wo.z), s.pdf = 1, s.flags = BSDF_SPECULAR_REFLECTION;
			return s;
		}
		V3 wh = dist.Sample(wo, u0, u1);
		if (dot(wo, wh) < 0) return s;
		V3 wi = Reflect(wo, wh);
		if (!SameHemisphere(wo, wi)) return s;
		s.f = f(wo, wi), s.wi = wi;
		s.pdf	= dist.Pdf(wo, wh) / (4 * dot(wo, wh));
		s.flags = BSDF_GLOSSY_REFLECTION;
		return s;
	}
};

// MicrofacetBtdf with KRR_USE_DISNEY, microfacet.h:285-395
struct MicrofacetBtdf {
	Spec T, disneyR0;
	float etaT, metallic;
	GGX<true> dist;
	KRR_HD Spec Fr(V3 wo, V3 wh) const { return DisneyFresnel(disneyR0, metallic, etaT, dot(wo, wh)); }
	KRR_HD Spec f(V3 wo, V3 wi) const {
		if (dist.isDelta()) return sp(0);
		if (SameHemisphere(wo, wi)) return sp(0);
		float cosThetaO = wo.z, cosThetaI = wi.z;
		if (cosThetaI == 0 || cosThetaO == 0) return sp(0);
		float eta = CosTheta(wo) > 0 ? etaT : 1 / etaT;
		V3 wh = normalize(wo + wi * eta);
		if (wh.z < 0) wh = -wh;
		if (dot(wo, wh) * dot(wi, wh) > 0) return sp(0);
		Spec F = Fr(wo, wh);
		float sqrtDenom = dot(wo, wh) + eta * dot(wi, wh);
		Spec ft = (sp(1) - F) * T * fabsf(dist.D(wh) * dist.G(wo, wi) * AbsDot(wi, wh) * AbsDot(wo, wh) / (cosThetaI * cosThetaO * pow2(sqrtDenom)));
		ft /= pow2(eta); // TransportMode::Radiance
		return ft;
	}
	KRR_HD float pdf(V3 wo, V3 wi) const {
		if (dist.isDelta()) return 0;
		if (SameHemisphere(wo, wi)) return 0;
		float eta = CosTheta(wo) > 0 ? etaT : 1 / etaT;
		V3 wh = normalize(wo + wi * eta);
		if (dot(wo, wh) * dot(wi, wh) > 0) return 0;
		float sqrtDenom = dot(wo, wh) + eta * dot(wi, wh);
		float dwh_dwi	= fabsf((eta * eta * dot(wi, wh)) / (sqrtDenom * sqrtDenom));
		return dist.Pdf(wo, wh) * dwh_dwi;
	}
	KRR_HD BSDFSample sample(V3 wo, Pcg &sg) const {
		BSDFSample s = emptySample();
		if (wo.z == 0) return s;
		float eta;
		if (dist.isDelta()) {
			V3 wi, wh = mk3(0, 0, copysignf(1, wo.z));
			if (!Refract(wo, wh, etaT, &eta, &wi)) return s;
			Spec ft = (sp(1) - Fr(wo, wh)) * T / AbsCosTheta(wi);
			ft /= pow2(eta);
			s.f = ft, s.wi = wi, s.pdf = 1, s.flags = BSDF_SPECULAR_TRANSMISSION;
			return s;
		}
		float u0 = sg.get1D(), u1 = sg.get1D();
		V3 wh = dist.Sample(wo, u0, u1);
		if (!Refract(wo, wh, etaT, &eta, &s.wi)) return emptySample();
		s.pdf = pdf(wo, s.wi), s.f = f(wo, s.wi), s.flags = BSDF_GLOSSY_TRANSMISSION;
		return s;
	}
};

// ---- materials/disney.h ----
KRR_HD float SchlickR0FromEta(float eta) { return pow2(eta - 1) / pow2(eta + 1); }
KRR_HD float SchlickWeight(float cosTheta) { return pow5(clampf(1.f - cosTheta, 0.f, 1.f)); }

template <int MT> struct Bsdf;

// what setup() needs besides the shading data
struct BsdfSetupCtx {
	V3 woWorld;				 // DisneyBsdf::setup reads AbsCosTheta(intr.wo) of the WORLD-space wo (disney.h:263)
	const Wavelengths *w;
	const ColorSpaceDev *cs;
};

template <> struct Bsdf<MAT_DISNEY> { // disney.h:193-370
	Spec diffR;	   // DisneyDiffuse::R == DisneyRetro::R = diffuseWeight * c
	float roughness;
	MicrofacetBrdf brdf;
	MicrofacetBtdf btdf;
	float pDiffuse, pSpecTrans, pSpecRefl;

	KRR_DEV void setup(const ShadingData &sd, const BsdfSetupCtx &ctx) {
		Spec c = sd.diffuse;
		float metallicWeight = sd.metallic, e = sd.IoR, strans = sd.specularTransmission;
		float diffuseWeight = (1 - metallicWeight) * (1 - strans);
		roughness = sd.roughness;
		float l	  = lum(c, *ctx.w, *ctx.cs);
		Spec Ctint = sp(1);
		if (l > 0) Ctint = c / l;
		bool hasDiffuse = diffuseWeight > 0;
		diffR = hasDiffuse ? diffuseWeight * c : sp(0);
		float aspect = sqrtf(1 - sd.anisotropic * .9f);
		float ax = fmaxf(.001f, pow2(roughness) / aspect), ay = fmaxf(.001f, pow2(roughness) * aspect);
		Spec Cspec0 = sd.specular;
		if (!any(sd.specular)) Cspec0 = lerpS(SchlickR0FromEta(e) * lerpS(sp(1), Ctint, 1.f), c, metallicWeight);
		brdf.R = sp(1), brdf.eta = e, brdf.dist.set(ax, ay), brdf.disneyR0 = Cspec0, brdf.metallic = metallicWeight;
		bool hasTrans = strans > 0;
		btdf.T = sp(0), btdf.etaT = 1.5f, btdf.dist.set(ax, ay), btdf.disneyR0 = Cspec0, btdf.metallic = metallicWeight;
		if (hasTrans) { btdf.T = strans * sqrtS(c); btdf.etaT = fmaxf(1.01f, e); }
		float approxFresnel = lum(DisneyFresnel(Cspec0, metallicWeight, e, fabsf(ctx.woWorld.z)), *ctx.w, *ctx.cs);
		pDiffuse   = hasDiffuse ? mean(sd.diffuse) * (1 - metallicWeight) * (1 - sd.specularTransmission) : 0;
		pSpecRefl  = mean(lerpS(sd.specular, sp(1), approxFresnel)) * (1 - sd.specularTransmission * (1 - metallicWeight));
		pSpecTrans = hasTrans ? (1 - approxFresnel) * (1 - metallicWeight) * sd.specularTransmission : 0;
		float totalWt = pDiffuse + pSpecRefl + pSpecTrans;
		if (totalWt > 0) pDiffuse /= totalWt, pSpecRefl /= totalWt, pSpecTrans /= totalWt;
	}
	KRR_DEV Spec diffuseF(V3 wo, V3 wi) const { // DisneyDiffuse::f + DisneyRetro::f, disney.h:34-47, 77-98
		Spec val = sp(0);
		float Fo = SchlickWeight(AbsCosTheta(wo)), Fi = SchlickWeight(AbsCosTheta(wi));
		if (SameHemisphere(wo, wi)) val += diffR * kInvPi * (1 - Fo / 2) * (1 - Fi / 2);
		V3 wh = wi + wo;
		if (!(wh.x == 0 && wh.y == 0 && wh.z == 0)) {
			wh = normalize(wh);
			float cosThetaD = dot(wi, wh);
			float Rr = 2 * roughness * cosThetaD * cosThetaD;
			val += diffR * kInvPi * Rr * (Fo + Fi + Fo * Fi * (Rr - 1));
		}
		return val;
	}
	KRR_DEV Spec f(V3 wo, V3 wi) const {
		Spec val = sp(0);
		bool reflect = SameHemisphere(wo, wi);
		if (pDiffuse > 0 && reflect) val += diffuseF(wo, wi);
		if (pSpecRefl > 0 && reflect) val += brdf.f(wo, wi);
		if (pSpecTrans > 0 && !reflect) val += btdf.f(wo, wi);
		return val;
	}
	KRR_DEV float pdf(V3 wo, V3 wi) const {
		float val = 0;
		bool reflect = SameHemisphere(wo, wi);
		if (pDiffuse > 0 && reflect) val += pDiffuse * AbsCosTheta(wi) * kInvPi;
		if (pSpecRefl > 0 && reflect) val += pSpecRefl * brdf.pdf(wo, wi);
		if (pSpecTrans > 0 && !reflect) val += pSpecTrans * btdf.pdf(wo, wi);
		return val;
	}
	// f(wo, wi) and pdf(wo, wi) in one pass: the same expressions as f() / pdf() above, evaluated in the
	// same order (so the values are bit-identical), but the half vector, D(wh), G1(wo) and the hemisphere
	// tests of the reflection lobe are computed once instead of once per function.  The scatter stage
	// calls this from ONE loop for both directions it has to evaluate (light sample, cosine sample), so
	// the code exists once in the kernel -- the stage is instruction-fetch bound.
	KRR_DEV void eval(V3 wo, V3 wi, Spec &fOut, float &pdfOut) const {
		Spec val   = sp(0);
		float pval = 0;
		const bool reflect = SameHemisphere(wo, wi);
		if (pDiffuse > 0 && reflect) {
			val += diffuseF(wo, wi);
			pval += pDiffuse * AbsCosTheta(wi) * kInvPi;
		}
		if (pSpecRefl > 0 && reflect) {
			const GGX<true> &dist = brdf.dist;
			if (!dist.isDelta()) {
				float cosThetaO = AbsCosTheta(wo), cosThetaI = AbsCosTheta(wi);
				V3 whRaw = wi + wo;
				V3 wh	 = normalize(whRaw);
				const float D = dist.D(wh), G1o = dist.G1(wo);
				if (!(cosThetaI == 0 || cosThetaO == 0) && anyNonZero(whRaw)) {
					Spec F = brdf.Fr(wo, wh);
					val += D * (G1o * dist.G1(wi)) * F / (4 * cosThetaI * cosThetaO);
				}
				pval += pSpecRefl * ((D * G1o * fabsf(dot(wo, wh)) / AbsCosTheta(wo)) / (4 * dot(wo, wh)));
			}
		}
		if (pSpecTrans > 0 && !reflect) {
			val += btdf.f(wo, wi);
			pval += pSpecTrans * btdf.pdf(wo, wi);
		}
		fOut = val, pdfOut = pval;
	}
	// first draw of sample(): which lobe
	KRR_DEV int pickLobe(Pcg &sg) const { return sampleDiscrete3(pDiffuse, pSpecRefl, pSpecTrans, sg.get1D()); }
	// sample() for the specular lobes (comp 1, 2); the diffuse lobe (comp 0) is a cosine sample + eval()
	KRR_DEV BSDFSample sampleSpecular(int comp, V3 wo, Pcg &sg) const {
		BSDFSample s;
		if (comp == 1) {
			s = brdf.sample(wo, sg);
			s.pdf *= pSpecRefl;
			if (pDiffuse && (s.flags & BSDF_SMOOTH)) {
				s.f += diffuseF(wo, s.wi);
				s.pdf += pDiffuse * AbsCosTheta(s.wi) * kInvPi;
			}
		} else {
			s = btdf.sample(wo, sg);
			s.pdf *= pSpecTrans;
		}
		return s;
	}
	KRR_DEV BSDFSample sample(V3 wo, Pcg &sg) const {
		BSDFSample s = emptySample();
		int comp = sampleDiscrete3(pDiffuse, pSpecRefl, pSpecTrans, sg.get1D());
		if (comp == 0) {
			float u0 = sg.get1D(), u1 = sg.get1D();
			V3 wi = cosineSampleHemisphere(u0, u1);
			if (wo.z < 0) wi.z *= -1;
			s.pdf = pdf(wo, wi), s.f = f(wo, wi), s.wi = wi, s.flags = BSDF_DIFFUSE_REFLECTION;
		} else if (comp == 1) {
			s = brdf.sample(wo, sg);
			s.pdf *= pSpecRefl;
			if (pDiffuse && (s.flags & BSDF_SMOOTH)) {
				s.f += diffuseF(wo, s.wi);
				s.pdf += pDiffuse * AbsCosTheta(s.wi) * kInvPi;
			}
		} else {
			s = btdf.sample(wo, sg);
			s.pdf *= pSpecTrans;
		}
		return s;
	}
};

template <> struct Bsdf<MAT_DIFFUSE> { // diffuse.h:18-58
	Spec diffuse;
	KRR_DEV void setup(const ShadingData &sd, const BsdfSetupCtx &) { diffuse = sd.diffuse; }
	KRR_DEV Spec f(V3 wo, V3 wi) const { return SameHemisphere(wo, wi) ? diffuse * kInvPi : sp(0); }
	KRR_DEV float pdf(V3 wo, V3 wi) const { return SameHemisphere(wo, wi) ? fabsf(wi.z) * kInvPi : 0.f; }
	KRR_DEV BSDFSample sample(V3 wo, Pcg &sg) const {
		BSDFSample s;
		float u0 = sg.get1D(), u1 = sg.get1D();
		V3 wi = cosineSampleHemisphere(u0, u1);
		s.wi  = mk3(wi.x, wi.y, wi.z * wo.z > 0 ? wi.z : -wi.z); // ToSameHemisphere(wi, wo)
		s.f = f(wo, s.wi), s.pdf = pdf(wo, s.wi), s.flags = BSDF_DIFFUSE_REFLECTION;
		return s;
	}
};

template <> struct Bsdf<MAT_NULL> { // null.h:18-46
	KRR_DEV void setup(const ShadingData &, const BsdfSetupCtx &) {}
	KRR_DEV Spec f(V3, V3) const { return sp(0); }
	KRR_DEV float pdf(V3, V3) const { return 0; }
	KRR_DEV BSDFSample sample(V3 wo, Pcg &) const { return BSDFSample{sp(1) / AbsCosTheta(wo), -wo, 1.f, BSDF_NULL}; }
};

template <> struct Bsdf<MAT_DIELECTRIC> { // dielectric.h:75-261
	float eta;
	Spec baseColor;
	GGX<false> dist;
	KRR_DEV void setup(const ShadingData &sd, const BsdfSetupCtx &) {
		baseColor = sd.diffuse * sd.specularTransmission;
		eta		  = sd.IoR;
		float alpha = pow2(sd.roughness);
		dist.set(alpha, alpha);
	}
	KRR_DEV Spec f(V3 wo, V3 wi) const {
		if (eta == 1 || dist.isDelta()) return sp(0);
		float cosTheta_o = CosTheta(wo), cosTheta_i = CosTheta(wi);
		bool reflect = cosTheta_i * cosTheta_o > 0;
		float etap = 1;
		if (!reflect) etap = cosTheta_o > 0 ? eta : (1 / eta);
		V3 wm = wi * etap + wo;
		if (cosTheta_i == 0 || cosTheta_o == 0 || !anyNonZero(wm)) return sp(0);
		wm = FaceForward(normalize(wm), mk3(0, 0, 1));
		if (dot(wm, wi) * cosTheta_i < 0 || dot(wm, wo) * cosTheta_o < 0) return sp(0);
		float F = FrDielectric(copysignf(dot(wo, wm), wo.z), eta);
		Spec R = baseColor * F, T = baseColor * (1 - F);
		if (reflect) return dist.D(wm) * dist.G(wo, wi) * R / fabsf(4 * cosTheta_i * cosTheta_o);
		float denom = pow2(dot(wi, wm) + dot(wo, wm) / etap) * cosTheta_i * cosTheta_o;
		Spec ft = T * dist.D(wm) * dist.G(wo, wi) * fabsf(dot(wi, wm) * dot(wo, wm) / denom);
		ft /= pow2(etap);
		return ft;
	}
	KRR_DEV float pdf(V3 wo, V3 wi) const {
		if (eta == 1 || dist.isDelta()) return 0;
		float cosTheta_o = CosTheta(wo), cosTheta_i = CosTheta(wi);
		bool reflect = cosTheta_i * cosTheta_o > 0;
		float etap = 1;
		if (!reflect) etap = cosTheta_o > 0 ? eta : (1 / eta);
		V3 wm = wi * etap + wo;
		if (cosTheta_i == 0 || cosTheta_o == 0 || dot(wm, wm) == 0) return 0;
		wm = FaceForward(normalize(wm), mk3(0, 0, 1));
		if (dot(wm, wi) * cosTheta_i < 0 || dot(wm, wo) * cosTheta_o < 0) return 0;
		float F = FrDielectric(CosTheta(wo), eta); // (sic) pdf uses cos(wo), dielectric.h:233
		Spec R = baseColor * F, T = baseColor * (1 - F);
		float pr = mean(R), pt = mean(T);
		if (pr == 0 && pt == 0) return 0;
		if (reflect) return dist.Pdf(wo, wm) / (4 * fabsf(dot(wo, wm))) * pr / (pr + pt);
		float denom = pow2(dot(wi, wm) + dot(wo, wm) / etap);
		float dwm_dwi = fabsf(dot(wi, wm)) / denom;
		return dist.Pdf(wo, wm) * dwm_dwi * pt / (pr + pt);
	}
	KRR_DEV BSDFSample sample(V3 wo, Pcg &sg) const {
		if (eta == 1 || dist.isDelta()) {
			float F = FrDielectric(CosTheta(wo), eta);
			Spec R = baseColor * F, T = baseColor * (1 - F);
			float pr = mean(R), pt = mean(T);
			if (pr == 0 && pt == 0) return emptySample();
			if (sg.get1D() < pr / (pr + pt)) {
				V3 wi = mk3(-wo.x, -wo.y, wo.z);
				return BSDFSample{R / AbsCosTheta(wi), wi, pr / (pr + pt), BSDF_SPECULAR_REFLECTION};
			}
			V3 wi; float etap;
			if (!Refract(wo, mk3(0, 0, copysignf(1, wo.z)), eta, &etap, &wi)) return emptySample();
			Spec ft = T / AbsCosTheta(wi);
			ft /= pow2(etap);
			return BSDFSample{ft, wi, pt / (pr + pt), BSDF_SPECULAR_TRANSMISSION};
		}
		float u0 = sg.get1D(), u1 = sg.get1D();
		V3 wm	= dist.Sample(wo, u0, u1);
		float F = FrDielectric(copysignf(dot(wo, wm), wo.z), eta);
		Spec R = baseColor * F, T = baseColor * (1 - F);
		float pr = mean(R), pt = mean(T);
		if (pr == 0 && pt == 0) return emptySample();
		if (sg.get1D() < pr / (pr + pt)) {
			V3 wi = Reflect(wo, wm);
			if (!SameHemisphere(wo, wi)) return emptySample();
			float pdf = dist.Pdf(wo, wm) / (4 * fabsf(dot(wo, wm))) * pr / (pr + pt);
			Spec f = dist.D(wm) * dist.G(wo, wi) * R / (4 * CosTheta(wi) * CosTheta(wo));
			return BSDFSample{f, wi, pdf, BSDF_GLOSSY_REFLECTION};
		}
		float etap; V3 wi = mk3(0, 0, 0);
		bool tir = !Refract(wo, wm, eta, &etap, &wi);
		if (SameHemisphere(wo, wi) || wi.z == 0 || tir) return emptySample();
		float denom = pow2(dot(wi, wm) + dot(wo, wm) / etap);
		float dwm_dwi = fabsf(dot(wi, wm)) / denom;
		float pdf = dist.Pdf(wo, wm) * dwm_dwi * pt / (pr + pt);
		Spec ft = T * dist.D(wm) * dist.G(wo, wi) * fabsf(dot(wi, wm) * dot(wo, wm) / (CosTheta(wi) * CosTheta(wo) * denom));
		ft /= pow2(etap);
		return BSDFSample{ft, wi, pdf, BSDF_GLOSSY_TRANSMISSION};
	}
};

template <> struct Bsdf<MAT_CONDUCTOR> { // conductor.h:6-99
	Spec eta, k;
	GGX<false> dist;
	KRR_DEV void setup(const ShadingData &sd, const BsdfSetupCtx &) {
		Spec reflectance = cwiseMin(sd.diffuse, 0.9999f);
		float aspect = sqrtf(1 - sd.anisotropic * .9f);
		dist.set(fmaxf(.001f, pow2(sd.roughness) / aspect), fmaxf(.001f, pow2(sd.roughness) * aspect));
		eta = sd.hasEta ? sd.etaSpec : sp(sd.IoR);
		// (sic) conductor.h:14,29: `k_spec` is a default-constructed local, so spectralK is never used
		k = 2 * sqrtS(reflectance) / sqrtS(cwiseMax(sp(1) - reflectance, 0.f));
	}
	KRR_DEV Spec f(V3 wo, V3 wi) const {
		if (!SameHemisphere(wo, wi)) return sp(0);
		if (dist.isDelta()) return sp(0);
		float cosTheta_o = AbsCosTheta(wo), cosTheta_i = AbsCosTheta(wi);
		if (cosTheta_i == 0 || cosTheta_o == 0) return sp(0);
		V3 wm = wi + wo;
		if (dot(wm, wm) == 0) return sp(0);
		wm = normalize(wm);
		Spec F = FrComplexS(AbsDot(wo, wm), eta, k);
		return dist.D(wm) * F * dist.G(wo, wi) / (4 * cosTheta_i * cosTheta_o);
	}
	KRR_DEV float pdf(V3 wo, V3 wi) const {
		if (!SameHemisphere(wo, wi)) return 0;
		if (dist.isDelta()) return 0;
		V3 wm = wo + wi;
		if (!anyNonZero(wm)) return 0;
		wm = FaceForward(normalize(wm), mk3(0, 0, 1));
		return dist.Pdf(wo, wm) / (4 * AbsDot(wo, wm));
	}
	KRR_DEV BSDFSample sample(V3 wo, Pcg &sg) const {
		if (dist.isDelta()) {
			V3 wi = mk3(-wo.x, -wo.y, wo.z);
			return BSDFSample{FrComplexS(AbsCosTheta(wi), eta, k) / AbsCosTheta(wi), wi, 1.f, BSDF_SPECULAR_REFLECTION};
		}
		if (wo.z == 0) return emptySample();
		float u0 = sg.get1D(), u1 = sg.get1D();
		V3 wm = dist.Sample(wo, u0, u1);
		V3 wi = Reflect(wo, wm);
		if (!SameHemisphere(wo, wi)) return emptySample();
		float pdf = dist.Pdf(wo, wm) / (4 * AbsDot(wo, wm));
		float cosTheta_o = AbsCosTheta(wo), cosTheta_i = AbsCosTheta(wi);
		if (cosTheta_i == 0 || cosTheta_o == 0) return emptySample();
		Spec F = FrComplexS(AbsDot(wo, wm), eta, k);
		Spec f = dist.D(wm) * F * dist.G(wo, wi) / (4 * cosTheta_i * cosTheta_o);
		return BSDFSample{f, wi, pdf, BSDF_GLOSSY_REFLECTION};
	}
};

} // namespace krr
