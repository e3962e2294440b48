// spectrum.cuh -- hero-wavelength sampling and RGB <-> spectrum conversion on the device.
// Follows reference src/render/spectrum.h:51-64 (sampleUniform), :488-529 (toXYZ/toRGB/fromRGB/lum),
// src/render/color.h:84-169 (RGBSigmoidPolynomial, RGBToSpectrumTable::operator()).
#pragma once
#include "krr_math.cuh"

namespace krr {

constexpr float kLambdaMin = 360.f, kLambdaMax = 830.f;
constexpr float kCIEYIntegral = 106.856895f;
constexpr int kSpecRes = 64;

struct ColorSpaceDev {
	const float *cieX, *cieY, *cieZ, *illum; // 471 each (360..830 nm)
	const float *zNodes;					 // 64
	const float *coeffs;					 // [3][64][64][64][3]
	float rgbFromXyz[9];
};

// Per-pixel wavelength state: lambda[0] plus a "secondary wavelengths terminated" flag (sign bit).
// lambda[1..3] and the pdfs are re-derived with the reference's own float operations
// (spectrum.h:55-63), which is exact, so 4 bytes replace the reference's 32-byte SampledWavelengths.
struct Wavelengths {
	float lambda[4];
	float pdf[4];
};
KRR_HD float sampleLambda0(float u) {
	// lerp(lambdaMin, lambdaMax, u) = (1 - u) * min + u * max, individually rounded
	return xadd(xmul(xsub(1.f, u), kLambdaMin), xmul(u, kLambdaMax));
}
KRR_HD Wavelengths expandWavelengths(float packed) {
	Wavelengths w;
	bool terminated = packed < 0;
	w.lambda[0]		= fabsf(packed);
	const float delta = (kLambdaMax - kLambdaMin) / 4;
#pragma unroll
	for (int i = 1; i < 4; i++) {
		w.lambda[i] = xadd(w.lambda[i - 1], delta);
		if (w.lambda[i] > kLambdaMax) w.lambda[i] = xsub(w.lambda[i], kLambdaMax - kLambdaMin);
	}
	const float p = 1.f / (kLambdaMax - kLambdaMin);
	if (terminated) { w.pdf[0] = p / 4; w.pdf[1] = w.pdf[2] = w.pdf[3] = 0; } // terminateSecondary, spectrum.h:70-74
	else w.pdf[0] = w.pdf[1] = w.pdf[2] = w.pdf[3] = p;
	return w;
}

struct SigmoidPoly {
	float c0, c1, c2;
	KRR_HD float operator()(float lambda) const { // color.h:91-93, 105-109
		float x = fmaf(lambda, fmaf(lambda, c0, c1), c2);
		if (isinf(x)) return x > 0 ? 1.f : 0.f;
		return .5f + x / (2 * sqrtf(1 + x * x));
	}
	KRR_HD float maxValue() const { // color.h:95-101
		float result = fmaxf((*this)(kLambdaMin), (*this)(kLambdaMax));
		float lambda = -c1 / (2 * c0);
		if (lambda >= kLambdaMin && lambda <= kLambdaMax) result = fmaxf(result, (*this)(lambda));
		return result;
	}
};

// RGBToSpectrumTable::operator(), color.h:128-169.  Host-callable: constant material / light colours
// are converted once at scene upload (identical float operations, x86 rounding = the oracle's).
KRR_HD SigmoidPoly rgbToCoeffs(const float *zNodes, const float *coeffs, float r, float g, float b) {
	float rgb[3] = {fmaxf(r, 0.f), fmaxf(g, 0.f), fmaxf(b, 0.f)}; // toRGBCoeffs: rgb.cwiseMax(0)
	if (rgb[0] == rgb[1] && rgb[1] == rgb[2]) return SigmoidPoly{0, 0, (rgb[0] - .5f) / sqrtf(rgb[0] * (1 - rgb[0]))};
	int maxc = 0;
	if (rgb[1] > rgb[maxc]) maxc = 1;
	if (rgb[2] > rgb[maxc]) maxc = 2;
	float z = rgb[maxc];
	float x = rgb[(maxc + 1) % 3] * (kSpecRes - 1) / z;
	float y = rgb[(maxc + 2) % 3] * (kSpecRes - 1) / z;
	int xi = min((int) x, kSpecRes - 2), yi = min((int) y, kSpecRes - 2);
	// findInterval(res, zNodes[i] < z), src/util/math_utils.h:94-106
	int size = kSpecRes - 2, first = 1;
	while (size > 0) {
		int half = size >> 1, middle = first + half;
		bool pr = zNodes[middle] < z;
		first = pr ? middle + 1 : first;
		size  = pr ? size - (half + 1) : half;
	}
	int zi	 = max(min(first - 1, kSpecRes - 2), 0);
	float dx = x - xi, dy = y - yi, dz = (z - zNodes[zi]) / (zNodes[zi + 1] - zNodes[zi]);
	float c[3];
#pragma unroll
	for (int i = 0; i < 3; ++i) {
		auto co = [&](int ddx, int ddy, int ddz) {
			return coeffs[((((size_t) maxc * kSpecRes + (zi + ddz)) * kSpecRes + (yi + ddy)) * kSpecRes + (xi + ddx)) * 3 + i];
		};
		c[i] = lerpf(lerpf(lerpf(co(0, 0, 0), co(1, 0, 0), dx), lerpf(co(0, 1, 0), co(1, 1, 0), dx), dy),
					 lerpf(lerpf(co(0, 0, 1), co(1, 0, 1), dx), lerpf(co(0, 1, 1), co(1, 1, 1), dx), dy), dz);
	}
	return SigmoidPoly{c[0], c[1], c[2]};
}

// A colour converted to sigmoid coefficients: RGBBounded uses scale 1; RGBUnbounded / Illuminant use
// scale = 2 * max(rgb), rsp = coeffs(rgb / scale) (spectrum.h:409-412)
struct RgbSpectrum { SigmoidPoly rsp; float scale; };
KRR_HD RgbSpectrum makeBounded(const float *zn, const float *co, float r, float g, float b) {
	return RgbSpectrum{rgbToCoeffs(zn, co, r, g, b), 1.f};
}
KRR_HD RgbSpectrum makeUnbounded(const float *zn, const float *co, float r, float g, float b) {
	float scale = 2 * fmaxf(r, fmaxf(g, b));
	RgbSpectrum s;
	s.scale = scale;
	s.rsp	= scale ? rgbToCoeffs(zn, co, r / scale, g / scale, b / scale) : rgbToCoeffs(zn, co, 0, 0, 0);
	return s;
}
KRR_HD Spec sampleBounded(const RgbSpectrum &s, const Wavelengths &w) {
	return make_float4(s.rsp(w.lambda[0]), s.rsp(w.lambda[1]), s.rsp(w.lambda[2]), s.rsp(w.lambda[3]));
}
KRR_HD Spec sampleUnbounded(const RgbSpectrum &s, const Wavelengths &w) {
	return make_float4(s.scale * s.rsp(w.lambda[0]), s.scale * s.rsp(w.lambda[1]), s.scale * s.rsp(w.lambda[2]), s.scale * s.rsp(w.lambda[3]));
}
// DenselySampledSpectrum::sample, spectrum.h:296-304
KRR_DEV float denseAt(const float *values, float lambda) {
	int offset = (int) lroundf(lambda) - 360;
	return (offset < 0 || offset >= 471) ? 0.f : __ldg(values + offset);
}
KRR_DEV Spec sampleDense(const float *values, const Wavelengths &w) {
	return make_float4(denseAt(values, w.lambda[0]), denseAt(values, w.lambda[1]), denseAt(values, w.lambda[2]), denseAt(values, w.lambda[3]));
}
KRR_DEV Spec sampleIlluminant(const RgbSpectrum &s, const Wavelengths &w, const ColorSpaceDev &cs) {
	return sampleUnbounded(s, w) * sampleDense(cs.illum, w); // spectrum.h:516-518
}
KRR_HD float safeDivf(float x, float y) { return y == 0 ? 0.f : x / y; }
KRR_HD Spec safeDiv(Spec s, const float *pdf) {
	return make_float4(safeDivf(s.x, pdf[0]), safeDivf(s.y, pdf[1]), safeDivf(s.z, pdf[2]), safeDivf(s.w, pdf[3]));
}
// RGBColorSpace::toRGB(SampledSpectrum, lambda), spectrum.h:471-486
KRR_DEV void toRGB(Spec s, const Wavelengths &w, const ColorSpaceDev &cs, float rgb[3]) {
	s = safeDiv(s, w.pdf);
	Spec X = sampleDense(cs.cieX, w), Y = sampleDense(cs.cieY, w), Z = sampleDense(cs.cieZ, w);
	float xyz[3] = {mean(X * s) / kCIEYIntegral, mean(Y * s) / kCIEYIntegral, mean(Z * s) / kCIEYIntegral};
#pragma unroll
	for (int i = 0; i < 3; i++)
		rgb[i] = cs.rgbFromXyz[i * 3] * xyz[0] + cs.rgbFromXyz[i * 3 + 1] * xyz[1] + cs.rgbFromXyz[i * 3 + 2] * xyz[2];
}
// RGBColorSpace::lum(SampledSpectrum), spectrum.h:537-541
KRR_DEV float lum(Spec s, const Wavelengths &w, const ColorSpaceDev &cs) {
	Spec Ys = sampleDense(cs.cieY, w);
	return mean(safeDiv(Ys * s, w.pdf)) / kCIEYIntegral;
}
KRR_HD float luminanceRGB(float r, float g, float b) { return r * 0.299f + g * 0.587f + b * 0.114f; }

} // namespace krr
