// medium.cuh -- participating media on the device.
// Follows reference src/render/media.h (MajorantIterator :41-106, HomogeneousMedium :108-142,
// NanoVDBMedium<float> :145-227 with a DENSE float grid standing in for the NanoVDB tree),
// src/render/media.cpp (majorant grid build :18-75, HG phase function :83-111) and
// src/render/wavefront/wavefront.h:34-78 (sampleT_maj).
#pragma once
#include "bsdf.cuh"
#include "scene.cuh"
#include "spectrum.cuh"

namespace krr {

constexpr int kMajRes = 64; // majorantGridRes, media.h:223

struct MediumPoint { Spec sigma_a, sigma_s, Le; };

// NanoVDBGrid<T>::getValue (util/volume.h:83-87): worldToIndexF + trilinear SampleFromVoxels, background 0 outside the
// grid.  NC = floats per voxel (1: density, 3: RGB albedo grid), `comp` the component
template <int NC> KRR_DEV float gridValue(const MediumRec &m, const float *__restrict__ grid, V3 p, int comp) {
	float ix = (p.x - m.boundsMin[0]) / (m.boundsMax[0] - m.boundsMin[0]) * m.res[0];
	float iy = (p.y - m.boundsMin[1]) / (m.boundsMax[1] - m.boundsMin[1]) * m.res[1];
	float iz = (p.z - m.boundsMin[2]) / (m.boundsMax[2] - m.boundsMin[2]) * m.res[2];
	float fx = floorf(ix), fy = floorf(iy), fz = floorf(iz);
	int x0 = (int) fx, y0 = (int) fy, z0 = (int) fz;
	float wx = ix - fx, wy = iy - fy, wz = iz - fz;
	auto at = [&](int x, int y, int z) -> float {
		if (x < 0 || y < 0 || z < 0 || x >= m.res[0] || y >= m.res[1] || z >= m.res[2]) return 0.f;
		return __ldg(grid + NC * (x + (size_t) m.res[0] * (y + (size_t) m.res[1] * z)) + comp);
	};
	auto mix = [](float a, float b, float t) { return a + t * (b - a); };
	float c00 = mix(at(x0, y0, z0), at(x0 + 1, y0, z0), wx), c10 = mix(at(x0, y0 + 1, z0), at(x0 + 1, y0 + 1, z0), wx);
	float c01 = mix(at(x0, y0, z0 + 1), at(x0 + 1, y0, z0 + 1), wx), c11 = mix(at(x0, y0 + 1, z0 + 1), at(x0 + 1, y0 + 1, z0 + 1), wx);
	return mix(mix(c00, c10, wy), mix(c01, c11, wy), wz);
}
KRR_DEV float gridDensity(const MediumRec &m, const float *__restrict__ pool, V3 p) { return gridValue<1>(m, pool + m.densityOff, p, 0); }

// Medium::samplePoint (media.h:121-126, 161-173).  Constant colours were converted to sigmoid
// coefficients at upload (MediumRec::*Spec), as for materials.
KRR_DEV MediumPoint mediumSamplePoint(const MediumRec &m, const SceneDev &sc, V3 p, const Wavelengths &wl) {
	MediumPoint mp;
	Spec sigma_t = sampleUnbounded(m.sigmaTSpec, wl);
	if (m.type == 0) { // homogeneous: albedo is RGBUnbounded, Le RGBIlluminant
		Spec sigma_s = sigma_t * sampleUnbounded(m.albedoUSpec, wl);
		mp.sigma_a = sigma_t - sigma_s, mp.sigma_s = sigma_s;
		mp.Le = (m.Le[0] != 0 || m.Le[1] != 0 || m.Le[2] != 0) ? sampleIlluminant(m.LeSpec, wl, sc.cs) : sp(0);
		return mp;
	}
	V3 pm = xfPoint(m.inv, p);
	sigma_t = sigma_t * (gridDensity(m, sc.densityPool, pm) * m.scale);
	Spec sigma_s;
	if (m.albedoOff >= 0) { // albedoGrid.getValue(p) -> Spectrum::fromRGB(..., RGBBounded): a table lookup per sample (media.h:168-170)
		const float *ag = sc.densityPool + m.albedoOff;
		const RgbSpectrum a = makeBounded(sc.cs.zNodes, sc.cs.coeffs, gridValue<3>(m, ag, pm, 0), gridValue<3>(m, ag, pm, 1), gridValue<3>(m, ag, pm, 2));
		sigma_s = sigma_t * sampleBounded(a, wl);
	} else sigma_s = sigma_t * sampleBounded(m.albedoBSpec, wl); // grid medium: albedo is RGBBounded
	mp.sigma_a = sigma_t - sigma_s, mp.sigma_s = sigma_s, mp.Le = sp(0);
	return mp;
}

// MajorantIterator, media.h:41-106 (DDA over the 64^3 max-density grid; homogeneous: one segment)
struct MajorantIter {
	Spec sigma_t;
	float tMin, tMax;
	const float *grid; // nullptr = homogeneous
	float nextCrossingT[3], deltaT[3];
	int step[3], voxelLimit[3], voxel[3];

	KRR_DEV void initEmpty() { tMin = 3.402823466e+38f, tMax = 1.175494351e-38f, grid = nullptr; }
	KRR_DEV void init(V3 o, V3 d, float tMin_, float tMax_, Spec sigma_t_, const float *grid_, const float bmin[3], const float bmax[3]) {
		tMin = tMin_, tMax = tMax_, sigma_t = sigma_t_, grid = grid_;
		if (!grid) return;
		float ro[3], rd[3];
		for (int a = 0; a < 3; a++) {
			float diag = bmax[a] - bmin[a];
			ro[a] = (o[a] - bmin[a]) / diag; // bounds.offset(origin)
			rd[a] = d[a] / diag;
		}
		for (int a = 0; a < 3; a++) {
			float gi = ro[a] + rd[a] * tMin;
			voxel[a]  = (int) clampf(gi * kMajRes, 0.f, kMajRes - 1.f);
			deltaT[a] = 1.f / (fabsf(rd[a]) * kMajRes);
			if (rd[a] == -0.f) rd[a] = 0.f;
			if (rd[a] >= 0) {
				float nextVoxelPos = float(voxel[a] + 1) / kMajRes;
				nextCrossingT[a]   = tMin + (nextVoxelPos - gi) / rd[a];
				step[a] = 1, voxelLimit[a] = kMajRes;
			} else {
				float nextVoxelPos = float(voxel[a]) / kMajRes;
				nextCrossingT[a]   = tMin + (nextVoxelPos - gi) / rd[a];
				step[a] = -1, voxelLimit[a] = -1;
			}
		}
	}
	KRR_DEV bool next(float &segMin, float &segMax, Spec &sigma_maj) {
		if (tMin >= tMax) return false;
		if (!grid) {
			segMin = tMin, segMax = tMax, sigma_maj = sigma_t;
			tMin = tMax;
			return true;
		}
		int bits = ((nextCrossingT[0] < nextCrossingT[1]) << 2) + ((nextCrossingT[0] < nextCrossingT[2]) << 1) +
				   ((nextCrossingT[1] < nextCrossingT[2]));
		// cmpToAxis = {2, 1, 2, 1, 2, 2, 0, 0}
		const int stepAxis = (bits == 0 || bits == 2 || bits == 4 || bits == 5) ? 2 : ((bits == 1 || bits == 3) ? 1 : 0);
		float crossing	 = stepAxis == 0 ? nextCrossingT[0] : (stepAxis == 1 ? nextCrossingT[1] : nextCrossingT[2]);
		float tVoxelExit = fminf(tMax, crossing);
		sigma_maj = sigma_t * __ldg(grid + voxel[0] + kMajRes * (voxel[1] + kMajRes * voxel[2]));
		segMin = tMin, segMax = tVoxelExit;
		tMin = tVoxelExit;
		if (crossing > tMax) tMin = tMax;
#pragma unroll
		for (int a = 0; a < 3; a++)
			if (a == stepAxis) {
				voxel[a] += step[a];
				if (voxel[a] == voxelLimit[a]) tMin = tMax;
				nextCrossingT[a] += deltaT[a];
			}
		return true;
	}
};

// Medium::sampleRay (media.h:128-132, 175-187; AABB::intersect krrmath/aabb.h:58-76); o, d world space, |d| = 1
KRR_DEV void mediumSampleRay(const MediumRec &m, const SceneDev &sc, V3 o, V3 d, float raytMax, const Wavelengths &wl, MajorantIter &it) {
	Spec sigma_t = sampleUnbounded(m.sigmaTSpec, wl);
	if (m.type == 0) {
		it.init(o, d, 0.f, raytMax, sigma_t, nullptr, m.boundsMin, m.boundsMax);
		return;
	}
	V3 lo = xfPoint(m.inv, o), ld = xfVector(m.inv, d);
	float t0 = 0, t1 = raytMax;
	for (int i = 0; i < 3; i++) {
		float inv = 1 / ld[i];
		float tn = (m.boundsMin[i] - lo[i]) * inv, tf = (m.boundsMax[i] - lo[i]) * inv;
		if (tn > tf) { float s = tn; tn = tf; tf = s; }
		t0 = tn > t0 ? tn : t0;
		t1 = tf < t1 ? tf : t1;
		if (t0 > t1) { it.initEmpty(); return; }
	}
	it.init(lo, ld, t0, t1, sigma_t * m.scale, sc.densityPool + m.majorantOff, m.boundsMin, m.boundsMax);
}

// sampleT_maj, wavefront.h:34-78.  callback(p, mp, sigma_maj, T_maj) -> keep going?
template <typename F>
KRR_DEV Spec sampleT_maj(const MediumRec &m, const SceneDev &sc, V3 o, V3 d, float tMax, Pcg &rng, const Wavelengths &wl, F callback) {
	tMax *= length(d);
	d = normalize(d);
	Spec T_maj = sp(1);
	MajorantIter it;
	mediumSampleRay(m, sc, o, d, tMax, wl, it);
	float segMin, segMax;
	Spec sigma_maj;
	while (it.next(segMin, segMax, sigma_maj)) {
		if (sigma_maj.x == 0) { // channel = lambda.mainIndex() = 0 in the spectral build
			float dt = segMax - segMin;
			if (isinf(dt)) dt = 3.402823466e+38f;
			T_maj *= expS(sigma_maj * -dt);
			continue;
		}
		float tMin = segMin;
		while (true) {
			float t = tMin + sampleExponential(rng.get1D(), sigma_maj.x);
			if (t < segMax) {
				T_maj *= expS(sigma_maj * -(t - tMin));
				V3 p = o + d * t;
				MediumPoint mp = mediumSamplePoint(m, sc, p, wl);
				if (!callback(p, mp, sigma_maj, T_maj)) return sp(1);
				T_maj = sp(1);
				tMin  = t;
			} else {
				float dt = segMax - tMin;
				if (isinf(dt)) dt = 3.402823466e+38f;
				T_maj *= expS(sigma_maj * -dt);
				break;
			}
		}
	}
	return T_maj;
}

// HGPhaseFunction, media.cpp:83-111
KRR_DEV float hgP(float g_, V3 wo, V3 wi) {
	float g		= clampf(g_, -.99f, .99f);
	float denom = 1 + pow2(g) + 2 * g * dot(wo, wi);
	return kInv4Pi * (1 - pow2(g)) / (denom * safe_sqrt(denom));
}
KRR_DEV V3 perpendicular(V3 u) { // getPerpendicular, util/math_utils.h:120-130
	V3 a = mk3(fabsf(u.x), fabsf(u.y), fabsf(u.z));
	uint32_t uyx = (a.x - a.y) < 0 ? 1 : 0, uzx = (a.x - a.z) < 0 ? 1 : 0, uzy = (a.y - a.z) < 0 ? 1 : 0;
	uint32_t xm = uyx & uzx, ym = (1 ^ xm) & uzy, zm = 1 ^ (xm | ym);
	return normalize(cross(u, mk3((float) xm, (float) ym, (float) zm)));
}
KRR_DEV void hgSample(float g_, V3 wo, float u0, float u1, V3 &wi, float &p, float &pdf) {
	float g = clampf(g_, -.99f, .99f);
	float cosTheta;
	if (fabsf(g) < 1e-3f) cosTheta = 1 - 2 * u0;
	else cosTheta = -1 / (2 * g) * (1 + pow2(g) - pow2((1 - pow2(g)) / (1 + g - 2 * g * u0)));
	float sinTheta = safe_sqrt(1 - pow2(cosTheta));
	float phi	   = k2Pi * u1;
	V3 n = normalize(wo);
	V3 T = perpendicular(n), B = normalize(cross(n, T)); // Frame(n), raytracing.h:63-66
	V3 l = mk3(sinTheta * cosf(phi), sinTheta * sinf(phi), cosTheta);
	wi	 = T * l.x + B * l.y + n * l.z;
	pdf	 = hgP(g_, n, wi);
	p	 = pdf;
}

} // namespace krr
