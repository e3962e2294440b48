/* backend_ref.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Implements oracle/oracle_leaf.h by calling the REFERENCE's own KRR_CALLABLE classes, compiled
 * host-side with g++ from the headers where they lie under /root/reference/src (include paths
 * only; nothing is copied into this repository).  Built by oracle/build_oracle.py into
 * oracle/_ref/ (git-ignored).  See SURVEY.md section 8(c) for the compat layer this needs.
 */
#include <cuda_runtime.h>
#include <cstdlib>
#include <cstring>
#include <cstdio>

/* ---- fake CUDA runtime (host malloc/memcpy) so the reference's containers work on the CPU ---- */
extern "C" {
cudaError_t cudaMallocManaged(void **p, size_t n, unsigned int) {
	*p = aligned_alloc(256, (n + 255) / 256 * 256);
	return cudaSuccess;
}
cudaError_t cudaMalloc(void **p, size_t n) {
	*p = aligned_alloc(256, (n + 255) / 256 * 256);
	return cudaSuccess;
}
cudaError_t cudaFree(void *p) { free(p); return cudaSuccess; }
cudaError_t cudaMemcpy(void *d, const void *s, size_t n, cudaMemcpyKind) {
	memcpy(d, s, n);
	return cudaSuccess;
}
/* host build: the "symbol" is a null pointer value; host code reads spec::x etc. directly */
cudaError_t cudaMemcpyToSymbol(const void *, const void *, size_t, size_t, cudaMemcpyKind) {
	return cudaSuccess;
}
cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }
cudaError_t cudaGetLastError() { return cudaSuccess; }
const char *cudaGetErrorName(cudaError_t) { return "shim"; }
const char *cudaGetErrorString(cudaError_t) { return "shim"; }
}

#include "common.h"
#include "logger.h"
#include "device/memory.h"
#include "render/spectrum.h"
#include "render/bsdf.h"
#include "render/media.h"
#include "core/light.h"
#include "core/camera.h"
#include "core/sampler.h"
#include "core/mesh.h"

#include "oracle_leaf.h"

namespace krr {
void Logger::log(Level level, const string &msg, bool terminate) {
	if ((int) level >= 3) fprintf(stderr, "[krr-ref log %d] %s\n", (int) level, msg.c_str());
	if (terminate) abort();
}
CUDATrackedMemory CUDATrackedMemory::singleton;
} // namespace krr

using namespace krr;

static bool g_init = false;
static Allocator *g_alloc;

extern "C" int ol_init(void) {
	if (g_init) return 0;
	gpu::set_default_resource(&CUDATrackedMemory::singleton);
	g_alloc = new Allocator(&CUDATrackedMemory::singleton);
	spec::init(*g_alloc);
	RGBToSpectrumTable::init(*g_alloc);
	RGBColorSpace::init(*g_alloc);
	g_init = true;
	return 0;
}

extern "C" const char *ol_backend_name(void) { return "reference"; }

static inline Vector3f V3(const float *p) { return Vector3f(p[0], p[1], p[2]); }
static inline void S3(float *o, const Vector3f &v) { o[0] = v[0]; o[1] = v[1]; o[2] = v[2]; }
static inline void S4(float *o, const Spectrum &s) { for (int i = 0; i < 4; i++) o[i] = s[i]; }
static inline Spectrum SP(const float *p) {
	Spectrum s;
	for (int i = 0; i < 4; i++) s[i] = p[i];
	return s;
}

/* SampledWavelengths keeps lambda/pdfs private; it is a standard-layout pair of float[4] arrays */
static inline SampledWavelengths SW(const float *lambda, const float *pdf) {
	SampledWavelengths swl = SampledWavelengths::sampleUniform(0.f);
	static_assert(sizeof(SampledWavelengths) == 32, "layout");
	float *raw = reinterpret_cast<float *>(&swl);
	for (int i = 0; i < 4; i++) raw[i] = lambda[i], raw[4 + i] = pdf ? pdf[i] : (1.f / (cLambdaMax - cLambdaMin));
	return swl;
}

/* ---------------- sampler ---------------- */
static_assert(sizeof(PCGSampler) == sizeof(OlSampler), "PCG layout");
extern "C" void ol_pcg_set_pixel_sample(OlSampler *s, uint32_t px, uint32_t py, uint32_t idx) {
	reinterpret_cast<PCGSampler *>(s)->setPixelSample(Vector2ui(px, py), idx);
}
extern "C" void ol_pcg_advance(OlSampler *s, int64_t delta) {
	reinterpret_cast<PCGSampler *>(s)->advance(delta);
}
extern "C" float ol_pcg_get1d(OlSampler *s) { return reinterpret_cast<PCGSampler *>(s)->get1D(); }

/* ---------------- spectrum ---------------- */
extern "C" void ol_sample_wavelengths(float u, float lambda[4], float pdf[4]) {
	SampledWavelengths swl = SampledWavelengths::sampleUniform(u);
	for (int i = 0; i < 4; i++) lambda[i] = swl[i], pdf[i] = swl.pdf()[i];
}
extern "C" void ol_from_rgb(const float rgb[3], int type, const float lambda[4], float out[4]) {
	ol_init();
	Spectrum s = Spectrum::fromRGB(RGB(rgb[0], rgb[1], rgb[2]), (SpectrumType) type,
								   SW(lambda, nullptr), *RGBColorSpace::sRGB);
	S4(out, s);
}
extern "C" void ol_to_rgb(const float s[4], const float lambda[4], const float pdf[4], float rgb[3]) {
	ol_init();
	RGB c = SP(s).toRGB(SW(lambda, pdf), *RGBColorSpace::sRGB);
	rgb[0] = c[0], rgb[1] = c[1], rgb[2] = c[2];
}
extern "C" float ol_lum(const float s[4], const float lambda[4], const float pdf[4]) {
	ol_init();
	return RGBColorSpace::sRGB->lum(SP(s), SW(lambda, pdf));
}

/* ---------------- camera ---------------- */
static Affine3f A12(const float *m) {
	/* krr::Transform::matrix() returns a COPY (krrmath/transform.h:44-47): fill a plain matrix and
	 * construct the transform from it */
	Matrix4f mat = Matrix4f::Identity();
	for (int r = 0; r < 3; r++)
		for (int c = 0; c < 4; c++) mat(r, c) = m[r * 4 + c];
	return Affine3f(mat);
}
extern "C" void ol_camera_ray(const OlCamera *cam, int px, int py, int w, int h, const float cs[5],
							  float o[3], float d[3], float *time) {
	rt::CameraData c;
	c.filmSize		= Vector2f(cam->filmSize[0], cam->filmSize[1]);
	c.focalLength	= cam->focalLength;
	c.focalDistance = cam->focalDistance;
	c.lensRadius	= cam->lensRadius;
	c.aspectRatio	= cam->aspectRatio;
	c.shutterOpen	= cam->shutterOpen;
	c.shutterTime	= cam->shutterTime;
	c.transform		= Transformation(A12(cam->transform));
	rt::CameraSample s{Vector2f(cs[0], cs[1]), Vector2f(cs[2], cs[3]), cs[4]};
	Ray r = c.getRay(Vector2i(px, py), Vector2i(w, h), s);
	S3(o, r.origin);
	S3(d, r.dir);
	*time = r.time;
}

/* ---------------- BSDF ---------------- */
static void fillIntr(SurfaceInteraction &intr, const OlShading *sd) {
	ol_init();
	intr.n		   = Vector3f(0, 0, 1);
	intr.tangent   = Vector3f(1, 0, 0);
	intr.bitangent = Vector3f(0, 1, 0);
	intr.wo		   = V3(sd->woWorld);
	intr.sd.IoR	   = sd->IoR;
	intr.sd.diffuse				 = SP(sd->diffuse);
	intr.sd.specular			 = SP(sd->specular);
	intr.sd.specularTransmission = sd->specularTransmission;
	intr.sd.roughness			 = sd->roughness;
	intr.sd.metallic			 = sd->metallic;
	intr.sd.anisotropic			 = sd->anisotropic;
	intr.sd.bsdfType			 = (MaterialType) sd->bsdfType;
	intr.lambda					 = SW(sd->lambda, sd->pdf);
	/* material: only colour space and spectral eta/k are read by BSDF::setup.  One scratch
	 * MaterialData per host thread (no per-call allocation: this adapter must not slow the
	 * reference's code down when it is timed as the CPU baseline). */
	static thread_local rt::MaterialData mat;
	static thread_local ConstantSpectrum etaC(1.5f), kC(0.f);
	mat.mColorSpace = RGBColorSpace::sRGB;
	mat.mMaterialParams.spectralEta = Spectra();
	mat.mMaterialParams.spectralK	= Spectra();
	if (sd->etaKind == 1) { etaC = ConstantSpectrum(sd->etaValue[0]); mat.mMaterialParams.spectralEta = &etaC; }
	if (sd->kKind == 1) { kC = ConstantSpectrum(sd->kValue[0]); mat.mMaterialParams.spectralK = &kC; }
	intr.material = &mat;
}
extern "C" int ol_bsdf_type(const OlShading *sd) {
	BSDFData d;
	d.IoR = sd->IoR; d.diffuse = SP(sd->diffuse); d.specular = SP(sd->specular);
	d.specularTransmission = sd->specularTransmission; d.roughness = sd->roughness;
	d.metallic = sd->metallic; d.anisotropic = sd->anisotropic; d.bsdfType = (MaterialType) sd->bsdfType;
	return (int) d.getBsdfType();
}
extern "C" void ol_bsdf_f_pdf(const OlShading *sd, const float wo[3], const float wi[3], float f[4],
							  float *pdf) {
	SurfaceInteraction intr;
	fillIntr(intr, sd);
	BSDF bsdf(intr);
	S4(f, bsdf.f(V3(wo), V3(wi)));
	*pdf = bsdf.pdf(V3(wo), V3(wi));
}
extern "C" void ol_bsdf_sample(const OlShading *sd, const float wo[3], OlSampler *s, float f[4],
							   float wi[3], float *pdf, int *flags) {
	SurfaceInteraction intr;
	fillIntr(intr, sd);
	BSDF bsdf(intr);
	Sampler sampler = reinterpret_cast<PCGSampler *>(s);
	BSDFSample bs	= bsdf.sample(V3(wo), sampler);
	S4(f, bs.f);
	S3(wi, bs.wi);
	*pdf   = bs.pdf;
	*flags = (int) bs.flags;
}

/* ---------------- area light ---------------- */
struct RefTriLight {
	rt::MeshData mesh;
	rt::InstanceData inst;
	Triangle tri;
	rt::DiffuseAreaLight light;
};
static RefTriLight *makeTri(const OlTriLight *l) {
	ol_init();
	/* one scratch one-triangle mesh per host thread, buffers allocated once (TypedBuffer never
	 * frees -- device/buffer.h:123-127 -- so per-call allocation would leak) */
	static thread_local RefTriLight *r = nullptr;
	if (!r) {
		r = new RefTriLight();
		std::vector<Vector3f> Z(3, Vector3f(0, 0, 0));
		std::vector<Vector3i> I{Vector3i(0, 1, 2)};
		r->mesh.positions.alloc_and_copy_from_host(Z);
		r->mesh.normals.alloc_and_copy_from_host(Z);
		r->mesh.indices.alloc_and_copy_from_host(I);
	}
	for (int c = 0; c < 3; c++) r->mesh.positions[c] = V3(l->p[c]), r->mesh.normals[c] = V3(l->n[c]);
	r->inst.mesh	  = &r->mesh;
	r->inst.transform = Transformation(A12(l->xform));
	r->tri			  = Triangle(0, &r->inst);
	r->light = rt::DiffuseAreaLight(Shape(&r->tri), RGB(l->Le[0], l->Le[1], l->Le[2]), l->twoSided,
									l->scale, RGBColorSpace::sRGB);
	return r;
}
extern "C" void ol_arealight_sample_li(const OlTriLight *l, const float u[2], const float ctxP[3],
									   const float ctxN[3], const float lambda[4], float p[3],
									   float n[3], float L[4], float *pdf) {
	RefTriLight *r = makeTri(l);
	rt::LightSample ls =
		r->light.sampleLi(Vector2f(u[0], u[1]), {V3(ctxP), V3(ctxN)}, SW(lambda, nullptr));
	S3(p, ls.intr.p);
	S3(n, ls.intr.n);
	S4(L, ls.L);
	*pdf = ls.pdf;
}
extern "C" void ol_arealight_L(const OlTriLight *l, const float p[3], const float n[3],
							   const float w[3], const float lambda[4], float L[4]) {
	RefTriLight *r = makeTri(l);
	S4(L, r->light.L(V3(p), V3(n), Vector2f(0, 0), V3(w), SW(lambda, nullptr)));
}
extern "C" float ol_arealight_pdf_li(const OlTriLight *l, const float p[3], const float n[3],
									 const float ctxP[3], const float ctxN[3]) {
	RefTriLight *r = makeTri(l);
	Interaction intr(V3(p), V3(n), Vector2f(0, 0));
	float pdf = r->light.pdfLi(intr, {V3(ctxP), V3(ctxN)});
	return pdf;
}

/* ---------------- analytic lights ---------------- */
static Matrix3f M9(const float *m) {
	Matrix3f r;
	for (int i = 0; i < 3; i++)
		for (int j = 0; j < 3; j++) r(i, j) = m[i * 3 + j];
	return r;
}
extern "C" void ol_light_sample_li(const OlLight *l, const float u[2], const float ctxP[3],
								   const float lambda[4], float p[3], float L[4], float *pdf) {
	ol_init();
	RGB c(l->color[0], l->color[1], l->color[2]);
	rt::LightSampleContext ctx{V3(ctxP), Vector3f(0, 0, 0)};
	SampledWavelengths swl = SW(lambda, nullptr);
	rt::LightSample ls;
	switch (l->type) {
		case 0: ls = rt::PointLight(V3(l->position), c, l->scale).sampleLi(Vector2f(u[0], u[1]), ctx, swl); break;
		case 1: ls = rt::DirectionalLight(M9(l->rotation), c, l->scale, l->sceneRadius).sampleLi(Vector2f(u[0], u[1]), ctx, swl); break;
		case 4: ls = rt::InfiniteLight(M9(l->rotation), c, l->scale, l->sceneRadius).sampleLi(Vector2f(u[0], u[1]), ctx, swl); break;
		default: {
			/* spot: constructor takes cone angles in degrees */
			float inner = std::acos(l->cosInner) * 180.f / M_PI, outer = std::acos(l->cosOuter) * 180.f / M_PI;
			ls = rt::SpotLight(Transformation(A12(l->xform)), c, l->scale, inner, outer)
					 .sampleLi(Vector2f(u[0], u[1]), ctx, swl);
		}
	}
	S3(p, ls.intr.p);
	S4(L, ls.L);
	*pdf = ls.pdf;
}
extern "C" void ol_inflight_Li(const OlLight *l, const float wi[3], const float lambda[4], float L[4]) {
	ol_init();
	RGB c(l->color[0], l->color[1], l->color[2]);
	S4(L, rt::InfiniteLight(M9(l->rotation), c, l->scale, l->sceneRadius).Li(V3(wi), SW(lambda, nullptr)));
}

/* ---------------- media ---------------- */
extern "C" float ol_hg_p(float g, const float wo[3], const float wi[3]) {
	return HGPhaseFunction(g).p(V3(wo), V3(wi));
}
extern "C" void ol_hg_sample(float g, const float wo[3], const float u[2], float wi[3], float *p,
							 float *pdf) {
	PhaseFunctionSample ps = HGPhaseFunction(g).sample(V3(wo), Vector2f(u[0], u[1]));
	S3(wi, ps.wi);
	*p	 = ps.p;
	*pdf = ps.pdf;
}

extern "C" void ol_homogeneous_sample_point(const float sigma_t[3], const float albedo[3], const float Le[3],
											 const float lambda[4], float sigma_a[4], float sigma_s[4], float LeOut[4]) {
	ol_init();
	HomogeneousMedium m(RGB(sigma_t[0], sigma_t[1], sigma_t[2]), RGB(albedo[0], albedo[1], albedo[2]),
						RGB(Le[0], Le[1], Le[2]), 0.f, RGBColorSpace::sRGB);
	MediumProperties mp = m.samplePoint(Vector3f(0, 0, 0), SW(lambda, nullptr));
	S4(sigma_a, mp.sigma_a), S4(sigma_s, mp.sigma_s), S4(LeOut, mp.Le);
}
extern "C" int ol_majorant_segments(const float boundsMin[3], const float boundsMax[3], const int res[3],
									 const float *voxels, const float o[3], const float d[3], float tMin, float tMax,
									 const float sigma_t[4], float *out6, int cap) {
	MajorantGrid grid(AABB3f(V3(boundsMin), V3(boundsMax)), Vector3i(res[0], res[1], res[2]));
	grid.voxels = TypedBufferView<float>(voxels, (size_t) res[0] * res[1] * res[2]);
	Ray ray{V3(o), V3(d)};
	MajorantIterator it(ray, tMin, tMax, SP(sigma_t), voxels ? &grid : nullptr);
	int n = 0;
	while (true) {
		gpu::optional<MajorantSegment> seg = it.next();
		if (!seg) break;
		if (n < cap) {
			out6[6 * n] = seg->tMin, out6[6 * n + 1] = seg->tMax;
			S4(out6 + 6 * n + 2, seg->sigma_maj);
		}
		n++;
	}
	return n;
}

/* getMetallic lives in render/shading.h, which cannot be compiled host-side (OptiX intrinsics);
 * restated from shading.h:17-30 using the reference's own luminance(). */
extern "C" float ol_get_metallic(const float diffuse[3], const float spec[3]) {
	float d = luminance(RGB(diffuse[0], diffuse[1], diffuse[2]));
	float s = luminance(RGB(spec[0], spec[1], spec[2]));
	if (s == 0) return 0;
	float b	   = s + d - 0.08f;
	float c	   = 0.04f - s;
	float root = krr::sqrt(b * b - 0.16f * c);
	float m	   = (root - b) * 12.5f;
	return krr::max(0.f, m);
}

/* ---- spectral tables the reference host application owns; dumped once for the product ---- */
extern "C" int ol_ref_dump_spectral(const char *path) {
	ol_init();
	FILE *f = fopen(path, "wb");
	if (!f) return -1;
	const RGBColorSpace *cs = RGBColorSpace::sRGB;
	uint32_t magic = 0x4b525253u /* 'KRRS' */, version = 1, nLambda = 471, res = 64;
	fwrite(&magic, 4, 1, f); fwrite(&version, 4, 1, f); fwrite(&nLambda, 4, 1, f); fwrite(&res, 4, 1, f);
	for (const DenselySampledSpectrum *d : {cs->CIE_X, cs->CIE_Y, cs->CIE_Z, &cs->illuminant})
		for (int l = 360; l <= 830; l++) { float v = (*d)((float) l); fwrite(&v, 4, 1, f); }
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { float v = cs->XYZFromRGB(i, j); fwrite(&v, 4, 1, f); }
	for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { float v = cs->RGBFromXYZ(i, j); fwrite(&v, 4, 1, f); }
	/* RGBToSpectrumTable keeps zNodes/coeffs private: standard layout {const float*, const CoefficientArray*} */
	struct Raw { const float *z; const RGBToSpectrumTable::CoefficientArray *c; };
	const Raw *raw = reinterpret_cast<const Raw *>(RGBToSpectrumTable::sRGB);
	fwrite(raw->z, 4, 64, f);
	fwrite(raw->c, 4, 3 * 64 * 64 * 64 * 3, f);
	fclose(f);
	return 0;
}
