/* oracle_leaf.h -- TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Leaf-function C API of the CPU oracle.  Two backends implement it:
 *   oracle/ref/backend_ref.cpp   -> calls the reference's own KRR_CALLABLE classes, compiled
 *                                   host-side from /root/reference/src (built into oracle/_ref/)
 *   oracle/port/backend_port.cpp -> plain C++ restatement of the same functions, each citing the
 *                                   reference file:line it follows
 * The wavefront driver (oracle/driver.cpp) is written once against this API and linked to either
 * backend, so the whole path can be run through the reference's code or through the port, and the
 * two can be compared function by function.
 *
 * All structs are POD; vectors are float[3]; spectra are float[4] (KRR_N_SPECTRUM_SAMPLES = 4,
 * reference src/core/config.in.h:15).
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- PCG sampler: reference src/core/sampler.h:13-91 ---- */
typedef struct OlSampler { uint64_t state, inc; } OlSampler;
void  ol_pcg_set_pixel_sample(OlSampler *s, uint32_t px, uint32_t py, uint32_t sampleIndex);
void  ol_pcg_advance(OlSampler *s, int64_t delta);
float ol_pcg_get1d(OlSampler *s);

/* ---- wavelengths + colour: reference src/render/spectrum.h:51-64, 488-529; color.h:84-169 ---- */
void  ol_sample_wavelengths(float u, float lambda[4], float pdf[4]);
/* type: 0 RGBBounded, 1 RGBUnbounded, 2 RGBIlluminant (src/render/color.h:32-36) */
void  ol_from_rgb(const float rgb[3], int type, const float lambda[4], float out[4]);
void  ol_to_rgb(const float s[4], const float lambda[4], const float pdf[4], float rgb[3]);
float ol_lum(const float s[4], const float lambda[4], const float pdf[4]);

/* ---- camera: reference src/core/camera.h:20-58 ---- */
typedef struct OlCamera {
	float filmSize[2];
	float focalLength, focalDistance, lensRadius, aspectRatio, shutterOpen, shutterTime;
	float transform[12]; /* 3x4 row-major camera->world */
} OlCamera;
/* cs = {pFilm.x, pFilm.y, pLens.x, pLens.y, time} */
void ol_camera_ray(const OlCamera *cam, int px, int py, int w, int h, const float cs[5],
				   float o[3], float d[3], float *time);

/* ---- BSDF: reference src/render/bsdf.h:19-54, shared.h:34-73, materials/ ---- */
typedef struct OlShading {
	float IoR;
	float diffuse[4];
	float specular[4];
	float specularTransmission, roughness, metallic, anisotropic;
	int   bsdfType;			   /* MaterialType: 0 null 1 diffuse 2 dielectric 3 conductor 4 disney */
	float woWorld[3];		   /* DisneyBsdf::setup reads AbsCosTheta(intr.wo) of the WORLD wo (disney.h:263) */
	float lambda[4], pdf[4];
	/* conductor: spectral eta / k (0 none, 1 constant(a), 2 named spectrum handled by backend) */
	int   etaKind; float etaValue[4];
	int   kKind;   float kValue[4];
} OlShading;
int  ol_bsdf_type(const OlShading *sd); /* BSDFData::getBsdfType, shared.h:46-73 */
void ol_bsdf_f_pdf(const OlShading *sd, const float wo[3], const float wi[3], float f[4], float *pdf);
void ol_bsdf_sample(const OlShading *sd, const float wo[3], OlSampler *s, float f[4], float wi[3],
					float *pdf, int *flags);

/* ---- diffuse area light on one triangle: reference src/core/light.h:152-206, shape.h:23-132,
 *      created per emissive triangle as mesh.cpp:39-59 (Le/=max, scale=max, one-sided) ---- */
typedef struct OlTriLight {
	float p[3][3];	 /* object-space vertex positions */
	float n[3][3];	 /* object-space vertex normals */
	float xform[12]; /* instance object->world 3x4 row-major */
	float Le[3];
	float scale;
	int	  twoSided;
} OlTriLight;
void  ol_arealight_sample_li(const OlTriLight *l, const float u[2], const float ctxP[3],
							 const float ctxN[3], const float lambda[4], float p[3], float n[3],
							 float L[4], float *pdf);
void  ol_arealight_L(const OlTriLight *l, const float p[3], const float n[3], const float w[3],
					 const float lambda[4], float L[4]);
float ol_arealight_pdf_li(const OlTriLight *l, const float p[3], const float n[3],
						  const float ctxP[3], const float ctxN[3]);

/* ---- analytic lights: reference src/core/light.h:30-150, 208-259 ---- */
typedef struct OlLight {
	int	  type; /* 0 point, 1 directional, 2 spot, 4 infinite (index in rt::Light's type list, light.h:261-263) */
	float color[3];
	float scale;
	float position[3];
	float rotation[9]; /* row-major 3x3 */
	float sceneRadius;
	float cosInner, cosOuter;
	float xform[12], xformInv[12]; /* spot */
} OlLight;
void ol_light_sample_li(const OlLight *l, const float u[2], const float ctxP[3], const float lambda[4],
						float p[3], float L[4], float *pdf);
void ol_inflight_Li(const OlLight *l, const float wi[3], const float lambda[4], float L[4]);

/* ---- media: reference src/render/media.h/.cpp, phase.h ---- */
float ol_hg_p(float g, const float wo[3], const float wi[3]);
void  ol_hg_sample(float g, const float wo[3], const float u[2], float wi[3], float *p, float *pdf);

/* HomogeneousMedium::samplePoint, reference src/render/media.h:121-126 (Le: RGBIlluminant, :139-141) */
void  ol_homogeneous_sample_point(const float sigma_t[3], const float albedo[3], const float Le[3],
								  const float lambda[4], float sigma_a[4], float sigma_s[4], float LeOut[4]);
/* MajorantIterator over a MajorantGrid (reference src/render/media.h:41-106); voxels == NULL runs the
 * homogeneous variant.  o/d are in MEDIUM space.  Writes up to cap segments as {tMin, tMax, sigma_maj[4]}
 * (6 floats each) and returns the number of segments the iterator produced. */
int	  ol_majorant_segments(const float boundsMin[3], const float boundsMax[3], const int res[3],
						   const float *voxels, const float o[3], const float d[3], float tMin, float tMax,
						   const float sigma_t[4], float *out6, int cap);

/* misc */
float		ol_get_metallic(const float diffuse[3], const float spec[3]); /* shading.h:17-30 (restated in both) */
const char *ol_backend_name(void);
int			ol_init(void);

#ifdef __cplusplus
}
#endif
