/* krr_wfpt.h -- C ABI of the B200-native WavefrontPathTracer pass.
 *
 * The reference (cuteday/KiRaRay) has NO C ABI: its pass is the C++ class `WavefrontPathTracer :
 * RenderPass` (reference src/render/wavefront/integrator.h:24-104) driven through the RenderPass
 * virtuals (src/core/renderpass.h:138-200).  BASELINE.json's north_star puts a thin C ABI *under*
 * that class; each entry point below is derived 1:1 from the virtual (or member) it backs, cited
 * at the declaration.  The C++17 host class that keeps the reference's plugin surface lives in
 * kiraray_b200/host/ and calls ONLY these functions (see INTEGRATION.md for the binding a KiRaRay
 * maintainer would add).
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns 0 on success
 * and a negative KRR_E_* code on failure (never throws, never exits -- the reference's
 * Log(Fatal)/CUDA_CHECK -> exit(1), src/core/logger.cpp:100, is replaced by error returns);
 * krr_wfpt_last_error() gives the message.  A handle belongs to the CUDA device that was current
 * at krr_wfpt_create(); it is not thread-safe; all work is ordered on the caller's stream: a call
 * returns with everything it enqueued ordered before whatever the caller enqueues next on that stream
 * (krr_wfpt_render may run part of a frame on an internal stream that is forked from and joined back
 * to the caller's stream with events inside the call -- pass parameter "bands": 1 to switch that off).
 * Host arrays passed to set_* are copied before the call returns.
 */
#ifndef KRR_WFPT_H
#define KRR_WFPT_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

#define KRR_WFPT_ABI_VERSION 6

enum {
	KRR_OK			  = 0,
	KRR_E_INVALID	  = -1, /* bad argument / bad JSON */
	KRR_E_CUDA		  = -2, /* CUDA runtime error (message has the cudaError string) */
	KRR_E_STATE		  = -3, /* call order violated (e.g. render before set_scene/resize) */
	KRR_E_UNSUPPORTED = -4,
};

/* MaterialType, reference src/core/raytracing.h:26-33 */
enum { KRR_MAT_NULL = 0, KRR_MAT_DIFFUSE = 1, KRR_MAT_DIELECTRIC = 2, KRR_MAT_CONDUCTOR = 3, KRR_MAT_DISNEY = 4 };
/* Material::ShadingModel, reference src/core/texture.h:124-127 */
enum { KRR_SHADING_METALLIC_ROUGHNESS = 0, KRR_SHADING_SPECULAR_GLOSSINESS = 1 };
/* Material::TextureType, reference src/core/texture.h:115-122 */
enum { KRR_TEX_DIFFUSE = 0, KRR_TEX_SPECULAR = 1, KRR_TEX_EMISSIVE = 2, KRR_TEX_NORMAL = 3, KRR_TEX_TRANSMISSION = 4, KRR_TEX_COUNT = 5 };
/* index of the light class in rt::Light's tagged-pointer type list, reference src/core/light.h:261-263 */
enum { KRR_LIGHT_POINT = 0, KRR_LIGHT_DIRECTIONAL = 1, KRR_LIGHT_SPOT = 2, KRR_LIGHT_DIFFUSE_AREA = 3, KRR_LIGHT_INFINITE = 4 };
/* spectrum kinds for spectral eta / k, reference src/render/spectrum.h:92-96 */
enum { KRR_SPEC_NONE = 0, KRR_SPEC_CONSTANT = 1, KRR_SPEC_CAUCHY = 2, KRR_SPEC_SELLMEIER = 3, KRR_SPEC_TABULATED = 4 };
enum { KRR_MEDIUM_HOMOGENEOUS = 0, KRR_MEDIUM_GRID = 1 };

/* rt::TextureData, reference src/core/texture.h:184-204: constant value and/or an RGBA32F image
 * (bilinear, wrap addressing, as texture.cpp:229-246). */
typedef struct KrrTextureDesc {
	int32_t		 valid;		/* TextureData::mValid */
	float		 value[4];	/* TextureData::mValue */
	const float *image;		/* optional RGBA32F texels, row-major, NULL = constant texture */
	int32_t		 width, height;
} KrrTextureDesc;

typedef struct KrrSpectrumDesc {
	int32_t		 kind;		/* KRR_SPEC_* */
	float		 a[3], b[3];/* constant: a[0]; cauchy: a[0], b[0]; sellmeier: b = a[], c = b[] */
	const float *lambdas;	/* tabulated (piecewise linear): n samples */
	const float *values;
	int32_t		 n;
} KrrSpectrumDesc;

/* rt::MaterialData, reference src/core/texture.h:206-224 (+ Material::MaterialParams :129-137) */
typedef struct KrrMaterialDesc {
	float			diffuse[4];
	float			specular[4];
	float			specular_transmission;
	float			anisotropic;
	float			ior;
	KrrSpectrumDesc spectral_eta, spectral_k;
	KrrTextureDesc	textures[KRR_TEX_COUNT];
	int32_t			bsdf_type;	   /* KRR_MAT_* */
	int32_t			shading_model; /* KRR_SHADING_* */
	int32_t			color_space;   /* ColorSpaceType; only 0 (sRGB) is accepted */
} KrrMaterialDesc;

/* rt::MeshData, reference src/core/mesh.h:24-35 */
typedef struct KrrMeshDesc {
	const float	  *positions; /* 3 * n_vertices */
	const float	  *normals;	  /* 3 * n_vertices or NULL */
	const float	  *texcoords; /* 2 * n_vertices or NULL */
	const float	  *tangents;  /* 3 * n_vertices or NULL */
	const int32_t *indices;	  /* 3 * n_triangles */
	int32_t		   n_vertices, n_triangles;
	int32_t		   material;	   /* index into materials, -1 = null material (medium interface) */
	int32_t		   medium_inside;  /* index into media or -1 */
	int32_t		   medium_outside; /* index into media or -1 */
	float		   Le[3];		   /* Mesh::Le, mesh-specific emission (pbrt import), src/core/mesh.h:92 */
} KrrMeshDesc;

/* one keyed SRT sample for motion blur / animation: scale, quaternion (x,y,z,w), translation
 * (reference resamples node SRTs to regular steps, src/core/device/optix.cpp:400-471) */
typedef struct KrrSRT { float s[3]; float q[4]; float t[3]; } KrrSRT;

/* One node of the scene graph ABOVE a mesh instance, for multi-level scenes with motion blur
 * (OptixSceneMultiLevel::buildIASForNode, reference src/core/device/optix.cpp:472-563): a node is
 * either a static local transform (the OptixInstance transform of its IAS entry) or, when it is
 * animated and motion blur is on, an SRT motion transform with keys regularly spaced over
 * [time_begin, time_end] that REPLACES the local transform (optix.cpp:400-470, 484-486, 540-561).
 * The object->world transform of an instance at ray time t is the product of its chain,
 * root first: M(t) = M_root(t) * ... * M_node(t).  Key interpolation follows OptiX SRT motion
 * transforms: scale and translation linearly, the quaternion linearly and then normalised; time is
 * clamped to [time_begin, time_end]. */
typedef struct KrrTransformNodeDesc {
	int32_t		  parent;		 /* index into transform_nodes, -1 = top of the chain */
	float		  transform[12]; /* static local transform (used when n_motion_keys < 2) */
	int32_t		  n_motion_keys; /* >= 2: SRT motion transform */
	const KrrSRT *motion_keys;
	float		  time_begin, time_end; /* MotionKeyframes::startTime / endTime */
} KrrTransformNodeDesc;

/* rt::InstanceData, reference src/core/mesh.h:37-58: mesh pointer + object->world transform
 * (3x4 row-major, the layout of krr::Affine3f, src/core/math/include/krrmath/transform.h:12).
 * `transform` is always the node's GLOBAL transform at the current animation time
 * (InstanceData::transform, mesh.cpp:61): static instances are traced with it, and lights of
 * emissive instances use it even when the instance moves, as in the reference (light.h:152-206). */
typedef struct KrrInstanceDesc {
	int32_t		  mesh;
	float		  transform[12];
	int32_t		  n_motion_keys; /* single-level shorthand: >= 2 keys uniformly spaced over
									[options.starttime, options.endtime] = one motion node, no parent */
	const KrrSRT *motion_keys;
	int32_t		  transform_node; /* index into KrrSceneDesc::transform_nodes of the node that holds
									 this mesh instance, or -1 (then `transform` / motion_keys apply) */
} KrrInstanceDesc;

/* analytic scene lights (point / directional / spot / infinite), reference src/core/light.h:30-259.
 * Diffuse area lights are NOT listed here: like the reference (src/core/mesh.cpp:39-59) the pass
 * creates one per triangle of every instance whose material has a constant emissive texture. */
typedef struct KrrLightDesc {
	int32_t		   type;  /* KRR_LIGHT_* (not DIFFUSE_AREA) */
	float		   color[3];
	float		   scale;
	float		   transform[12]; /* node global transform: position = translation, direction = R * +Z */
	float		   inner_cone_deg, outer_cone_deg;
	float		   scene_radius;  /* root bounding-box diagonal length (src/core/light.cpp:20-21, 32-33) */
	KrrTextureDesc texture;		  /* infinite light lat-long image (optional) */
} KrrLightDesc;

/* media: HomogeneousMedium / dense-grid stand-in for NanoVDBMedium<float>, reference src/render/media.h:108-227 */
typedef struct KrrMediumDesc {
	int32_t		 type; /* KRR_MEDIUM_* */
	float		 sigma_t[3], albedo[3], Le[3];
	float		 g;
	/* grid medium */
	float		 transform[12];	 /* medium -> world */
	float		 bounds_min[3], bounds_max[3]; /* medium-space bounds of the density grid */
	int32_t		 res[3];
	const float *density;		 /* res[0]*res[1]*res[2], x fastest */
	float		 scale;			 /* density scale */
	/* optional (NULL = none): RGB single-scattering albedo per voxel, 3 floats per voxel, same resolution, bounds and
	 * trilinear lookup as `density` (NanoVDBMedium::albedoGrid, media.h:168-170: when present it replaces `albedo`) */
	const float *albedo_grid;
} KrrMediumDesc;

/* OptixSceneParameters, reference src/core/device/scene.h:34-47 */
typedef struct KrrSceneOptions {
	int32_t animated, multilevel, motionblur;
	float	starttime, endtime;
} KrrSceneOptions;

typedef struct KrrSceneDesc {
	const KrrMeshDesc	  *meshes;	  int32_t n_meshes;
	const KrrInstanceDesc *instances; int32_t n_instances;
	const KrrMaterialDesc *materials; int32_t n_materials;
	const KrrLightDesc	  *lights;	  int32_t n_lights;
	const KrrMediumDesc	  *media;	  int32_t n_media;
	KrrSceneOptions		   options;
	const KrrTransformNodeDesc *transform_nodes; int32_t n_transform_nodes; /* only read when options.motionblur */
} KrrSceneDesc;

/* rt::CameraData, reference src/core/camera.h:20-30 */
typedef struct KrrCameraData {
	float	film_size[2];
	float	focal_length, focal_distance, lens_radius, aspect_ratio, shutter_open, shutter_time;
	float	transform[12]; /* camera -> world, 3x4 row-major */
	int32_t medium;		   /* index of the medium the camera is in, or -1 */
} KrrCameraData;

/* RGBColorSpace + RGBToSpectrumTable + CIE curves: global data the reference host application owns
 * and hands to every pass (LaunchParameters::colorSpace, src/render/wavefront/wavefront.h:28;
 * MaterialData::mColorSpace).  Layouts as reference src/render/color.h:113-114, spectrum.h:375-378. */
typedef struct KrrColorSpaceData {
	const float *cie_x, *cie_y, *cie_z; /* 471 samples, 360..830 nm, 1 nm step */
	const float *illuminant;			/* 471 samples (normalised std illuminant, D65 for sRGB) */
	float		 xyz_from_rgb[9], rgb_from_xyz[9]; /* row-major */
	const float *z_nodes;				/* 64 */
	const float *coeffs;				/* [3][64][64][64][3] */
} KrrColorSpaceData;

/* per-stage counters of the last render() (reference PROFILE() stage names, integrator.cpp:51-167) */
#define KRR_MAX_DEPTH_STATS 64
typedef struct KrrStats {
	uint64_t camera_rays;
	uint64_t closest_rays;	   /* items popped from the ray queues (primary + bounce) */
	uint64_t shadow_rays;	   /* items popped from the shadow queue */
	uint64_t scatter_items, hit_light_items, miss_items, medium_sample_items, medium_scatter_items;
	uint64_t closest_by_depth[KRR_MAX_DEPTH_STATS];
	uint64_t shadow_by_depth[KRR_MAX_DEPTH_STATS];
	uint64_t kernel_launches;  /* kernels this handle launched in the last render()+begin_frame() */
	uint64_t bvh_nodes, bvh_triangles, tlas_nodes;
} KrrStats;

typedef struct KrrWfpt KrrWfpt;

/* WavefrontPathTracer() + from_json, integrator.h:29, 96-103.  params_json is the pass's "params"
 * object: {"nee": true, "enable_medium": true, "max_depth": 10, "rr": 0.8, "enable_clamp": false,
 * "clamp_max": 1000.0}; plus "spp" (samplesPerPixel, UI-only in the reference, integrator.h:80,
 * integrator.cpp:271) -- NULL or "{}" gives the reference defaults. */
int krr_wfpt_create(const char *params_json, KrrWfpt **out);
/* Scheduling parameters the same JSON object accepts.  None is a reference parameter and none changes a result: each is
 * pinned by a bit-identical-film test (tests/test_gpu_*.py); DESIGN.md section 4 has the measurements behind the defaults.
 *   "frame_batch": F (1..64, default 1)  one render() carries frame indices frame_index .. frame_index + F - 1 through the
 *                  same launches and returns the mean of their films (takes effect at the next resize / set_scene)
 *   "bands": 0..4                        see krr_wfpt_render below
 *   "fuse_stages": true                  2 launches per depth (hit / miss in the scatter launch, shadow + next closest in
 *                                        one trace launch) instead of the reference's 4; off with participating media
 *   "rr_in_trace": true                  Russian roulette evaluated when the closest stage routes a hit
 *   "implicit_depth0": true              depth-0 ray items store origin and direction only
 *   "merge_static" / "flatten_static": true   static instances share one world-space BLAS (set_scene)
 *   "flat_blas_max": 48                  a BLAS of at most this many triangles is a flat list (set_scene)
 *   "refill": 0 (= automatic)            idle lanes of a trace warp that trigger finalisation + refill
 *   "tail_depth": -1 (= automatic, off)  from this loop depth on ONE launch finishes every path
 *   "sort_rays": -1 (= off), "sort_key"  trace the rays of depth >= 1 in (octant, origin Morton code) order (experiment)
 *   "l2_persist_mb": 0                   keep the first megabytes of the BVH node pool as persisting L2 lines (experiment)
 *   "pdl": false                         programmatic dependent launch between the stage kernels
 *   "debug_taps": false                  record camera samples and depth-0 hits for the parity taps */
void krr_wfpt_destroy(KrrWfpt *h);
int krr_wfpt_set_params(KrrWfpt *h, const char *params_json);

/* KRR_DEFAULT_COLORSPACE / spec::init, src/render/spectrum.cpp:107-157, 326-360 */
int krr_wfpt_set_color_space(KrrWfpt *h, const KrrColorSpaceData *cs);

/* WavefrontPathTracer::setScene, integrator.cpp:186-203 (scene upload device/scene.cpp:28-173 and
 * acceleration-structure build device/optix.cpp:143-250, 357-398 happen inside). */
int krr_wfpt_set_scene(KrrWfpt *h, const KrrSceneDesc *scene);

/* WavefrontPathTracer::resize, integrator.cpp:181-184 */
int krr_wfpt_resize(KrrWfpt *h, int32_t width, int32_t height);

/* Scene::update -> RTScene::updateAccelStructure (TLAS refit), device/optix.cpp:346-354, 618-669.
 * transforms: n x 12 floats (3x4 row-major).  Normally a refit enqueued on cuda_stream (no host synchronisation).
 * The FIRST time an instance moves that set_scene had merged into the static world-space BLAS ("merge_static" /
 * "flatten_static"), it is taken out of it: one device synchronisation and a rebuild of the acceleration structures
 * (milliseconds to ~1 s for tens of millions of triangles), once per such instance set.  Scenes that animate instances
 * from the start should mark them by passing motion keys / transform nodes, or switch the two parameters off. */
int krr_wfpt_update_instances(KrrWfpt *h, const int32_t *instance_ids, const float *transforms, int32_t n, void *cuda_stream);

/* WavefrontPathTracer::beginFrame, integrator.cpp:205-221.  frame_index is DeviceManager's frame
 * counter (first frame is 1, src/core/window.cpp:457). */
int krr_wfpt_begin_frame(KrrWfpt *h, uint64_t frame_index, const KrrCameraData *camera, void *cuda_stream);

/* WavefrontPathTracer::render, integrator.cpp:223-267.  film: device pointer to width*height
 * float4 (RGBA32F, alpha 1), row H-1-y like CudaRenderTarget::write (device/cuda.h:33-45). */
int krr_wfpt_render(KrrWfpt *h, float *film_rgba_device, void *cuda_stream);
/* Scheduling parameter of the pass (not a reference parameter; results do not depend on it): "bands" = number
 * of interleaved row sets a frame is rendered as, each with its own queues on its own stream; 0 (default)
 * = automatic: 2 for a scene that is one flat triangle list, else 1; at most 4. */

/* Same, with a HOST film buffer: render + device->host copy + stream synchronise. */
int krr_wfpt_render_to_host(KrrWfpt *h, float *film_rgba_host, void *cuda_stream);

/* Pipelined read-back for callers that consume frames on the host (the reference reads its film back only
 * to save it, AccumulatePass / RenderContext::readback): render as above, then copy the film to
 * film_rgba_host (pinned memory for a truly asynchronous copy) on an internal copy stream, WITHOUT
 * synchronising; the caller's stream is free to render the next frame meanwhile (two internal device films
 * alternate).  The host buffer holds the frame once krr_wfpt_wait_host() has returned. */
int krr_wfpt_render_to_host_async(KrrWfpt *h, float *film_rgba_host, void *cuda_stream);
int krr_wfpt_wait_host(KrrWfpt *h);

/* ---- multi-GPU: scene replicated, work split by image tile x spp slice, ONE exchange step ----
 * The reference is single-device (src/core/device/context.cpp:37-40); these entry points are the product's own.
 * One handle per GPU; the film-reduction communicator is NCCL over NVLink (resolved with dlopen at the first call:
 * KRR_E_UNSUPPORTED when libnccl.so.2 is absent).  Either every process calls krr_wfpt_comm_init_rank with the id
 * rank 0 obtained from krr_wfpt_comm_unique_id (one process per GPU), or ONE process passes all its handles to
 * krr_wfpt_comm_init_all (one host thread per handle afterwards: kiraray_b200/host MultiDeviceRenderApp). */
int krr_wfpt_comm_unique_id(uint8_t *out128);
int krr_wfpt_comm_init_rank(KrrWfpt *h, const uint8_t *id128, int32_t world, int32_t rank);
int krr_wfpt_comm_init_all(KrrWfpt **handles, int32_t n);
int krr_wfpt_comm_destroy(KrrWfpt *h);
/* film (device RGBA32F, W x H) of every rank is summed onto `root` in place (ncclReduce, ordered on the stream),
 * then multiplied by `scale` on the root (1 / spp slices).  World size 1: only the scale. */
int krr_wfpt_reduce_film(KrrWfpt *h, float *film_rgba_device, int32_t root, float scale, void *cuda_stream);
/* krr_wfpt_render + krr_wfpt_reduce_film + (root only) the pipelined read-back of krr_wfpt_render_to_host_async;
 * film_rgba_host may be NULL on the other ranks.  krr_wfpt_wait_host() as above. */
int krr_wfpt_render_reduce_to_host_async(KrrWfpt *h, float *film_rgba_host, int32_t root, float scale, void *cuda_stream);

/* Multi-GPU work split by image tile (no reference counterpart: single device,
 * device/context.cpp:37-40).  This handle renders pixel rows [row_begin,row_end) only; the other
 * rows of the film are written as zeros, so films of disjoint tiles ADD to the full film.  The
 * spp axis is split by FRAME: give each GPU its own frame_index (the reference accumulates spp
 * across frames in AccumulatePass, accumulate.cu:30-79) and sum / average the films (NCCL). */
int krr_wfpt_set_partition(KrrWfpt *h, int32_t row_begin, int32_t row_end);

int krr_wfpt_get_stats(KrrWfpt *h, KrrStats *out);

/* Per-stage device time, the counterpart of the reference's PROFILE("...") scopes
 * (integrator.cpp:51,68,79,93,111,167; src/render/profiler/profiler.h:212-235): when enabled, every
 * stage launch is bracketed by CUDA events on the launching stream.  get_stage_times synchronises
 * on the last event and returns, per stage, the summed milliseconds and the number of launches
 * since profiling was enabled (or since the last reset). */
enum { KRR_STAGE_CAMERA = 0, KRR_STAGE_CLOSEST = 1, KRR_STAGE_HIT_MISS = 2, KRR_STAGE_SCATTER = 3, KRR_STAGE_SHADOW = 4,
	   KRR_STAGE_RESOLVE = 5, KRR_STAGE_MEDIUM = 6,
	   KRR_STAGE_TRACE = 7, /* fused launch: shadow rays of depth d + closest rays of depth d + 1 */
	   KRR_STAGE_TAIL = 8,	/* tail launch: every remaining bounce of the paths still alive at "tail_depth" */
	   KRR_STAGE_COUNT = 9 };
int krr_wfpt_set_profiling(KrrWfpt *h, int32_t enable);
int krr_wfpt_get_stage_times(KrrWfpt *h, double *ms, int32_t *launches, int32_t reset);
/* the same events launch by launch, in issue order: stage id and milliseconds of each; returns the
 * number of launches recorded (<= capacity are written) */
int krr_wfpt_get_launch_times(KrrWfpt *h, int32_t *stage, float *ms, int32_t capacity);

/* ---- parity / debug taps (read-only views of device state; used by tests and smoke) ---- */
/* depth-0 hit per pixel of the LAST sample rendered: instance id and primitive id (-1 = miss) */
int krr_wfpt_debug_first_hits(KrrWfpt *h, int32_t *instance_ids_host, int32_t *prim_ids_host);
/* PixelState after begin_frame (+ camera sample after render): sampler state (2 x u64), lambda[4],
 * camera sample[5] per pixel; any pointer may be NULL */
int krr_wfpt_debug_pixel_state(KrrWfpt *h, uint64_t *sampler_host, float *lambda_host, float *camera_sample_host);
/* Capture integer fields of the queues at (sample_id, depth) during the next render(): call before
 * render(); afterwards fetch with krr_wfpt_debug_queue.  queue: 0 ray(current), 1 miss, 2 hitLight,
 * 3 scatter, 4 shadow, 5 next ray.  Fields per item: pixelId, depth, bsdfType, aux (light index /
 * material type / -1).  Returns the item count (>= 0) or an error. */
int krr_wfpt_debug_capture(KrrWfpt *h, int32_t sample_id, int32_t depth);
int krr_wfpt_debug_queue(KrrWfpt *h, int32_t queue, int32_t *items4_host, int32_t capacity);

/* ---- leaf-function taps: run the DEVICE implementations of the reference's KRR_CALLABLE leaf
 * functions (BSDF variant src/render/bsdf.h:19-54 + materials/, lights src/core/light.h:30-259,
 * colour src/render/spectrum.h:488-529, camera src/core/camera.h:32-58) on caller-supplied inputs,
 * one CUDA thread per query.  They exist so that parity tests can compare every BSDF / light /
 * colour routine with the reference's own code value by value; the render path never calls them. */
typedef struct KrrLeafBsdfQuery {
	float	 ior, diffuse[4], specular[4], specular_transmission, roughness, metallic, anisotropic;
	int32_t	 bsdf_type;			 /* KRR_MAT_* */
	float	 wo[3], wi[3];		 /* local shading frame (n = +z); wo doubles as the world-space wo */
	float	 wavelength_u;		 /* SampledWavelengths::sampleUniform(u) */
	uint32_t seed_px, seed_py, seed_index; /* PCGSampler::setPixelSample((px,py), index) for sample() */
	int32_t	 eta_kind;			 /* 0 none, 1 constant spectral eta (conductor) */
	float	 eta;
} KrrLeafBsdfQuery;
typedef struct KrrLeafBsdfResult {
	int32_t type_flags;			 /* BSDFData::getBsdfType */
	float	f[4], pdf;			 /* BSDF::f(wo, wi), BSDF::pdf(wo, wi) */
	float	s_f[4], s_wi[3], s_pdf; /* BSDF::sample(wo, sampler) */
	int32_t s_flags;
} KrrLeafBsdfResult;
int krr_wfpt_debug_eval_bsdf(KrrWfpt *h, const KrrLeafBsdfQuery *queries_host, int32_t n, KrrLeafBsdfResult *results_host);

typedef struct KrrLeafLightQuery {
	int32_t type;				 /* KRR_LIGHT_* ; DIFFUSE_AREA uses the triangle fields */
	float	p[3][3], n[3][3];	 /* triangle, object space */
	float	transform[12];		 /* object->world (triangle) / node transform (analytic) */
	float	color[3], scale;	 /* area light: Le (already divided by its max) and scale */
	int32_t two_sided;
	float	scene_radius, cos_inner, cos_outer;
	float	u[2], ctx_p[3], ctx_n[3], wi[3];
	float	wavelength_u;
} KrrLeafLightQuery;
typedef struct KrrLeafLightResult {
	float p[3], n[3], L[4], pdf; /* sampleLi */
	float L_eval[4];			 /* area: L(p, n, uv, normalize(ctx_p - p)); infinite: Li(wi) */
	float pdf_li;				 /* area: pdfLi(sampled point, ctx) */
} KrrLeafLightResult;
int krr_wfpt_debug_eval_light(KrrWfpt *h, const KrrLeafLightQuery *queries_host, int32_t n, KrrLeafLightResult *results_host);

/* in: rgb[3], wavelength_u, spectrum[4]  (8 floats per query);
 * out: fromRGB as RGBBounded[4], RGBUnbounded[4], RGBIlluminant[4], toRGB(spectrum)[3], lum(spectrum), lambda[4]  (20 floats) */
int krr_wfpt_debug_eval_color(KrrWfpt *h, const float *in8_host, int32_t n, float *out20_host);
/* CameraData::getRay: in px, py (as floats), camera sample[5] (7 floats); out origin[3], dir[3], time (7 floats) */
int krr_wfpt_debug_camera_rays(KrrWfpt *h, const KrrCameraData *camera, int32_t width, int32_t height, const float *in7_host, int32_t n, float *out7_host);

/* object->world (12 floats) and world->object (12 floats) of instance instance_ids[i] for a ray that
 * carries times[i]: what optixGetObjectToWorldTransformMatrix / optixGetWorldToObjectTransformMatrix
 * report at a hit (getInstanceTransform, src/render/shading.h:70-76), evaluated on the device */
int krr_wfpt_debug_instance_xf(KrrWfpt *h, const int32_t *instance_ids_host, const float *times_host, int32_t n, float *out24_host);

/* ---- the reference's MegakernelPathTracer (src/render/megakernel/device.cu:50-195) on the
 * same handle: one launch, one lane per pixel, power-heuristic MIS; uses the handle's scene, size and the
 * params nee / max_depth / rr / spp.  film: device float4[w*h] (sum over spp, like the reference).  An
 * independent estimator for cross-validation, not a fast path. */
int krr_wfpt_render_megakernel(KrrWfpt *h, uint64_t frame_index, const KrrCameraData *camera, float *film_device, void *cuda_stream);

/* ---- next row: AccumulatePass kernel, src/render/passes/accumulate/accumulate.cu:30-52 ----
 * accum, film: device float4[n_pixels]; film is replaced by the running average. */
int krr_accumulate_f32(float *accum, float *film, int64_t n_pixels, uint64_t accum_count,
					   uint64_t max_accum_count, int32_t moving_average, void *cuda_stream);

/* the same in double precision ("precision": "double", accumulate.h:17, accumulate.cu:68-75): accum is
 * device double4[n_pixels] */
int krr_accumulate_f64(double *accum, float *film, int64_t n_pixels, uint64_t accum_count,
					   uint64_t max_accum_count, int32_t moving_average, void *cuda_stream);
/* AccumulatePass::saveImage (accumulate.cu:90-111): accumulated sum * (1 / accum_count) as RGBA32F into HOST
 * memory; accum is the float4 (is_double = 0) or double4 (is_double = 1) device buffer.  Synchronises. */
int krr_accumulate_read_average(const void *accum, int32_t is_double, uint64_t accum_count, int64_t n_pixels,
								float *out_rgba_host, void *cuda_stream);

/* ---- ErrorMeasurePass metric, src/render/passes/errormeasure/metrics.cu:64-128 ----
 * film, reference: device float4[n_pixels]; *result_host = mean over pixels of the per-pixel error (mean over
 * RGB, reference pixels with inf/NaN count 0, per-pixel error clamped at 100).  Synchronises the stream. */
enum { KRR_METRIC_MSE = 0, KRR_METRIC_MAPE = 1, KRR_METRIC_SMAPE = 2, KRR_METRIC_REL_MSE = 3 };
int krr_error_metric_f32(const float *film, const float *reference, int64_t n_pixels, int32_t metric,
						 double *result_host, void *cuda_stream);

/* ---- ToneMappingPass, src/render/passes/tonemapping/tonemapping.cu:11-85 ----
 * film: device float4[n_pixels], rewritten in place: rgb * exposure -> operator -> optional pow(1/2.2); alpha = 1 */
enum { KRR_TONEMAP_LINEAR = 0, KRR_TONEMAP_REINHARD = 1, KRR_TONEMAP_ACES = 2, KRR_TONEMAP_UNCHARTED2 = 3, KRR_TONEMAP_HEJIHABLE = 4 };
int krr_tonemap_f32(float *film, int64_t n_pixels, int32_t tonemap_operator, float exposure, int32_t use_gamma, void *cuda_stream);

const char *krr_wfpt_last_error(void);
int			krr_wfpt_abi_version(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* KRR_WFPT_H */
