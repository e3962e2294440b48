/* driver.h -- TEST INFRASTRUCTURE ONLY.
 * C API of the CPU oracle's wavefront driver (oracle/driver.cpp).  The driver restates the stage
 * bodies of the reference's WavefrontPathTracer (src/render/wavefront/integrator.cpp, device.cu,
 * medium.cpp, render/shading.h) on the CPU, depth-first per pixel, on top of oracle_leaf.h.
 */
#pragma once
#include <stdint.h>
#include "krr_wfpt.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct OrcParams {
	int32_t nee, enable_medium, max_depth, enable_clamp, spp;
	float	rr, clamp_max;
	int32_t use_bvh;  /* 0: brute-force intersector (ID validation); 1: median-split BVH over the static instances, EVERY moving instance
					   * visited (what the kernels' motion boxes are verified against); 2: + a BVH over the moving instances' boxes for
					   * the render's ray-time window (proven motion bound restated on the host: the fast CPU baseline) */
	int32_t threads;  /* 0 = all cores */
	int32_t row_begin, row_end;		/* rows to render, row_end <= 0 -> all */
} OrcParams;

typedef struct OrcScene OrcScene;

OrcScene *orc_scene_create(const KrrSceneDesc *desc);
void	  orc_scene_destroy(OrcScene *s);
int32_t	  orc_scene_num_lights(const OrcScene *s);

/* Renders one frame (beginFrame + render, integrator.cpp:205-267).
 *  film            : w*h*4 floats, row H-1-y (device/cuda.h:33-36), may be NULL
 *  first_hits      : w*h*2 int32 (instance, primitive) of the depth-0 hit of the LAST sample, may be NULL
 *  sampler_state   : w*h*2 uint64 after beginFrame (before the first camera sample), may be NULL
 *  lambda          : w*h*4 floats, may be NULL
 *  camera_sample   : w*h*5 floats of the LAST sample, may be NULL
 * capture: if capture_sample >= 0, integer fields (pixelId, depth, bsdfType, aux) of the items each
 * queue holds at (capture_sample, capture_depth) are appended to cap_items[q] (q: 0 ray, 1 miss,
 * 2 hitLight, 3 scatter, 4 shadow, 5 next ray); cap_counts[q] receives the counts. cap_items[q]
 * must hold w*h*4 int32 each.
 * returns wall seconds spent in the render loop (scene build excluded), < 0 on error. */
double orc_render(const OrcScene *s, const OrcParams *p, const KrrCameraData *cam, int32_t w, int32_t h,
				  uint64_t frame_index, float *film, int32_t *first_hits, uint64_t *sampler_state,
				  float *lambda, float *camera_sample, KrrStats *stats, int32_t capture_sample,
				  int32_t capture_depth, int32_t *cap_items[6], int32_t cap_counts[6]);

/* object->world / world->object of an instance for a ray carrying `time` (the motion spec) */
void orc_instance_xf(const OrcScene *s, int32_t inst, float time, float m[12], float inv[12]);

/* the build's own ray/triangle routine, exposed so tests can compare it bit-for-bit with the GPU's:
 * returns 1 on hit and writes t,u,v */
int orc_intersect_triangle(const float o[3], const float d[3], const float v0[3], const float v1[3],
						   const float v2[3], float tmax, float *t, float *u, float *v);

#ifdef __cplusplus
}
#endif
