/* krr_host_c.h -- C entry points of the C++17 host layer (kiraray_b200/host), for drivers that are
 * not C++ (the Python tests and bench.py bind these with ctypes).  The host layer mirrors the
 * reference's RenderApp/RenderPass/SceneImporter surface (src/main/renderer.cpp:84-122, 258-316;
 * src/core/renderpass.h:138-273; src/scene/krrscene.cpp:8-349); loading a config touches no GPU. */
#ifndef KRR_HOST_C_H
#define KRR_HOST_C_H
#include "krr_wfpt.h"
#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default)
#endif

typedef struct KrrHostApp KrrHostApp;

/* directory holding spectral_srgb.bin (colour-space tables) */
int krr_host_set_data_dir(const char *dir);
/* RenderApp::loadConfigFrom / loadConfig.  config: path to a JSON file (is_path != 0) or JSON text;
 * asset_root: directory model paths are resolved against (NULL: the config's directory). */
int krr_host_app_create(const char *config, int is_path, const char *asset_root, KrrHostApp **out);
void krr_host_app_destroy(KrrHostApp *app);
int krr_host_app_get_resolution(KrrHostApp *app, int32_t *w, int32_t *h);
int krr_host_app_set_resolution(KrrHostApp *app, int32_t w, int32_t h);
/* flat scene view (valid until the app is destroyed or the scene changes) */
const KrrSceneDesc *krr_host_app_scene_desc(KrrHostApp *app);
/* camera as the pass will receive it at the current resolution (Scene::update + aspect ratio) */
int krr_host_app_get_camera(KrrHostApp *app, double time_seconds, KrrCameraData *out);
/* WavefrontPathTracer params as JSON (integrator.h:86-94 keys + "spp"); returns length or <0 */
int krr_host_app_get_wfpt_params(KrrHostApp *app, char *buf, int32_t capacity);
int krr_host_app_set_wfpt_params(KrrHostApp *app, const char *params_json);
/* GPU: RenderApp::initialize (first call) + n_frames x {++frameIndex; beginFrame; render; endFrame}
 * over all enabled passes; then reads the film back (film_host may be NULL). */
int krr_host_app_render_frames(KrrHostApp *app, int32_t n_frames, float *film_host);
/* the pass's C-ABI handle (NULL before the first render) and frame counter */
KrrWfpt *krr_host_app_wfpt_handle(KrrHostApp *app);
uint64_t krr_host_app_frame_index(KrrHostApp *app);
/* RenderApp main loop without a window (DeviceManager::runMessageLoop, window.cpp:450-485): frames until a
 * pass requests the exit (AccumulatePass "exit_on_finish" with a spent "task" budget) or max_frames (0 = no
 * cap); then RenderApp::finalize -> finalize() on every pass ("save_on_finish", ErrorMeasure "save").
 * Returns the number of frames rendered (>= 0) or a negative error. */
int krr_host_app_run(KrrHostApp *app, int32_t max_frames, int32_t finalize);
/* File::outputDir() of this app ("output_dir" of the config, else <config dir>/output) */
int krr_host_app_set_output_dir(KrrHostApp *app, const char *dir);
/* a pass's parameters as JSON (to_json of the pass), by pass name; returns length or < 0 */
int krr_host_app_get_pass_json(KrrHostApp *app, const char *pass_name, char *buf, int32_t capacity);
/* AccumulatePass: frames accumulated so far / accumulated average read back (RGBA32F, film layout) */
int64_t krr_host_app_accum_count(KrrHostApp *app);
int krr_host_app_read_accumulated(KrrHostApp *app, float *rgba_host);
/* ErrorMeasurePass: set the reference from memory (film layout), ask for an evaluation in the next frame
 * (the UI's "Evaluate" button), read the last result; value_out may be NULL when nothing was evaluated */
int krr_host_app_set_reference(KrrHostApp *app, const float *rgba_host, int32_t w, int32_t h);
int krr_host_app_evaluate_next_frame(KrrHostApp *app);
int krr_host_app_last_error_metric(KrrHostApp *app, double *value_out, int32_t *n_evaluations_out);

/* HDR image files (Image::loadImage / saveImage, src/core/texture.cpp:27-118): .exr and .pfm, no GPU.
 * load: first call with rgba_host = NULL to get the size.  flip: vertical flip as in the reference's API.
 * save: reference_channel_order != 0 writes EXR planes the way KiRaRay's save_exr does (image.cpp). */
int krr_host_image_load(const char *path, int32_t flip, int32_t *w, int32_t *h, float *rgba_host);
int krr_host_image_save(const char *path, const float *rgba_host, int32_t w, int32_t h, int32_t flip, int32_t reference_channel_order);
/* EXR writer with explicit storage options (half / float, none / ZIP) */
int krr_host_image_save_exr(const char *path, const float *rgba_host, int32_t w, int32_t h, int32_t half_precision, int32_t zip);
const char *krr_host_last_error(void);

/* One process, N GPUs (kiraray_b200/host/multi_device.cpp MultiDeviceRenderApp): one pass handle per device, one host
 * thread per handle, work = `tiles` image tiles x (n_devices / tiles) spp slices, film sum-reduce onto rank 0 with
 * krr_wfpt_reduce_film (NCCL over NVLink).  devices[i] = CUDA device of rank i; ranks that share a device (single-GPU
 * test boxes) are reduced on the host instead.  krr_multi_render: `steps` frames per spp slice from first_frame on;
 * the reduced film of the last step is written to film_host (RGBA32F, may be NULL). */
typedef struct KrrMultiApp KrrMultiApp;
int krr_multi_create(const KrrSceneDesc *scene, const char *params_json, int32_t w, int32_t h, const int32_t *devices, int32_t n_devices, int32_t tiles, KrrMultiApp **out);
void krr_multi_destroy(KrrMultiApp *app);
int krr_multi_uses_nccl(KrrMultiApp *app);
int krr_multi_render(KrrMultiApp *app, const KrrCameraData *cam, uint64_t first_frame, int32_t steps, float *film_host, double *ms_total, uint64_t *rays_last_step);
const char *krr_multi_last_error(void);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif
