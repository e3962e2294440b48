// host_c_api.cpp -- C entry points of the host layer (include/krr_host_c.h).
#include "krr_host.h"
#include "krr_host_c.h"

#include <algorithm>
#include <cstring>

using namespace krr;

struct KrrHostApp {
	RenderApp app;
};

static thread_local std::string gErr;
#define KRR_TRY try {
#define KRR_CATCH } catch (const std::exception &e) { gErr = e.what(); return KRR_E_INVALID; } return KRR_OK;

extern "C" const char *krr_host_last_error(void) { return gErr.c_str(); }

extern "C" int krr_host_set_data_dir(const char *dir) { setDataDir(dir ? dir : ""); return KRR_OK; }

extern "C" int krr_host_app_create(const char *config, int is_path, const char *asset_root, KrrHostApp **out) {
	KRR_TRY
	if (!config || !out) throw std::runtime_error("null argument");
	auto *a = new KrrHostApp();
	try {
		if (is_path && !asset_root) a->app.loadConfigFrom(config);
		else {
			json j;
			std::string base = asset_root ? asset_root : ".";
			if (is_path) {
				FILE *f = fopen(config, "rb");
				if (!f) throw std::runtime_error(std::string("cannot open config ") + config);
				std::string text;
				char buf[4096];
				size_t n;
				while ((n = fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
				fclose(f);
				j = json::parse(text);
			} else j = json::parse(config);
			a->app.loadConfig(j, base);
		}
		if (!a->app.scene()) throw std::runtime_error("config has no scene");
	} catch (...) { delete a; throw; }
	*out = a;
	KRR_CATCH
}

extern "C" void krr_host_app_destroy(KrrHostApp *app) { delete app; }

extern "C" int krr_host_app_get_resolution(KrrHostApp *app, int32_t *w, int32_t *h) {
	*w = app->app.frameSize().x, *h = app->app.frameSize().y;
	return KRR_OK;
}

extern "C" int krr_host_app_set_resolution(KrrHostApp *app, int32_t w, int32_t h) {
	KRR_TRY
	if (w <= 0 || h <= 0) throw std::runtime_error("bad resolution");
	// before initialize() this only records the size; afterwards it resizes film and passes
	struct Peek : RenderApp { using RenderApp::RenderApp; };
	if (app->app.context()->getColorDevice()) app->app.resize(Vector2i{w, h});
	else {
		json j = json::object();
		json r = json::array();
		r.push_back(json(w)), r.push_back(json(h));
		j["resolution"] = r;
		app->app.loadConfig(j, ".");
	}
	KRR_CATCH
}

extern "C" const KrrSceneDesc *krr_host_app_scene_desc(KrrHostApp *app) {
	try { return &app->app.scene()->desc(); } catch (const std::exception &e) { gErr = e.what(); return nullptr; }
}

extern "C" int krr_host_app_get_camera(KrrHostApp *app, double t, KrrCameraData *out) {
	KRR_TRY
	auto s = app->app.scene();
	s->setAspectRatio((float) app->app.frameSize().x / app->app.frameSize().y);
	s->update(app->app.frameIndex(), t);
	*out = s->camera;
	KRR_CATCH
}

extern "C" int krr_host_app_get_wfpt_params(KrrHostApp *app, char *buf, int32_t cap) {
	auto p = app->app.findPass<WavefrontPathTracer>();
	if (!p) { gErr = "config has no WavefrontPathTracer pass"; return KRR_E_STATE; }
	std::string s = p->toJson().dump();
	if ((int) s.size() + 1 > cap) return KRR_E_INVALID;
	memcpy(buf, s.c_str(), s.size() + 1);
	return (int) s.size();
}

extern "C" int krr_host_app_set_wfpt_params(KrrHostApp *app, const char *params) {
	KRR_TRY
	auto p = app->app.findPass<WavefrontPathTracer>();
	if (!p) throw std::runtime_error("config has no WavefrontPathTracer pass");
	json cur = p->toJson(), upd = json::parse(params);
	for (auto &kv : upd.members()) cur[kv.first] = kv.second;
	p->fromJson(cur);
	KRR_CATCH
}

extern "C" int krr_host_app_render_frames(KrrHostApp *app, int32_t n, float *film) {
	KRR_TRY
	for (int i = 0; i < n; i++) app->app.renderFrame(0.0);
	if (film) {
		std::vector<float> host;
		app->app.readFilm(host);
		memcpy(film, host.data(), host.size() * 4);
	}
	KRR_CATCH
}

extern "C" KrrWfpt *krr_host_app_wfpt_handle(KrrHostApp *app) {
	auto p = app->app.findPass<WavefrontPathTracer>();
	return p ? p->handle() : nullptr;
}

extern "C" uint64_t krr_host_app_frame_index(KrrHostApp *app) { return app->app.frameIndex(); }

extern "C" int krr_host_app_run(KrrHostApp *app, int32_t maxFrames, int32_t finalize) {
	try {
		size_t n = app->app.run((size_t) std::max(maxFrames, 0));
		if (finalize) app->app.finalize();
		return (int) n;
	} catch (const std::exception &e) { gErr = e.what(); return KRR_E_INVALID; }
}

extern "C" int krr_host_app_set_output_dir(KrrHostApp *app, const char *dir) {
	KRR_TRY
	if (!dir) throw std::runtime_error("null argument");
	app->app.setOutputDir(dir);
	KRR_CATCH
}

extern "C" int krr_host_app_get_pass_json(KrrHostApp *app, const char *name, char *buf, int32_t cap) {
	try {
		for (auto &p : app->app.passes())
			if (p->getName() == name) {
				std::string s = p->toJson().dump();
				if ((int) s.size() + 1 > cap) throw std::runtime_error("buffer too small");
				memcpy(buf, s.c_str(), s.size() + 1);
				return (int) s.size();
			}
		throw std::runtime_error(std::string("no pass named ") + name);
	} catch (const std::exception &e) { gErr = e.what(); return KRR_E_INVALID; }
}

extern "C" int64_t krr_host_app_accum_count(KrrHostApp *app) {
	auto p = app->app.findPass<AccumulatePass>();
	return p ? (int64_t) p->accumCount() : -1;
}

extern "C" int krr_host_app_read_accumulated(KrrHostApp *app, float *rgba) {
	KRR_TRY
	auto p = app->app.findPass<AccumulatePass>();
	if (!p || !rgba) throw std::runtime_error("no AccumulatePass / null buffer");
	Image img;
	if (!p->readAverage(img)) throw std::runtime_error("nothing accumulated yet");
	memcpy(rgba, img.rgba.data(), img.rgba.size() * 4);
	KRR_CATCH
}

extern "C" int krr_host_app_set_reference(KrrHostApp *app, const float *rgba, int32_t w, int32_t h) {
	KRR_TRY
	auto p = app->app.findPass<ErrorMeasurePass>();
	if (!p || !rgba || w <= 0 || h <= 0) throw std::runtime_error("no ErrorMeasurePass / bad image");
	Image img;
	img.width = w, img.height = h;
	img.rgba.assign(rgba, rgba + (size_t) w * h * 4);
	p->setReferenceImage(img);
	KRR_CATCH
}

extern "C" int krr_host_app_evaluate_next_frame(KrrHostApp *app) {
	KRR_TRY
	auto p = app->app.findPass<ErrorMeasurePass>();
	if (!p) throw std::runtime_error("no ErrorMeasurePass");
	p->evaluateNextFrame();
	KRR_CATCH
}

extern "C" int krr_host_app_last_error_metric(KrrHostApp *app, double *value, int32_t *nEval) {
	KRR_TRY
	auto p = app->app.findPass<ErrorMeasurePass>();
	if (!p) throw std::runtime_error("no ErrorMeasurePass");
	if (value) *value = p->lastValue();
	if (nEval) *nEval = (int32_t) p->results().size();
	KRR_CATCH
}

extern "C" int krr_host_image_load(const char *path, int32_t flip, int32_t *w, int32_t *h, float *rgba) {
	KRR_TRY
	if (!path || !w || !h) throw std::runtime_error("null argument");
	Image img;
	std::string err;
	if (!loadImage(path, img, flip != 0, &err)) throw std::runtime_error(err);
	*w = img.width, *h = img.height;
	if (rgba) memcpy(rgba, img.rgba.data(), img.rgba.size() * 4);
	KRR_CATCH
}

static Image wrapImage(const float *rgba, int32_t w, int32_t h) {
	if (!rgba || w <= 0 || h <= 0) throw std::runtime_error("bad image");
	Image img;
	img.width = w, img.height = h;
	img.rgba.assign(rgba, rgba + (size_t) w * h * 4);
	return img;
}

extern "C" int krr_host_image_save(const char *path, const float *rgba, int32_t w, int32_t h, int32_t flip, int32_t refOrder) {
	KRR_TRY
	if (!path) throw std::runtime_error("null argument");
	std::string err;
	if (!saveImage(path, wrapImage(rgba, w, h), flip != 0, &err, refOrder != 0)) throw std::runtime_error(err);
	KRR_CATCH
}

extern "C" int krr_host_image_save_exr(const char *path, const float *rgba, int32_t w, int32_t h, int32_t half, int32_t zip) {
	KRR_TRY
	if (!path) throw std::runtime_error("null argument");
	std::string err;
	if (!saveEXR(path, wrapImage(rgba, w, h), half != 0, zip != 0, &err)) throw std::runtime_error(err);
	KRR_CATCH
}
