// multi_device.cpp -- one process, N GPUs: the multi-GPU driver of the host layer.
//
// The reference is single-device (src/core/device/context.cpp:37-40 pins device 0).  Pixels and frames are
// independent in its integrator (per-pixel PCG stream and accumulator, integrator.cpp:213-220; frames are averaged
// by AccumulatePass, accumulate.cu:30-52), so the work of a render is a T x S grid: T image tiles (contiguous row
// bands, krr_wfpt_set_partition: a rank renders its rows and writes zeros elsewhere, so tile films ADD) and S spp
// slices (rank s renders frame indices first + (s + k S) F ...), with the scene replicated on every GPU.  The one
// exchange step is the film accumulation: ncclReduce over NVLink onto rank 0 (krr_wfpt_reduce_film), x 1/S.
//
// MultiDeviceRenderApp owns one KrrWfpt handle per device and drives each from its own host thread (a handle is
// not thread-safe, different handles are independent: include/krr_wfpt.h); the communicator is created for all
// handles at once (krr_wfpt_comm_init_all).  Rank 0 reads the reduced film back with the pipelined
// krr_wfpt_render_reduce_to_host_async, so the read-back of step k overlaps the rendering of step k + 1.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>

#include "krr_host.h"
#include "krr_host_c.h"

namespace krr {

namespace {
struct Barrier { // C++17: no std::barrier
	std::mutex m;
	std::condition_variable cv;
	int n, count = 0, gen = 0;
	explicit Barrier(int n_) : n(n_) {}
	void wait() {
		std::unique_lock<std::mutex> lk(m);
		const int g = gen;
		if (++count == n) { count = 0, gen++; cv.notify_all(); }
		else cv.wait(lk, [&] { return gen != g; });
	}
};
} // namespace

struct MultiDeviceRenderApp::Impl {
	struct Rank {
		int device = 0, tile = 0, slice = 0, rowBegin = 0, rowEnd = 0;
		KrrWfpt *h = nullptr;
		cudaStream_t stream = nullptr;
		float *film = nullptr; // device film (same-device fallback only)
		string error;
	};
	std::vector<Rank> ranks;
	int width = 0, height = 0, tiles = 1, slices = 1, frameBatch = 1;
	bool nccl = false;
};

MultiDeviceRenderApp::MultiDeviceRenderApp() : m(new Impl) {}
MultiDeviceRenderApp::~MultiDeviceRenderApp() {
	for (auto &r : m->ranks) {
		cudaSetDevice(r.device);
		if (r.h) krr_wfpt_destroy(r.h);
		if (r.stream) cudaStreamDestroy(r.stream);
		if (r.film) cudaFree(r.film);
	}
	delete m;
}

int MultiDeviceRenderApp::size() const { return (int) m->ranks.size(); }
KrrWfpt *MultiDeviceRenderApp::handle(int rank) { return m->ranks[rank].h; }
bool MultiDeviceRenderApp::usesNccl() const { return m->nccl; }

void MultiDeviceRenderApp::init(const KrrSceneDesc *scene, const string &paramsJson, int width, int height, const std::vector<int> &devices, int tiles) {
	const int n = (int) devices.size();
	if (n < 1) throw std::runtime_error("MultiDeviceRenderApp: no devices");
	if (tiles < 1 || n % tiles) throw std::runtime_error("MultiDeviceRenderApp: tiles must divide the number of devices");
	if (tiles > height) throw std::runtime_error("MultiDeviceRenderApp: more tiles than rows");
	m->width = width, m->height = height, m->tiles = tiles, m->slices = n / tiles;
	m->frameBatch = 1;
	if (!paramsJson.empty()) m->frameBatch = std::max(1, json::parse(paramsJson).value("frame_batch", 1));
	m->ranks.resize(n);
	bool distinct = true;
	for (int i = 0; i < n; i++)
		for (int j = 0; j < i; j++) distinct &= devices[i] != devices[j];
	// every rank sets itself up on its own thread (scene upload + BVH build run concurrently on the devices)
	std::vector<std::thread> th;
	for (int i = 0; i < n; i++) {
		Impl::Rank &r = m->ranks[i];
		r.device = devices[i], r.tile = i % tiles, r.slice = i / tiles;
		r.rowBegin = (int) ((int64_t) r.tile * height / tiles), r.rowEnd = (int) ((int64_t) (r.tile + 1) * height / tiles);
		th.emplace_back([&r, scene, &paramsJson, width, height, tiles] {
			auto ck = [&](int rc, const char *what) { if (rc != KRR_OK && r.error.empty()) r.error = string(what) + ": " + krr_wfpt_last_error(); return rc == KRR_OK; };
			if (cudaSetDevice(r.device) != cudaSuccess) { r.error = "cudaSetDevice failed"; return; }
			if (!ck(krr_wfpt_create(paramsJson.empty() ? nullptr : paramsJson.c_str(), &r.h), "create")) return;
			if (!ck(krr_wfpt_set_color_space(r.h, &defaultColorSpace()), "set_color_space")) return;
			if (!ck(krr_wfpt_set_scene(r.h, scene), "set_scene")) return;
			if (!ck(krr_wfpt_resize(r.h, width, height), "resize")) return;
			if (tiles > 1 && !ck(krr_wfpt_set_partition(r.h, r.rowBegin, r.rowEnd), "set_partition")) return;
			cudaStreamCreateWithFlags(&r.stream, cudaStreamNonBlocking);
		});
	}
	for (auto &t : th) t.join();
	for (auto &r : m->ranks)
		if (!r.error.empty()) throw std::runtime_error("MultiDeviceRenderApp rank on device " + std::to_string(r.device) + ": " + r.error);
	if (n > 1 && distinct) {
		std::vector<KrrWfpt *> hs;
		for (auto &r : m->ranks) hs.push_back(r.h);
		if (krr_wfpt_comm_init_all(hs.data(), n) != KRR_OK) throw std::runtime_error(string("krr_wfpt_comm_init_all: ") + krr_wfpt_last_error());
		m->nccl = true;
	} else if (n > 1) {
		// several ranks on ONE device (a single-GPU box: the thread-safety tests): no NCCL rank may share a
		// device, the films are summed on the host instead
		for (auto &r : m->ranks) {
			cudaSetDevice(r.device);
			if (cudaMalloc((void **) &r.film, (size_t) width * height * 16) != cudaSuccess) throw std::runtime_error("MultiDeviceRenderApp: film allocation failed");
		}
	}
}

MultiDeviceRenderApp::Result MultiDeviceRenderApp::render(const KrrCameraData &cam, uint64_t firstFrame, int steps, float *filmHost) {
	const int n = size();
	const size_t nFloats = (size_t) m->width * m->height * 4;
	Result res;
	if (steps < 1) return res;
	Barrier barrier(n);
	std::vector<double> ms(n, 0.0);
	std::vector<uint64_t> rays(n, 0);
	std::vector<std::vector<float>> hostFilms(m->nccl || n == 1 ? 0 : n);
	for (auto &f : hostFilms) f.resize(nFloats);
	std::vector<float> scratch[2];
	if (!filmHost) scratch[0].resize(nFloats);
	float *out = filmHost ? filmHost : scratch[0].data();
	const float scale = 1.f / (float) m->slices;
	std::vector<std::thread> th;
	for (int i = 0; i < n; i++)
		th.emplace_back([&, i] {
			Impl::Rank &r = m->ranks[i];
			auto ck = [&](int rc, const char *what) { if (rc != KRR_OK && r.error.empty()) r.error = string(what) + ": " + krr_wfpt_last_error(); };
			cudaSetDevice(r.device);
			barrier.wait();
			const auto t0 = std::chrono::steady_clock::now();
			for (int k = 0; k < steps; k++) {
				const uint64_t frame = firstFrame + ((uint64_t) r.slice + (uint64_t) k * m->slices) * m->frameBatch;
				ck(krr_wfpt_begin_frame(r.h, frame, &cam, r.stream), "begin_frame");
				if (m->nccl || n == 1) ck(krr_wfpt_render_reduce_to_host_async(r.h, i == 0 ? out : nullptr, 0, scale, r.stream), "render_reduce_to_host_async");
				else {
					ck(krr_wfpt_render(r.h, r.film, r.stream), "render");
					cudaMemcpyAsync(hostFilms[i].data(), r.film, nFloats * 4, cudaMemcpyDeviceToHost, r.stream);
				}
			}
			cudaStreamSynchronize(r.stream);
			ck(krr_wfpt_wait_host(r.h), "wait_host");
			ms[i] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
			KrrStats st;
			if (krr_wfpt_get_stats(r.h, &st) == KRR_OK) rays[i] = st.closest_rays + st.shadow_rays;
			else ck(KRR_E_CUDA, "get_stats");
		});
	for (auto &t : th) t.join();
	for (auto &r : m->ranks)
		if (!r.error.empty()) { string e = r.error; r.error.clear(); throw std::runtime_error("MultiDeviceRenderApp: " + e); }
	if (!hostFilms.empty()) { // same-device ranks: the reduction on the host, in rank order
		for (size_t p = 0; p < nFloats; p++) {
			float s = hostFilms[0][p];
			for (int i = 1; i < n; i++) s += hostFilms[i][p];
			out[p] = s * scale;
		}
	}
	for (int i = 0; i < n; i++) res.msTotal = std::max(res.msTotal, ms[i]), res.raysLastStep += rays[i];
	res.steps = steps;
	return res;
}

} // namespace krr

// ---- C entry points (include/krr_host_c.h) ----
using namespace krr;
struct KrrMultiApp { MultiDeviceRenderApp app; };
static thread_local std::string gMultiErr;
extern "C" const char *krr_multi_last_error(void) { return gMultiErr.c_str(); }

extern "C" int krr_multi_create(const KrrSceneDesc *scene, const char *params_json, int32_t w, int32_t h, const int32_t *devices, int32_t n_devices, int32_t tiles, KrrMultiApp **out) {
	if (!scene || !devices || !out || n_devices < 1 || w <= 0 || h <= 0) { gMultiErr = "bad argument"; return KRR_E_INVALID; }
	auto *a = new KrrMultiApp();
	try {
		a->app.init(scene, params_json ? params_json : "", w, h, std::vector<int>(devices, devices + n_devices), tiles);
	} catch (const std::exception &e) { gMultiErr = e.what(); delete a; return KRR_E_INVALID; }
	*out = a;
	return KRR_OK;
}
extern "C" void krr_multi_destroy(KrrMultiApp *a) { delete a; }
extern "C" int krr_multi_uses_nccl(KrrMultiApp *a) { return a && a->app.usesNccl() ? 1 : 0; }
extern "C" int krr_multi_render(KrrMultiApp *a, const KrrCameraData *cam, uint64_t first_frame, int32_t steps, float *film_host, double *ms_total, uint64_t *rays_last_step) {
	if (!a || !cam) { gMultiErr = "bad argument"; return KRR_E_INVALID; }
	try {
		auto r = a->app.render(*cam, first_frame, steps, film_host);
		if (ms_total) *ms_total = r.msTotal;
		if (rays_last_step) *rays_last_step = r.raysLastStep;
	} catch (const std::exception &e) { gMultiErr = e.what(); return KRR_E_CUDA; }
	return KRR_OK;
}
