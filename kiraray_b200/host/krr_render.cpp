// krr_render -- headless counterpart of the reference's executable (src/main/kiraray.cpp:5-32):
//   krr_render <config.json> [max_frames] [film.exr|film.pfm]
// environment: KRR_DATA_DIR = directory of spectral_srgb.bin; KRR_ASSET_ROOT = directory model paths are resolved
// against (default: the config's directory, as in the reference)
// loads the config (RenderApp::loadConfigFrom), runs the frame loop until a pass asks for the exit (AccumulatePass
// "exit_on_finish" with a spent "task" budget) or max_frames, finalises the passes ("save_on_finish", ErrorMeasure
// "save") and optionally writes the accumulated (or last) film.  Everything goes through the C entry points of the
// host layer (include/krr_host_c.h); there is no window, UI or swap chain.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "krr_host_c.h"

static int fail(const char *what) {
	std::fprintf(stderr, "krr_render: %s: %s\n", what, krr_host_last_error());
	return EXIT_FAILURE;
}

int main(int argc, char **argv) {
	if (argc < 2) {
		std::fprintf(stderr, "usage: krr_render <config.json> [max_frames] [film.exr|film.pfm]\n");
		return EXIT_FAILURE;
	}
	const char *config  = argv[1];
	const int maxFrames = argc > 2 ? std::atoi(argv[2]) : 0;
	const char *out		= argc > 3 ? argv[3] : nullptr;
	if (const char *dir = std::getenv("KRR_DATA_DIR")) krr_host_set_data_dir(dir);

	KrrHostApp *app = nullptr;
	if (krr_host_app_create(config, 1, std::getenv("KRR_ASSET_ROOT"), &app) != KRR_OK) return fail("loading the config");
	std::fprintf(stderr, "krr_render: using config file %s\n", config);
	const int frames = krr_host_app_run(app, maxFrames, /*finalize=*/1);
	if (frames < 0) {
		krr_host_app_destroy(app);
		return fail("rendering");
	}
	std::fprintf(stderr, "krr_render: %d frames rendered\n", frames);
	int rc = EXIT_SUCCESS;
	if (out) {
		int32_t w = 0, h = 0;
		krr_host_app_get_resolution(app, &w, &h);
		std::vector<float> film((size_t) w * h * 4);
		// the accumulated average when an AccumulatePass ran, else the last frame
		if (!(krr_host_app_accum_count(app) > 0 && krr_host_app_read_accumulated(app, film.data()) == KRR_OK) &&
			krr_host_app_render_frames(app, 0, film.data()) != KRR_OK)
			rc = fail("reading the film");
		else if (krr_host_image_save(out, film.data(), w, h, /*flip=*/0, /*reference_channel_order=*/1) != KRR_OK)
			rc = fail("writing the film");
	}
	krr_host_app_destroy(app);
	return rc;
}
