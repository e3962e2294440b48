// krr_host.h -- C++17 host layer that keeps the reference's plugin surface for the
// WavefrontPathTracer pass and talks to the CUDA kernels ONLY through the C ABI (include/krr_wfpt.h).
//
// Mirrors (same names, argument meaning and defaults):
//   RenderPass / RenderPassFactory / KRR_REGISTER_PASS_DEC|DEF   reference src/core/renderpass.h:138-273
//   RenderContext (film = RGBA32F colour target)                 src/core/renderpass.h:19-136
//   WavefrontPathTracer + its JSON params                        src/render/wavefront/integrator.h:24-104
//   RenderApp::loadConfig / render loop (headless)               src/main/renderer.cpp:84-122, 258-316
//   SceneImporter (KRR JSON scene schema)                        src/scene/krrscene.cpp:8-349
//   OBJ/MTL material mapping                                     src/scene/assimp.cpp:60-65, 93-226
//   Camera / OrbitCameraController                               src/core/camera.{h,cpp}
//   AccumulatePass / ErrorMeasurePass / ToneMappingPass          src/render/passes/{accumulate,errormeasure,tonemapping}
//   Image (EXR / PFM load + save)                                src/core/texture.cpp:27-118, src/util/image.cpp
// Windowing, Vulkan interop, UI and the other importers are out of scope.
#pragma once
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "json.h"
#include "krr_wfpt.h"

namespace krr {

using json	 = Json;
using string = std::string;

struct Vector2i { int x = 0, y = 0; int operator[](int i) const { return i ? y : x; } };

// ------------------------------------------------------------------------------------------------
// Scene: flattened host scene graph (instances carry global transforms), owns all arrays the
// KrrSceneDesc view points into.
// ------------------------------------------------------------------------------------------------
struct HostMesh {
	string name;
	std::vector<float> positions, normals, texcoords, tangents;
	std::vector<int32_t> indices;
	int material = -1, mediumInside = -1, mediumOutside = -1;
	float Le[3]	 = {0, 0, 0};
};

// one animated ancestor of an instance: world = ... * pre * SRT(t) * ...
struct AnimLink {
	float pre[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}; // static transforms between the previous animated node and this one
	std::vector<float> times;
	std::vector<KrrSRT> keys;
};

struct HostInstance {
	int mesh = 0;
	float transform[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
	std::vector<KrrSRT> motionKeys;
	// simple keyframe animation (reference src/core/animation.h): linear SRT keys over time
	std::vector<float> animTimes;
	std::vector<KrrSRT> animKeys;
	// transform of the (static) ancestors of an animated node: world = animParent * SRT(t)
	float animParent[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
	// animated ancestors, root first: world = prod(a.pre * SRT_a(t)) * animParent * [SRT(t) if the node has keys]
	std::vector<AnimLink> animAncestors;
};

struct HostMaterial {
	string name;
	KrrMaterialDesc desc{};
	std::vector<float> etaLambdas, etaValues, kLambdas, kValues;
	std::vector<std::vector<float>> images; // per texture slot
};

struct HostMedium {
	string name;
	KrrMediumDesc desc{};
	std::vector<float> density;
	std::vector<float> albedoGrid; // optional RGB albedo per voxel (KrrMediumDesc::albedo_grid)
	float boundMin[3] = {0, 0, 0}, boundMax[3] = {0, 0, 0};
	bool hasBound = false;
};

// OrbitCameraController::CameraControllerData, src/core/camera.h:149-156
struct CameraControllerData {
	float target[3] = {0, 0, 0};
	float radius = 5, pitch = 0, yaw = 0;
};

class Scene {
public:
	using SharedPtr = std::shared_ptr<Scene>;

	std::vector<HostMesh> meshes;
	std::vector<HostInstance> instances;
	std::vector<HostMaterial> materials;
	std::vector<KrrLightDesc> lights;
	std::vector<std::vector<float>> lightImages; // RGBA32F lat-long images of infinite lights (lights[i].texture.image points here)
	std::vector<HostMedium> media;
	KrrSceneOptions options{1, 0, 0, 0.f, 1.f};
	KrrCameraData camera;
	CameraControllerData cameraController;
	bool hasCameraController = false;
	bool animated			 = false;

	Scene();
	// Scene::update, src/core/scene.cpp:17-31: animate -> camera controller -> camera
	bool update(size_t frameIndex, double currentTime);
	void setAspectRatio(float aspect); // renderer.cpp:28 + Camera::update (camera.cpp:6-17)
	// flat view for the C ABI; pointers stay valid until the scene is modified
	const KrrSceneDesc &desc();
	void boundingBox(float lo[3], float hi[3]) const;
	uint64_t version() const { return mVersion; }
	void touch() { mVersion++; }
	// instance ids whose transform changed in the last update() (drives TLAS refit)
	std::vector<int32_t> updatedInstances;

private:
	KrrSceneDesc mDesc{};
	std::vector<KrrMeshDesc> mMeshDescs;
	std::vector<KrrInstanceDesc> mInstanceDescs;
	std::vector<KrrMaterialDesc> mMaterialDescs;
	std::vector<KrrMediumDesc> mMediumDescs;
	uint64_t mVersion = 1;
};

// ------------------------------------------------------------------------------------------------
// Importers
// ------------------------------------------------------------------------------------------------
class SceneImporter {
public:
	// src/scene/krrscene.cpp:253-305
	static bool import(const json &j, Scene::SharedPtr scene, const string &baseDir);
	static bool loadModel(const string &filepath, Scene::SharedPtr scene, const float nodeTransform[12],
						  const string &baseDir);
	// the "environment" key of a scene / application config (krrscene.cpp:267-274, renderer.cpp:295-302):
	// an InfiniteLight with a lat-long image at the scene root
	static bool addEnvironment(const string &texture, Scene::SharedPtr scene, const string &baseDir);
};
bool loadObj(const string &filepath, Scene &scene, const float nodeTransform[12]);
// minimal glTF 2.0 (gltf.cpp): meshes, node hierarchy, pbrMetallicRoughness materials, PNG textures, TRS animation
bool loadGltf(const string &filepath, Scene &scene, const float nodeTransform[12]);

// ------------------------------------------------------------------------------------------------
// HDR images (image.cpp): RGBA32F, row 0 first -- the layout of the film buffer
// ------------------------------------------------------------------------------------------------
struct Image {
	int width = 0, height = 0;
	std::vector<float> rgba;
	bool isValid() const { return width > 0 && height > 0 && rgba.size() == (size_t) width * height * 4; }
	void flipVertically();
};
// Image::loadImage(path, flip, srgb) / saveImage(path, flip), texture.cpp:27-118: .exr and .pfm
bool loadImage(const string &path, Image &img, bool flip, string *err = nullptr);
// referenceChannelOrder: write the EXR channels the way the reference's save_exr does (see image.cpp)
bool saveImage(const string &path, const Image &img, bool flip, string *err = nullptr, bool referenceChannelOrder = true);
bool loadEXR(const string &path, Image &img, string *err = nullptr);
bool saveEXR(const string &path, const Image &img, bool halfPrecision, bool zip, string *err = nullptr);
bool loadPFM(const string &path, Image &img, string *err = nullptr);
bool loadPNG(const string &path, Image &img, bool srgb, string *err = nullptr); // 8-bit, non-interlaced (gltf.cpp)
bool savePFM(const string &path, const Image &img, string *err = nullptr);

// ------------------------------------------------------------------------------------------------
// RenderContext / RenderPass / factory
// ------------------------------------------------------------------------------------------------
class RenderContext {
public:
	// headless stand-in for RenderTexture/CudaRenderTarget (renderpass.h:19-136, device/cuda.h:17-53):
	// a linear device buffer of float4[W*H], RGBA32F like renderpass.cpp:44
	float *getColorDevice() const { return mColor; }
	Vector2i getSize() const { return mSize; }
	void *getStream() const { return mStream; }
	void resize(Vector2i size);
	void readback(std::vector<float> &host) const;
	~RenderContext();

private:
	float *mColor = nullptr;
	Vector2i mSize;
	void *mStream = nullptr;
};

class RenderApp;

class RenderPass {
public:
	using SharedPtr = std::shared_ptr<RenderPass>;
	RenderPass()		  = default;
	virtual ~RenderPass() = default;

	virtual void resize(const Vector2i &size) { mFrameSize = size; }
	virtual void setEnable(bool enable) { mEnable = enable; }
	virtual void setScene(Scene::SharedPtr scene) { mScene = scene; }
	virtual Scene::SharedPtr getScene() { return mScene; }
	virtual void tick(float elapsedSeconds) {}
	virtual void beginFrame(RenderContext *context) {}
	virtual void render(RenderContext *context) {}
	virtual void endFrame(RenderContext *context) {}
	virtual void renderUI() {}
	virtual void initialize() {}
	virtual void finalize() {}
	virtual bool isCudaPass() const { return true; }
	virtual string getName() const { return "RenderPass"; }
	virtual bool enabled() const { return mEnable; }
	virtual json toJson() const { return json::object(); }
	// the reference pulls these from DeviceManager (renderpass.cpp:120-126); the headless app sets them
	void setFrameIndex(size_t i) { mFrameIndex = i; }
	// gpContext->getGlobalConfig() / File::outputDir() of the reference, handed down by RenderApp
	void setApp(RenderApp *app) { mApp = app; }

protected:
	size_t getFrameIndex() const { return mFrameIndex; }
	Vector2i getFrameSize() const { return mFrameSize; }
	bool mEnable = true;
	Scene::SharedPtr mScene;
	size_t mFrameIndex = 0;
	Vector2i mFrameSize;
	RenderApp *mApp = nullptr;
};

class RenderPassFactory {
public:
	typedef std::map<string, std::function<RenderPass::SharedPtr(void)>> map_type;
	typedef std::map<string, std::function<RenderPass::SharedPtr(const json &)>> configured_map_type;
	static RenderPass::SharedPtr createInstance(std::string const &s);
	static RenderPass::SharedPtr deserizeInstance(std::string const &s, const json &serde);
	static std::shared_ptr<map_type> getMap();
	static std::shared_ptr<configured_map_type> getConfiguredMap();
};

template <typename T> class RenderPassRegister : RenderPassFactory {
public:
	RenderPassRegister(const string &s) {
		getMap()->insert(std::make_pair(s, []() -> RenderPass::SharedPtr { return std::make_shared<T>(); }));
		getConfiguredMap()->insert(std::make_pair(s, [](const json &j) -> RenderPass::SharedPtr {
			auto p = std::make_shared<T>();
			p->fromJson(j);
			return p;
		}));
	}
};
#define KRR_REGISTER_PASS_DEC(name) static RenderPassRegister<name> reg;
#define KRR_REGISTER_PASS_DEF(name) RenderPassRegister<name> name::reg(#name);

// ------------------------------------------------------------------------------------------------
// The pass (integrator.h:24-104).  Every method forwards to the C ABI.
// ------------------------------------------------------------------------------------------------
class WavefrontPathTracer : public RenderPass {
public:
	using SharedPtr = std::shared_ptr<WavefrontPathTracer>;
	KRR_REGISTER_PASS_DEC(WavefrontPathTracer);

	WavefrontPathTracer() = default;
	~WavefrontPathTracer() override;

	void resize(const Vector2i &size) override;
	void setScene(Scene::SharedPtr scene) override;
	void beginFrame(RenderContext *context) override;
	void render(RenderContext *context) override;
	void initialize() override;
	string getName() const override { return "WavefrontPathTracer"; }

	void fromJson(const json &j);
	json toJson() const override;
	KrrWfpt *handle() { return mHandle; }
	KrrStats stats();

	// path tracing parameters (integrator.h:78-88)
	int samplesPerPixel{1};
	int maxDepth{10};
	float probRR{0.8f};
	bool enableNEE{true};
	bool enableMedium{true};
	bool enableClamp{false};
	float clampMax{1e3f};

private:
	void ensureHandle();
	void pushParams();
	KrrWfpt *mHandle = nullptr;
	uint64_t mSceneVersion = 0;
};

// MegakernelPathTracer: src/render/megakernel/pathtracer.{h,cpp}; JSON nee, max_depth, rr
class MegakernelPathTracer : public RenderPass {
public:
	KRR_REGISTER_PASS_DEC(MegakernelPathTracer);
	~MegakernelPathTracer() override;
	void resize(const Vector2i &size) override;
	void setScene(Scene::SharedPtr scene) override;
	void render(RenderContext *context) override;
	string getName() const override { return "MegakernelPathTracer"; }
	void fromJson(const json &j);
	json toJson() const override;
	bool enableNEE{true};
	int maxDepth{10};
	float probRR{0.8f};
	int samplesPerPixel{1};

private:
	void ensureHandle();
	KrrWfpt *mHandle = nullptr;
};

// AccumulatePass: src/render/passes/accumulate/accumulate.{h,cu}
// JSON: spp, mode ("accumulate" | "moving average"), precision ("float" | "double"), save_on_finish,
// exit_on_finish, save_every, task {"type": "spp" | "time", "value": N} (accumulate.h:31-57, util/task.h)
class AccumulatePass : public RenderPass {
public:
	KRR_REGISTER_PASS_DEC(AccumulatePass);
	enum class Mode { Accumulate, MovingAverage, Count };
	enum class Precision { Float, Double, Count };
	enum class BudgetType { None, Spp, Time };
	void resize(const Vector2i &size) override;
	void render(RenderContext *context) override;
	void endFrame(RenderContext *context) override;
	void finalize() override;
	void reset();
	string getName() const override { return "AccumulatePass"; }
	void fromJson(const json &j);
	json toJson() const override;
	~AccumulatePass() override;
	size_t accumCount() const { return mAccumCount; }
	// the accumulated average as an image (AccumulatePass::saveImage, accumulate.cu:90-111)
	bool readAverage(Image &img);
	bool saveImage(const string &path);
	// RenderTask (util/task.h:21-60)
	float progress() const;
	bool finished() const { return progress() >= 1.f; }
	size_t maxAccumCount = 0; // "spp": 0 = unlimited
	Mode mode			 = Mode::Accumulate;
	Precision precision	 = Precision::Float;
	bool saveOnFinish = false, exitOnFinish = false;
	size_t saveEvery	  = 0;
	BudgetType budgetType = BudgetType::None;
	double budgetValue	  = 0;

private:
	void *mAccum = nullptr; // float4 or double4 per pixel
	size_t mAccumCount = 0, mTaskSpp = 0;
	double mTaskStart = 0, mTaskNow = 0;
};

// ErrorMeasurePass: src/render/passes/errormeasure/{errormeasure.cpp,metrics.cu}
// JSON: metric ("mse" | "mape" | "smape" | "rel_mse"), reference (image path; also the global config's
// "reference"), continuous, interval, log, save
class ErrorMeasurePass : public RenderPass {
public:
	KRR_REGISTER_PASS_DEC(ErrorMeasurePass);
	enum class ErrorMetric { MSE, MAPE, SMAPE, RelMSE, Count };
	void beginFrame(RenderContext *context) override;
	void render(RenderContext *context) override;
	void finalize() override;
	string getName() const override { return "ErrorMeasurePass"; }
	void fromJson(const json &j);
	json toJson() const override;
	~ErrorMeasurePass() override;
	bool loadReferenceImage(const string &path);
	void setReferenceImage(const Image &img); // already in film layout (no file, no permutation)
	void evaluateNextFrame() { mNeedsEvaluate = true; } // the UI's "Evaluate" button
	const json &lastResult() const { return mLastResult; }
	double lastValue() const { return mLastValue; }
	struct EvaluationData { size_t timestep; double timepoint; json metrics; };
	const std::vector<EvaluationData> &results() const { return mEvaluationResults; }
	ErrorMetric metric = ErrorMetric::RelMSE;
	bool continuousEvaluate = false, logResults = false, saveResults = false;
	size_t evaluateInterval = 1;

private:
	void reset();
	Image mReferenceImage;
	float *mReferenceDevice = nullptr;
	string mReferenceImagePath;
	json mLastResult;
	double mLastValue = 0, mStartTime = 0;
	bool mNeedsEvaluate = false;
	size_t mFrameNumber = 0;
	std::vector<EvaluationData> mEvaluationResults;
};

// ToneMappingPass: src/render/passes/tonemapping/tonemapping.{h,cu}
// JSON: exposure, operator ("linear" | "reinhard" | "aces" | "uncharted2" | "hejihable"), gamma
class ToneMappingPass : public RenderPass {
public:
	KRR_REGISTER_PASS_DEC(ToneMappingPass);
	enum class Operator { Linear = 0, Reinhard, Aces, Uncharted2, HejiHable, NumsOperators };
	void render(RenderContext *context) override;
	string getName() const override { return "ToneMappingPass"; }
	void fromJson(const json &j);
	json toJson() const override;
	void setOperator(Operator op) { mOperator = op; }
	Operator getOperator() const { return mOperator; }
	bool useGamma = true;
	float exposureCompensation = 1.f;

private:
	Operator mOperator = Operator::Linear;
};

// ------------------------------------------------------------------------------------------------
// Headless RenderApp (renderer.cpp): loads the same JSON config, drives beginFrame/render/endFrame
// ------------------------------------------------------------------------------------------------
class RenderApp {
public:
	void loadConfigFrom(const string &path);
	void loadConfig(const json &config, const string &baseDir);
	void setScene(Scene::SharedPtr scene);
	void resize(Vector2i size);
	void initialize();
	// one iteration of DeviceManager::runMessageLoop (window.cpp:450-485): ++frameIndex; tick; render
	void renderFrame(double timeSeconds = 0);
	void readFilm(std::vector<float> &rgba) { mContext.readback(rgba); }
	RenderContext *context() { return &mContext; }
	Scene::SharedPtr scene() { return mScene; }
	std::vector<RenderPass::SharedPtr> &passes() { return mRenderPasses; }
	template <typename T> std::shared_ptr<T> findPass() {
		for (auto &p : mRenderPasses) if (auto t = std::dynamic_pointer_cast<T>(p)) return t;
		return nullptr;
	}
	size_t frameIndex() const { return mFrameIndex; }
	Vector2i frameSize() const { return mSize; }
	// RenderApp::finalize (renderer.cpp:333-336): finalize() on every pass
	void finalize();
	// gpContext->requestExit() / shouldQuit() (context.h): set by AccumulatePass when its budget is spent
	void requestExit() { mExit = true; }
	bool shouldQuit() const { return mExit; }
	// runs frames until a pass requests the exit (or maxFrames); returns the number of frames rendered
	size_t run(size_t maxFrames = 0);
	const json &globalConfig() const { return mConfig; }
	// File::outputDir(): "output_dir" of the config, else "<base dir>/output"
	string outputDir() const { return mOutputDir; }
	void setOutputDir(const string &d) { mOutputDir = d; }

private:
	std::vector<RenderPass::SharedPtr> mRenderPasses;
	Scene::SharedPtr mScene;
	RenderContext mContext;
	Vector2i mSize{1280, 720};
	size_t mFrameIndex = 0;
	bool mInitialized  = false;
	bool mExit		   = false;
	json mConfig;
	string mOutputDir = "output";
};

// ------------------------------------------------------------------------------------------------
// One process, N GPUs (host/multi_device.cpp): scene replicated, T image tiles x S spp slices, one host thread per
// device, film sum-reduce with the product's NCCL entry point (krr_wfpt_reduce_film).  No reference counterpart.
// ------------------------------------------------------------------------------------------------
class MultiDeviceRenderApp {
public:
	struct Result { double msTotal = 0; uint64_t raysLastStep = 0; int steps = 0; };
	MultiDeviceRenderApp();
	~MultiDeviceRenderApp();
	MultiDeviceRenderApp(const MultiDeviceRenderApp &) = delete;
	// devices[i] = CUDA device of rank i; tiles = T (must divide the number of devices), S = devices / T
	void init(const KrrSceneDesc *scene, const string &paramsJson, int width, int height, const std::vector<int> &devices, int tiles);
	// `steps` frames (frame batches) per spp slice starting at firstFrame; the reduced film of the LAST step lands
	// in filmHost (may be null).  Returns the wall time of the slowest rank and the rays of the last step.
	Result render(const KrrCameraData &cam, uint64_t firstFrame, int steps, float *filmHost);
	int size() const;
	KrrWfpt *handle(int rank);
	bool usesNccl() const;

private:
	struct Impl;
	Impl *m;
};

// colour-space tables (kiraray_b200/data/spectral_srgb.bin); throws when missing
const KrrColorSpaceData &defaultColorSpace();
void setDataDir(const string &dir);

} // namespace krr
