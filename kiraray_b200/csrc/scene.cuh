// scene.cuh -- device-side scene layout (flat arrays, no pointer-chasing object graph).
// Replaces rt::SceneData / MeshData / InstanceData / MaterialData / Light (reference
// src/core/device/scene.h:22-32, src/core/mesh.h:24-58, src/core/texture.h:184-224,
// src/core/light.h:261-293).  Indices instead of pointers: 4-byte handles, and records are sized
// in multiples of 16 bytes so they load as float4.
#pragma once
#include "krr_math.cuh"
#include "motion.cuh"
#include "spectrum.cuh"

namespace krr {

struct MeshRec {
	int32_t posOff;	 // into positions (float3 units)
	int32_t idxOff;	 // into indices (int3 units)
	int32_t nrmOff;	 // into normals, -1 = none
	int32_t uvOff;	 // into texcoords (float2 units), -1 = none
	int32_t tanOff;	 // into tangents, -1 = none
	int32_t nTri;
	int32_t material;  // -1 = null material
	int32_t mediumIn, mediumOut; // -1 = none
	int32_t blasRoot;  // node index of this mesh's BLAS root in the node pool
	int32_t triBase;   // first triangle of this mesh in the BVH triangle pool
	int32_t pad;
};

struct InstRec {
	Xf xf, inv;
	int32_t mesh;
	int32_t lightBase; // first triangle light of this instance, -1 = not emissive
	int32_t motion;	   // moving instance: first node of its transform chain (motion.cuh), -1 = static
	int32_t blasRoot;  // node index of the mesh's BLAS root (copied from MeshRec for one less hop)
};

// texture handle: constant value or an RGBA32F image in the texel pool (bilinear, wrap)
struct TexRec {
	float value[4];
	int32_t valid;
	int32_t texOff; // float4 offset into texel pool, -1 = constant
	int32_t width, height;
};

struct SpectrumRec { // spectral eta / k (KrrSpectrumDesc)
	int32_t kind;
	float a[3], b[3];
	int32_t tabOff, n; // tabulated: offset (floats) into spectrum table pool: n lambdas then n values
};

struct MatRec {
	float diffuse[4], specular[4];
	float specularTransmission, anisotropic, ior;
	int32_t bsdfType, shadingModel;
	TexRec tex[5];
	SpectrumRec eta, k;
	// constant-colour fast path: sigmoid coefficients of diffuse / specular RGB precomputed at upload
	// with the same float operations fromRGB() performs (valid when the corresponding texture is
	// absent or constant); saves 2 x 24 dependent table reads per hit (shading.h:222-223)
	RgbSpectrum diffuseSpec, specularSpec;
	float constDiffuse[3], constSpecular[4];
	int32_t constColours; // 1 = both colours are constants
};

// one emissive triangle = one DiffuseAreaLight (src/core/mesh.cpp:39-59); vertices pre-gathered
struct TriLightRec {
	float p[3][3]; // object space
	float n[3][3];
	int32_t inst;
	float scale;
	float Le[3];	   // the colour L() evaluates (see lights.cuh)
	RgbSpectrum LeSpec; // unbounded coefficients of Le
	int32_t twoSided;
	int32_t hasNormals;
};

struct AnalyticLightRec { // point / directional / spot / infinite (src/core/light.h:30-259)
	int32_t type;
	float color[3];
	float scale;
	float position[3];
	float rotation[9];
	float sceneRadius;
	float cosInner, cosOuter;
	Xf xf, inv;
	RgbSpectrum colorSpec; // unbounded coefficients of color (tint)
	TexRec image;
};

struct LightRec { int32_t type, index; }; // index into triLights or analytic

struct MediumRec {
	int32_t type;
	float sigma_t[3], albedo[3], Le[3];
	float g;
	RgbSpectrum sigmaTSpec, albedoUSpec, albedoBSpec, LeSpec;
	Xf xf, inv;
	float boundsMin[3], boundsMax[3];
	int32_t res[3];
	int32_t densityOff;	 // float offset into density pool
	int32_t majorantOff; // float offset (64^3 majorant grid)
	float scale;
	int32_t albedoOff;	 // float offset of the RGB albedo grid (3 floats per voxel, density lattice), -1 = constant albedo
};

struct SceneDev {
	const float *positions, *normals, *texcoords, *tangents;
	const int32_t *indices;
	const MeshRec *meshes;
	const InstRec *instances;
	const MatRec *materials;
	const LightRec *lights;
	const TriLightRec *triLights;
	const AnalyticLightRec *analytic;
	const int32_t *infiniteLights; // indices into analytic
	const MediumRec *media;
	const float *densityPool;
	const float4 *texels;
	const float *spectrumTables;
	const float *motionKeys;	 // KrrSRT as 10 floats
	const XformNodeRec *xnodes;	 // transform chains of the moving instances
	const float4 *motionFlat;	 // per instance: flat motion record (motion.cuh kMotionFlatStride float4), null = none
	int32_t nMeshes, nInstances, nMaterials, nLights, nInfinite, nMedia;
	float motionStart, motionEnd;
	int32_t hasMotion;
	ColorSpaceDev cs;
};

} // namespace krr
