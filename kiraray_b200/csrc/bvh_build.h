// bvh_build.h -- host interface of the device BVH builder (see bvh_build.cu).
#pragma once
#include <cuda_runtime.h>
#include "bvh.cuh"
#include "scene.cuh"

namespace krr {

// transform chains of the moving instances + the ray-time window their TLAS boxes must cover
struct MotionWindow {
	const XformNodeRec *xnodes = nullptr;
	const float *keys		   = nullptr;
	float w0 = 0.f, w1 = 0.f;
};

class BvhBuilder {
public:
	BvhBuilder();
	~BvhBuilder();
	BvhBuilder(const BvhBuilder &) = delete;
	// Builds one BLAS per mesh (object space) and a TLAS over the instances.  dPositions / dIndices /
	// dInstances are device arrays; hMeshes is the host copy of the mesh records (offsets, counts).
	// hMerge (optional, per instance): 1 = static instance with an exactly-identity transform, to be
	// merged into ONE world-space BLAS that enters the TLAS as pseudo-instance `nInstances`; dInstances
	// must then hold nInstances + 1 records, the last one being that pseudo-instance (identity,
	// mesh = nMeshes, blasRoot patched by the caller from mergedRoot()).
	bool build(const float *dPositions, const int32_t *dIndices, const MeshRec *hMeshes, int nMeshes, const InstRec *dInstances,
			   const InstRec *hInstances, int nInstances, const uint8_t *hMerge, int flatMax, const MotionWindow &motion, cudaStream_t stream,
			   char *err);
	int flatCount() const; // BLASes stored as flat triangle lists (<= flatMax triangles; see BvhDev::flats)
	int mergedRoot() const;		// node index of the merged BLAS root, -1 = none
	int mergedTriCount() const;
	// Re-fits the TLAS after instance transforms (or the motion window) changed; topology kept, no
	// host synchronisation.
	bool refitTlas(const InstRec *dInstances, cudaStream_t stream, char *err, const MotionWindow *window = nullptr);
	int refitLaunches() const;
	BvhDev device() const;
	int blasRoot(int mesh) const;
	int triBase(int mesh) const;
	int nodeCount() const;
	int tlasNodeCount() const;
	int triCount() const;

private:
	struct Impl;
	Impl *m;
};

} // namespace krr
