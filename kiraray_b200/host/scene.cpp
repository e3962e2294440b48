// scene.cpp -- host scene, KRR JSON scene importer, OBJ/MTL reader.
// Follows reference src/scene/krrscene.cpp:8-349 (schema), src/scene/assimp.cpp:60-65, 93-226
// (OBJ material mapping), src/core/camera.cpp:6-61 (camera + orbit controller), src/core/scene.cpp.
#include "krr_host.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace krr {

void proceduralDensity(std::vector<float> &out, const int res[3], uint64_t seed);

namespace {
struct D3 { double x, y, z; };

void identity12(float m[12]) {
	const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
	memcpy(m, I, sizeof I);
}
// c = a * b for 3x4 affine (row-major)
void mul12(const float a[12], const float b[12], float c[12]) {
	float r[12];
	for (int i = 0; i < 3; i++) {
		for (int j = 0; j < 4; j++) {
			double s = 0;
			for (int k = 0; k < 3; k++) s += (double) a[i * 4 + k] * b[k * 4 + j];
			if (j == 3) s += a[i * 4 + 3];
			r[i * 4 + j] = (float) s;
		}
	}
	memcpy(c, r, sizeof r);
}
void quatToMat(const float q[4] /*x,y,z,w*/, double R[9]) {
	double x = q[0], y = q[1], z = q[2], w = q[3];
	double n = std::sqrt(x * x + y * y + z * z + w * w);
	if (n > 0) x /= n, y /= n, z /= n, w /= n;
	R[0] = 1 - 2 * (y * y + z * z), R[1] = 2 * (x * y - z * w), R[2] = 2 * (x * z + y * w);
	R[3] = 2 * (x * y + z * w), R[4] = 1 - 2 * (x * x + z * z), R[5] = 2 * (y * z - x * w);
	R[6] = 2 * (x * z - y * w), R[7] = 2 * (y * z + x * w), R[8] = 1 - 2 * (x * x + y * y);
}
// T * R * S, the SceneGraphNode local transform (src/core/scenenode.cpp)
void srtToMat(const KrrSRT &k, float m[12]) {
	double R[9];
	quatToMat(k.q, R);
	for (int i = 0; i < 3; i++) {
		for (int j = 0; j < 3; j++) m[i * 4 + j] = (float) (R[i * 3 + j] * k.s[j]);
		m[i * 4 + 3] = k.t[i];
	}
}
float luminance3(const float c[3]) { return c[0] * 0.299f + c[1] * 0.587f + c[2] * 0.114f; }
string dirOf(const string &p) {
	size_t s = p.find_last_of("/\\");
	return s == string::npos ? string(".") : p.substr(0, s);
}
bool isAbs(const string &p) { return !p.empty() && p[0] == '/'; }
string joinPath(const string &base, const string &p) { return isAbs(p) ? p : base + "/" + p; }
bool fileExists(const string &p) { std::ifstream f(p); return f.good(); }
} // namespace

Scene::Scene() {
	memset(&camera, 0, sizeof camera);
	// rt::CameraData defaults, src/core/camera.h:21-28
	camera.film_size[0] = 42.666667f, camera.film_size[1] = 24.0f;
	camera.focal_length = 21, camera.focal_distance = 10, camera.lens_radius = 0;
	camera.aspect_ratio = 1.777777f, camera.shutter_open = 0, camera.shutter_time = 0;
	identity12(camera.transform);
	camera.medium = -1;
}

void Scene::setAspectRatio(float aspect) {
	camera.aspect_ratio = aspect;
	camera.film_size[0] = camera.aspect_ratio * camera.film_size[1]; // mPreserveHeight, camera.cpp:11
}

bool Scene::update(size_t frameIndex, double t) {
	bool changed = false;
	updatedInstances.clear();
	if (animated) {
		// keyframe animation of instance SRTs (src/core/animation.cpp: linear interpolation, clamp)
		// SRT keys at time t: linear interpolation, clamped at both ends; rotations by nlerp along the shortest arc
		auto evalKeys = [](const std::vector<float> &times, const std::vector<KrrSRT> &keys, float tt, float m[12]) {
			if (times.size() == 1) { srtToMat(keys[0], m); return; } // a single key: constant
			size_t k = 0;
			while (k + 2 < times.size() && tt >= times[k + 1]) k++;
			float t0 = times[k], t1 = times[k + 1];
			float a	 = t1 > t0 ? std::min(1.f, std::max(0.f, (tt - t0) / (t1 - t0))) : 0.f;
			const KrrSRT &A = keys[k], &B = keys[k + 1];
			KrrSRT s;
			for (int c = 0; c < 3; c++) s.s[c] = (1 - a) * A.s[c] + a * B.s[c], s.t[c] = (1 - a) * A.t[c] + a * B.t[c];
			float d = A.q[0] * B.q[0] + A.q[1] * B.q[1] + A.q[2] * B.q[2] + A.q[3] * B.q[3];
			for (int c = 0; c < 4; c++) s.q[c] = (1 - a) * A.q[c] + a * (d < 0 ? -B.q[c] : B.q[c]);
			srtToMat(s, m);
		};
		for (size_t i = 0; i < instances.size(); i++) {
			HostInstance &in = instances[i];
			const bool own = !in.animTimes.empty() && (in.animTimes.size() >= 2 || !in.animAncestors.empty());
			if (!own && in.animAncestors.empty()) continue;
			float m[12], k[12];
			identity12(m);
			for (const AnimLink &a : in.animAncestors) { // animated ancestors, root first
				mul12(m, a.pre, m);
				if (!a.times.empty()) { evalKeys(a.times, a.keys, (float) t, k); mul12(m, k, m); }
			}
			mul12(m, in.animParent, m);
			if (own) { evalKeys(in.animTimes, in.animKeys, (float) t, k); mul12(m, k, m); }
			if (memcmp(m, in.transform, sizeof m)) {
				memcpy(in.transform, m, sizeof m);
				updatedInstances.push_back((int32_t) i);
				changed = true;
			}
		}
	}
	if (hasCameraController) {
		// OrbitCameraController::update, camera.cpp:46-61:
		//   rotate = AngleAxis(yaw,Y) * AngleAxis(roll=0,Z) * AngleAxis(pitch,X); forward = rotate * -Z
		//   pos = target - forward * radius
		double cy = std::cos((double) cameraController.yaw), sy = std::sin((double) cameraController.yaw);
		double cp = std::cos((double) cameraController.pitch), sp = std::sin((double) cameraController.pitch);
		double R[9] = {cy, sy * sp, sy * cp, 0, cp, -sp, -sy, cy * sp, cy * cp}; // Ry * Rx
		double f[3] = {-R[2], -R[5], -R[8]};
		float m[12];
		for (int i = 0; i < 3; i++) {
			for (int j = 0; j < 3; j++) m[i * 4 + j] = (float) R[i * 3 + j];
			m[i * 4 + 3] = (float) (cameraController.target[i] - f[i] * cameraController.radius);
		}
		if (memcmp(m, camera.transform, sizeof m)) changed = true;
		memcpy(camera.transform, m, sizeof m);
	}
	// Camera::update (camera.cpp:19-27): the camera is inside a medium whose node bbox contains it
	camera.medium = -1;
	for (size_t i = 0; i < media.size(); i++) {
		if (!media[i].hasBound) continue;
		bool in = true;
		for (int k = 0; k < 3; k++) {
			float p = camera.transform[k * 4 + 3];
			in &= p >= media[i].boundMin[k] && p <= media[i].boundMax[k];
		}
		if (in) camera.medium = (int32_t) i;
	}
	return changed;
}

void Scene::boundingBox(float lo[3], float hi[3]) const {
	for (int k = 0; k < 3; k++) lo[k] = 1e30f, hi[k] = -1e30f;
	for (const HostInstance &in : instances) {
		const HostMesh &m = meshes[in.mesh];
		float mlo[3] = {1e30f, 1e30f, 1e30f}, mhi[3] = {-1e30f, -1e30f, -1e30f};
		for (size_t v = 0; v < m.positions.size() / 3; v++)
			for (int k = 0; k < 3; k++) mlo[k] = std::min(mlo[k], m.positions[3 * v + k]), mhi[k] = std::max(mhi[k], m.positions[3 * v + k]);
		// transformed local AABB (8 corners), like the scene graph's global bounding boxes
		for (int c = 0; c < 8; c++) {
			float p[3] = {c & 1 ? mhi[0] : mlo[0], c & 2 ? mhi[1] : mlo[1], c & 4 ? mhi[2] : mlo[2]};
			for (int k = 0; k < 3; k++) {
				float w = in.transform[k * 4] * p[0] + in.transform[k * 4 + 1] * p[1] + in.transform[k * 4 + 2] * p[2] + in.transform[k * 4 + 3];
				lo[k] = std::min(lo[k], w), hi[k] = std::max(hi[k], w);
			}
		}
	}
}

const KrrSceneDesc &Scene::desc() {
	mMeshDescs.resize(meshes.size());
	for (size_t i = 0; i < meshes.size(); i++) {
		HostMesh &m	   = meshes[i];
		KrrMeshDesc &d = mMeshDescs[i];
		d.positions	   = m.positions.data();
		d.normals	   = m.normals.empty() ? nullptr : m.normals.data();
		d.texcoords	   = m.texcoords.empty() ? nullptr : m.texcoords.data();
		d.tangents	   = m.tangents.empty() ? nullptr : m.tangents.data();
		d.indices	   = m.indices.data();
		d.n_vertices   = (int32_t) (m.positions.size() / 3);
		d.n_triangles  = (int32_t) (m.indices.size() / 3);
		d.material = m.material, d.medium_inside = m.mediumInside, d.medium_outside = m.mediumOutside;
		memcpy(d.Le, m.Le, 12);
	}
	mInstanceDescs.resize(instances.size());
	for (size_t i = 0; i < instances.size(); i++) {
		KrrInstanceDesc &d = mInstanceDescs[i];
		d.mesh = instances[i].mesh;
		memcpy(d.transform, instances[i].transform, 48);
		d.n_motion_keys = (int32_t) instances[i].motionKeys.size();
		d.motion_keys	= instances[i].motionKeys.empty() ? nullptr : instances[i].motionKeys.data();
		d.transform_node = -1;
	}
	mMaterialDescs.resize(materials.size());
	for (size_t i = 0; i < materials.size(); i++) {
		HostMaterial &m = materials[i];
		if (m.desc.spectral_eta.kind == KRR_SPEC_TABULATED) {
			m.desc.spectral_eta.lambdas = m.etaLambdas.data(), m.desc.spectral_eta.values = m.etaValues.data();
			m.desc.spectral_eta.n = (int32_t) m.etaLambdas.size();
		}
		if (m.desc.spectral_k.kind == KRR_SPEC_TABULATED) {
			m.desc.spectral_k.lambdas = m.kLambdas.data(), m.desc.spectral_k.values = m.kValues.data();
			m.desc.spectral_k.n = (int32_t) m.kLambdas.size();
		}
		for (size_t t = 0; t < m.images.size() && t < KRR_TEX_COUNT; t++)
			if (!m.images[t].empty()) m.desc.textures[t].image = m.images[t].data();
		mMaterialDescs[i] = m.desc;
	}
	mMediumDescs.resize(media.size());
	for (size_t i = 0; i < media.size(); i++) {
		if (!media[i].density.empty()) media[i].desc.density = media[i].density.data();
		if (!media[i].albedoGrid.empty()) media[i].desc.albedo_grid = media[i].albedoGrid.data();
		mMediumDescs[i] = media[i].desc;
	}
	// sceneRadius of directional/infinite lights: root bounding-box diagonal (light.cpp:20-21,32-33)
	float lo[3], hi[3];
	boundingBox(lo, hi);
	float diag = instances.empty() ? 0.f : std::sqrt((hi[0] - lo[0]) * (hi[0] - lo[0]) + (hi[1] - lo[1]) * (hi[1] - lo[1]) + (hi[2] - lo[2]) * (hi[2] - lo[2]));
	for (auto &l : lights) l.scene_radius = diag;
	mDesc.meshes = mMeshDescs.data(), mDesc.n_meshes = (int32_t) mMeshDescs.size();
	mDesc.instances = mInstanceDescs.data(), mDesc.n_instances = (int32_t) mInstanceDescs.size();
	mDesc.materials = mMaterialDescs.data(), mDesc.n_materials = (int32_t) mMaterialDescs.size();
	mDesc.lights = lights.data(), mDesc.n_lights = (int32_t) lights.size();
	mDesc.media = mMediumDescs.data(), mDesc.n_media = (int32_t) mMediumDescs.size();
	mDesc.options = options;
	return mDesc;
}

// ------------------------------------------------------------------------------------------------
// OBJ / MTL
// ------------------------------------------------------------------------------------------------
namespace {

// convertSpecPowerToRoughness, assimp.cpp:60-65
float convertSpecPowerToRoughness(float specPower) {
	if (specPower >= 1000) return 0;
	return std::min(1.f, std::max(0.f, std::sqrt(2.0f / (specPower + 2.0f))));
}

struct MtlEntry {
	string name;
	float Kd[3] = {0.6f, 0.6f, 0.6f}, Ks[3] = {0, 0, 0}, Ke[3] = {0, 0, 0}, Tf[3] = {1, 1, 1};
	float Ns = 0, Ni = 1, d = 1;
	bool hasNs = false;
};

// createMaterial for ImportMode::OBJ, assimp.cpp:93-226
HostMaterial materialFromMtl(const MtlEntry &e) {
	HostMaterial m;
	m.name = e.name;
	KrrMaterialDesc &d = m.desc;
	memset(&d, 0, sizeof d);
	d.diffuse[0] = d.diffuse[1] = d.diffuse[2] = d.diffuse[3] = 1; // MaterialParams defaults, texture.h:129-137
	d.ior = 1.5f;
	d.bsdf_type		= KRR_MAT_DISNEY;					// Material::mBsdfType default, texture.h:165
	d.shading_model = KRR_SHADING_SPECULAR_GLOSSINESS;	// OBJ import, assimp.cpp:222-224
	d.color_space	= 0;
	d.diffuse[3] = e.d; // opacity
	if (e.d == 0.f) { d.specular_transmission = 1 - e.d; d.bsdf_type = KRR_MAT_DIELECTRIC; }
	// shininess -> glossiness (assimp always exports Ns for OBJ materials)
	d.specular[3] = 1.f - convertSpecPowerToRoughness(e.Ns);
	d.ior = e.Ni;
	// AI_MATKEY_COLOR_TRANSPARENT (Tf, default 1): transmission = 1 - Tf
	float tr[3] = {1 - e.Tf[0], 1 - e.Tf[1], 1 - e.Tf[2]};
	d.specular_transmission = luminance3(tr);
	if (luminance3(tr) > 1 - 1e-5f) d.bsdf_type = KRR_MAT_DIELECTRIC;
	for (int k = 0; k < 3; k++) d.diffuse[k] = e.Kd[k], d.specular[k] = e.Ks[k];
	if (e.Ke[0] != 0 || e.Ke[1] != 0 || e.Ke[2] != 0) {
		// setConstantTexture(Emissive, RGBA(emissive, 1)), assimp.cpp:178-181
		KrrTextureDesc &t = d.textures[KRR_TEX_EMISSIVE];
		t.valid = 1;
		t.value[0] = e.Ke[0], t.value[1] = e.Ke[1], t.value[2] = e.Ke[2], t.value[3] = 1;
	}
	return m;
}

std::vector<MtlEntry> parseMtl(const string &path) {
	std::vector<MtlEntry> out;
	std::ifstream f(path);
	string line;
	while (std::getline(f, line)) {
		size_t h = line.find('#');
		if (h != string::npos) line = line.substr(0, h);
		std::istringstream ss(line);
		string tok;
		if (!(ss >> tok)) continue;
		if (tok == "newmtl") { MtlEntry e; ss >> e.name; out.push_back(e); continue; }
		if (out.empty()) continue;
		MtlEntry &e = out.back();
		if (tok == "Kd") ss >> e.Kd[0] >> e.Kd[1] >> e.Kd[2];
		else if (tok == "Ks") ss >> e.Ks[0] >> e.Ks[1] >> e.Ks[2];
		else if (tok == "Ke") ss >> e.Ke[0] >> e.Ke[1] >> e.Ke[2];
		else if (tok == "Tf") ss >> e.Tf[0] >> e.Tf[1] >> e.Tf[2];
		else if (tok == "Ns") { ss >> e.Ns; e.hasNs = true; }
		else if (tok == "Ni") ss >> e.Ni;
		else if (tok == "d") ss >> e.d;
		else if (tok == "Tr") { float tr; ss >> tr; e.d = 1 - tr; }
	}
	return out;
}

struct ObjGroup {
	string name;
	int material = -1;
	std::vector<int> faceStart; // into corner arrays
	std::vector<int> vi, ti, ni; // per corner indices (0-based, -1 = none)
};

} // namespace

bool loadObj(const string &filepath, Scene &scene, const float nodeTransform[12]) {
	std::ifstream f(filepath);
	if (!f.good()) return false;
	std::vector<float> V, VT, VN;
	std::vector<ObjGroup> groups;
	std::map<string, int> mtlIndex;
	string curGroup = "default";
	int curMtl		= -1;
	bool needNew	= true;
	string line;
	const string dir = dirOf(filepath);
	while (std::getline(f, line)) {
		size_t h = line.find('#');
		if (h != string::npos) line = line.substr(0, h);
		std::istringstream ss(line);
		string tok;
		if (!(ss >> tok)) continue;
		if (tok == "v") { float x, y, z; ss >> x >> y >> z; V.insert(V.end(), {x, y, z}); }
		else if (tok == "vt") { float u = 0, v = 0; ss >> u >> v; VT.insert(VT.end(), {u, 1.f - v}); /* aiProcess_FlipUVs */ }
		else if (tok == "vn") { float x, y, z; ss >> x >> y >> z; VN.insert(VN.end(), {x, y, z}); }
		else if (tok == "mtllib") {
			string name; ss >> name;
			for (const MtlEntry &e : parseMtl(joinPath(dir, name))) {
				mtlIndex[e.name] = (int) scene.materials.size();
				scene.materials.push_back(materialFromMtl(e));
			}
		} else if (tok == "g" || tok == "o") { ss >> curGroup; needNew = true; }
		else if (tok == "usemtl") {
			string name; ss >> name;
			auto it = mtlIndex.find(name);
			curMtl	= it == mtlIndex.end() ? -1 : it->second;
			needNew = true;
		} else if (tok == "f") {
			if (needNew) { groups.push_back(ObjGroup{curGroup, curMtl, {}, {}, {}, {}}); needNew = false; }
			ObjGroup &g = groups.back();
			std::vector<int> fv, ft, fn;
			string c;
			while (ss >> c) {
				int vi = 0, ti = 0, ni = 0;
				// v, v/t, v//n, v/t/n
				size_t s1 = c.find('/');
				vi = std::atoi(c.substr(0, s1).c_str());
				if (s1 != string::npos) {
					size_t s2 = c.find('/', s1 + 1);
					string t = c.substr(s1 + 1, s2 == string::npos ? string::npos : s2 - s1 - 1);
					if (!t.empty()) ti = std::atoi(t.c_str());
					if (s2 != string::npos) ni = std::atoi(c.substr(s2 + 1).c_str());
				}
				auto fix = [](int i, size_t n) { return i > 0 ? i - 1 : i < 0 ? (int) n + i : -1; };
				fv.push_back(fix(vi, V.size() / 3));
				ft.push_back(fix(ti, VT.size() / 2));
				fn.push_back(fix(ni, VN.size() / 3));
			}
			// aiProcess_Triangulate: fan
			for (size_t k = 1; k + 1 < fv.size(); k++) {
				g.faceStart.push_back((int) g.vi.size());
				for (size_t c3 : {(size_t) 0, k, k + 1}) { g.vi.push_back(fv[c3]); g.ti.push_back(ft[c3]); g.ni.push_back(fn[c3]); }
			}
		}
	}
	for (ObjGroup &g : groups) {
		if (g.vi.empty()) continue;
		HostMesh mesh;
		mesh.name	  = g.name;
		mesh.material = g.material;
		const size_t nc = g.vi.size(), nf = nc / 3;
		bool hasN = true, hasT = true;
		for (size_t c = 0; c < nc; c++) hasN &= g.ni[c] >= 0, hasT &= g.ti[c] >= 0;
		// face normals (normalised), used for aiProcess_GenSmoothNormals (max angle 60 deg)
		std::vector<D3> fnrm(nf);
		for (size_t fi = 0; fi < nf; fi++) {
			const float *a = &V[3 * g.vi[3 * fi]], *b = &V[3 * g.vi[3 * fi + 1]], *c = &V[3 * g.vi[3 * fi + 2]];
			double e1[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, e2[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
			D3 n{e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
			double l = std::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
			if (l > 0) n.x /= l, n.y /= l, n.z /= l;
			fnrm[fi] = n;
		}
		// faces incident on each referenced position
		std::map<int, std::vector<int>> incident;
		if (!hasN) for (size_t c = 0; c < nc; c++) incident[g.vi[c]].push_back((int) (c / 3));
		const double cosLimit = std::cos(60.0 * M_PI / 180.0);
		// aiProcess_JoinIdenticalVertices: merge corners with identical attributes
		struct Key { float p[3], n[3], t[2]; bool operator<(const Key &o) const { return memcmp(this, &o, sizeof(Key)) < 0; } };
		std::map<Key, int> remap;
		for (size_t c = 0; c < nc; c++) {
			Key k;
			memset(&k, 0, sizeof k);
			memcpy(k.p, &V[3 * g.vi[c]], 12);
			if (hasN) memcpy(k.n, &VN[3 * g.ni[c]], 12);
			else {
				const D3 &fn = fnrm[c / 3];
				D3 acc{0, 0, 0};
				for (int of : incident[g.vi[c]]) {
					const D3 &on = fnrm[of];
					if (fn.x * on.x + fn.y * on.y + fn.z * on.z >= cosLimit) acc.x += on.x, acc.y += on.y, acc.z += on.z;
				}
				double l = std::sqrt(acc.x * acc.x + acc.y * acc.y + acc.z * acc.z);
				if (l > 0) acc.x /= l, acc.y /= l, acc.z /= l;
				k.n[0] = (float) acc.x, k.n[1] = (float) acc.y, k.n[2] = (float) acc.z;
			}
			if (hasT) memcpy(k.t, &VT[2 * g.ti[c]], 8);
			auto it = remap.find(k);
			int id;
			if (it == remap.end()) {
				id = (int) (mesh.positions.size() / 3);
				remap[k] = id;
				mesh.positions.insert(mesh.positions.end(), k.p, k.p + 3);
				mesh.normals.insert(mesh.normals.end(), k.n, k.n + 3);
				if (hasT) mesh.texcoords.insert(mesh.texcoords.end(), k.t, k.t + 2);
			} else id = it->second;
			mesh.indices.push_back(id);
		}
		if (hasT) {
			// aiProcess_CalcTangentSpace (needs UVs): per-vertex accumulated face tangents
			mesh.tangents.assign(mesh.positions.size(), 0.f);
			for (size_t fi = 0; fi < nf; fi++) {
				int i0 = mesh.indices[3 * fi], i1 = mesh.indices[3 * fi + 1], i2 = mesh.indices[3 * fi + 2];
				const float *p0 = &mesh.positions[3 * i0], *p1 = &mesh.positions[3 * i1], *p2 = &mesh.positions[3 * i2];
				const float *t0 = &mesh.texcoords[2 * i0], *t1 = &mesh.texcoords[2 * i1], *t2 = &mesh.texcoords[2 * i2];
				float du1 = t1[0] - t0[0], dv1 = t1[1] - t0[1], du2 = t2[0] - t0[0], dv2 = t2[1] - t0[1];
				float det = du1 * dv2 - du2 * dv1;
				float r	  = det != 0 ? 1.f / det : 0.f;
				for (int k = 0; k < 3; k++) {
					float tk = ((p1[k] - p0[k]) * dv2 - (p2[k] - p0[k]) * dv1) * r;
					mesh.tangents[3 * i0 + k] += tk, mesh.tangents[3 * i1 + k] += tk, mesh.tangents[3 * i2 + k] += tk;
				}
			}
			for (size_t v = 0; v < mesh.positions.size() / 3; v++) {
				float *t = &mesh.tangents[3 * v];
				float l	 = std::sqrt(t[0] * t[0] + t[1] * t[1] + t[2] * t[2]);
				if (l > 0) t[0] /= l, t[1] /= l, t[2] /= l;
				else t[0] = 1, t[1] = 0, t[2] = 0;
			}
		}
		HostInstance inst;
		inst.mesh = (int) scene.meshes.size();
		memcpy(inst.transform, nodeTransform, 48);
		scene.meshes.push_back(std::move(mesh));
		scene.instances.push_back(inst);
	}
	scene.touch();
	return true;
}

// ------------------------------------------------------------------------------------------------
// KRR JSON scene
// ------------------------------------------------------------------------------------------------
namespace {

void nodeTransformFromJson(const json &j, const float parent[12], float out[12]) {
	KrrSRT s{{1, 1, 1}, {0, 0, 0, 1}, {0, 0, 0}};
	j.getFloats<3>("translate", s.t);
	j.getFloats<3>("scale", s.s);
	float q[4];
	if (j.getFloats<4>("rotate", q)) { s.q[3] = q[0], s.q[0] = q[1], s.q[1] = q[2], s.q[2] = q[3]; } // [w,x,y,z]
	float local[12];
	srtToMat(s, local);
	mul12(parent, local, out);
}

// SceneLight::setDirection (scenegraph.cpp:489-503): node rotation = inverse(look_at(0, dir, +Y))
void rotationFromDirection(const float dir[3], double R[9]) {
	double f[3] = {dir[0], dir[1], dir[2]};
	double l	= std::sqrt(f[0] * f[0] + f[1] * f[1] + f[2] * f[2]);
	for (double &c : f) c /= l;
	double up[3] = {0, 1, 0};
	double s[3]	 = {f[1] * up[2] - f[2] * up[1], f[2] * up[0] - f[0] * up[2], f[0] * up[1] - f[1] * up[0]};
	l = std::sqrt(s[0] * s[0] + s[1] * s[1] + s[2] * s[2]);
	for (double &c : s) c /= l;
	double u[3] = {s[1] * f[2] - s[2] * f[1], s[2] * f[0] - s[0] * f[2], s[0] * f[1] - s[1] * f[0]};
	// columns (s, u, -f)
	for (int i = 0; i < 3; i++) R[i * 3 + 0] = s[i], R[i * 3 + 1] = u[i], R[i * 3 + 2] = -f[i];
}

// Image::loadImage path rule (texture.cpp:34-38): "$name" is a built-in texture, File::textureDir() / name
// (common/assets/textures of the reference's tree); here: $KRR_TEXTURE_DIR, then assets/textures and
// common/assets/textures under the asset root
string resolveTexturePath(const string &name, const string &baseDir) {
	size_t d = name.find('$');
	if (d == string::npos) return joinPath(baseDir, name);
	string file = name.substr(d + 1);
	std::vector<string> dirs;
	if (const char *e = getenv("KRR_TEXTURE_DIR")) dirs.push_back(e);
	dirs.push_back(joinPath(baseDir, "assets/textures"));
	dirs.push_back(joinPath(baseDir, "common/assets/textures"));
	for (const string &dir : dirs)
		if (fileExists(dir + "/" + file)) return dir + "/" + file;
	return dirs.back() + "/" + file;
}

// Texture::createFromFile for the lat-long image of an infinite light.  The reference logs an error and
// carries on with an invalid texture when the file cannot be read (texture.h:95-99): the light then emits
// its tint; same here, with the message on stderr.
void setLightImage(Scene &scene, KrrLightDesc &l, const string &texture, const string &baseDir) {
	Image img;
	string err, path = resolveTexturePath(texture, baseDir);
	if (!loadImage(path, img, false, &err) || !img.isValid()) {
		fprintf(stderr, "[krr_host] failed to load texture %s: %s\n", path.c_str(), err.c_str());
		return;
	}
	scene.lightImages.push_back(std::move(img.rgba));
	l.texture.valid	 = 1;
	l.texture.value[0] = l.texture.value[1] = l.texture.value[2] = l.texture.value[3] = 1;
	l.texture.image	 = scene.lightImages.back().data();
	l.texture.width = img.width, l.texture.height = img.height;
}

bool loadLight(Scene &scene, const json &params, const float nodeXf[12], const string &baseDir) {
	string type = params.value("type", "infinite");
	KrrLightDesc l;
	memset(&l, 0, sizeof l);
	l.scale = params.value("scale", 1.f);
	l.color[0] = l.color[1] = l.color[2] = 1;
	params.getFloats<3>("color", l.color);
	memcpy(l.transform, nodeXf, 48);
	if (type == "point") l.type = KRR_LIGHT_POINT;
	else if (type == "directional") l.type = KRR_LIGHT_DIRECTIONAL;
	else if (type == "spotlight") {
		l.type = KRR_LIGHT_SPOT;
		l.inner_cone_deg = params.value("inner_cone", 30.f);
		l.outer_cone_deg = params.value("outer_cone", 45.f);
	} else if (type == "infinite") {
		l.type = KRR_LIGHT_INFINITE;
		// krr::InfiniteLight(color, scale) uploads through the texture constructor with tint = 1
		// (light.cpp:35-38): a texture-less JSON light has an invalid image -> Li = tint = (1,1,1)
		string texture = params.value("texture", string());
		if (!texture.empty()) setLightImage(scene, l, texture, baseDir); // krrscene.cpp:44-48
	} else return false;
	float v[3];
	if (params.getFloats<3>("position", v)) { l.transform[3] = v[0], l.transform[7] = v[1], l.transform[11] = v[2]; }
	if (params.getFloats<3>("direction", v)) {
		double R[9];
		rotationFromDirection(v, R);
		for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) l.transform[i * 4 + j] = (float) R[i * 3 + j];
	}
	scene.lights.push_back(l);
	return true;
}

void addBoxMesh(Scene &scene, const float lo[3], const float hi[3], int mediumInside, const float xf[12], int mediumOutside = -1) {
	// krrscene.cpp:95-113
	HostMesh mesh;
	mesh.name = "Medium box";
	const int idx[12][3] = {{4, 2, 0}, {2, 7, 3}, {6, 5, 7}, {1, 7, 5}, {0, 3, 1}, {4, 1, 5},
							{4, 6, 2}, {2, 6, 7}, {6, 4, 5}, {1, 3, 7}, {0, 2, 3}, {4, 0, 1}};
	const float P[8][3] = {{hi[0], hi[1], lo[2]}, {hi[0], lo[1], lo[2]}, {hi[0], hi[1], hi[2]}, {hi[0], lo[1], hi[2]},
						   {lo[0], hi[1], lo[2]}, {lo[0], lo[1], lo[2]}, {lo[0], hi[1], hi[2]}, {lo[0], lo[1], hi[2]}};
	for (auto &p : P) mesh.positions.insert(mesh.positions.end(), p, p + 3);
	for (auto &t : idx) mesh.indices.insert(mesh.indices.end(), t, t + 3);
	mesh.material	  = -1;
	mesh.mediumInside = mediumInside;
	mesh.mediumOutside = mediumOutside; // a medium box nested in another medium ("outside": index of the enclosing medium)
	HostInstance inst;
	inst.mesh = (int) scene.meshes.size();
	memcpy(inst.transform, xf, 48);
	scene.meshes.push_back(std::move(mesh));
	scene.instances.push_back(inst);
}

bool loadMedium(Scene &scene, const json &params, const float nodeXf[12]) {
	string type = params.value("type", "homogeneous");
	HostMedium m;
	memset(&m.desc, 0, sizeof m.desc);
	identity12(m.desc.transform);
	if (type == "homogeneous") {
		m.desc.type = KRR_MEDIUM_HOMOGENEOUS;
		float sigma_t[3] = {1, 1, 1}, albedo[3] = {0.5f, 0.5f, 0.5f}, Le[3] = {0, 0, 0};
		params.getFloats<3>("sigma_t", sigma_t);
		params.getFloats<3>("albedo", albedo);
		params.getFloats<3>("Le", Le);
		if (params.contains("sigma_a") || params.contains("sigma_s")) {
			float sa[3] = {0.5f, 0.5f, 0.5f}, ss[3] = {0.5f, 0.5f, 0.5f};
			params.getFloats<3>("sigma_a", sa);
			params.getFloats<3>("sigma_s", ss);
			for (int k = 0; k < 3; k++) sigma_t[k] = sa[k] + ss[k], albedo[k] = ss[k] / sigma_t[k];
		}
		memcpy(m.desc.sigma_t, sigma_t, 12), memcpy(m.desc.albedo, albedo, 12), memcpy(m.desc.Le, Le, 12);
		m.desc.g = params.value("g", 0.f);
		int id	 = (int) scene.media.size();
		if (params.contains("bound")) {
			const json &b = params.at("bound"); // AABB3f: [[min],[max]]
			float lo[3], hi[3];
			for (int k = 0; k < 3; k++) lo[k] = (float) b.at(0).at(k).asNumber(), hi[k] = (float) b.at(1).at(k).asNumber();
			m.hasBound = true;
			for (int k = 0; k < 3; k++) m.boundMin[k] = lo[k] + nodeXf[k * 4 + 3], m.boundMax[k] = hi[k] + nodeXf[k * 4 + 3];
			scene.media.push_back(m);
			addBoxMesh(scene, lo, hi, id, nodeXf, params.value("outside", -1));
		} else {
			scene.media.push_back(m);
			if (params.contains("meshes"))
				for (const json &n : params.at("meshes").items())
					for (HostMesh &mesh : scene.meshes) if (mesh.name == n.asString()) mesh.mediumInside = id;
			if (params.contains("meshes_outside"))
				for (const json &n : params.at("meshes_outside").items())
					for (HostMesh &mesh : scene.meshes) if (mesh.name == n.asString()) mesh.mediumOutside = id;
		}
		return true;
	}
	if (type == "heterogeneous" || type == "grid") {
		// The reference loads OpenVDB files here (out of scope); a procedural dense grid stands in:
		// {"type":"grid","res":[x,y,z],"bound":[[..],[..]],"procedural":"blobs","seed":7272, ...}
		m.desc.type = KRR_MEDIUM_GRID;
		float sigma_t[3] = {1, 1, 1}, albedo[3] = {0.5f, 0.5f, 0.5f};
		params.getFloats<3>("sigma_t", sigma_t);
		params.getFloats<3>("albedo", albedo);
		memcpy(m.desc.sigma_t, sigma_t, 12), memcpy(m.desc.albedo, albedo, 12);
		m.desc.g	 = params.value("g", 0.f);
		m.desc.scale = params.value("scale", 1.f);
		int res[3]	 = {64, 64, 64};
		float rf[3];
		if (params.getFloats<3>("res", rf)) for (int k = 0; k < 3; k++) res[k] = (int) rf[k];
		float lo[3] = {-1, -1, -1}, hi[3] = {1, 1, 1};
		if (params.contains("bound")) {
			const json &b = params.at("bound");
			for (int k = 0; k < 3; k++) lo[k] = (float) b.at(0).at(k).asNumber(), hi[k] = (float) b.at(1).at(k).asNumber();
		}
		memcpy(m.desc.bounds_min, lo, 12), memcpy(m.desc.bounds_max, hi, 12);
		memcpy(m.desc.res, res, 12);
		memcpy(m.desc.transform, nodeXf, 48);
		proceduralDensity(m.density, res, (uint64_t) params.value("seed", 7272));
		if (params.contains("albedo_gradient")) {
			// RGB albedo grid (NanoVDBMedium::albedoGrid; the reference reads it from the .vdb file): a linear blend of two
			// colours along x, {"albedo_gradient": [[r,g,b],[r,g,b]]}, on the density lattice
			const json &g = params.at("albedo_gradient");
			float a[3], b[3];
			for (int k = 0; k < 3; k++) a[k] = (float) g.at(0).at(k).asNumber(), b[k] = (float) g.at(1).at(k).asNumber();
			m.albedoGrid.resize((size_t) 3 * res[0] * res[1] * res[2]);
			for (int z = 0; z < res[2]; z++)
				for (int y = 0; y < res[1]; y++)
					for (int x = 0; x < res[0]; x++) {
						const float t = res[0] > 1 ? (float) x / (float) (res[0] - 1) : 0.f;
						float *v = &m.albedoGrid[3 * ((size_t) x + (size_t) res[0] * ((size_t) y + (size_t) res[1] * z))];
						for (int k = 0; k < 3; k++) v[k] = a[k] + t * (b[k] - a[k]);
					}
		}
		m.hasBound = true;
		for (int k = 0; k < 3; k++) m.boundMin[k] = lo[k] + nodeXf[k * 4 + 3], m.boundMax[k] = hi[k] + nodeXf[k * 4 + 3];
		int id = (int) scene.media.size();
		scene.media.push_back(std::move(m));
		addBoxMesh(scene, lo, hi, id, nodeXf, params.value("outside", -1));
		return true;
	}
	return false;
}

int materialTypeFromString(const string &s) {
	if (s == "null") return KRR_MAT_NULL;
	if (s == "diffuse") return KRR_MAT_DIFFUSE;
	if (s == "conductor") return KRR_MAT_CONDUCTOR;
	if (s == "dielectric") return KRR_MAT_DIELECTRIC;
	if (s == "disney") return KRR_MAT_DISNEY;
	return KRR_MAT_DIFFUSE;
}

KrrSpectrumDesc extractSpectrum(const json &j) {
	KrrSpectrumDesc s;
	memset(&s, 0, sizeof s);
	if (j.isObject()) {
		string type = j.value("type", "null");
		if (type == "constant_spectrum") { s.kind = KRR_SPEC_CONSTANT; s.a[0] = j.value("value", 1.5f); }
		else if (type == "cauchy_spectrum") { s.kind = KRR_SPEC_CAUCHY; s.a[0] = j.value("a", 1.5f); s.b[0] = j.value("b", 0.f); }
		else if (type == "sellmeier_spectrum") { s.kind = KRR_SPEC_SELLMEIER; j.getFloats<3>("b", s.a); j.getFloats<3>("c", s.b); }
	}
	return s; // named spectra (metal-Au-eta, ...) need the reference's measured tables: unsupported
}

float spectrumMax(const KrrSpectrumDesc &s) {
	switch (s.kind) {
		case KRR_SPEC_CONSTANT: return s.a[0];
		case KRR_SPEC_CAUCHY: return s.a[0] + (s.a[0] + s.b[0] / (0.36f * 0.36f)); // CauchyIoRSpectrum::maxValue, spectrum.h:121
		case KRR_SPEC_SELLMEIER: return 1.f;
		default: return 1.5f;
	}
}

// loadMaterials, krrscene.cpp:158-231
void loadMaterials(Scene &scene, const json &j) {
	for (const json &m : j.items()) {
		string name = m.value("name", "Untitled");
		HostMaterial *mat = nullptr;
		for (HostMaterial &hm : scene.materials) if (hm.name == name) mat = &hm;
		if (!mat) {
			scene.materials.emplace_back();
			mat = &scene.materials.back();
			memset(&mat->desc, 0, sizeof mat->desc);
			mat->name = name;
			mat->desc.diffuse[3] = 1;
			mat->desc.ior = 1.5f;
		}
		json params = m.value("params", json::object());
		KrrMaterialDesc &d = mat->desc;
		float v[3] = {1, 1, 1};
		params.getFloats<3>("diffuse", v);
		memcpy(d.diffuse, v, 12);
		float s[3] = {0, 0, 0};
		params.getFloats<3>("specular", s);
		memcpy(d.specular, s, 12);
		d.specular[3]			= 1 - params.value("roughness", 1.f);
		d.specular_transmission = params.value("specular_transmission", 0.f);
		d.anisotropic			= params.value("anisotropic", 0.f);
		if (params.contains("eta")) {
			const json &e = params.at("eta");
			if (e.isFloat()) d.ior = (float) e.asNumber();
			else { d.spectral_eta = extractSpectrum(e); d.ior = spectrumMax(d.spectral_eta); }
		}
		if (params.contains("k")) d.spectral_k = extractSpectrum(params.at("k"));
		d.bsdf_type		= materialTypeFromString(m.value("bsdf", "diffuse"));
		d.shading_model = KRR_SHADING_SPECULAR_GLOSSINESS;
		d.color_space	= 0;
	}
}

bool importNode(const json &j, Scene::SharedPtr scene, const float parent[12], const string &baseDir) {
	if (j.isArray()) {
		for (const json &m : j.items()) importNode(m, scene, parent, baseDir);
	} else if (j.isObject()) {
		float xf[12];
		nodeTransformFromJson(j, parent, xf);
		json params = j.value("params", json::object());
		if (j.contains("model")) importNode(j.at("model"), scene, xf, baseDir);
		else {
			string type = j.value("type", "model");
			if (type == "medium") return loadMedium(*scene, params, xf);
			if (type == "light") return loadLight(*scene, params, xf, baseDir);
			return false;
		}
	} else if (j.isString()) {
		return SceneImporter::loadModel(j.asString(), scene, parent, baseDir);
	} else return false;
	return true;
}

} // namespace

bool SceneImporter::loadModel(const string &filepath, Scene::SharedPtr scene, const float xf[12], const string &baseDir) {
	string path = joinPath(baseDir, filepath);
	if (!fileExists(path)) {
		// the reference's configs name assets relative to its source tree ("common/assets/..."):
		// fall back to this repository's assets/ directory by file name
		size_t s = filepath.find("scenes/");
		if (s != string::npos) path = joinPath(baseDir, "assets/" + filepath.substr(s + 7));
	}
	size_t dot = path.find_last_of('.');
	string ext = dot == string::npos ? "" : path.substr(dot);
	// the reference logs "Failed to load" and carries on with an empty scene, which then dies in the
	// OptiX build; here a missing asset is an error the C entry point returns
	if (!fileExists(path)) throw std::runtime_error("cannot open model " + path);
	if (ext == ".obj") return loadObj(path, *scene, xf);
	if (ext == ".gltf" || ext == ".glb") return loadGltf(path, *scene, xf);
	if (ext == ".json") {
		std::ifstream f(path);
		std::stringstream ss;
		ss << f.rdbuf();
		return SceneImporter::import(json::parse(ss.str()), scene, dirOf(path));
	}
	return false; // FBX / pbrt / VDB importers are out of scope
}

bool SceneImporter::addEnvironment(const string &texture, Scene::SharedPtr scene, const string &baseDir) {
	KrrLightDesc l;
	memset(&l, 0, sizeof l);
	l.type	= KRR_LIGHT_INFINITE;
	l.scale = 1;
	l.color[0] = l.color[1] = l.color[2] = 1;
	identity12(l.transform);
	setLightImage(*scene, l, texture, baseDir);
	scene->lights.push_back(l);
	scene->touch();
	return l.texture.image != nullptr;
}

bool SceneImporter::import(const json &j, Scene::SharedPtr scene, const string &baseDir) {
	float I[12];
	identity12(I);
	if (j.contains("camera")) {
		json c	 = j.at("camera").value("mData", json::object());
		auto &cd = scene->camera;
		cd.focal_length	  = c.value("focalLength", cd.focal_length);
		cd.focal_distance = c.value("focalDistance", cd.focal_distance);
		cd.lens_radius	  = c.value("lensRadius", cd.lens_radius);
		cd.aspect_ratio	  = c.value("aspectRatio", cd.aspect_ratio);
		cd.shutter_open	  = c.value("shutterOpen", cd.shutter_open);
		cd.shutter_time	  = c.value("shutterTime", cd.shutter_time);
		if (j.contains("cameraController")) {
			json cc	 = j.at("cameraController").value("mData", json::object());
			auto &cr = scene->cameraController;
			cr.radius = cc.value("radius", cr.radius);
			cr.pitch  = cc.value("pitch", cr.pitch);
			cr.yaw	  = cc.value("yaw", cr.yaw);
			cc.getFloats<3>("target", cr.target);
		}
		scene->hasCameraController = true;
	}
	if (j.contains("model")) importNode(j.at("model"), scene, I, baseDir);
	if (j.contains("environment")) addEnvironment(j.at("environment").asString(), scene, baseDir);
	if (j.contains("media"))
		for (const json &m : j.at("media").items()) loadMedium(*scene, m.at("params"), I);
	if (j.contains("materials")) loadMaterials(*scene, j.at("materials"));
	if (j.contains("options")) {
		const json &o = j.at("options");
		scene->animated			  = o.value("animated", false);
		scene->options.animated	  = o.value("animated", true);
		scene->options.multilevel = o.value("multilevel", false);
		scene->options.motionblur = o.value("motionblur", false);
		scene->options.starttime  = o.value("starttime", 0.f);
		scene->options.endtime	  = o.value("endtime", 1.f);
	}
	scene->touch();
	return true;
}

// deterministic procedural density (config 4): sum of 8 Gaussian blobs + 3-octave value noise
static inline uint32_t pcgHash(uint32_t v) {
	uint32_t s = v * 747796405u + 2891336453u;
	uint32_t w = ((s >> ((s >> 28u) + 4u)) ^ s) * 277803737u;
	return (w >> 22u) ^ w;
}
static inline float hashf(uint32_t x, uint32_t y, uint32_t z, uint32_t seed) {
	return pcgHash(x + pcgHash(y + pcgHash(z + pcgHash(seed)))) * (1.f / 4294967296.f);
}
static float valueNoise(float x, float y, float z, uint32_t seed) {
	int xi = (int) std::floor(x), yi = (int) std::floor(y), zi = (int) std::floor(z);
	float fx = x - xi, fy = y - yi, fz = z - zi;
	auto s = [](float t) { return t * t * (3 - 2 * t); };
	float wx = s(fx), wy = s(fy), wz = s(fz), r = 0;
	for (int c = 0; c < 8; c++) {
		float w = (c & 1 ? wx : 1 - wx) * (c & 2 ? wy : 1 - wy) * (c & 4 ? wz : 1 - wz);
		r += w * hashf(xi + (c & 1), yi + ((c >> 1) & 1), zi + ((c >> 2) & 1), seed);
	}
	return r;
}
void proceduralDensity(std::vector<float> &out, const int res[3], uint64_t seed) {
	out.resize((size_t) res[0] * res[1] * res[2]);
	float blobs[8][4];
	for (int b = 0; b < 8; b++) {
		for (int k = 0; k < 3; k++) blobs[b][k] = 0.2f + 0.6f * hashf(b, k, 17, (uint32_t) seed);
		blobs[b][3] = 0.08f + 0.12f * hashf(b, 3, 17, (uint32_t) seed);
	}
	float mx = 0;
	for (int z = 0; z < res[2]; z++)
		for (int y = 0; y < res[1]; y++)
			for (int x = 0; x < res[0]; x++) {
				float p[3] = {(x + 0.5f) / res[0], (y + 0.5f) / res[1], (z + 0.5f) / res[2]};
				float d = 0;
				for (auto &b : blobs) {
					float r2 = (p[0] - b[0]) * (p[0] - b[0]) + (p[1] - b[1]) * (p[1] - b[1]) + (p[2] - b[2]) * (p[2] - b[2]);
					d += std::exp(-r2 / (2 * b[3] * b[3]));
				}
				float n = 0, amp = 0.5f, fr = 4;
				for (int o = 0; o < 3; o++, amp *= 0.5f, fr *= 2) n += amp * valueNoise(p[0] * fr, p[1] * fr, p[2] * fr, (uint32_t) seed + o);
				d = std::max(0.f, d * (0.4f + n));
				out[x + (size_t) res[0] * (y + (size_t) res[1] * z)] = d;
				mx = std::max(mx, d);
			}
	if (mx > 0) for (float &v : out) v /= mx; // max density 1
}

} // namespace krr
