// gltf.cpp -- minimal glTF 2.0 importer + PNG decoder of the host layer.
//
// The reference imports glTF through Assimp (src/scene/assimp.cpp:240-420) and maps materials in
// createMaterial(..., ImportMode::GLTF2) (assimp.cpp:93-226).  Assimp and stb_image are third-party
// libraries; this file reads the subset of the format the reference's assets use, straight from the glTF
// 2.0 specification:
//   * .gltf JSON with external .bin buffers, or binary .glb (JSON + BIN chunks, images in buffer views); no data: URIs; accessors of every component type,
//     strided buffer views, normalised integers;
//   * triangle primitives with POSITION, NORMAL, TEXCOORD_0, TANGENT and indices; one HostMesh per primitive;
//   * the node hierarchy (matrix or T * R * S), flattened to one instance per mesh node;
//   * pbrMetallicRoughness materials mapped exactly like createMaterial does (base colour -> diffuse,
//     roughness -> specular.g, metallic -> specular.b, MetallicRoughness shading model, emissive factor ->
//     constant emissive texture, KHR_materials_transmission / _emissive_strength), textures from 8-bit PNG
//     files (decoded with zlib; sRGB flag as Material::determineSrgb decides, texture.cpp:161-176);
//   * LINEAR / STEP animation samplers on translation / rotation / scale of any node (src/core/animation.cpp):
//     keys are resampled at the union of the node's channel times; an instance carries the chain of its animated
//     ancestors and the static transforms between them.
// Anything else (skins, morph targets, cameras, KTX / JPEG images) is reported or skipped
// with a message on stderr.
#include "krr_host.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>
#include <stdexcept>

namespace krr {

namespace {

string dirOfPath(const string &p) {
	size_t s = p.find_last_of("/\\");
	return s == string::npos ? string(".") : p.substr(0, s);
}
bool readAll(const string &path, std::vector<unsigned char> &out) {
	std::ifstream f(path, std::ios::binary | std::ios::ate);
	if (!f.good()) return false;
	std::streamsize n = f.tellg();
	f.seekg(0);
	out.resize((size_t) n);
	return n == 0 || (bool) f.read((char *) out.data(), n);
}
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
void trsToMat(const float t[3], const float q[4], const float s[3], float m[12]) {
	const double x = q[0], y = q[1], z = q[2], w = q[3];
	const double R[9] = {1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w), 2 * (x * y + z * w), 1 - 2 * (x * x + z * z),
						 2 * (y * z - x * w),	  2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)};
	for (int i = 0; i < 3; i++) {
		for (int j = 0; j < 3; j++) m[i * 4 + j] = (float) (R[i * 3 + j] * s[j]);
		m[i * 4 + 3] = t[i];
	}
}
float srgbToLinear(float c) { return c <= 0.04045f ? c / 12.92f : std::pow((c + 0.055f) / 1.055f, 2.4f); }

} // namespace

// =================================================================================================
// PNG: 8-bit grey / grey+alpha / RGB / RGBA / palette, non-interlaced (PNG specification, 2nd edition)
bool decodePNG(const std::vector<unsigned char> &buf, const string &path, Image &img, bool srgb, string *err);
bool loadPNG(const string &path, Image &img, bool srgb, string *err) {
	std::vector<unsigned char> buf;
	if (!readAll(path, buf)) { if (err) *err = path + ": cannot open"; return false; }
	return decodePNG(buf, path, img, srgb, err);
}
// the same from memory (images embedded in a .glb through a bufferView); `path` only names it in messages
bool decodePNG(const std::vector<unsigned char> &buf, const string &path, Image &img, bool srgb, string *err) {
	auto fail = [&](const string &m) { if (err) *err = path + ": " + m; return false; };
	static const unsigned char sig[8] = {0x89, 'P', 'N', 'G', 0x0d, 0x0a, 0x1a, 0x0a};
	if (buf.size() < 8 || memcmp(buf.data(), sig, 8)) return fail("not a PNG file");
	auto be32 = [&](size_t p) { return (uint32_t) buf[p] << 24 | (uint32_t) buf[p + 1] << 16 | (uint32_t) buf[p + 2] << 8 | buf[p + 3]; };
	uint32_t w = 0, h = 0;
	int depth = 0, ctype = 0, interlace = 0;
	std::vector<unsigned char> idat, palette, trns;
	for (size_t pos = 8; pos + 12 <= buf.size();) {
		const uint32_t len = be32(pos);
		const string type((const char *) &buf[pos + 4], 4);
		if (pos + 12 + len > buf.size()) return fail("truncated chunk");
		const unsigned char *d = &buf[pos + 8];
		if (type == "IHDR") { w = be32(pos + 8), h = be32(pos + 12), depth = d[8], ctype = d[9], interlace = d[12]; }
		else if (type == "PLTE") palette.assign(d, d + len);
		else if (type == "tRNS") trns.assign(d, d + len);
		else if (type == "IDAT") idat.insert(idat.end(), d, d + len);
		else if (type == "IEND") break;
		pos += 12 + len;
	}
	if (!w || !h) return fail("no IHDR");
	if (interlace) return fail("interlaced PNG images are not supported");
	if (depth != 1 && depth != 2 && depth != 4 && depth != 8 && depth != 16) return fail("bad bit depth");
	const int nch = ctype == 0 ? 1 : ctype == 2 ? 3 : ctype == 3 ? 1 : ctype == 4 ? 2 : ctype == 6 ? 4 : 0;
	if (!nch) return fail("bad colour type");
	if (depth < 8 && ctype != 0 && ctype != 3) return fail("bad bit depth for the colour type");
	const size_t rowBytes = ((size_t) w * nch * depth + 7) / 8; // bytes of one filtered scanline
	const size_t bpp	  = std::max<size_t>(1, (size_t) nch * depth / 8);
	std::vector<unsigned char> raw((rowBytes + 1) * h);
	uLongf got = (uLongf) raw.size();
	if (uncompress(raw.data(), &got, idat.data(), (uLong) idat.size()) != Z_OK || got != raw.size()) return fail("zlib error");
	std::vector<unsigned char> lines(rowBytes * h);
	for (uint32_t y = 0; y < h; y++) { // undo the scanline filters (byte-wise, distance bpp)
		const unsigned char *in = &raw[(rowBytes + 1) * y + 1];
		unsigned char *out = &lines[rowBytes * y];
		const unsigned char *up = y ? &lines[rowBytes * (y - 1)] : nullptr;
		const int ft = raw[(rowBytes + 1) * y];
		for (size_t x = 0; x < rowBytes; x++) {
			const int a = x >= bpp ? out[x - bpp] : 0, b = up ? up[x] : 0, c = (up && x >= bpp) ? up[x - bpp] : 0;
			int v = in[x];
			switch (ft) {
				case 1: v += a; break;
				case 2: v += b; break;
				case 3: v += (a + b) / 2; break;
				case 4: {
					const int p = a + b - c, pa = std::abs(p - a), pb = std::abs(p - b), pc = std::abs(p - c);
					v += (pa <= pb && pa <= pc) ? a : (pb <= pc ? b : c);
					break;
				}
				default: break;
			}
			out[x] = (unsigned char) v;
		}
	}
	// samples -> 8 bits per channel (16-bit: the high byte; 1/2/4-bit grey: scaled to 0..255; palette indices as is)
	const size_t stride = (size_t) w * nch;
	std::vector<unsigned char> pix(stride * h);
	for (uint32_t y = 0; y < h; y++)
		for (size_t x = 0; x < stride; x++) {
			const unsigned char *ln = &lines[rowBytes * y];
			unsigned v;
			if (depth == 8) v = ln[x];
			else if (depth == 16) v = ln[2 * x];
			else {
				const size_t bit = x * depth;
				v = (ln[bit / 8] >> (8 - depth - bit % 8)) & ((1u << depth) - 1);
				if (ctype == 0) v = v * 255 / ((1u << depth) - 1);
			}
			pix[stride * y + x] = (unsigned char) v;
		}
	img.width = (int) w, img.height = (int) h;
	img.rgba.resize((size_t) w * h * 4);
	for (size_t i = 0; i < (size_t) w * h; i++) {
		unsigned char r, g, b, a = 255;
		const unsigned char *p = &pix[i * nch];
		if (ctype == 0) r = g = b = p[0];
		else if (ctype == 4) r = g = b = p[0], a = p[1];
		else if (ctype == 3) {
			if ((size_t) p[0] * 3 + 2 >= palette.size()) return fail("palette index out of range");
			r = palette[p[0] * 3], g = palette[p[0] * 3 + 1], b = palette[p[0] * 3 + 2];
			if (p[0] < trns.size()) a = trns[p[0]];
		} else { r = p[0], g = p[1], b = p[2]; if (nch == 4) a = p[3]; }
		float c[4] = {r / 255.f, g / 255.f, b / 255.f, a / 255.f};
		if (srgb) for (int k = 0; k < 3; k++) c[k] = srgbToLinear(c[k]); // texDesc.sRGB, texture.cpp:249
		memcpy(&img.rgba[i * 4], c, 16);
	}
	return true;
}

// =================================================================================================
namespace {

struct Gltf {
	json doc;
	string dir;
	std::vector<std::vector<unsigned char>> buffers;

	const unsigned char *viewData(int view, size_t &stride, size_t &size) {
		const json &bv = doc.at("bufferViews").at((size_t) view);
		const int b = (int) bv.at("buffer").asNumber();
		const size_t off = bv.contains("byteOffset") ? (size_t) bv.at("byteOffset").asNumber() : 0;
		size   = (size_t) bv.at("byteLength").asNumber();
		stride = bv.contains("byteStride") ? (size_t) bv.at("byteStride").asNumber() : 0;
		if (b < 0 || (size_t) b >= buffers.size() || off + size > buffers[b].size()) throw std::runtime_error("glTF: buffer view out of range");
		return buffers[b].data() + off;
	}
	// accessor -> floats, `comps` per element (normalised integers are scaled as the specification says)
	std::vector<float> readFloats(int accessor, int &comps, size_t &count) {
		const json &a = doc.at("accessors").at((size_t) accessor);
		const string type = a.at("type").asString();
		comps = type == "SCALAR" ? 1 : type == "VEC2" ? 2 : type == "VEC3" ? 3 : type == "VEC4" ? 4 : type == "MAT4" ? 16 : 0;
		if (!comps) throw std::runtime_error("glTF: unsupported accessor type " + type);
		count = (size_t) a.at("count").asNumber();
		const int ct = (int) a.at("componentType").asNumber();
		const size_t cs = ct == 5120 || ct == 5121 ? 1 : ct == 5122 || ct == 5123 ? 2 : 4;
		const bool norm = a.value("normalized", false);
		std::vector<float> out(count * comps, 0.f);
		if (!a.contains("bufferView")) return out; // all zeros (sparse accessors are not supported)
		size_t stride, size;
		const unsigned char *base = viewData((int) a.at("bufferView").asNumber(), stride, size);
		const size_t off = a.contains("byteOffset") ? (size_t) a.at("byteOffset").asNumber() : 0;
		if (!stride) stride = cs * comps;
		if (count && off + stride * (count - 1) + cs * comps > size) throw std::runtime_error("glTF: accessor out of range");
		for (size_t i = 0; i < count; i++)
			for (int c = 0; c < comps; c++) {
				const unsigned char *p = base + off + stride * i + cs * c;
				float v;
				switch (ct) {
					case 5120: { int8_t x; memcpy(&x, p, 1); v = norm ? std::max(x / 127.f, -1.f) : x; break; }
					case 5121: { uint8_t x; memcpy(&x, p, 1); v = norm ? x / 255.f : x; break; }
					case 5122: { int16_t x; memcpy(&x, p, 2); v = norm ? std::max(x / 32767.f, -1.f) : x; break; }
					case 5123: { uint16_t x; memcpy(&x, p, 2); v = norm ? x / 65535.f : x; break; }
					case 5125: { uint32_t x; memcpy(&x, p, 4); v = (float) x; break; }
					case 5126: memcpy(&v, p, 4); break;
					default: throw std::runtime_error("glTF: unsupported component type");
				}
				out[i * comps + c] = v;
			}
		return out;
	}
};

void nodeLocal(const json &n, float m[12]) {
	if (n.contains("matrix")) { // column-major 4x4
		const json &a = n.at("matrix");
		for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) m[r * 4 + c] = (float) a.at((size_t) (c * 4 + r)).asNumber();
		return;
	}
	float t[3] = {0, 0, 0}, q[4] = {0, 0, 0, 1}, s[3] = {1, 1, 1};
	if (n.contains("translation")) for (int k = 0; k < 3; k++) t[k] = (float) n.at("translation").at((size_t) k).asNumber();
	if (n.contains("rotation")) for (int k = 0; k < 4; k++) q[k] = (float) n.at("rotation").at((size_t) k).asNumber();
	if (n.contains("scale")) for (int k = 0; k < 3; k++) s[k] = (float) n.at("scale").at((size_t) k).asNumber();
	trsToMat(t, q, s, m);
}

// createMaterial(..., ImportMode::GLTF2), assimp.cpp:93-226
HostMaterial materialFromGltf(Gltf &g, const json &m, int index) {
	HostMaterial out;
	out.name = m.value("name", "unnamed");
	KrrMaterialDesc &d = out.desc;
	memset(&d, 0, sizeof d);
	d.diffuse[0] = d.diffuse[1] = d.diffuse[2] = d.diffuse[3] = 1; // MaterialParams defaults, texture.h:129-137
	d.ior			= 1.5f;
	d.bsdf_type		= KRR_MAT_DISNEY;
	d.shading_model = KRR_SHADING_METALLIC_ROUGHNESS; // assimp.cpp:222-224
	d.color_space	= 0;
	out.images.resize(KRR_TEX_COUNT);
	auto loadTex = [&](const json &ref, int slot, bool srgb) {
		const int ti	= (int) ref.at("index").asNumber();
		const json &tex = g.doc.at("textures").at((size_t) ti);
		if (!tex.contains("source")) return;
		const json &im = g.doc.at("images").at((size_t) tex.at("source").asNumber());
		Image img;
		string err;
		bool ok = false;
		if (im.contains("uri") && im.at("uri").asString().rfind("data:", 0) != 0) ok = loadPNG(g.dir + "/" + im.at("uri").asString(), img, srgb, &err);
		else if (im.contains("bufferView")) { // image stored in a buffer (.glb)
			const json &bv = g.doc.at("bufferViews").at((size_t) im.at("bufferView").asNumber());
			const size_t off = (size_t) bv.value("byteOffset", 0.0), len = (size_t) bv.at("byteLength").asNumber();
			const int b = (int) bv.at("buffer").asNumber();
			if (b < 0 || (size_t) b >= g.buffers.size() || off + len > g.buffers[b].size()) err = "image buffer view out of range";
			else ok = decodePNG(std::vector<unsigned char>(g.buffers[b].begin() + off, g.buffers[b].begin() + off + len), "embedded image", img, srgb, &err);
		} else err = "data: URIs are not supported";
		if (!ok) { fprintf(stderr, "glTF: material %d: texture skipped (%s)\n", index, err.c_str()); return; }
		out.images[slot] = img.rgba;
		KrrTextureDesc &t = d.textures[slot];
		t.valid = 1, t.width = img.width, t.height = img.height;
		t.value[0] = t.value[1] = t.value[2] = t.value[3] = 1;
	};
	float metallic = 1, roughness = 1; // glTF defaults; Assimp reports them, so the reference always sets both
	if (m.contains("pbrMetallicRoughness")) {
		const json &p = m.at("pbrMetallicRoughness");
		if (p.contains("baseColorFactor")) for (int k = 0; k < 4; k++) d.diffuse[k] = (float) p.at("baseColorFactor").at((size_t) k).asNumber();
		metallic  = p.value("metallicFactor", 1.f);
		roughness = p.value("roughnessFactor", 1.f);
		// determineSrgb runs while the material still has its default SpecularGlossiness model (the model is
		// only set at the end of createMaterial), so the metallic-roughness texture is read as sRGB too
		if (p.contains("baseColorTexture")) loadTex(p.at("baseColorTexture"), KRR_TEX_DIFFUSE, true);
		if (p.contains("metallicRoughnessTexture")) loadTex(p.at("metallicRoughnessTexture"), KRR_TEX_SPECULAR, true);
	}
	d.specular[1] = roughness, d.specular[2] = metallic;
	if (m.contains("normalTexture")) loadTex(m.at("normalTexture"), KRR_TEX_NORMAL, false);
	if (m.contains("emissiveTexture")) loadTex(m.at("emissiveTexture"), KRR_TEX_EMISSIVE, true);
	float emissive[3] = {0, 0, 0}, strength = 0;
	if (m.contains("emissiveFactor")) for (int k = 0; k < 3; k++) emissive[k] = (float) m.at("emissiveFactor").at((size_t) k).asNumber();
	if (m.contains("extensions")) {
		const json &e = m.at("extensions");
		if (e.contains("KHR_materials_emissive_strength")) {
			strength = e.at("KHR_materials_emissive_strength").value("emissiveStrength", 1.f);
			for (float &c : emissive) c *= strength; // assimp.cpp:173-176
		}
		if (e.contains("KHR_materials_transmission")) {
			const float tr			= e.at("KHR_materials_transmission").value("transmissionFactor", 0.f);
			d.specular_transmission = tr;
			if (tr > 1 - 1e-5f) d.bsdf_type = KRR_MAT_DIELECTRIC;
		}
		if (e.contains("KHR_materials_ior")) d.ior = e.at("KHR_materials_ior").value("ior", 1.5f);
	}
	if ((emissive[0] != 0 || emissive[1] != 0 || emissive[2] != 0) && !d.textures[KRR_TEX_EMISSIVE].valid) {
		KrrTextureDesc &t = d.textures[KRR_TEX_EMISSIVE]; // setConstantTexture(Emissive, RGBA(emissive, 1))
		t.valid = 1, t.value[0] = emissive[0], t.value[1] = emissive[1], t.value[2] = emissive[2], t.value[3] = 1;
	}
	return out;
}

} // namespace

bool loadGltf(const string &filepath, Scene &scene, const float nodeTransform[12]) {
	Gltf g;
	std::vector<unsigned char> glbBin; // BIN chunk of a .glb: the buffer that has no uri
	bool haveGlbBin = false;
	{
		std::vector<unsigned char> file;
		if (!readAll(filepath, file)) throw std::runtime_error("cannot open model " + filepath);
		string text;
		auto u32 = [&](size_t o) { uint32_t v; memcpy(&v, &file[o], 4); return v; };
		if (file.size() >= 12 && u32(0) == 0x46546C67u) { // binary glTF: 12-byte header, then chunks (length, type, data)
			if (u32(4) != 2) throw std::runtime_error("glTF: unsupported .glb version in " + filepath);
			size_t pos = 12, end = std::min<size_t>(u32(8), file.size());
			while (pos + 8 <= end) {
				const size_t len = u32(pos);
				const uint32_t type = u32(pos + 4);
				if (pos + 8 + len > end) throw std::runtime_error("glTF: truncated .glb chunk in " + filepath);
				if (type == 0x4E4F534Au) text.assign((const char *) &file[pos + 8], len);
				else if (type == 0x004E4942u && !haveGlbBin) glbBin.assign(file.begin() + pos + 8, file.begin() + pos + 8 + len), haveGlbBin = true;
				pos += 8 + ((len + 3) & ~(size_t) 3);
			}
			if (text.empty()) throw std::runtime_error("glTF: .glb without a JSON chunk: " + filepath);
		} else text.assign(file.begin(), file.end());
		g.doc = json::parse(text);
	}
	g.dir = dirOfPath(filepath);
	if (g.doc.contains("buffers"))
		for (const json &b : g.doc.at("buffers").items()) {
			if (!b.contains("uri")) {
				if (!haveGlbBin) throw std::runtime_error("glTF: buffer without uri outside a .glb (" + filepath + ")");
				g.buffers.push_back(glbBin);
				continue;
			}
			if (b.at("uri").asString().rfind("data:", 0) == 0) throw std::runtime_error("glTF: data: URIs are not supported (" + filepath + ")");
			std::vector<unsigned char> data;
			if (!readAll(g.dir + "/" + b.at("uri").asString(), data)) throw std::runtime_error("glTF: cannot open buffer " + b.at("uri").asString());
			g.buffers.push_back(std::move(data));
		}
	// materials
	const int matBase = (int) scene.materials.size();
	int nMat = 0;
	if (g.doc.contains("materials"))
		for (const json &m : g.doc.at("materials").items()) scene.materials.push_back(materialFromGltf(g, m, nMat++));
	int defaultMat = -1;
	// meshes: one HostMesh per triangle primitive
	std::vector<std::vector<int>> meshPrims;
	if (g.doc.contains("meshes"))
		for (const json &m : g.doc.at("meshes").items()) {
			std::vector<int> prims;
			for (const json &p : m.at("primitives").items()) {
				if (p.contains("mode") && (int) p.at("mode").asNumber() != 4) { fprintf(stderr, "glTF: non-triangle primitive skipped\n"); continue; }
				const json &at = p.at("attributes");
				if (!at.contains("POSITION")) continue;
				HostMesh hm;
				hm.name = m.value("name", "mesh");
				int comps;
				size_t nv, n;
				hm.positions = g.readFloats((int) at.at("POSITION").asNumber(), comps, nv);
				if (comps != 3) throw std::runtime_error("glTF: POSITION must be VEC3");
				if (at.contains("NORMAL")) { hm.normals = g.readFloats((int) at.at("NORMAL").asNumber(), comps, n); if (n != nv || comps != 3) hm.normals.clear(); }
				if (at.contains("TEXCOORD_0")) { hm.texcoords = g.readFloats((int) at.at("TEXCOORD_0").asNumber(), comps, n); if (n != nv || comps != 2) hm.texcoords.clear(); }
				if (at.contains("TANGENT")) {
					std::vector<float> t4 = g.readFloats((int) at.at("TANGENT").asNumber(), comps, n);
					if (n == nv && comps == 4) { hm.tangents.resize(nv * 3); for (size_t i = 0; i < nv; i++) for (int k = 0; k < 3; k++) hm.tangents[3 * i + k] = t4[4 * i + k]; }
				}
				if (p.contains("indices")) {
					std::vector<float> idx = g.readFloats((int) p.at("indices").asNumber(), comps, n);
					hm.indices.resize(n / 3 * 3);
					for (size_t i = 0; i < hm.indices.size(); i++) {
						hm.indices[i] = (int32_t) idx[i];
						if (hm.indices[i] < 0 || (size_t) hm.indices[i] >= nv) throw std::runtime_error("glTF: vertex index out of range");
					}
				} else { hm.indices.resize(nv / 3 * 3); for (size_t i = 0; i < hm.indices.size(); i++) hm.indices[i] = (int32_t) i; }
				if (hm.indices.empty()) continue;
				if (p.contains("material")) hm.material = matBase + (int) p.at("material").asNumber();
				else {
					if (defaultMat < 0) { // the glTF default material: white, metallic 1, roughness 1
						json none = json::object();
						scene.materials.push_back(materialFromGltf(g, none, -1));
						defaultMat = (int) scene.materials.size() - 1;
					}
					hm.material = defaultMat;
				}
				prims.push_back((int) scene.meshes.size());
				scene.meshes.push_back(std::move(hm));
			}
			meshPrims.push_back(prims);
		}
	// animation channels per node
	struct Channel { string path; std::vector<float> times, values; bool step; };
	std::vector<std::vector<Channel>> nodeChannels(g.doc.contains("nodes") ? g.doc.at("nodes").size() : 0);
	if (g.doc.contains("animations"))
		for (const json &an : g.doc.at("animations").items())
			for (const json &ch : an.at("channels").items()) {
				const json &tg = ch.at("target");
				if (!tg.contains("node")) continue;
				const json &sm = an.at("samplers").at((size_t) ch.at("sampler").asNumber());
				Channel c;
				c.path = tg.at("path").asString();
				c.step = sm.value("interpolation", "LINEAR") == "STEP";
				if (sm.value("interpolation", "LINEAR") == "CUBICSPLINE") { fprintf(stderr, "glTF: CUBICSPLINE animation skipped\n"); continue; }
				int comps;
				size_t n, nvals;
				c.times	 = g.readFloats((int) sm.at("input").asNumber(), comps, n);
				c.values = g.readFloats((int) sm.at("output").asNumber(), comps, nvals);
				if (n < 1 || nvals != n || (c.path != "translation" && c.path != "rotation" && c.path != "scale")) continue;
				nodeChannels[(size_t) tg.at("node").asNumber()].push_back(c);
			}
	// scene graph -> instances
	struct Walk {
		Gltf &g; Scene &scene; std::vector<std::vector<int>> &meshPrims; std::vector<std::vector<Channel>> &chan;
		// keys of an animated node at the union of its channel times; every key holds the node's full local TRS
		void buildKeys(int ni, const json &n, std::vector<float> &outTimes, std::vector<KrrSRT> &outKeys) {
			std::set<float> times;
			for (const Channel &c : chan[ni]) times.insert(c.times.begin(), c.times.end());
			float t0[3] = {0, 0, 0}, q0[4] = {0, 0, 0, 1}, s0[3] = {1, 1, 1};
			if (n.contains("translation")) for (int k = 0; k < 3; k++) t0[k] = (float) n.at("translation").at((size_t) k).asNumber();
			if (n.contains("rotation")) for (int k = 0; k < 4; k++) q0[k] = (float) n.at("rotation").at((size_t) k).asNumber();
			if (n.contains("scale")) for (int k = 0; k < 3; k++) s0[k] = (float) n.at("scale").at((size_t) k).asNumber();
			for (float t : times) {
				KrrSRT key;
				memcpy(key.t, t0, 12), memcpy(key.q, q0, 16), memcpy(key.s, s0, 12);
				for (const Channel &c : chan[ni]) {
					const int w = c.path == "rotation" ? 4 : 3;
					size_t k = 0;
					while (k + 2 < c.times.size() && t >= c.times[k + 1]) k++;
					float v[4];
					if (c.times.size() == 1 || t <= c.times[0]) memcpy(v, &c.values[0], w * 4);
					else if (t >= c.times.back()) memcpy(v, &c.values[(c.times.size() - 1) * w], w * 4);
					else {
						const float a = c.step ? 0.f : (t - c.times[k]) / (c.times[k + 1] - c.times[k]);
						const float *A = &c.values[k * w], *B = &c.values[(k + 1) * w];
						float dq = 0;
						if (w == 4) for (int j = 0; j < 4; j++) dq += A[j] * B[j];
						for (int j = 0; j < w; j++) v[j] = (1 - a) * A[j] + a * ((w == 4 && dq < 0) ? -B[j] : B[j]);
					}
					if (c.path == "translation") memcpy(key.t, v, 12);
					else if (c.path == "scale") memcpy(key.s, v, 12);
					else {
						float len = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2] + v[3] * v[3]);
						for (int j = 0; j < 4; j++) key.q[j] = len > 0 ? v[j] / len : (j == 3 ? 1.f : 0.f);
					}
				}
				outTimes.push_back(t);
				outKeys.push_back(key);
			}
		}
		// parent: world transform of the parent at rest; chain: the animated ancestors (root first); since: product of
		// the static node transforms below the last animated ancestor (below the root when there is none)
		void node(int ni, const float parent[12], const std::vector<AnimLink> &chain, const float since[12]) {
			const json &n = g.doc.at("nodes").at((size_t) ni);
			float local[12], world[12];
			nodeLocal(n, local);
			mul12(parent, local, world);
			const bool moves = !chan[ni].empty();
			std::vector<float> times;
			std::vector<KrrSRT> keys;
			if (moves) buildKeys(ni, n, times, keys);
			if (n.contains("mesh")) {
				for (int mi : meshPrims[(size_t) n.at("mesh").asNumber()]) {
					HostInstance in;
					in.mesh = mi;
					memcpy(in.transform, world, sizeof world);
					if (moves || !chain.empty()) {
						in.animAncestors = chain;
						if (moves) { // world = chain * since * SRT(t)
							in.animTimes = times, in.animKeys = keys;
							memcpy(in.animParent, since, sizeof in.animParent);
						} else mul12(since, local, in.animParent); // world = chain * (since * local)
						scene.animated = true;
					}
					scene.instances.push_back(in);
				}
			}
			if (n.contains("children")) {
				std::vector<AnimLink> below = chain;
				float sinceBelow[12];
				if (moves) {
					AnimLink link;
					memcpy(link.pre, since, sizeof link.pre);
					link.times = times, link.keys = keys;
					below.push_back(std::move(link));
					const float I[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
					memcpy(sinceBelow, I, sizeof I);
				} else mul12(since, local, sinceBelow);
				for (const json &c : n.at("children").items()) node((int) c.asNumber(), world, below, sinceBelow);
			}
		}
	} walk{g, scene, meshPrims, nodeChannels};
	if (g.doc.contains("scenes")) {
		const size_t si = g.doc.contains("scene") ? (size_t) g.doc.at("scene").asNumber() : 0;
		for (const json &r : g.doc.at("scenes").at(si).at("nodes").items()) walk.node((int) r.asNumber(), nodeTransform, {}, nodeTransform);
	} else if (g.doc.contains("nodes")) { // no scene: every node that is nobody's child is a root
		std::vector<char> isChild(g.doc.at("nodes").size(), 0);
		for (const json &n : g.doc.at("nodes").items())
			if (n.contains("children")) for (const json &c : n.at("children").items()) isChild[(size_t) c.asNumber()] = 1;
		for (size_t i = 0; i < isChild.size(); i++) if (!isChild[i]) walk.node((int) i, nodeTransform, {}, nodeTransform);
	}
	scene.touch();
	return true;
}

} // namespace krr
