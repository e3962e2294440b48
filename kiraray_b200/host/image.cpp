// image.cpp -- HDR image files of the passes around the path tracer: OpenEXR (scanline) and PFM.
//
// Reference: Image::loadImage / saveImage (src/core/texture.cpp:27-118), tinyexr::save_exr / load_exr
// (src/util/image.cpp:29-104, 106-330) and pfm::ReadImagePFM (src/util/image.cpp:345-430).  tinyexr and
// stb are third-party code the reference vendors; here the two formats the render passes actually use
// (AccumulatePass::saveImage writes .exr, ErrorMeasurePass::loadReferenceImage reads .exr / .pfm) are
// implemented directly from the file-format specifications:
//   * EXR: single-part scanline files; pixel types HALF / FLOAT / UINT; compression NONE, ZIPS, ZIP (read + write), PIZ (read)
//     (zlib); any line order; channels by name (R, G, B, A; one channel -> grey).  Tiled, multi-part, deep
//     and the lossy / wavelet codecs are reported as errors.
//   * PFM: "PF" / "Pf", either endianness, bottom-up rows.
// Images are RGBA32F, row 0 first, exactly the layout of the film buffer (RenderContext).
#include "krr_host.h"

#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

namespace krr {

namespace {

// ---- half <-> float (IEEE 754 binary16, round to nearest even) ----
uint16_t floatToHalf(float f) {
	uint32_t x;
	memcpy(&x, &f, 4);
	uint32_t sign = (x >> 16) & 0x8000u, mant = x & 0x7fffffu;
	int exp = (int) ((x >> 23) & 0xff) - 127 + 15;
	if (((x >> 23) & 0xff) == 0xff) return (uint16_t) (sign | 0x7c00u | (mant ? 0x200u | (mant >> 13) : 0)); // inf / nan
	if (exp >= 31) return (uint16_t) (sign | 0x7c00u);																	  // overflow -> inf
	if (exp <= 0) {																											  // subnormal or zero
		if (exp < -10) return (uint16_t) sign;
		mant |= 0x800000u;
		int shift	  = 14 - exp;
		uint32_t half = mant >> shift, rem = mant & ((1u << shift) - 1), mid = 1u << (shift - 1);
		if (rem > mid || (rem == mid && (half & 1))) half++;
		return (uint16_t) (sign | half);
	}
	uint32_t half = (uint32_t) (exp << 10) | (mant >> 13), rem = mant & 0x1fffu;
	if (rem > 0x1000u || (rem == 0x1000u && (half & 1))) half++; // may carry into the exponent: still correct
	return (uint16_t) (sign | half);
}
float halfToFloat(uint16_t h) {
	uint32_t sign = (uint32_t) (h & 0x8000u) << 16, exp = (h >> 10) & 0x1f, mant = h & 0x3ffu, x;
	if (exp == 0) {
		if (mant == 0) x = sign;
		else {
			int e = -1;
			do { mant <<= 1; e++; } while (!(mant & 0x400u));
			x = sign | (uint32_t) (127 - 15 - e) << 23 | (mant & 0x3ffu) << 13;
		}
	} else if (exp == 31) x = sign | 0x7f800000u | mant << 13;
	else x = sign | (exp + 127 - 15) << 23 | mant << 13;
	float f;
	memcpy(&f, &x, 4);
	return f;
}

bool endsWith(const string &s, const char *suffix) {
	size_t n = strlen(suffix);
	if (s.size() < n) return false;
	for (size_t i = 0; i < n; i++)
		if (tolower((unsigned char) s[s.size() - n + i]) != suffix[i]) return false;
	return true;
}

bool fail(string *err, const string &msg) {
	if (err) *err = msg;
	return false;
}

bool readFile(const string &path, std::vector<unsigned char> &out) {
	std::ifstream f(path, std::ios::binary | std::ios::ate);
	if (!f.good()) return false;
	std::streamsize n = f.tellg();
	f.seekg(0);
	out.resize((size_t) n);
	return n == 0 || (bool) f.read((char *) out.data(), n);
}

// ---- EXR ZIP codec: byte de-interleave + delta predictor around zlib (OpenEXR ImfZip) ----
void zipPredictEncode(const unsigned char *raw, size_t n, std::vector<unsigned char> &tmp) {
	tmp.resize(n);
	size_t half = (n + 1) / 2;
	for (size_t i = 0; i < n; i++) tmp[(i & 1) ? half + i / 2 : i / 2] = raw[i];
	int p = tmp.empty() ? 0 : tmp[0];
	for (size_t i = 1; i < n; i++) {
		int d  = (int) tmp[i] - p + (128 + 256);
		p	   = tmp[i];
		tmp[i] = (unsigned char) d;
	}
}
void zipPredictDecode(std::vector<unsigned char> &tmp, unsigned char *raw) {
	size_t n = tmp.size();
	for (size_t i = 1; i < n; i++) tmp[i] = (unsigned char) ((int) tmp[i - 1] + (int) tmp[i] - 128);
	size_t half = (n + 1) / 2;
	for (size_t i = 0; i < n; i++) raw[i] = (i & 1) ? tmp[half + i / 2] : tmp[i / 2];
}

struct ExrChannel {
	string name;
	int type = 1; // 0 uint, 1 half, 2 float
	int xs = 1, ys = 1;
};

} // namespace

// =================================================================================================
bool loadPFM(const string &path, Image &img, string *err) {
	std::vector<unsigned char> buf;
	if (!readFile(path, buf)) return fail(err, "cannot open " + path);
	size_t pos = 0;
	auto word  = [&](string &w) {
		 w.clear();
		 while (pos < buf.size() && isspace(buf[pos])) pos++;
		 while (pos < buf.size() && !isspace(buf[pos])) w.push_back((char) buf[pos++]);
		 return !w.empty();
	};
	string w;
	if (!word(w) || (w != "PF" && w != "Pf")) return fail(err, path + ": not a PFM file");
	const int nc = w == "PF" ? 3 : 1;
	string ws, hs, ss;
	if (!word(ws) || !word(hs) || !word(ss)) return fail(err, path + ": truncated PFM header");
	const int width = atoi(ws.c_str()), height = atoi(hs.c_str());
	const float scale = (float) atof(ss.c_str());
	pos++; // the single whitespace byte that ends the header
	if (width <= 0 || height <= 0 || scale == 0) return fail(err, path + ": bad PFM header");
	const size_t need = (size_t) width * height * nc * 4;
	if (buf.size() - pos < need) return fail(err, path + ": truncated PFM data");
	const bool fileLittle = scale < 0.f;
	const uint16_t one	  = 1;
	const bool hostLittle = *(const unsigned char *) &one == 1;
	img.width = width, img.height = height;
	img.rgba.assign((size_t) width * height * 4, 1.f);
	for (int fy = 0; fy < height; fy++) { // P*M stores the bottom row first
		const int y = height - 1 - fy;
		for (int x = 0; x < width; x++)
			for (int c = 0; c < nc; c++) {
				unsigned char b[4];
				memcpy(b, &buf[pos + (((size_t) fy * width + x) * nc + c) * 4], 4);
				if (hostLittle != fileLittle) { std::swap(b[0], b[3]); std::swap(b[1], b[2]); }
				float v;
				memcpy(&v, b, 4);
				if (std::fabs(scale) != 1.f) v *= std::fabs(scale);
				float *px = &img.rgba[((size_t) y * width + x) * 4];
				if (nc == 1) px[0] = px[1] = px[2] = px[3] = v; // RGBA(data[i]), image.cpp:415-417
				else px[c] = v;
			}
	}
	return true;
}

bool savePFM(const string &path, const Image &img, string *err) {
	FILE *f = fopen(path.c_str(), "wb");
	if (!f) return fail(err, "cannot write " + path);
	fprintf(f, "PF\n%d %d\n-1.0\n", img.width, img.height);
	std::vector<float> row((size_t) img.width * 3);
	for (int fy = 0; fy < img.height; fy++) {
		const float *src = &img.rgba[(size_t) (img.height - 1 - fy) * img.width * 4];
		for (int x = 0; x < img.width; x++)
			for (int c = 0; c < 3; c++) row[3 * x + c] = src[4 * x + c];
		fwrite(row.data(), 4, row.size(), f);
	}
	fclose(f);
	return true;
}

// =================================================================================================

// ---- EXR PIZ codec, decoding side (OpenEXR ImfPizCompressor / ImfHuf / ImfWav, restated from the
// published format): a chunk is  minNonZero u16 | maxNonZero u16 | bitmap[min..max] | length i32 |
// Huffman stream.  Decoding: canonical Huffman with a run-length symbol -> 16-bit words, inverse 2-D Haar
// wavelet per channel (per 16-bit half of a 32-bit sample), inverse value look-up table, then the
// channel-planar block is re-interleaved scanline by scanline.  The reference's sky.exr is PIZ. ----
namespace piz {
constexpr int kEncBits = 16, kDecBits = 14, kEncSize = (1 << kEncBits) + 1, kDecSize = 1 << kDecBits, kDecMask = kDecSize - 1;
constexpr int kShortZeroRun = 59, kLongZeroRun = 63, kShortestLongRun = 2 + kLongZeroRun - kShortZeroRun;

struct Dec {
	int len = 0;			// short code: its length, symbol in lit
	int lit = 0;
	std::vector<int> longs; // long codes that share this 14-bit prefix
};

struct BitReader {
	const unsigned char *p, *end;
	uint64_t c = 0;
	int lc	   = 0;
	bool ok	   = true;
	uint32_t get(int n) {
		while (lc < n) {
			if (p >= end) { ok = false; return 0; }
			c = (c << 8) | *p++, lc += 8;
		}
		lc -= n;
		return (uint32_t) ((c >> lc) & ((1u << n) - 1));
	}
};

// code lengths (6 bits each, zero runs packed) -> canonical codes: hcode[i] = length | code << 6
bool unpackEncTable(const unsigned char *&ptr, const unsigned char *end, int im, int iM, std::vector<uint64_t> &hcode) {
	hcode.assign(kEncSize, 0);
	BitReader br{ptr, end};
	for (; im <= iM; im++) {
		uint64_t l = hcode[im] = br.get(6);
		if (!br.ok) return false;
		if (l == (uint64_t) kLongZeroRun) {
			int zerun = (int) br.get(8) + kShortestLongRun;
			if (!br.ok || im + zerun > iM + 1) return false;
			while (zerun--) hcode[im++] = 0;
			im--;
		} else if (l >= (uint64_t) kShortZeroRun) {
			int zerun = (int) l - kShortZeroRun + 2;
			if (im + zerun > iM + 1) return false;
			while (zerun--) hcode[im++] = 0;
			im--;
		}
	}
	ptr = br.p;
	uint64_t n[59] = {};
	for (int i = 0; i < kEncSize; i++) n[hcode[i]]++;
	uint64_t c = 0;
	for (int i = 58; i > 0; --i) {
		uint64_t nc = (c + n[i]) >> 1;
		n[i] = c, c = nc;
	}
	for (int i = 0; i < kEncSize; i++) {
		uint64_t l = hcode[i];
		if (l > 0) hcode[i] = l | (n[l]++ << 6);
	}
	return true;
}

bool buildDecTable(const std::vector<uint64_t> &hcode, int im, int iM, std::vector<Dec> &dec) {
	dec.assign(kDecSize, Dec());
	for (; im <= iM; im++) {
		const uint64_t c = hcode[im] >> 6;
		const int l		 = (int) (hcode[im] & 63);
		if (c >> l) return false;
		if (l > kDecBits) {
			Dec &pl = dec[c >> (l - kDecBits)];
			if (pl.len) return false;
			pl.longs.push_back(im);
		} else if (l) {
			Dec *pl = &dec[c << (kDecBits - l)];
			for (uint64_t i = 1ull << (kDecBits - l); i > 0; i--, pl++) {
				if (pl->len || !pl->longs.empty()) return false;
				pl->len = l, pl->lit = im;
			}
		}
	}
	return true;
}

bool decode(const std::vector<uint64_t> &hcode, const std::vector<Dec> &dec, const unsigned char *in, const unsigned char *fileEnd, int nBits, int rlc,
			size_t no, uint16_t *out) {
	uint64_t c = 0;
	int lc	   = 0;
	uint16_t *const ob = out, *const oe = out + no;
	const unsigned char *ie = in + (nBits + 7) / 8;
	if (ie > fileEnd) return false;
	auto emit = [&](int po) { // a literal, or the run-length symbol followed by an 8-bit repeat count
		if (po == rlc) {
			if (lc < 8) {
				if (in >= ie) return false;
				c = (c << 8) | *in++, lc += 8;
			}
			lc -= 8;
			int cs = (int) ((c >> lc) & 0xff);
			if (out + cs > oe || out == ob) return false;
			const uint16_t s = out[-1];
			while (cs-- > 0) *out++ = s;
		} else if (out < oe) *out++ = (uint16_t) po;
		else return false;
		return true;
	};
	while (in < ie) {
		c = (c << 8) | *in++, lc += 8;
		while (lc >= kDecBits) {
			const Dec &pl = dec[(c >> (lc - kDecBits)) & kDecMask];
			if (pl.len) {
				lc -= pl.len;
				if (!emit(pl.lit)) return false;
			} else {
				if (pl.longs.empty()) return false;
				size_t j = 0;
				for (; j < pl.longs.size(); j++) {
					const int sym = pl.longs[j], l = (int) (hcode[sym] & 63);
					while (lc < l && in < ie) c = (c << 8) | *in++, lc += 8;
					if (lc >= l && (hcode[sym] >> 6) == ((c >> (lc - l)) & ((1ull << l) - 1))) {
						lc -= l;
						if (!emit(sym)) return false;
						break;
					}
				}
				if (j == pl.longs.size()) return false;
			}
		}
	}
	const int i = (8 - nBits) & 7; // bits of the last byte that are padding
	c >>= i, lc -= i;
	while (lc > 0) {
		const Dec &pl = dec[(c << (kDecBits - lc)) & kDecMask];
		if (!pl.len) return false;
		lc -= pl.len;
		if (!emit(pl.lit)) return false;
	}
	return out == oe;
}

bool hufUncompress(const unsigned char *in, size_t nIn, uint16_t *out, size_t nOut) {
	if (nIn == 0) return nOut == 0;
	if (nIn < 20) return false;
	auto u32 = [&](size_t o) { uint32_t v; memcpy(&v, in + o, 4); return v; };
	const int im = (int) u32(0), iM = (int) u32(4), nBits = (int) u32(12);
	if (im < 0 || im >= kEncSize || iM < 0 || iM >= kEncSize || nBits < 0) return false;
	const unsigned char *ptr = in + 20, *end = in + nIn;
	std::vector<uint64_t> hcode;
	std::vector<Dec> dec;
	if (!unpackEncTable(ptr, end, im, iM, hcode)) return false;
	if ((size_t) (nBits + 7) / 8 > (size_t) (end - ptr)) return false;
	if (!buildDecTable(hcode, im, iM, dec)) return false;
	return decode(hcode, dec, ptr, end, nBits, iM, nOut, out);
}

// inverse wavelet butterflies: 14-bit data (signed arithmetic) and full 16-bit data (modulo arithmetic)
inline void wdec14(uint16_t l, uint16_t h, uint16_t &a, uint16_t &b) {
	const int16_t ls = (int16_t) l, hs = (int16_t) h;
	const int hi = hs, ai = ls + (hi & 1) + (hi >> 1);
	a = (uint16_t) (int16_t) ai, b = (uint16_t) (int16_t) (ai - hi);
}
inline void wdec16(uint16_t l, uint16_t h, uint16_t &a, uint16_t &b) {
	const int m = l, d = h;
	const int bb = (m - (d >> 1)) & 0xffff, aa = (d + bb - (1 << 15)) & 0xffff;
	b = (uint16_t) bb, a = (uint16_t) aa;
}
void wav2Decode(uint16_t *in, int nx, int ox, int ny, int oy, uint16_t mx) {
	const bool w14 = mx < (1 << 14);
	const int n = nx > ny ? ny : nx;
	int p = 1, p2;
	while (p <= n) p <<= 1;
	p >>= 1, p2 = p, p >>= 1;
	auto dec2 = [&](uint16_t l, uint16_t h, uint16_t &a, uint16_t &b) { w14 ? wdec14(l, h, a, b) : wdec16(l, h, a, b); };
	while (p >= 1) {
		uint16_t *py = in, *ey = in + (ptrdiff_t) oy * (ny - p2);
		const ptrdiff_t oy1 = (ptrdiff_t) oy * p, oy2 = (ptrdiff_t) oy * p2, ox1 = (ptrdiff_t) ox * p, ox2 = (ptrdiff_t) ox * p2;
		uint16_t i00, i01, i10, i11;
		for (; py <= ey; py += oy2) {
			uint16_t *px = py, *ex = py + (ptrdiff_t) ox * (nx - p2);
			for (; px <= ex; px += ox2) {
				uint16_t *p01 = px + ox1, *p10 = px + oy1, *p11 = p10 + ox1;
				dec2(*px, *p10, i00, i10);
				dec2(*p01, *p11, i01, i11);
				dec2(i00, i01, *px, *p01);
				dec2(i10, i11, *p10, *p11);
			}
			if (nx & p) {
				uint16_t *p10 = px + oy1;
				dec2(*px, *p10, i00, *p10);
				*px = i00;
			}
		}
		if (ny & p) {
			uint16_t *px = py, *ex = py + (ptrdiff_t) ox * (nx - p2);
			for (; px <= ex; px += ox2) {
				uint16_t *p01 = px + ox1;
				dec2(*px, *p01, i00, *p01);
				*px = i00;
			}
		}
		p2 = p, p >>= 1;
	}
}

// one chunk -> the uncompressed scanline layout (per line: per channel: width samples)
bool uncompressChunk(const unsigned char *src, size_t size, const std::vector<ExrChannel> &channels, int width, int lines, unsigned char *raw) {
	if (size < 4) return false;
	uint16_t minNonZero, maxNonZero;
	memcpy(&minNonZero, src, 2), memcpy(&maxNonZero, src + 2, 2);
	size_t pos = 4;
	std::vector<unsigned char> bitmap(8192, 0);
	if (minNonZero <= maxNonZero) {
		if (maxNonZero >= 8192) return false;
		const size_t nb = (size_t) maxNonZero - minNonZero + 1;
		if (pos + nb > size) return false;
		memcpy(&bitmap[minNonZero], src + pos, nb);
		pos += nb;
	}
	std::vector<uint16_t> lut(65536, 0);
	int k = 0;
	for (int i = 0; i < 65536; i++)
		if (i == 0 || (bitmap[i >> 3] & (1 << (i & 7)))) lut[k++] = (uint16_t) i;
	const uint16_t maxValue = (uint16_t) (k - 1);
	if (pos + 4 > size) return false;
	int32_t length;
	memcpy(&length, src + pos, 4);
	pos += 4;
	if (length < 0 || pos + (size_t) length > size) return false;
	size_t total = 0;
	for (const ExrChannel &c : channels) total += (size_t) width * lines * (c.type == 1 ? 1 : 2);
	std::vector<uint16_t> tmp(total);
	if (!hufUncompress(src + pos, (size_t) length, tmp.data(), total)) return false;
	std::vector<size_t> start(channels.size());
	size_t at = 0;
	for (size_t c = 0; c < channels.size(); c++) {
		const int sz = channels[c].type == 1 ? 1 : 2;
		start[c]	 = at;
		for (int j = 0; j < sz; j++) wav2Decode(tmp.data() + at + j, width, sz, lines, width * sz, maxValue);
		at += (size_t) width * lines * sz;
	}
	for (uint16_t &v : tmp) v = lut[v];
	unsigned char *out = raw;
	for (int l = 0; l < lines; l++)
		for (size_t c = 0; c < channels.size(); c++) {
			const size_t n = (size_t) width * (channels[c].type == 1 ? 1 : 2);
			memcpy(out, tmp.data() + start[c] + (size_t) l * n, n * 2);
			out += n * 2;
		}
	return true;
}
} // namespace piz

bool loadEXR(const string &path, Image &img, string *err) {
	std::vector<unsigned char> buf;
	if (!readFile(path, buf)) return fail(err, "cannot open " + path);
	if (buf.size() < 8 || buf[0] != 0x76 || buf[1] != 0x2f || buf[2] != 0x31 || buf[3] != 0x01) return fail(err, path + ": not an OpenEXR file");
	const uint32_t flags = buf[5] | buf[6] << 8 | buf[7] << 16;
	if (flags & 0x1a) return fail(err, path + ": tiled / deep / multi-part EXR files are not supported");
	size_t pos = 8;
	auto cstr  = [&](string &s) {
		 s.clear();
		 while (pos < buf.size() && buf[pos]) s.push_back((char) buf[pos++]);
		 pos++;
		 return pos <= buf.size();
	};
	auto i32 = [&](size_t p) { int32_t v; memcpy(&v, &buf[p], 4); return v; };
	std::vector<ExrChannel> channels;
	int compression = -1, lineOrder = 0, win[4] = {0, 0, -1, -1};
	while (true) {
		string name, type;
		if (!cstr(name)) return fail(err, path + ": truncated EXR header");
		if (name.empty()) break;
		if (!cstr(type) || pos + 4 > buf.size()) return fail(err, path + ": truncated EXR header");
		const int size = i32(pos);
		pos += 4;
		if (size < 0 || pos + (size_t) size > buf.size()) return fail(err, path + ": truncated EXR attribute " + name);
		const size_t end = pos + (size_t) size;
		if (name == "channels") {
			size_t p = pos;
			while (p < end && buf[p]) {
				ExrChannel ch;
				while (p < end && buf[p]) ch.name.push_back((char) buf[p++]);
				p++;
				if (p + 16 > end) return fail(err, path + ": bad channel list");
				ch.type = i32(p), ch.xs = i32(p + 8), ch.ys = i32(p + 12);
				p += 16;
				channels.push_back(ch);
			}
		} else if (name == "compression") compression = buf[pos];
		else if (name == "dataWindow") for (int k = 0; k < 4; k++) win[k] = i32(pos + 4 * k);
		else if (name == "lineOrder") lineOrder = buf[pos];
		pos = end;
	}
	(void) lineOrder; // every chunk carries its y coordinate
	if (channels.empty() || win[2] < win[0] || win[3] < win[1]) return fail(err, path + ": EXR header lacks channels / dataWindow");
	if (compression != 0 && compression != 2 && compression != 3 && compression != 4)
		return fail(err, path + ": EXR compression " + std::to_string(compression) + " is not supported (NONE, ZIPS, ZIP, PIZ are)");
	for (const ExrChannel &c : channels)
		if (c.xs != 1 || c.ys != 1 || c.type < 0 || c.type > 2) return fail(err, path + ": sub-sampled or unknown-type EXR channels are not supported");
	const int width = win[2] - win[0] + 1, height = win[3] - win[1] + 1;
	const int linesPerBlock = compression == 4 ? 32 : compression == 3 ? 16 : 1;
	const int nBlocks		= (height + linesPerBlock - 1) / linesPerBlock;
	size_t lineBytes = 0;
	for (const ExrChannel &c : channels) lineBytes += (size_t) width * (c.type == 1 ? 2 : 4);
	if (pos + (size_t) nBlocks * 8 > buf.size()) return fail(err, path + ": truncated EXR offset table");
	// channel -> RGBA slot (tinyexr::load_exr: by name; a single channel is replicated)
	std::vector<int> slot(channels.size(), -1);
	for (size_t c = 0; c < channels.size(); c++) {
		const string &n = channels[c].name;
		slot[c] = n == "R" ? 0 : n == "G" ? 1 : n == "B" ? 2 : n == "A" ? 3 : -1;
	}
	const bool grey = channels.size() == 1;
	img.width = width, img.height = height;
	img.rgba.assign((size_t) width * height * 4, 0.f);
	bool haveA = false;
	for (size_t c = 0; c < channels.size(); c++) haveA |= slot[c] == 3;
	if (!haveA && !grey)
		for (size_t i = 0; i < (size_t) width * height; i++) img.rgba[4 * i + 3] = 1.f;
	std::vector<unsigned char> raw, tmp;
	for (int b = 0; b < nBlocks; b++) {
		uint64_t off;
		memcpy(&off, &buf[pos + (size_t) b * 8], 8);
		if (off + 8 > buf.size()) return fail(err, path + ": bad EXR chunk offset");
		const int y0 = i32(off) - win[1], size = i32(off + 4);
		if (y0 < 0 || y0 >= height || size < 0 || off + 8 + (size_t) size > buf.size()) return fail(err, path + ": bad EXR chunk");
		const int lines	  = std::min(linesPerBlock, height - y0);
		const size_t want = lineBytes * lines;
		raw.resize(want);
		const unsigned char *src = &buf[off + 8];
		if (compression == 0 || (size_t) size == want) memcpy(raw.data(), src, std::min(want, (size_t) size));
		else if (compression == 4) {
			if (!piz::uncompressChunk(src, (size_t) size, channels, width, lines, raw.data())) return fail(err, path + ": corrupt PIZ data in EXR chunk");
		} else {
			tmp.resize(want);
			uLongf got = (uLongf) want;
			if (uncompress(tmp.data(), &got, src, (uLong) size) != Z_OK || got != want) return fail(err, path + ": zlib error in EXR chunk");
			zipPredictDecode(tmp, raw.data());
		}
		const unsigned char *p = raw.data();
		for (int l = 0; l < lines; l++)
			for (size_t c = 0; c < channels.size(); c++) {
				const int t = channels[c].type;
				for (int x = 0; x < width; x++) {
					float v;
					if (t == 1) { uint16_t h; memcpy(&h, p, 2); p += 2; v = halfToFloat(h); }
					else if (t == 2) { memcpy(&v, p, 4); p += 4; }
					else { uint32_t u; memcpy(&u, p, 4); p += 4; v = (float) u; }
					float *px = &img.rgba[((size_t) (y0 + l) * width + x) * 4];
					if (grey) px[0] = px[1] = px[2] = px[3] = v;
					else if (slot[c] >= 0) px[slot[c]] = v;
				}
			}
	}
	return true;
}

// halfPrecision: tinyexr::save_exr stores HALF (requested_pixel_types, image.cpp:86-91); FLOAT keeps the
// film's bits.  zip: ZIP (16-line blocks) instead of no compression.
bool saveEXR(const string &path, const Image &img, bool halfPrecision, bool zip, string *err) {
	FILE *f = fopen(path.c_str(), "wb");
	if (!f) return fail(err, "cannot write " + path);
	std::vector<unsigned char> hdr;
	auto put  = [&](const void *p, size_t n) { hdr.insert(hdr.end(), (const unsigned char *) p, (const unsigned char *) p + n); };
	auto puts = [&](const char *s) { put(s, strlen(s) + 1); };
	auto puti = [&](int32_t v) { put(&v, 4); };
	auto putf = [&](float v) { put(&v, 4); };
	const unsigned char magic[8] = {0x76, 0x2f, 0x31, 0x01, 2, 0, 0, 0};
	put(magic, 8);
	const char *names[4] = {"A", "B", "G", "R"}; // the channel list of an EXR file is sorted by name
	const int srcOf[4]	 = {3, 2, 1, 0};
	puts("channels"), puts("chlist"), puti(4 * 18 + 1);
	for (int c = 0; c < 4; c++) {
		puts(names[c]);
		puti(halfPrecision ? 1 : 2);
		const unsigned char lin[4] = {0, 0, 0, 0};
		put(lin, 4);
		puti(1), puti(1);
	}
	hdr.push_back(0);
	puts("compression"), puts("compression"), puti(1), hdr.push_back(zip ? 3 : 0);
	puts("dataWindow"), puts("box2i"), puti(16), puti(0), puti(0), puti(img.width - 1), puti(img.height - 1);
	puts("displayWindow"), puts("box2i"), puti(16), puti(0), puti(0), puti(img.width - 1), puti(img.height - 1);
	puts("lineOrder"), puts("lineOrder"), puti(1), hdr.push_back(0);
	puts("pixelAspectRatio"), puts("float"), puti(4), putf(1.f);
	puts("screenWindowCenter"), puts("v2f"), puti(8), putf(0.f), putf(0.f);
	puts("screenWindowWidth"), puts("float"), puti(4), putf(1.f);
	hdr.push_back(0);
	const int linesPerBlock = zip ? 16 : 1;
	const int nBlocks		= (img.height + linesPerBlock - 1) / linesPerBlock;
	const size_t px			= halfPrecision ? 2 : 4, lineBytes = (size_t) img.width * 4 * px;
	std::vector<uint64_t> offsets(nBlocks);
	std::vector<unsigned char> body, raw, tmp, comp;
	uint64_t cursor = hdr.size() + (uint64_t) nBlocks * 8;
	for (int b = 0; b < nBlocks; b++) {
		const int y0 = b * linesPerBlock, lines = std::min(linesPerBlock, img.height - y0);
		raw.resize(lineBytes * lines);
		unsigned char *p = raw.data();
		for (int l = 0; l < lines; l++)
			for (int c = 0; c < 4; c++)
				for (int x = 0; x < img.width; x++) {
					const float v = img.rgba[((size_t) (y0 + l) * img.width + x) * 4 + srcOf[c]];
					if (halfPrecision) { uint16_t h = floatToHalf(v); memcpy(p, &h, 2); p += 2; }
					else { memcpy(p, &v, 4); p += 4; }
				}
		const unsigned char *out = raw.data();
		size_t outSize			 = raw.size();
		if (zip) {
			zipPredictEncode(raw.data(), raw.size(), tmp);
			uLongf bound = compressBound((uLong) tmp.size());
			comp.resize(bound);
			if (compress(comp.data(), &bound, tmp.data(), (uLong) tmp.size()) == Z_OK && bound < raw.size()) out = comp.data(), outSize = bound;
		}
		offsets[b] = cursor;
		int32_t y = y0, sz = (int32_t) outSize;
		body.insert(body.end(), (unsigned char *) &y, (unsigned char *) &y + 4);
		body.insert(body.end(), (unsigned char *) &sz, (unsigned char *) &sz + 4);
		body.insert(body.end(), out, out + outSize);
		cursor += 8 + outSize;
	}
	bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size() && fwrite(offsets.data(), 8, offsets.size(), f) == offsets.size() &&
			  fwrite(body.data(), 1, body.size(), f) == body.size();
	fclose(f);
	return ok ? true : fail(err, "short write to " + path);
}

// =================================================================================================
// Image::loadImage(filepath, flip, srgb), texture.cpp:27-87 (HDR formats only)
bool loadImage(const string &path, Image &img, bool flip, string *err) {
	bool ok;
	if (endsWith(path, ".pfm")) ok = loadPFM(path, img, err);
	else if (endsWith(path, ".exr")) ok = loadEXR(path, img, err);
	else return fail(err, "unsupported image format: " + path + " (.exr and .pfm are)");
	if (ok && flip) img.flipVertically();
	return ok;
}

// Image::saveImage(filepath, flip), texture.cpp:89-118.  The reference's save_exr hands the planes to
// tinyexr in reversed order (image.cpp:45-47), so the file's channel X does not hold the image's X: file
// A = image R, file R = image G, file G = image B, file B = image A.  ErrorMeasurePass undoes exactly
// that permutation when it loads a reference image (errormeasure.cpp:96-105).  `referenceChannelOrder`
// reproduces it, so files written here and files written by KiRaRay are interchangeable.
bool saveImage(const string &path, const Image &img, bool flip, string *err, bool referenceChannelOrder) {
	Image tmp = img;
	if (flip) tmp.flipVertically();
	if (endsWith(path, ".pfm")) return savePFM(path, tmp, err);
	if (!endsWith(path, ".exr")) return fail(err, "unsupported image format: " + path + " (.exr and .pfm are)");
	if (referenceChannelOrder)
		for (size_t i = 0; i < tmp.rgba.size(); i += 4) {
			float r = tmp.rgba[i], g = tmp.rgba[i + 1], b = tmp.rgba[i + 2], a = tmp.rgba[i + 3];
			tmp.rgba[i] = g, tmp.rgba[i + 1] = b, tmp.rgba[i + 2] = a, tmp.rgba[i + 3] = r; // file R,G,B,A <- image G,B,A,R
		}
	return saveEXR(path, tmp, true, false, err);
}

void Image::flipVertically() {
	std::vector<float> row((size_t) width * 4);
	for (int y = 0; y < height / 2; y++) {
		float *a = &rgba[(size_t) y * width * 4], *b = &rgba[(size_t) (height - 1 - y) * width * 4];
		memcpy(row.data(), a, row.size() * 4), memcpy(a, b, row.size() * 4), memcpy(b, row.data(), row.size() * 4);
	}
}

} // namespace krr
