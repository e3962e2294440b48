// json.h -- minimal JSON value / parser / writer for the host layer.
// The reference uses nlohmann::json (src/ext/json); only the subset its configs need is provided:
// objects, arrays, strings, numbers, booleans, null; `value(key, default)`, `contains`, `at`.
#pragma once
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace krr {

class Json {
public:
	enum class Type { Null, Bool, Number, String, Array, Object };
	using array_t  = std::vector<Json>;
	using object_t = std::vector<std::pair<std::string, Json>>; // keeps insertion order

	Json() = default;
	Json(bool b) : mType(Type::Bool), mBool(b) {}
	Json(double d) : mType(Type::Number), mNum(d) {}
	Json(int i) : mType(Type::Number), mNum(i), mIsInt(true) {}
	Json(long long i) : mType(Type::Number), mNum((double) i), mIsInt(true) {}
	Json(const char *s) : mType(Type::String), mStr(s) {}
	Json(const std::string &s) : mType(Type::String), mStr(s) {}
	static Json array() { Json j; j.mType = Type::Array; return j; }
	static Json object() { Json j; j.mType = Type::Object; return j; }

	Type type() const { return mType; }
	bool isNull() const { return mType == Type::Null; }
	bool isObject() const { return mType == Type::Object; }
	bool isArray() const { return mType == Type::Array; }
	bool isString() const { return mType == Type::String; }
	bool isNumber() const { return mType == Type::Number; }
	bool isBool() const { return mType == Type::Bool; }
	bool isFloat() const { return mType == Type::Number && !mIsInt; }

	bool contains(const std::string &k) const { return find(k) != nullptr; }
	const Json &at(const std::string &k) const {
		const Json *j = find(k);
		if (!j) throw std::runtime_error("json: missing key '" + k + "'");
		return *j;
	}
	const Json &operator[](const std::string &k) const { return at(k); }
	Json &operator[](const std::string &k) {
		if (mType == Type::Null) mType = Type::Object;
		for (auto &kv : mObj) if (kv.first == k) return kv.second;
		mObj.emplace_back(k, Json());
		return mObj.back().second;
	}
	const Json &at(size_t i) const {
		if (mType != Type::Array || i >= mArr.size()) throw std::runtime_error("json: bad array index");
		return mArr[i];
	}
	const Json &operator[](size_t i) const { return at(i); }
	size_t size() const { return mType == Type::Array ? mArr.size() : mType == Type::Object ? mObj.size() : 0; }
	const array_t &items() const { return mArr; }
	const object_t &members() const { return mObj; }
	void push_back(const Json &j) { if (mType == Type::Null) mType = Type::Array; mArr.push_back(j); }

	double asNumber() const { if (mType == Type::Bool) return mBool; if (mType != Type::Number) throw std::runtime_error("json: not a number"); return mNum; }
	bool asBool() const { if (mType == Type::Number) return mNum != 0; if (mType != Type::Bool) throw std::runtime_error("json: not a bool"); return mBool; }
	const std::string &asString() const { if (mType != Type::String) throw std::runtime_error("json: not a string"); return mStr; }

	// value(key, default): nlohmann semantics -- default when the key is absent
	double value(const std::string &k, double d) const { const Json *j = find(k); return j && !j->isNull() ? j->asNumber() : d; }
	float value(const std::string &k, float d) const { return (float) value(k, (double) d); }
	int value(const std::string &k, int d) const { const Json *j = find(k); return j && !j->isNull() ? (int) j->asNumber() : d; }
	bool value(const std::string &k, bool d) const { const Json *j = find(k); return j && !j->isNull() ? j->asBool() : d; }
	std::string value(const std::string &k, const char *d) const { const Json *j = find(k); return j && j->isString() ? j->asString() : std::string(d); }
	std::string value(const std::string &k, const std::string &d) const { return value(k, d.c_str()); }
	Json value(const std::string &k, const Json &d) const { const Json *j = find(k); return j ? *j : d; }
	// fixed-size float arrays ("translate": [x,y,z], ...)
	template <int N> bool getFloats(const std::string &k, float (&out)[N]) const {
		const Json *j = find(k);
		if (!j || !j->isArray() || j->size() != (size_t) N) return false;
		for (int i = 0; i < N; i++) out[i] = (float) j->mArr[i].asNumber();
		return true;
	}

	static Json parse(const std::string &text) {
		Parser p{text.c_str(), text.c_str() + text.size()};
		Json j = p.parseValue();
		p.skipWs();
		if (p.cur != p.end) throw std::runtime_error("json: trailing characters");
		return j;
	}

	std::string dump(int indent = -1) const {
		std::ostringstream os;
		write(os, indent, 0);
		return os.str();
	}

private:
	const Json *find(const std::string &k) const {
		if (mType != Type::Object) return nullptr;
		for (auto &kv : mObj) if (kv.first == k) return &kv.second;
		return nullptr;
	}

	struct Parser {
		const char *cur, *end;
		void skipWs() {
			while (cur < end) {
				if (*cur == ' ' || *cur == '\t' || *cur == '\n' || *cur == '\r') cur++;
				else if (*cur == '/' && cur + 1 < end && cur[1] == '/') { while (cur < end && *cur != '\n') cur++; }
				else break;
			}
		}
		[[noreturn]] void fail(const char *what) { throw std::runtime_error(std::string("json: ") + what); }
		Json parseValue() {
			skipWs();
			if (cur >= end) fail("unexpected end");
			char c = *cur;
			if (c == '{') return parseObject();
			if (c == '[') return parseArray();
			if (c == '"') return Json(parseString());
			if (c == 't' && end - cur >= 4 && !strncmp(cur, "true", 4)) { cur += 4; return Json(true); }
			if (c == 'f' && end - cur >= 5 && !strncmp(cur, "false", 5)) { cur += 5; return Json(false); }
			if (c == 'n' && end - cur >= 4 && !strncmp(cur, "null", 4)) { cur += 4; return Json(); }
			return parseNumber();
		}
		static int strncmp(const char *a, const char *b, size_t n) { for (size_t i = 0; i < n; i++) if (a[i] != b[i]) return 1; return 0; }
		Json parseNumber() {
			const char *s = cur;
			bool isInt = true;
			if (cur < end && (*cur == '-' || *cur == '+')) cur++;
			while (cur < end && ((*cur >= '0' && *cur <= '9') || *cur == '.' || *cur == 'e' || *cur == 'E' || *cur == '-' || *cur == '+')) {
				if (*cur == '.' || *cur == 'e' || *cur == 'E') isInt = false;
				cur++;
			}
			if (s == cur) fail("bad value");
			std::string t(s, cur);
			Json j(std::strtod(t.c_str(), nullptr));
			j.mIsInt = isInt;
			return j;
		}
		std::string parseString() {
			std::string out;
			cur++; // opening quote
			while (cur < end && *cur != '"') {
				if (*cur == '\\') {
					cur++;
					if (cur >= end) fail("bad escape");
					switch (*cur) {
						case 'n': out += '\n'; break; case 't': out += '\t'; break; case 'r': out += '\r'; break;
						case 'b': out += '\b'; break; case 'f': out += '\f'; break;
						case 'u': { // basic-plane only
							if (end - cur < 5) fail("bad \\u escape");
							unsigned v = (unsigned) std::strtoul(std::string(cur + 1, cur + 5).c_str(), nullptr, 16);
							if (v < 0x80) out += (char) v;
							else if (v < 0x800) { out += (char) (0xC0 | (v >> 6)); out += (char) (0x80 | (v & 0x3F)); }
							else { out += (char) (0xE0 | (v >> 12)); out += (char) (0x80 | ((v >> 6) & 0x3F)); out += (char) (0x80 | (v & 0x3F)); }
							cur += 4;
							break;
						}
						default: out += *cur;
					}
					cur++;
				} else out += *cur++;
			}
			if (cur >= end) fail("unterminated string");
			cur++;
			return out;
		}
		Json parseArray() {
			Json j = Json::array();
			cur++;
			skipWs();
			if (cur < end && *cur == ']') { cur++; return j; }
			while (true) {
				j.mArr.push_back(parseValue());
				skipWs();
				if (cur >= end) fail("unterminated array");
				if (*cur == ',') { cur++; continue; }
				if (*cur == ']') { cur++; break; }
				fail("expected , or ]");
			}
			return j;
		}
		Json parseObject() {
			Json j = Json::object();
			cur++;
			skipWs();
			if (cur < end && *cur == '}') { cur++; return j; }
			while (true) {
				skipWs();
				if (cur >= end || *cur != '"') fail("expected string key");
				std::string k = parseString();
				skipWs();
				if (cur >= end || *cur != ':') fail("expected :");
				cur++;
				j.mObj.emplace_back(k, parseValue());
				skipWs();
				if (cur >= end) fail("unterminated object");
				if (*cur == ',') { cur++; continue; }
				if (*cur == '}') { cur++; break; }
				fail("expected , or }");
			}
			return j;
		}
	};

	void write(std::ostream &os, int indent, int level) const {
		auto nl = [&](int l) { if (indent >= 0) { os << '\n'; for (int i = 0; i < indent * l; i++) os << ' '; } };
		switch (mType) {
			case Type::Null: os << "null"; break;
			case Type::Bool: os << (mBool ? "true" : "false"); break;
			case Type::Number: {
				char buf[40];
				if (mIsInt && std::fabs(mNum) < 9e15) snprintf(buf, sizeof buf, "%lld", (long long) mNum);
				else snprintf(buf, sizeof buf, "%.9g", mNum);
				os << buf;
				break;
			}
			case Type::String: {
				os << '"';
				for (char c : mStr) {
					if (c == '"' || c == '\\') os << '\\' << c;
					else if (c == '\n') os << "\\n";
					else if (c == '\t') os << "\\t";
					else os << c;
				}
				os << '"';
				break;
			}
			case Type::Array:
				os << '[';
				for (size_t i = 0; i < mArr.size(); i++) { if (i) os << ','; nl(level + 1); mArr[i].write(os, indent, level + 1); }
				if (!mArr.empty()) nl(level);
				os << ']';
				break;
			case Type::Object:
				os << '{';
				for (size_t i = 0; i < mObj.size(); i++) {
					if (i) os << ',';
					nl(level + 1);
					os << '"' << mObj[i].first << "\":" << (indent >= 0 ? " " : "");
					mObj[i].second.write(os, indent, level + 1);
				}
				if (!mObj.empty()) nl(level);
				os << '}';
				break;
		}
	}

	Type mType = Type::Null;
	bool mBool = false, mIsInt = false;
	double mNum = 0;
	std::string mStr;
	array_t mArr;
	object_t mObj;
};

} // namespace krr
