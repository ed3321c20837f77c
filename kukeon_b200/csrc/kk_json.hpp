// Minimal strict JSON reader (RFC 8259) for safetensors headers and model.safetensors.index.json.
// Keeps object members in file order; numbers that are integral and fit are kept as uint64/int64.
#pragma once
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "kk_common.hpp"

namespace kk {

struct JsonValue {
  enum Kind { Null, Bool, Int, Double, String, Array, Object } kind = Null;
  bool b = false;
  bool neg = false;      // Int: sign
  uint64_t u = 0;        // Int: magnitude
  double d = 0.0;
  std::string s;
  std::vector<JsonValue> arr;
  std::vector<std::pair<std::string, JsonValue>> obj;

  const JsonValue* find(const char* key) const {
    for (auto& kv : obj)
      if (kv.first == key) return &kv.second;
    return nullptr;
  }
  bool is_uint() const { return kind == Int && !neg; }
};

class JsonParser {
 public:
  JsonParser(const char* p, size_t n) : p_(p), end_(p + n) {}

  JsonValue parse_document() {
    JsonValue v = parse_value(0);
    skip_ws();
    if (p_ != end_) fail(KK_EFORMAT, "json: trailing characters at byte %zu", pos());
    return v;
  }

 private:
  const char* p_;
  const char* end_;
  const char* begin_ = p_;
  size_t pos() const { return (size_t)(p_ - begin_); }

  void skip_ws() {
    while (p_ < end_ && (*p_ == ' ' || *p_ == '\t' || *p_ == '\n' || *p_ == '\r')) ++p_;
  }
  char peek() {
    if (p_ >= end_) fail(KK_EFORMAT, "json: unexpected end of input");
    return *p_;
  }
  void expect(char c) {
    if (peek() != c) fail(KK_EFORMAT, "json: expected '%c' at byte %zu", c, pos());
    ++p_;
  }

  JsonValue parse_value(int depth) {
    if (depth > 64) fail(KK_EFORMAT, "json: nesting too deep");
    skip_ws();
    char c = peek();
    JsonValue v;
    if (c == '{') {
      v.kind = JsonValue::Object;
      ++p_;
      skip_ws();
      if (peek() == '}') { ++p_; return v; }
      for (;;) {
        skip_ws();
        std::string key = parse_string();
        skip_ws();
        expect(':');
        JsonValue child = parse_value(depth + 1);
        v.obj.emplace_back(std::move(key), std::move(child));
        skip_ws();
        if (peek() == ',') { ++p_; continue; }
        expect('}');
        return v;
      }
    } else if (c == '[') {
      v.kind = JsonValue::Array;
      ++p_;
      skip_ws();
      if (peek() == ']') { ++p_; return v; }
      for (;;) {
        v.arr.push_back(parse_value(depth + 1));
        skip_ws();
        if (peek() == ',') { ++p_; continue; }
        expect(']');
        return v;
      }
    } else if (c == '"') {
      v.kind = JsonValue::String;
      v.s = parse_string();
      return v;
    } else if (c == 't' || c == 'f' || c == 'n') {
      auto lit = [&](const char* w) {
        size_t n = strlen(w);
        if ((size_t)(end_ - p_) < n || memcmp(p_, w, n) != 0)
          fail(KK_EFORMAT, "json: bad literal at byte %zu", pos());
        p_ += n;
      };
      if (c == 't') { lit("true"); v.kind = JsonValue::Bool; v.b = true; }
      else if (c == 'f') { lit("false"); v.kind = JsonValue::Bool; v.b = false; }
      else { lit("null"); v.kind = JsonValue::Null; }
      return v;
    } else if (c == '-' || (c >= '0' && c <= '9')) {
      return parse_number();
    }
    fail(KK_EFORMAT, "json: unexpected character 0x%02x at byte %zu", (unsigned char)c, pos());
  }

  JsonValue parse_number() {
    JsonValue v;
    const char* s = p_;
    bool neg = false;
    if (*p_ == '-') { neg = true; ++p_; }
    if (p_ >= end_ || *p_ < '0' || *p_ > '9') fail(KK_EFORMAT, "json: bad number at byte %zu", pos());
    bool integral = true, overflow = false;
    uint64_t mag = 0;
    if (*p_ == '0') { ++p_; }
    else {
      while (p_ < end_ && *p_ >= '0' && *p_ <= '9') {
        uint64_t dgt = (uint64_t)(*p_ - '0');
        if (mag > (UINT64_MAX - dgt) / 10) overflow = true;
        else mag = mag * 10 + dgt;
        ++p_;
      }
    }
    if (p_ < end_ && *p_ == '.') {
      integral = false; ++p_;
      if (p_ >= end_ || *p_ < '0' || *p_ > '9') fail(KK_EFORMAT, "json: bad fraction at byte %zu", pos());
      while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_;
    }
    if (p_ < end_ && (*p_ == 'e' || *p_ == 'E')) {
      integral = false; ++p_;
      if (p_ < end_ && (*p_ == '+' || *p_ == '-')) ++p_;
      if (p_ >= end_ || *p_ < '0' || *p_ > '9') fail(KK_EFORMAT, "json: bad exponent at byte %zu", pos());
      while (p_ < end_ && *p_ >= '0' && *p_ <= '9') ++p_;
    }
    if (integral && !overflow) {
      v.kind = JsonValue::Int; v.neg = neg && mag != 0; v.u = mag;
      v.d = neg ? -(double)mag : (double)mag;
    } else {
      v.kind = JsonValue::Double;
      v.d = strtod(std::string(s, p_).c_str(), nullptr);
    }
    return v;
  }

  static void append_utf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) out.push_back((char)cp);
    else if (cp < 0x800) { out.push_back((char)(0xC0 | (cp >> 6))); out.push_back((char)(0x80 | (cp & 0x3F))); }
    else if (cp < 0x10000) {
      out.push_back((char)(0xE0 | (cp >> 12))); out.push_back((char)(0x80 | ((cp >> 6) & 0x3F)));
      out.push_back((char)(0x80 | (cp & 0x3F)));
    } else {
      out.push_back((char)(0xF0 | (cp >> 18))); out.push_back((char)(0x80 | ((cp >> 12) & 0x3F)));
      out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))); out.push_back((char)(0x80 | (cp & 0x3F)));
    }
  }

  uint32_t parse_hex4() {
    if (end_ - p_ < 4) fail(KK_EFORMAT, "json: truncated \\u escape");
    uint32_t v = 0;
    for (int i = 0; i < 4; ++i) {
      char c = *p_++;
      v <<= 4;
      if (c >= '0' && c <= '9') v |= (uint32_t)(c - '0');
      else if (c >= 'a' && c <= 'f') v |= (uint32_t)(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') v |= (uint32_t)(c - 'A' + 10);
      else fail(KK_EFORMAT, "json: bad \\u escape at byte %zu", pos());
    }
    return v;
  }

  std::string parse_string() {
    expect('"');
    std::string out;
    for (;;) {
      if (p_ >= end_) fail(KK_EFORMAT, "json: unterminated string");
      unsigned char c = (unsigned char)*p_++;
      if (c == '"') return out;
      if (c < 0x20) fail(KK_EFORMAT, "json: control character in string at byte %zu", pos());
      if (c != '\\') { out.push_back((char)c); continue; }
      if (p_ >= end_) fail(KK_EFORMAT, "json: unterminated escape");
      char e = *p_++;
      switch (e) {
        case '"': out.push_back('"'); break;
        case '\\': out.push_back('\\'); break;
        case '/': out.push_back('/'); break;
        case 'b': out.push_back('\b'); break;
        case 'f': out.push_back('\f'); break;
        case 'n': out.push_back('\n'); break;
        case 'r': out.push_back('\r'); break;
        case 't': out.push_back('\t'); break;
        case 'u': {
          uint32_t cp = parse_hex4();
          if (cp >= 0xD800 && cp <= 0xDBFF) {
            if (end_ - p_ >= 6 && p_[0] == '\\' && p_[1] == 'u') {
              p_ += 2;
              uint32_t lo = parse_hex4();
              if (lo < 0xDC00 || lo > 0xDFFF) fail(KK_EFORMAT, "json: bad surrogate pair");
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            } else fail(KK_EFORMAT, "json: lone surrogate");
          } else if (cp >= 0xDC00 && cp <= 0xDFFF) fail(KK_EFORMAT, "json: lone surrogate");
          append_utf8(out, cp);
          break;
        }
        default: fail(KK_EFORMAT, "json: bad escape '\\%c'", e);
      }
    }
  }
};

// Escape a UTF-8 string for embedding in JSON output (manifest, stats).
inline std::string json_escape(const std::string& s) {
  std::string o;
  o.reserve(s.size() + 2);
  for (unsigned char c : s) {
    switch (c) {
      case '"': o += "\\\""; break;
      case '\\': o += "\\\\"; break;
      case '\n': o += "\\n"; break;
      case '\r': o += "\\r"; break;
      case '\t': o += "\\t"; break;
      default:
        if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04x", c); o += b; }
        else o.push_back((char)c);
    }
  }
  return o;
}

}  // namespace kk
