// Minimal JSON reader for the three Plonky2 fixture files (types/deserialize.go, types/common_data.go).
// Numbers keep their source text: Goldilocks elements go up to 2^64 - 1 and must never pass through double.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace gpw {
namespace json {

struct Value {
  enum Type { Null, Bool, Number, String, Array, Object } type = Null;
  std::string str;  // Number (raw text) / String
  bool b = false;
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj;

  const Value& operator[](const char* key) const {
    if (type != Object) throw std::runtime_error(std::string("json: not an object when looking up ") + key);
    for (const auto& kv : obj)
      if (kv.first == key) return kv.second;
    throw std::runtime_error(std::string("json: missing key ") + key);
  }
  uint64_t u64() const {
    if (type != Number) throw std::runtime_error("json: expected a number");
    uint64_t v = 0;
    for (char c : str) {
      if (c < '0' || c > '9') throw std::runtime_error("json: expected an unsigned integer, got " + str);
      uint64_t nv = v * 10 + (uint64_t)(c - '0');
      if (nv / 10 != v) throw std::runtime_error("json: integer overflows u64: " + str);
      v = nv;
    }
    return v;
  }
  bool boolean() const {
    if (type != Bool) throw std::runtime_error("json: expected a bool");
    return b;
  }
};

class Parser {
 public:
  explicit Parser(const std::string& s) : s_(s) {}
  Value parse_all() {
    Value v = parse_value();
    ws();
    if (i_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  size_t i_ = 0;
  [[noreturn]] void fail(const char* what) { throw std::runtime_error(std::string("json: ") + what + " at offset " + std::to_string(i_)); }
  void ws() {
    while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\n' || s_[i_] == '\t' || s_[i_] == '\r')) i_++;
  }
  Value parse_value() {
    ws();
    if (i_ >= s_.size()) fail("unexpected end");
    char c = s_[i_];
    Value v;
    if (c == '{') {
      v.type = Value::Object;
      i_++;
      ws();
      if (i_ < s_.size() && s_[i_] == '}') { i_++; return v; }
      for (;;) {
        ws();
        Value k = parse_string();
        ws();
        if (i_ >= s_.size() || s_[i_] != ':') fail("expected ':'");
        i_++;
        v.obj.emplace_back(k.str, parse_value());
        ws();
        if (i_ < s_.size() && s_[i_] == ',') { i_++; continue; }
        if (i_ < s_.size() && s_[i_] == '}') { i_++; break; }
        fail("expected ',' or '}'");
      }
      return v;
    }
    if (c == '[') {
      v.type = Value::Array;
      i_++;
      ws();
      if (i_ < s_.size() && s_[i_] == ']') { i_++; return v; }
      for (;;) {
        v.arr.push_back(parse_value());
        ws();
        if (i_ < s_.size() && s_[i_] == ',') { i_++; continue; }
        if (i_ < s_.size() && s_[i_] == ']') { i_++; break; }
        fail("expected ',' or ']'");
      }
      return v;
    }
    if (c == '"') return parse_string();
    if (s_.compare(i_, 4, "true") == 0) { i_ += 4; v.type = Value::Bool; v.b = true; return v; }
    if (s_.compare(i_, 5, "false") == 0) { i_ += 5; v.type = Value::Bool; v.b = false; return v; }
    if (s_.compare(i_, 4, "null") == 0) { i_ += 4; return v; }
    size_t j = i_;
    while (j < s_.size() && (s_[j] == '-' || s_[j] == '+' || s_[j] == '.' || s_[j] == 'e' || s_[j] == 'E' || (s_[j] >= '0' && s_[j] <= '9'))) j++;
    if (j == i_) fail("unexpected character");
    v.type = Value::Number;
    v.str = s_.substr(i_, j - i_);
    i_ = j;
    return v;
  }
  Value parse_string() {
    if (i_ >= s_.size() || s_[i_] != '"') fail("expected string");
    i_++;
    Value v;
    v.type = Value::String;
    while (i_ < s_.size() && s_[i_] != '"') {
      if (s_[i_] == '\\') {
        i_++;
        if (i_ >= s_.size()) fail("bad escape");
        char e = s_[i_];
        v.str += (e == 'n' ? '\n' : e == 't' ? '\t' : e);
      } else {
        v.str += s_[i_];
      }
      i_++;
    }
    if (i_ >= s_.size()) fail("unterminated string");
    i_++;
    return v;
  }
};

inline Value parse(const std::string& s) { return Parser(s).parse_all(); }

}  // namespace json
}  // namespace gpw
