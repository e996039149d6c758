// mini_json.h — a small recursive-descent JSON reader for the executable's config
// files (the reference vendors nlohmann/json for this, src/json.hpp; the 37 keys read by
// from_json(Param), src/main.cc:500-562, only need numbers, arrays and objects).
#ifndef HMC_MINI_JSON_H
#define HMC_MINI_JSON_H

#include <cctype>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace mini_json {

struct Value;
using ValuePtr = std::shared_ptr<Value>;

struct Value {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0.0;
  std::string str;
  std::vector<ValuePtr> arr;
  std::map<std::string, ValuePtr> obj;

  bool contains(const std::string &k) const { return kind == Object && obj.count(k) != 0; }
  const Value &at(const std::string &k) const {
    if (kind != Object) throw std::runtime_error("json: not an object");
    auto it = obj.find(k);
    if (it == obj.end()) throw std::runtime_error("json: key '" + k + "' not found");
    return *it->second;
  }
  double number() const {
    if (kind != Number) throw std::runtime_error("json: not a number");
    return num;
  }
  const std::vector<ValuePtr> &array() const {
    if (kind != Array) throw std::runtime_error("json: not an array");
    return arr;
  }
};

class Parser {
public:
  explicit Parser(const std::string &text) : s(text) {}
  ValuePtr parse() {
    ValuePtr v = value();
    ws();
    if (i != s.size()) fail("trailing characters");
    return v;
  }

private:
  const std::string &s;
  size_t i = 0;

  [[noreturn]] void fail(const std::string &m) const {
    throw std::runtime_error("json parse error at byte " + std::to_string(i) + ": " + m);
  }
  void ws() {
    while (i < s.size() && std::isspace(static_cast<unsigned char>(s[i]))) i++;
  }
  bool eat(char c) {
    ws();
    if (i < s.size() && s[i] == c) {
      i++;
      return true;
    }
    return false;
  }
  ValuePtr value() {
    ws();
    if (i >= s.size()) fail("unexpected end");
    auto v = std::make_shared<Value>();
    const char c = s[i];
    if (c == '{') {
      i++;
      v->kind = Value::Object;
      if (eat('}')) return v;
      do {
        ws();
        const std::string k = string();
        if (!eat(':')) fail("expected ':'");
        v->obj[k] = value();
      } while (eat(','));
      if (!eat('}')) fail("expected '}'");
    } else if (c == '[') {
      i++;
      v->kind = Value::Array;
      if (eat(']')) return v;
      do v->arr.push_back(value());
      while (eat(','));
      if (!eat(']')) fail("expected ']'");
    } else if (c == '"') {
      v->kind = Value::String;
      v->str = string();
    } else if (s.compare(i, 4, "true") == 0) {
      v->kind = Value::Bool;
      v->b = true;
      i += 4;
    } else if (s.compare(i, 5, "false") == 0) {
      v->kind = Value::Bool;
      i += 5;
    } else if (s.compare(i, 4, "null") == 0) {
      i += 4;
    } else {
      const char *b = s.c_str() + i;
      char *e = nullptr;
      v->num = std::strtod(b, &e);
      if (e == b) fail("bad token");
      v->kind = Value::Number;
      i += size_t(e - b);
    }
    return v;
  }
  std::string string() {
    if (i >= s.size() || s[i] != '"') fail("expected string");
    i++;
    std::string out;
    while (i < s.size() && s[i] != '"') {
      if (s[i] == '\\' && i + 1 < s.size()) {
        i++;
        switch (s[i]) {
        case 'n': out += '\n'; break;
        case 't': out += '\t'; break;
        case 'u': i += 4; out += '?'; break;
        default: out += s[i];
        }
      } else {
        out += s[i];
      }
      i++;
    }
    if (i >= s.size()) fail("unterminated string");
    i++;
    return out;
  }
};

inline ValuePtr parse(const std::string &text) { return Parser(text).parse(); }

} // namespace mini_json
#endif
